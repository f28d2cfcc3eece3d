// vilf/EstimationMapping.hpp — host-side mirror of the reference's `EstimationMapping` class
// (src/visual_inertial_lidar/feature_tracker/include/EstimationMapping.hpp:71-403): same method names, argument
// meaning, order of effects and public members; the work runs in libvilf_cuda.so (sm_100a) through include/vilf.h.
//
//   reference member                                          here -> C ABI
//   initParameter(ros::NodeHandle&)               EM:80-92    initParameter(NH&)  (template: ros::NodeHandle or vilf::ParamMap)
//   allocateMemory()                              EM:94-103   allocateMemory()    host-side clouds only; device state is the Session
//   localMapInited(edge, surf)                    EM:105-115  vilf_map_init (features resident) | vilf_map_init_points
//   optimation_processing(edge, surf)             EM:235-296  vilf_update   (features resident) | vilf_update_points
//   EdgeCostFactor(cloud, problem, loss)          EM:117-172  EdgeCostFactor(cloud)  vilf_factors; the "problem" is this object's factor list
//   SurfCostFactor(cloud, problem, loss)          EM:174-232  SurfCostFactor(cloud)  vilf_factors
//   ceres::Solve(options, &problem, &summary)     EM:283      SolveProblem()         vilf_solve (device LM, Ceres' schedule)
//   (prediction, first lines of the above)        EM:238-243  predictPose()          vilf_predict
//   createSubMap(edge, surf)                      EM:298-352  vilf_set_pose + vilf_create_submap
//   pointAssociaToMap(pi, po)                     EM:355-363  host arithmetic on parameter_opti (fp64 transform, fp32 store)
//   getMapCloud(reg, noreg) / getMapCloud(noreg)  EM:365-375  vilf_get_cloud(4 / 5)
//   parameter_opti, globalOdom, globalOdom_last   EM:383-388  refreshed from the device after every call that moves them
//   localMapEdge, localMapSurf                    EM:394-395  device resident; syncMapsToHost() fills the host clouds on demand
//
// ceres::Problem / ceres::LossFunction do not exist any more (the solver is a device kernel); the two factor methods
// keep their name, their order of effects (transform with the current pose, 5-NN in the current local map, fit,
// gate, append a factor) and their diagnostics ("not enough edge feature." / "no enough surf feature.", both below 20
// factors, EM:168-171, :227-230).  optimation_processing() does all of it on the device in one submission.
//
// F-LOAM names (BASELINE.json north_star) are provided as aliases on OdomEstimationClass at the end of this file.
#pragma once

#include <cmath>
#include <cstring>
#include <iostream>
#include <vector>

#include "cloud.hpp"
#include "featureExtraction.hpp"
#include "session.hpp"

namespace vilf {

class EstimationMapping {
 public:
  EstimationMapping() : edgeMapLeafSize(0.2), surfMapLeafSize(0.4), sess_(std::make_shared<Session>()), trust_resident_(true) {
    const double id[7] = {0, 0, 0, 1, 0, 0, 0};
    std::memcpy(parameter_opti, id, sizeof(id));
    allocateMemory();
  }

  // EM:80-92.  Same parameter names and the same code defaults as the reference.
  template <class NH>
  void initParameter(NH& nh) {
    nh.template param<double>("/EdgeLeafSize", edgeMapLeafSize, 0.2);
    nh.template param<double>("/SurfLeafSize", surfMapLeafSize, 0.4);
    globalOdom = Isometry3d::Identity();
    globalOdom_last = Isometry3d::Identity();
    allocateMemory();
    applyConfig();
  }

  void allocateMemory() {  // EM:94-103
    cloudRegistered = make_cloud();
    cloudNoRegistered = make_cloud();
    localMapEdge = make_cloud();
    localMapSurf = make_cloud();
  }

  // Put featureExtraction and EstimationMapping on ONE device session: features stay resident between the stages.
  void shareSession(featureExtraction& fe) {
    fe.shareSession(sess_);
    applyConfig();
  }
  // the same for any stage-1 class with shareSession(const SessionPtr&) (vilf::featureExtract, the ring-field extractor)
  template <class Stage1>
  void shareSessionWith(Stage1& fe) {
    fe.shareSession(sess_);
    applyConfig();
  }
  const SessionPtr& session() const { return sess_; }
  void setTrustResident(bool on) { trust_resident_ = on; }  // false: always upload the clouds that are passed in

  // EM:105-115
  void localMapInited(const CloudPtr& edge_cloud, const CloudPtr& surf_cloud) {
    applyConfig();
    vilf_handle* h = sess_->handle();
    if (resident(edge_cloud, surf_cloud)) {
      sess_->check(vilf_map_init(h), "localMapInited");
    } else {
      pack_cloud(*edge_cloud, edge_);
      pack_cloud(*surf_cloud, surf_);
      sess_->check(vilf_map_init_points(h, edge_.data(), (int)edge_cloud->points.size(), surf_.data(), (int)surf_cloud->points.size()), "localMapInited");
    }
    sess_->invalidate_resident();
    refreshPose();
  }

  // EM:235-296: predict, voxel-filter the scan features, 2 x (associate + fit + LM solve), write the pose back, createSubMap.
  void optimation_processing(const CloudPtr& edgeCloud_In, const CloudPtr& surfCloud_In) {
    applyConfig();
    vilf_handle* h = sess_->handle();
    double pose[7];
    if (resident(edgeCloud_In, surfCloud_In)) {
      sess_->check(vilf_update(h, pose), "optimation_processing");
    } else {
      pack_cloud(*edgeCloud_In, edge_);
      pack_cloud(*surfCloud_In, surf_);
      sess_->check(vilf_update_points(h, edge_.data(), (int)edgeCloud_In->points.size(), surf_.data(), (int)surfCloud_In->points.size(), pose), "optimation_processing");
    }
    sess_->invalidate_resident();
    refreshPose();
    int32_t c[8];
    sess_->check(vilf_get_counts(h, c), "optimation_processing");
    double rows[8 * 8];
    int nrows = 0;
    sess_->check(vilf_get_solves(h, rows, 8, &nrows), "optimation_processing");
    if (nrows == 0) {  // EM:286-289
      std::cout << "localMapEdge->points.size() = " << prev_map_[0] << " , localMapSurf->points.size() = " << prev_map_[1] << std::endl;
      std::cout << "not enough feature points in local map to associate." << std::endl;
    }
    for (int r = 0; r < nrows; ++r) {
      if (rows[8 * r + 0] < 20) std::cout << "not enough edge feature." << std::endl;  // EM:168-171
      if (rows[8 * r + 1] < 20) std::cout << "no enough surf feature." << std::endl;  // EM:227-230
    }
    prev_map_[0] = c[4];
    prev_map_[1] = c[5];
  }

  // ---- the pieces optimation_processing is made of, callable one by one like the reference's public methods ----
  // Voxel filters of EM:246-251 (voxelEdgeFilter / voxelSurfFilter) on an explicit cloud.
  void voxelFilter(const CloudPtr& in, double leaf, CloudPtr& out) {
    applyConfig();
    vilf_handle* h = sess_->handle();
    pack_cloud(*in, edge_);
    const int n = (int)in->points.size();
    surf_.resize((std::size_t)(n > 0 ? n : 1) * 4);
    int n_out = 0, guard = 0;
    sess_->check(vilf_voxel_downsample(h, edge_.data(), n, (float)leaf, surf_.data(), n > 0 ? n : 1, &n_out, &guard), "voxelFilter");
    out->clear();
    append_cloud(*out, surf_.data(), (std::size_t)n_out);
  }
  // EM:238-243: constant-velocity prediction (globalOdom, globalOdom_last, parameter_opti).
  // Step-by-step path only: optimation_processing() predicts by itself (EM:238-243), so predictPose() followed by
  // optimation_processing() would apply the constant-velocity step twice.
  void predictPose() {
    double pose[7];
    sess_->check(vilf_predict(sess_->handle(), pose), "predictPose");
    refreshPose();
  }
  void ProblemReset() {  // EM:263-268: new ceres::Problem
    pab_.clear();
    pnd_.clear();
  }
  // EM:117-172 without the ceres::Problem argument; returns the number of residual blocks added.
  int EdgeCostFactor(const CloudPtr& edge_cloud) {
    const int n = (int)edge_cloud->points.size();
    int added = 0;
    if (n > 0) {
      vilf_handle* h = sess_->handle();
      pack_cloud(*edge_cloud, edge_);
      std::vector<uint8_t> valid((std::size_t)n);
      std::vector<double> ab((std::size_t)n * 6);
      sess_->check(vilf_factors(h, parameter_opti, edge_.data(), n, nullptr, 0, valid.data(), ab.data(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr), "EdgeCostFactor");
      for (int i = 0; i < n; ++i) {
        if (!valid[i]) continue;
        for (int k = 0; k < 3; ++k) pab_.push_back((double)edge_[4 * i + k]);
        for (int k = 0; k < 6; ++k) pab_.push_back(ab[(std::size_t)i * 6 + k]);
        ++added;
      }
    }
    if (added < 20) std::cout << "not enough edge feature." << std::endl;
    return added;
  }
  // EM:174-232
  int SurfCostFactor(const CloudPtr& surf_cloud) {
    const int n = (int)surf_cloud->points.size();
    int added = 0;
    if (n > 0) {
      vilf_handle* h = sess_->handle();
      pack_cloud(*surf_cloud, surf_);
      std::vector<uint8_t> valid((std::size_t)n);
      std::vector<double> nd((std::size_t)n * 4);
      sess_->check(vilf_factors(h, parameter_opti, nullptr, 0, surf_.data(), n, nullptr, nullptr, nullptr, nullptr, valid.data(), nd.data(), nullptr, nullptr), "SurfCostFactor");
      for (int i = 0; i < n; ++i) {
        if (!valid[i]) continue;
        for (int k = 0; k < 3; ++k) pnd_.push_back((double)surf_[4 * i + k]);
        for (int k = 0; k < 4; ++k) pnd_.push_back(nd[(std::size_t)i * 4 + k]);
        ++added;
      }
    }
    if (added < 20) std::cout << "no enough surf feature." << std::endl;  // EM:227-230
    return added;
  }
  // EM:275-283: ceres::Solve(DENSE_QR, max_num_iterations = 4) on the factors added since ProblemReset(); updates
  // parameter_opti in place like Ceres does.  Returns the termination code (vilf.h: vilf_solve).
  int SolveProblem(int max_num_iterations = 4) {
    vilf_handle* h = sess_->handle();
    int nrows = 0, term = 0;
    sess_->check(vilf_solve(h, parameter_opti, pab_.data(), (int)(pab_.size() / 9), pnd_.data(), (int)(pnd_.size() / 7), max_num_iterations, nullptr, 0, &nrows, &term), "SolveProblem");
    return term;
  }
  // EM:291-293 + EM:298-352 at the current parameter_opti.
  void createSubMap(const CloudPtr& edge_cloud, const CloudPtr& surf_cloud) {
    vilf_handle* h = sess_->handle();
    sess_->check(vilf_set_pose(h, parameter_opti, 1), "createSubMap");
    pack_cloud(*edge_cloud, edge_);
    pack_cloud(*surf_cloud, surf_);
    sess_->check(vilf_create_submap(h, edge_.data(), (int)edge_cloud->points.size(), surf_.data(), (int)surf_cloud->points.size()), "createSubMap");
    sess_->invalidate_resident();
    refreshPose();
  }

  // EM:355-363: fp64 rotate + translate (Eigen's q * v: v + w*(2 q x v) + q x (2 q x v)), fp32 store.
  void pointAssociaToMap(PointType const* const p_in, PointType* const p_out) const {
    const double* q = parameter_opti;
    const double v[3] = {p_in->x, p_in->y, p_in->z};
    double uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
    for (int i = 0; i < 3; ++i) uv[i] += uv[i];
    const double c[3] = {q[1] * uv[2] - q[2] * uv[1], q[2] * uv[0] - q[0] * uv[2], q[0] * uv[1] - q[1] * uv[0]};
    p_out->x = (float)(((v[0] + q[3] * uv[0]) + c[0]) + parameter_opti[4]);
    p_out->y = (float)(((v[1] + q[3] * uv[1]) + c[1]) + parameter_opti[5]);
    p_out->z = (float)(((v[2] + q[3] * uv[2]) + c[2]) + parameter_opti[6]);
    p_out->intensity = p_in->intensity;
  }

  // EM:365-375
  void getMapCloud(CloudPtr& MapRsgistered, CloudPtr& MapNoRegistered) {
    fetch(4, *cloudRegistered);
    fetch(5, *cloudNoRegistered);
    *MapRsgistered = *cloudRegistered;
    *MapNoRegistered = *cloudNoRegistered;
  }
  void getMapCloud(CloudPtr& MapRsgistered) {  // the reference returns the UN-registered cloud here (EM:371-375)
    fetch(5, *cloudNoRegistered);
    *MapRsgistered = *cloudNoRegistered;
  }
  // localMapEdge / localMapSurf live on the device; copy them into the public host clouds on demand.
  void syncMapsToHost() {
    fetch(0, *localMapEdge);
    fetch(1, *localMapSurf);
  }

 public:  // EM:378-401
  double edgeMapLeafSize;
  double surfMapLeafSize;
  double parameter_opti[7];  // q (x, y, z, w), t
  Isometry3d globalOdom;
  Isometry3d globalOdom_last;
  CloudPtr cloudRegistered;
  CloudPtr cloudNoRegistered;
  CloudPtr localMapEdge;
  CloudPtr localMapSurf;

 private:
  void applyConfig() {
    if (sess_->created()) return;
    vilf_config& c = sess_->config();
    c.edge_leaf = edgeMapLeafSize;
    c.surf_leaf = surfMapLeafSize;
  }
  bool resident(const CloudPtr& e, const CloudPtr& s) const {
    if (!trust_resident_ || sess_->resident_tag() == 0) return false;
    const Session::Resident& r = sess_->resident;
    if (r.tag != sess_->resident_tag() || e->points.size() != r.n_edge || s->points.size() != r.n_surf) return false;
    return Session::content_hash(*e) == r.hash_edge && Session::content_hash(*s) == r.hash_surf;
  }
  void refreshPose() {
    double rt[12], st[31];
    sess_->check(vilf_get_pose(sess_->handle(), parameter_opti, rt), "vilf_get_pose");
    iso_from_rt12(globalOdom, rt);
    sess_->check(vilf_state_export(sess_->handle(), st), "vilf_state_export");
    iso_from_rt12(globalOdom_last, st + 19);
  }
  void fetch(int which, Cloud& out) {
    vilf_handle* h = sess_->handle();
    int n = 0;
    sess_->check(vilf_get_cloud(h, which, nullptr, 0, &n), "vilf_get_cloud");
    edge_.resize((std::size_t)(n > 0 ? n : 1) * 4);
    sess_->check(vilf_get_cloud(h, which, edge_.data(), n > 0 ? n : 1, &n), "vilf_get_cloud");
    out.clear();
    append_cloud(out, edge_.data(), (std::size_t)n);
  }

  SessionPtr sess_;
  bool trust_resident_;
  std::vector<float> edge_, surf_;
  std::vector<double> pab_, pnd_;  // the "ceres::Problem": 9 doubles (p, a, b) per edge factor, 7 (p, n, d) per surf factor
  int prev_map_[2] = {0, 0};
};

// F-LOAM's names for the same surface (BASELINE.json north_star).
class OdomEstimationClass : public EstimationMapping {
 public:
  void init(double edge_resolution, double surf_resolution) {
    ParamMap p;
    p.set("/EdgeLeafSize", edge_resolution);
    p.set("/SurfLeafSize", surf_resolution);
    initParameter(p);
  }
  void initMapWithPoints(const CloudPtr& edge_in, const CloudPtr& surf_in) { localMapInited(edge_in, surf_in); }
  void updatePointsToMap(const CloudPtr& edge_in, const CloudPtr& surf_in) { optimation_processing(edge_in, surf_in); }
  int addEdgeCostFactor(const CloudPtr& pc_in) { return EdgeCostFactor(pc_in); }
  int addSurfCostFactor(const CloudPtr& pc_in) { return SurfCostFactor(pc_in); }
  void addPointsToMap(const CloudPtr& downsampledEdgeCloud, const CloudPtr& downsampledSurfCloud) { createSubMap(downsampledEdgeCloud, downsampledSurfCloud); }
  void pointAssociateToMap(PointType const* const pi, PointType* const po) const { pointAssociaToMap(pi, po); }
  void downSamplingToMap(const CloudPtr& edge_pc_in, CloudPtr& edge_pc_out, const CloudPtr& surf_pc_in, CloudPtr& surf_pc_out) {
    voxelFilter(edge_pc_in, edgeMapLeafSize, edge_pc_out);
    voxelFilter(surf_pc_in, surfMapLeafSize, surf_pc_out);
  }
  void getMap(CloudPtr& laserCloudMap) {
    syncMapsToHost();
    *laserCloudMap = *localMapEdge;
    *laserCloudMap += *localMapSurf;
  }
  Isometry3d& odom() { return globalOdom; }
};

}  // namespace vilf

#ifndef VILF_NO_GLOBAL_NAMES
using vilf::EstimationMapping;  // feature_tracker_node.cpp:19 `EstimationMapping Estimator;`
#endif
