// vilf/featureExtraction.hpp — host-side mirror of the reference's `featureExtraction` class
// (src/visual_inertial_lidar/feature_tracker/include/featureExtraction.hpp:35-246), same method names, argument
// meaning and output order; the work runs in libvilf_cuda.so (sm_100a kernels) through the C ABI of include/vilf.h.
//
//   reference member                                   here
//   initParam(ros::NodeHandle&)              FE:43-52  initParam(NH&)            (template: ros::NodeHandle or vilf::ParamMap)
//   extractFeature(cloud_in, edge, surf)     FE:223    extractFeature(...)       vilf_feature_extract + vilf_get_features
//   getLaserCloud / featureEdge_Surf /       FE:54-220 device kernels (k_extract.cu); not callable separately, as in
//   featureExtractionFromSector                        the reference nothing outside the class calls them
//
// F-LOAM names (BASELINE.json north_star): LaserProcessingClass::featureExtraction == extractFeature (alias below).
#pragma once

#include <cstring>
#include <vector>

#include "cloud.hpp"
#include "session.hpp"

namespace vilf {

class featureExtraction {
 public:
  featureExtraction() : N_SCANS(64), edgeThreshold(0.1), surfThreshold(0.1), SurfLeafSize(0.4), lidarMinDis(3.0), lidarMaxDis(100.0), sess_(std::make_shared<Session>()) {}

  // FE:43-52.  Same parameter names and the same code defaults as the reference.
  template <class NH>
  void initParam(NH& nh) {
    nh.template param<int>("/N_SCAN", N_SCANS, 64);
    nh.template param<double>("/lidarMinRange", lidarMinDis, 3.0);
    nh.template param<double>("/lidarMaxRange", lidarMaxDis, 100.0);
    nh.template param<double>("/edgeThreshold", edgeThreshold, 0.1);
    nh.template param<double>("/surfThreshold", surfThreshold, 0.1);  // read but unused, as in the reference
    nh.template param<double>("/SurfLeafSize", SurfLeafSize, 0.4);    // read but unused, as in the reference
    applyConfig();
  }

  // FE:223-232.  Appends the edge and surf features of `cloud_in` to the caller-owned clouds, in the reference's
  // order: rings ascending, 6 sectors per ring, edges in pick order (largest curvature first), surfs in ascending
  // curvature.  The features also stay resident on the device for EstimationMapping (see session.hpp).
  void extractFeature(const CloudPtr& cloud_in, CloudPtr& cloud_Edge, CloudPtr& cloud_Surf) {
    applyConfig();
    pack_cloud(*cloud_in, scan_);
    const int n = static_cast<int>(cloud_in->points.size());
    if (!sess_->created() && n > sess_->config().max_scan_points) sess_->config().max_scan_points = n + n / 4;
    vilf_handle* h = sess_->handle();
    int ne = 0, ns = 0;
    sess_->check(vilf_feature_extract(h, scan_.data(), n, nullptr, &ne, &ns), "extractFeature");
    edge_.resize(static_cast<std::size_t>(ne > 0 ? ne : 1) * 4);
    surf_.resize(static_cast<std::size_t>(ns > 0 ? ns : 1) * 4);
    int got = 0;
    sess_->check(vilf_get_features(h, 0, edge_.data(), nullptr, ne > 0 ? ne : 1, &got), "extractFeature(edge)");
    sess_->check(vilf_get_features(h, 1, surf_.data(), nullptr, ns > 0 ? ns : 1, &got), "extractFeature(surf)");
    const bool fresh = cloud_Edge->points.empty() && cloud_Surf->points.empty();
    append_cloud(*cloud_Edge, edge_.data(), static_cast<std::size_t>(ne));
    append_cloud(*cloud_Surf, surf_.data(), static_cast<std::size_t>(ns));
    if (fresh) {  // the device copy equals the caller's clouds: remember that (session.hpp)
      Session::Resident& r = sess_->resident;
      r.tag = sess_->new_tag();
      r.n_edge = static_cast<std::size_t>(ne);
      r.n_surf = static_cast<std::size_t>(ns);
      r.hash_edge = Session::content_hash(*cloud_Edge);
      r.hash_surf = Session::content_hash(*cloud_Surf);
    } else {
      sess_->invalidate_resident();
    }
  }

  const SessionPtr& session() const { return sess_; }
  void shareSession(const SessionPtr& s) {
    sess_ = s;
    applyConfig();
  }

 private:
  void applyConfig() {
    if (sess_->created()) return;  // fixed once the device state exists
    vilf_config& c = sess_->config();
    c.n_scan = N_SCANS;
    c.n_rings = N_SCANS;
    c.lidar_min = lidarMinDis;
    c.lidar_max = lidarMaxDis;
    c.edge_threshold = edgeThreshold;
  }

  int N_SCANS;
  double edgeThreshold;
  double surfThreshold;
  double SurfLeafSize;
  double lidarMinDis, lidarMaxDis;
  SessionPtr sess_;
  std::vector<float> scan_, edge_, surf_;
};

// F-LOAM's name for the same stage (north_star): LaserProcessingClass::featureExtraction(pc_in, pc_out_edge, pc_out_surf).
class LaserProcessingClass : public featureExtraction {
 public:
  void featureExtraction(const CloudPtr& pc_in, CloudPtr& pc_out_edge, CloudPtr& pc_out_surf) { extractFeature(pc_in, pc_out_edge, pc_out_surf); }
};

}  // namespace vilf

#ifndef VILF_NO_GLOBAL_NAMES
// The reference's node instantiates `featureExtraction featureExtractFactor;` at file scope
// (feature_tracker_node.cpp:18): make the unqualified name resolve when this header replaces the reference's.
using vilf::featureExtraction;
#endif
