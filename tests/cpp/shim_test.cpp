// Drives the host-side mirror classes (include/vilf/featureExtraction.hpp, include/vilf/EstimationMapping.hpp) the way
// the reference's ROS node does (feature_tracker_node.cpp:339-389), three ways over the same scans:
//   A  node order, ONE shared device session (features stay resident between the two classes)
//   B  node order, independent objects exactly as the reference declares them (features cross the host)
//   C  the public pieces one by one: predictPose, voxelFilter, 2 x (ProblemReset, EdgeCostFactor, SurfCostFactor,
//      SolveProblem), createSubMap  — the body of optimation_processing (EM:235-296) spelled out by the caller
// and writes the per-frame poses of A; A == B bit for bit, C == A to 1e-9 (the factor list of C is compacted, so the
// reduction order of the normal equations differs).  The Python test compares A with the CPU oracle.
//
// usage: shim_test <scans.bin> <n_scan> <poses_out.bin>
//   scans.bin: int32 frames, then per frame: int32 n, float32[n][4]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "vilf/EstimationMapping.hpp"
#include "vilf/featureDepth.hpp"
#include "vilf/nodeOutputs.hpp"
#include "vilf/Scancontext.hpp"
#include "vilf/featureExtract.hpp"

using vilf::Cloud;
using vilf::CloudPtr;

static std::vector<CloudPtr> read_scans(const char* path) {
  std::vector<CloudPtr> out;
  FILE* f = std::fopen(path, "rb");
  if (!f) { std::perror(path); std::exit(2); }
  int frames = 0;
  if (std::fread(&frames, 4, 1, f) != 1) std::exit(2);
  for (int i = 0; i < frames; ++i) {
    int n = 0;
    if (std::fread(&n, 4, 1, f) != 1) std::exit(2);
    std::vector<float> buf((size_t)n * 4);
    if (n && std::fread(buf.data(), 16, (size_t)n, f) != (size_t)n) std::exit(2);
    CloudPtr c = vilf::make_cloud();
    vilf::append_cloud(*c, buf.data(), (size_t)n);
    out.push_back(c);
  }
  std::fclose(f);
  return out;
}

// The payloads the node publishes (NODE:385-446): chaining the /Odometry relative poses must reproduce the /path pose, and
// the /GlobalMap message must decode to the cloud it was made from.
static void check_outputs(vilf::NodeOutputs& out, const EstimationMapping& est, const CloudPtr& MapCloud, double chain[7]) {
  out.update(est.globalOdom);
  const vilf::PoseMsg& r = out.odometry;
  const vilf::PoseMsg& a = out.pathPose;
  const double q[4] = {chain[0], chain[1], chain[2], chain[3]}, t[3] = {chain[4], chain[5], chain[6]};
  // chain <- chain * relative
  const double v[3] = {r.x, r.y, r.z};
  double uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
  for (int c = 0; c < 3; ++c) uv[c] *= 2;
  const double w[3] = {q[1] * uv[2] - q[2] * uv[1], q[2] * uv[0] - q[0] * uv[2], q[0] * uv[1] - q[1] * uv[0]};
  for (int c = 0; c < 3; ++c) chain[4 + c] = t[c] + v[c] + q[3] * uv[c] + w[c];
  chain[0] = q[3] * r.qx + q[0] * r.qw + q[1] * r.qz - q[2] * r.qy;
  chain[1] = q[3] * r.qy + q[1] * r.qw + q[2] * r.qx - q[0] * r.qz;
  chain[2] = q[3] * r.qz + q[2] * r.qw + q[0] * r.qy - q[1] * r.qx;
  chain[3] = q[3] * r.qw - q[0] * r.qx - q[1] * r.qy - q[2] * r.qz;
  const double ref[7] = {a.qx, a.qy, a.qz, a.qw, a.x, a.y, a.z};
  const double sgn = (chain[3] * ref[3] + chain[0] * ref[0] + chain[1] * ref[1] + chain[2] * ref[2]) < 0 ? -1.0 : 1.0;
  for (int c = 0; c < 7; ++c) {
    const double d = std::fabs((c < 4 ? sgn : 1.0) * chain[c] - ref[c]);
    if (!(d < 1e-9)) { std::fprintf(stderr, "chained /Odometry differs from /path by %.3e in component %d\n", d, c); std::exit(1); }
  }
  const double* po = est.parameter_opti;
  const double sgn2 = (po[0] * ref[0] + po[1] * ref[1] + po[2] * ref[2] + po[3] * ref[3]) < 0 ? -1.0 : 1.0;
  for (int c = 0; c < 7; ++c) {
    if (!(std::fabs(ref[c] - (c < 4 ? sgn2 : 1.0) * po[c]) < 1e-12)) { std::fprintf(stderr, "/path pose differs from parameter_opti in component %d\n", c); std::exit(1); }
  }
  vilf::PointCloud2 msg;
  vilf::to_pointcloud2(*MapCloud, msg);
  CloudPtr back = vilf::make_cloud();
  vilf::from_pointcloud2(msg, *back);
  if (msg.point_step != 32 || msg.width != MapCloud->points.size() || back->points.size() != MapCloud->points.size()) { std::fprintf(stderr, "bad /GlobalMap message\n"); std::exit(1); }
  for (size_t i = 0; i < back->points.size(); ++i) {
    const vilf::PointType &p = back->points[i], &o = MapCloud->points[i];
    if (p.x != o.x || p.y != o.y || p.z != o.z || p.intensity != o.intensity) { std::fprintf(stderr, "/GlobalMap message does not decode to MapCloud\n"); std::exit(1); }
  }
}

static void node_loop(featureExtraction& fe, EstimationMapping& est, const std::vector<CloudPtr>& scans, std::vector<double>& poses) {
  bool init_pub = false;
  vilf::NodeOutputs outputs;
  double chain[7] = {0, 0, 0, 1, 0, 0, 0};
  for (size_t i = 0; i < scans.size(); ++i) {
    CloudPtr MapCloud = vilf::make_cloud(), edge = vilf::make_cloud(), surf = vilf::make_cloud();
    fe.extractFeature(scans[i], edge, surf);              // NODE:346
    if (!init_pub) {
      init_pub = true;
      est.localMapInited(edge, surf);                     // NODE:373
      outputs.reset();                                    // NODE:376-377
    } else {
      est.optimation_processing(edge, surf);              // NODE:384
      est.getMapCloud(MapCloud);                          // NODE:385
      if (MapCloud->points.empty()) { std::fprintf(stderr, "empty /GlobalMap cloud at frame %zu\n", i); std::exit(1); }
      check_outputs(outputs, est, MapCloud, chain);      // NODE:388-446
    }
    for (int k = 0; k < 7; ++k) poses.push_back(est.parameter_opti[k]);
  }
}

int main(int argc, char** argv) {
  if (argc < 4) { std::fprintf(stderr, "usage: shim_test scans.bin n_scan poses_out.bin\n"); return 2; }
  try {
    std::vector<CloudPtr> scans = read_scans(argv[1]);
    vilf::ParamMap nh;  // config/kitti/velodyne_param_64.yaml:9-23
    nh.set("/N_SCAN", std::atof(argv[2]));
    nh.set("/lidarMinRange", 3.0); nh.set("/lidarMaxRange", 90.0); nh.set("/edgeThreshold", 0.1); nh.set("/surfThreshold", 0.1);
    nh.set("/EdgeLeafSize", 0.4); nh.set("/SurfLeafSize", 0.8);

    // A: shared session
    std::vector<double> pa, pb, pc;
    {
      featureExtraction featureExtractFactor;   // NODE:18
      EstimationMapping Estimator;              // NODE:19
      featureExtractFactor.initParam(nh);       // NODE:509
      Estimator.initParameter(nh);              // NODE:510
      Estimator.shareSession(featureExtractFactor);
      node_loop(featureExtractFactor, Estimator, scans, pa);
      // depth for visual features (NODE:348-366): from the scan that is still resident, and from a host-prepared cloud
      const double T[16] = {0, -1, 0, 0.05, 0, 0, -1, -0.1, 1, 0, 0, 0.02, 0, 0, 0, 1};  // LIDAR_CAMERA_EX, row-major
      std::vector<vilf::Point32> feats;
      for (int i = 0; i < 60; ++i) feats.push_back(vilf::Point32{-0.9f + 0.03f * i, -0.3f + 0.005f * i, 1.0f});
      std::vector<float> d1 = vilf::getFeatureDepthFromScan(*featureExtractFactor.session(), T, feats);
      CloudPtr cam = vilf::make_cloud();
      const Cloud& full = *scans.back();
      for (size_t i = 0; i < full.points.size(); ++i) {  // NODE:351-361 on the host
        const vilf::PointType p = full.points[i];
        if (p.x > 0 && std::fabs(p.y / p.x) <= 10 && std::fabs(p.z / p.x) <= 10) {
          vilf::PointType q = p;
          const double x = p.x, y = p.y, z = p.z;
          q.x = (float)(T[0] * x + T[1] * y + T[2] * z + T[3]);
          q.y = (float)(T[4] * x + T[5] * y + T[6] * z + T[7]);
          q.z = (float)(T[8] * x + T[9] * y + T[10] * z + T[11]);
          cam->push_back(q);
        }
      }
      std::vector<float> d2 = vilf::getFeatureDepth(*featureExtractFactor.session(), cam, feats);
      int with_depth = 0;
      for (size_t i = 0; i < feats.size(); ++i) {
        if (d1[i] != d2[i]) { std::fprintf(stderr, "depth from the resident scan and from the host-prepared cloud differ at %zu\n", i); return 1; }
        with_depth += d1[i] > 0;
      }
      std::printf("depth association: %d of %zu features have lidar depth\n", with_depth, feats.size());
      if (with_depth == 0) { std::fprintf(stderr, "no feature received a depth\n"); return 1; }
    }
    // B: independent objects
    {
      featureExtraction featureExtractFactor;
      EstimationMapping Estimator;
      featureExtractFactor.initParam(nh);
      Estimator.initParameter(nh);
      node_loop(featureExtractFactor, Estimator, scans, pb);
    }
    // C: F-LOAM names + the pieces one by one
    {
      vilf::LaserProcessingClass laserProcessing;
      vilf::OdomEstimationClass odomEstimation;
      laserProcessing.initParam(nh);
      odomEstimation.init(0.4, 0.8);
      for (size_t i = 0; i < scans.size(); ++i) {
        CloudPtr edge = vilf::make_cloud(), surf = vilf::make_cloud();
        laserProcessing.featureExtraction(scans[i], edge, surf);
        if (i == 0) {
          odomEstimation.initMapWithPoints(edge, surf);
        } else {
          odomEstimation.predictPose();
          CloudPtr de = vilf::make_cloud(), ds = vilf::make_cloud();
          odomEstimation.downSamplingToMap(edge, de, surf, ds);
          for (int it = 0; it < 2; ++it) {
            odomEstimation.ProblemReset();
            odomEstimation.addEdgeCostFactor(de);
            odomEstimation.addSurfCostFactor(ds);
            odomEstimation.SolveProblem(4);
          }
          odomEstimation.addPointsToMap(de, ds);
        }
        for (int k = 0; k < 7; ++k) pc.push_back(odomEstimation.parameter_opti[k]);
      }
      CloudPtr m = vilf::make_cloud();
      odomEstimation.getMap(m);
      if (m->points.empty()) { std::fprintf(stderr, "empty map\n"); return 1; }
      vilf::PointType pi = scans[0]->points[0], po;
      odomEstimation.pointAssociateToMap(&pi, &po);
      if (!(std::isfinite(po.x) && std::isfinite(po.y) && std::isfinite(po.z))) return 1;
    }
    {  // D: the loop detector of global_fusion (poseGraphOptimization.cpp:553, :600-603) on the same scans, as key frames
      vilf::SCManager scManager;
      scManager.NUM_EXCLUDE_RECENT = 2; scManager.TREE_MAKING_PERIOD_ = 1;
      scManager.setSCdistThres(0.3); scManager.setMaximumRadius(80.0);
      for (size_t i = 0; i < scans.size(); ++i) scManager.makeAndSaveScancontextAndKeys(*scans[i]);
      scManager.makeAndSaveScancontextAndKeys(*scans[0]);  // revisit of key frame 0
      std::pair<int, float> r = scManager.detectLoopClosureID();
      std::pair<double, int> d = scManager.distanceBtnScanContext((int)scans.size(), 0);
      if (r.first != 0 || r.second != 0.0f || d.first != 0.0 || d.second != 0 || scManager.size() != (int)scans.size() + 1) {
        std::fprintf(stderr, "SCManager mirror: expected loop 0 at distance 0, got id %d yaw %g dist %g shift %d\n", r.first, r.second, d.first, d.second);
        return 1;
      }
      std::vector<double> a = scManager.getScancontext(0), b = scManager.getScancontext(-1);
      if (a != b || scManager.distanceBtnScanContext(a, b).first != 0.0) { std::fprintf(stderr, "SCManager mirror: revisit descriptor differs\n"); return 1; }
    }
    if (argc > 4) {  // E (shim_test ... rings): the ring-field extractor class (featureExtract.hpp) in front of the same estimator
      const int n_rings = std::atoi(argv[4]);
      vilf::ParamMap nh2;
      nh2.set("N_SCAN", n_rings); nh2.set("Horizon_SCAN", 1800); nh2.set("lidarMaxRange", 90.0);
      nh2.set("/EdgeLeafSize", 0.4); nh2.set("/SurfLeafSize", 0.8);
      featureExtract extractor;
      EstimationMapping Estimator;
      extractor.initParam(nh2);
      Estimator.initParameter(nh2);
      Estimator.shareSessionWith(extractor);
      for (size_t i = 0; i < scans.size(); ++i) {
        std::vector<vilf::PointXYZIRT> pts(scans[i]->points.size());
        for (size_t k = 0; k < pts.size(); ++k) {
          const vilf::PointType& p = scans[i]->points[k];
          const double el = std::atan2((double)p.z, std::sqrt((double)p.x * p.x + (double)p.y * p.y)) * 180.0 / M_PI;
          int ring = (int)std::floor((el + 92.0 / 3.0) * 3.0 / 4.0);  // the synthetic 32-beam layout (FE:85)
          pts[k].x = p.x; pts[k].y = p.y; pts[k].z = p.z; pts[k].intensity = p.intensity;
          pts[k].ring = (std::uint16_t)(ring < 0 ? 0 : ring >= n_rings ? n_rings - 1 : ring);
        }
        CloudPtr e = vilf::make_cloud(), s = vilf::make_cloud();
        extractor.extractFeatureFromPoints(pts, e, s);
        if (e->points.empty() || s->points.empty()) { std::fprintf(stderr, "featureExtract mirror: no features at frame %zu\n", i); return 1; }
        if (i == 0) Estimator.localMapInited(e, s); else Estimator.optimation_processing(e, s);
      }
      if (!(std::fabs(Estimator.parameter_opti[4]) > 0.5)) { std::fprintf(stderr, "featureExtract mirror: the estimator did not move (x = %g)\n", Estimator.parameter_opti[4]); return 1; }
      std::printf("featureExtract + EstimationMapping: x = %.3f after %zu frames\n", Estimator.parameter_opti[4], scans.size());
    }
    double dab = 0, dac = 0;
    for (size_t i = 0; i < pa.size(); ++i) {
      dab = std::fmax(dab, std::fabs(pa[i] - pb[i]));
      dac = std::fmax(dac, std::fabs(pa[i] - pc[i]));
    }
    std::printf("frames %zu  max|A-B| %.3e  max|A-C| %.3e\n", scans.size(), dab, dac);
    FILE* f = std::fopen(argv[3], "wb");
    if (!f) { std::perror(argv[3]); return 2; }
    std::fwrite(pa.data(), sizeof(double), pa.size(), f);
    std::fclose(f);
    if (dab != 0.0) { std::fprintf(stderr, "shared-session and host-round-trip runs differ\n"); return 1; }
    if (!(dac < 1e-9)) { std::fprintf(stderr, "piecewise run differs from optimation_processing\n"); return 1; }
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "shim_test: %s\n", e.what());
    return 3;
  }
}
