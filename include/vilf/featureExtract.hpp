// vilf/featureExtract.hpp — host-side mirror of the reference's `featureExtract` class, its ring-field / range-image stage 1
// (src/visual_inertial_lidar/feature_tracker/include/featureExtract.hpp:41-391, FX below): same method names, parameter names,
// defaults and output order; the work runs in libvilf_cuda.so (k_rangeimage.cu) through the C ABI with VILF_FLAG_RANGE_IMAGE.
//
//   reference member                                            here
//   initParam(ros::NodeHandle&)                       FX:83-93   initParam(NH&)   (template: ros::NodeHandle or vilf::ParamMap)
//   extractFeature(PointCloud2&, edge, surf)          FX:96-115  extractFeature(msg, edge, surf): x / y / z / intensity / ring fields
//                                                                 of the message -> vilf_feature_extract(xyzi, ring) + vilf_get_features
//   (after pcl::moveFromROSMsg, FX:99)                            extractFeatureFromPoints(points_with_ring, edge, surf) for an already converted cloud
//   projectPointCloud / inverProjectCloud / extractSmoothness /   device kernels; not callable separately (they are members nothing
//   markBadPoints / featureEdge_Surf                  FX:118-370  outside the class calls)
//
// The reference's extractFeature is declared bool and returns nothing (FX:96-115); here it returns true.  The features stay resident
// on the device, so an EstimationMapping sharing the session (shareSession) consumes them without a second upload.
#pragma once

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "cloud.hpp"
#include "session.hpp"

namespace vilf {

struct VelodynePointXYZIRT {  // FX:26-32 without PCL's padding
  float x, y, z, intensity;
  std::uint16_t ring;
};
typedef VelodynePointXYZIRT PointXYZIRT;

class featureExtract {
 public:
  featureExtract() : N_SCAN(64), Horizon_SCAN(1800), downsampleRate(1), edgeThreshold(1.0), surfThreshold(0.1), SurfLeafSize(0.4), lidarMinDis(3.0), lidarMaxDis(200.0),
                     sess_(std::make_shared<Session>()) {}

  // FX:83-93: same names (no leading slash here, as in the reference) and code defaults
  template <class NH>
  void initParam(NH& nh) {
    nh.template param<int>("Horizon_SCAN", Horizon_SCAN, 1800);
    nh.template param<int>("N_SCAN", N_SCAN, 64);
    nh.template param<int>("downsampleRate", downsampleRate, 1);
    nh.template param<double>("lidarMinRange", lidarMinDis, 3.0);
    nh.template param<double>("lidarMaxRange", lidarMaxDis, 200.0);
    nh.template param<double>("edgeThreshold", edgeThreshold, 1.0);
    nh.template param<double>("surfThreshold", surfThreshold, 0.1);
    nh.template param<double>("SurfLeafSize", SurfLeafSize, 0.4);  // read; the per-ring voxel filter is commented out in the reference (FX:215-219)
    applyConfig();
  }

  // FX:96-115 on a sensor_msgs::PointCloud2-like message (.data, .width, .height, .point_step, .fields[i].{name, offset}) carrying
  // float32 x, y, z, intensity and uint16 ring.
  template <class MsgT>
  bool extractFeature(MsgT& cloud_Msg, CloudPtr& cloud_Edge, CloudPtr& cloud_Surf) {
    int ox = -1, oy = -1, oz = -1, oi = -1, orr = -1;
    for (std::size_t f = 0; f < cloud_Msg.fields.size(); ++f) {
      const std::string& nm = cloud_Msg.fields[f].name;
      if (nm == "x") ox = (int)cloud_Msg.fields[f].offset;
      else if (nm == "y") oy = (int)cloud_Msg.fields[f].offset;
      else if (nm == "z") oz = (int)cloud_Msg.fields[f].offset;
      else if (nm == "intensity") oi = (int)cloud_Msg.fields[f].offset;
      else if (nm == "ring") orr = (int)cloud_Msg.fields[f].offset;
    }
    if (orr < 0) throw std::runtime_error("featureExtract::extractFeature: the message has no ring field");
    const int n = (int)(cloud_Msg.width * cloud_Msg.height);
    scan_.resize((std::size_t)(n > 0 ? n : 1) * 4);
    ring_.resize((std::size_t)(n > 0 ? n : 1));
    if (vilf_pack_pointcloud2(cloud_Msg.data.data(), n, (int)cloud_Msg.point_step, ox, oy, oz, oi, scan_.data()) != 0)
      throw std::runtime_error("featureExtract::extractFeature: the message has no float32 x / y / z fields");
    for (int i = 0; i < n; ++i) std::memcpy(&ring_[(std::size_t)i], cloud_Msg.data.data() + (std::size_t)i * cloud_Msg.point_step + orr, 2);
    return run(n, cloud_Edge, cloud_Surf);
  }
  // the same on an already converted cloud of points with .x .y .z .intensity .ring (pcl::PointCloud<VelodynePointXYZIRT>, FX:99)
  template <class P>
  bool extractFeatureFromPoints(const std::vector<P>& points, CloudPtr& cloud_Edge, CloudPtr& cloud_Surf) {
    const int n = (int)points.size();
    scan_.resize((std::size_t)(n > 0 ? n : 1) * 4);
    ring_.resize((std::size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) {
      scan_[4 * (std::size_t)i + 0] = points[(std::size_t)i].x; scan_[4 * (std::size_t)i + 1] = points[(std::size_t)i].y;
      scan_[4 * (std::size_t)i + 2] = points[(std::size_t)i].z; scan_[4 * (std::size_t)i + 3] = points[(std::size_t)i].intensity;
      ring_[(std::size_t)i] = points[(std::size_t)i].ring;
    }
    return run(n, cloud_Edge, cloud_Surf);
  }

  const SessionPtr& session() const { return sess_; }
  void shareSession(const SessionPtr& s) {
    sess_ = s;
    applyConfig();
  }

 private:
  bool run(int n, CloudPtr& cloud_Edge, CloudPtr& cloud_Surf) {
    applyConfig();
    if (!sess_->created() && n > sess_->config().max_scan_points) sess_->config().max_scan_points = n + n / 4;
    vilf_handle* h = sess_->handle();
    int ne = 0, ns = 0;
    sess_->check(vilf_feature_extract(h, scan_.data(), n, ring_.data(), &ne, &ns), "featureExtract::extractFeature");
    edge_.resize((std::size_t)(ne > 0 ? ne : 1) * 4);
    surf_.resize((std::size_t)(ns > 0 ? ns : 1) * 4);
    int got = 0;
    sess_->check(vilf_get_features(h, 0, edge_.data(), nullptr, ne > 0 ? ne : 1, &got), "featureExtract::extractFeature(edge)");
    sess_->check(vilf_get_features(h, 1, surf_.data(), nullptr, ns > 0 ? ns : 1, &got), "featureExtract::extractFeature(surf)");
    const bool fresh = cloud_Edge->points.empty() && cloud_Surf->points.empty();
    append_cloud(*cloud_Edge, edge_.data(), (std::size_t)ne);   // FX:149 push_back order
    append_cloud(*cloud_Surf, surf_.data(), (std::size_t)ns);   // FX:222 ring by ring
    if (fresh) {
      Session::Resident& r = sess_->resident;
      r.tag = sess_->new_tag();
      r.n_edge = (std::size_t)ne; r.n_surf = (std::size_t)ns;
      r.hash_edge = Session::content_hash(*cloud_Edge);
      r.hash_surf = Session::content_hash(*cloud_Surf);
    } else {
      sess_->invalidate_resident();
    }
    return true;
  }
  void applyConfig() {
    if (sess_->created()) return;  // fixed once the device state exists
    vilf_config& c = sess_->config();
    c.n_scan = 0;                  // ring ids come with every scan (FX:341)
    c.n_rings = N_SCAN;
    c.horizon_scan = Horizon_SCAN;
    c.downsample_rate = downsampleRate;
    c.lidar_min = lidarMinDis;
    c.lidar_max = lidarMaxDis;
    c.ri_edge_threshold = edgeThreshold;
    c.ri_surf_threshold = surfThreshold;
    c.flags |= VILF_FLAG_RANGE_IMAGE;
  }

  int N_SCAN;
  int Horizon_SCAN;
  int downsampleRate;
  double edgeThreshold;
  double surfThreshold;
  double SurfLeafSize;
  double lidarMinDis, lidarMaxDis;
  SessionPtr sess_;
  std::vector<float> scan_, edge_, surf_;
  std::vector<std::uint16_t> ring_;
};

}  // namespace vilf

#ifndef VILF_NO_GLOBAL_NAMES
using vilf::featureExtract;
#endif
