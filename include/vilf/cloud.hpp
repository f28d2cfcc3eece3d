// vilf/cloud.hpp — the few PCL / Eigen / ROS types the reference's hot-path classes expose, without PCL, Eigen or ROS.
//
// The reference's classes (featureExtraction.hpp, EstimationMapping.hpp) speak pcl::PointCloud<pcl::PointXYZI>::Ptr,
// Eigen::Isometry3d and ros::NodeHandle.  A ROS build defines VILF_WITH_PCL and/or VILF_WITH_EIGEN and gets exactly
// those types; a bare build (this repository's tests, any non-ROS host) gets layout-compatible stand-ins declared
// here.  Either way the point payload that crosses the C ABI is the packed float[n][4] = x, y, z, intensity.
#pragma once

#include <cstddef>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../vilf.h"

#ifdef VILF_WITH_PCL
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#endif
#ifdef VILF_WITH_EIGEN
#include <Eigen/Geometry>
#endif

namespace vilf {

#ifdef VILF_WITH_PCL
typedef pcl::PointXYZI PointType;  // common.h:25
typedef pcl::PointCloud<PointType> Cloud;
typedef Cloud::Ptr CloudPtr;
inline CloudPtr make_cloud() { return CloudPtr(new Cloud()); }
#else
// Payload of pcl::PointXYZI (common.h:25) without PCL's SSE padding: 16 bytes.
struct PointXYZI {
  float x, y, z, intensity;
};
typedef PointXYZI PointType;
// The subset of pcl::PointCloud the reference's hot path uses: points, size(), clear(), push_back(), operator+=.
template <class P>
struct PointCloud {
  typedef std::shared_ptr<PointCloud<P> > Ptr;
  std::vector<P> points;
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); }
  void push_back(const P& p) { points.push_back(p); }
  PointCloud& operator+=(const PointCloud& o) {
    points.insert(points.end(), o.points.begin(), o.points.end());
    return *this;
  }
};
typedef PointCloud<PointType> Cloud;
typedef Cloud::Ptr CloudPtr;
inline CloudPtr make_cloud() { return std::make_shared<Cloud>(); }
#endif

// Any cloud whose points have .x .y .z .intensity -> packed float[n][4] and back (appending, like push_back).
template <class CloudT>
inline void pack_cloud(const CloudT& c, std::vector<float>& out) {
  const std::size_t n = c.points.size();
  out.resize(n * 4);
  for (std::size_t i = 0; i < n; ++i) {
    out[4 * i + 0] = c.points[i].x;
    out[4 * i + 1] = c.points[i].y;
    out[4 * i + 2] = c.points[i].z;
    out[4 * i + 3] = c.points[i].intensity;
  }
}
template <class CloudT>
inline void append_cloud(CloudT& c, const float* xyzi, std::size_t n) {
  for (std::size_t i = 0; i < n; ++i) {
    PointType p;
    std::memset(&p, 0, sizeof(p));
    p.x = xyzi[4 * i + 0];
    p.y = xyzi[4 * i + 1];
    p.z = xyzi[4 * i + 2];
    p.intensity = xyzi[4 * i + 3];
    c.push_back(p);
  }
}

// sensor_msgs::PointCloud2 (any type with .data, .width, .height, .point_step, .fields[i].{name, offset}) -> cloud, the
// job of pcl::fromROSMsg at feature_tracker_node.cpp:339-340, through vilf_pack_pointcloud2.
template <class MsgT, class CloudT>
inline void from_pointcloud2(const MsgT& msg, CloudT& cloud) {
  int ox = -1, oy = -1, oz = -1, oi = -1;
  for (std::size_t f = 0; f < msg.fields.size(); ++f) {
    if (msg.fields[f].name == "x") ox = (int)msg.fields[f].offset;
    else if (msg.fields[f].name == "y") oy = (int)msg.fields[f].offset;
    else if (msg.fields[f].name == "z") oz = (int)msg.fields[f].offset;
    else if (msg.fields[f].name == "intensity") oi = (int)msg.fields[f].offset;
  }
  const int n = (int)(msg.width * msg.height);
  std::vector<float> packed((std::size_t)(n > 0 ? n : 1) * 4);
  if (vilf_pack_pointcloud2(msg.data.data(), n, (int)msg.point_step, ox, oy, oz, oi, packed.data()) != 0)
    throw std::runtime_error("from_pointcloud2: the message has no float32 x / y / z fields");
  cloud.clear();
  append_cloud(cloud, packed.data(), (std::size_t)n);
}

#ifdef VILF_WITH_EIGEN
typedef Eigen::Isometry3d Isometry3d;
inline void iso_from_rt12(Isometry3d& T, const double rt[12]) {
  T = Isometry3d::Identity();
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T.linear()(i, j) = rt[3 * i + j];
    T.translation()(i) = rt[9 + i];
  }
}
inline void iso_to_rt12(const Isometry3d& T, double rt[12]) {
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) rt[3 * i + j] = T.linear()(i, j);
    rt[9 + i] = T.translation()(i);
  }
}
#else
// Stand-in for Eigen::Isometry3d as the node reads it (feature_tracker_node.cpp:388-389): rotation + translation.
struct Isometry3d {
  double R[9];  // row-major
  double t[3];
  Isometry3d() { setIdentity(); }
  void setIdentity() {
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    t[0] = t[1] = t[2] = 0.0;
  }
  static Isometry3d Identity() { return Isometry3d(); }
  const double* rotation() const { return R; }
  const double* translation() const { return t; }
};
inline void iso_from_rt12(Isometry3d& T, const double rt[12]) {
  std::memcpy(T.R, rt, 9 * sizeof(double));
  std::memcpy(T.t, rt + 9, 3 * sizeof(double));
}
inline void iso_to_rt12(const Isometry3d& T, double rt[12]) {
  std::memcpy(rt, T.R, 9 * sizeof(double));
  std::memcpy(rt + 9, T.t, 3 * sizeof(double));
}
#endif

// Stand-in for ros::NodeHandle::param<T>(name, variable, default) (featureExtraction.hpp:45-50,
// EstimationMapping.hpp:82-83): a flat name -> value table, e.g. filled from the reference's YAML
// (config/kitti/velodyne_param_64.yaml:9-23).  initParam / initParameter are templates, so a real ros::NodeHandle
// works unchanged.
class ParamMap {
 public:
  void set(const std::string& name, double v) { values_[name] = v; }
  template <class T>
  bool param(const std::string& name, T& var, const T& def) const {
    std::map<std::string, double>::const_iterator it = values_.find(name);
    if (it == values_.end()) {
      var = def;
      return false;
    }
    var = static_cast<T>(it->second);
    return true;
  }

 private:
  std::map<std::string, double> values_;
};

}  // namespace vilf
