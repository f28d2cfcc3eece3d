// vilf/nodeOutputs.hpp — host-side mirror of what the reference's node does with the result of the lidar path each frame
// (src/visual_inertial_lidar/feature_tracker/feature_tracker_node.cpp:385-446): the pose read from Estimator.globalOdom,
// the RELATIVE pose it publishes on /Odometry, the absolute pose appended to /path, and the /GlobalMap cloud message.
// ROS publishing itself (tf broadcaster, publishers, stamps) stays in the node; these are the payloads.
//
//   reference (NODE)                                                     here -> C ABI
//   :376-377  q_last = identity, t_last = 0 (first frame)                NodeOutputs::reset()
//   :388-389  q_estimator(globalOdom.rotation()), t_estimator            NodeOutputs::update(globalOdom)   vilf_node_outputs
//   :400-401  q_relative, t_relative                                     .odometry  (laserOdometry.pose.pose, :409-415)
//   :423-436  laserOdometryPath / laserPath.poses.push_back              .pathPose, .path
//   :439-440  pcl::toROSMsg(*MapCloud, MapCloudMsg)                      to_pointcloud2(cloud, msg)        vilf_unpack_pointcloud2
//   :445-446  q_last = q_estimator, t_last = t_estimator                 inside update()
#pragma once

#include <stdexcept>
#include <string>
#include <vector>

#include "cloud.hpp"

namespace vilf {

struct PoseMsg {  // geometry_msgs::Pose: orientation x y z w, position x y z
  double qx, qy, qz, qw, x, y, z;
};

class NodeOutputs {
 public:
  NodeOutputs() { reset(); }
  void reset() {  // NODE:376-377
    const double id[7] = {0, 0, 0, 1, 0, 0, 0};
    for (int i = 0; i < 7; ++i) last_[i] = id[i];
    path.clear();
  }
  void update(const Isometry3d& globalOdom) {
    double rt[12], rel[7], abs[7];
    iso_to_rt12(globalOdom, rt);
    if (vilf_node_outputs(rt, last_, rel, abs) != VILF_OK) throw std::runtime_error("vilf_node_outputs failed");
    odometry = PoseMsg{rel[0], rel[1], rel[2], rel[3], rel[4], rel[5], rel[6]};
    pathPose = PoseMsg{abs[0], abs[1], abs[2], abs[3], abs[4], abs[5], abs[6]};
    path.push_back(pathPose);
  }
  PoseMsg odometry;           // /Odometry: pose of the current frame in the previous one
  PoseMsg pathPose;           // /path, tf world -> body
  std::vector<PoseMsg> path;  // laserPath.poses
  const double* last() const { return last_; }

 private:
  double last_[7];  // q_last (x y z w), t_last
};

// Stand-ins for sensor_msgs::PointField / PointCloud2 on a host without ROS; to_pointcloud2 is a template and fills the
// real message type just the same.
struct PointField {
  std::string name;
  unsigned offset;
  unsigned char datatype;  // 7 = FLOAT32
  unsigned count;
};
struct PointCloud2 {
  unsigned height, width;
  std::vector<PointField> fields;
  bool is_bigendian;
  unsigned point_step, row_step;
  std::vector<unsigned char> data;
  bool is_dense;
};

// pcl::toROSMsg for a PointXYZI cloud (NODE:439-440): height 1, fields x y z intensity at 0 4 8 16, point_step 32.
template <class CloudT, class MsgT>
inline void to_pointcloud2(const CloudT& cloud, MsgT& msg) {
  std::vector<float> packed;
  pack_cloud(cloud, packed);
  const unsigned n = (unsigned)cloud.points.size();
  static const char* const names[4] = {"x", "y", "z", "intensity"};
  static const unsigned offs[4] = {0, 4, 8, 16};
  msg.height = 1;
  msg.width = n;
  msg.fields.resize(4);
  for (int f = 0; f < 4; ++f) {
    msg.fields[f].name = names[f];
    msg.fields[f].offset = offs[f];
    msg.fields[f].datatype = 7;
    msg.fields[f].count = 1;
  }
  msg.is_bigendian = false;
  msg.point_step = 32;
  msg.row_step = 32 * n;
  msg.is_dense = true;
  msg.data.resize((std::size_t)32 * n);
  float dummy_in[4] = {0, 0, 0, 0};
  unsigned char dummy_out[32];
  if (vilf_unpack_pointcloud2(n ? packed.data() : dummy_in, (int)n, 32, 0, 4, 8, 16, n ? msg.data.data() : dummy_out) != VILF_OK)
    throw std::runtime_error("to_pointcloud2 failed");
}

}  // namespace vilf
