// vilf/featureDepth.hpp — host-side mirror of the lidar-depth association the reference's node runs for its visual
// features: getFeatureDepth (src/visual_inertial_lidar/feature_tracker/feature_tracker_node.cpp:54-199) and the camera-frame
// preparation of the scan in front of it (NODE:348-361).  The drawing of the depth image (step 4.5, NODE:164-197) is
// OpenCV / ROS publishing and stays in the node.
//
//   reference                                                       here -> C ABI
//   getFeatureDepth(depth_cloud_local, show_img, features_2d)       getFeatureDepth(session, depth_cloud_local, features_2d)   vilf_feature_depth(cloud)
//   NODE:348-361 filter + transformPointCloud, then getFeatureDepth getFeatureDepthFromScan(session, LIDAR_CAMERA_EX, features_2d)  vilf_feature_depth(NULL, T)
//
// The second form uses the scan that featureExtraction::extractFeature already uploaded for the odometry (same session):
// only the features go up and the depths come down.
#pragma once

#include <vector>

#include "cloud.hpp"
#include "session.hpp"

namespace vilf {

struct Point32 {  // geometry_msgs::Point32 as the node fills it (normalised image coordinates, z == 1)
  float x, y, z;
};

constexpr int kDepthNumBins = 360;  // NODE:51

// Returns the `values` of the reference's sensor_msgs::ChannelFloat32 "depth": one entry per feature, -1 = no depth.
template <class FeatureT>
inline std::vector<float> getFeatureDepth(Session& sess, const CloudPtr& depth_cloud_local, const std::vector<FeatureT>& features_2d) {
  std::vector<float> cloud, feats(features_2d.size() * 3), depth(features_2d.size(), -1.0f);
  pack_cloud(*depth_cloud_local, cloud);
  for (std::size_t i = 0; i < features_2d.size(); ++i) { feats[3 * i] = features_2d[i].x; feats[3 * i + 1] = features_2d[i].y; feats[3 * i + 2] = features_2d[i].z; }
  if (features_2d.empty()) return depth;
  sess.check(vilf_feature_depth(sess.handle(), cloud.empty() ? feats.data() : cloud.data(), (int)depth_cloud_local->points.size(), nullptr, feats.data(),
                                (int)features_2d.size(), kDepthNumBins, depth.data(), nullptr, nullptr),
             "getFeatureDepth");
  return depth;
}

// LIDAR_CAMERA_EX: 4x4 row-major (Eigen::Matrix4d is column-major: pass its transpose's data, or fill row by row).
template <class FeatureT>
inline std::vector<float> getFeatureDepthFromScan(Session& sess, const double LIDAR_CAMERA_EX[16], const std::vector<FeatureT>& features_2d) {
  std::vector<float> feats(features_2d.size() * 3), depth(features_2d.size(), -1.0f);
  for (std::size_t i = 0; i < features_2d.size(); ++i) { feats[3 * i] = features_2d[i].x; feats[3 * i + 1] = features_2d[i].y; feats[3 * i + 2] = features_2d[i].z; }
  if (features_2d.empty()) return depth;
  sess.check(vilf_feature_depth(sess.handle(), nullptr, 0, LIDAR_CAMERA_EX, feats.data(), (int)features_2d.size(), kDepthNumBins, depth.data(), nullptr, nullptr),
             "getFeatureDepthFromScan");
  return depth;
}

}  // namespace vilf
