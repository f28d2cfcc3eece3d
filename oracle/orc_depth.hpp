// TEST INFRASTRUCTURE — CPU oracle for the lidar-depth association of visual features.  NOT product code.
// PARITY UNPINNED like the rest of the oracle (the reference has no tests and cannot be built here).
//
// Restates, line by line, from /root/reference/src/visual_inertial_lidar/feature_tracker/feature_tracker_node.cpp:
//   camera_cloud()   NODE:348-361  field-of-view filter (p.x > 0, |p.y/p.x| <= 10, |p.z/p.x| <= 10) followed by
//                    pcl::transformPointCloud(cloud, cloud, LIDAR_CAMERA_EX) — PCL 1.7.2 common/impl/transforms.hpp:
//                    out = float(T(r,0)*x + T(r,1)*y + T(r,2)*z + T(r,3)) evaluated left to right in the scalar type of the
//                    matrix (double, parameters.h:45)
//   feature_depth()  NODE:54-140   steps 4.1-4.4 of getFeatureDepth: features and cloud on the unit sphere, 3-NN
//                    (pcl::KdTreeFLANN, exact, fp32 L2_Simple), plane through the three neighbours intersected with the
//                    feature ray, the range sanity rules, depth = z of the scaled feature when > 2.0
// The 3-NN is a brute-force search with the oracle's canonical (d^2, index) order — the same neighbours FLANN's exact
// search returns, up to equal-distance ties (tie class T2).  num_bins = 360 (NODE:51).
#pragma once
#include <cmath>
#include <vector>
#include "orc_pipeline.hpp"

namespace orc {

inline void camera_cloud(const P4* scan, int n, const double T[16], Cloud& out) {
  out.clear();
  for (int i = 0; i < n; ++i) {
    const P4 p = scan[i];
    if (p.x > 0 && std::abs(p.y / p.x) <= 10 && std::abs(p.z / p.x) <= 10) {  // NODE:351-355 (arctan(10) = 84 deg)
      const double x = p.x, y = p.y, z = p.z;
      P4 q;
      q.x = static_cast<float>(T[0] * x + T[1] * y + T[2] * z + T[3]);
      q.y = static_cast<float>(T[4] * x + T[5] * y + T[6] * z + T[7]);
      q.z = static_cast<float>(T[8] * x + T[9] * y + T[10] * z + T[11]);
      q.i = p.i;
      out.push_back(q);
    }
  }
}

// feats: [m][3] normalised image coordinates (x, y, 1); depth_out[m] = -1 where no reliable depth exists (NODE:57-58)
inline void feature_depth(const P4* cloud, int n, const float* feats, int m, int num_bins, float* depth_out, int* nn_out /*[m][3] or null*/) {
  for (int i = 0; i < m; ++i) depth_out[i] = -1.0f;
  if (nn_out) for (int i = 0; i < 3 * m; ++i) nn_out[i] = -1;
  // 4.2: cloud on the unit sphere, range kept in intensity (NODE:78-89)
  const float bin_res = 180.0f / (float)num_bins;
  std::vector<P4> sph((size_t)n);
  for (int i = 0; i < n; ++i) {
    P4 p = cloud[i];
    const float range = std::sqrt(p.x * p.x + p.y * p.y + p.z * p.z);  // pointDistance, common.h:54-57
    p.x /= range; p.y /= range; p.z /= range;
    p.i = range;
    sph[(size_t)i] = p;
  }
  if (n < 10) return;  // NODE:91-95
  const float thr = (float)std::pow(std::sin(bin_res / 180.0 * M_PI) * 5.0, 2);  // NODE:103
  for (int f = 0; f < m; ++f) {
    // 4.1: feature on the unit sphere (Eigen::Vector3f::normalize: v /= sqrt(x^2 + y^2 + z^2), NODE:64-66)
    float vx = feats[3 * f], vy = feats[3 * f + 1], vz = feats[3 * f + 2];
    const float nrm = std::sqrt(vx * vx + vy * vy + vz * vz);
    vx /= nrm; vy /= nrm; vz /= nrm;
    // 4.3: exact 3-NN, fp32 ((dx*dx)+dy*dy)+dz*dz, ascending (d^2, index)
    float bd[3] = {3.4e38f, 3.4e38f, 3.4e38f};
    int bi[3] = {-1, -1, -1};
    for (int i = 0; i < n; ++i) {
      const float dx = vx - sph[(size_t)i].x, dy = vy - sph[(size_t)i].y, dz = vz - sph[(size_t)i].z;
      const float d = dx * dx + dy * dy + dz * dz;
      if (d < bd[2]) {  // strictly: first visited wins among equals, and i ascends
        int k = 2;
        while (k > 0 && d < bd[k - 1]) { bd[k] = bd[k - 1]; bi[k] = bi[k - 1]; --k; }
        bd[k] = d; bi[k] = i;
      }
    }
    if (nn_out) for (int k = 0; k < 3; ++k) nn_out[3 * f + k] = bi[k];
    if (bi[2] >= 0 && bd[2] < thr) {  // NODE:107
      const P4 &a = sph[(size_t)bi[0]], &b = sph[(size_t)bi[1]], &c = sph[(size_t)bi[2]];
      const float r1 = a.i, r2 = b.i, r3 = c.i;
      const float A[3] = {a.x * r1, a.y * r1, a.z * r1}, B[3] = {b.x * r2, b.y * r2, b.z * r2}, Cc[3] = {c.x * r3, c.y * r3, c.z * r3};
      const float ab[3] = {A[0] - B[0], A[1] - B[1], A[2] - B[2]}, bc[3] = {B[0] - Cc[0], B[1] - Cc[1], B[2] - Cc[2]};
      const float N[3] = {ab[1] * bc[2] - ab[2] * bc[1], ab[2] * bc[0] - ab[0] * bc[2], ab[0] * bc[1] - ab[1] * bc[0]};  // (A-B) x (B-C), NODE:129
      float s = (N[0] * A[0] + N[1] * A[1] + N[2] * A[2]) / (N[0] * vx + N[1] * vy + N[2] * vz);  // NODE:130-131
      const float min_depth = std::min(r1, std::min(r2, r3)), max_depth = std::max(r1, std::max(r2, r3));
      if (max_depth - min_depth > 2 || s <= 0.5) continue;  // NODE:135-137 (a NaN s passes both tests, as in the reference)
      else if (s - max_depth > 0) s = max_depth;
      else if (s - min_depth < 0) s = min_depth;
      const float z = vz * s;  // NODE:146-148: intensity = z of the scaled feature
      if (z > 2.0) depth_out[f] = z;  // NODE:155-159
    }
  }
}

}  // namespace orc
