// TEST INFRASTRUCTURE — CPU restatement of the reference's ScanContext loop-candidate detector:
// src/global_fusion/include/Scancontext/Scancontext.h (SCManager) and the helpers it uses from
// src/global_fusion/include/common.h (xy2theta :79-92, circshift :95-114, eig2stdvec :117+).
//
// PARITY UNPINNED: the reference ships no vectors for this either.  First-party arithmetic is restated line by line (float /
// double types as the reference's expressions have them under `using namespace std`, common.h:22: sqrt / atan of float arguments
// are the float overloads).  Third-party: Eigen 3.3.7 reductions (mean(), norm(), dot()) are restated as sequential sums — Eigen
// vectorises them in a different association order, so descriptor keys and distances are parity "to rounding" (1e-12 relative),
// not bit-exact; the descriptor itself (a max of floats per bin) is exact.  The ring-key search of detectLoopClosureID uses a
// nanoflann kd-tree in the reference (exact 3-NN, L2, float keys); restated as an exact brute-force 3-NN with (distance, index)
// order.  pcl::IterativeClosestPoint (poseGraphOptimization.cpp:376-444) is NOT restated.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>

namespace orc {

struct SCParams {
  double lidar_height = 2.0;     // Scancontext.h:313
  int num_ring = 20;             // :315
  int num_sector = 60;           // :316
  double max_radius = 80.0;      // :318 (setMaximumRadius)
  int num_exclude_recent = 30;   // :323
  int num_candidates = 3;        // :324
  double search_ratio = 0.1;     // :327
  double dist_thres = 0.2;       // :329 (setSCdistThres)
  int tree_making_period = 30;   // :331
};

// common.h:79-92.  The four quadrant tests use bitwise & on bools, the results are double expressions returned as float.
inline float sc_xy2theta(const float& x, const float& y) {
  if ((x >= 0) & (y >= 0)) return (float)((180 / M_PI) * std::atan(y / x));
  if ((x < 0) & (y >= 0)) return (float)(180 - ((180 / M_PI) * std::atan(y / (-x))));
  if ((x < 0) & (y < 0)) return (float)(180 + ((180 / M_PI) * std::atan(y / x)));
  if ((x >= 0) & (y < 0)) return (float)(360 - ((180 / M_PI) * std::atan((-y) / x)));
  return std::nanf("");  // NaN input: the reference falls off the end of the function (undefined); every comparison below is false
}

// Scancontext.h:42-83.  desc is num_ring x num_sector, row-major here: desc[r * num_sector + s].  pts: x, y, z, intensity.
inline void sc_make(const SCParams& P, const float* pts, int n, double* desc) {
  const int NO_POINT = -1000;
  const int R = P.num_ring, S = P.num_sector;
  for (int i = 0; i < R * S; ++i) desc[i] = NO_POINT;
  for (int i = 0; i < n; ++i) {
    float px = pts[4 * i + 0], py = pts[4 * i + 1];
    float pz = (float)(pts[4 * i + 2] + P.lidar_height);  // pt.z is a float member: the double sum is rounded (:56)
    float azim_range = std::sqrt(px * px + py * py);       // :59
    float azim_angle = sc_xy2theta(px, py);                // :60
    if (azim_range > P.max_radius) continue;               // :63
    int ring_idx = std::max(std::min(R, int(std::ceil((azim_range / P.max_radius) * R))), 1);    // :66
    int sctor_idx = std::max(std::min(S, int(std::ceil((azim_angle / 360.0) * S))), 1);          // :67
    double& d = desc[(ring_idx - 1) * S + (sctor_idx - 1)];
    if (d < pz) d = pz;  // :70-71
  }
  for (int i = 0; i < R * S; ++i)
    if (desc[i] == NO_POINT) desc[i] = 0;  // :75-78
}

// :86-99 rowwise mean, :102-115 columnwise mean (sequential sums, see the header)
inline void sc_ringkey(const SCParams& P, const double* desc, double* key) {
  for (int r = 0; r < P.num_ring; ++r) {
    double s = 0;
    for (int c = 0; c < P.num_sector; ++c) s += desc[r * P.num_sector + c];
    key[r] = s / P.num_sector;
  }
}
inline void sc_sectorkey(const SCParams& P, const double* desc, double* key) {
  for (int c = 0; c < P.num_sector; ++c) {
    double s = 0;
    for (int r = 0; r < P.num_ring; ++r) s += desc[r * P.num_sector + c];
    key[c] = s / P.num_ring;
  }
}

// :119-138.  circshift moves column c to (c + shift) % cols, so shifted[c] = v[(c - shift) mod cols].
inline int sc_fast_align(const SCParams& P, const double* vkey1, const double* vkey2) {
  const int S = P.num_sector;
  int argmin = 0;
  double mn = 10000000;
  for (int sh = 0; sh < S; ++sh) {
    double s = 0;
    for (int c = 0; c < S; ++c) { double d = vkey1[c] - vkey2[((c - sh) % S + S) % S]; s += d * d; }
    double nrm = std::sqrt(s);
    if (nrm < mn) { argmin = sh; mn = nrm; }
  }
  return argmin;
}

// :140-161 with sc2 circularly shifted by `shift` columns
inline double sc_dist_direct(const SCParams& P, const double* sc1, const double* sc2, int shift) {
  const int R = P.num_ring, S = P.num_sector;
  int num_eff = 0;
  double sum_sim = 0;
  for (int c = 0; c < S; ++c) {
    const int c2 = ((c - shift) % S + S) % S;
    double n1 = 0, n2 = 0, dot = 0;
    for (int r = 0; r < R; ++r) { double a = sc1[r * S + c], b = sc2[r * S + c2]; n1 += a * a; n2 += b * b; dot += a * b; }
    n1 = std::sqrt(n1); n2 = std::sqrt(n2);
    if ((n1 == 0) | (n2 == 0)) continue;  // :149
    sum_sim = sum_sim + dot / (n1 * n2);
    num_eff = num_eff + 1;
  }
  return 1.0 - sum_sim / num_eff;  // (0 / 0 = NaN when no sector pair is populated, as in the reference)
}

// :163-193
inline std::pair<double, int> sc_distance(const SCParams& P, const double* sc1, const double* sc2) {
  const int S = P.num_sector;
  std::vector<double> v1(S), v2(S);
  sc_sectorkey(P, sc1, v1.data());
  sc_sectorkey(P, sc2, v2.data());
  const int a = sc_fast_align(P, v1.data(), v2.data());
  const int radius = (int)std::round(0.5 * P.search_ratio * S);
  std::vector<int> space{a};
  for (int ii = 1; ii < radius + 1; ++ii) { space.push_back((a + ii + S) % S); space.push_back((a - ii + S) % S); }
  std::sort(space.begin(), space.end());
  int argmin = 0;
  double mn = 10000000;
  for (int sh : space) {
    double d = sc_dist_direct(P, sc1, sc2, sh);
    if (d < mn) { argmin = sh; mn = d; }
  }
  return {mn, argmin};
}

inline float sc_key_dist(const float* a, const float* b, int dim) {
  float result = 0;
  int d = 0;
  for (; d + 3 < dim; d += 4) {
    const float d0 = a[d] - b[d], d1 = a[d + 1] - b[d + 1], d2 = a[d + 2] - b[d + 2], d3 = a[d + 3] - b[d + 3];
    result += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  for (; d < dim; ++d) { const float d0 = a[d] - b[d]; result += d0 * d0; }
  return result;
}

// SCManager state + :196-299
struct SCManager {
  SCParams P;
  std::vector<std::vector<double>> descs;
  std::vector<std::vector<float>> invkeys;       // eig2stdvec: double -> float
  std::vector<std::vector<float>> tree_keys;     // snapshot the kd-tree was built from (:229-231)
  int tree_counter = 0;

  void add(const float* pts, int n) {  // makeAndSaveScancontextAndKeys :196-208
    std::vector<double> d((size_t)P.num_ring * P.num_sector), k(P.num_ring);
    sc_make(P, pts, n, d.data());
    sc_ringkey(P, d.data(), k.data());
    descs.push_back(d);
    invkeys.emplace_back(k.begin(), k.end());
  }
  // detectLoopClosureID :210-299 -> loop id (-1: none), yaw difference [rad]; min_dist / nn_idx as printed
  std::pair<int, float> detect(double* min_dist_out = nullptr, int* nn_idx_out = nullptr) {
    if (min_dist_out) *min_dist_out = 10000000;
    if (nn_idx_out) *nn_idx_out = 0;
    if ((int)invkeys.size() < P.num_exclude_recent + 1) return {-1, 0.0f};
    const std::vector<float>& cur = invkeys.back();
    const std::vector<double>& cur_desc = descs.back();
    if (tree_counter % P.tree_making_period == 0) tree_keys.assign(invkeys.begin(), invkeys.end() - P.num_exclude_recent);
    tree_counter = tree_counter + 1;
    // exact num_candidates-NN over the snapshot; squared L2 in float in nanoflann's metric_L2 order (L2_Adaptor::evalMetric,
    // nanoflann.hpp: groups of four differences, result += d0*d0 + d1*d1 + d2*d2 + d3*d3, then the remainder one by one)
    std::vector<std::pair<float, int>> best;
    for (int i = 0; i < (int)tree_keys.size(); ++i) best.emplace_back(sc_key_dist(cur.data(), tree_keys[i].data(), P.num_ring), i);
    std::sort(best.begin(), best.end());
    std::vector<size_t> cand(P.num_candidates, 0);  // slots the search does not fill stay 0 (:247)
    for (int c = 0; c < P.num_candidates && c < (int)best.size(); ++c) cand[c] = best[c].second;
    double min_dist = 10000000;
    int nn_align = 0, nn_idx = 0;
    for (int c = 0; c < P.num_candidates; ++c) {
      auto r = sc_distance(P, cur_desc.data(), descs[cand[c]].data());
      if (r.first < min_dist) { min_dist = r.first; nn_align = r.second; nn_idx = (int)cand[c]; }
    }
    int loop_id = min_dist < P.dist_thres ? nn_idx : -1;
    if (min_dist_out) *min_dist_out = min_dist;
    if (nn_idx_out) *nn_idx_out = nn_idx;
    const double unit_sector_angle = 360.0 / double(P.num_sector);
    float yaw = (float)(nn_align * unit_sector_angle * M_PI / 180.0);  // deg2rad (common.h:49-52)
    return {loop_id, yaw};
  }
};

}  // namespace orc
