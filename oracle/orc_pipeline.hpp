// TEST INFRASTRUCTURE — CPU oracle for the lidar-odometry hot path.  NOT product code.
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use anything under oracle/.
//
// PARITY UNPINNED: the reference ships no tests / golden vectors and cannot be
// built here.  First-party logic is restated line by line from
//   featureExtraction.hpp (FE), EstimationMapping.hpp (EM), lidarFactor.hpp (LF), common.h (CM)
// under /root/reference/src/visual_inertial_lidar/feature_tracker/include/.  Third-party arithmetic that
// is NOT in the reference tree is restated from the published algorithms of the versions the
// reference's README.md:22-30 names: PCL 1.7.2 (VoxelGrid, CropBox, KdTreeFLANN -> FLANN
// KDTreeSingleIndex), Eigen 3.3.7, Ceres 2.0.0 (TrustRegionMinimizer + LevenbergMarquardtStrategy +
// DenseQRSolver + HuberLoss/Corrector).  The kd-tree is the one piece that IS pinned: it is validated
// index-for-index against the reference's vendored nanoflann 1.3.2 (oracle/_ref, see Makefile).
#pragma once
#include <cstdint>
#include <cstdio>
#include <chrono>
#include <limits>
#include <vector>
#include "orc_math.hpp"

namespace orc {

struct P4 { float x, y, z, i; };
typedef std::vector<P4> Cloud;

struct Config {
  int n_scan = 64;              // FE:45  (/N_SCAN); 0 => explicit ring ids, n_rings rings
  int n_rings = 64;
  double lidar_min = 3.0;       // FE:46, config/kitti/velodyne_param_64.yaml:12
  double lidar_max = 90.0;      // FE:47, yaml:13
  double edge_threshold = 0.1;  // FE:48
  double edge_leaf = 0.4;       // EM:82, yaml:21
  double surf_leaf = 0.8;       // EM:83, yaml:22
  double crop_half = 100.0;     // EM:327-332
  double knn_gate = 1.0;        // EM:129, :189
  double huber = 0.1;           // EM:263
  int outer_iters = 2;          // EM:260
  int lm_max_iters = 4;         // EM:277
  int voxel_order = 0;          // 0: within-voxel order = input order (stable); 1: std::sort like PCL (unstable)
  // Alternates of the third-party restatements, for the sensitivity envelope only (tools/sensitivity.py); all 0 = the oracle proper.
  int eig_alg = 0;              // EM:150: 0 cyclic Jacobi, 1 Eigen 3.3.7's tridiagonal QR iteration (orc_math.hpp: eig3_sym_eigen)
  int plane_alg = 0;            // EM:198: 0 column-pivoted Householder QR, 1 Householder QR without column pivoting (Eigen's HouseholderQR)
  int centroid_div = 0;         // voxel centroid: 0 `sum / count` (Eigen >= 3.3 operator/=), 1 `sum * (1 / count)` (Eigen 3.2 operator/=)
  int lm_solver = 0;            // EM:283: 0 Householder QR of [J; D] (Ceres DENSE_QR), 1 normal equations + Cholesky (what the CUDA path does)
  int knn_ties = 0;             // tie class T2 (equal fp32 squared distances): 0 = FLANN / nanoflann order (first visited wins,
                                // the reference's behaviour); 1 = canonical (d^2, map index) order, the CUDA path's rule
};

// ------------------------------------------------------------------------------------------------
// Stage 1 — featureExtraction (FE:54-232)
// ------------------------------------------------------------------------------------------------
struct Smooth { double value; size_t ind; };

// FE:54-110.  Returns ring id or -1 (dropped).  `ring` is used when cfg.n_scan == 0.
inline int ring_of(const Config& c, const P4& p, int explicit_ring) {
  // Non-finite returns: FE:56-57 only fills an index vector, so they stay in the cloud; for N_SCANS 16/32/64 the x86 build
  // drops them anyway (comparisons with NaN are false, int(NaN) == INT_MIN fails FE:78 / :86 / :98; +-inf fails the range or
  // the scanID test).  The explicit-ring and "wrong scan number" modes would push NaN points into a ring and hand NaN to
  // std::sort's comparator (undefined behaviour); oracle and CUDA path both drop them in every mode instead.
  if (!(std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z))) return -1;
  float dxy = std::sqrt(p.x * p.x + p.y * p.y);  // CM:59-62 (float sqrt)
  double distance = dxy;
  if (distance < c.lidar_min || distance > c.lidar_max) return -1;  // FE:70
  if (c.n_scan == 0) return (explicit_ring >= 0 && explicit_ring < c.n_rings) ? explicit_ring : -1;
  double angle = std::atan(p.z / distance) * 180 / M_PI;  // FE:73
  int id = 0;
  if (c.n_scan == 16) {
    id = int((angle + 15) / 2 + 0.5);  // FE:77
    if (id > 15 || id < 0) return -1;
  } else if (c.n_scan == 32) {
    id = int((angle + 92.0 / 3.0) * 3.0 / 4.0);  // FE:85
    if (id > 31 || id < 0) return -1;
  } else if (c.n_scan == 64) {
    if (angle >= -8.83) id = int((2 - angle) * 3.0 + 0.5);  // FE:93-96
    else id = 64 / 2 + int((-8.83 - angle) * 2.0 + 0.5);
    if (angle > 2 || angle < -24.33 || id > 63 || id < 0) return -1;  // FE:98
  } else {
    id = 0;  // FE:103-106 "wrong scan number": everything lands in ring 0
  }
  return id;
}

// FE:112-173.  ring = the ring cloud, src = index of every ring point in the input scan.
inline void extract_sector(const Config& c, const Cloud& ring, const std::vector<int>& src, std::vector<Smooth>& sub,
                           Cloud& edge, std::vector<int>& edge_src, Cloud& surf, std::vector<int>& surf_src) {
  // FE:115 std::sort is unstable; equal curvatures (tie class T1) are ordered by index here.
  std::sort(sub.begin(), sub.end(), [](const Smooth& a, const Smooth& b) { return a.value < b.value || (a.value == b.value && a.ind < b.ind); });
  int picked_num = 0;
  std::vector<int> picked;
  for (int i = (int)sub.size() - 1; i >= 0; i--) {
    int ind = (int)sub[i].ind;
    if (std::find(picked.begin(), picked.end(), ind) == picked.end()) {
      if (sub[i].value <= c.edge_threshold) break;  // FE:125
      picked_num++;
      picked.push_back(ind);
      if (picked_num <= 20) { edge.push_back(ring[ind]); edge_src.push_back(src[ind]); }  // FE:131-133
      else break;
      for (int k = 1; k <= 5; k++) {  // FE:138-148
        double dx = ring[ind + k].x - ring[ind + k - 1].x;
        double dy = ring[ind + k].y - ring[ind + k - 1].y;
        double dz = ring[ind + k].z - ring[ind + k - 1].z;
        if (dx * dx + dy * dy + dz * dz > 0.05) break;
        picked.push_back(ind + k);
      }
      for (int l = -1; l >= -5; l--) {  // FE:150-160
        double dx = ring[ind + l].x - ring[ind + l + 1].x;
        double dy = ring[ind + l].y - ring[ind + l + 1].y;
        double dz = ring[ind + l].z - ring[ind + l + 1].z;
        if (dx * dx + dy * dy + dz * dz > 0.05) break;
        picked.push_back(ind + l);
      }
    }
  }
  for (int i = 0; i <= (int)sub.size() - 1; i++) {  // FE:165-172
    int ind = (int)sub[i].ind;
    if (std::find(picked.begin(), picked.end(), ind) == picked.end()) { surf.push_back(ring[ind]); surf_src.push_back(src[ind]); }
  }
}

// FE:223-232 extractFeature = getLaserCloud + featureEdge_Surf.  Outputs are APPENDED (FE:133, :170).
inline void extract_features(const Config& c, const P4* in, int n, const uint16_t* ring_ids, Cloud& edge, std::vector<int>& edge_src,
                             Cloud& surf, std::vector<int>& surf_src) {
  int R = c.n_scan == 0 ? c.n_rings : c.n_scan;
  if (c.n_scan != 0 && c.n_scan != 16 && c.n_scan != 32 && c.n_scan != 64) R = std::max(1, c.n_scan);
  std::vector<Cloud> rings(R);
  std::vector<std::vector<int>> srcs(R);
  for (int i = 0; i < n; ++i) {
    int id = ring_of(c, in[i], ring_ids ? (int)ring_ids[i] : -1);
    if (id < 0) continue;
    rings[id].push_back(in[i]);  // FE:108 arrival order
    srcs[id].push_back(i);
  }
  std::vector<Smooth> curv;
  for (int r = 0; r < R; ++r) {  // FE:175-220
    const Cloud& rc = rings[r];
    if (rc.size() < 131) continue;  // FE:179
    curv.clear();
    size_t smooth_size = rc.size() - 5;
    for (size_t j = 5; j < smooth_size; j++) {
      // FE:190-198: every operand is float (10 * float stays float) -> fp32 left-to-right, then widened.
      float fx = rc[j - 5].x + rc[j - 4].x + rc[j - 3].x + rc[j - 2].x + rc[j - 1].x - 10 * rc[j].x + rc[j + 1].x + rc[j + 2].x + rc[j + 3].x + rc[j + 4].x + rc[j + 5].x;
      float fy = rc[j - 5].y + rc[j - 4].y + rc[j - 3].y + rc[j - 2].y + rc[j - 1].y - 10 * rc[j].y + rc[j + 1].y + rc[j + 2].y + rc[j + 3].y + rc[j + 4].y + rc[j + 5].y;
      float fz = rc[j - 5].z + rc[j - 4].z + rc[j - 3].z + rc[j - 2].z + rc[j - 1].z - 10 * rc[j].z + rc[j + 1].z + rc[j + 2].z + rc[j + 3].z + rc[j + 4].z + rc[j + 5].z;
      double dx = fx, dy = fy, dz = fz;
      curv.push_back({dx * dx + dy * dy + dz * dz, j});  // FE:200
    }
    size_t cloud_size = smooth_size - 5;  // FE:205
    for (int j = 0; j < 6; j++) {
      int len = (int)(cloud_size / 6);
      int start = len * j, end = len * (j + 1) - 1;
      if (j == 5) end = (int)cloud_size - 1;
      std::vector<Smooth> sub(curv.begin() + start, curv.begin() + end);  // FE:215 half-open: element `end` is dropped
      extract_sector(c, rc, srcs[r], sub, edge, edge_src, surf, surf_src);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// pcl::VoxelGrid<PointXYZI>::applyFilter (PCL 1.7.2 filters/impl/voxel_grid.hpp), call sites EM:248-251, :347-350
// downsample_all_data_ = true (default): centroid of x,y,z,intensity accumulated in fp32, divided by float(count).
// ------------------------------------------------------------------------------------------------
struct VoxKey { unsigned int idx; unsigned int pt; };

inline bool voxel_grid(const Cloud& in, float leaf, int order_mode, Cloud& out, std::vector<int>* first_pt = nullptr) {
  out.clear();
  if (first_pt) first_pt->clear();
  if (in.empty()) return true;
  const float inv = 1.0f / leaf;  // inverse_leaf_size_ = Array4f::Ones() / leaf_size_.array()
  float mn[3] = {in[0].x, in[0].y, in[0].z}, mx[3] = {in[0].x, in[0].y, in[0].z};
  for (const P4& p : in) {  // getMinMax3D
    mn[0] = std::min(mn[0], p.x); mx[0] = std::max(mx[0], p.x);
    mn[1] = std::min(mn[1], p.y); mx[1] = std::max(mx[1], p.y);
    mn[2] = std::min(mn[2], p.z); mx[2] = std::max(mx[2], p.z);
  }
  int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1, dy = (int64_t)((mx[1] - mn[1]) * inv) + 1, dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > (int64_t)std::numeric_limits<int32_t>::max()) {  // "Leaf size is too small": output = input
    out = in;
    return false;
  }
  int min_b[3], max_b[3], div_b[3];
  for (int a = 0; a < 3; ++a) {
    min_b[a] = (int)std::floor(mn[a] * inv);
    max_b[a] = (int)std::floor(mx[a] * inv);
    div_b[a] = max_b[a] - min_b[a] + 1;
  }
  const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
  std::vector<VoxKey> keys;
  keys.reserve(in.size());
  for (size_t i = 0; i < in.size(); ++i) {
    int i0 = (int)(std::floor(in[i].x * inv) - (float)min_b[0]);
    int i1 = (int)(std::floor(in[i].y * inv) - (float)min_b[1]);
    int i2 = (int)(std::floor(in[i].z * inv) - (float)min_b[2]);
    keys.push_back({(unsigned)(i0 * mul[0] + i1 * mul[1] + i2 * mul[2]), (unsigned)i});
  }
  if (order_mode & 1)  // what PCL does: std::sort on idx only; within-voxel order unspecified (tie class T3)
    std::sort(keys.begin(), keys.end(), [](const VoxKey& a, const VoxKey& b) { return a.idx < b.idx; });
  else  // canonical: input order inside a voxel
    std::stable_sort(keys.begin(), keys.end(), [](const VoxKey& a, const VoxKey& b) { return a.idx < b.idx; });
  size_t index = 0;
  while (index < keys.size()) {
    size_t i = index + 1;
    while (i < keys.size() && keys[i].idx == keys[index].idx) ++i;
    float c[4] = {0, 0, 0, 0};
    for (size_t li = index; li < i; ++li) {
      const P4& p = in[keys[li].pt];
      c[0] += p.x; c[1] += p.y; c[2] += p.z; c[3] += p.i;
    }
    const float cnt = (float)(i - index);
    if (order_mode & 2) { const float rc = 1.0f / cnt; out.push_back({c[0] * rc, c[1] * rc, c[2] * rc, c[3] * rc}); }  // sensitivity alternate
    else out.push_back({c[0] / cnt, c[1] / cnt, c[2] / cnt, c[3] / cnt});
    if (first_pt) first_pt->push_back((int)keys[index].pt);
    index = i;
  }
  return true;
}

// pcl::CropBox<PointXYZI>::applyFilter, call site EM:335-344: closed AABB, order preserving, bounds cast to float.
inline void crop_box(const Cloud& in, const double mn[3], const double mx[3], Cloud& out) {
  const float lo[3] = {(float)mn[0], (float)mn[1], (float)mn[2]}, hi[3] = {(float)mx[0], (float)mx[1], (float)mx[2]};
  out.clear();
  for (const P4& p : in) {
    if (p.x < lo[0] || p.y < lo[1] || p.z < lo[2] || p.x > hi[0] || p.y > hi[1] || p.z > hi[2]) continue;
    out.push_back(p);
  }
}

// ------------------------------------------------------------------------------------------------
// pcl::KdTreeFLANN -> FLANN KDTreeSingleIndex (leaf 15, L2_Simple<float>, eps 0, sorted), EM:256-257, :128, :185.
// Restated from the algorithm as vendored in the reference tree (nanoflann 1.3.2, a FLANN single-index
// derivative: src/global_fusion/include/Scancontext/nanoflann.hpp:858-1003 build, :1347-1410 search,
// :175-199 result set).  Validated index-for-index against that file (tests/test_oracle_kdtree.py).
// ------------------------------------------------------------------------------------------------
class KdTree {
 public:
  struct Node { int child1, child2; int left, right; int divfeat; float divlow, divhigh; };
  struct Interval { float low, high; };

  void set_canonical_ties(bool on) { canonical_ = on; }
  void build(const Cloud* cloud, int leaf_max = 15) {
    pts_ = cloud;
    leaf_max_ = leaf_max;
    nodes_.clear();
    const int n = (int)cloud->size();
    vind_.resize(n);
    for (int i = 0; i < n; ++i) vind_[i] = i;
    if (n == 0) { root_ = -1; return; }
    for (int a = 0; a < 3; ++a) root_bbox_[a].low = root_bbox_[a].high = get(0, a);
    for (int k = 1; k < n; ++k)
      for (int a = 0; a < 3; ++a) {
        if (get(k, a) < root_bbox_[a].low) root_bbox_[a].low = get(k, a);
        if (get(k, a) > root_bbox_[a].high) root_bbox_[a].high = get(k, a);
      }
    Interval bb[3] = {root_bbox_[0], root_bbox_[1], root_bbox_[2]};
    root_ = divide(0, n, bb);
    // nanoflann keeps root_bbox from computeBoundingBox (not the tightened one); identical for the root.
  }

  // exact k-NN: ascending squared distances, strict '<' insertion (first visited wins ties, tie class T2)
  void knn(const float q[3], int k, int* idx, float* d2) const {
    for (int i = 0; i < k; ++i) { idx[i] = -1; d2[i] = std::numeric_limits<float>::max(); }
    if (root_ < 0) return;
    int count = 0;
    float dists[3] = {0, 0, 0};
    float distsq = 0;
    for (int a = 0; a < 3; ++a) {
      if (q[a] < root_bbox_[a].low) { dists[a] = (q[a] - root_bbox_[a].low) * (q[a] - root_bbox_[a].low); distsq += dists[a]; }
      if (q[a] > root_bbox_[a].high) { dists[a] = (q[a] - root_bbox_[a].high) * (q[a] - root_bbox_[a].high); distsq += dists[a]; }
    }
    search(q, root_, distsq, dists, k, idx, d2, count);
  }
  size_t node_count() const { return nodes_.size(); }

 private:
  float get(int i, int a) const { const P4& p = (*pts_)[i]; return a == 0 ? p.x : (a == 1 ? p.y : p.z); }

  int divide(int left, int right, Interval* bbox) {
    int id = (int)nodes_.size();
    nodes_.push_back(Node());
    if (right - left <= leaf_max_) {
      Node nd; nd.child1 = nd.child2 = -1; nd.left = left; nd.right = right; nd.divfeat = 0; nd.divlow = nd.divhigh = 0;
      for (int a = 0; a < 3; ++a) bbox[a].low = bbox[a].high = get(vind_[left], a);
      for (int k = left + 1; k < right; ++k)
        for (int a = 0; a < 3; ++a) {
          if (bbox[a].low > get(vind_[k], a)) bbox[a].low = get(vind_[k], a);
          if (bbox[a].high < get(vind_[k], a)) bbox[a].high = get(vind_[k], a);
        }
      nodes_[id] = nd;
    } else {
      int idx, cutfeat; float cutval;
      middle_split(&vind_[left], right - left, idx, cutfeat, cutval, bbox);
      Interval lb[3] = {bbox[0], bbox[1], bbox[2]};
      lb[cutfeat].high = cutval;
      int c1 = divide(left, left + idx, lb);
      Interval rb[3] = {bbox[0], bbox[1], bbox[2]};
      rb[cutfeat].low = cutval;
      int c2 = divide(left + idx, right, rb);
      Node nd; nd.child1 = c1; nd.child2 = c2; nd.left = nd.right = 0; nd.divfeat = cutfeat;
      nd.divlow = lb[cutfeat].high; nd.divhigh = rb[cutfeat].low;
      nodes_[id] = nd;
      for (int a = 0; a < 3; ++a) { bbox[a].low = std::min(lb[a].low, rb[a].low); bbox[a].high = std::max(lb[a].high, rb[a].high); }
    }
    return id;
  }
  void min_max(const int* ind, int count, int a, float& mn, float& mx) const {
    mn = mx = get(ind[0], a);
    for (int i = 1; i < count; ++i) { float v = get(ind[i], a); if (v < mn) mn = v; if (v > mx) mx = v; }
  }
  void middle_split(int* ind, int count, int& index, int& cutfeat, float& cutval, const Interval* bbox) {
    const float EPS = 0.00001f;
    float max_span = bbox[0].high - bbox[0].low;
    for (int a = 1; a < 3; ++a) { float span = bbox[a].high - bbox[a].low; if (span > max_span) max_span = span; }
    float max_spread = -1;
    cutfeat = 0;
    for (int a = 0; a < 3; ++a) {
      float span = bbox[a].high - bbox[a].low;
      if (span > (1 - EPS) * max_span) {
        float mn, mx; min_max(ind, count, a, mn, mx);
        float spread = mx - mn;
        if (spread > max_spread) { cutfeat = a; max_spread = spread; }
      }
    }
    float split_val = (bbox[cutfeat].low + bbox[cutfeat].high) / 2;
    float mn, mx; min_max(ind, count, cutfeat, mn, mx);
    if (split_val < mn) cutval = mn; else if (split_val > mx) cutval = mx; else cutval = split_val;
    int lim1, lim2;
    plane_split(ind, count, cutfeat, cutval, lim1, lim2);
    if (lim1 > count / 2) index = lim1; else if (lim2 < count / 2) index = lim2; else index = count / 2;
  }
  void plane_split(int* ind, int count, int cutfeat, float cutval, int& lim1, int& lim2) {
    int left = 0, right = count - 1;
    for (;;) {
      while (left <= right && get(ind[left], cutfeat) < cutval) ++left;
      while (right && left <= right && get(ind[right], cutfeat) >= cutval) --right;
      if (left > right || !right) break;
      std::swap(ind[left], ind[right]); ++left; --right;
    }
    lim1 = left;
    right = count - 1;
    for (;;) {
      while (left <= right && get(ind[left], cutfeat) <= cutval) ++left;
      while (right && left <= right && get(ind[right], cutfeat) > cutval) --right;
      if (left > right || !right) break;
      std::swap(ind[left], ind[right]); ++left; --right;
    }
    lim2 = left;
  }
  void add_point(float dist, int index, int k, int* idx, float* d2, int& count) const {
    int i;
    for (i = count; i > 0; --i) {
      if (d2[i - 1] > dist || (canonical_ && d2[i - 1] == dist && idx[i - 1] > index)) { if (i < k) { d2[i] = d2[i - 1]; idx[i] = idx[i - 1]; } }
      else break;
    }
    if (i < k) { d2[i] = dist; idx[i] = index; }
    if (count < k) count++;
  }
  void search(const float q[3], int node, float mindistsq, float* dists, int k, int* idx, float* d2, int& count) const {
    const Node& nd = nodes_[node];
    if (nd.child1 < 0 && nd.child2 < 0) {
      float worst = d2[k - 1];
      for (int i = nd.left; i < nd.right; ++i) {
        const int index = vind_[i];
        const P4& p = (*pts_)[index];
        float dist = 0;  // L2_Simple: result += diff*diff, x then y then z
        { float d = q[0] - p.x; dist += d * d; }
        { float d = q[1] - p.y; dist += d * d; }
        { float d = q[2] - p.z; dist += d * d; }
        // canonical ties: an equal distance with a lower index displaces the current k-th (the subtree pruning test below
        // already visits every branch whose lower bound EQUALS the k-th distance, so no tied candidate is missed)
        if (dist < worst || (canonical_ && count == k && dist == d2[k - 1] && index < idx[k - 1])) add_point(dist, index, k, idx, d2, count);
      }
      return;
    }
    int a = nd.divfeat;
    float val = q[a];
    float diff1 = val - nd.divlow, diff2 = val - nd.divhigh;
    int best, other; float cut;
    if ((diff1 + diff2) < 0) { best = nd.child1; other = nd.child2; cut = (val - nd.divhigh) * (val - nd.divhigh); }
    else { best = nd.child2; other = nd.child1; cut = (val - nd.divlow) * (val - nd.divlow); }
    search(q, best, mindistsq, dists, k, idx, d2, count);
    float dst = dists[a];
    mindistsq = mindistsq + cut - dst;
    dists[a] = cut;
    if (mindistsq * 1.0f <= d2[k - 1]) search(q, other, mindistsq, dists, k, idx, d2, count);
    dists[a] = dst;
  }

  const Cloud* pts_ = nullptr;
  bool canonical_ = false;
  int leaf_max_ = 15;
  int root_ = -1;
  std::vector<int> vind_;
  std::vector<Node> nodes_;
  Interval root_bbox_[3];
};

// ------------------------------------------------------------------------------------------------
// Factors (EM:117-232) and their residual / Jacobian (LF:21-52, :79-102)
// ------------------------------------------------------------------------------------------------
struct EdgeFactor { V3 p, a, b; };
struct SurfFactor { V3 p, n; double d; };

// EM:355-363
inline P4 associate(Quat q, V3 t, const P4& in) {
  V3 w = qrot(q, V3{in.x, in.y, in.z}) + t;
  return {(float)w.x, (float)w.y, (float)w.z, in.i};
}

// EM:117-172.  nn_idx/nn_d2 (5 per point) and valid (1 per point) are optional diagnostics.
inline void edge_factors(const Config& c, Quat q, V3 t, const Cloud& edge, const Cloud& map, const KdTree& tree, std::vector<EdgeFactor>& out,
                         int* nn_idx = nullptr, float* nn_d2 = nullptr, uint8_t* valid = nullptr) {
  for (size_t i = 0; i < edge.size(); ++i) {
    P4 w = associate(q, t, edge[i]);
    int idx[5]; float d2[5];
    const float qq[3] = {w.x, w.y, w.z};
    tree.knn(qq, 5, idx, d2);
    if (nn_idx) { std::memcpy(nn_idx + 5 * i, idx, sizeof(idx)); std::memcpy(nn_d2 + 5 * i, d2, sizeof(d2)); }
    if (valid) valid[i] = 0;
    if (d2[4] < c.knn_gate) {
      V3 near[5], center{0, 0, 0};
      for (int j = 0; j < 5; ++j) { near[j] = {map[idx[j]].x, map[idx[j]].y, map[idx[j]].z}; center = center + near[j]; }
      center = {center.x / 5.0, center.y / 5.0, center.z / 5.0};
      M3 cov{};  // EM:143-148 (not divided by 5)
      for (int j = 0; j < 5; ++j) {
        V3 z = near[j] - center;
        const double v[3] = {z.x, z.y, z.z};
        for (int r = 0; r < 3; ++r)
          for (int cc = 0; cc < 3; ++cc) cov.m[r][cc] = cov.m[r][cc] + v[r] * v[cc];
      }
      double w3[3]; M3 V;
      if (c.eig_alg == 1) eig3_sym_eigen(cov, w3, V); else eig3_sym(cov, w3, V);
      V3 dir{V.m[0][2], V.m[1][2], V.m[2][2]};  // EM:151
      if (w3[2] > 3 * w3[1]) {                    // EM:153
        EdgeFactor f;
        f.p = {edge[i].x, edge[i].y, edge[i].z};
        f.a = 0.1 * dir + center;   // EM:156
        f.b = -0.1 * dir + center;  // EM:157
        out.push_back(f);
        if (valid) valid[i] = 1;
      }
    }
  }
}

// EM:174-232
inline void surf_factors(const Config& c, Quat q, V3 t, const Cloud& surf, const Cloud& map, const KdTree& tree, std::vector<SurfFactor>& out,
                         int* nn_idx = nullptr, float* nn_d2 = nullptr, uint8_t* valid = nullptr) {
  for (size_t i = 0; i < surf.size(); ++i) {
    P4 w = associate(q, t, surf[i]);
    int idx[5]; float d2[5];
    const float qq[3] = {w.x, w.y, w.z};
    tree.knn(qq, 5, idx, d2);
    if (nn_idx) { std::memcpy(nn_idx + 5 * i, idx, sizeof(idx)); std::memcpy(nn_d2 + 5 * i, d2, sizeof(d2)); }
    if (valid) valid[i] = 0;
    if (d2[4] < c.knn_gate) {
      double A[5][3], b[5];
      for (int j = 0; j < 5; ++j) { A[j][0] = map[idx[j]].x; A[j][1] = map[idx[j]].y; A[j][2] = map[idx[j]].z; b[j] = -1.0; }
      V3 n = lstsq5x3_colpiv(A, b, c.plane_alg != 1);  // EM:198
      double nn = norm(n);
      double d = 1.0 / nn;           // EM:199
      n = {n.x / nn, n.y / nn, n.z / nn};  // EM:200 normalize(): v /= norm
      bool ok = true;
      for (int j = 0; j < 5; ++j)
        if (std::fabs(n.x * (double)map[idx[j]].x + n.y * (double)map[idx[j]].y + n.z * (double)map[idx[j]].z + d) > 0.2) { ok = false; break; }
      if (ok) {
        out.push_back({V3{surf[i].x, surf[i].y, surf[i].z}, n, d});
        if (valid) valid[i] = 1;
      }
    }
  }
}

// LF:21-52: residual (3) and 3x6 local Jacobian (row-major) of one edge factor at pose (q,t).
inline void edge_eval(const EdgeFactor& f, Quat q, V3 t, double r[3], double* J /*18 or null*/) {
  V3 lp = qrot(q, f.p) + t;
  V3 nu = cross(lp - f.a, lp - f.b);
  V3 ab = f.a - f.b;
  double abn = norm(ab);
  r[0] = nu.x / abn; r[1] = nu.y / abn; r[2] = nu.z / abn;
  if (J) {
    M3 sl = skew(lp), sab = skew(ab);
    double Jse3[3][6];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) { Jse3[i][j] = -sl.m[i][j]; Jse3[i][3 + j] = (i == j) ? 1.0 : 0.0; }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 6; ++j) {
        double s = (-sab.m[i][0]) * Jse3[0][j] + (-sab.m[i][1]) * Jse3[1][j] + (-sab.m[i][2]) * Jse3[2][j];
        J[i * 6 + j] = s / abn;
      }
  }
}
// LF:79-102
inline void surf_eval(const SurfFactor& f, Quat q, V3 t, double r[1], double* J /*6 or null*/) {
  V3 pw = qrot(q, f.p) + t;
  r[0] = dot(f.n, pw) + f.d;
  if (J) {
    M3 sp = skew(pw);
    double Jse3[3][6];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) { Jse3[i][j] = -sp.m[i][j]; Jse3[i][3 + j] = (i == j) ? 1.0 : 0.0; }
    for (int j = 0; j < 6; ++j) J[j] = f.n.x * Jse3[0][j] + f.n.y * Jse3[1][j] + f.n.z * Jse3[2][j];
  }
}

// ------------------------------------------------------------------------------------------------
// ceres::Solve restated (Ceres 2.0.0; EM:263-283).  One 7-parameter block with LocalSE3Parameterization
// (EM:20-69: Plus = left se(3) perturbation, ComputeJacobian = [I6;0] so the local Jacobian is the first
// six columns), HuberLoss(0.1) on every block, TRUST_REGION / LEVENBERG_MARQUARDT / DENSE_QR,
// max_num_iterations 4, all other options default (function_tolerance 1e-6, gradient_tolerance 1e-10,
// parameter_tolerance 1e-8, initial radius 1e4, max radius 1e16, min radius 1e-32, min_relative_decrease
// 1e-3, min/max LM diagonal 1e-6/1e32, jacobi_scaling on, monotonic steps).
// ------------------------------------------------------------------------------------------------
struct LmIter {            // one row of the per-iteration trace
  int iteration;
  int step_valid, step_successful;
  double cost, candidate_cost, model_cost_change, relative_decrease, radius, step_norm, gradient_max_norm;
  double x[7];
};
struct SolveTrace {
  int n_edge = 0, n_surf = 0;
  int termination = 0;  // 0 max iterations, 1 parameter tol, 2 function tol, 3 gradient tol, 4 no residuals, 5 other
  std::vector<LmIter> iters;
  double H0[21], g0[6], cost0;  // unscaled normal equations at iteration 0 (upper triangle row-major)
};

struct Problem {
  const std::vector<EdgeFactor>* edges;
  const std::vector<SurfFactor>* surfs;
  double huber;
  int rows() const { return 3 * (int)edges->size() + (int)surfs->size(); }

  // ResidualBlock::Evaluate + Corrector (rho'' <= 0 for Huber => scale r and J by sqrt(rho'), alpha = 0)
  // Jm: column-major (rows x 6) with leading dimension ld, or null (cost only).
  double evaluate(const double x[7], double* res, double* Jm, int ld, double* grad) const {
    Quat q{x[0], x[1], x[2], x[3]};
    V3 t{x[4], x[5], x[6]};
    const double a = huber, b = huber * huber;
    double cost = 0;
    if (grad) for (int j = 0; j < 6; ++j) grad[j] = 0;
    int row = 0;
    for (const EdgeFactor& f : *edges) {
      double r[3], J[18];
      edge_eval(f, q, t, r, Jm ? J : nullptr);
      double s = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
      double rho0, rho1;
      if (s > b) { double rr = std::sqrt(s); rho0 = 2 * a * rr - b; rho1 = std::max(std::numeric_limits<double>::min(), a / rr); }
      else { rho0 = s; rho1 = 1.0; }
      cost += 0.5 * rho0;
      if (res || Jm) {
        double sq = std::sqrt(rho1);
        for (int i = 0; i < 3; ++i) {
          if (Jm) for (int j = 0; j < 6; ++j) { J[i * 6 + j] *= sq; Jm[(size_t)j * ld + row + i] = J[i * 6 + j]; }
          r[i] *= sq;
          if (res) res[row + i] = r[i];
        }
        if (grad && Jm) for (int j = 0; j < 6; ++j) grad[j] += J[j] * r[0] + J[6 + j] * r[1] + J[12 + j] * r[2];
      }
      row += 3;
    }
    for (const SurfFactor& f : *surfs) {
      double r[1], J[6];
      surf_eval(f, q, t, r, Jm ? J : nullptr);
      double s = r[0] * r[0];
      double rho0, rho1;
      if (s > b) { double rr = std::sqrt(s); rho0 = 2 * a * rr - b; rho1 = std::max(std::numeric_limits<double>::min(), a / rr); }
      else { rho0 = s; rho1 = 1.0; }
      cost += 0.5 * rho0;
      if (res || Jm) {
        double sq = std::sqrt(rho1);
        if (Jm) for (int j = 0; j < 6; ++j) { J[j] *= sq; Jm[(size_t)j * ld + row] = J[j]; }
        r[0] *= sq;
        if (res) res[row] = r[0];
        if (grad && Jm) for (int j = 0; j < 6; ++j) grad[j] += J[j] * r[0];
      }
      row += 1;
    }
    return cost;
  }
};

// Eigen::HouseholderQR<ColMajor>::compute + solve as used by ceres DenseQRSolver: least squares of the
// (m x 6) column-major matrix A (destroyed) against rhs (destroyed, length m); x[6] out.
inline void dense_qr_solve(double* A, int m, int ld, double* rhs, double x[6]) {
  const int n = 6;
  double tau[6];
  for (int k = 0; k < n && k < m; ++k) {
    double* col = A + (size_t)k * ld;
    double tail = 0;
    for (int i = k + 1; i < m; ++i) tail += col[i] * col[i];
    double c0 = col[k], beta;
    if (tail <= std::numeric_limits<double>::min()) { tau[k] = 0; beta = c0; for (int i = k + 1; i < m; ++i) col[i] = 0; }
    else {
      beta = std::sqrt(c0 * c0 + tail);
      if (c0 >= 0) beta = -beta;
      const double den = c0 - beta;
      for (int i = k + 1; i < m; ++i) col[i] /= den;
      tau[k] = (beta - c0) / beta;
    }
    col[k] = beta;
    for (int j = k + 1; j < n; ++j) {
      double* cj = A + (size_t)j * ld;
      double tmp = cj[k];
      for (int i = k + 1; i < m; ++i) tmp += col[i] * cj[i];
      cj[k] -= tau[k] * tmp;
      for (int i = k + 1; i < m; ++i) cj[i] -= tau[k] * col[i] * tmp;
    }
    double tmp = rhs[k];
    for (int i = k + 1; i < m; ++i) tmp += col[i] * rhs[i];
    rhs[k] -= tau[k] * tmp;
    for (int i = k + 1; i < m; ++i) rhs[i] -= tau[k] * col[i] * tmp;
  }
  for (int k = n - 1; k >= 0; --k) {
    double s = rhs[k];
    for (int j = k + 1; j < n; ++j) s -= A[(size_t)j * ld + k] * x[j];
    x[k] = s / A[(size_t)k * ld + k];
  }
}

// Sensitivity alternate of the line above (Config.lm_solver = 1): the same minimiser from (J^T J + D^2) y = J^T r by Cholesky.
inline void normal_cholesky_solve(const double* J, int m, int ld, const double* r, const double D[6], double x[6]) {
  double H[6][6], g[6];
  for (int i = 0; i < 6; ++i) {
    double s = 0;
    for (int k = 0; k < m; ++k) s += J[(size_t)i * ld + k] * r[k];
    g[i] = s;
    for (int j = i; j < 6; ++j) { double t = 0; for (int k = 0; k < m; ++k) t += J[(size_t)i * ld + k] * J[(size_t)j * ld + k]; H[i][j] = H[j][i] = t; }
    H[i][i] += D[i] * D[i];
  }
  double L[6][6] = {};
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = H[i][j];
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
      L[i][j] = i == j ? std::sqrt(s) : s / L[j][j];
    }
  double y[6];
  for (int i = 0; i < 6; ++i) { double s = g[i]; for (int k = 0; k < i; ++k) s -= L[i][k] * y[k]; y[i] = s / L[i][i]; }
  for (int i = 5; i >= 0; --i) { double s = y[i]; for (int k = i + 1; k < 6; ++k) s -= L[k][i] * x[k]; x[i] = s / L[i][i]; }
}

inline double norm7(const double* v) { double s = 0; for (int i = 0; i < 7; ++i) s += v[i] * v[i]; return std::sqrt(s); }

// TrustRegionMinimizer::Minimize (ceres 2.0 internal/ceres/trust_region_minimizer.cc) on this problem.
inline void ceres_solve(const Problem& prob, double params[7], int max_iters, SolveTrace* trace, int lm_solver = 0) {
  const int m = prob.rows();
  if (trace) { trace->n_edge = (int)prob.edges->size(); trace->n_surf = (int)prob.surfs->size(); trace->iters.clear(); }
  if (m == 0) { if (trace) trace->termination = 4; return; }
  const int ld = m + 6;
  std::vector<double> Jm((size_t)ld * 6), Jwork((size_t)ld * 6), res(m), rhs(ld), model(m);
  double x[7], cand[7], grad[6], scale[6], diag[6], lm_diag[6], step[6], delta[6];
  std::memcpy(x, params, sizeof(x));
  double x_norm = norm7(x);
  double x_cost = 0, candidate_cost = 0, minimum_cost = std::numeric_limits<double>::max();
  double radius = 1e4, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  double model_cost_change = 0;
  double grad_max = 0;
  int num_invalid = 0;

  auto eval_grad_jac = [&](int iteration) {  // EvaluateGradientAndJacobian
    x_cost = prob.evaluate(x, res.data(), Jm.data(), ld, grad);
    if (iteration == 0) {
      if (trace) {
        int kk = 0;
        for (int i = 0; i < 6; ++i)
          for (int j = i; j < 6; ++j) { double s = 0; for (int r = 0; r < m; ++r) s += Jm[(size_t)i * ld + r] * Jm[(size_t)j * ld + r]; trace->H0[kk++] = s; }
        for (int j = 0; j < 6; ++j) trace->g0[j] = grad[j];
        trace->cost0 = x_cost;
      }
      for (int j = 0; j < 6; ++j) { double s = 0; for (int r = 0; r < m; ++r) s += Jm[(size_t)j * ld + r] * Jm[(size_t)j * ld + r]; scale[j] = 1.0 / (1.0 + std::sqrt(s)); }
    }
    for (int j = 0; j < 6; ++j) for (int r = 0; r < m; ++r) Jm[(size_t)j * ld + r] *= scale[j];
    double ng[6], proj[7];
    for (int j = 0; j < 6; ++j) ng[j] = -grad[j];
    se3_plus(x, ng, proj);
    grad_max = 0;
    for (int i = 0; i < 7; ++i) grad_max = std::max(grad_max, std::fabs(x[i] - proj[i]));
  };
  auto record = [&](int it, int valid, int succ, double cost, double rel, double stepn) {
    if (!trace) return;
    LmIter r; r.iteration = it; r.step_valid = valid; r.step_successful = succ; r.cost = cost; r.candidate_cost = candidate_cost;
    r.model_cost_change = model_cost_change; r.relative_decrease = rel; r.radius = radius; r.step_norm = stepn; r.gradient_max_norm = grad_max;
    std::memcpy(r.x, x, sizeof(x));
    trace->iters.push_back(r);
  };

  // IterationZero
  eval_grad_jac(0);
  bool step_successful = true;
  int iteration = 0;
  record(0, 1, 1, x_cost, 0, 0);
  int termination = 0;
  for (;;) {
    // FinalizeIterationAndCheckIfMinimizerCanContinue
    if (step_successful && x_cost < minimum_cost) { minimum_cost = x_cost; std::memcpy(params, x, sizeof(x)); }
    if (iteration >= max_iters) { termination = 0; break; }
    if (step_successful && grad_max <= 1e-10) { termination = 3; break; }
    if (radius <= 1e-32) { termination = 5; break; }
    ++iteration;
    // ComputeTrustRegionStep -> LevenbergMarquardtStrategy::ComputeStep
    if (!reuse_diagonal) {
      for (int j = 0; j < 6; ++j) { double s = 0; for (int r = 0; r < m; ++r) s += Jm[(size_t)j * ld + r] * Jm[(size_t)j * ld + r]; diag[j] = std::min(std::max(s, 1e-6), 1e32); }
    }
    for (int j = 0; j < 6; ++j) lm_diag[j] = std::sqrt(diag[j] / radius);
    // DenseQRSolver: [J; D] y = [r; 0]  (then step = -y)
    Jwork = Jm;
    for (int j = 0; j < 6; ++j) { for (int i = 0; i < 6; ++i) Jwork[(size_t)j * ld + m + i] = 0; Jwork[(size_t)j * ld + m + j] = lm_diag[j]; }
    for (int r = 0; r < m; ++r) rhs[r] = res[r];
    for (int i = 0; i < 6; ++i) rhs[m + i] = 0;
    if (lm_solver == 1) normal_cholesky_solve(Jm.data(), m, ld, res.data(), lm_diag, step);  // sensitivity alternate
    else dense_qr_solve(Jwork.data(), ld, ld, rhs.data(), step);
    reuse_diagonal = true;
    bool finite = true;
    for (int j = 0; j < 6; ++j) { step[j] = -step[j]; if (!std::isfinite(step[j])) finite = false; }
    bool step_valid = false;
    if (finite) {
      for (int r = 0; r < m; ++r) { double s = 0; for (int j = 0; j < 6; ++j) s += Jm[(size_t)j * ld + r] * step[j]; model[r] = s; }
      double mc = 0;
      for (int r = 0; r < m; ++r) mc += model[r] * (res[r] + model[r] / 2.0);
      model_cost_change = -mc;
      step_valid = model_cost_change > 0.0;
    }
    if (!step_valid) {  // HandleInvalidStep
      if (++num_invalid >= 5) { termination = 5; record(iteration, 0, 0, x_cost, 0, 0); break; }
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      step_successful = false;
      record(iteration, 0, 0, x_cost, 0, 0);
      continue;
    }
    num_invalid = 0;
    for (int j = 0; j < 6; ++j) delta[j] = step[j] * scale[j];
    // ComputeCandidatePointAndEvaluateCost
    se3_plus(x, delta, cand);
    candidate_cost = prob.evaluate(cand, nullptr, nullptr, 0, nullptr);
    // ParameterToleranceReached
    double sn = 0; for (int i = 0; i < 7; ++i) sn += (x[i] - cand[i]) * (x[i] - cand[i]);
    sn = std::sqrt(sn);
    if (sn <= 1e-8 * (x_norm + 1e-8)) { termination = 1; record(iteration, 1, 0, x_cost, 0, sn); break; }
    // FunctionToleranceReached
    double cost_change = x_cost - candidate_cost;
    if (std::fabs(cost_change) <= 1e-6 * x_cost) { termination = 2; record(iteration, 1, 0, x_cost, 0, sn); break; }
    // IsStepSuccessful (monotonic: StepQuality = cost_change / model_cost_change)
    double rel = cost_change / model_cost_change;
    if (rel > 1e-3) {  // HandleSuccessfulStep
      std::memcpy(x, cand, sizeof(x));
      x_norm = norm7(x);
      eval_grad_jac(iteration);
      step_successful = true;
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rel - 1.0, 3));
      radius = std::min(1e16, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
      record(iteration, 1, 1, x_cost, rel, sn);
    } else {  // HandleUnsuccessfulStep
      step_successful = false;
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      record(iteration, 1, 0, x_cost, rel, sn);
    }
  }
  if (trace) trace->termination = termination;
}

// ------------------------------------------------------------------------------------------------
// EstimationMapping (EM:71-403)
// ------------------------------------------------------------------------------------------------
struct Timing { double ds = 0, kdbuild = 0, assoc = 0, solve = 0, map = 0, extract = 0; int frames = 0; };
inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

class Odometry {
 public:
  explicit Odometry(const Config& c) : cfg(c) { odom = iso_identity(); odom_last = iso_identity(); }  // EM:80-92

  // EM:105-115
  void init_map(const Cloud& edge, const Cloud& surf) {
    map_edge.insert(map_edge.end(), edge.begin(), edge.end());
    map_surf.insert(map_surf.end(), surf.begin(), surf.end());
    registered.insert(registered.end(), edge.begin(), edge.end());
    registered.insert(registered.end(), surf.begin(), surf.end());
    no_registered.insert(no_registered.end(), edge.begin(), edge.end());
    no_registered.insert(no_registered.end(), surf.begin(), surf.end());
  }

  // EM:235-296
  void update(const Cloud& edge_in, const Cloud& surf_in) {
    double t0 = now_s();
    Iso est = iso_mul(odom, iso_mul(iso_inv(odom_last), odom));  // EM:238
    odom_last = odom;
    odom = est;
    Quat q = mat2q(odom.R);  // EM:242
    x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
    x[4] = odom.t.x; x[5] = odom.t.y; x[6] = odom.t.z;
    voxel_grid(edge_in, (float)cfg.edge_leaf, vox_mode(), ds_edge);  // EM:248-251
    voxel_grid(surf_in, (float)cfg.surf_leaf, vox_mode(), ds_surf);
    double t1 = now_s();
    timing.ds += t1 - t0;
    traces.clear();
    if (map_edge.size() > 10 && map_surf.size() > 50) {  // EM:254
      tree_edge.set_canonical_ties(cfg.knn_ties == 1);
      tree_surf.set_canonical_ties(cfg.knn_ties == 1);
      tree_edge.build(&map_edge);
      tree_surf.build(&map_surf);
      double t2 = now_s();
      timing.kdbuild += t2 - t1;
      for (int iter = 0; iter < cfg.outer_iters; ++iter) {
        double ta = now_s();
        std::vector<EdgeFactor> ef;
        std::vector<SurfFactor> sf;
        Quat qq{x[0], x[1], x[2], x[3]};
        V3 tt{x[4], x[5], x[6]};
        edge_factors(cfg, qq, tt, ds_edge, map_edge, tree_edge, ef);
        surf_factors(cfg, qq, tt, ds_surf, map_surf, tree_surf, sf);
        double tb = now_s();
        timing.assoc += tb - ta;
        Problem prob{&ef, &sf, cfg.huber};
        SolveTrace tr;
        ceres_solve(prob, x, cfg.lm_max_iters, &tr, cfg.lm_solver);
        traces.push_back(tr);
        timing.solve += now_s() - tb;
      }
    }
    double t3 = now_s();
    odom = iso_identity();  // EM:291-293
    odom.R = qmat(Quat{x[0], x[1], x[2], x[3]});
    odom.t = {x[4], x[5], x[6]};
    create_submap();
    timing.map += now_s() - t3;
    timing.frames++;
  }

  // EM:298-352
  void create_submap() {
    registered.clear();
    no_registered.clear();
    no_registered.insert(no_registered.end(), ds_edge.begin(), ds_edge.end());
    no_registered.insert(no_registered.end(), ds_surf.begin(), ds_surf.end());
    Quat q{x[0], x[1], x[2], x[3]};
    V3 t{x[4], x[5], x[6]};
    for (const P4& p : ds_edge) { P4 w = associate(q, t, p); map_edge.push_back(w); registered.push_back(w); }
    for (const P4& p : ds_surf) { P4 w = associate(q, t, p); map_surf.push_back(w); registered.push_back(w); }
    const double mn[3] = {odom.t.x - cfg.crop_half, odom.t.y - cfg.crop_half, odom.t.z - cfg.crop_half};
    const double mx[3] = {odom.t.x + cfg.crop_half, odom.t.y + cfg.crop_half, odom.t.z + cfg.crop_half};
    Cloud ce, cs;
    crop_box(map_edge, mn, mx, ce);
    crop_box(map_surf, mn, mx, cs);
    voxel_grid(ce, (float)cfg.edge_leaf, vox_mode(), map_edge);
    voxel_grid(cs, (float)cfg.surf_leaf, vox_mode(), map_surf);
  }

  int vox_mode() const { return (cfg.voxel_order & 1) | (cfg.centroid_div ? 2 : 0); }
  Config cfg;
  double x[7] = {0, 0, 0, 1, 0, 0, 0};  // EM:383 parameter_opti
  Iso odom, odom_last;
  Cloud map_edge, map_surf, registered, no_registered, ds_edge, ds_surf;
  KdTree tree_edge, tree_surf;
  std::vector<SolveTrace> traces;
  Timing timing;
};

}  // namespace orc
