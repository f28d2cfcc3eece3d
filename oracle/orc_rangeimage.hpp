// TEST INFRASTRUCTURE — CPU restatement of the reference's ring-field / range-image feature extractor:
// src/visual_inertial_lidar/feature_tracker/include/featureExtract.hpp (class featureExtract, FX below), the alternative
// stage 1 for sensors whose driver supplies ring ids (any beam count), with the helpers it uses from
// src/visual_inertial_lidar/feature_tracker/include/common.h (pointDistance :54-57, DistanceXY :59-62, rad2deg :44-47).
//
// PARITY UNPINNED (the reference holds no vectors for it and the class is not instantiated by any node).  First-party logic is
// restated line by line, including what the code actually does rather than what it seems to intend:
//  * curvature and the occlusion marks run over the FLATTENED cloud, across ring boundaries (FX:236-290);
//  * sectors are sorted over [sp, ep) but walked over [sp, ep]: position ep keeps its natural entry and is visited first (FX:134-138);
//  * the 21st qualifying edge candidate of a sector stops the walk without being marked (FX:147-155);
//  * the "surf" marks of FX:178-203 never decide an output themselves (every non-edge position of a sector is emitted, FX:206-210)
//    but they do block edge candidates of the NEXT sector near the boundary, so they are replayed;
//  * ring 0's first sector starts at position 4, which extractSmoothness (i >= 5) never writes: the entry there is the
//    value-initialised {0, ind 0} of the first call (it sorts first and stays; curvature[0] = 0 is never an edge).  Its surf mark reads
//    pointColInd[-1] (out of bounds, FX:195) — undefined in the reference, treated as "break" here; no output depends on it.
// Deviations kept on BOTH sides (oracle and CUDA path), as for the ring-angle extractor: non-finite returns are dropped on entry
// (the reference would store a NaN range and hand NaN to std::sort); equal curvatures are ordered by position (std::sort is
// unstable: tie class T1); clouds of 10 or fewer image points yield nothing (the reference's size_t loop bounds underflow).
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "orc_pipeline.hpp"

namespace orc {

struct RIParams {
  int n_scan = 64;              // FX:86
  int horizon_scan = 1800;      // FX:85
  int downsample_rate = 1;      // FX:87
  double lidar_min = 3.0;       // FX:88
  double lidar_max = 200.0;     // FX:89
  double edge_threshold = 1.0;  // FX:90
  double surf_threshold = 0.1;  // FX:91
};

struct RIDebug {  // intermediate arrays for stage-level comparisons
  std::vector<int> src;          // semanticCloud index -> input index
  std::vector<int> col;          // pointColInd
  std::vector<float> range;      // pointRange
  std::vector<float> curvature;  // cloudCurvature
  std::vector<int> picked;       // cloudNeighborPicked after markBadPoints
  std::vector<int> start, end;   // startRingIndex / endRingIndex
};

inline void ri_extract(const RIParams& P, const P4* pts, const uint16_t* ring, int n, Cloud& edge, std::vector<int>& edge_src, Cloud& surf,
                       std::vector<int>& surf_src, RIDebug* dbg = nullptr) {
  edge.clear(); surf.clear(); edge_src.clear(); surf_src.clear();
  const int R = P.n_scan, H = P.horizon_scan;
  // ---- projectPointCloud FX:322-370: first point to reach an image cell keeps it ----
  std::vector<int> owner((size_t)R * H, -1);
  std::vector<float> range_mat((size_t)R * H, FLT_MAX);
  const float ang_res_x = (float)(360.0 / (float)H);  // FX:350 (static float)
  for (int i = 0; i < n; ++i) {
    const P4& p = pts[i];
    if (!(std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z))) continue;  // see the header
    const float range = std::sqrt(p.x * p.x + p.y * p.y + p.z * p.z);  // pointDistance
    const float range_xy = std::sqrt(p.x * p.x + p.y * p.y);           // DistanceXY
    if (range_xy < P.lidar_min || range_xy > P.lidar_max) continue;    // FX:337
    const int row = ring[i];
    if (row < 0 || row >= R) continue;                                 // FX:342
    if (row % P.downsample_rate != 0) continue;                        // FX:346
    const float horizon_angle = (float)((double)std::atan2(p.x, p.y) * 180.0 / M_PI);  // FX:349: float atan2, rad2deg in double, stored as float
    int column = (int)(-std::round((horizon_angle - 90.0) / ang_res_x) + H / 2);       // FX:352
    if (column >= H) column -= H;
    if (column < 0 || column >= H) continue;
    if (range_mat[(size_t)row * H + column] != FLT_MAX) continue;      // FX:360
    range_mat[(size_t)row * H + column] = range;
    owner[(size_t)row * H + column] = i;
  }
  // ---- inverProjectCloud FX:293-318 ----
  std::vector<int> src, col, start(R), end(R);
  std::vector<float> rng;
  int count = 0;
  for (int i = 0; i < R; ++i) {
    start[i] = count - 1 + 5;
    for (int j = 0; j < H; ++j)
      if (range_mat[(size_t)i * H + j] != FLT_MAX) { col.push_back(j); rng.push_back(range_mat[(size_t)i * H + j]); src.push_back(owner[(size_t)i * H + j]); ++count; }
    end[i] = count - 1 - 5;
  }
  const int size = count;
  std::vector<float> curv(size > 0 ? size : 1, 0.f);
  std::vector<int> picked(size > 0 ? size : 1, 0), label(size > 0 ? size : 1, 0);
  if (dbg) { dbg->src = src; dbg->col = col; dbg->range = rng; dbg->start = start; dbg->end = end; }
  if (size <= 10) { if (dbg) { dbg->curvature.assign(curv.begin(), curv.begin() + size); dbg->picked.assign(picked.begin(), picked.begin() + size); } return; }
  // ---- extractSmoothness FX:268-290 ----
  struct Sm { float value; int ind; };
  std::vector<Sm> sm(size, Sm{0.f, 0});
  for (int i = 5; i < size - 5; ++i) {
    const float d = rng[i - 5] + rng[i - 4] + rng[i - 3] + rng[i - 2] + rng[i - 1] + rng[i + 5] + rng[i + 4] + rng[i + 3] + rng[i + 2] + rng[i + 1] - rng[i] * 10;
    curv[i] = d * d;
    sm[i].ind = i; sm[i].value = curv[i];
  }
  // ---- markBadPoints FX:230-265 ----
  for (int i = 5; i < size - 6; ++i) {
    const float depth1 = rng[i], depth2 = rng[i + 1];
    const int cdiff = std::abs(int(col[i + 1] - col[i]));
    if (cdiff < 10) {
      if (depth1 - depth2 > 0.3) { for (int k = 0; k <= 5; ++k) picked[i - k] = 1; }
      else if (depth2 - depth1 > 0.3) { for (int k = 1; k <= 6; ++k) picked[i + k] = 1; }
    }
    const float diff1 = std::abs(float(rng[i - 1] - rng[i])), diff2 = std::abs(float(rng[i + 1] - rng[i]));
    if (diff1 > 0.02 * rng[i] && diff2 > 0.02 * rng[i]) picked[i] = 1;
  }
  if (dbg) { dbg->curvature = curv; dbg->picked = picked; }
  // ---- featureEdge_Surf FX:115-227 ----
  auto mark_neighbours = [&](int ind) {
    for (int l = 1; l <= 5; ++l) {
      if (ind + l >= size) break;  // (never reached: ind <= size - 7)
      if (std::abs(int(col[ind + l] - col[ind + l - 1])) > 10) break;
      picked[ind + l] = 1;
    }
    for (int l = -1; l >= -5; --l) {
      if (ind + l < 0) break;      // pointColInd[-1]: out of bounds in the reference, see the header
      if (std::abs(int(col[ind + l] - col[ind + l + 1])) > 10) break;
      picked[ind + l] = 1;
    }
  };
  for (int i = 0; i < R; ++i) {
    for (int j = 0; j < 6; ++j) {
      const int sp = (start[i] * (6 - j) + end[i] * j) / 6;
      const int ep = (start[i] * (5 - j) + end[i] * (j + 1)) / 6 - 1;
      if (sp >= ep) continue;
      std::sort(sm.begin() + sp, sm.begin() + ep, [](const Sm& a, const Sm& b) { return a.value < b.value || (a.value == b.value && a.ind < b.ind); });
      int largest = 0;
      for (int k = ep; k >= sp; --k) {
        const int ind = sm[k].ind;
        if (picked[ind] == 0 && curv[ind] > P.edge_threshold) {
          ++largest;
          if (largest <= 20) { label[ind] = 1; edge.push_back(pts[src[ind]]); edge_src.push_back(src[ind]); }
          else break;
          picked[ind] = 1;
          mark_neighbours(ind);
        }
      }
      for (int k = sp; k < ep; ++k) {
        const int ind = sm[k].ind;
        if (picked[ind] == 0 && curv[ind] < P.surf_threshold) {
          label[ind] = -1;
          picked[ind] = 1;
          mark_neighbours(ind);
        }
      }
      for (int k = sp; k <= ep; ++k)
        if (label[k] <= 0) { surf.push_back(pts[src[k]]); surf_src.push_back(src[k]); }
    }
  }
}

}  // namespace orc
