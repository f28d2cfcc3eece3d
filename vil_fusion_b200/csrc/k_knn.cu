// Stage 2/3 — data association: pcl::KdTreeFLANN::setInputCloud / nearestKSearch(k=5) (EM:256-257, :128,
// :185) replaced by an exact-within-the-gate spatial-hash grid, fused with the per-point line / plane fit of
// EdgeCostFactor / SurfCostFactor (EM:117-232).
//
// Grid: cell edge c = a power of two >= sqrt(knn_gate) (1.0 m for the reference's gate of 1.0 m^2), integer
// cell coordinates floor(x / c) are exact in fp32, bucket = hash(cell) & (H - 1), H = pow2 >= 2 * points.
// Build = counting sort (zero, count with atomics, exclusive scan, scatter): 16 B read + 16 B written per
// map point + 8 B per bucket.  The reordered copy carries the original map index in .w, so neighbour indices
// are reported in the reference's map order.
//
// Query: one thread per feature point walks the 27 buckets around its cell (float4 loads through the read-only path; the
// lanes of a warp are spatial neighbours and share buckets).  Every point closer than c is inside those 27 cells, hence
// every neighbour with d^2 < knn_gate is found: the result is exact wherever the reference uses it (EM:129/:189 reject the
// feature when d^2[4] >= 1); candidates at or beyond the gate are dropped on sight (their slots stay idx -1, d^2 FLT_MAX).
// Distances: ((dx*dx)+dy*dy)+dz*dz in fp32 without contraction = FLANN L2_Simple<float>.  Top-5 kept in registers,
// ordered by (d^2, index): deterministic tie break (tie class T2).
#include "vilf_internal.cuh"

namespace vilf {

__device__ __forceinline__ uint32_t cell_hash(int x, int y, int z) {
  return ((uint32_t)x * 73856093u) ^ ((uint32_t)y * 19349663u) ^ ((uint32_t)z * 83492791u);
}
__device__ __forceinline__ int grid_buckets(int n, int hcap) {
  int h = 1024;
  while (h < 2 * n && h < hcap) h <<= 1;
  return h;
}
__device__ __forceinline__ void cell_of(float4 p, float inv_cell, int& cx, int& cy, int& cz) {
  cx = (int)floorf(fmul(p.x, inv_cell));
  cy = (int)floorf(fmul(p.y, inv_cell));
  cz = (int)floorf(fmul(p.z, inv_cell));
}

__global__ void __launch_bounds__(256) k_grid_zero(const GridJob* __restrict__ jobs) {
  const GridJob& J = jobs[blockIdx.y];
  const int H = grid_buckets(*J.n, J.hcap);
  for (int i = blockIdx.x * 256 + threadIdx.x; i <= H; i += gridDim.x * 256) J.start[i] = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) *J.hvar = H;
}

__global__ void __launch_bounds__(256) k_grid_count(const GridJob* __restrict__ jobs) {
  const GridJob& J = jobs[blockIdx.y];
  const float inv_cell = J.inv_cell;
  const int n = *J.n;
  const uint32_t hm = (uint32_t)grid_buckets(n, J.hcap) - 1u;
  // the atomic returns the point's slot inside its bucket, so a thread waits for a full L2 round trip per point:
  // four independent points in flight per thread
  const int stride = gridDim.x * 256;
  for (int i0 = blockIdx.x * 256 + threadIdx.x; i0 < n; i0 += 4 * stride) {
    uint32_t r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * stride;
      r[u] = 0;
      if (i < n) {
        int cx, cy, cz;
        cell_of(J.pts[i], inv_cell, cx, cy, cz);
        r[u] = atomicAdd(&J.start[cell_hash(cx, cy, cz) & hm], 1u);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * stride;
      if (i < n) J.rank[i] = r[u];
    }
  }
}

__device__ __forceinline__ void grid_chunk(int total, int b, int& beg, int& end) {
  int chunk = (total + GRID_G - 1) / GRID_G;
  chunk = (chunk + 255) / 256 * 256;
  beg = min(total, b * chunk);
  end = min(total, beg + chunk);
}

__global__ void __launch_bounds__(256) k_grid_scan_partial(const GridJob* __restrict__ jobs) {
  const GridJob& J = jobs[blockIdx.y];
  const int total = grid_buckets(*J.n, J.hcap) + 1;
  int beg, end;
  grid_chunk(total, blockIdx.x, beg, end);
  uint32_t s = 0;
  for (int i = beg + threadIdx.x; i < end; i += 256) s += J.start[i];
  __shared__ uint32_t red[8];
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < 8; ++w) t += red[w];
    J.partial[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) k_grid_scan_final(const GridJob* __restrict__ jobs) {
  const GridJob& J = jobs[blockIdx.y];
  const int total = grid_buckets(*J.n, J.hcap) + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ uint32_t red[256];
  __shared__ uint32_t wsum[8];
  uint32_t pre = 0;
  for (int b = tid; b < (int)blockIdx.x; b += 256) pre += J.partial[b];
  red[tid] = pre;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) { if (tid < off) red[tid] += red[tid + off]; __syncthreads(); }
  uint32_t run = red[0];
  int beg, end;
  grid_chunk(total, blockIdx.x, beg, end);
  // tiles of 2048 counters: every thread scans 8 consecutive ones serially, the 256 thread sums are scanned by the block
  for (int base = beg; base < end; base += 2048) {
    const int i0 = base + tid * 8;
    uint32_t v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = i0 + u < end ? J.start[i0 + u] : 0u;
    uint32_t sum = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) sum += v[u];
    uint32_t inc = sum;  // inclusive warp scan of the thread sums
    for (int off = 1; off < 32; off <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += t; }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { const uint32_t c = wsum[w]; if (w < warp) woff += c; tot += c; }
    uint32_t ex = run + woff + inc - sum;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (i0 + u < end) J.start[i0 + u] = ex;
      ex += v[u];
    }
    run += tot;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_grid_scatter(const GridJob* __restrict__ jobs) {
  const GridJob& J = jobs[blockIdx.y];
  const float inv_cell = J.inv_cell;
  const int n = *J.n;
  const uint32_t hm = (uint32_t)grid_buckets(n, J.hcap) - 1u;
  const int stride = gridDim.x * 256;
  for (int i0 = blockIdx.x * 256 + threadIdx.x; i0 < n; i0 += 4 * stride) {  // four dependent (bucket start, slot) lookups in flight
    float4 p[4];
    uint32_t pos[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * stride;
      pos[u] = 0;
      if (i < n) {
        p[u] = J.pts[i];
        int cx, cy, cz;
        cell_of(p[u], inv_cell, cx, cy, cz);
        pos[u] = J.start[cell_hash(cx, cy, cz) & hm] + J.rank[i];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * stride;
      if (i < n) J.sorted[pos[u]] = make_float4(p[u].x, p[u].y, p[u].z, __int_as_float(i));
    }
  }
}

void launch_grid_build(const Launch& L, const GridJob* jobs_dev, int njobs, const ConfigDev&) {
  dim3 g(GRID_G, njobs);
  k_grid_zero<<<g, 256, 0, L.st>>>(jobs_dev);
  L.tick(K_GRID_ZERO);
  k_grid_count<<<g, 256, 0, L.st>>>(jobs_dev);
  L.tick(K_GRID_COUNT);
  k_grid_scan_partial<<<g, 256, 0, L.st>>>(jobs_dev);
  L.tick(K_GRID_SCAN_PARTIAL);
  k_grid_scan_final<<<g, 256, 0, L.st>>>(jobs_dev);
  L.tick(K_GRID_SCAN_FINAL);
  k_grid_scatter<<<g, 256, 0, L.st>>>(jobs_dev);
  L.tick(K_GRID_SCATTER);
}

// ------------------------------------------------------------------------------------------------
// 5-NN, one THREAD per query
// ------------------------------------------------------------------------------------------------
// The first version gave every query a whole warp (27 lanes looked up one bucket each, candidates were read 32 at a time
// and inserted through warp shuffles): 231 warp instructions per query, issue bound (profiles/r1b, r1d).  With ~30
// candidates per query there is not enough work to amortise the warp-wide bookkeeping, so each thread now walks its own 27
// buckets.  Consecutive queries are spatial neighbours (the voxel filter emits them in voxel order), hence the lanes of a
// warp read the same buckets and points: the loads hit L1 and the per-lane trip counts are similar.
//
// Exactness: a bucket holds every point of its cell (plus, rarely, points of cells that hash alike; those are farther than
// one cell edge >= sqrt(gate), unless the colliding cell is itself one of the 27 — then the same bucket is visited twice and
// the duplicate is rejected by index).  So every map point with d^2 < gate is seen at least once and kept at most once.
struct Top5 {
  float d[5];
  int id[5];
};
// (d^2, index) lexicographic order, evaluated without branches
__device__ __forceinline__ bool closer(float d, int id, float bd, int bid) { return (d < bd) | ((d == bd) & (id < bid)); }

__device__ __forceinline__ void top5_insert(Top5& t, float cd, int ci) {  // precondition: (cd, ci) is closer than t[4] and not in t
  t.d[4] = cd; t.id[4] = ci;
#pragma unroll
  for (int k = 4; k > 0; --k) {
    const bool sw = closer(t.d[k], t.id[k], t.d[k - 1], t.id[k - 1]);
    const float dk = sw ? t.d[k - 1] : t.d[k], dk1 = sw ? t.d[k] : t.d[k - 1];
    const int ik = sw ? t.id[k - 1] : t.id[k], ik1 = sw ? t.id[k] : t.id[k - 1];
    t.d[k] = dk; t.d[k - 1] = dk1; t.id[k] = ik; t.id[k - 1] = ik1;
  }
}

// EIGHT lanes per query (four queries per warp).  A whole warp per query was issue bound (515 warp instructions per query
// for ~30 candidates, profiles/r1b), one thread per query latency bound (27 x 2 dependent L2 round trips per thread, 9.7 of
// 32 lanes active, profiles/r1f).  Here lane sl of a group owns buckets sl, sl+8, sl+16, sl+24 of the 27: their bounds
// are 8 independent loads, the first two candidates of each of its buckets 8 more, so a query costs ~3 dependent round
// trips.  Every lane keeps the best five of ITS candidates; five rounds of a group arg-min (xor shuffles over 8 lanes)
// merge them, and every lane holding the winner pops it — which also removes duplicates that reach two lanes through
// colliding buckets.  Same (d^2, index) order as before, so the result is bit-identical.
constexpr int KNN_THREADS = 128;
#ifndef VILF_KNN_GROUP
#define VILF_KNN_GROUP 8
#endif
constexpr int KNN_GROUP = VILF_KNN_GROUP;          // lanes per query (8: see above; 4 was measured)
constexpr int KNN_SLOTS = (27 + KNN_GROUP - 1) / KNN_GROUP;  // buckets of the first 27 cells per lane

__device__ __forceinline__ void knn_consider(Top5& best, float gate_f, float qx, float qy, float qz, const float4 c) {
  const float ddx = fsub(qx, c.x), ddy = fsub(qy, c.y), ddz = fsub(qz, c.z);
  const float cd = fadd(fadd(fmul(ddx, ddx), fmul(ddy, ddy)), fmul(ddz, ddz));  // FLANN L2_Simple<float>
  const int ci = __float_as_int(c.w);
  // neighbours at or beyond the gate can never be used (EM:129 / :189 reject the feature), so they never enter the list
  if ((cd < gate_f) & closer(cd, ci, best.d[4], best.id[4])) {
    const bool dup = (ci == best.id[0]) | (ci == best.id[1]) | (ci == best.id[2]) | (ci == best.id[3]);  // (ci != id[4]: strictly closer)
    if (!dup) top5_insert(best, cd, ci);
  }
}

// Walk one bucket: the first PF candidates were loaded by the caller, the rest follow.
template <int PF>
__device__ __forceinline__ void knn_bucket(Top5& best, float gate_f, float qx, float qy, float qz, const float4* __restrict__ sorted, uint32_t s, uint32_t e,
                                           const float4 (&c)[PF]) {
#pragma unroll
  for (int i = 0; i < PF; ++i)
    if (s + i < e) knn_consider(best, gate_f, qx, qy, qz, c[i]);
  for (uint32_t p = s + PF; p < e; ++p) knn_consider(best, gate_f, qx, qy, qz, __ldg(sorted + p));
}

// Called by all 32 lanes; the lanes of a group (lane >> 3) pass the same query.  Returns, replicated in every lane of the group,
// the five nearest neighbours in ascending (d^2, index) order (FLT_MAX / INT_MAX where fewer than five lie inside the gate).
//
// Cells of Chebyshev distance <= 1 (27 buckets) come first.  Grids of dense maps have cells finer than the gate radius
// (G.rings > 1): further shells follow, and the walk stops as soon as the fifth-best distance is certainly smaller than the
// distance to anything in an unvisited shell: after shell r every unseen point is farther than (r + f) * cell, f = the
// query's distance to the nearest face of its own cell in cell units.  The test uses an upper bound of the group's true
// fifth distance (the smallest fifth distance any single lane holds) and a 1e-5 relative margin for the fp32 rounding of
// computed distances, so it can only stop late, never early: the result is the same exact 5-NN as the full walk.
__device__ __forceinline__ void group_knn5(const GridJob& G, float gate_f, float qx, float qy, float qz, bool active, float (&rd)[5], int (&ri)[5]) {
  constexpr int PF = KNN_GROUP >= 8 ? 2 : 1;
  const int sl = threadIdx.x & (KNN_GROUP - 1);
  Top5 best;
#pragma unroll
  for (int k = 0; k < 5; ++k) { best.d[k] = FLT_MAX; best.id[k] = INT_MAX; }
  const float inv_cell = G.inv_cell;
  const int rings = G.rings;
  int qcx = 0, qcy = 0, qcz = 0;
  uint32_t hm = 0;
  const uint32_t* __restrict__ start = G.start;
  const float4* __restrict__ sorted = G.sorted;
  if (active) {
    hm = (uint32_t)(*G.hvar) - 1u;
    cell_of(make_float4(qx, qy, qz, 0.f), inv_cell, qcx, qcy, qcz);
    uint32_t s[KNN_SLOTS], e[KNN_SLOTS];
#pragma unroll
    for (int j = 0; j < KNN_SLOTS; ++j) {
      const int c = sl + KNN_GROUP * j;  // cell (c % 3 - 1, (c / 3) % 3 - 1, c / 9 - 1)
      s[j] = 0; e[j] = 0;
      if (c < 27) {
        const uint32_t h = cell_hash(qcx + c % 3 - 1, qcy + (c / 3) % 3 - 1, qcz + c / 9 - 1) & hm;
        s[j] = __ldg(start + h);
        e[j] = __ldg(start + h + 1);
      }
    }
    float4 c[KNN_SLOTS][PF];
#pragma unroll
    for (int j = 0; j < KNN_SLOTS; ++j)
#pragma unroll
      for (int i = 0; i < PF; ++i)
        if (s[j] + i < e[j]) c[j][i] = __ldg(sorted + s[j] + i);
#pragma unroll
    for (int j = 0; j < KNN_SLOTS; ++j) knn_bucket<PF>(best, gate_f, qx, qy, qz, sorted, s[j], e[j], c[j]);
  }
  // the groups of a warp may use different grids (edge / surf): the shell loop runs to the warp-wide maximum
  const int rmax = __reduce_max_sync(0xffffffffu, active ? rings : 1);
  if (rmax > 1) {
    float fmin_cells = 0.f;
    if (active) {
      const float ux = fmul(qx, inv_cell), uy = fmul(qy, inv_cell), uz = fmul(qz, inv_cell);
      const float fx = fsub(ux, floorf(ux)), fy = fsub(uy, floorf(uy)), fz = fsub(uz, floorf(uz));
      fmin_cells = fminf(fminf(fminf(fx, 1.f - fx), fminf(fy, 1.f - fy)), fminf(fz, 1.f - fz));
    }
    const double cell = 1.0 / (double)inv_cell;
    for (int r = 1; r < rmax; ++r) {  // shells 0..r are done; is shell r + 1 needed?
      float ub = best.d[4];            // a lane's own fifth distance bounds the group's from above
#pragma unroll
      for (int off = 1; off < KNN_GROUP; off <<= 1) ub = fminf(ub, __shfl_xor_sync(0xffffffffu, ub, off));
      const double reach = ((double)r + (double)fmin_cells) * cell;
      const bool done = !active || r >= rings || (double)ub < reach * reach * (1.0 - 1e-5);
      if (__all_sync(0xffffffffu, done)) break;
      if (!done) {
        const int R = r + 1, side = 2 * R + 1, ncell = side * side * side;
        for (int t = sl; t < ncell; t += KNN_GROUP) {
          const int dx = t % side - R, dy = (t / side) % side - R, dz = t / (side * side) - R;
          if (max(max(abs(dx), abs(dy)), abs(dz)) != R) continue;  // interior: visited before
          const uint32_t h = cell_hash(qcx + dx, qcy + dy, qcz + dz) & hm;
          const uint32_t s0 = __ldg(start + h), e0 = __ldg(start + h + 1);
          for (uint32_t p = s0; p < e0; ++p) knn_consider(best, gate_f, qx, qy, qz, __ldg(sorted + p));
        }
      }
    }
  }
#pragma unroll
  for (int round = 0; round < 5; ++round) {
    float wd = best.d[0];
    int wi = best.id[0];
#pragma unroll
    for (int off = 1; off < KNN_GROUP; off <<= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, wd, off);
      const int oi = __shfl_xor_sync(0xffffffffu, wi, off);
      const bool take = closer(od, oi, wd, wi);
      wd = take ? od : wd;
      wi = take ? oi : wi;
    }
    rd[round] = wd; ri[round] = wi;  // known to every lane of the group
    const bool pop = (best.id[0] == wi) & (wi != INT_MAX);  // every lane holding the winner drops it
#pragma unroll
    for (int k = 0; k < 4; ++k) { best.d[k] = pop ? best.d[k + 1] : best.d[k]; best.id[k] = pop ? best.id[k + 1] : best.id[k]; }
    best.d[4] = pop ? FLT_MAX : best.d[4];
    best.id[4] = pop ? INT_MAX : best.id[4];
  }
}

// ------------------------------------------------------------------------------------------------
// line / plane fits (fp64, same operation order as the oracle's restatement of Eigen)
// ------------------------------------------------------------------------------------------------
// Symmetric 3x3 eigen-decomposition by cyclic Jacobi (stands in for Eigen::SelfAdjointEigenSolver, EM:150).
// Returns the two largest eigenvalues and the eigenvector of the largest.
__device__ void eig3_largest(const double C[3][3], double& w_mid, double& w_max, D3& dir) {
  double a[3][3], V[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) { a[i][j] = C[i][j]; V[i][j] = i == j ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 60; ++sweep) {
    const double off = dadd(dadd(dmul(a[0][1], a[0][1]), dmul(a[0][2], a[0][2])), dmul(a[1][2], a[1][2]));
    const double diag = dadd(dadd(dmul(a[0][0], a[0][0]), dmul(a[1][1], a[1][1])), dmul(a[2][2], a[2][2]));
    if (off <= dmul(1e-34, diag) || off == 0.0) break;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        const double theta = dsub(a[q][q], a[p][p]) / dmul(2.0, a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / dadd(fabs(theta), sqrt(dadd(dmul(theta, theta), 1.0)));
        const double c = 1.0 / sqrt(dadd(dmul(t, t), 1.0)), s = dmul(t, c);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = dsub(dmul(c, akp), dmul(s, akq));
          a[k][q] = dadd(dmul(s, akp), dmul(c, akq));
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = dsub(dmul(c, apk), dmul(s, aqk));
          a[q][k] = dadd(dmul(s, apk), dmul(c, aqk));
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = dsub(dmul(c, vkp), dmul(s, vkq));
          V[k][q] = dadd(dmul(s, vkp), dmul(c, vkq));
        }
      }
  }
  const double w0 = a[0][0], w1 = a[1][1], w2 = a[2][2];
  // ascending order with the insertion-sort tie behaviour of the oracle (first of equals stays first)
  int i0 = 0, i1 = 1, i2 = 2;
  double s0 = w0, s1 = w1, s2 = w2;
  if (s1 < s0) { double t = s0; s0 = s1; s1 = t; int ti = i0; i0 = i1; i1 = ti; }
  if (s2 < s1) {
    double t = s1; s1 = s2; s2 = t; int ti = i1; i1 = i2; i2 = ti;
    if (s1 < s0) { t = s0; s0 = s1; s1 = t; ti = i0; i0 = i1; i1 = ti; }
  }
  w_mid = s1; w_max = s2;
  dir = i2 == 0 ? d3(V[0][0], V[1][0], V[2][0]) : (i2 == 1 ? d3(V[0][1], V[1][1], V[2][1]) : d3(V[0][2], V[1][2], V[2][2]));
}

// 5x3 least squares by Householder QR with column pivoting (stands in for colPivHouseholderQr().solve, EM:198).
__device__ D3 lstsq5x3(const double Ain[5][3]) {
  double A[5][3], b[5];
#pragma unroll
  for (int i = 0; i < 5; ++i) { b[i] = -1.0; for (int j = 0; j < 3; ++j) A[i][j] = Ain[i][j]; }
  int perm[3] = {0, 1, 2};
  double maxcol = 0;
  for (int j = 0; j < 3; ++j) {
    double s = 0;
    for (int i = 0; i < 5; ++i) s = dadd(s, dmul(A[i][j], A[i][j]));
    maxcol = fmax(maxcol, sqrt(s));
  }
  const double me = dmul(maxcol, 2.220446049250313e-16);
  const double thr = dmul(me, me) / 5.0;
  int npiv = 3;
  for (int k = 0; k < 3; ++k) {
    int bestj = k;
    double bn = -1;
    for (int j = k; j < 3; ++j) {
      double s = 0;
      for (int i = k; i < 5; ++i) s = dadd(s, dmul(A[i][j], A[i][j]));
      if (s > bn) { bn = s; bestj = j; }
    }
    if (npiv == 3 && bn < dmul(thr, (double)(5 - k))) { npiv = k; break; }
    if (bestj != k) {
      for (int i = 0; i < 5; ++i) { const double t = A[i][k]; A[i][k] = A[i][bestj]; A[i][bestj] = t; }
      const int tp = perm[k]; perm[k] = perm[bestj]; perm[bestj] = tp;
    }
    double tail = 0;
    for (int i = k + 1; i < 5; ++i) tail = dadd(tail, dmul(A[i][k], A[i][k]));
    const double c0 = A[k][k];
    double beta, tau, v[5];
    if (tail <= 2.2250738585072014e-308) {
      tau = 0; beta = c0;
      for (int i = k + 1; i < 5; ++i) v[i] = 0;
    } else {
      beta = sqrt(dadd(dmul(c0, c0), tail));
      if (c0 >= 0) beta = -beta;
      for (int i = k + 1; i < 5; ++i) v[i] = A[i][k] / dsub(c0, beta);
      tau = dsub(beta, c0) / beta;
    }
    A[k][k] = beta;
    for (int i = k + 1; i < 5; ++i) A[i][k] = 0;
    for (int j = k + 1; j < 3; ++j) {
      double tmp = A[k][j];
      for (int i = k + 1; i < 5; ++i) tmp = dadd(tmp, dmul(v[i], A[i][j]));
      A[k][j] = dsub(A[k][j], dmul(tau, tmp));
      for (int i = k + 1; i < 5; ++i) A[i][j] = dsub(A[i][j], dmul(dmul(tau, v[i]), tmp));
    }
    double tmp = b[k];
    for (int i = k + 1; i < 5; ++i) tmp = dadd(tmp, dmul(v[i], b[i]));
    b[k] = dsub(b[k], dmul(tau, tmp));
    for (int i = k + 1; i < 5; ++i) b[i] = dsub(b[i], dmul(dmul(tau, v[i]), tmp));
  }
  double y[3] = {0, 0, 0};
  for (int k = npiv - 1; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < npiv; ++j) s = dsub(s, dmul(A[k][j], y[j]));
    y[k] = s / A[k][k];
  }
  double n[3];
  for (int k = 0; k < 3; ++k) n[perm[k]] = y[k];
  return d3(n[0], n[1], n[2]);
}

// Association, kernel 1: eight lanes per voxel-filtered feature point: transform (EM:355-363) + 5-NN (EM:128 / :185).
__global__ void __launch_bounds__(KNN_THREADS) k_knn_assoc(LaneDev* lanes, const GridJob* __restrict__ grid_jobs, int lane0, ConfigDev cfg,
                                                            const double* pose_override) {
  const int ln = lane0 + blockIdx.y;
  const LaneDev& L = lanes[ln];
  LaneVars& V = *L.v;
  const int me = V.n_map[0], ms = V.n_map[1];
  if (!(me > 10 && ms > 50)) return;  // EM:254
  const int ne = V.n_ds[0], ns = V.n_ds[1];
  double x[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) x[i] = pose_override ? pose_override[i] : V.x[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) V.opt_ran = 1;
  const int sl = threadIdx.x & (KNN_GROUP - 1);
  const int gpb = KNN_THREADS / KNN_GROUP;  // queries per CTA per sweep
  const int nq = ne + ns;
  for (int base = blockIdx.x * gpb; base < nq; base += gridDim.x * gpb) {  // whole warps iterate together (shuffles inside)
    const int q = base + (threadIdx.x / KNN_GROUP);
    const bool active = q < nq;
    const int w = (active && q >= ne) ? 1 : 0;
    const int k = active ? (q < ne ? q : q - ne) : 0;
    float4 pw = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) pw = associate(x, L.ds[w][k]);
    float rd[5];
    int ri[5];
    group_knn5(grid_jobs[ln * 2 + w], cfg.knn_gate_f, pw.x, pw.y, pw.z, active, rd, ri);
#pragma unroll
    for (int j = 0; j < 5; ++j)
      if (active && (j % KNN_GROUP) == sl) {
        L.nn_idx[w][k * 5 + j] = ri[j] == INT_MAX ? -1 : ri[j];
        L.nn_d2[w][k * 5 + j] = rd[j];
      }
  }
}

// Association, kernel 2: one THREAD per feature point (so the FP64 pipe runs full warps): the 3x3 eigen line fit
// (EM:131-163) or the 5x3 least-squares plane fit (EM:187-222) on the five neighbours, and the factor record.
__global__ void __launch_bounds__(128) k_fit(LaneDev* lanes, int lane0, int cur, ConfigDev cfg) {
  const LaneDev& L = lanes[lane0 + blockIdx.y];
  const LaneVars& V = *L.v;
  if (!(V.n_map[0] > 10 && V.n_map[1] > 50)) return;  // EM:254
  const int ne = V.n_ds[0], ns = V.n_ds[1];
  for (int q = blockIdx.x * 128 + threadIdx.x; q < ne + ns; q += gridDim.x * 128) {
    const int w = q < ne ? 0 : 1;
    const int k = q < ne ? q : q - ne;
    uint8_t valid = 0;
    if ((double)L.nn_d2[w][k * 5 + 4] < cfg.knn_gate) {  // EM:129 / :189
      const float4* map = L.map[w][cur];
      const float4 p = L.ds[w][k];
      D3 nb[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) { const float4 m = map[L.nn_idx[w][k * 5 + j]]; nb[j] = d3((double)m.x, (double)m.y, (double)m.z); }
      if (w == 0) {
        D3 center = d3(0, 0, 0);
#pragma unroll
        for (int j = 0; j < 5; ++j) center = center + nb[j];
        center = d3(center.x / 5.0, center.y / 5.0, center.z / 5.0);
        double cov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const D3 z = nb[j] - center;
          const double v[3] = {z.x, z.y, z.z};
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) cov[r][c] = dadd(cov[r][c], dmul(v[r], v[c]));
        }
        double w_mid, w_max;
        D3 dir;
        eig3_largest(cov, w_mid, w_max, dir);
        if (w_max > dmul(3.0, w_mid)) {  // EM:153
          const D3 a = 0.1 * dir + center, b = -0.1 * dir + center;  // EM:156-157
          double* o = L.edge_pab + (size_t)k * 9;
          o[0] = p.x; o[1] = p.y; o[2] = p.z;
          o[3] = a.x; o[4] = a.y; o[5] = a.z;
          o[6] = b.x; o[7] = b.y; o[8] = b.z;
          valid = 1;
        }
      } else {
        double A[5][3];
#pragma unroll
        for (int j = 0; j < 5; ++j) { A[j][0] = nb[j].x; A[j][1] = nb[j].y; A[j][2] = nb[j].z; }
        D3 n = lstsq5x3(A);
        const double nn = norm3(n);
        const double d = 1.0 / nn;            // EM:199
        n = d3(n.x / nn, n.y / nn, n.z / nn);  // EM:200
        bool okp = true;
#pragma unroll
        for (int j = 0; j < 5; ++j)
          if (fabs(dadd(dadd(dadd(dmul(n.x, nb[j].x), dmul(n.y, nb[j].y)), dmul(n.z, nb[j].z)), d)) > 0.2) okp = false;  // EM:203-213
        if (okp) {
          double* o = L.surf_pnd + (size_t)k * 7;
          o[0] = p.x; o[1] = p.y; o[2] = p.z;
          o[3] = n.x; o[4] = n.y; o[5] = n.z; o[6] = d;
          valid = 1;
        }
      }
    }
    L.fvalid[w][k] = valid;
  }
}

void launch_knn_fit(const Launch& L, LaneDev* lanes, const GridJob* grid_jobs, int lane0, int nlanes, int cur, const ConfigDev& cfg,
                    const double* pose_override) {
  dim3 g(KNN_G * 4, nlanes);
  k_knn_assoc<<<g, KNN_THREADS, 0, L.st>>>(lanes, grid_jobs, lane0, cfg, pose_override);
  L.tick(K_KNN_FIT);
  launch_fit(L, lanes, lane0, nlanes, cur, cfg);
}
void launch_fit(const Launch& L, LaneDev* lanes, int lane0, int nlanes, int cur, const ConfigDev& cfg) {
  dim3 g2(FIT_G, nlanes);
  k_fit<<<g2, 128, 0, L.st>>>(lanes, lane0, cur, cfg);
  L.tick(K_FIT);
}

// nearestKSearch alone against an explicit map (test entry point vilf_knn5).
__global__ void __launch_bounds__(KNN_THREADS) k_knn_only(const GridJob* __restrict__ job, const float4* __restrict__ q, const int* nq_dev, int* idx,
                                                           float* d2, float gate_f) {
  const int nq = *nq_dev;
  const int sl = threadIdx.x & (KNN_GROUP - 1);
  const int gpb = KNN_THREADS / KNN_GROUP;
  for (int base = blockIdx.x * gpb; base < nq; base += gridDim.x * gpb) {
    const int i = base + (threadIdx.x / KNN_GROUP);
    const bool active = i < nq;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) p = q[i];
    float rd[5];
    int ri[5];
    group_knn5(*job, gate_f, p.x, p.y, p.z, active, rd, ri);
#pragma unroll
    for (int j = 0; j < 5; ++j)
      if (active && (j % KNN_GROUP) == sl) {
        idx[i * 5 + j] = ri[j] == INT_MAX ? -1 : ri[j];
        d2[i * 5 + j] = rd[j];
      }
  }
}

void launch_knn_only(const Launch& L, const GridJob* job_dev, const float4* q, const int* nq_dev, int* idx, float* d2, const ConfigDev& cfg) {
  k_knn_only<<<KNN_G * 4, KNN_THREADS, 0, L.st>>>(job_dev, q, nq_dev, idx, d2, cfg.knn_gate_f);
  L.tick(K_KNN_ONLY);
}

}  // namespace vilf
