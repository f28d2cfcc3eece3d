// Stage 1 — featureExtraction::extractFeature (FE:223-232) on the device.
//
//   k_frame_reset        per-frame scratch reset + constant-velocity pose prediction (EM:238-243)
//   k_ring_partition     getLaserCloud (FE:54-110): range gate + vertical-angle -> ring id, and the stable partition of
//                        the scan by ring (arrival order kept inside a ring, FE:108), one cluster per sequence
//   k_sector_warp<EPL>   featureEdge_Surf + featureExtractionFromSector (FE:112-220): one WARP per (ring, sector):
//                        11-tap fp32 curvature, register bitonic sort of 32-bit (curvature prefix, index) keys with
//                        exact tie repair, warp-serial greedy edge pick with +-5 neighbour suppression, surf =
//                        everything not picked
//   k_sector_select      the same per CTA with a shared-memory sort: fallback for sectors longer than 512 elements
//   k_compact_features   concatenates the per-sector lists in (ring, sector) order = the reference's
//                        push_back order of cloud_Edge / cloud_Surf
#include "k_sort.cuh"
#include <cooperative_groups.h>

namespace vilf {

// ------------------------------------------------------------------------------------------------
// frame reset + prediction
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mat3_mul(const double* a, const double* b, double* r) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      r[i * 3 + j] = dadd(dadd(dmul(a[i * 3 + 0], b[0 * 3 + j]), dmul(a[i * 3 + 1], b[1 * 3 + j])), dmul(a[i * 3 + 2], b[2 * 3 + j]));
}
__device__ __forceinline__ void mat3_vec(const double* a, const double* v, double* r) {
#pragma unroll
  for (int i = 0; i < 3; ++i) r[i] = dadd(dadd(dmul(a[i * 3 + 0], v[0]), dmul(a[i * 3 + 1], v[1])), dmul(a[i * 3 + 2], v[2]));
}
// Eigen Transform<double,3,Isometry> product: (A*B).R = A.R*B.R, (A*B).t = A.R*B.t + A.t
__device__ __forceinline__ void iso_mul(const double* a, const double* b, double* r) {
  mat3_mul(a, b, r);
  double v[3];
  mat3_vec(a, b + 9, v);
#pragma unroll
  for (int i = 0; i < 3; ++i) r[9 + i] = dadd(v[i], a[9 + i]);
}
__device__ __forceinline__ void iso_inv(const double* a, double* r) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r[i * 3 + j] = a[j * 3 + i];
  double v[3];
  mat3_vec(r, a + 9, v);
#pragma unroll
  for (int i = 0; i < 3; ++i) r[9 + i] = -v[i];
}
// Eigen Quaternion(Matrix3) assignment (EM:242)
__device__ void mat_to_quat(const double* a, double* q /*x y z w*/) {
  double t = dadd(dadd(a[0], a[4]), a[8]);
  if (t > 0) {
    t = sqrt(dadd(t, 1.0));
    q[3] = dmul(0.5, t);
    t = 0.5 / t;
    q[0] = dmul(dsub(a[7], a[5]), t);
    q[1] = dmul(dsub(a[2], a[6]), t);
    q[2] = dmul(dsub(a[3], a[1]), t);
  } else {
    int i = 0;
    if (a[4] > a[0]) i = 1;
    if (a[8] > a[i * 3 + i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(dadd(dsub(dsub(a[i * 3 + i], a[j * 3 + j]), a[k * 3 + k]), 1.0));
    q[i] = dmul(0.5, t);
    t = 0.5 / t;
    q[3] = dmul(dsub(a[k * 3 + j], a[j * 3 + k]), t);
    q[j] = dmul(dadd(a[j * 3 + i], a[i * 3 + j]), t);
    q[k] = dmul(dadd(a[k * 3 + i], a[i * 3 + k]), t);
  }
}

__global__ void k_frame_reset(LaneDev* lanes, int lane0, int nlanes, VoxVars* vv, int vv_per_lane, int predict) {
  const int t = threadIdx.x;
  if (t < nlanes * vv_per_lane) {
    VoxVars& v = vv[(size_t)lane0 * vv_per_lane + t];
    v.bbox[0] = v.bbox[1] = v.bbox[2] = INT_MAX;
    v.bbox[3] = v.bbox[4] = v.bbox[5] = INT_MIN;
    v.n_valid = 0;
    v.guard = 0;
  }
  if (t < nlanes) {
    LaneVars& L = *lanes[lane0 + t].v;
    L.opt_ran = 0;
    L.status = 0;  // error bits describe the current frame
    if (predict) {  // EM:238-243
      double inv[12], rel[12], est[12], od[12], ol[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) { od[i] = L.odom[i]; ol[i] = L.odom_last[i]; }
      iso_inv(ol, inv);
      iso_mul(inv, od, rel);
      iso_mul(od, rel, est);
#pragma unroll
      for (int i = 0; i < 12; ++i) { L.odom_last[i] = od[i]; L.odom[i] = est[i]; }
      double q[4];
      mat_to_quat(est, q);
      L.x[0] = q[0]; L.x[1] = q[1]; L.x[2] = q[2]; L.x[3] = q[3];
      L.x[4] = est[9]; L.x[5] = est[10]; L.x[6] = est[11];
      SolveTraceDev* tr = lanes[lane0 + t].trace;
      for (int o = 0; o < MAX_OUTER; ++o) { tr[o].n_rows = 0; tr[o].n_edge = 0; tr[o].n_surf = 0; tr[o].termination = -1; }
    }
  }
}

void launch_frame_reset(const Launch& L, LaneDev* lanes, int lane0, int nlanes, VoxVars* vv, int vv_per_lane, int predict) {
  k_frame_reset<<<1, 256, 0, L.st>>>(lanes, lane0, nlanes, vv, vv_per_lane, predict);
  L.tick(K_RESET);
}

// ------------------------------------------------------------------------------------------------
// ring classification as a radix key (FE:54-110)
// ------------------------------------------------------------------------------------------------
struct KeyGenRing {
  const LaneDev* lanes;
  int lane0, sel;
  ConfigDev cfg;
  __device__ int prepare(int) const { return 8; }
  __device__ uint32_t key(int job, int i) const {
    const LaneDev& L = lanes[lane0 + job];
    const float4 p = L.scan[sel][i];
    // Non-finite returns (organized clouds of real drivers carry NaN rows): removeNaNFromPointCloud at FE:56-57 only fills an
    // index vector, the points stay in cloud_in; the x86 reference then drops them because every comparison with NaN is false
    // and int(NaN) == INT_MIN fails the scanID range test (FE:78, :86, :98).  (int)NaN is 0 on the device, so drop them here.
    if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) return 255u;
    const float dxy = __fsqrt_rn(fadd(fmul(p.x, p.x), fmul(p.y, p.y)));  // DistanceXY, CM:59-62 (fp32)
    const double distance = (double)dxy;
    if (distance < cfg.lidar_min || distance > cfg.lidar_max) return 255u;  // FE:70
    if (cfg.n_scan == 0) {
      const int r = (int)L.ring_in[sel][i];
      return (r < cfg.n_rings) ? (uint32_t)r : 255u;
    }
    if (cfg.n_scan == 64) {
      // Certified fp32 fast path: the ring id is a step function of the vertical angle; an fp32 angle (error < 1e-5 deg)
      // decides it whenever the argument of every int() / comparison of FE:93-98 is farther than 1e-3 from a decision
      // point (99.8 % of the returns); the rest take the reference's fp64 expression below.
      const float af = atanf(__fdiv_rn(p.z, dxy)) * 57.29577951308232f;
      const float u = af >= -8.83f ? (2.0f - af) * 3.0f + 0.5f : (-8.83f - af) * 2.0f + 0.5f;
      const float fr = u - floorf(u);
      const bool safe = fr > 1e-3f && fr < 1.0f - 1e-3f && fabsf(af + 8.83f) > 1e-3f && fabsf(af - 2.0f) > 1e-3f && fabsf(af + 24.33f) > 1e-3f;
      if (safe) {
        if (af > 2.0f || af < -24.33f) return 255u;
        const int idf = af >= -8.83f ? (int)u : 32 + (int)u;
        return (idf > 63 || idf < 0) ? 255u : (uint32_t)idf;
      }
    }
    const double angle = atan((double)p.z / distance) * 180 / M_PI;  // FE:73
    int id = 0;
    if (cfg.n_scan == 16) {
      id = (int)((angle + 15) / 2 + 0.5);  // FE:77
      if (id > 15 || id < 0) return 255u;
    } else if (cfg.n_scan == 32) {
      id = (int)((angle + 92.0 / 3.0) * 3.0 / 4.0);  // FE:85
      if (id > 31 || id < 0) return 255u;
    } else if (cfg.n_scan == 64) {
      if (angle >= -8.83) id = (int)((2 - angle) * 3.0 + 0.5);  // FE:93-96
      else id = 64 / 2 + (int)((-8.83 - angle) * 2.0 + 0.5);
      if (angle > 2 || angle < -24.33 || id > 63 || id < 0) return 255u;  // FE:98
    } else {
      id = 0;  // FE:103-106: "wrong scan number", everything lands in ring 0
    }
    return (uint32_t)id;
  }
};

// ------------------------------------------------------------------------------------------------
// ring partition in one cluster kernel
// ------------------------------------------------------------------------------------------------
// getLaserCloud's "push_back to ring cloud r in arrival order" (FE:75-108) is a stable partition of the scan by an 8-bit
// key.  The generic radix machinery (k_sort_keyhist + k_sort_scatter: 512-digit tables, a 256-thread Hillis-Steele scan, an
// O(CTAs^2) histogram read, 160 instructions per point) is far more than this needs.  One 8-CTA cluster per sequence:
// every warp owns a contiguous slice of the scan, counts its ring ids (keys kept in global memory as one byte per point),
// the CTAs exchange their 256-bin histograms through distributed shared memory, and the warps then write the scan indices
// to their ring's range in arrival order.  Outputs: ring_sort.val[1] (indices in (ring, arrival) order), digit_start[257].
namespace cgx = cooperative_groups;
constexpr int RP_CL = 8, RP_CT = 512, RP_CW = RP_CT / 32;

struct RingShared {
  uint32_t wcnt[RP_CW][256];
  uint32_t cta_hist[256];
  uint32_t scan[RP_CW];
};

__global__ void __cluster_dims__(RP_CL, 1, 1) __launch_bounds__(RP_CT, 2) k_ring_partition(LaneDev* lanes, int lane0, int sel, ConfigDev cfg) {
  cgx::cluster_group cluster = cgx::this_cluster();
  const int rank = (int)cluster.block_rank();
  const LaneDev& L = lanes[lane0 + blockIdx.y];
  const int n = L.v->n_scan[sel];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ RingShared S;
  KeyGenRing gen;
  gen.lanes = lanes; gen.lane0 = lane0; gen.sel = sel; gen.cfg = cfg;
  uint8_t* keys = reinterpret_cast<uint8_t*>(L.ring_sort.key[0]);  // one byte per point (scratch of the generic sort job)
  uint32_t* __restrict__ perm = L.ring_sort.val[1];
  int cchunk = (n + RP_CL - 1) / RP_CL;
  cchunk = (cchunk + RP_CT - 1) / RP_CT * RP_CT;
  const int wchunk = cchunk / RP_CW;  // multiple of 32
  const int cbeg = min(n, rank * cchunk), cend = min(n, cbeg + cchunk);
  const int wbeg = min(cend, cbeg + warp * wchunk), wend = min(cend, wbeg + wchunk);
  for (int i = tid; i < RP_CW * 256; i += RP_CT) (&S.wcnt[0][0])[i] = 0;
  __syncthreads();
  constexpr int RB = 4;  // 32-point groups in flight per warp (every batch costs one exposed L2 round trip)
  for (int base = wbeg; base < wend; base += 32 * RB) {
    uint32_t k[RB];
#pragma unroll
    for (int it = 0; it < RB; ++it) {
      const int i = base + it * 32 + lane;
      k[it] = i < wend ? gen.key(blockIdx.y, i) : 0u;
    }
#pragma unroll
    for (int it = 0; it < RB; ++it) {
      const int i = base + it * 32 + lane;
      if (i < wend) { keys[i] = (uint8_t)k[it]; atomicAdd(&S.wcnt[warp][k[it]], 1u); }
    }
  }
  __syncthreads();
  if (tid < 256) {
    uint32_t s = 0;
#pragma unroll
    for (int w = 0; w < RP_CW; ++w) s += S.wcnt[w][tid];
    S.cta_hist[tid] = s;
  }
  cluster.sync();
  {  // ring start = rings below; this CTA's first slot in ring `tid` = ring start + same ring in lower-rank CTAs
    uint32_t tot = 0, pre = 0;
    if (tid < 256) {
#pragma unroll
      for (int c = 0; c < RP_CL; ++c) {
        const uint32_t v = cluster.map_shared_rank(&S, c)->cta_hist[tid];
        if (c < rank) pre += v;
        tot += v;
      }
    }
    // exclusive scan over the 256 rings (threads >= 256 contribute 0)
    uint32_t inc = tot;
    for (int off = 1; off < 32; off <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += t; }
    if (lane == 31) S.scan[warp] = inc;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < RP_CW; ++w) woff += w < warp ? S.scan[w] : 0u;
    const uint32_t start = woff + inc - tot;
    if (tid < 256) {
      if (rank == 0) {
        L.ring_sort.digit_start[tid] = start;
        if (tid == 255) L.ring_sort.digit_start[256] = start + tot;
      }
      uint32_t run = start + pre;
#pragma unroll
      for (int w = 0; w < RP_CW; ++w) {
        const uint32_t c = S.wcnt[w][tid];
        S.wcnt[w][tid] = run;
        run += c;
      }
    }
  }
  __syncthreads();
  for (int base = wbeg; base < wend; base += 32 * RB) {  // in arrival order, 32 points at a time, RB groups loaded ahead
    uint32_t dk[RB];
#pragma unroll
    for (int it = 0; it < RB; ++it) {
      const int i = base + it * 32 + lane;
      dk[it] = i < wend ? (uint32_t)keys[i] : 0u;
    }
#pragma unroll
    for (int it = 0; it < RB; ++it) {
      const int i = base + it * 32 + lane;
      const bool ok = i < wend;
      const uint32_t d = dk[it];
      const unsigned act = __ballot_sync(0xffffffffu, ok);
      unsigned peers = 0, lower = 0;
      uint32_t before = 0;
      if (ok) {
        peers = __match_any_sync(act, d);
        lower = peers & ((1u << lane) - 1u);
        before = S.wcnt[warp][d];
      }
      __syncwarp();
      if (ok) {
        if (lower == 0) S.wcnt[warp][d] = before + __popc(peers);
        perm[before + __popc(lower)] = (uint32_t)i;
      }
      __syncwarp();
    }
  }
  cluster.sync();  // no CTA may exit while a peer can still read its shared memory
}

// ------------------------------------------------------------------------------------------------
// per-(ring, sector) selection
// ------------------------------------------------------------------------------------------------
// dynamic shared memory of the selection kernel for sectors of up to ms elements (ms + 10 staged points; the sort
// network runs over np = the power of two >= ms)
static inline int sector_pow2(int ms) { int p = 64; while (p < ms) p <<= 1; return p; }
static inline size_t sector_smem(int ms) {
  const size_t pts = (size_t)ms + 10, np = (size_t)sector_pow2(ms);
  return sizeof(float4) * pts + sizeof(unsigned long long) * ms + sizeof(uint32_t) * pts + sizeof(uint16_t) * np + (pts + 15) / 16 * 16 + 16;
}

__device__ __forceinline__ bool sector_range(int n_r, int s, int& start, int& m) {
  if (n_r < 131) return false;                 // FE:179
  const int cloud_size = n_r - 10;             // FE:185, :205
  const int len = cloud_size / SECTORS;        // FE:208
  start = len * s;                             // FE:209
  int end = len * (s + 1) - 1;                 // FE:210
  if (s == SECTORS - 1) end = cloud_size - 1;  // FE:213
  m = end - start;                             // FE:215: half-open copy, element `end` is dropped
  return m > 0;
}

__global__ void __launch_bounds__(256) k_sector_select(LaneDev* lanes, int lane0, int sel, ConfigDev cfg) {
  const LaneDev& L = lanes[lane0 + blockIdx.y];
  const int r = blockIdx.x / SECTORS, s = blockIdx.x % SECTORS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t* rs = L.ring_sort.digit_start;
  const int ring_beg = (int)rs[r];
  const int n_r = (int)rs[r + 1] - ring_beg;
  int start = 0, m = 0;
  bool ok = sector_range(n_r, s, start, m);
  const int MS = cfg.max_sector;  // multiple of 8
  if (ok && m > MS) {
    if (tid == 0) atomicOr(&L.v->status, ST_SECTOR_TOO_LONG);
    ok = false;
  }
  if (!ok) {
    if (tid == 0) L.sec_cnt[blockIdx.x] = make_int2(0, 0);
    return;
  }
  extern __shared__ __align__(16) unsigned char smem[];
  float4* pts = reinterpret_cast<float4*>(smem);                              // ring points start .. start+m+9
  unsigned long long* val = reinterpret_cast<unsigned long long*>(smem + sizeof(float4) * (MS + 10));  // curvature bits of element e
  uint32_t* srcs = reinterpret_cast<uint32_t*>(val + MS);                     // scan index of each staged point
  uint16_t* sorted = reinterpret_cast<uint16_t*>(srcs + MS + 10);             // elements in ascending (curvature, index); NP entries
  uint8_t* picked = reinterpret_cast<uint8_t*>(sorted + cfg.sector_np);       // cloudNeighborPicked as flags

  const uint32_t* perm = L.ring_sort.val[1];  // scan indices in (ring, arrival) order after the single pass
  for (int k = tid; k < m + 10; k += 256) {
    const uint32_t src = perm[ring_beg + start + k];
    srcs[k] = src;
    pts[k] = L.scan[sel][src];
    picked[k] = 0;
  }
  __syncthreads();
  // FE:190-200: fp32 left-to-right 11-tap sums, squared in fp64.  Element e <-> ring index 5+start+e <-> pts[e+5].
  // The curvature is a non-negative double, so its bit pattern orders like the value: kept as a 64-bit integer key.
  for (int e = tid; e < m; e += 256) {
    const float4* p = pts + e + 5;
    float fx = fadd(fadd(fadd(fadd(p[-5].x, p[-4].x), p[-3].x), p[-2].x), p[-1].x);
    fx = fsub(fx, fmul(10.0f, p[0].x));
    fx = fadd(fadd(fadd(fadd(fadd(fx, p[1].x), p[2].x), p[3].x), p[4].x), p[5].x);
    float fy = fadd(fadd(fadd(fadd(p[-5].y, p[-4].y), p[-3].y), p[-2].y), p[-1].y);
    fy = fsub(fy, fmul(10.0f, p[0].y));
    fy = fadd(fadd(fadd(fadd(fadd(fy, p[1].y), p[2].y), p[3].y), p[4].y), p[5].y);
    float fz = fadd(fadd(fadd(fadd(p[-5].z, p[-4].z), p[-3].z), p[-2].z), p[-1].z);
    fz = fsub(fz, fmul(10.0f, p[0].z));
    fz = fadd(fadd(fadd(fadd(fadd(fz, p[1].z), p[2].z), p[3].z), p[4].z), p[5].z);
    const double dx = fx, dy = fy, dz = fz;
    const double c = dadd(dadd(dmul(dx, dx), dmul(dy, dy)), dmul(dz, dz));
    val[e] = (unsigned long long)__double_as_longlong(c == 0.0 ? 0.0 : c);  // -0.0 cannot occur (sum of squares); non-finite returns never reach a ring (KeyGenRing::key)
  }
  // std::sort ascending by curvature (FE:115); equal curvatures ordered by index (tie class T1 made canonical): a bitonic
  // network over the element INDICES (keys stay in place), NP = power of two >= m, padding entries sort to the end.
  // O(NP log^2 NP) compare-exchanges instead of the m^2 comparisons of a rank sort (profiles/r1b: 23 k warp instructions
  // per sector); steps with partner distance <= 32 stay inside one warp's 64-element block and need only __syncwarp.
  int NP = 64;
  while (NP < m) NP <<= 1;
  for (int e = tid; e < NP; e += 256) sorted[e] = e < m ? (uint16_t)e : (uint16_t)0xffffu;
  __syncthreads();
  for (int k = 2; k <= NP; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (NP >> 1); t += 256) {
        const int i = 2 * t - (t & (j - 1));
        const int l = i + j;
        const uint32_t ea = sorted[i], eb = sorted[l];
        const unsigned long long ka = ea == 0xffffu ? ~0ull : val[ea], kb = eb == 0xffffu ? ~0ull : val[eb];
        const bool gt = ka > kb || (ka == kb && ea > eb);
        if (gt == ((i & k) == 0)) { sorted[i] = (uint16_t)eb; sorted[l] = (uint16_t)ea; }
      }
      if (j > 32) __syncthreads(); else __syncwarp();
    }
    __syncthreads();
  }

  int n_edge = 0;
  if (warp == 0) {  // FE:120-163, serial in pick order, 32 candidates examined per step
    int i = m - 1;
    int cnt = 0;
    while (i >= 0) {
      const int pos = i - lane;
      int e = 0;
      bool un = false;
      if (pos >= 0) { e = sorted[pos]; un = picked[e + 5] == 0; }
      const unsigned msk = __ballot_sync(0xffffffffu, un);
      if (msk == 0) { i -= 32; continue; }
      const int l = __ffs(msk) - 1;
      const int esel = __shfl_sync(0xffffffffu, e, l);
      if (__longlong_as_double((long long)val[esel]) <= cfg.edge_threshold) break;  // FE:125
      ++cnt;                                        // FE:128
      const int li = esel + 5;
      if (lane == 0) picked[li] = 1;                // FE:129
      if (cnt > EDGES_PER_SECTOR) break;            // FE:131-136: the 21st is consumed, not emitted
      if (lane == 0) {
        L.sec_edge[blockIdx.x * EDGES_PER_SECTOR + cnt - 1] = pts[li];
        L.sec_edge_src[blockIdx.x * EDGES_PER_SECTOR + cnt - 1] = (int)srcs[li];
      }
      n_edge = cnt;
      bool stop = false;  // FE:138-160: gap^2 between consecutive ring points, fp32 difference squared in fp64
      if (lane < 5 || (lane >= 8 && lane < 13)) {
        const int k = lane < 5 ? lane + 1 : -(lane - 8 + 1);
        const float4 a = pts[li + k];
        const float4 b = lane < 5 ? pts[li + k - 1] : pts[li + k + 1];
        const double gx = fsub(a.x, b.x), gy = fsub(a.y, b.y), gz = fsub(a.z, b.z);
        stop = dadd(dadd(dmul(gx, gx), dmul(gy, gy)), dmul(gz, gz)) > 0.05;
      }
      const unsigned sm = __ballot_sync(0xffffffffu, stop);
      const unsigned f = sm & 0x1fu, bk = (sm >> 8) & 0x1fu;
      const int nf = f ? __ffs(f) - 1 : 5, nbk = bk ? __ffs(bk) - 1 : 5;
      if (lane < nf) picked[li + lane + 1] = 1;
      if (lane >= 8 && lane - 8 < nbk) picked[li - (lane - 8 + 1)] = 1;
      __syncwarp();
      i = i - l - 1;
    }
  }
  __syncthreads();
  // FE:165-172: everything not picked, in ascending-curvature order -> surf
  __shared__ int wsum[8];
  const int sbase = ring_beg + 5 + start;
  int run = 0;
  for (int base = 0; base < m; base += 256) {
    const int i = base + tid;
    int e = 0;
    bool keep = false;
    if (i < m) { e = sorted[i]; keep = picked[e + 5] == 0; }
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) wsum[warp] = __popc(b);
    __syncthreads();
    int off = 0, total = 0;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) { const int c = wsum[ww]; if (ww < warp) off += c; total += c; }
    if (keep) {
      const int dst = sbase + run + off + __popc(b & ((1u << lane) - 1u));
      L.sec_surf[dst] = pts[e + 5];
      L.sec_surf_src[dst] = (int)srcs[e + 5];
    }
    run += total;
    __syncthreads();
  }
  if (tid == 0) L.sec_cnt[blockIdx.x] = make_int2(n_edge, run);
}

// ------------------------------------------------------------------------------------------------
// warp-per-sector selection (sectors of up to 32 * EPL elements): no CTA barrier anywhere
// ------------------------------------------------------------------------------------------------
// The sort keys live in REGISTERS: lane l holds the EPL consecutive network positions l*EPL .. l*EPL+EPL-1 as ONE 32-bit
// word each (23-bit curvature prefix, 9-bit element index).  Bitonic steps with partner distance < EPL are register min / max
// pairs, larger distances are warp shuffles.  History (warp instructions per sector, the kernel is issue bound): rank sort
// in shared memory 23 k, bitonic network in shared memory 17 k, register network on (u64 key, index) with short-circuit
// compares 18 k, the same branch-free 10 k, 32-bit keys + tie fix-up ~6 k (profiles/r1d, r1e, r1h).  Packing (fp32 key << 32 |
// index) into one 64-bit word did not help: the selects dominate, not the compares.  The greedy pick and the surf
// compaction are warp-serial as before.
constexpr int SEC_WPC = 4;  // sectors (warps) per CTA
__host__ __device__ constexpr size_t sec_warp_bytes(int ms, int np) {  // shared memory of one warp: sectors of <= ms elements, network of np
  return ((sizeof(float4) * (ms + 10) + sizeof(double) * ms + sizeof(uint32_t) * (ms + 10) + sizeof(uint16_t) * np + (ms + 16)) + 15) / 16 * 16;
}

// 32-bit keys: (top 23 bits of the fp32-rounded curvature) << 9 | index.  Rounding and truncation are monotone, so two
// elements whose 23-bit prefixes differ are ordered exactly like their fp64 curvatures, and a compare-exchange is one
// min + one max (4-5 instructions instead of ~20 for the (u64, index) pair: this kernel is instruction-issue bound).
// Elements with EQUAL prefixes end up adjacent, ordered by index only; fix_tie_runs() then orders every such run by the
// exact (fp64 curvature, index) key with an odd-even transposition inside the run (runs are 2-3 elements long; roughly
// every second sector has one).
template <int EPL, int J>
__device__ __forceinline__ void k32_local_step(uint32_t (&v)[EPL], int lane, int k) {
#pragma unroll
  for (int r = 0; r < EPL; ++r)
    if ((r & J) == 0) {
      const uint32_t lo = min(v[r], v[r | J]), hi = max(v[r], v[r | J]);
      const bool asc = ((lane * EPL + r) & k) == 0;
      v[r] = asc ? lo : hi;
      v[r | J] = asc ? hi : lo;
    }
}
template <int EPL>
__device__ __forceinline__ void k32_sort(uint32_t (&v)[EPL], int lane) {
  constexpr int NP = 32 * EPL;
  for (int k = 2; k <= NP; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= EPL) {
        const int lj = j / EPL;
        const bool lower = (lane & lj) == 0;
#pragma unroll
        for (int q = 0; q < EPL; ++q) {
          const uint32_t o = __shfl_xor_sync(0xffffffffu, v[q], lj);
          const bool keep_min = (((lane * EPL + q) & k) == 0) == lower;
          v[q] = keep_min ? min(v[q], o) : max(v[q], o);
        }
      } else {
        if (EPL > 1 && j == 1) k32_local_step<EPL, 1>(v, lane, k);
        if (EPL > 2 && j == 2) k32_local_step<EPL, (EPL > 2 ? 2 : 1)>(v, lane, k);
        if (EPL > 4 && j == 4) k32_local_step<EPL, (EPL > 4 ? 4 : 1)>(v, lane, k);
        if (EPL > 8 && j == 8) k32_local_step<EPL, (EPL > 8 ? 8 : 1)>(v, lane, k);
        if (EPL > 16 && j == 16) k32_local_step<EPL, (EPL > 16 ? 16 : 1)>(v, lane, k);
      }
    }
  }
}
__device__ __forceinline__ uint32_t curv_prefix(double c) { return __float_as_uint(__double2float_rn(c)) >> 8; }  // 23 bits (sign is 0)

// sorted[0..m): elements ordered by (23-bit prefix, index).  Order every run of equal prefixes by (fp64 value, index).
// Returns false when a run is too long for the transposition passes (the caller then falls back to an exact rank sort).
// Ties are a handful of isolated pairs per sector: each lane remembers which of its positions p = lane + 32 k start a tied
// pair (p, p + 1) -- swapping two elements of one run does not change which pairs are tied -- and the odd-even
// transposition rounds visit only those, stopping after an odd and an even round in a row without a swap.
__device__ __forceinline__ bool fix_tie_runs(const double* val, uint16_t* sorted, int m, int lane) {
  constexpr int MAX_RUN = 8;
  unsigned tied = 0;  // bit k: positions lane + 32 k and lane + 32 k + 1 share the prefix
  bool too_long = false;
  for (int p = lane, k = 0; p + 1 < m; p += 32, ++k) {
    const uint32_t a = curv_prefix(val[sorted[p]]), b = curv_prefix(val[sorted[p + 1]]);
    if (a == b) {
      tied |= 1u << k;
      if (p + MAX_RUN < m && curv_prefix(val[sorted[p + MAX_RUN]]) == a) too_long = true;
    }
  }
  if (__any_sync(0xffffffffu, too_long)) return false;
  if (!__any_sync(0xffffffffu, tied != 0)) return true;
  bool prev_swapped = true;
  for (int round = 0; round < MAX_RUN; ++round) {  // runs of up to MAX_RUN elements are sorted after MAX_RUN rounds at the latest
    bool swapped = false;
    unsigned todo = ((lane ^ round) & 1) == 0 ? tied : 0u;  // the parity of p is the parity of the lane
    while (todo) {
      const int p = lane + 32 * (__ffs(todo) - 1);
      todo &= todo - 1;
      const int ea = sorted[p], eb = sorted[p + 1];
      const double va = val[ea], vb = val[eb];
      if ((va > vb) | ((va == vb) & (ea > eb))) { sorted[p] = (uint16_t)eb; sorted[p + 1] = (uint16_t)ea; swapped = true; }
    }
    __syncwarp();
    swapped = __any_sync(0xffffffffu, swapped);
    if (!swapped && !prev_swapped) break;
    prev_swapped = swapped;
  }
  return true;
}

// Exact fallback for a sector with a long run of equal prefixes: rank sort on the fp64 curvatures, (value, index) order.
__device__ __noinline__ void exact_rank_sort(const double* val, int m, int lane, uint16_t* sorted) {
  for (int e = lane; e < m; e += 32) {
    const double v = val[e];
    int rank = 0;
    for (int k = 0; k < m; ++k) {
      const double vk = val[k];
      rank += ((vk < v) | ((vk == v) & (k < e))) ? 1 : 0;
    }
    sorted[rank] = (uint16_t)e;
  }
}

template <int EPL>
__global__ void __launch_bounds__(32 * SEC_WPC) k_sector_warp(LaneDev* lanes, int lane0, int sel, ConfigDev cfg, int n_sectors) {
  const LaneDev& L = lanes[lane0 + blockIdx.y];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sec = blockIdx.x * SEC_WPC + warp;
  if (sec >= n_sectors) return;
  const int r = sec / SECTORS, s = sec % SECTORS;
  const uint32_t* rs = L.ring_sort.digit_start;
  const int ring_beg = (int)rs[r];
  const int n_r = (int)rs[r + 1] - ring_beg;
  int start = 0, m = 0;
  bool ok = sector_range(n_r, s, start, m);
  constexpr int NP = 32 * EPL;
  if (ok && m > min(cfg.max_sector, NP)) {
    if (lane == 0) atomicOr(&L.v->status, ST_SECTOR_TOO_LONG);
    ok = false;
  }
  if (!ok) {
    if (lane == 0) L.sec_cnt[sec] = make_int2(0, 0);
    return;
  }
  extern __shared__ __align__(16) unsigned char smem[];
  const int MS = min(cfg.max_sector, NP);  // multiple of 8
  unsigned char* base_sm = smem + (size_t)warp * sec_warp_bytes(MS, NP);
  float4* pts = reinterpret_cast<float4*>(base_sm);                        // ring points start .. start+m+9
  double* val = reinterpret_cast<double*>(base_sm + sizeof(float4) * (MS + 10));  // curvature of element e
  uint32_t* srcs = reinterpret_cast<uint32_t*>(val + MS);                  // scan index of each staged point
  uint16_t* sorted = reinterpret_cast<uint16_t*>(srcs + MS + 10);          // elements in ascending (curvature, index)
  uint8_t* picked = reinterpret_cast<uint8_t*>(sorted + NP);               // cloudNeighborPicked as flags

  const uint32_t* perm = L.ring_sort.val[1];  // scan indices in (ring, arrival) order after the partition
  for (int k0 = lane; k0 < m + 10; k0 += 32 * 6) {  // index, then point: two dependent loads; six of each in flight per lane
    uint32_t src[6];
    float4 pt[6];
#pragma unroll
    for (int u = 0; u < 6; ++u) src[u] = k0 + 32 * u < m + 10 ? perm[ring_beg + start + k0 + 32 * u] : 0u;
#pragma unroll
    for (int u = 0; u < 6; ++u) if (k0 + 32 * u < m + 10) pt[u] = L.scan[sel][src[u]];
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      const int k = k0 + 32 * u;
      if (k < m + 10) { srcs[k] = src[u]; pts[k] = pt[u]; picked[k] = 0; }
    }
  }
  __syncwarp();
  // FE:190-200: fp32 left-to-right 11-tap sums, squared in fp64.  Element e <-> ring index 5+start+e <-> pts[e+5].
  for (int e = lane; e < m; e += 32) {
    const float4* p = pts + e + 5;
    float fx = fadd(fadd(fadd(fadd(p[-5].x, p[-4].x), p[-3].x), p[-2].x), p[-1].x);
    fx = fsub(fx, fmul(10.0f, p[0].x));
    fx = fadd(fadd(fadd(fadd(fadd(fx, p[1].x), p[2].x), p[3].x), p[4].x), p[5].x);
    float fy = fadd(fadd(fadd(fadd(p[-5].y, p[-4].y), p[-3].y), p[-2].y), p[-1].y);
    fy = fsub(fy, fmul(10.0f, p[0].y));
    fy = fadd(fadd(fadd(fadd(fadd(fy, p[1].y), p[2].y), p[3].y), p[4].y), p[5].y);
    float fz = fadd(fadd(fadd(fadd(p[-5].z, p[-4].z), p[-3].z), p[-2].z), p[-1].z);
    fz = fsub(fz, fmul(10.0f, p[0].z));
    fz = fadd(fadd(fadd(fadd(fadd(fz, p[1].z), p[2].z), p[3].z), p[4].z), p[5].z);
    const double dx = fx, dy = fy, dz = fz;
    const double c = dadd(dadd(dmul(dx, dx), dmul(dy, dy)), dmul(dz, dz));
    val[e] = c == 0.0 ? 0.0 : c;  // (+0.0; non-finite returns never reach a ring, KeyGenRing::key)
  }
  __syncwarp();
  // std::sort ascending by curvature (FE:115); equal curvatures ordered by index (tie class T1 made canonical).
  static_assert(EPL <= 16, "the element index must fit the 9 low bits of the 32-bit sort key");
  uint32_t v[EPL];
#pragma unroll
  for (int q = 0; q < EPL; ++q) {
    const int p = lane * EPL + q;
    v[q] = p < m ? ((curv_prefix(val[p]) << 9) | (uint32_t)p) : 0xffffffffu;  // padding sorts behind the m elements
  }
  k32_sort<EPL>(v, lane);
#pragma unroll
  for (int q = 0; q < EPL; ++q) sorted[lane * EPL + q] = (uint16_t)(v[q] & 0x1ffu);  // (padding entries are never read: positions >= m)
  __syncwarp();
  if (!fix_tie_runs(val, sorted, m, lane)) {
    __syncwarp();
    exact_rank_sort(val, m, lane, sorted);
  }
  __syncwarp();

  int n_edge = 0;
  {  // FE:120-163, serial in pick order, 32 candidates examined per step
    int i = m - 1;
    int cnt = 0;
    while (i >= 0) {
      const int pos = i - lane;
      int e = 0;
      bool un = false;
      if (pos >= 0) { e = sorted[pos]; un = picked[e + 5] == 0; }
      const unsigned msk = __ballot_sync(0xffffffffu, un);
      if (msk == 0) { i -= 32; continue; }
      const int l = __ffs(msk) - 1;
      const int esel = __shfl_sync(0xffffffffu, e, l);
      if (val[esel] <= cfg.edge_threshold) break;  // FE:125
      ++cnt;                                        // FE:128
      const int li = esel + 5;
      if (lane == 0) picked[li] = 1;                // FE:129
      if (cnt > EDGES_PER_SECTOR) break;            // FE:131-136: the 21st is consumed, not emitted
      if (lane == 0) {
        L.sec_edge[sec * EDGES_PER_SECTOR + cnt - 1] = pts[li];
        L.sec_edge_src[sec * EDGES_PER_SECTOR + cnt - 1] = (int)srcs[li];
      }
      n_edge = cnt;
      bool stop = false;  // FE:138-160: gap^2 between consecutive ring points, fp32 difference squared in fp64
      if (lane < 5 || (lane >= 8 && lane < 13)) {
        const int k = lane < 5 ? lane + 1 : -(lane - 8 + 1);
        const float4 a = pts[li + k];
        const float4 b = lane < 5 ? pts[li + k - 1] : pts[li + k + 1];
        const double gx = fsub(a.x, b.x), gy = fsub(a.y, b.y), gz = fsub(a.z, b.z);
        stop = dadd(dadd(dmul(gx, gx), dmul(gy, gy)), dmul(gz, gz)) > 0.05;
      }
      const unsigned sm = __ballot_sync(0xffffffffu, stop);
      const unsigned f = sm & 0x1fu, bk = (sm >> 8) & 0x1fu;
      const int nf = f ? __ffs(f) - 1 : 5, nbk = bk ? __ffs(bk) - 1 : 5;
      if (lane < nf) picked[li + lane + 1] = 1;
      if (lane >= 8 && lane - 8 < nbk) picked[li - (lane - 8 + 1)] = 1;
      __syncwarp();
      i = i - l - 1;
    }
  }
  __syncwarp();
  // FE:165-172: everything not picked, in ascending-curvature order -> surf
  const int sbase = ring_beg + 5 + start;
  int run = 0;
  for (int base = 0; base < m; base += 32) {
    const int i = base + lane;
    int e = 0;
    bool keep = false;
    if (i < m) { e = sorted[i]; keep = picked[e + 5] == 0; }
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int dst = sbase + run + __popc(b & ((1u << lane) - 1u));
      L.sec_surf[dst] = pts[e + 5];
      L.sec_surf_src[dst] = (int)srcs[e + 5];
    }
    run += __popc(b);
  }
  if (lane == 0) L.sec_cnt[sec] = make_int2(n_edge, run);
}

template <int EPL>
static void launch_sector_warp(const Launch& L, LaneDev* lanes, int lane0, int nlanes, int sel, const ConfigDev& cfg) {
  constexpr int NP = 32 * EPL;
  const size_t PER_WARP = sec_warp_bytes(cfg.max_sector < NP ? cfg.max_sector : NP, NP);
  const int n_sectors = cfg.rings_total * SECTORS;
  dim3 g((n_sectors + SEC_WPC - 1) / SEC_WPC, nlanes);
  k_sector_warp<EPL><<<g, 32 * SEC_WPC, PER_WARP * SEC_WPC, L.st>>>(lanes, lane0, sel, cfg, n_sectors);
  L.tick(K_SECTOR);
}

// One CTA per ring: concatenates its 6 per-sector lists behind those of the lower rings = the reference's push_back
// order of cloud_Edge / cloud_Surf (FE:133, :170), and reduces getMinMax3D of both clouds for the voxel filter (EM:248-251).
__global__ void __launch_bounds__(256) k_compact_features(LaneDev* lanes, int lane0, ConfigDev cfg) {
  const LaneDev& L = lanes[lane0 + blockIdx.y];
  const int tid = threadIdx.x, lane = tid & 31;
  const int r = blockIdx.x, sec0 = r * SECTORS;
  __shared__ int s_off[2];
  __shared__ float bb_sm[7 * 8];
  if (tid < 32) {  // features emitted by the sectors of the lower rings
    int se = 0, ss = 0;
    for (int i = lane; i < sec0; i += 32) { const int2 c = L.sec_cnt[i]; se += c.x; ss += c.y; }
    for (int off = 16; off > 0; off >>= 1) { se += __shfl_xor_sync(0xffffffffu, se, off); ss += __shfl_xor_sync(0xffffffffu, ss, off); }
    if (lane == 0) { s_off[0] = se; s_off[1] = ss; }
  }
  __syncthreads();
  const int oe = s_off[0], os = s_off[1];
  const uint32_t* rs = L.ring_sort.digit_start;
  const int ring_beg = (int)rs[r];
  const int n_r = (int)rs[r + 1] - ring_beg;
  // the ring's 6 sector lists as ONE flat index range each (edges, surfs): a single round of loads instead of six
  __shared__ int s_cnt[2][SECTORS + 1], s_src[SECTORS];  // exclusive prefix of the per-sector counts; first staged surf element
  if (tid == 0) {
    int ae = 0, as = 0;
    for (int s = 0; s < SECTORS; ++s) {
      const int2 c = L.sec_cnt[sec0 + s];
      int start = 0, m = 0;
      sector_range(n_r, s, start, m);
      s_cnt[0][s] = ae; s_cnt[1][s] = as; s_src[s] = ring_beg + 5 + start;
      ae += c.x; as += c.y;
    }
    s_cnt[0][SECTORS] = ae; s_cnt[1][SECTORS] = as;
  }
  __syncthreads();
  float emn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, emx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  float smn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, smx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  int ecnt = 0, scnt = 0;
  const int te = s_cnt[0][SECTORS], ts = s_cnt[1][SECTORS];
  for (int k = tid; k < te; k += 256) {  // <= 120 edges
    int s = 0;
#pragma unroll
    for (int q = 1; q < SECTORS; ++q) s += k >= s_cnt[0][q] ? 1 : 0;
    const int src = (sec0 + s) * EDGES_PER_SECTOR + (k - s_cnt[0][s]);
    const float4 p = L.sec_edge[src];
    L.feat[0][oe + k] = p;
    L.feat_src[0][oe + k] = L.sec_edge_src[src];
    emn[0] = fminf(emn[0], p.x); emx[0] = fmaxf(emx[0], p.x); emn[1] = fminf(emn[1], p.y); emx[1] = fmaxf(emx[1], p.y);
    emn[2] = fminf(emn[2], p.z); emx[2] = fmaxf(emx[2], p.z); ++ecnt;
  }
  for (int k0 = tid; k0 < ts; k0 += 256 * 4) {  // four points in flight per thread (the loads are L2 round trips)
    float4 p[4];
    int si[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = k0 + 256 * u;
      if (k < ts) {
        int s = 0;
#pragma unroll
        for (int q = 1; q < SECTORS; ++q) s += k >= s_cnt[1][q] ? 1 : 0;
        const int src = s_src[s] + (k - s_cnt[1][s]);
        p[u] = L.sec_surf[src];
        si[u] = L.sec_surf_src[src];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = k0 + 256 * u;
      if (k < ts) {
        L.feat[1][os + k] = p[u];
        L.feat_src[1][os + k] = si[u];
        smn[0] = fminf(smn[0], p[u].x); smx[0] = fmaxf(smx[0], p[u].x); smn[1] = fminf(smn[1], p[u].y); smx[1] = fmaxf(smx[1], p[u].y);
        smn[2] = fminf(smn[2], p[u].z); smx[2] = fmaxf(smx[2], p[u].z); ++scnt;
      }
    }
  }
  bbox_commit(L.vv + 0, emn, emx, ecnt, bb_sm);
  bbox_commit(L.vv + 1, smn, smx, scnt, bb_sm);
  if (blockIdx.x == gridDim.x - 1 && tid == 0) { L.v->n_edge = oe + te; L.v->n_surf = os + ts; }
}

// Opt-in shared-memory sizes of the selection kernels.  Function attributes are per device: called once per context
// (vilf_create*), never cached in a process-wide flag.
cudaError_t init_extract_kernels() {
  cudaError_t e = cudaFuncSetAttribute(k_sector_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sector_smem(MAX_SECTOR));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sector_warp<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sec_warp_bytes(128, 128) * SEC_WPC));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sector_warp<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sec_warp_bytes(256, 256) * SEC_WPC));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_sector_warp<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sec_warp_bytes(512, 512) * SEC_WPC));
  return e;
}

void launch_extract(const Launch& L, LaneDev* lanes, const SortJob* ring_jobs, int lane0, int nlanes, int sel, const ConfigDev& cfg) {
  const size_t SEC_SMEM = sector_smem(cfg.max_sector);
  KeyGenRing gen;
  gen.lanes = lanes; gen.lane0 = lane0; gen.sel = sel; gen.cfg = cfg;
  if (cfg.flags_no_cluster) {  // generic radix machinery (also the reference implementation of the partition for the tests)
    dim3 gs(SORT_G, nlanes);
    k_sort_keyhist<KeyGenRing><<<gs, SORT_THREADS, 0, L.st>>>(ring_jobs + lane0, gen);
    L.tick(K_RING_KEYHIST);
    launch_sort_scatter(L, ring_jobs + lane0, nlanes, 0);
  } else {
    dim3 gr(RP_CL, nlanes);
    k_ring_partition<<<gr, RP_CT, 0, L.st>>>(lanes, lane0, sel, cfg);
    L.tick(K_RING_PARTITION);
  }
  if (cfg.max_sector <= 128) launch_sector_warp<4>(L, lanes, lane0, nlanes, sel, cfg);
  else if (cfg.max_sector <= 256) launch_sector_warp<8>(L, lanes, lane0, nlanes, sel, cfg);
  else if (cfg.max_sector <= 512) launch_sector_warp<16>(L, lanes, lane0, nlanes, sel, cfg);
  else {  // very long rings: one CTA per sector with the sort network in shared memory
    dim3 g2(cfg.rings_total * SECTORS, nlanes);
    k_sector_select<<<g2, 256, SEC_SMEM, L.st>>>(lanes, lane0, sel, cfg);
    L.tick(K_SECTOR);
  }
  dim3 g3(cfg.rings_total, nlanes);
  k_compact_features<<<g3, 256, 0, L.st>>>(lanes, lane0, cfg);
  L.tick(K_COMPACT);
}

}  // namespace vilf
