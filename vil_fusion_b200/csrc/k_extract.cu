// Stage 1 — featureExtraction::extractFeature (FE:223-232) on the device.
//
//   k_frame_reset        per-frame scratch reset + constant-velocity pose prediction (EM:238-243)
//   k_sort_keyhist<KeyGenRing>  getLaserCloud (FE:54-110): range gate + vertical-angle -> ring id, as the
//                        8-bit key of a single stable radix pass (arrival order kept inside a ring, FE:108)
//   k_sector_select      featureEdge_Surf + featureExtractionFromSector (FE:112-220): one CTA per
//                        (ring, sector): 11-tap fp32 curvature, rank-sort in shared memory, warp-serial
//                        greedy edge pick with +-5 neighbour suppression, surf = everything not picked
//   k_compact_features   concatenates the per-sector lists in (ring, sector) order = the reference's
//                        push_back order of cloud_Edge / cloud_Surf
#include "k_sort.cuh"

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
    const float dxy = __fsqrt_rn(fadd(fmul(p.x, p.x), fmul(p.y, p.y)));  // DistanceXY, CM:59-62 (fp32)
    const double distance = (double)dxy;
    if (distance < cfg.lidar_min || distance > cfg.lidar_max) return 255u;  // FE:70
    if (cfg.n_scan == 0) {
      const int r = (int)L.ring_in[sel][i];
      return (r < cfg.n_rings) ? (uint32_t)r : 255u;
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
// per-(ring, sector) selection
// ------------------------------------------------------------------------------------------------
// dynamic shared memory of the selection kernel for sectors of up to ms elements (ms + 10 staged points)
static inline size_t sector_smem(int ms) {
  const size_t pts = (size_t)ms + 10;
  return sizeof(float4) * pts + sizeof(double) * ms + sizeof(uint32_t) * pts + sizeof(uint16_t) * ms + (pts + 15) / 16 * 16 + 16;
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
  double* val = reinterpret_cast<double*>(smem + sizeof(float4) * (MS + 10)); // curvature of element e
  uint32_t* srcs = reinterpret_cast<uint32_t*>(val + MS);                     // scan index of each staged point
  uint16_t* sorted = reinterpret_cast<uint16_t*>(srcs + MS + 10);             // elements in ascending (curvature, index)
  uint8_t* picked = reinterpret_cast<uint8_t*>(sorted + MS);                  // cloudNeighborPicked as flags

  const uint32_t* perm = L.ring_sort.val[1];  // scan indices in (ring, arrival) order after the single pass
  for (int k = tid; k < m + 10; k += 256) {
    const uint32_t src = perm[ring_beg + start + k];
    srcs[k] = src;
    pts[k] = L.scan[sel][src];
    picked[k] = 0;
  }
  __syncthreads();
  // FE:190-200: fp32 left-to-right 11-tap sums, squared in fp64.  Element e <-> ring index 5+start+e <-> pts[e+5].
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
    val[e] = dadd(dadd(dmul(dx, dx), dmul(dy, dy)), dmul(dz, dz));
  }
  __syncthreads();
  // std::sort ascending by curvature (FE:115); equal curvatures ordered by index (tie class T1 made canonical).
  for (int e = tid; e < m; e += 256) {
    const double v = val[e];
    int rank = 0;
    for (int k = 0; k < m; ++k) {
      const double vk = val[k];
      rank += (vk < v || (vk == v && k < e)) ? 1 : 0;
    }
    sorted[rank] = (uint16_t)e;
  }
  __syncthreads();

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
      if (val[esel] <= cfg.edge_threshold) break;  // FE:125
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

__global__ void __launch_bounds__(256) k_compact_features(LaneDev* lanes, int lane0, ConfigDev cfg) {
  const LaneDev& L = lanes[lane0 + blockIdx.y];
  const int tid = threadIdx.x;
  __shared__ int red_e[256], red_s[256];
  int se = 0, ss = 0;
  for (int i = tid; i < (int)blockIdx.x; i += 256) { const int2 c = L.sec_cnt[i]; se += c.x; ss += c.y; }
  red_e[tid] = se; red_s[tid] = ss;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (tid < off) { red_e[tid] += red_e[tid + off]; red_s[tid] += red_s[tid + off]; }
    __syncthreads();
  }
  const int oe = red_e[0], os = red_s[0];
  const int2 c = L.sec_cnt[blockIdx.x];
  const int r = blockIdx.x / SECTORS, s = blockIdx.x % SECTORS;
  const uint32_t* rs = L.ring_sort.digit_start;
  const int ring_beg = (int)rs[r];
  int start = 0, m = 0;
  sector_range((int)rs[r + 1] - ring_beg, s, start, m);
  const int sbase = ring_beg + 5 + start;
  // copy + getMinMax3D of both feature clouds for the voxel filter that follows (EM:248-251)
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  int cnt = 0;
  for (int k = tid; k < c.x; k += 256) {
    const float4 p = L.sec_edge[blockIdx.x * EDGES_PER_SECTOR + k];
    L.feat[0][oe + k] = p;
    L.feat_src[0][oe + k] = L.sec_edge_src[blockIdx.x * EDGES_PER_SECTOR + k];
    mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x); mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
    mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z); ++cnt;
  }
  __shared__ float bb_sm[7 * 8];
  bbox_commit(L.vv + 0, mn, mx, cnt, bb_sm);
  for (int a = 0; a < 3; ++a) { mn[a] = FLT_MAX; mx[a] = -FLT_MAX; }
  cnt = 0;
  for (int k = tid; k < c.y; k += 256) {
    const float4 p = L.sec_surf[sbase + k];
    L.feat[1][os + k] = p;
    L.feat_src[1][os + k] = L.sec_surf_src[sbase + k];
    mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x); mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
    mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z); ++cnt;
  }
  bbox_commit(L.vv + 1, mn, mx, cnt, bb_sm);
  if (blockIdx.x == gridDim.x - 1 && tid == 0) { L.v->n_edge = oe + c.x; L.v->n_surf = os + c.y; }
}

void launch_extract(const Launch& L, LaneDev* lanes, const SortJob* ring_jobs, int lane0, int nlanes, int sel, const ConfigDev& cfg) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_sector_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sector_smem(MAX_SECTOR));
    attr_set = true;
  }
  const size_t SEC_SMEM = sector_smem(cfg.max_sector);
  KeyGenRing gen;
  gen.lanes = lanes; gen.lane0 = lane0; gen.sel = sel; gen.cfg = cfg;
  dim3 gs(SORT_G, nlanes);
  k_sort_keyhist<KeyGenRing><<<gs, SORT_THREADS, 0, L.st>>>(ring_jobs + lane0, gen);
  L.tick(K_RING_KEYHIST);
  launch_sort_scatter(L, ring_jobs + lane0, nlanes, 0);
  dim3 g2(cfg.rings_total * SECTORS, nlanes);
  k_sector_select<<<g2, 256, SEC_SMEM, L.st>>>(lanes, lane0, sel, cfg);
  L.tick(K_SECTOR);
  k_compact_features<<<g2, 256, 0, L.st>>>(lanes, lane0, cfg);
  L.tick(K_COMPACT);
}

}  // namespace vilf
