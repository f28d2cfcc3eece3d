// Ring-field / range-image feature extractor: the alternative stage 1 for sensors whose driver supplies ring ids, any beam count
// (SURVEY.md §8f rank 4).  Reference: class featureExtract, src/visual_inertial_lidar/feature_tracker/include/featureExtract.hpp
// (FX below): projectPointCloud FX:322-370, inverProjectCloud FX:293-318, extractSmoothness FX:268-290, markBadPoints FX:230-265,
// featureEdge_Surf FX:115-227.  Selected with VILF_FLAG_RANGE_IMAGE; it fills the same feature buffers as the ring-angle extractor
// (k_extract.cu), so everything downstream is unchanged.
//
//   k_ri_clear     image owners <- empty, per-ring counters <- 0
//   k_ri_project   range gate, ring / downsample test, azimuth -> column; the FIRST point to reach a cell keeps it (FX:360) = the lowest
//                  input index: atomicMin on the cell's owner; the thread that finds the cell empty counts it for its row
//   k_ri_fill      one CTA per ring: row-major compaction of the occupied cells behind those of the lower rings (= the order
//                  inverProjectCloud pushes them): input index, column, range
//   k_ri_smooth    range curvature over the FLATTENED cloud (11 taps, fp32, the reference's summation order) and the occlusion /
//                  parallel-beam marks; the marks are idempotent stores of 1, so the scatter needs no ordering
//   k_ri_select    one CTA per ring, its six sectors in sequence (a sector's surf marks can block edge candidates of the next one):
//                  bitonic sort of (curvature, position) keys in shared memory, then the reference's two greedy walks on one
//                  thread over shared-memory copies of the ring's marks / columns / curvatures, cut short where the sorted order
//                  proves that no later entry can qualify
//   k_ri_emit      one CTA per ring: edges and surfs behind those of the lower rings (the reference's push_back order), bounding
//                  boxes for the voxel filter
// What the reference's code does at its corners (position 4 of ring 0, the unsorted last position of a sector, the 21st candidate)
// is spelled out in the CPU restatement the tests compare with; this file follows the same rules.
#include "vilf_internal.cuh"

namespace vilf {

constexpr int RI_NP = 1024;          // sort network size: sectors of up to 1024 positions (Horizon_SCAN <= 6138)
constexpr int RI_MAX_H = 6138;
constexpr int RI_INFO = 4;           // per-ring ints: occupied cells, edges, surfs, spare

__global__ void __launch_bounds__(256) k_ri_clear(LaneDev* lanes, int lane0, ConfigDev cfg) {
  const LaneDev& L = lanes[lane0 + blockIdx.y];
  const int cells = cfg.n_rings * cfg.horizon;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < cells; i += gridDim.x * 256) L.ri_owner[i] = INT_MAX;
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < (MAX_RINGS + 1) * RI_INFO; i += 256) L.ri_info[i] = 0;
}

__device__ __forceinline__ float ri_range(const float4 p) { return __fsqrt_rn(fadd(fadd(fmul(p.x, p.x), fmul(p.y, p.y)), fmul(p.z, p.z))); }  // pointDistance, common.h:54-57

__global__ void __launch_bounds__(256) k_ri_project(LaneDev* lanes, int lane0, int sel, ConfigDev cfg) {
  const LaneDev& L = lanes[lane0 + blockIdx.y];
  const int n = min(L.v->n_scan[sel], cfg.cap_scan);
  const float4* __restrict__ scan = L.scan[sel];
  const uint16_t* __restrict__ ring = L.ring_in[sel];
  const int H = cfg.horizon, R = cfg.n_rings;
  const float ang_res_x = (float)(360.0 / (double)(float)H);  // FX:350
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const float4 p = scan[i];
    if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) continue;
    const float rxy = __fsqrt_rn(fadd(fmul(p.x, p.x), fmul(p.y, p.y)));  // DistanceXY
    if ((double)rxy < cfg.lidar_min || (double)rxy > cfg.lidar_max) continue;  // FX:337
    const int row = ring[i];
    if (row >= R) continue;                                                     // FX:342
    if (row % cfg.ri_down != 0) continue;                                       // FX:346
    // FX:349: float atan2 (computed in double and rounded: the correctly rounded float), rad2deg in double, stored as float
    const float ang = (float)(dmul((double)(float)atan2((double)p.x, (double)p.y), 180.0) / M_PI);
    int column = (int)(-round(dsub((double)ang, 90.0) / (double)ang_res_x) + (double)(H / 2));  // FX:352
    if (column >= H) column -= H;
    if (column < 0 || column >= H) continue;
    const int old = atomicMin(&L.ri_owner[row * H + column], i);
    if (old == INT_MAX) atomicAdd(&L.ri_info[row * RI_INFO], 1);
  }
}

__device__ __forceinline__ int ri_block_scan(int v, int* buf, int* total) {  // exclusive, 256 threads
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
  for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += t; }
  __syncthreads();
  if (lane == 31) buf[warp] = inc;
  __syncthreads();
  int woff = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) { const int c = buf[w]; if (w < warp) woff += c; tot += c; }
  *total = tot;
  return woff + inc - v;
}

__device__ __forceinline__ int ri_base_of(const int* info, int r) {  // image points of the rings below r (warp 0 of the CTA)
  int s = 0;
  for (int q = threadIdx.x & 31; q < r; q += 32) s += info[q * RI_INFO];
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  return s;
}

__global__ void __launch_bounds__(256) k_ri_fill(LaneDev* lanes, int lane0, int sel, ConfigDev cfg) {
  const LaneDev& L = lanes[lane0 + blockIdx.y];
  const int r = blockIdx.x, H = cfg.horizon, tid = threadIdx.x;
  __shared__ int s_base, s_buf[8];
  if (tid < 32) { const int b = ri_base_of(L.ri_info, r); if (tid == 0) s_base = b; }
  __syncthreads();
  int base = s_base;
  const float4* __restrict__ scan = L.scan[sel];
  for (int j0 = 0; j0 < H; j0 += 256) {
    const int j = j0 + tid;
    const int own = j < H ? L.ri_owner[r * H + j] : INT_MAX;
    const int has = own != INT_MAX ? 1 : 0;
    int tot;
    const int off = ri_block_scan(has, s_buf, &tot);
    if (has && base + off < cfg.cap_scan) {
      const int idx = base + off;
      L.ri_src[idx] = own; L.ri_col[idx] = j; L.ri_range[idx] = ri_range(scan[own]);
      L.ri_curv[idx] = 0.f; L.ri_picked[idx] = 0; L.ri_label[idx] = 0;
    }
    base += tot;
  }
}

__global__ void __launch_bounds__(256) k_ri_smooth(LaneDev* lanes, int lane0, ConfigDev cfg) {
  const LaneDev& L = lanes[lane0 + blockIdx.y];
  __shared__ int s_size;
  if (threadIdx.x < 32) { const int b = ri_base_of(L.ri_info, cfg.n_rings); if (threadIdx.x == 0) s_size = b; }
  __syncthreads();
  const int size = min(s_size, cfg.cap_scan);
  const float* __restrict__ rg = L.ri_range;
  const int* __restrict__ col = L.ri_col;
  for (int i = 5 + blockIdx.x * 256 + threadIdx.x; i < size - 5; i += gridDim.x * 256) {
    const float c = rg[i];
    // FX:273-275, left to right in fp32
    float d = fadd(rg[i - 5], rg[i - 4]); d = fadd(d, rg[i - 3]); d = fadd(d, rg[i - 2]); d = fadd(d, rg[i - 1]);
    d = fadd(d, rg[i + 5]); d = fadd(d, rg[i + 4]); d = fadd(d, rg[i + 3]); d = fadd(d, rg[i + 2]); d = fadd(d, rg[i + 1]);
    d = fsub(d, fmul(c, 10.f));
    L.ri_curv[i] = fmul(d, d);
    if (i < size - 6) {  // markBadPoints FX:233-263
      const float d2 = rg[i + 1];
      const int cdiff = abs(col[i + 1] - col[i]);
      if (cdiff < 10) {
        if ((double)fsub(c, d2) > 0.3) { for (int k = 0; k <= 5; ++k) L.ri_picked[i - k] = 1; }
        else if ((double)fsub(d2, c) > 0.3) { for (int k = 1; k <= 6; ++k) L.ri_picked[i + k] = 1; }
      }
      const float diff1 = fabsf(fsub(rg[i - 1], c)), diff2 = fabsf(fsub(d2, c));
      const double thr = dmul(0.02, (double)c);
      if ((double)diff1 > thr && (double)diff2 > thr) L.ri_picked[i] = 1;
    }
  }
}

// shared memory of k_ri_select: keys[RI_NP] u64 | curv[H + 2] f32 | col[H + 2] u16 | picked[H + 2] u8 | label[H + 2] u8
static size_t ri_select_smem(int H) { return (size_t)RI_NP * 8 + (size_t)(H + 2) * (4 + 2 + 1 + 1) + 16; }

__global__ void __launch_bounds__(256) k_ri_select(LaneDev* lanes, int lane0, ConfigDev cfg) {
  extern __shared__ __align__(16) unsigned char ri_sm[];
  const LaneDev& L = lanes[lane0 + blockIdx.y];
  const int r = blockIdx.x, H = cfg.horizon, tid = threadIdx.x;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(ri_sm);
  float* s_curv = reinterpret_cast<float*>(keys + RI_NP);
  unsigned short* s_col = reinterpret_cast<unsigned short*>(s_curv + (H + 2));
  unsigned char* s_picked = reinterpret_cast<unsigned char*>(s_col + (H + 2));
  unsigned char* s_label = s_picked + (H + 2);
  __shared__ int s_base, s_ne, s_ns;
  if (tid < 32) { const int b = ri_base_of(L.ri_info, r); if (tid == 0) { s_base = b; s_ne = 0; s_ns = 0; } }
  __syncthreads();
  const int base = s_base;
  const int cnt = min(L.ri_info[r * RI_INFO], max(0, cfg.cap_scan - base));
  const int lo = base - 1;  // shared copies hold flattened indices [base - 1, base + cnt): rel = index - lo
  for (int t = tid; t <= cnt; t += 256) {
    const int idx = lo + t;
    const bool ok = idx >= 0;
    s_curv[t] = ok ? L.ri_curv[idx] : 0.f;
    s_col[t] = ok ? (unsigned short)L.ri_col[idx] : (unsigned short)0;
    s_picked[t] = ok ? L.ri_picked[idx] : (unsigned char)0;
    s_label[t] = 0;
  }
  __syncthreads();
  const int start = base - 1 + 5, end = base + cnt - 1 - 5;  // startRingIndex / endRingIndex FX:299, :315
  const double ethr = cfg.ri_edge_thr, sthr = cfg.ri_surf_thr;
  for (int j = 0; j < 6; ++j) {
    const int sp = (start * (6 - j) + end * j) / 6;                 // FX:127-128
    const int ep = (start * (5 - j) + end * (j + 1)) / 6 - 1;
    if (sp >= ep) continue;                                         // (uniform over the CTA)
    const int len = ep - sp;  // sorted part [sp, ep)
    int np = 32;
    while (np < len) np <<= 1;
    if (np > RI_NP) { if (tid == 0) atomicOr(&L.v->status, ST_SECTOR_TOO_LONG); continue; }
    // entry of position k: {curvature[k], k}; position 4 of ring 0 was never written by extractSmoothness: {0, index 0}
    for (int t = tid; t < np; t += 256) {
      unsigned long long key = ~0ull;
      if (t < len) {
        const int k = sp + t;
        const int ind = k < 5 ? 0 : k;
        const float v = k < 5 ? 0.f : s_curv[k - lo];
        key = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned)(ind - lo);  // curvature >= 0: its bit pattern orders like the value
      }
      keys[t] = key;
    }
    __syncthreads();
    for (int k2 = 2; k2 <= np; k2 <<= 1)
      for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
        for (int t = tid; t < np; t += 256) {
          const int x = t ^ j2;
          if (x > t) {
            const unsigned long long a = keys[t], b = keys[x];
            const bool up = (t & k2) == 0;
            if ((a > b) == up) { keys[t] = b; keys[x] = a; }
          }
        }
        __syncthreads();
      }
    if (tid == 0) {
      auto mark = [&](int rel) {  // FX:160-176: +-5 neighbours until a column gap > 10
        for (int l = 1; l <= 5; ++l) {
          if (rel + l > cnt) break;
          if (abs((int)s_col[rel + l] - (int)s_col[rel + l - 1]) > 10) break;
          s_picked[rel + l] = 1;
        }
        for (int l = -1; l >= -5; --l) {
          if (lo + rel + l < 0) break;  // pointColInd[-1] in the reference (only for the stale entry of ring 0)
          if (abs((int)s_col[rel + l] - (int)s_col[rel + l + 1]) > 10) break;
          s_picked[rel + l] = 1;
        }
      };
      int ne = s_ne;
      int largest = 0;
      for (int k = ep; k >= sp; --k) {  // FX:138-177: position ep (never sorted) first, then descending curvature
        int rel;
        float v;
        if (k == ep) { rel = ep - lo; v = s_curv[rel]; }
        else { const unsigned long long e = keys[k - sp]; rel = (int)(unsigned)e; v = __uint_as_float((unsigned)(e >> 32)); if (!((double)v > ethr)) break; }
        // (the break is exact: sorted entries carry curvature[ind] itself, so none of the remaining ones can pass FX:142)
        if (s_picked[rel] == 0 && (double)s_curv[rel] > ethr) {
          ++largest;
          if (largest <= 20) {
            s_label[rel] = 1;
            const int slot = r * (SECTORS * EDGES_PER_SECTOR) + ne++;
            const int src = L.ri_src[lo + rel];
            L.sec_edge_src[slot] = src;
          } else break;
          s_picked[rel] = 1;
          mark(rel);
        }
      }
      s_ne = ne;
      for (int k = sp; k < ep; ++k) {   // FX:180-204: ascending curvature
        const unsigned long long e = keys[k - sp];
        const int rel = (int)(unsigned)e;
        if (!((double)__uint_as_float((unsigned)(e >> 32)) < sthr)) break;  // exact for the same reason
        if (s_picked[rel] == 0 && (double)s_curv[rel] < sthr) {
          s_label[rel] = 2;  // cloudLabel = -1
          s_picked[rel] = 1;
          mark(rel);
        }
      }
    }
    __syncthreads();
    // FX:207-211: every position of [sp, ep] that is not an edge goes to the surf cloud, in position order
    int mine = 0;
    for (int k = sp + tid; k <= ep; k += 256)
      if (s_label[k - lo] != 1) { s_label[k - lo] |= 4; ++mine; }
    for (int off = 16; off > 0; off >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, off);
    if ((tid & 31) == 0 && mine) atomicAdd(&s_ns, mine);
    __syncthreads();
  }
  for (int t = tid + 1; t <= cnt; t += 256) L.ri_label[lo + t] = s_label[t];
  if (tid == 0) { L.ri_info[r * RI_INFO + 1] = s_ne; L.ri_info[r * RI_INFO + 2] = s_ns; }
}

__global__ void __launch_bounds__(256) k_ri_emit(LaneDev* lanes, int lane0, int sel, ConfigDev cfg) {
  const LaneDev& L = lanes[lane0 + blockIdx.y];
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  __shared__ int s_off[3], s_buf[8];
  __shared__ float bb_sm[7 * 8];
  if (tid < 32) {
    int b = 0, e = 0, s = 0;
    for (int q = lane; q < r; q += 32) { b += L.ri_info[q * RI_INFO]; e += L.ri_info[q * RI_INFO + 1]; s += L.ri_info[q * RI_INFO + 2]; }
    for (int off = 16; off > 0; off >>= 1) { b += __shfl_xor_sync(0xffffffffu, b, off); e += __shfl_xor_sync(0xffffffffu, e, off); s += __shfl_xor_sync(0xffffffffu, s, off); }
    if (lane == 0) { s_off[0] = b; s_off[1] = e; s_off[2] = s; }
  }
  __syncthreads();
  const int base = s_off[0], oe = s_off[1];
  int os = s_off[2];
  const int cnt = min(L.ri_info[r * RI_INFO], max(0, cfg.cap_scan - base));
  const int ne = L.ri_info[r * RI_INFO + 1], ns = L.ri_info[r * RI_INFO + 2];
  const float4* __restrict__ scan = L.scan[sel];
  float emn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, emx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  float smn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, smx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  int ecnt = 0, scnt = 0;
  for (int k = tid; k < ne; k += 256) {
    const int src = L.sec_edge_src[r * (SECTORS * EDGES_PER_SECTOR) + k];
    const float4 p = scan[src];
    L.feat[0][oe + k] = p; L.feat_src[0][oe + k] = src;
    emn[0] = fminf(emn[0], p.x); emx[0] = fmaxf(emx[0], p.x); emn[1] = fminf(emn[1], p.y); emx[1] = fmaxf(emx[1], p.y);
    emn[2] = fminf(emn[2], p.z); emx[2] = fmaxf(emx[2], p.z); ++ecnt;
  }
  const int os0 = os;
  for (int t0 = 0; t0 < cnt; t0 += 256) {
    const int t = t0 + tid;
    const int idx = base + t;
    const int has = (t < cnt && (L.ri_label[idx] & 4)) ? 1 : 0;
    int tot;
    const int off = ri_block_scan(has, s_buf, &tot);
    if (has) {
      const int src = L.ri_src[idx];
      const float4 p = scan[src];
      L.feat[1][os + off] = p; L.feat_src[1][os + off] = src;
      smn[0] = fminf(smn[0], p.x); smx[0] = fmaxf(smx[0], p.x); smn[1] = fminf(smn[1], p.y); smx[1] = fmaxf(smx[1], p.y);
      smn[2] = fminf(smn[2], p.z); smx[2] = fmaxf(smx[2], p.z); ++scnt;
    }
    os += tot;
  }
  bbox_commit(L.vv + 0, emn, emx, ecnt, bb_sm);
  bbox_commit(L.vv + 1, smn, smx, scnt, bb_sm);
  if (r == gridDim.x - 1 && tid == 0) { L.v->n_edge = oe + ne; L.v->n_surf = os0 + ns; }
}

cudaError_t init_rangeimage_kernels() {
  return cudaFuncSetAttribute(k_ri_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ri_select_smem(RI_MAX_H));
}

void launch_extract_range_image(const Launch& L, LaneDev* lanes, int lane0, int nlanes, int sel, const ConfigDev& cfg) {
  dim3 g(148, nlanes), gr(cfg.n_rings, nlanes);
  k_ri_clear<<<g, 256, 0, L.st>>>(lanes, lane0, cfg);
  L.tick(K_RANGE_IMAGE);
  k_ri_project<<<g, 256, 0, L.st>>>(lanes, lane0, sel, cfg);
  L.tick(K_RANGE_IMAGE);
  k_ri_fill<<<gr, 256, 0, L.st>>>(lanes, lane0, sel, cfg);
  L.tick(K_RANGE_IMAGE);
  k_ri_smooth<<<g, 256, 0, L.st>>>(lanes, lane0, cfg);
  L.tick(K_RANGE_IMAGE);
  k_ri_select<<<gr, 256, ri_select_smem(cfg.horizon), L.st>>>(lanes, lane0, cfg);
  L.tick(K_RI_SELECT);
  k_ri_emit<<<gr, 256, 0, L.st>>>(lanes, lane0, sel, cfg);
  L.tick(K_RANGE_IMAGE);
}

}  // namespace vilf
