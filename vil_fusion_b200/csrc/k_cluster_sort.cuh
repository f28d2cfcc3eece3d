// Pieces shared by the one-cluster-per-cloud kernels (k_cluster.cu: voxel filter; k_cellmap.cu: sort of a frame's new map
// points): cluster geometry, the per-CTA shared block, a CTA scan and the stable LSD radix sort over distributed shared memory.
#pragma once
#include "k_voxel.cuh"
#include <cooperative_groups.h>

namespace vilf {

namespace cg = cooperative_groups;

#ifndef VILF_VOX_CLUSTER
#define VILF_VOX_CLUSTER 8
#endif
constexpr int CL = VILF_VOX_CLUSTER;  // CTAs per cluster (8 = portable maximum)
constexpr int CT = 512;            // threads per CTA (== SORT_RADIX: one thread per digit)
constexpr int CW = CT / 32;        // warps per CTA
constexpr int NB = 8;              // 32-key groups a warp loads ahead in the sort sweeps (every batch costs one exposed L2 round trip)

__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define PHASE_MARK(k) do { if (rank == 0 && tid == 0) J.vv->t[k] = gtimer(); } while (0)

struct VoxShared {
  uint32_t wcnt[CW][SORT_RADIX];   // per-warp digit counters (histogram, then running scatter offsets)
  uint32_t cta_hist[SORT_RADIX];   // this CTA's digit totals; read by the peers over DSMEM
  uint32_t scan[CW];
  float bb[8];                     // partial bounding box + count of this CTA
  int heads;                       // heads in this CTA's chunk
  int wsum[CW];
  uint32_t part;                   // grid build: this CTA's bucket-count total
};

// Exclusive scan of one value per thread over the CTA (CT threads); every thread gets the CTA total.  Ends synchronised.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* buf, uint32_t* total) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t inc = v;
  for (int off = 1; off < 32; off <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += t; }
  __syncthreads();  // buf may still be read from a previous call
  if (lane == 31) buf[warp] = inc;
  __syncthreads();
  uint32_t woff = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < CW; ++w) { const uint32_t c = buf[w]; if (w < warp) woff += c; tot += c; }
  if (total) *total = tot;
  return woff + inc - v;
}

// P stable LSD radix passes over (key, index) pairs of one cluster.  Warp `warp` of CTA `rank` owns elements [wbeg, wend) (contiguous,
// in order over the cluster); pass 0 materialises the keys with keyfn(i).  Result in sort.pair[P & 1].  Ends with a cluster barrier.
template <class KeyFn>
__device__ __forceinline__ void cluster_radix_sort(cg::cluster_group& cluster, VoxShared& S, const SortJob& sort, int wbeg, int wend, int bits, KeyFn keyfn) {
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int P = sort_passes(bits, sort.npass);
  const int w = sort_width(bits, P);
  const uint32_t mask = (1u << w) - 1u;
  const int nd_used = 1 << w;  // digits in use (<= SORT_RADIX)
  for (int pass = 0; pass < P; ++pass) {
    const int shift = pass * w;
    const uint2* pin = sort.pair[pass & 1];   // (key, value) pairs, moved with one 8-byte access each
    uint2* __restrict__ pout = sort.pair[(pass + 1) & 1];
    // sweep 1: warp-private digit histogram (pass 0 also materialises the keys); NB x 32 keys in flight per warp
    for (int i = tid; i < CW * SORT_RADIX; i += CT) (&S.wcnt[0][0])[i] = 0;
    __syncthreads();
    for (int base = wbeg; base < wend; base += 32 * NB) {
      uint32_t k[NB];
#pragma unroll
      for (int it = 0; it < NB; ++it) {
        const int i = base + it * 32 + lane;
        k[it] = 0;
        if (i < wend) {
          if (pass == 0) { k[it] = keyfn(i); sort.pair[0][i] = make_uint2(k[it], (uint32_t)i); }
          else k[it] = __ldcg(pin + i).x;  // written by other CTAs in the previous pass: L2, and no L1 allocation (see k_voxel.cuh)
        }
      }
#pragma unroll
      for (int it = 0; it < NB; ++it)
        if (base + it * 32 + lane < wend) atomicAdd(&S.wcnt[warp][(k[it] >> shift) & mask], 1u);
    }
    __syncthreads();
    {
      uint32_t s = 0;
      if (tid < nd_used) {
#pragma unroll
        for (int ww = 0; ww < CW; ++ww) s += S.wcnt[ww][tid];
      }
      S.cta_hist[tid] = s;
    }
    cluster.sync();
    // offsets of digit `tid`: exclusive scan over digits + same-digit counts of lower-rank CTAs and lower warps
    {
      uint32_t tot = 0, pre = 0;
      if (tid < nd_used) {
        for (int c = 0; c < CL; ++c) {
          const uint32_t v = cluster.map_shared_rank(&S, c)->cta_hist[tid];
          if (c < rank) pre += v;
          tot += v;
        }
      }
      uint32_t run = block_excl_scan(tot, S.scan, nullptr) + pre;
      if (tid < nd_used) {
#pragma unroll
        for (int ww = 0; ww < CW; ++ww) {
          const uint32_t c = S.wcnt[ww][tid];
          S.wcnt[ww][tid] = run;
          run += c;
        }
      }
    }
    __syncthreads();
    // sweep 2: stable rank inside the warp's sub-chunk, 32 keys at a time in order (NB x 32 loaded ahead), and scatter
    for (int base = wbeg; base < wend; base += 32 * NB) {
      uint32_t k[NB], v[NB];
#pragma unroll
      for (int it = 0; it < NB; ++it) {
        const int i = base + it * 32 + lane;
        k[it] = 0; v[it] = (uint32_t)i;
        if (i < wend) {
          const uint2 e = __ldcg(pin + i);
          k[it] = e.x; v[it] = e.y;
        }
      }
#pragma unroll
      for (int it = 0; it < NB; ++it) {
        const bool ok = base + it * 32 + lane < wend;
        const uint32_t d = (k[it] >> shift) & mask;
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
          const uint32_t pos = before + __popc(lower);
          pout[pos] = make_uint2(k[it], v[it]);
        }
        __syncwarp();
      }
    }
    cluster.sync();  // scattered pairs visible to the whole cluster; cta_hist may be overwritten
  }

}

}  // namespace vilf
