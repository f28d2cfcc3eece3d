// Stable LSD radix sort of (u32 key, u32 value) pairs, batched over jobs (blockIdx.y).
//
// Used three ways on the hot path: (1) one 8-bit pass on ring ids = the stable "push_back to the ring
// cloud in arrival order" of getLaserCloud (FE:108); (2) four passes on PCL voxel indices for
// pcl::VoxelGrid (stands in for the std::sort of cloud_point_index_idx in voxel_grid.hpp, EM:248-251,
// :347-350) — stability keeps the points of a voxel in input order, which fixes the fp32 summation order
// of the centroid (tie class T3 removed).
//
// Geometry: a job with n keys uses geff = ceil(n / chunk) CTAs, each owning one contiguous chunk
// (>= 2048 keys).  Pass = histogram kernel (per-CTA digit counts -> hist[cta][digit]) + scatter kernel
// (every CTA re-derives its digit offsets from the hist table, ranks its keys stably with
// __match_any_sync and per-warp counters, and writes the pairs to the other buffer).
// HBM/L2 traffic per pass: 2 x 8 B read + 8 B write per pair.
#pragma once
#include "vilf_internal.cuh"

namespace vilf {

__device__ __forceinline__ void sort_geometry(int n, int& chunk, int& geff) {
  int c = (n + SORT_G - 1) / SORT_G;
  if (c < SORT_MIN_CHUNK) c = SORT_MIN_CHUNK;
  c = (c + SORT_TILE - 1) / SORT_TILE * SORT_TILE;
  chunk = c;
  geff = (n + c - 1) / c;
}
__device__ __forceinline__ int sort_width(int bits, int npass) {
  int w = (bits + npass - 1) / npass;
  return w < 1 ? 1 : (w > 8 ? 8 : w);
}

// Key generators: prepare() is called by every thread of the CTA (may __syncthreads) and returns the
// number of significant key bits; key(i) produces the key of element i.  KeyGenNone reads stored keys.
struct KeyGenNone {};

// Histogram kernel of pass `pass`.  With a key generator (pass 0 only) it also materialises keys and
// the identity payload.
template <class KeyGen, bool GEN>
__global__ void __launch_bounds__(SORT_THREADS) k_sort_hist(const SortJob* __restrict__ jobs, int pass, KeyGen gen) {
  const SortJob& J = jobs[blockIdx.y];
  const int n = *J.n;
  int chunk, geff;
  sort_geometry(n, chunk, geff);
  if ((int)blockIdx.x >= geff) return;
  __shared__ uint32_t sh[256];
  sh[threadIdx.x] = 0;
  int bits;
  if constexpr (GEN) {
    bits = gen.prepare(blockIdx.y);
  } else {
    bits = J.bits ? *J.bits : J.fixed_bits;
  }
  __syncthreads();
  const int w = sort_width(bits, J.npass);
  const uint32_t mask = (1u << w) - 1u;
  const int shift = pass * w;
  const int beg = blockIdx.x * chunk;
  const int end = min(n, beg + chunk);
  uint32_t* kbuf = J.key[pass & 1];
  uint32_t* vbuf = J.val[pass & 1];
  for (int i = beg + threadIdx.x; i < end; i += SORT_THREADS) {
    uint32_t k;
    if constexpr (GEN) {
      k = gen.key(blockIdx.y, i);
      kbuf[i] = k;
      vbuf[i] = (uint32_t)i;
    } else {
      k = kbuf[i];
    }
    atomicAdd(&sh[(k >> shift) & mask], 1u);
  }
  __syncthreads();
  J.hist[blockIdx.x * 256 + threadIdx.x] = sh[threadIdx.x];
}

}  // namespace vilf
