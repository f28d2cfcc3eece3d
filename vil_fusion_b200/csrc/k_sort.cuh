// Stable LSD radix sort of (u32 key, u32 value) pairs, batched over jobs (blockIdx.y).
//
// Used two ways on the hot path: (1) one 8-bit pass on ring ids = the stable "push_back to the ring cloud in
// arrival order" of getLaserCloud (FE:108); (2) 3-4 passes on PCL voxel indices for pcl::VoxelGrid (stands in for
// the std::sort of cloud_point_index_idx in voxel_grid.hpp, EM:248-251, :347-350) — stability keeps the points
// of a voxel in input order, which fixes the fp32 summation order of the centroid (tie class T3 removed).
//
// Geometry: a job with n keys uses geff = ceil(n / chunk) CTAs, each owning one contiguous chunk (>= 2048 keys).
// Digits are up to 9 bits; P = ceil(bits / 9) passes of w = ceil(bits / P) bits.  Launch sequence:
//   k_sort_keyhist   key generator (ring id / voxel index) + pass-0 digit counts per CTA -> hist[0][cta][digit];
//                    also zeroes the tables of the later passes
//   k_sort_scatter x 4  (pass >= P exits at once): every CTA re-derives its digit offsets from hist[pass], ranks
//                    its keys stably (__match_any_sync + per-warp counters), writes the pairs to the other
//                    buffer and — fused — counts the NEXT pass's digit per destination CTA with L2 atomics, so no
//                    separate histogram kernel runs after pass 0.
// Result: key[P & 1], val[P & 1].  Traffic per pass: 8 B read + 8 B written per pair.
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
__device__ __forceinline__ int sort_passes(int bits, int max_pass) {
  int p = (bits + SORT_RADIX_BITS - 1) / SORT_RADIX_BITS;
  return p < 1 ? 1 : (p > max_pass ? max_pass : p);
}
__device__ __forceinline__ int sort_width(int bits, int passes) {
  int w = (bits + passes - 1) / passes;
  return w < 1 ? 1 : (w > SORT_RADIX_BITS ? SORT_RADIX_BITS : w);
}
__device__ __forceinline__ uint32_t* sort_hist(const SortJob& J, int pass, int cta) {
  return J.hist + ((size_t)pass * SORT_G + cta) * SORT_RADIX;
}

// Key generators: prepare() is called by every thread of the CTA and returns the number of significant key bits;
// key(job, i) produces the key of element i.
template <class KeyGen>
__global__ void __launch_bounds__(SORT_THREADS) k_sort_keyhist(const SortJob* __restrict__ jobs, KeyGen gen) {
  const SortJob& J = jobs[blockIdx.y];
  const int n = *J.n;
  int chunk, geff;
  sort_geometry(n, chunk, geff);
  if ((int)blockIdx.x >= geff) return;
  __shared__ uint32_t sh[SORT_RADIX];
  sh[threadIdx.x] = 0;
  sh[threadIdx.x + SORT_THREADS] = 0;
  const int bits = gen.prepare(blockIdx.y);
  __syncthreads();
  const int P = sort_passes(bits, J.npass);
  const int w = sort_width(bits, P);
  const uint32_t mask = (1u << w) - 1u;
  const int beg = blockIdx.x * chunk;
  const int end = min(n, beg + chunk);
  uint32_t* kbuf = J.key[0];
  uint32_t* vbuf = J.val[0];
  for (int i = beg + threadIdx.x; i < end; i += SORT_THREADS) {
    const uint32_t k = gen.key(blockIdx.y, i);
    kbuf[i] = k;
    vbuf[i] = (uint32_t)i;
    atomicAdd(&sh[k & mask], 1u);
  }
  __syncthreads();
  uint32_t* h0 = sort_hist(J, 0, blockIdx.x);
  h0[threadIdx.x] = sh[threadIdx.x];
  h0[threadIdx.x + SORT_THREADS] = sh[threadIdx.x + SORT_THREADS];
  for (int p = 1; p < P; ++p) {
    uint32_t* hp = sort_hist(J, p, blockIdx.x);
    hp[threadIdx.x] = 0;
    hp[threadIdx.x + SORT_THREADS] = 0;
  }
}

}  // namespace vilf
