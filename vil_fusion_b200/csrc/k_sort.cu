// Radix-sort scatter kernel.  See k_sort.cuh.
#include "k_sort.cuh"

namespace vilf {

__global__ void __launch_bounds__(SORT_THREADS) k_sort_scatter(const SortJob* __restrict__ jobs, int pass) {
  const SortJob& J = jobs[blockIdx.y];
  const int n = *J.n;
  int chunk, geff;
  sort_geometry(n, chunk, geff);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (geff == 0 && blockIdx.x == 0 && pass == 0 && J.digit_start != nullptr) {  // empty input: all digit ranges are empty
    J.digit_start[tid] = 0;
    if (tid == 255) J.digit_start[256] = 0;
  }
  if ((int)blockIdx.x >= geff) return;
  const int bits = J.bits ? *J.bits : J.fixed_bits;
  const int P = sort_passes(bits, J.npass);
  if (pass >= P) return;
  const int w = sort_width(bits, P);
  const uint32_t mask = (1u << w) - 1u;
  const int shift = pass * w;
  const bool count_next = pass + 1 < P;
  const int shift_next = shift + w;

  __shared__ uint32_t base[SORT_RADIX];        // running global offset of this CTA per digit
  __shared__ uint32_t wcnt[8][SORT_RADIX];     // per-warp digit counters of the current tile
  __shared__ uint32_t scan_buf[SORT_THREADS];

  // 1. digit offsets of this CTA: exclusive scan of the digit totals + counts of the CTAs before it (digits 2t, 2t+1).
  const int nd = 1 << w;  // digits in use (<= SORT_RADIX): everything below touches only those
  uint32_t tot0 = 0, tot1 = 0, pre0 = 0, pre1 = 0;
  if (2 * tid < nd) {
    for (int b = 0; b < geff; ++b) {
      const uint2 v = *reinterpret_cast<const uint2*>(sort_hist(J, pass, b) + 2 * tid);
      if (b < (int)blockIdx.x) { pre0 += v.x; pre1 += v.y; }
      tot0 += v.x; tot1 += v.y;
    }
  }
  scan_buf[tid] = tot0 + tot1;
  __syncthreads();
  for (int off = 1; off < SORT_THREADS; off <<= 1) {  // Hillis-Steele inclusive scan over the 256 digit pairs
    const uint32_t add = tid >= off ? scan_buf[tid - off] : 0u;
    __syncthreads();
    scan_buf[tid] += add;
    __syncthreads();
  }
  const uint32_t excl0 = scan_buf[tid] - (tot0 + tot1), excl1 = excl0 + tot0;
  base[2 * tid] = excl0 + pre0;
  base[2 * tid + 1] = excl1 + pre1;
  if (pass == 0 && J.digit_start != nullptr && blockIdx.x == 0 && tid < 128) {  // 8-bit ring keys: digits 0..255
    J.digit_start[2 * tid] = excl0;
    J.digit_start[2 * tid + 1] = excl1;
    if (tid == 127) J.digit_start[256] = excl1 + tot1;
  }
  __syncthreads();

  const uint32_t* __restrict__ kin = J.key[pass & 1];
  const uint32_t* __restrict__ vin = J.val[pass & 1];
  uint32_t* __restrict__ kout = J.key[(pass + 1) & 1];
  uint32_t* __restrict__ vout = J.val[(pass + 1) & 1];
  const int beg = blockIdx.x * chunk;
  const int end = min(n, beg + chunk);

  for (int tile = beg; tile < end; tile += SORT_TILE) {
    for (int i = tid; i < 8 * nd; i += SORT_THREADS) wcnt[i >> w][i & (int)mask] = 0;
    __syncthreads();
    uint32_t k[4], v[4], r[4], d[4];
    bool ok[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {  // each warp ranks its own 128 consecutive keys, 32 at a time, in order
      const int i = tile + warp * 128 + it * 32 + lane;
      ok[it] = i < end;
      k[it] = 0; v[it] = 0; d[it] = 0; r[it] = 0;
      if (ok[it]) { k[it] = kin[i]; v[it] = vin[i]; d[it] = (k[it] >> shift) & mask; }
      const unsigned act = __ballot_sync(0xffffffffu, ok[it]);
      unsigned peers = 0, lower = 0;
      uint32_t before = 0;
      if (ok[it]) {
        peers = __match_any_sync(act, d[it]);
        lower = peers & ((1u << lane) - 1u);
        before = wcnt[warp][d[it]];
        r[it] = before + __popc(lower);
      }
      __syncwarp();
      if (ok[it] && lower == 0) wcnt[warp][d[it]] = before + __popc(peers);
      __syncwarp();
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // exclusive prefix over the 8 warps for digits tid and tid + 256, on top of the running base
      const int dg = tid + h * SORT_THREADS;
      if (dg >= nd) break;
      uint32_t run = base[dg];
#pragma unroll
      for (int ww = 0; ww < 8; ++ww) {
        const uint32_t c = wcnt[ww][dg];
        wcnt[ww][dg] = run;
        run += c;
      }
      base[dg] = run;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      if (ok[it]) {
        const uint32_t pos = wcnt[warp][d[it]] + r[it];
        kout[pos] = k[it];
        vout[pos] = v[it];
        if (count_next) atomicAdd(sort_hist(J, pass + 1, (int)(pos / (uint32_t)chunk)) + ((k[it] >> shift_next) & mask), 1u);
      }
    }
    __syncthreads();
  }
}

void launch_sort_scatter(const Launch& L, const SortJob* jobs_dev, int njobs, int pass) {
  dim3 grid(SORT_G, njobs);
  k_sort_scatter<<<grid, SORT_THREADS, 0, L.st>>>(jobs_dev, pass);
  L.tick(K_SORT_SCATTER);
}

}  // namespace vilf
