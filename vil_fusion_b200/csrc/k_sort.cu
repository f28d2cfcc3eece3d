// Radix-sort scatter kernel and the launchers of the generic (stored-key) passes.  See k_sort.cuh.
#include "k_sort.cuh"

namespace vilf {

__global__ void __launch_bounds__(SORT_THREADS) k_sort_scatter(const SortJob* __restrict__ jobs, int pass) {
  const SortJob& J = jobs[blockIdx.y];
  const int n = *J.n;
  int chunk, geff;
  sort_geometry(n, chunk, geff);
  if (geff == 0 && blockIdx.x == 0 && pass == 0 && J.digit_start != nullptr) {  // empty input: all digit ranges are empty
    J.digit_start[threadIdx.x] = 0;
    if (threadIdx.x == 255) J.digit_start[256] = 0;
  }
  if ((int)blockIdx.x >= geff) return;
  const int bits = J.bits ? *J.bits : J.fixed_bits;
  const int w = sort_width(bits, J.npass);
  const uint32_t mask = (1u << w) - 1u;
  const int shift = pass * w;
  const int nb = 1 << w;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  __shared__ uint32_t base[256];        // running global offset of this CTA per digit
  __shared__ uint32_t wcnt[8][256];     // per-warp digit counters of the current tile
  __shared__ uint32_t scan_buf[256];

  // 1. digit offsets of this CTA: exclusive scan of the digit totals + counts of the CTAs before it.
  uint32_t tot = 0, pre = 0;
  if (tid < nb) {
    for (int b = 0; b < geff; ++b) {
      uint32_t v = J.hist[b * 256 + tid];
      if (b < (int)blockIdx.x) pre += v;
      tot += v;
    }
  }
  scan_buf[tid] = tot;
  __syncthreads();
  for (int off = 1; off < 256; off <<= 1) {  // Hillis-Steele inclusive scan over 256 digits
    uint32_t add = tid >= off ? scan_buf[tid - off] : 0u;
    __syncthreads();
    scan_buf[tid] += add;
    __syncthreads();
  }
  const uint32_t excl = scan_buf[tid] - tot;
  base[tid] = excl + pre;
  if (pass == 0 && J.digit_start != nullptr && blockIdx.x == 0) {
    J.digit_start[tid] = excl;
    if (tid == 255) J.digit_start[256] = excl + tot;
  }
  __syncthreads();

  const uint32_t* __restrict__ kin = J.key[pass & 1];
  const uint32_t* __restrict__ vin = J.val[pass & 1];
  uint32_t* __restrict__ kout = J.key[(pass + 1) & 1];
  uint32_t* __restrict__ vout = J.val[(pass + 1) & 1];
  const int beg = blockIdx.x * chunk;
  const int end = min(n, beg + chunk);

  for (int tile = beg; tile < end; tile += SORT_TILE) {
    for (int i = tid; i < 8 * 256; i += SORT_THREADS) (&wcnt[0][0])[i] = 0;
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
    {  // exclusive prefix over the 8 warps for digit `tid`, on top of the running base
      uint32_t run = base[tid];
#pragma unroll
      for (int ww = 0; ww < 8; ++ww) {
        uint32_t c = wcnt[ww][tid];
        wcnt[ww][tid] = run;
        run += c;
      }
      base[tid] = run;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      if (ok[it]) {
        const uint32_t pos = wcnt[warp][d[it]] + r[it];
        kout[pos] = k[it];
        vout[pos] = v[it];
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

void launch_sort_pass(const Launch& L, const SortJob* jobs_dev, int njobs, int pass) {
  dim3 grid(SORT_G, njobs);
  k_sort_hist<KeyGenNone, false><<<grid, SORT_THREADS, 0, L.st>>>(jobs_dev, pass, KeyGenNone());
  L.tick(K_SORT_HIST);
  launch_sort_scatter(L, jobs_dev, njobs, pass);
}

}  // namespace vilf
