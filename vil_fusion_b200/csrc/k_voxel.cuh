// Pieces of pcl::VoxelGrid / pcl::CropBox shared by the grid-wide path (k_voxel.cu, any size) and the
// one-cluster-per-cloud path (k_cluster.cu, clouds up to a few 1e5 points): voxel-index key generator and the
// centroid emitter.  Both paths therefore produce bit-identical clouds.
#pragma once
#include "k_sort.cuh"

namespace vilf {

__device__ __forceinline__ void crop_bounds(const VoxJob& J, float lo[3], float hi[3]) {
  for (int a = 0; a < 3; ++a) {
    if (J.crop == 2) { lo[a] = J.crop_lo[a]; hi[a] = J.crop_hi[a]; continue; }
    lo[a] = (float)dsub(J.crop_center[a], J.crop_half);  // EM:327-336: bounds in fp64, stored in an Eigen::Vector4f
    hi[a] = (float)dadd(J.crop_center[a], J.crop_half);
  }
}
__device__ __forceinline__ bool outside(const float4 p, const float lo[3], const float hi[3]) {  // pcl::CropBox: closed box
  return p.x < lo[0] || p.y < lo[1] || p.z < lo[2] || p.x > hi[0] || p.y > hi[1] || p.z > hi[2];
}

// PCL 1.7.2 voxel_grid.hpp: min_b = floor(min * inv), div_b = max_b - min_b + 1, the int32 guard, and the per-point
// index  ijk = (int)(floor(p * inv) - (float)min_b),  idx = i + j * dx + k * dx * dy  (all fp32).
struct KeyGenVoxel {
  const VoxJob* jobs;
  float inv;
  float lo[3], hi[3];
  int min_b[3], mul[3], total, guard, crop;
  const float4* in;

  // bounding box (of the points inside the crop box) -> grid; returns the significant bits of the keys 0..total
  __device__ int setup(const VoxJob& J, const float mn[3], const float mx[3], int n_valid, bool publish) {
    in = J.in;
    crop = J.crop;
    if (crop) crop_bounds(J, lo, hi);
    inv = 1.0f / J.leaf;  // inverse_leaf_size_ = Array4f::Ones() / leaf_size_.array()
    int bits = 1;
    guard = 0; total = 1;
    min_b[0] = min_b[1] = min_b[2] = 0; mul[0] = mul[1] = mul[2] = 0;
    int div_b[3] = {1, 1, 1};
    if (J.passthrough) {
      // pcl::CropBox::filter alone (test entry point): kept points get key 0, the rest the sentinel 1; the
      // stable sort then is an order-preserving compaction and every kept point is its own output.
      guard = 1;
    } else if (n_valid > 0) {
      const long long dx = (long long)(fmul(fsub(mx[0], mn[0]), inv)) + 1;
      const long long dy = (long long)(fmul(fsub(mx[1], mn[1]), inv)) + 1;
      const long long dz = (long long)(fmul(fsub(mx[2], mn[2]), inv)) + 1;
      if (dx * dy * dz > (long long)INT_MAX) {
        guard = 1;  // "Leaf size is too small for the input dataset": PCL returns the input cloud
        total = 1;
      } else {
        for (int a = 0; a < 3; ++a) {
          min_b[a] = (int)floorf(fmul(mn[a], inv));
          const int max_b = (int)floorf(fmul(mx[a], inv));
          div_b[a] = max_b - min_b[a] + 1;
        }
        mul[0] = 1; mul[1] = div_b[0]; mul[2] = div_b[0] * div_b[1];
        total = div_b[0] * div_b[1] * div_b[2];
      }
      bits = 32 - __clz(total);  // keys are 0..total (total = sentinel of cropped-out points)
    }
    if (publish) {
      VoxVars& W = *J.vv;
      for (int a = 0; a < 3; ++a) { W.min_b[a] = min_b[a]; W.div_b[a] = div_b[a]; }
      W.bits = bits; W.guard = guard; W.total = total;
    }
    return bits;
  }
  __device__ int prepare(int job) {  // grid path: the bounding box was reduced into VoxVars by an earlier kernel
    const VoxJob& J = jobs[job];
    const VoxVars& V = *J.vv;
    float mn[3], mx[3];
    for (int a = 0; a < 3; ++a) { mn[a] = ord2f(V.bbox[a]); mx[a] = ord2f(V.bbox[3 + a]); }
    return setup(J, mn, mx, V.n_valid, blockIdx.x == 0 && threadIdx.x == 0);
  }
  __device__ uint32_t key(int, int i) const {
    const float4 p = in[i];
    if (crop && outside(p, lo, hi)) return (uint32_t)total;
    if (guard) return 0u;
    const int i0 = (int)fsub(floorf(fmul(p.x, inv)), (float)min_b[0]);
    const int i1 = (int)fsub(floorf(fmul(p.y, inv)), (float)min_b[1]);
    const int i2 = (int)fsub(floorf(fmul(p.z, inv)), (float)min_b[2]);
    return (uint32_t)(i0 * mul[0] + i1 * mul[1] + i2 * mul[2]);
  }
};

// Sorted (voxel key, input index) pairs as the two paths store them.
struct KvPairs {   // cluster path: interleaved pairs written by other CTAs of the same kernel -> explicit ld.global.ca (see k_cluster.cu)
  const uint2* kv;
  __device__ __forceinline__ uint2 operator()(int i) const { return __ldca(kv + i); }
};
struct KvSplit {   // grid-wide path: separate key / value arrays written by earlier kernels
  const uint32_t* key;
  const uint32_t* val;
  __device__ __forceinline__ uint2 operator()(int i) const { return make_uint2(key[i], val[i]); }
};

// Centroids of the voxels whose first sorted element ("head") lies in [beg, end), by ONE warp, 32 sorted positions per
// step: coalesced (key, index) loads and the gathered points of the NEXT step are in flight while the current 32 points,
// staged in shared memory, are summed — every run by the lane at its first element, sequentially in input order, i.e.
// PCL's `centroid += point; ...; centroid /= count` in fp32 (same order as the grid-wide emitter => identical bits).  A run
// that is still open at the end of a step is carried into the next one (also past `end`: it still belongs to this warp).
// `dst0` = output index of the first run that starts in the range.  All loads are ld.global.ca (see k_voxel.cuh).
// centroid /= count.  A voxel with one point (most voxels of a local map) keeps the point: (0 + p) / 1 == p bit for bit, and
// skipping four IEEE divisions matters because this phase is instruction bound (150 warp instructions per 32 positions).
__device__ __forceinline__ float4 centroid_of(const float4 acc, int cnt) {
  if (cnt == 1) return acc;
  const float fc = (float)cnt;
  return make_float4(__fdiv_rn(acc.x, fc), __fdiv_rn(acc.y, fc), __fdiv_rn(acc.z, fc), __fdiv_rn(acc.w, fc));
}

template <class KV>
__device__ __forceinline__ void emit_range(const VoxJob& J, const KV kv, int nv, int guard, int beg, int end, int dst0, float4* stage) {
  if (beg >= end) return;  // uniform over the warp
  const int lane = threadIdx.x & 31;
  const unsigned FULL = 0xffffffffu;
  int dst = dst0;
  bool open = false;  // carried run (all of this is uniform over the warp)
  float4 cacc = make_float4(0.f, 0.f, 0.f, 0.f);
  int ccnt = 0, cdst = 0;
  uint32_t prevk = beg > 0 ? kv(beg - 1).x : 0u;
  uint32_t k = 0;
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
  bool valid = beg + lane < nv;
  if (valid) { const uint2 e = kv(beg + lane); k = e.x; p = __ldca(J.in + e.y); }
  // software pipeline, two dependent loads deep: the (key, index) pair of step s + 2 and the gathered point of step s + 1 are
  // in flight while step s is summed, so a step waits for one L2 round trip at most, not for the dependent pair
  bool nvalid = beg + 32 + lane < nv;
  uint2 ne = make_uint2(0u, 0u);
  if (nvalid) ne = kv(beg + 32 + lane);
  float4* const stage_base = stage;
  int buf = 0;
  for (int base = beg;; base += 32) {
    // The staged points go to the buffer the PREVIOUS step did not use, so one warp barrier per step suffices (the barrier of
    // step s + 1 orders the reads of step s before the writes of step s + 2).  The barrier also waits for this warp's
    // outstanding global loads and stores (measured: two L2 round trips per step when loads were issued before it and
    // stores sat before a second barrier), hence: stage, barrier, THEN issue the next step's loads, sum, store.
    stage = stage_base + 32 * buf;
    buf ^= 1;
    stage[lane] = p;
    __syncwarp();
    const uint32_t nk = ne.x;
    float4 np = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nvalid) np = __ldca(J.in + ne.y);
    const bool nnvalid = base + 64 + lane < nv;
    uint2 nne = make_uint2(0u, 0u);
    if (nnvalid) nne = kv(base + 64 + lane);
    // heads of this step
    uint32_t kl = __shfl_up_sync(FULL, k, 1);
    if (lane == 0) kl = prevk;
    const bool head = valid && (guard || base + lane == 0 || k != kl);
    const unsigned hm = __ballot_sync(FULL, head);
    const unsigned beyond = __ballot_sync(FULL, head && base + lane >= end);  // heads of the next warp's range
    const int nval = __popc(__ballot_sync(FULL, valid));                      // valid lanes are a prefix
    const int lim = min(beyond ? __ffs(beyond) - 1 : 32, nval);               // lanes [0, lim) are in play
    const unsigned starts = lim >= 32 ? hm : (hm & ((1u << lim) - 1u));
    const bool cont0 = open && lim > 0 && !(starts & 1u);                     // the carried run continues at lane 0
    if (open && !cont0) {                                                     // ... or it ended with the previous step
      if (lane == 0 && cdst < J.cap_out) J.out[cdst] = centroid_of(cacc, ccnt);
      open = false;
    }
    const unsigned all_starts = starts | (cont0 ? 1u : 0u);
    const bool is_start = lane < lim && ((all_starts >> lane) & 1u);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int cnt = 0, e = 0, my_dst = 0;
    if (all_starts == 1u && lim == 32) {
      // the whole step lies inside ONE run (dense returns near the sensor: hundreds of points per voxel): lanes 0..3 sum
      // one component each over the 32 staged points, loads first, then the dependent fp32 adds in input order
      float a = 0.f;
      if (lane < 4) {
        if (cont0) a = lane == 0 ? cacc.x : lane == 1 ? cacc.y : lane == 2 ? cacc.z : cacc.w;
        const float* sp = reinterpret_cast<const float*>(stage) + lane;
        float vv[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) vv[t] = sp[4 * t];
#pragma unroll
        for (int t = 0; t < 32; ++t) a = fadd(a, vv[t]);
      }
      acc.x = __shfl_sync(FULL, a, 0); acc.y = __shfl_sync(FULL, a, 1); acc.z = __shfl_sync(FULL, a, 2); acc.w = __shfl_sync(FULL, a, 3);
      e = 32;
      cnt = (cont0 ? ccnt : 0) + 32;
      my_dst = cont0 ? cdst : dst;
    } else if (is_start) {
      const unsigned higher = lane == 31 ? 0u : (all_starts & ~((2u << lane) - 1u));
      e = higher ? min(__ffs(higher) - 1, lim) : lim;  // exclusive end of my run inside this step
      if (lane == 0 && cont0) { acc = cacc; cnt = ccnt; my_dst = cdst; }
      else my_dst = dst + __popc(starts & ((1u << lane) - 1u));
      cnt += e - lane;
      int t = lane;
      for (; t + 4 <= e; t += 4) {
        const float4 q0 = stage[t], q1 = stage[t + 1], q2 = stage[t + 2], q3 = stage[t + 3];
        acc.x = fadd(fadd(fadd(fadd(acc.x, q0.x), q1.x), q2.x), q3.x);
        acc.y = fadd(fadd(fadd(fadd(acc.y, q0.y), q1.y), q2.y), q3.y);
        acc.z = fadd(fadd(fadd(fadd(acc.z, q0.z), q1.z), q2.z), q3.z);
        acc.w = fadd(fadd(fadd(fadd(acc.w, q0.w), q1.w), q2.w), q3.w);
      }
      for (; t < e; ++t) {
        const float4 q = stage[t];
        acc.x = fadd(acc.x, q.x); acc.y = fadd(acc.y, q.y); acc.z = fadd(acc.z, q.z); acc.w = fadd(acc.w, q.w);
      }
    }
    // a run reaching the end of a FULL step may continue: carry it; every other run is complete
    const bool carry = is_start && e == 32 && lim == 32;
    if (is_start && !carry && my_dst < J.cap_out) J.out[my_dst] = centroid_of(acc, cnt);
    dst += __popc(starts);
    if (lim < 32) break;
    const unsigned cm = __ballot_sync(FULL, carry);  // the last run of a full step, if this warp owns any run in it
    if (cm == 0u) {                                  // nothing owned yet (the step continues a previous warp's run)
      if (base + 32 >= end) break;                   // ... and no head of this range is left
      prevk = __shfl_sync(FULL, k, 31);
      k = nk; p = np; valid = nvalid; ne = nne; nvalid = nnvalid;
      continue;
    }
    const int cl = __ffs(cm) - 1;
    cacc.x = __shfl_sync(FULL, acc.x, cl); cacc.y = __shfl_sync(FULL, acc.y, cl); cacc.z = __shfl_sync(FULL, acc.z, cl); cacc.w = __shfl_sync(FULL, acc.w, cl);
    ccnt = __shfl_sync(FULL, cnt, cl);
    cdst = __shfl_sync(FULL, my_dst, cl);
    open = true;
    prevk = __shfl_sync(FULL, k, 31);
    k = nk; p = np; valid = nvalid; ne = nne; nvalid = nnvalid;
  }
}


}  // namespace vilf
