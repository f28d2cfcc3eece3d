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

// Centroids of the voxels whose first sorted element ("head") sits in this warp's 32 consecutive positions.
// PCL: centroid = Zero; centroid += point (input order inside the voxel, kept by the stable sort); centroid /= count.
// Short runs are summed by the head's own thread.  Long runs (dense scan returns falling into one voxel: up to ~1000
// points near the sensor) are walked by the whole warp: 32 coalesced (index, point) loads per step, prefetched one
// step ahead, staged in shared memory, then lanes 0..3 each add one component sequentially — the same fp32 order.
// Must be called by all 32 lanes; `stage` = 32 float4 of shared memory private to the warp.
// (Grid-wide path: the sorted pairs come from earlier kernels.  The cluster path has its own emitter, k_cluster.cu: emit_range.)
__device__ __forceinline__ void emit_centroids(const VoxJob& J, const uint32_t* __restrict__ key, const uint32_t* __restrict__ val, int n, int guard,
                                               int i, bool head, int dst, float4* stage) {
  constexpr int SHORT_RUN = 8;
  const int lane = threadIdx.x & 31;
  uint32_t k0 = 0;
  bool long_run = false;
  if (head && dst < J.cap_out) {
    k0 = key[i];
    if (!guard && i + SHORT_RUN < n && key[i + SHORT_RUN] == k0) {
      long_run = true;
    } else if (guard) {
      J.out[dst] = J.in[val[i]];  // PCL's guard path returns the input points untouched
    } else {
      float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
      int cnt = 0;
      for (int j = i; j < n && (j == i || key[j] == k0); ++j) {
        const float4 p = J.in[val[j]];
        c.x = fadd(c.x, p.x); c.y = fadd(c.y, p.y); c.z = fadd(c.z, p.z); c.w = fadd(c.w, p.w);
        ++cnt;
      }
      const float fc = (float)cnt;
      J.out[dst] = make_float4(__fdiv_rn(c.x, fc), __fdiv_rn(c.y, fc), __fdiv_rn(c.z, fc), __fdiv_rn(c.w, fc));
    }
  }
  unsigned lm = __ballot_sync(0xffffffffu, long_run);
  while (lm) {
    const int l = __ffs(lm) - 1;
    lm &= lm - 1;
    const int start = __shfl_sync(0xffffffffu, i, l);
    const uint32_t ks = __shfl_sync(0xffffffffu, k0, l);
    const int d = __shfl_sync(0xffffffffu, dst, l);
    float acc = 0.f;  // lane c < 4 accumulates component c
    int cnt = 0;
    int j = start + lane;
    bool valid = j < n && key[j] == ks;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) p = J.in[val[j]];
    for (;;) {
      const int len = __popc(__ballot_sync(0xffffffffu, valid));  // sorted keys: the valid lanes are a prefix
      stage[lane] = p;
      // prefetch the next 32 while this step's serial adds run
      bool valid2 = false;
      float4 p2 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (len == 32) {
        j += 32;
        valid2 = j < n && key[j] == ks;
        if (valid2) p2 = J.in[val[j]];
      }
      __syncwarp();
      if (lane < 4) {
        const float* s = reinterpret_cast<const float*>(stage) + lane;
        if (len == 32) {
          float v[32];
#pragma unroll
          for (int t = 0; t < 32; ++t) v[t] = s[4 * t];
#pragma unroll
          for (int t = 0; t < 32; ++t) acc = fadd(acc, v[t]);
        } else {
          for (int t = 0; t < len; ++t) acc = fadd(acc, s[4 * t]);
        }
      }
      __syncwarp();
      cnt += len;
      if (len < 32) break;
      valid = valid2;
      p = p2;
    }
    const float fc = (float)cnt;
    const float q = __fdiv_rn(acc, fc);
    const float qx = __shfl_sync(0xffffffffu, q, 0), qy = __shfl_sync(0xffffffffu, q, 1), qz = __shfl_sync(0xffffffffu, q, 2), qw = __shfl_sync(0xffffffffu, q, 3);
    if (lane == 0) J.out[d] = make_float4(qx, qy, qz, qw);
  }
}

}  // namespace vilf
