// Lidar depth for visual features — the caller-side stage next to the odometry path (SURVEY.md §8f / DESIGN.md §0 row f):
// getFeatureDepth steps 4.1-4.4 (feature_tracker_node.cpp:54-140) and the camera-frame preparation of the scan that
// precedes it (NODE:348-361).  The reference builds a pcl::KdTreeFLANN over ~5e4 unit-sphere points to answer ~150 3-NN
// queries; here the scan that is ALREADY resident in HBM for the odometry is filtered, transformed and normalised by one
// kernel, and every feature is answered by one CTA that streams the prepared points (coalesced float4, L2-resident) and
// keeps the three nearest in registers — exact by construction, no tree.
//
//   k_depth_cloud   NODE:351-361 + NODE:79-89: field-of-view test, pcl::transformPointCloud(LIDAR_CAMERA_EX) in fp64,
//                   range = sqrt(x^2+y^2+z^2) (fp32), unit vector; warp-aggregated compaction (the slot order is arbitrary,
//                   ties are broken by the original index, which orders like the reference's filtered cloud)
//   k_depth_query   NODE:61-75 + NODE:98-160: feature on the unit sphere, exact 3-NN with FLANN's fp32 L2_Simple, plane
//                   through the neighbours intersected with the feature ray, the reference's range sanity rules
#include "vilf_internal.cuh"

namespace vilf {

constexpr int DEPTH_THREADS = 256;

__global__ void __launch_bounds__(256) k_depth_cloud(const float4* __restrict__ in, const int* __restrict__ n_dev, int from_scan, const double* __restrict__ T,
                                                      float4* __restrict__ sph, int* __restrict__ oidx, int* __restrict__ count) {
  const int n = *n_dev;
  const int lane = threadIdx.x & 31;
  double t[12];
  if (from_scan) {
#pragma unroll
    for (int i = 0; i < 12; ++i) t[i] = T[i];
  }
  for (int base = blockIdx.x * 256; base < n; base += gridDim.x * 256) {
    const int i = base + threadIdx.x;
    bool keep = false;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) {
      p = in[i];
      keep = true;
      if (from_scan) {
        keep = p.x > 0.f && fabsf(__fdiv_rn(p.y, p.x)) <= 10.f && fabsf(__fdiv_rn(p.z, p.x)) <= 10.f;  // NODE:353
        const double x = p.x, y = p.y, z = p.z;  // PCL 1.7.2 transforms.hpp: left-to-right in the matrix scalar type (double)
        p.x = (float)dadd(dadd(dadd(dmul(t[0], x), dmul(t[1], y)), dmul(t[2], z)), t[3]);
        p.y = (float)dadd(dadd(dadd(dmul(t[4], x), dmul(t[5], y)), dmul(t[6], z)), t[7]);
        p.z = (float)dadd(dadd(dadd(dmul(t[8], x), dmul(t[9], y)), dmul(t[10], z)), t[11]);
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    int slot0 = 0;
    if (lane == 0 && m) slot0 = atomicAdd(count, __popc(m));
    slot0 = __shfl_sync(0xffffffffu, slot0, 0);
    if (keep) {
      const float range = __fsqrt_rn(fadd(fadd(fmul(p.x, p.x), fmul(p.y, p.y)), fmul(p.z, p.z)));  // pointDistance, common.h:54-57
      const int slot = slot0 + __popc(m & ((1u << lane) - 1u));
      sph[slot] = make_float4(__fdiv_rn(p.x, range), __fdiv_rn(p.y, range), __fdiv_rn(p.z, range), range);  // NODE:84-87
      oidx[slot] = i;
    }
  }
}

struct Top3 {
  float d[3];
  int id[3], slot[3];
};
__device__ __forceinline__ bool closer3(float d, int id, float bd, int bid) { return (d < bd) | ((d == bd) & (id < bid)); }
__device__ __forceinline__ void top3_insert(Top3& t, float cd, int ci, int cs) {  // precondition: (cd, ci) closer than t[2]
  t.d[2] = cd; t.id[2] = ci; t.slot[2] = cs;
#pragma unroll
  for (int k = 2; k > 0; --k) {
    const bool sw = closer3(t.d[k], t.id[k], t.d[k - 1], t.id[k - 1]);
    const float dk = sw ? t.d[k - 1] : t.d[k], dk1 = sw ? t.d[k] : t.d[k - 1];
    const int ik = sw ? t.id[k - 1] : t.id[k], ik1 = sw ? t.id[k] : t.id[k - 1];
    const int sk = sw ? t.slot[k - 1] : t.slot[k], sk1 = sw ? t.slot[k] : t.slot[k - 1];
    t.d[k] = dk; t.d[k - 1] = dk1; t.id[k] = ik; t.id[k - 1] = ik1; t.slot[k] = sk; t.slot[k - 1] = sk1;
  }
}

__global__ void __launch_bounds__(DEPTH_THREADS) k_depth_query(const float4* __restrict__ sph, const int* __restrict__ oidx, const int* __restrict__ count,
                                                                const float* __restrict__ feats, int m, float thr, float* __restrict__ depth_out,
                                                                int* __restrict__ nn_out) {
  const int f = blockIdx.x;
  if (f >= m) return;
  const int n = *count;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // 4.1: Eigen::Vector3f::normalize (NODE:64-66)
  float vx = feats[3 * f], vy = feats[3 * f + 1], vz = feats[3 * f + 2];
  const float nrm = __fsqrt_rn(fadd(fadd(fmul(vx, vx), fmul(vy, vy)), fmul(vz, vz)));
  vx = __fdiv_rn(vx, nrm); vy = __fdiv_rn(vy, nrm); vz = __fdiv_rn(vz, nrm);
  Top3 best;
#pragma unroll
  for (int k = 0; k < 3; ++k) { best.d[k] = FLT_MAX; best.id[k] = INT_MAX; best.slot[k] = -1; }
  for (int i = tid; i < n; i += DEPTH_THREADS) {
    const float4 c = __ldg(sph + i);
    const int id = __ldg(oidx + i);
    const float dx = fsub(vx, c.x), dy = fsub(vy, c.y), dz = fsub(vz, c.z);
    const float d = fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));  // FLANN L2_Simple<float>
    if (closer3(d, id, best.d[2], best.id[2])) top3_insert(best, d, id, i);
  }
  // merge the 256 lists: three rounds of a warp arg-min, then thread 0 merges the 8 warp results
  __shared__ float sd[8][3];
  __shared__ int sid[8][3], sslot[8][3];
#pragma unroll
  for (int round = 0; round < 3; ++round) {
    float wd = best.d[0];
    int wi = best.id[0], ws = best.slot[0];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, wd, off);
      const int oi = __shfl_xor_sync(0xffffffffu, wi, off);
      const int os = __shfl_xor_sync(0xffffffffu, ws, off);
      const bool take = closer3(od, oi, wd, wi);
      wd = take ? od : wd; wi = take ? oi : wi; ws = take ? os : ws;
    }
    if (lane == 0) { sd[warp][round] = wd; sid[warp][round] = wi; sslot[warp][round] = ws; }
    const bool pop = (best.id[0] == wi) & (wi != INT_MAX);
    best.d[0] = pop ? best.d[1] : best.d[0]; best.id[0] = pop ? best.id[1] : best.id[0]; best.slot[0] = pop ? best.slot[1] : best.slot[0];
    best.d[1] = pop ? best.d[2] : best.d[1]; best.id[1] = pop ? best.id[2] : best.id[1]; best.slot[1] = pop ? best.slot[2] : best.slot[1];
    best.d[2] = pop ? FLT_MAX : best.d[2]; best.id[2] = pop ? INT_MAX : best.id[2]; best.slot[2] = pop ? -1 : best.slot[2];
  }
  __syncthreads();
  if (tid != 0) return;
  Top3 r;
#pragma unroll
  for (int k = 0; k < 3; ++k) { r.d[k] = FLT_MAX; r.id[k] = INT_MAX; r.slot[k] = -1; }
  for (int w = 0; w < DEPTH_THREADS / 32; ++w)
    for (int k = 0; k < 3; ++k)
      if (sid[w][k] != INT_MAX && closer3(sd[w][k], sid[w][k], r.d[2], r.id[2])) top3_insert(r, sd[w][k], sid[w][k], sslot[w][k]);
  float depth = -1.0f;  // NODE:57-58
  if (nn_out) for (int k = 0; k < 3; ++k) nn_out[3 * f + k] = r.id[k] == INT_MAX ? -1 : r.id[k];
  if (n >= 10 && r.slot[2] >= 0 && r.d[2] < thr) {  // NODE:91-95, :107
    const float4 a = sph[r.slot[0]], b = sph[r.slot[1]], c = sph[r.slot[2]];
    const float r1 = a.w, r2 = b.w, r3 = c.w;
    const float A0 = fmul(a.x, r1), A1 = fmul(a.y, r1), A2 = fmul(a.z, r1);
    const float B0 = fmul(b.x, r2), B1 = fmul(b.y, r2), B2 = fmul(b.z, r2);
    const float C0 = fmul(c.x, r3), C1 = fmul(c.y, r3), C2 = fmul(c.z, r3);
    const float ab0 = fsub(A0, B0), ab1 = fsub(A1, B1), ab2 = fsub(A2, B2);
    const float bc0 = fsub(B0, C0), bc1 = fsub(B1, C1), bc2 = fsub(B2, C2);
    const float N0 = fsub(fmul(ab1, bc2), fmul(ab2, bc1)), N1 = fsub(fmul(ab2, bc0), fmul(ab0, bc2)), N2 = fsub(fmul(ab0, bc1), fmul(ab1, bc0));  // NODE:129
    float s = __fdiv_rn(fadd(fadd(fmul(N0, A0), fmul(N1, A1)), fmul(N2, A2)), fadd(fadd(fmul(N0, vx), fmul(N1, vy)), fmul(N2, vz)));  // NODE:130-131
    const float min_depth = fminf(r1, fminf(r2, r3)), max_depth = fmaxf(r1, fmaxf(r2, r3));
    bool ok = true;
    if (fsub(max_depth, min_depth) > 2.f || s <= 0.5f) ok = false;  // NODE:135-137
    else if (fsub(s, max_depth) > 0.f) s = max_depth;
    else if (fsub(s, min_depth) < 0.f) s = min_depth;
    if (ok) {
      const float z = fmul(vz, s);  // NODE:146-148
      if (z > 2.0f) depth = z;      // NODE:155-159
    }
  }
  depth_out[f] = depth;
}

void launch_depth(const Launch& L, const float4* in, const int* n_dev, int from_scan, const double* T_dev, float4* sph, int* oidx, int* count,
                  const float* feats_dev, int m, float thr, float* depth_dev, int* nn_dev) {
  cudaMemsetAsync(count, 0, sizeof(int), L.st);
  k_depth_cloud<<<148 * 2, 256, 0, L.st>>>(in, n_dev, from_scan, T_dev, sph, oidx, count);
  L.tick(K_DEPTH_CLOUD);
  if (m > 0) {
    k_depth_query<<<m, DEPTH_THREADS, 0, L.st>>>(sph, oidx, count, feats_dev, m, thr, depth_dev, nn_dev);
    L.tick(K_DEPTH_QUERY);
  }
}

}  // namespace vilf
