// pcl::VoxelGrid<PointXYZI>::applyFilter (PCL 1.7.2 filters/impl/voxel_grid.hpp; call sites EM:248-251 for
// the scan features and EM:347-350 for the local maps) and pcl::CropBox (EM:335-344), batched over jobs.
//
//   k_vox_bbox      crop-box test (closed AABB, bounds cast to fp32) + getMinMax3D over the kept points
//   k_sort_hist<KeyGenVoxel>  min_b/div_b, PCL's int32 guard, fp32 voxel index per point (cropped-out points get
//                   the sentinel key dx*dy*dz and sort to the tail), pass-0 digit histogram
//   3 x (hist, scatter) + scatter of pass 0     stable radix sort by voxel index (k_sort.cu)
//   k_vox_heads     first point of every occupied voxel, counted per CTA chunk
//   k_vox_centroid  one thread per occupied voxel walks its points in input order and accumulates
//                   x, y, z, intensity in fp32 exactly like PCL's `centroid += ...; centroid /= count`;
//                   output order = ascending voxel index = PCL's output order
//   k_map_append    createSubMap step 1 (EM:308-324): world-transform the filtered scan features and append
//   k_map_init      localMapInited (EM:105-115)
//
// Algorithmic bytes per point: 16 B read + 16 B/voxel written; the sort's (key,index) traffic
// (4 passes x 24 B) is implementation overhead.
#include "k_sort.cuh"

namespace vilf {

__device__ __forceinline__ bool in_crop(const VoxJob& J, const float4 p, const float lo[3], const float hi[3]) {
  return !(p.x < lo[0] || p.y < lo[1] || p.z < lo[2] || p.x > hi[0] || p.y > hi[1] || p.z > hi[2]);
}
__device__ __forceinline__ void crop_bounds(const VoxJob& J, float lo[3], float hi[3]) {
  for (int a = 0; a < 3; ++a) {
    if (J.crop == 2) { lo[a] = J.crop_lo[a]; hi[a] = J.crop_hi[a]; continue; }
    lo[a] = (float)dsub(J.crop_center[a], J.crop_half);  // EM:327-336: bounds in fp64, stored in an Eigen::Vector4f
    hi[a] = (float)dadd(J.crop_center[a], J.crop_half);
  }
}

__global__ void __launch_bounds__(256) k_vox_bbox(const VoxJob* __restrict__ jobs) {
  const VoxJob& J = jobs[blockIdx.y];
  const int n = *J.n_in;
  float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  if (J.crop) crop_bounds(J, lo, hi);
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  int cnt = 0;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const float4 p = J.in[i];
    if (J.crop && !in_crop(J, p, lo, hi)) continue;
    ++cnt;
    mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
    mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
    mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
  }
  for (int off = 16; off > 0; off >>= 1) {
    for (int a = 0; a < 3; ++a) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], off));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], off));
    }
    cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
  }
  if ((threadIdx.x & 31) == 0 && cnt > 0) {
    for (int a = 0; a < 3; ++a) {
      atomicMin(&J.vv->bbox[a], f2ord(mn[a]));
      atomicMax(&J.vv->bbox[3 + a], f2ord(mx[a]));
    }
    atomicAdd(&J.vv->n_valid, cnt);
  }
}

struct KeyGenVoxel {
  const VoxJob* jobs;
  // per-CTA state (set by prepare)
  float inv;
  float lo[3], hi[3];
  int min_b[3], mul[3], total, guard, crop;
  const float4* in;

  __device__ int prepare(int job) {
    const VoxJob& J = jobs[job];
    const VoxVars& V = *J.vv;
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
    } else if (V.n_valid > 0) {
      float mn[3], mx[3];
      for (int a = 0; a < 3; ++a) { mn[a] = ord2f(V.bbox[a]); mx[a] = ord2f(V.bbox[3 + a]); }
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
      bits = 32 - __clz(total);  // keys are 0..total (total = sentinel)
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      VoxVars& W = *J.vv;
      for (int a = 0; a < 3; ++a) { W.min_b[a] = min_b[a]; W.div_b[a] = div_b[a]; }
      W.bits = bits; W.guard = guard; W.total = total;
    }
    return bits;
  }
  __device__ uint32_t key(int, int i) const {
    const float4 p = in[i];
    if (crop && (p.x < lo[0] || p.y < lo[1] || p.z < lo[2] || p.x > hi[0] || p.y > hi[1] || p.z > hi[2])) return (uint32_t)total;
    if (guard) return 0u;
    const int i0 = (int)fsub(floorf(fmul(p.x, inv)), (float)min_b[0]);
    const int i1 = (int)fsub(floorf(fmul(p.y, inv)), (float)min_b[1]);
    const int i2 = (int)fsub(floorf(fmul(p.z, inv)), (float)min_b[2]);
    return (uint32_t)(i0 * mul[0] + i1 * mul[1] + i2 * mul[2]);
  }
};

__device__ __forceinline__ void vox_chunk(int n, int b, int& beg, int& end) {
  int chunk = (n + VOX_G - 1) / VOX_G;
  chunk = (chunk + 255) / 256 * 256;
  beg = min(n, b * chunk);
  end = min(n, beg + chunk);
}

__global__ void __launch_bounds__(256) k_vox_heads(const VoxJob* __restrict__ jobs) {
  const VoxJob& J = jobs[blockIdx.y];
  const int n = J.vv->n_valid;
  const int guard = J.vv->guard;
  const uint32_t* key = J.sort.key[0];
  int beg, end;
  vox_chunk(n, blockIdx.x, beg, end);
  int cnt = 0;
  for (int i = beg + threadIdx.x; i < end; i += 256) cnt += (guard || i == 0 || key[i] != key[i - 1]) ? 1 : 0;
  __shared__ int red[8];
  for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < 8; ++w) s += red[w];
    J.head_cnt[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(256) k_vox_centroid(const VoxJob* __restrict__ jobs) {
  const VoxJob& J = jobs[blockIdx.y];
  const int n = J.vv->n_valid;
  const int guard = J.vv->guard;
  const uint32_t* __restrict__ key = J.sort.key[0];
  const uint32_t* __restrict__ val = J.sort.val[0];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ int red[256];
  __shared__ int wsum[8];
  // heads in the chunks before this CTA, and in total
  int pre = 0, tot = 0;
  for (int b = tid; b < VOX_G; b += 256) { const int c = J.head_cnt[b]; if (b < (int)blockIdx.x) pre += c; tot += c; }
  red[tid] = pre;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) { if (tid < off) red[tid] += red[tid + off]; __syncthreads(); }
  pre = red[0];
  __syncthreads();
  if (blockIdx.x == 0) {
    red[tid] = tot;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) { if (tid < off) red[tid] += red[tid + off]; __syncthreads(); }
    if (tid == 0) {
      int total = red[0];
      if (total > J.cap_out) { atomicOr(J.status, ST_MAP_CAPACITY); total = J.cap_out; }
      *J.n_out = total;
    }
    __syncthreads();
  }
  int beg, end;
  vox_chunk(n, blockIdx.x, beg, end);
  int run = pre;
  for (int base = beg; base < end; base += 256) {
    const int i = base + tid;
    const bool head = i < end && (guard || i == 0 || key[i] != key[i - 1]);
    const unsigned b = __ballot_sync(0xffffffffu, head);
    if (lane == 0) wsum[warp] = __popc(b);
    __syncthreads();
    int off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { const int c = wsum[w]; if (w < warp) off += c; total += c; }
    if (head) {
      const int dst = run + off + __popc(b & ((1u << lane) - 1u));
      if (dst < J.cap_out) {
        const uint32_t k0 = key[i];
        float4 c = J.in[val[i]];
        int cnt = 1;
        if (!guard) {
          for (int j = i + 1; j < n && key[j] == k0; ++j) {  // input order inside the voxel (stable sort)
            const float4 p = J.in[val[j]];
            c.x = fadd(c.x, p.x); c.y = fadd(c.y, p.y); c.z = fadd(c.z, p.z); c.w = fadd(c.w, p.w);
            ++cnt;
          }
          const float fc = (float)cnt;
          c.x = __fdiv_rn(c.x, fc); c.y = __fdiv_rn(c.y, fc); c.z = __fdiv_rn(c.z, fc); c.w = __fdiv_rn(c.w, fc);
        }
        J.out[dst] = c;
      }
    }
    run += total;
    __syncthreads();
  }
}

void launch_voxel(const Launch& L, const VoxJob* jobs_dev, int njobs, const SortJob* sort_jobs_dev) {
  dim3 gv(VOX_G, njobs);
  k_vox_bbox<<<gv, 256, 0, L.st>>>(jobs_dev);
  L.tick(K_VOX_BBOX);
  KeyGenVoxel gen;
  gen.jobs = jobs_dev;
  dim3 gs(SORT_G, njobs);
  k_sort_hist<KeyGenVoxel, true><<<gs, SORT_THREADS, 0, L.st>>>(sort_jobs_dev, 0, gen);
  L.tick(K_VOX_KEYHIST);
  launch_sort_scatter(L, sort_jobs_dev, njobs, 0);
  for (int pass = 1; pass < 4; ++pass) launch_sort_pass(L, sort_jobs_dev, njobs, pass);
  k_vox_heads<<<gv, 256, 0, L.st>>>(jobs_dev);
  L.tick(K_VOX_HEADS);
  k_vox_centroid<<<gv, 256, 0, L.st>>>(jobs_dev);
  L.tick(K_VOX_CENTROID);
}

// ------------------------------------------------------------------------------------------------
// map maintenance helpers
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_map_append(LaneDev* lanes, int lane0, int cur, ConfigDev cfg) {
  const LaneDev& L = lanes[lane0 + blockIdx.y];
  LaneVars& V = *L.v;
  const int ne = V.n_ds[0], ns = V.n_ds[1];
  const int me = V.n_map[0], ms = V.n_map[1];
  const int cap = cfg.cap_map + cfg.cap_scan;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < ne + ns; i += gridDim.x * 256) {
    const int w = i < ne ? 0 : 1;
    const int k = i < ne ? i : i - ne;
    const int dst = (w ? ms : me) + k;
    if (dst < cap) L.map[w][cur][dst] = associate(V.x, L.ds[w][k]);  // EM:313-314, :321-322
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    V.n_cat[0] = min(cap, me + ne);
    V.n_cat[1] = min(cap, ms + ns);
    if (me + ne > cap || ms + ns > cap) atomicOr(&V.status, ST_MAP_CAPACITY);
  }
}

__global__ void __launch_bounds__(256) k_map_init(LaneDev* lanes, int lane0, int cur, ConfigDev cfg) {
  const LaneDev& L = lanes[lane0 + blockIdx.y];
  LaneVars& V = *L.v;
  const int ne = V.n_edge, ns = V.n_surf;
  const int me = V.n_map[0], ms = V.n_map[1];  // EM:107-108 append (maps are empty on the first frame)
  const int cap = cfg.cap_map + cfg.cap_scan;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < ne + ns; i += gridDim.x * 256) {
    const int w = i < ne ? 0 : 1;
    const int k = i < ne ? i : i - ne;
    const int dst = (w ? ms : me) + k;
    if (dst < cap) L.map[w][cur][dst] = L.feat[w][k];
  }
  __threadfence();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (me + ne > cap || ms + ns > cap) atomicOr(&V.status, ST_MAP_CAPACITY);
  }
}
__global__ void k_map_init_commit(LaneDev* lanes, int lane0, ConfigDev cfg) {
  const int t = threadIdx.x;
  LaneVars& V = *lanes[lane0 + t].v;
  const int cap = cfg.cap_map + cfg.cap_scan;
  V.n_map[0] = min(cap, V.n_map[0] + V.n_edge);
  V.n_map[1] = min(cap, V.n_map[1] + V.n_surf);
  V.n_ds[0] = 0; V.n_ds[1] = 0;
}

void launch_map_append(const Launch& L, LaneDev* lanes, int lane0, int nlanes, int cur, const ConfigDev& cfg) {
  dim3 g(64, nlanes);
  k_map_append<<<g, 256, 0, L.st>>>(lanes, lane0, cur, cfg);
  L.tick(K_MAP_APPEND);
}
void launch_map_init(const Launch& L, LaneDev* lanes, int lane0, int nlanes, int cur, const ConfigDev& cfg) {
  dim3 g(148, nlanes);
  k_map_init<<<g, 256, 0, L.st>>>(lanes, lane0, cur, cfg);
  L.tick(K_MAP_INIT);
  k_map_init_commit<<<1, nlanes, 0, L.st>>>(lanes, lane0, cfg);
  L.tick(K_MAP_INIT);
}

}  // namespace vilf
