// pcl::VoxelGrid<PointXYZI>::applyFilter (PCL 1.7.2 filters/impl/voxel_grid.hpp; call sites EM:248-251 for
// the scan features and EM:347-350 for the local maps) and pcl::CropBox (EM:335-344), batched over jobs.
//
//   k_vox_bbox      crop-box test (closed AABB, bounds cast to fp32) + getMinMax3D over the kept points
//   k_sort_keyhist<KeyGenVoxel>  min_b/div_b, PCL's int32 guard, fp32 voxel index per point (cropped-out points get
//                   the sentinel key dx*dy*dz and sort to the tail), pass-0 digit histogram
//   k_sort_scatter x 3-4   stable radix sort by voxel index, next-pass histogram fused (k_sort.cu)
//   k_vox_heads     first point of every occupied voxel, counted per CTA chunk
//   k_vox_centroid  every warp walks a contiguous range of the sorted points 32 at a time and accumulates x, y, z,
//                   intensity of every voxel in input order in fp32, exactly like PCL's `centroid += ...; centroid /= count`;
//                   output order = ascending voxel index = PCL's output order
//   k_map_append    createSubMap step 1 (EM:308-324): world-transform the filtered scan features and append
//   k_map_init      localMapInited (EM:105-115)
//
// Algorithmic bytes per point: 16 B read + 16 B/voxel written; the sort's (key,index) traffic
// (3-4 passes x 16 B) is implementation overhead.
#include "k_voxel.cuh"

namespace vilf {

__global__ void __launch_bounds__(256) k_vox_bbox(const VoxJob* __restrict__ jobs) {
  const VoxJob& J = jobs[blockIdx.y];
  const int n = *J.n_in;
  float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  if (J.crop) crop_bounds(J, lo, hi);
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  int cnt = 0;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const float4 p = J.in[i];
    if (J.crop && outside(p, lo, hi)) continue;
    ++cnt;
    mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
    mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
    mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
  }
  __shared__ float bb_sm[7 * 8];
  bbox_commit(J.vv, mn, mx, cnt, bb_sm);
}

__device__ __forceinline__ void vox_chunk(int n, int b, int& beg, int& end) {
  int chunk = (n + VOX_G - 1) / VOX_G;
  chunk = (chunk + 255) / 256 * 256;
  beg = min(n, b * chunk);
  end = min(n, beg + chunk);
}

// Sorted positions are split into VOX_G CTA chunks of 8 contiguous warp ranges each (multiples of 32 positions).
__device__ __forceinline__ void vox_warp_range(int n, int cta, int warp, int& beg, int& end) {
  int cb, ce;
  vox_chunk(n, cta, cb, ce);
  int chunk = (n + VOX_G - 1) / VOX_G;
  chunk = (chunk + 255) / 256 * 256;
  const int wchunk = chunk / 8;
  beg = min(ce, cb + warp * wchunk);
  end = min(ce, beg + wchunk);
}

__global__ void __launch_bounds__(256) k_vox_heads(const VoxJob* __restrict__ jobs) {
  const VoxJob& J = jobs[blockIdx.y];
  const int n = J.vv->n_valid;
  const int guard = J.vv->guard | J.emit_all;
  const uint32_t* key = J.sort.key[sort_passes(J.vv->bits, J.sort.npass) & 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int beg, end;
  vox_warp_range(n, blockIdx.x, warp, beg, end);
  int cnt = 0;
  for (int i = beg + lane; i < end; i += 32) cnt += (guard || i == 0 || key[i] != key[i - 1]) ? 1 : 0;
  for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
  if (lane == 0) J.head_cnt[blockIdx.x * 8 + warp] = cnt;  // heads (first points of occupied voxels) per warp range
}

// Every warp emits the centroids of the voxels whose head lies in its range (k_voxel.cuh: emit_range — the same
// warp-collective, software-pipelined emitter as the cluster path, hence identical bits), at the output index given by
// the number of heads in all earlier ranges.
__global__ void __launch_bounds__(256) k_vox_centroid(const VoxJob* __restrict__ jobs) {
  const VoxJob& J = jobs[blockIdx.y];
  const int n = J.vv->n_valid;
  const int guard = J.vv->guard | J.emit_all;
  const int res = sort_passes(J.vv->bits, J.sort.npass) & 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ int red[8];
  __shared__ __align__(16) float4 stage[8][64];
  const int first = blockIdx.x * 8;  // index of this CTA's first warp range
  int pre = 0, tot = 0;
  for (int b = tid; b < VOX_G * 8; b += 256) { const int c = J.head_cnt[b]; if (b < first) pre += c; tot += c; }
  // block sums of `pre` (heads before this CTA) and `tot` (all heads)
  for (int off = 16; off > 0; off >>= 1) { pre += __shfl_xor_sync(0xffffffffu, pre, off); tot += __shfl_xor_sync(0xffffffffu, tot, off); }
  __shared__ int rp[8], rt[8];
  if (lane == 0) { rp[warp] = pre; rt[warp] = tot; }
  __syncthreads();
  pre = 0; tot = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) { pre += rp[w]; tot += rt[w]; }
  if (tid < 8) red[tid] = J.head_cnt[first + tid];
  __syncthreads();
  for (int w = 0; w < warp; ++w) pre += red[w];
  if (blockIdx.x == 0 && tid == 0) {
    int total = tot;
    if (total > J.cap_out) { atomicOr(J.status, ST_MAP_CAPACITY); total = J.cap_out; }
    *J.n_out = total;
  }
  int beg, end;
  vox_warp_range(n, blockIdx.x, warp, beg, end);
  emit_range(J, KvSplit{J.sort.key[res], J.sort.val[res]}, n, guard, beg, end, pre, &stage[warp][0]);
}

void launch_voxel(const Launch& L, const VoxJob* jobs_dev, int njobs, const SortJob* sort_jobs_dev, bool bbox_done) {
  dim3 gv(VOX_G, njobs);
  if (!bbox_done) {  // otherwise the producer of the input cloud already reduced the bounding box (k_compact_features / k_map_append)
    k_vox_bbox<<<gv, 256, 0, L.st>>>(jobs_dev);
    L.tick(K_VOX_BBOX);
  }
  KeyGenVoxel gen;
  gen.jobs = jobs_dev;
  dim3 gs(SORT_G, njobs);
  k_sort_keyhist<KeyGenVoxel><<<gs, SORT_THREADS, 0, L.st>>>(sort_jobs_dev, gen);
  L.tick(K_VOX_KEYHIST);
  for (int pass = 0; pass < 4; ++pass) launch_sort_scatter(L, sort_jobs_dev, njobs, pass);
  k_vox_heads<<<gv, 256, 0, L.st>>>(jobs_dev);
  L.tick(K_VOX_HEADS);
  k_vox_centroid<<<gv, 256, 0, L.st>>>(jobs_dev);
  L.tick(K_VOX_CENTROID);
}

// ------------------------------------------------------------------------------------------------
// map maintenance helpers
// ------------------------------------------------------------------------------------------------
// createSubMap step 1 (EM:308-324) fused with the crop-box test and getMinMax3D of step 2-4's input: the voxel-filtered
// scan features are transformed with the final pose and appended; every point of map + appended (old and new) that lies
// inside the crop box (EM:327-344) contributes to the bounding box the voxel filter needs.
__global__ void __launch_bounds__(256) k_map_append(LaneDev* lanes, int lane0, int cur, ConfigDev cfg) {
  const LaneDev& L = lanes[lane0 + blockIdx.y];
  LaneVars& V = *L.v;
  const int cap = cfg.cap_map + cfg.cap_scan;
  float lo[3], hi[3];
  for (int a = 0; a < 3; ++a) {
    lo[a] = (float)dsub(V.x[4 + a], cfg.crop_half);
    hi[a] = (float)dadd(V.x[4 + a], cfg.crop_half);
  }
  __shared__ float bb_sm[7 * 8];
  for (int w = 0; w < 2; ++w) {
    const int nd = V.n_ds[w], m = V.n_map[w];
    const int tot = min(cap, m + nd);
    float4* map = L.map[w][cur];
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    int cnt = 0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < tot; i += gridDim.x * 256) {
      float4 p;
      if (i < m) {
        p = map[i];
      } else {
        p = associate(V.x, L.ds[w][i - m]);  // EM:313-314, :321-322
        map[i] = p;
      }
      if (p.x < lo[0] || p.y < lo[1] || p.z < lo[2] || p.x > hi[0] || p.y > hi[1] || p.z > hi[2]) continue;
      ++cnt;
      mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
      mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
      mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
    }
    bbox_commit(L.vv + 2 + w, mn, mx, cnt, bb_sm);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      V.n_cat[w] = tot;
      if (m + nd > cap) atomicOr(&V.status, ST_MAP_CAPACITY);
    }
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
  dim3 g(148, nlanes);
  k_map_append<<<g, 256, 0, L.st>>>(lanes, lane0, cur, cfg);
  L.tick(K_MAP_APPEND);
}
void launch_map_init_commit(const Launch& L, LaneDev* lanes, int lane0, int nlanes, const ConfigDev& cfg) {
  k_map_init_commit<<<1, nlanes, 0, L.st>>>(lanes, lane0, cfg);
  L.tick(K_MAP_INIT);
}
void launch_map_init(const Launch& L, LaneDev* lanes, int lane0, int nlanes, int cur, const ConfigDev& cfg) {
  dim3 g(148, nlanes);
  k_map_init<<<g, 256, 0, L.st>>>(lanes, lane0, cur, cfg);
  L.tick(K_MAP_INIT);
  k_map_init_commit<<<1, nlanes, 0, L.st>>>(lanes, lane0, cfg);
  L.tick(K_MAP_INIT);
}

}  // namespace vilf
