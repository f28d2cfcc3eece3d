// One thread-block CLUSTER per cloud: the latency-optimised path for clouds of up to CLUSTER_MAX_POINTS points (every
// per-frame cloud of an HDL-64 / 128-beam sequence).  The grid-wide kernels of k_voxel.cu / k_knn.cu need 1 + 7 + 5
// dependent launches for map append + voxel filter + grid build, and at these sizes they are launch / tail / atomics
// bound (profiles/r1b_frame_S16.txt: 377 us of a 1275 us frame).  Here the phases of one job run inside ONE kernel on 8
// SMs, separated by cluster barriers instead of kernel boundaries, and the CTAs exchange digit histograms / head counts
// / partial sums through distributed shared memory instead of global tables.  Independent jobs (edge + surf cloud of
// every sequence of a batch) are separate clusters of the same launch (blockIdx.y), so a batch still fills the GPU.
//
//   k_voxel_cluster   [createSubMap append (EM:308-324)] -> pcl::CropBox + getMinMax3D -> PCL voxel keys -> P stable
//                     radix passes (<= 9-bit digits) -> heads -> centroids (EM:248-251, :327-350)
//                     [-> spatial-hash build of the filtered map for the 5-NN search (EM:256-257)].
//                     Same key generator, same stable order and same centroid emitter as the grid-wide path, hence
//                     bit-identical output.
//   k_grid_cluster    the spatial-hash build alone (first frame, state import).
#include "k_cluster_sort.cuh"

namespace vilf {

__device__ __forceinline__ uint32_t cell_hash_c(int x, int y, int z) {
  return ((uint32_t)x * 73856093u) ^ ((uint32_t)y * 19349663u) ^ ((uint32_t)z * 83492791u);
}

// Spatial-hash build (same table layout as k_knn.cu's grid-wide build) by the whole cluster.  `S` = this CTA's shared block.
__device__ void grid_build_cluster(cg::cluster_group& cluster, const GridJob& G, int n, VoxShared& S) {
  const float inv_cell = G.inv_cell;
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x;
  int H = 1024;
  while (H < 2 * n && H < G.hcap) H <<= 1;
  const uint32_t hm = (uint32_t)H - 1u;
  const int total = H + 1;
  const int gtid = rank * CT + tid, gstride = CL * CT;
  for (int i = gtid; i < total; i += gstride) G.start[i] = 0;
  if (gtid == 0) *G.hvar = H;
  cluster.sync();
  for (int i = gtid; i < n; i += gstride) {
    const float4 p = __ldcg(G.pts + i);
    const int cx = (int)floorf(fmul(p.x, inv_cell)), cy = (int)floorf(fmul(p.y, inv_cell)), cz = (int)floorf(fmul(p.z, inv_cell));
    G.rank[i] = atomicAdd(&G.start[cell_hash_c(cx, cy, cz) & hm], 1u);
  }
  cluster.sync();
  // exclusive scan of the bucket counts: CTA `rank` owns one contiguous chunk
  int chunk = (total + CL - 1) / CL;
  chunk = (chunk + CT - 1) / CT * CT;
  const int beg = min(total, rank * chunk), end = min(total, beg + chunk);
  {
    uint32_t s = 0;
    for (int i = beg + tid; i < end; i += CT) s += __ldcg(G.start + i);
    uint32_t tot = 0;
    block_excl_scan(s, S.scan, &tot);
    if (tid == 0) S.part = tot;
  }
  cluster.sync();
  uint32_t run = 0;
  for (int c = 0; c < rank; ++c) run += cluster.map_shared_rank(&S, c)->part;
  for (int base = beg; base < end; base += CT) {
    const int i = base + tid;
    const uint32_t v = i < end ? __ldcg(G.start + i) : 0u;
    uint32_t tot = 0;
    const uint32_t ex = block_excl_scan(v, S.scan, &tot);
    if (i < end) G.start[i] = run + ex;
    run += tot;
  }
  cluster.sync();
  for (int i = gtid; i < n; i += gstride) {
    const float4 p = __ldcg(G.pts + i);
    const int cx = (int)floorf(fmul(p.x, inv_cell)), cy = (int)floorf(fmul(p.y, inv_cell)), cz = (int)floorf(fmul(p.z, inv_cell));
    const uint32_t pos = __ldcg(G.start + (cell_hash_c(cx, cy, cz) & hm)) + G.rank[i];
    G.sorted[pos] = make_float4(p.x, p.y, p.z, __int_as_float(i));
  }
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(CT, 2) k_voxel_cluster(const VoxJob* __restrict__ jobs, int bbox_done) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const VoxJob& J = jobs[blockIdx.y];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ VoxShared S;

  PHASE_MARK(0);
  // ---- input size; map jobs first append the voxel-filtered scan features at the final pose (EM:308-324) ----
  int n, m_old = 0;
  if (J.app_src != nullptr) {
    m_old = *J.app_n_map;
    const int nd = *J.app_n;
    n = min(J.app_cap, m_old + nd);
    if (rank == 0 && tid == 0) {
      *const_cast<int*>(J.n_in) = n;
      if (m_old + nd > J.app_cap) atomicOr(J.status, ST_MAP_CAPACITY);
    }
  } else {
    n = *J.n_in;
  }

  // ---- geometry: CTA `rank` owns one contiguous chunk, each of its warps one contiguous sub-chunk (stable order) ----
  int cchunk = (n + CL - 1) / CL;
  cchunk = (cchunk + CT - 1) / CT * CT;
  const int wchunk = cchunk / CW;  // multiple of 32
  const int cbeg = min(n, rank * cchunk);
  const int cend = min(n, cbeg + cchunk);
  const int wbeg = min(cend, cbeg + warp * wchunk);
  const int wend = min(cend, wbeg + wchunk);

  // ---- phase 0: [append] + bounding box of the points inside the crop box (getMinMax3D) ----
  float mn[3], mx[3];
  int n_valid;
  if (bbox_done && J.app_src == nullptr) {
    for (int a = 0; a < 3; ++a) { mn[a] = ord2f(J.vv->bbox[a]); mx[a] = ord2f(J.vv->bbox[3 + a]); }
    n_valid = J.vv->n_valid;
  } else {
    float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    if (J.crop) crop_bounds(J, lo, hi);
    double x[7];
    if (J.app_src != nullptr) {
#pragma unroll
      for (int i = 0; i < 7; ++i) x[i] = J.app_pose[i];
    }
    for (int a = 0; a < 3; ++a) { mn[a] = FLT_MAX; mx[a] = -FLT_MAX; }
    int cnt = 0;
    float4* inw = const_cast<float4*>(J.in);
    for (int i = cbeg + tid; i < cend; i += CT) {
      float4 p;
      if (i < m_old || J.app_src == nullptr) {
        p = J.in[i];
      } else {
        p = associate(x, J.app_src[i - m_old]);  // EM:313-314, :321-322
        inw[i] = p;
      }
      if (J.crop && outside(p, lo, hi)) continue;
      ++cnt;
      mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
      mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
      mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
    }
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], off));
        mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], off));
      }
      cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    }
    float* sm = reinterpret_cast<float*>(&S.wcnt[0][0]);  // scratch, free until the first pass
    if (lane == 0) {
      for (int a = 0; a < 3; ++a) { sm[warp * 7 + a] = mn[a]; sm[warp * 7 + 3 + a] = mx[a]; }
      sm[warp * 7 + 6] = __int_as_float(cnt);
    }
    __syncthreads();
    if (tid == 0) {
      int total = 0;
      for (int w = 0; w < CW; ++w) {
        for (int a = 0; a < 3; ++a) { mn[a] = fminf(mn[a], sm[w * 7 + a]); mx[a] = fmaxf(mx[a], sm[w * 7 + 3 + a]); }
        total += __float_as_int(sm[w * 7 + 6]);
      }
      for (int a = 0; a < 3; ++a) { S.bb[a] = mn[a]; S.bb[3 + a] = mx[a]; }
      S.bb[6] = __int_as_float(total);
    }
    cluster.sync();  // partial boxes visible; appended points visible to the whole cluster
    n_valid = 0;
    for (int a = 0; a < 3; ++a) { mn[a] = FLT_MAX; mx[a] = -FLT_MAX; }
    for (int c = 0; c < CL; ++c) {
      const float* rb = cluster.map_shared_rank(&S, c)->bb;
      for (int a = 0; a < 3; ++a) { mn[a] = fminf(mn[a], rb[a]); mx[a] = fmaxf(mx[a], rb[3 + a]); }
      n_valid += __float_as_int(rb[6]);
    }
    if (rank == 0 && tid == 0) {
      for (int a = 0; a < 3; ++a) { J.vv->bbox[a] = f2ord(mn[a]); J.vv->bbox[3 + a] = f2ord(mx[a]); }
      J.vv->n_valid = n_valid;
    }
    __syncthreads();  // S.wcnt scratch is reused below
  }

  PHASE_MARK(1);
  // ---- phase 1: voxel keys + stable radix sort ----
  KeyGenVoxel gen;
  gen.jobs = jobs;
  const int bits = gen.setup(J, mn, mx, n_valid, rank == 0 && tid == 0);
  const int guard = gen.guard;
  const int P = sort_passes(bits, J.sort.npass);
  cluster_radix_sort(cluster, S, J.sort, wbeg, wend, bits, [&](int i) { return gen.key(0, i); });

  PHASE_MARK(2);
  // ---- phase 2: heads (first point of every occupied voxel) and centroids, in ascending voxel order ----
  const uint2* kv = J.sort.pair[P & 1];  // first L1-allocating reads of this buffer in this kernel: ld.global.ca is coherent here
  const int nv = n_valid;  // cropped-out points carry the sentinel key and sit behind the valid ones
  int hchunk = (nv + CL - 1) / CL;
  hchunk = (hchunk + CT - 1) / CT * CT;
  const int hwchunk = hchunk / CW;
  const int hbeg = min(nv, rank * hchunk), hend = min(nv, hbeg + hchunk);
  const int hwbeg = min(hend, hbeg + warp * hwchunk), hwend = min(hend, hwbeg + hwchunk);
  {
    int cnt = 0;
    for (int i = hwbeg + lane; i < hwend; i += 32) cnt += (guard || i == 0 || __ldca(kv + i).x != __ldca(kv + i - 1).x) ? 1 : 0;
    for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    if (lane == 0) S.wsum[warp] = cnt;
    __syncthreads();
    if (tid == 0) {
      int s = 0;
      for (int ww = 0; ww < CW; ++ww) s += S.wsum[ww];
      S.heads = s;
    }
  }
  cluster.sync();
  int run = 0, total_heads = 0;
  for (int c = 0; c < CL; ++c) {
    const int h = cluster.map_shared_rank(&S, c)->heads;
    if (c < rank) run += h;
    total_heads += h;
  }
  for (int ww = 0; ww < warp; ++ww) run += S.wsum[ww];
  int n_out = total_heads;
  if (n_out > J.cap_out) n_out = J.cap_out;
  if (rank == 0 && tid == 0) {
    if (total_heads > J.cap_out) atomicOr(J.status, ST_MAP_CAPACITY);
    *J.n_out = n_out;
  }
  PHASE_MARK(3);
  // staging of the centroid emitter: two 32-point buffers per warp, in the digit-counter table the sort no longer needs
  emit_range(J, KvPairs{kv}, nv, guard, hwbeg, hwend, run, reinterpret_cast<float4*>(&S.wcnt[0][0]) + warp * 64);  // warps run independently

  // ---- phase 3 (map jobs): spatial hash of the filtered map for the next frame's 5-NN search ----
  if (J.grid != nullptr) {
    cluster.sync();  // all centroids written
    PHASE_MARK(4);
    grid_build_cluster(cluster, *J.grid, n_out, S);
  }
  cluster.sync();  // no CTA may exit while a peer can still read its shared memory
  PHASE_MARK(5);
}

void launch_voxel_cluster(const Launch& L, const VoxJob* jobs_dev, int njobs, bool bbox_done, const ConfigDev&) {
#if VILF_VOX_CLUSTER > 8
  cudaFuncSetAttribute(k_voxel_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);  // A/B builds only
#endif
  dim3 g(CL, njobs);
  k_voxel_cluster<<<g, CT, 0, L.st>>>(jobs_dev, bbox_done ? 1 : 0);
  L.tick(K_VOX_CLUSTER);
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(CT, 2) k_grid_cluster(const GridJob* __restrict__ jobs) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ VoxShared S;
  const GridJob& G = jobs[blockIdx.y];
  grid_build_cluster(cluster, G, *G.n, S);
  cluster.sync();
}

void launch_grid_cluster(const Launch& L, const GridJob* jobs_dev, int njobs, const ConfigDev&) {
#if VILF_VOX_CLUSTER > 8
  cudaFuncSetAttribute(k_grid_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);  // A/B builds only
#endif
  dim3 g(CL, njobs);
  k_grid_cluster<<<g, CT, 0, L.st>>>(jobs_dev);
  L.tick(K_GRID_CLUSTER);
}

}  // namespace vilf
