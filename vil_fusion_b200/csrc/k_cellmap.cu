// Cell-ordered local maps: ONE spatial order for pcl::VoxelGrid (EM:347-350), pcl::CropBox (EM:335-344), createSubMap's append
// (EM:308-324) and pcl::KdTreeFLANN (EM:256-257, :128, :185).
//
// Order.  Every map point has the voxel coordinate v = floor(p * inverse_leaf) PCL's filter gives it (fp32, voxel_grid.hpp);
// with k = 2^shift voxels per search-cell edge the 63-bit key is (cell z, cell y, cell x, voxel-in-cell z, y, x).  A voxel is a
// run of equal keys, a search cell is a contiguous range, and the three cells of an x-row are adjacent in memory.
//
// Update (k_new_xform .. k_merge<EMIT>).  The previous map is already in key order and voxel-filtered; a frame adds a few
// thousand points.  So instead of re-sorting ~1e6 points (the radix path: 3-4 passes of 16 B per point plus a gather) only the
// new points are sorted, and the map update is a MERGE: merge-path tiles of 1024 elements, each staged in shared memory once,
// crop box applied on the fly, every voxel run summed sequentially in fp32 in PCL's order (old points first, then the new
// ones, in input order), output = the other map buffer.  A counting pass sizes the tiles' outputs, a one-CTA scan turns the
// counts into offsets (no atomics, no look-back spinning), the emitting pass writes points and the cell table.  Algorithmic bytes: 16 B read per old point + 16 B written
// per kept voxel; the second read of the old map comes from L2.
//
// Search (k_knn_cell*).  One WARP per query: lane l probes neighbour cell l of the 27 (open addressing, verified by the cell of
// the first point of the range, which is a candidate anyway); the candidate ranges are concatenated with a warp scan and read
// 32 at a time as coalesced float4; selection of the five best uses redux.sync min on the distance bits — no per-lane lists, no
// merge tree.  Finer cells than the gate radius (dense maps) are walked shell by shell with the early exit of k_knn.cu.
// The search of the second outer iteration of a frame is seeded with the first one's neighbours: they bound the fifth distance
// before any cell is read, and cells whose box lies outside that ball are not probed (k_knn_cell_assoc<true>).
//
// The PCL order of the points (ascending voxel index) differs from the stored order; everything the reference's results depend
// on is order independent or handled explicitly: voxel sums run in input order inside a voxel (= PCL's stable order), exact
// distance ties are resolved by PCL rank (recomputed from the two points on the rare tie), and the API boundary
// (vilf_get_cloud, vilf_factors) converts to PCL order on demand.
#include "k_sort.cuh"
#include <cstdlib>
#include "k_voxel.cuh"
#include "k_cluster_sort.cuh"

namespace vilf {

// ------------------------------------------------------------------------------------------------
// keys
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void voxel_of(const float4 p, float inv, int& vx, int& vy, int& vz) {
  // PCL: static_cast<int>(floor(p * inverse_leaf) - min_b) with integer-valued floats below 2^24 == floor(p * inverse_leaf) - min_b
  vx = (int)floorf(fmul(p.x, inv));
  vy = (int)floorf(fmul(p.y, inv));
  vz = (int)floorf(fmul(p.z, inv));
}
__device__ __forceinline__ uint32_t vbias(int v) { return (uint32_t)(min(max(v, -VOX_BIAS), VOX_BIAS - 1) + VOX_BIAS); }
// 63-bit key: [cell z][cell y][cell x][sub z][sub y][sub x], 21 - s bits per cell field, s bits per sub field
__device__ __forceinline__ unsigned long long key64_of(int vx, int vy, int vz, int s) {
  const uint32_t x = vbias(vx), y = vbias(vy), z = vbias(vz);
  const uint32_t m = (1u << s) - 1u;
  const int cb = 21 - s;
  const unsigned long long cell = ((unsigned long long)(z >> s) << (2 * cb)) | ((unsigned long long)(y >> s) << cb) | (unsigned long long)(x >> s);
  const unsigned long long sub = ((unsigned long long)(z & m) << (2 * s)) | ((unsigned long long)(y & m) << s) | (unsigned long long)(x & m);
  return (cell << (3 * s)) | sub;
}
__device__ __forceinline__ unsigned long long key64_pt(const float4 p, const CellGeom& g) {
  int vx, vy, vz;
  voxel_of(p, g.inv_leaf, vx, vy, vz);
  return key64_of(vx, vy, vz, g.shift);
}
__device__ __forceinline__ unsigned long long cellkey_cells(uint32_t cx, uint32_t cy, uint32_t cz, int s) {
  const int cb = 21 - s;
  return ((unsigned long long)cz << (2 * cb)) | ((unsigned long long)cy << cb) | (unsigned long long)cx;
}
__device__ __forceinline__ uint32_t cell_slot(unsigned long long cellkey) { return (uint32_t)((cellkey * 0x9E3779B97F4A7C15ull) >> 32); }
constexpr uint32_t SLOT_EMPTY = 0xFFFFFFFFu;

__device__ __forceinline__ void table_insert(uint2* tab, uint32_t mask, unsigned long long cellkey, uint32_t start, uint32_t end) {
  uint32_t h = cell_slot(cellkey) & mask;
  for (;;) {
    const uint32_t old = atomicCAS(&tab[h].x, SLOT_EMPTY, start);
    if (old == SLOT_EMPTY) { tab[h].y = end; return; }
    h = (h + 1) & mask;
  }
}
__device__ __forceinline__ int table_size(int n, int hcap) {
  int h = 1024;
  while (h < 2 * n && h < hcap) h <<= 1;
  return h;
}

// ------------------------------------------------------------------------------------------------
// generic build: arbitrary cloud -> cell-ordered copy + original indices + cell table (first frame, state import, explicit maps)
// ------------------------------------------------------------------------------------------------
// k_cb_reset   voxel bounding box init, table clear
// k_cb_bbox    voxel-coordinate bounding box
// k_sort_keyhist<KeyGenCell> + k_sort_scatter x 4   stable radix sort by the key relative to the bounding box
// k_cb_gather  dst[i] = src[val[i]], orig[i] = val[i], cell heads -> table starts
// k_cb_ends    last point of every cell -> table ends
__global__ void __launch_bounds__(256) k_cb_reset(const CellBuildJob* __restrict__ jobs) {
  const CellBuildJob& J = jobs[blockIdx.y];
  const int n = *J.n;
  const int H = table_size(n, J.hcap);
  for (int i = blockIdx.x * 256 + threadIdx.x; i < H; i += gridDim.x * 256) J.table[i] = make_uint2(SLOT_EMPTY, 0u);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    J.meta[0] = H - 1; J.meta[1] = 1;
    J.meta[2] = J.meta[3] = J.meta[4] = INT_MAX;
    J.meta[5] = J.meta[6] = J.meta[7] = INT_MIN;
  }
}
__global__ void __launch_bounds__(256) k_cb_bbox(const CellBuildJob* __restrict__ jobs) {
  const CellBuildJob& J = jobs[blockIdx.y];
  const int n = *J.n;
  int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    int v[3];
    voxel_of(J.src[i], J.g.inv_leaf, v[0], v[1], v[2]);
#pragma unroll
    for (int a = 0; a < 3; ++a) { const int b = (int)vbias(v[a]); mn[a] = min(mn[a], b); mx[a] = max(mx[a], b); }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) { mn[a] = __reduce_min_sync(0xffffffffu, mn[a]); mx[a] = __reduce_max_sync(0xffffffffu, mx[a]); }
  if ((threadIdx.x & 31) == 0 && mn[0] != INT_MAX) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { atomicMin(&J.meta[2 + a], mn[a]); atomicMax(&J.meta[5 + a], mx[a]); }
  }
}

// Relative key of a point inside a voxel bounding box (biased voxel coordinates lo[], hi[]): cells numbered x-fastest over the
// box, then the voxel inside the cell.  Order-isomorphic to key64 for points of the box.
struct RelKey {
  uint32_t c0[3];   // first cell of the box per axis
  uint32_t nc[2];   // cells along x, y
  int s, bits;
  bool ok;
  __device__ void setup(const int lo[3], const int hi[3], int shift, bool any) {
    s = shift; ok = true; bits = 1;
    c0[0] = c0[1] = c0[2] = 0; nc[0] = nc[1] = 1;
    if (!any) return;
    unsigned long long tot = 1;
    uint32_t ncz = 1;
    for (int a = 0; a < 3; ++a) {
      c0[a] = (uint32_t)lo[a] >> s;
      const uint32_t cnt = ((uint32_t)hi[a] >> s) - c0[a] + 1u;
      if (a < 2) nc[a] = cnt; else ncz = cnt;
      tot *= cnt;
    }
    (void)ncz;
    tot <<= 3 * s;  // keys are 0 .. tot - 1, tot = sentinel of dropped points
    if (tot >= 0xFFFFFFFFull) { ok = false; return; }
    bits = 64 - __clzll(tot);
  }
  __device__ __forceinline__ uint32_t sentinel() const { return ok ? (bits >= 32 ? 0xFFFFFFFFu : (1u << bits) - 1u) : 0u; }
  __device__ __forceinline__ uint32_t key(int vx, int vy, int vz) const {
    if (!ok) return 0u;
    const uint32_t x = vbias(vx), y = vbias(vy), z = vbias(vz);
    const uint32_t m = (1u << s) - 1u;
    const uint32_t cell = (((z >> s) - c0[2]) * nc[1] + ((y >> s) - c0[1])) * nc[0] + ((x >> s) - c0[0]);
    return (cell << (3 * s)) | ((z & m) << (2 * s)) | ((y & m) << s) | (x & m);
  }
};

struct KeyGenCell {
  const CellBuildJob* jobs;
  RelKey rk;
  const float4* in;
  float inv;
  __device__ int prepare(int job) {
    const CellBuildJob& J = jobs[job];
    in = J.src; inv = J.g.inv_leaf;
    int lo[3], hi[3];
    for (int a = 0; a < 3; ++a) { lo[a] = J.meta[2 + a]; hi[a] = J.meta[5 + a]; }
    rk.setup(lo, hi, J.g.shift, lo[0] != INT_MAX);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      J.meta[8] = rk.bits;
      if (!rk.ok) atomicOr(J.status, ST_KEY_RANGE);
    }
    return rk.bits;
  }
  __device__ uint32_t key(int, int i) const {
    int vx, vy, vz;
    voxel_of(in[i], inv, vx, vy, vz);
    return rk.key(vx, vy, vz);
  }
};

__global__ void __launch_bounds__(256) k_cb_gather(const CellBuildJob* __restrict__ jobs) {
  const CellBuildJob& J = jobs[blockIdx.y];
  if (*J.status & ST_KEY_RANGE) return;  // keys are meaningless: the call fails with VILF_ERR_UNSUPPORTED
  const int n = *J.n;
  const int bits = J.meta[8];
  const int res = sort_passes(bits, J.sort.npass) & 1;
  const uint32_t* __restrict__ key = J.sort.key[res];
  const uint32_t* __restrict__ val = J.sort.val[res];
  const uint32_t mask = (uint32_t)J.meta[0];
  const int s3 = 3 * J.g.shift;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const uint32_t v = val[i];
    const float4 p = J.src[v];
    J.dst[i] = p;
    J.orig[i] = v;
    const uint32_t c = key[i] >> s3;
    if (i == 0 || (key[i - 1] >> s3) != c) {  // first point of a cell: claim a slot (the end follows in k_cb_ends)
      uint32_t h = cell_slot(key64_pt(p, J.g) >> s3) & mask;
      for (;;) {
        if (atomicCAS(&J.table[h].x, SLOT_EMPTY, (uint32_t)i) == SLOT_EMPTY) break;
        h = (h + 1) & mask;
      }
    }
  }
}
__global__ void __launch_bounds__(256) k_cb_ends(const CellBuildJob* __restrict__ jobs) {
  const CellBuildJob& J = jobs[blockIdx.y];
  if (*J.status & ST_KEY_RANGE) return;
  const int n = *J.n;
  const int bits = J.meta[8];
  const uint32_t* __restrict__ key = J.sort.key[sort_passes(bits, J.sort.npass) & 1];
  const uint32_t mask = (uint32_t)J.meta[0];
  const int s3 = 3 * J.g.shift;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    const uint32_t c = key[i] >> s3;
    if (i == n - 1 || (key[i + 1] >> s3) != c) {  // last point of a cell: find the slot its first point claimed
      const unsigned long long ck = key64_pt(J.dst[i], J.g) >> s3;
      uint32_t h = cell_slot(ck) & mask;
      for (uint32_t probes = 0; probes <= mask; ++probes) {  // (always found: k_cb_gather inserted the cell)
        const uint32_t st = J.table[h].x;
        if (st == SLOT_EMPTY) break;
        if ((key64_pt(J.dst[st], J.g) >> s3) == ck) { J.table[h].y = (uint32_t)i + 1u; break; }
        h = (h + 1) & mask;
      }
    }
  }
}

void launch_cell_build(const Launch& L, const CellBuildJob* jobs_dev, const SortJob* sort_jobs_dev, int njobs) {
  dim3 g(148, njobs);
  k_cb_reset<<<g, 256, 0, L.st>>>(jobs_dev);
  L.tick(K_CELL_BUILD);
  k_cb_bbox<<<g, 256, 0, L.st>>>(jobs_dev);
  L.tick(K_CELL_BUILD);
  KeyGenCell gen;
  gen.jobs = jobs_dev;
  dim3 gs(SORT_G, njobs);
  k_sort_keyhist<KeyGenCell><<<gs, SORT_THREADS, 0, L.st>>>(sort_jobs_dev, gen);
  L.tick(K_CELL_BUILD);
  for (int pass = 0; pass < 4; ++pass) launch_sort_scatter(L, sort_jobs_dev, njobs, pass);
  k_cb_gather<<<g, 256, 0, L.st>>>(jobs_dev);
  L.tick(K_CELL_BUILD);
  k_cb_ends<<<g, 256, 0, L.st>>>(jobs_dev);
  L.tick(K_CELL_BUILD);
}

__global__ void __launch_bounds__(256) k_cell_unpermute(const float4* __restrict__ pts, const uint32_t* __restrict__ orig, const int* n_dev, float4* out, int cap) {
  const int n = min(*n_dev, cap);
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) out[orig[i]] = pts[i];
}
void launch_cell_unpermute(const Launch& L, const float4* pts, const uint32_t* orig, const int* n_dev, float4* out, int cap) {
  k_cell_unpermute<<<296, 256, 0, L.st>>>(pts, orig, n_dev, out, cap);
  L.tick(K_CELL_BUILD);
}

// ------------------------------------------------------------------------------------------------
// map update, step 1: the new points
// ------------------------------------------------------------------------------------------------
// createSubMap's append (EM:308-324): transform the voxel-filtered scan features with the final pose (pointAssociaToMap,
// fp64 -> fp32), drop what the crop box (EM:327-344) would drop anyway, and find the voxel bounding box of the rest.
// newpts = [orphans of the previous update ..., transformed features ...].
__device__ __forceinline__ void merge_crop(const MergeJob& J, float lo[3], float hi[3]) {
  for (int a = 0; a < 3; ++a) {
    lo[a] = -FLT_MAX; hi[a] = FLT_MAX;
    if (J.crop_center) {
      lo[a] = (float)dsub(J.crop_center[a], J.crop_half);  // EM:327-336: bounds in fp64, stored in an Eigen::Vector4f
      hi[a] = (float)dadd(J.crop_center[a], J.crop_half);
    }
  }
}
__global__ void k_merge_reset(const MergeJob* __restrict__ jobs, int njobs) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= njobs) return;
  MergeVars& V = *jobs[j].mv;
  V.n_live = 0; V.n_live_all = 0;
  V.vb[0] = V.vb[1] = V.vb[2] = INT_MAX; V.vb[3] = V.vb[4] = V.vb[5] = INT_MIN;
  V.fb[0] = V.fb[1] = V.fb[2] = INT_MAX; V.fb[3] = V.fb[4] = V.fb[5] = INT_MIN;
}
__global__ void __launch_bounds__(256) k_new_xform(const MergeJob* __restrict__ jobs) {
  const MergeJob& J = jobs[blockIdx.y];
  MergeVars& V = *J.mv;
  const int n_orph = min(V.n_orph, ORPHAN_CAP);
  int n_src = *J.n_src;
  if (n_orph + n_src > J.cap_new) n_src = J.cap_new - n_orph;
  const int n = n_orph + n_src;
  float lo[3], hi[3];
  merge_crop(J, lo, hi);
  double x[7];
  if (J.pose) {
#pragma unroll
    for (int i = 0; i < 7; ++i) x[i] = J.pose[i];
  }
  int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
  int live = 0;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
    float4 p;
    if (i < n_orph) p = J.orphans[i];
    else { p = J.src[i - n_orph]; if (J.pose) p = associate(x, p); }
    J.newpts[i] = p;
    if (outside(p, lo, hi)) continue;
    ++live;
    int v[3];
    voxel_of(p, J.g.inv_leaf, v[0], v[1], v[2]);
#pragma unroll
    for (int a = 0; a < 3; ++a) { const int b = (int)vbias(v[a]); mn[a] = min(mn[a], b); mx[a] = max(mx[a], b); }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) { mn[a] = __reduce_min_sync(0xffffffffu, mn[a]); mx[a] = __reduce_max_sync(0xffffffffu, mx[a]); }
  live = __reduce_add_sync(0xffffffffu, live);
  if ((threadIdx.x & 31) == 0 && live > 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { atomicMin(&V.vb[a], mn[a]); atomicMax(&V.vb[3 + a], mx[a]); }
    atomicAdd(&V.n_live, live);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    V.n_in = n; V.n_orph_in = n_orph;
    if (n_orph + *J.n_src > J.cap_new) atomicOr(J.status, ST_SCAN_CAPACITY);
  }
}

struct KeyGenNew {
  const MergeJob* jobs;
  RelKey rk;
  const float4* in;
  float inv;
  float lo[3], hi[3];
  __device__ int prepare(int job) {
    const MergeJob& J = jobs[job];
    MergeVars& V = *J.mv;
    in = J.newpts; inv = J.g.inv_leaf;
    merge_crop(J, lo, hi);
    int l[3], h[3];
    for (int a = 0; a < 3; ++a) { l[a] = V.vb[a]; h[a] = V.vb[3 + a]; }
    rk.setup(l, h, J.g.shift, V.n_live > 0);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      V.bits = rk.bits;
      if (!rk.ok) atomicOr(J.status, ST_KEY_RANGE);
    }
    return rk.bits;
  }
  __device__ uint32_t key(int, int i) const {
    const float4 p = in[i];
    if (outside(p, lo, hi)) return rk.sentinel();  // sorts behind every live point
    int vx, vy, vz;
    voxel_of(p, inv, vx, vy, vz);
    return rk.key(vx, vy, vz);
  }
};

// The same three steps (transform + crop + voxel bounding box, keys, stable radix sort) plus the sorted copy, for up to
// CLUSTER_MAX_POINTS new points, in ONE kernel on one 8-CTA cluster per map: the grid-wide version above costs seven dependent
// launches for a few thousand points.  Phases are separated by cluster barriers, histograms and partial boxes travel over
// distributed shared memory (k_cluster_sort.cuh).
struct NewShared {
  VoxShared V;
  int vb[8];  // this CTA's voxel bounding box (min xyz, max xyz) and live count
};
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(CT, 2) k_new_cluster(const MergeJob* __restrict__ jobs) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const MergeJob& J = jobs[blockIdx.y];
  MergeVars& V = *J.mv;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ NewShared S;
  const int n_orph = min(V.n_orph, ORPHAN_CAP);
  const int n_src_all = *J.n_src;
  const int n_src = min(n_src_all, J.cap_new - n_orph);
  const int n = n_orph + n_src;
  int cchunk = (n + CL - 1) / CL;
  cchunk = (cchunk + CT - 1) / CT * CT;
  const int wchunk = cchunk / CW;  // multiple of 32
  const int cbeg = min(n, rank * cchunk), cend = min(n, cbeg + cchunk);
  const int wbeg = min(cend, cbeg + warp * wchunk), wend = min(cend, wbeg + wchunk);
  float lo[3], hi[3];
  merge_crop(J, lo, hi);
  const CellGeom g = J.g;
  // ---- phase 0: append (EM:308-324) + crop test + voxel bounding box ----
  {
    double x[7];
    if (J.pose) {
#pragma unroll
      for (int i = 0; i < 7; ++i) x[i] = J.pose[i];
    }
    int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
    int live = 0;
    for (int i = cbeg + tid; i < cend; i += CT) {
      float4 p;
      if (i < n_orph) p = J.orphans[i];
      else { p = J.src[i - n_orph]; if (J.pose) p = associate(x, p); }
      J.newpts[i] = p;
      if (outside(p, lo, hi)) continue;
      ++live;
      int v[3];
      voxel_of(p, g.inv_leaf, v[0], v[1], v[2]);
#pragma unroll
      for (int a = 0; a < 3; ++a) { const int b = (int)vbias(v[a]); mn[a] = min(mn[a], b); mx[a] = max(mx[a], b); }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) { mn[a] = __reduce_min_sync(0xffffffffu, mn[a]); mx[a] = __reduce_max_sync(0xffffffffu, mx[a]); }
    live = __reduce_add_sync(0xffffffffu, live);
    int* sm = reinterpret_cast<int*>(&S.V.wcnt[0][0]);  // scratch, free until the first pass
    if (lane == 0) {
#pragma unroll
      for (int a = 0; a < 3; ++a) { sm[warp * 7 + a] = mn[a]; sm[warp * 7 + 3 + a] = mx[a]; }
      sm[warp * 7 + 6] = live;
    }
    __syncthreads();
    if (tid == 0) {
      int tl = 0;
      for (int w = 0; w < CW; ++w) {
        for (int a = 0; a < 3; ++a) { mn[a] = min(mn[a], sm[w * 7 + a]); mx[a] = max(mx[a], sm[w * 7 + 3 + a]); }
        tl += sm[w * 7 + 6];
      }
      for (int a = 0; a < 3; ++a) { S.vb[a] = mn[a]; S.vb[3 + a] = mx[a]; }
      S.vb[6] = tl;
    }
    cluster.sync();  // partial boxes and the transformed points are visible to the whole cluster
  }
  int vlo[3] = {INT_MAX, INT_MAX, INT_MAX}, vhi[3] = {INT_MIN, INT_MIN, INT_MIN};
  int n_live = 0;
  for (int c = 0; c < CL; ++c) {
    const int* rb = cluster.map_shared_rank(&S, c)->vb;
    for (int a = 0; a < 3; ++a) { vlo[a] = min(vlo[a], rb[a]); vhi[a] = max(vhi[a], rb[3 + a]); }
    n_live += rb[6];
  }
  RelKey rk;
  rk.setup(vlo, vhi, g.shift, n_live > 0);
  if (rank == 0 && tid == 0) {
    V.n_in = n; V.n_orph_in = n_orph; V.n_live = n_live; V.bits = rk.bits;
    V.n_live_all = 0;
    V.fb[0] = V.fb[1] = V.fb[2] = INT_MAX; V.fb[3] = V.fb[4] = V.fb[5] = INT_MIN;
    if (n_orph + n_src_all > J.cap_new) atomicOr(J.status, ST_SCAN_CAPACITY);
    if (!rk.ok) atomicOr(J.status, ST_KEY_RANGE);
  }
  __syncthreads();  // the scratch in S.V.wcnt is reused by the sort
  // ---- phase 1: keys + stable radix sort (cropped-out points carry the sentinel and end up behind the live ones) ----
  const float4* __restrict__ np = J.newpts;
  cluster_radix_sort(cluster, S.V, J.sort, wbeg, wend, rk.bits, [&](int i) {
    const float4 p = __ldcg(np + i);
    if (outside(p, lo, hi)) return rk.sentinel();
    int vx, vy, vz;
    voxel_of(p, g.inv_leaf, vx, vy, vz);
    return rk.key(vx, vy, vz);
  });
  // ---- phase 2: sorted copy of the live points and their 63-bit keys ----
  const uint2* __restrict__ pr = J.sort.pair[sort_passes(rk.bits, J.sort.npass) & 1];
  for (int j = rank * CT + tid; j < n_live; j += CL * CT) {
    const float4 p = __ldcg(np + __ldcg(pr + j).y);
    J.nsorted[j] = p;
    J.nkey[j] = key64_pt(p, g);
  }
  cluster.sync();  // no CTA may exit while a peer can still read its shared memory
}

// ------------------------------------------------------------------------------------------------
// map update, step 2: merge-path partition + sorted copy of the new points
// ------------------------------------------------------------------------------------------------
// Tile t owns merged positions [t * MERGE_TILE, (t + 1) * MERGE_TILE); part[t] = old-map elements before that diagonal
// (ties: old elements first — PCL sums a voxel in cloud order, and the old map precedes the appended points, EM:313-323).
// One warp per diagonal: 32-ary search, so ~3 rounds of two dependent loads instead of ~15.
__global__ void __launch_bounds__(256) k_merge_partition(const MergeJob* __restrict__ jobs, int gather) {
  const MergeJob& J = jobs[blockIdx.y];
  MergeVars& V = *J.mv;
  const bool bad = (*J.status & ST_KEY_RANGE) != 0;  // the new points could not be keyed: the call fails, nothing is merged
  const int n_old = bad ? 0 : *J.n_map, n_new = bad ? 0 : V.n_live;
  const int n_tot = n_old + n_new;
  const int n_tiles = (n_tot + MERGE_TILE - 1) / MERGE_TILE;
  const int res = sort_passes(V.bits, J.sort.npass) & 1;
  const uint32_t* __restrict__ val = J.sort.val[res];
  // sorted copy of the live new points and their keys (coalesced loads in the merge), unless k_new_cluster made it
  for (int j = blockIdx.x * 256 + threadIdx.x; gather && j < n_new; j += gridDim.x * 256) {
    const float4 p = J.newpts[val[j]];
    J.nsorted[j] = p;
    J.nkey[j] = key64_pt(p, J.g);
  }
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * 256 + threadIdx.x) >> 5, nw = (gridDim.x * 256) >> 5;
  for (int t = wid; t <= n_tiles && t <= J.max_tiles; t += nw) {
    const long long d = min((long long)t * MERGE_TILE, (long long)n_tot);
    int lo = (int)max(0ll, d - n_new), hi = (int)min(d, (long long)n_old);
    // smallest a in [lo, hi] with NOT(old[a] <= new[d - a - 1]); the predicate is true on a prefix
    while (lo < hi) {
      const int span = hi - lo;
      const int step = (span + 31) / 32;
      const int a = lo + lane * step;
      bool more = false;
      if (a < hi) {
        const unsigned long long ko = key64_pt(J.old_pts[a], J.g);
        // (the cluster sort leaves interleaved pairs and has already written nkey; the grid-wide sort leaves split arrays and
        // nkey is being written by this very kernel)
        const unsigned long long kn = gather ? key64_pt(J.newpts[val[d - a - 1]], J.g) : J.nkey[d - a - 1];
        more = ko <= kn;
      }
      const unsigned m = __ballot_sync(0xffffffffu, more);   // a prefix of the lanes
      const int f = __popc(m);                               // first lane whose probe says "enough old elements"
      const int nlo = f > 0 ? lo + (f - 1) * step + 1 : lo;
      const int nhi = (f < 32 && lo + f * step < hi) ? lo + f * step : hi;
      lo = nlo; hi = nhi;
    }
    if (lane == 0) J.part[t] = (uint32_t)lo;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    V.n_old = n_old; V.n_tot = n_tot; V.n_tiles = n_tiles;
    const int H = table_size(n_tot, J.hcap);
    V.hmask = H - 1;
    V.n_orph = 0;  // consumed by k_new_xform; the emitting pass appends the next ones
    V.old_unique = J.meta[1] == 0 ? 1 : 0;  // (k_merge_scan clears meta[1] between the two merge passes: latch it here)
    if (n_tiles > J.max_tiles) atomicOr(J.status, ST_MAP_CAPACITY);
  }
}

// ------------------------------------------------------------------------------------------------
// map update, step 3: the merge (counting pass, tile scan, emitting pass)
// ------------------------------------------------------------------------------------------------
constexpr int MPER = MERGE_TILE / MERGE_THREADS;  // merged positions per thread (consecutive)
constexpr int MWARPS = MERGE_THREADS / 32;
struct MergeShared {
  float4 pts[MERGE_TILE];                 // old part [0, a), new part [a, a + b); later the tile's output points
  unsigned long long keys[MERGE_TILE];    // their keys; later the output points' cell keys
  unsigned short order[MERGE_TILE];       // merged position -> staged index; later the list of cell heads
  unsigned char live[MERGE_TILE];
  uint32_t scan[MWARPS];
  float bb[MWARPS][7];
};

__device__ __forceinline__ uint32_t merge_block_scan(uint32_t v, uint32_t* buf, uint32_t* total) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t inc = v;
  for (int off = 1; off < 32; off <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, off); if (lane >= off) inc += t; }
  __syncthreads();
  if (lane == 31) buf[warp] = inc;
  __syncthreads();
  uint32_t woff = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < MWARPS; ++w) { const uint32_t c = buf[w]; if (w < warp) woff += c; tot += c; }
  *total = tot;
  return woff + inc - v;
}

template <bool EMIT>
__global__ void __launch_bounds__(MERGE_THREADS, 4) k_merge(const MergeJob* __restrict__ jobs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  MergeShared& S = *reinterpret_cast<MergeShared*>(smem_raw);
  const MergeJob& J = jobs[blockIdx.y];
  MergeVars& V = *J.mv;
  const int n_tiles = min(V.n_tiles, J.max_tiles);
  const int t = blockIdx.x;
  const int tid = threadIdx.x;
  if (n_tiles == 0) {  // nothing in, nothing out
    if (EMIT && t == 0 && tid == 0) { *J.n_map = 0; J.meta[0] = V.hmask; J.meta[1] = 0; }
    if (!EMIT && t == 0) {
      const int H = V.hmask + 1;
      for (int i = tid; i < H; i += MERGE_THREADS) J.table[i] = make_uint2(SLOT_EMPTY, 0u);
    }
    return;
  }
  if (t >= n_tiles) return;
  const int n_old = V.n_old, n_new = V.n_live, n_tot = V.n_tot;
  const CellGeom g = J.g;
  const int s3 = 3 * g.shift;
  const int a0 = (int)J.part[t], a1 = (int)J.part[t + 1];
  TileOut TO;
  if (EMIT) TO = J.tout[t];
  const int d0 = t * MERGE_TILE, d1 = min(d0 + MERGE_TILE, n_tot);
  const int b0 = d0 - a0, b1 = d1 - a1;
  const int na = a1 - a0, nb = b1 - b0, nt = na + nb;
  float lo[3], hi[3];
  merge_crop(J, lo, hi);
  // key of the merged element in front of the tile (run continuation) — ~0 when there is none; loads issued ahead of the staging
  unsigned long long prev_key = ~0ull;
  {
    float4 po = make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned long long kn = 0;
    if (a0 > 0) po = J.old_pts[a0 - 1];
    if (b0 > 0) kn = J.nkey[b0 - 1];
    if (a0 > 0) prev_key = key64_pt(po, g);
    if (b0 > 0) prev_key = (prev_key == ~0ull || kn > prev_key) ? kn : prev_key;
  }

  if (!EMIT) {  // the counting pass also clears the cell table of the map being written
    const int H = V.hmask + 1;
    const int per = (H + n_tiles - 1) / n_tiles;
    const int hb = min(H, t * per), he = min(H, hb + per);
    for (int i = hb + tid; i < he; i += MERGE_THREADS) J.table[i] = make_uint2(SLOT_EMPTY, 0u);
  }

  // ---- stage the tile: old part, new part; bounding box of the live points (PCL guard) ----
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  int nlive = 0;
  {
    float4 pb[MPER];
    unsigned long long kb[MPER];
#pragma unroll
    for (int it = 0; it < MPER; ++it) {  // all loads of the tile in flight before the first key is computed
      const int i = it * MERGE_THREADS + tid;
      kb[it] = 0;
      if (i < na) pb[it] = J.old_pts[a0 + i];
      else if (i < nt) { pb[it] = J.nsorted[b0 + i - na]; kb[it] = J.nkey[b0 + i - na]; }
    }
#pragma unroll
    for (int it = 0; it < MPER; ++it) {
      const int i = it * MERGE_THREADS + tid;
      if (i >= nt) continue;
      const float4 p = pb[it];
      bool lv = true;
      unsigned long long k = kb[it];
      if (i < na) { k = key64_pt(p, g); lv = !outside(p, lo, hi); }
      S.pts[i] = p; S.keys[i] = k; S.live[i] = lv ? 1 : 0;
      if (!EMIT && lv) {
        ++nlive;
        mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
        mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
        mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
      }
    }
  }
  __syncthreads();

  // ---- merged order: old i -> i + #(new < key), new j -> j + #(old <= key) ----
  if (nb == 0) {
    for (int i = tid; i < nt; i += MERGE_THREADS) S.order[i] = (unsigned short)i;
  } else {
    for (int i = tid; i < nt; i += MERGE_THREADS) {
      const unsigned long long k = S.keys[i];
      int r;
      if (i < na) {
        int l = 0, h = nb;  // lower_bound in the new part
        while (l < h) { const int m = (l + h) >> 1; if (S.keys[na + m] < k) l = m + 1; else h = m; }
        r = i + l;
      } else {
        int l = 0, h = na;  // upper_bound in the old part
        while (l < h) { const int m = (l + h) >> 1; if (S.keys[m] <= k) l = m + 1; else h = m; }
        r = (i - na) + l;
      }
      S.order[r] = (unsigned short)i;
    }
  }
  __syncthreads();

  // ---- every run head sums its voxel: PCL's `centroid += point` in cloud order, fp32, then `/= count` ----
  // thread `tid` owns the consecutive merged positions [tid * MPER, tid * MPER + MPER)
  float4 outp[MPER];
  unsigned long long outk[MPER];
  uint32_t hasm = 0;
  // Fast path (most tiles of a big map): no new point falls into the tile and the old map has one point per voxel, so every live
  // point is a complete voxel of its own: centroid = (0 + p) / 1.  A run is an old point followed by the new points of its voxel, so
  // the only run that can leave such a tile is the one of its last point, when the next tile begins with new points of that voxel:
  // then the general path below runs (the head of a run sums all of it, also what lies behind its tile).
  bool fast = nb == 0 && nt > 0 && V.old_unique != 0;
  if (fast && d1 < n_tot && b1 < n_new) fast = J.nkey[b1] != S.keys[nt - 1];
  if (fast) {
    const int r0 = tid * MPER;
#pragma unroll
    for (int it = 0; it < MPER; ++it) {
      const int r = r0 + it;
      outk[it] = 0;
      outp[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nt && S.live[r]) {
        const float4 q = S.pts[r];
        hasm |= 1u << it;
        outp[it] = make_float4(fadd(0.f, q.x), fadd(0.f, q.y), fadd(0.f, q.z), fadd(0.f, q.w));
        outk[it] = S.keys[r] >> s3;
      }
    }
  } else {
    const int r0 = tid * MPER;
    unsigned long long pk = r0 == 0 ? prev_key : (r0 - 1 < nt ? S.keys[S.order[r0 - 1]] : 0ull);
#pragma unroll
    for (int it = 0; it < MPER; ++it) {
      const int r = r0 + it;
      outk[it] = 0;
      outp[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r >= nt) continue;
      const unsigned long long k = S.keys[S.order[r]];
      const bool head = k != pk;  // otherwise it continues a run: its head (in this or an earlier tile) sums it
      pk = k;
      if (!head) continue;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int cnt = 0;
      int rr = r;
      for (; rr < nt; ++rr) {
        const int li = S.order[rr];
        if (S.keys[li] != k) break;
        if (S.live[li]) {
          const float4 q = S.pts[li];
          acc.x = fadd(acc.x, q.x); acc.y = fadd(acc.y, q.y); acc.z = fadd(acc.z, q.z); acc.w = fadd(acc.w, q.w);
          ++cnt;
        }
      }
      if (rr == nt && d1 < n_tot) {  // the run may go on behind the tile: the rest of its old points first, then its new ones
        for (int a = a1; a < n_old; ++a) {
          const float4 q = J.old_pts[a];
          if (key64_pt(q, g) != k) break;
          if (!outside(q, lo, hi)) { acc.x = fadd(acc.x, q.x); acc.y = fadd(acc.y, q.y); acc.z = fadd(acc.z, q.z); acc.w = fadd(acc.w, q.w); ++cnt; }
        }
        for (int b = b1; b < n_new; ++b) {
          if (J.nkey[b] != k) break;
          const float4 q = J.nsorted[b];
          acc.x = fadd(acc.x, q.x); acc.y = fadd(acc.y, q.y); acc.z = fadd(acc.z, q.z); acc.w = fadd(acc.w, q.w); ++cnt;
        }
      }
      if (cnt == 0) continue;  // every point of the voxel left the crop box
      const float4 c = centroid_of(acc, cnt);
      if (cnt > 1 && key64_pt(c, g) != k) {
        // fp32 rounding put the centroid across a face of its voxel: stored here it would break the map's order.  PCL would
        // count it into the neighbouring voxel at the next update; it is set aside and re-inserted with the next new points.
        if (EMIT) {
          const int o = atomicAdd(&V.n_orph, 1);
          if (o < ORPHAN_CAP) J.orphans[o] = c; else atomicOr(J.status, ST_ORPHANS);
        }
        continue;
      }
      hasm |= 1u << it; outp[it] = c; outk[it] = k >> s3;
    }
  }
  __syncthreads();  // all reads of the staged tile are done: its storage now takes the compacted outputs

  // ---- compact the outputs in merged order ----
  uint32_t tot_out;
  {
    uint32_t o = merge_block_scan((uint32_t)__popc(hasm), S.scan, &tot_out);
#pragma unroll
    for (int it = 0; it < MPER; ++it)
      if (hasm & (1u << it)) { S.pts[o] = outp[it]; S.keys[o] = outk[it]; ++o; }
  }
  const int count = (int)tot_out;
  __syncthreads();

  if (!EMIT) {
    // bounding box of the live points -> MergeVars (7 atomics per tile)
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], off));
        mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], off));
      }
      nlive += __shfl_xor_sync(0xffffffffu, nlive, off);
    }
    if ((tid & 31) == 0) {
#pragma unroll
      for (int a = 0; a < 3; ++a) { S.bb[tid >> 5][a] = mn[a]; S.bb[tid >> 5][3 + a] = mx[a]; }
      S.bb[tid >> 5][6] = __int_as_float(nlive);
    }
    __syncthreads();
    if (tid == 0) {
      int tl = 0;
      for (int w = 0; w < MWARPS; ++w) {
        for (int a = 0; a < 3; ++a) { mn[a] = fminf(mn[a], S.bb[w][a]); mx[a] = fmaxf(mx[a], S.bb[w][3 + a]); }
        tl += __float_as_int(S.bb[w][6]);
      }
      if (tl > 0) {
        for (int a = 0; a < 3; ++a) { atomicMin(&V.fb[a], f2ord(mn[a])); atomicMax(&V.fb[3 + a], f2ord(mx[a])); }
        atomicAdd(&V.n_live_all, tl);
      }
      TileAgg A;
      A.count = count;
      A.first_cell = count ? S.keys[0] : 0ull;
      A.last_cell = count ? S.keys[count - 1] : 0ull;
      int lrs = 0;
      for (int o = count - 1; o > 0; --o)
        if (S.keys[o] != S.keys[o - 1]) { lrs = o; break; }
      A.last_run_start = lrs;
      J.agg[t] = A;
    }
    return;
  }

  // ---- emitting pass: points, then the cell table (head c closes the cell in front of it) ----
  const int base = TO.base;
  const uint32_t hmask = (uint32_t)V.hmask;
  for (int o = tid; o < count; o += MERGE_THREADS)
    if (base + o < J.cap_out) J.out_pts[base + o] = S.pts[o];
  // cell heads of this tile, compacted in order: thread `tid` looks at outputs [tid * MPER, tid * MPER + MPER)
  uint32_t nheads;
  {
    const int o0 = tid * MPER;
    uint32_t hm = 0;
#pragma unroll
    for (int it = 0; it < MPER; ++it) {
      const int o = o0 + it;
      if (o >= count) continue;
      const bool head = o > 0 ? (S.keys[o] != S.keys[o - 1]) : (!TO.have_prev || S.keys[0] != TO.carry_cell);
      if (head) hm |= 1u << it;
    }
    uint32_t c = merge_block_scan((uint32_t)__popc(hm), S.scan, &nheads);
#pragma unroll
    for (int it = 0; it < MPER; ++it)
      if (hm & (1u << it)) S.order[c++] = (unsigned short)(o0 + it);
    __syncthreads();
  }
  for (int c = tid; c < (int)nheads; c += MERGE_THREADS) {
    const int o = S.order[c];
    if (c > 0) {
      const int po = S.order[c - 1];
      table_insert(J.table, hmask, S.keys[po], (uint32_t)(base + po), (uint32_t)(base + o));
    } else if (TO.have_prev) {
      table_insert(J.table, hmask, TO.carry_cell, (uint32_t)TO.carry_start, (uint32_t)(base + o));
    }
  }
  if (t == n_tiles - 1 && tid == 0) {  // the last tile closes the last cell
    const int total = base + count;
    if (nheads > 0) { const int po = S.order[nheads - 1]; table_insert(J.table, hmask, S.keys[po], (uint32_t)(base + po), (uint32_t)total); }
    else if (TO.have_prev) table_insert(J.table, hmask, TO.carry_cell, (uint32_t)TO.carry_start, (uint32_t)total);
  }
}

// Between the two passes, one CTA per map: where every tile's outputs go (exclusive prefix of the counts) and which cell run is
// open in front of it (cell key + index of that run's first point; a run may span several tiles).  Also publishes the size of the
// new map and evaluates PCL's "leaf size is too small" guard on the bounding box the counting pass reduced.
constexpr int MSCAN_THREADS = 1024;
__global__ void __launch_bounds__(MSCAN_THREADS) k_merge_scan(const MergeJob* __restrict__ jobs) {
  const MergeJob& J = jobs[blockIdx.x];
  MergeVars& V = *J.mv;
  const int n_tiles = min(V.n_tiles, J.max_tiles);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int* s_base = reinterpret_cast<int*>(smem_raw);            // [n_tiles + 1]
  int* s_prev = s_base + (J.max_tiles + 2);                   // last non-empty tile in front of tile t, or -1
  __shared__ int wsum[32], wmax[32];
  __shared__ int run_sum, run_max;
  if (tid == 0) { run_sum = 0; run_max = -1; }
  __syncthreads();
  for (int b = 0; b < n_tiles; b += MSCAN_THREADS) {
    const int t = b + tid;
    const int c = t < n_tiles ? J.agg[t].count : 0;
    int inc = c, mxi = c > 0 ? t : -1;
    for (int off = 1; off < 32; off <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, inc, off), m = __shfl_up_sync(0xffffffffu, mxi, off);
      if (lane >= off) { inc += u; mxi = max(mxi, m); }
    }
    if (lane == 31) { wsum[warp] = inc; wmax[warp] = mxi; }
    __syncthreads();
    int woff = 0, wm = -1;
    for (int w = 0; w < warp; ++w) { woff += wsum[w]; wm = max(wm, wmax[w]); }
    const int incl = run_sum + woff + inc;          // inclusive prefix of the counts
    const int last = max(run_max, max(wm, mxi));    // last non-empty tile <= t
    const int pl = __shfl_up_sync(0xffffffffu, last, 1);  // exclusive "last non-empty": the inclusive value of the previous tile
    if (t < n_tiles) {
      s_base[t] = incl - c;
      s_prev[t] = lane > 0 ? pl : max(run_max, wm);
    }
    __syncthreads();
    if (tid == MSCAN_THREADS - 1) { run_sum = incl; run_max = last; }
    __syncthreads();
  }
  const int total = run_sum;
  for (int t = tid; t < n_tiles; t += MSCAN_THREADS) {
    TileOut O;
    O.base = s_base[t]; O.have_prev = 0; O.carry_start = 0; O.carry_cell = 0;
    int p = s_prev[t];
    if (p >= 0) {
      O.have_prev = 1;
      O.carry_cell = J.agg[p].last_cell;
      for (;;) {  // walk back to the tile in which the run of carry_cell starts
        const TileAgg A = J.agg[p];
        if (A.last_run_start > 0) { O.carry_start = s_base[p] + A.last_run_start; break; }
        O.carry_start = s_base[p];  // the whole tile is one run of this cell: it may have begun earlier
        const int pp = s_prev[p];
        if (pp < 0 || J.agg[pp].last_cell != O.carry_cell) break;
        p = pp;
      }
    }
    J.tout[t] = O;
  }
  if (tid == 0) {
    if (total > J.cap_out) atomicOr(J.status, ST_MAP_CAPACITY);
    *J.n_map = min(total, J.cap_out);
    J.meta[0] = V.hmask; J.meta[1] = 0;
    // PCL's "leaf size is too small" guard (voxel_grid.hpp): the reference would hand the cloud through unfiltered
    if (V.n_live_all > 0) {
      const float inv = J.g.inv_leaf;
      long long prod = 1;
      for (int a = 0; a < 3; ++a) prod *= (long long)(fmul(fsub(ord2f(V.fb[3 + a]), ord2f(V.fb[a])), inv)) + 1;
      if (prod > (long long)INT_MAX) atomicOr(J.status, ST_PCL_GUARD);
    }
  }
}

cudaError_t init_cellmap_kernels() {
  cudaError_t e = cudaFuncSetAttribute(k_merge<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MergeShared));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_merge<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MergeShared));
}

void launch_cell_update(const Launch& L, const MergeJob* jobs_dev, const SortJob* sort_jobs_dev, int njobs, int max_tiles, bool cluster_new) {
  dim3 g(74, njobs);
  if (cluster_new) {
    dim3 gc(CL, njobs);
    k_new_cluster<<<gc, CT, 0, L.st>>>(jobs_dev);
    L.tick(K_NEW_XFORM);
  } else {
    k_merge_reset<<<(njobs + 127) / 128, 128, 0, L.st>>>(jobs_dev, njobs);
    L.tick(K_NEW_XFORM);
    k_new_xform<<<g, 256, 0, L.st>>>(jobs_dev);
    L.tick(K_NEW_XFORM);
    KeyGenNew gen;
    gen.jobs = jobs_dev;
    dim3 gs(SORT_G, njobs);
    k_sort_keyhist<KeyGenNew><<<gs, SORT_THREADS, 0, L.st>>>(sort_jobs_dev, gen);
    L.tick(K_NEW_KEYHIST);
    for (int pass = 0; pass < 4; ++pass) launch_sort_scatter(L, sort_jobs_dev, njobs, pass);
  }
  k_merge_partition<<<g, 256, 0, L.st>>>(jobs_dev, cluster_new ? 0 : 1);
  L.tick(K_MERGE_PART);
  dim3 gm(max_tiles, njobs);
  k_merge<false><<<gm, MERGE_THREADS, sizeof(MergeShared), L.st>>>(jobs_dev);
  L.tick(K_MERGE_COUNT);
  k_merge_scan<<<njobs, MSCAN_THREADS, (size_t)(2 * max_tiles + 8) * sizeof(int), L.st>>>(jobs_dev);
  L.tick(K_MERGE_PART);
  k_merge<true><<<gm, MERGE_THREADS, sizeof(MergeShared), L.st>>>(jobs_dev);
  L.tick(K_MERGE_EMIT);
}

// ------------------------------------------------------------------------------------------------
// 5-NN: one warp per query
// ------------------------------------------------------------------------------------------------
struct CellMapView {
  const float4* pts;
  const uint2* table;
  const uint32_t* orig;  // PCL index of every point, or null when the PCL order is the voxel order (any map after an update)
  uint32_t hmask;
  CellGeom g;
  int n;
};

// Is map point a before map point b in the reference's map order?  Only consulted on exact fp32 distance ties (tie class T2).
// (Inlined on purpose: as a __noinline__ callee taking the view by reference it made k_knn_cell_assoc fault — the view is then
// selected at run time between two local structs and the generic pointer to it was wrong in the callee.)
__device__ __forceinline__ bool pcl_before(const CellMapView& M, int a, int b) {
  if (b == INT_MAX) return true;
  if (a == INT_MAX) return false;
  if (M.orig) return M.orig[a] < M.orig[b];
  int ax, ay, az, bx, by, bz;
  voxel_of(M.pts[a], M.g.inv_leaf, ax, ay, az);
  voxel_of(M.pts[b], M.g.inv_leaf, bx, by, bz);
  if (az != bz) return az < bz;  // PCL's output order: ascending i + j * dx + k * dx * dy
  if (ay != by) return ay < by;
  if (ax != bx) return ax < bx;
  return a < b;
}
// The warp's candidate pool lives in shared memory: POOL_CAP (distance, index) pairs.  Top five of pool[0, npool), nearest first,
// replicated in every lane: lane l looks after entries l and l + 32; five rounds of redux.sync min on the distance bits
// (non-negative floats order like their bit patterns), exact ties resolved by the reference's map order.
constexpr int POOL_CAP = 64;
__device__ __forceinline__ void pool_top5(const CellMapView& M, const uint2* pool, int npool, float (&rd)[5], int (&ri)[5]) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  __syncwarp();
  uint2 e0 = make_uint2(0xFFFFFFFFu, (uint32_t)INT_MAX), e1 = e0;
  if (lane < npool) e0 = pool[lane];
  if (lane + 32 < npool) e1 = pool[lane + 32];
  // keep the lane's better entry in e0
  if (e1.x < e0.x || (e1.x == e0.x && e1.x != 0xFFFFFFFFu && pcl_before(M, (int)e1.y, (int)e0.y))) { const uint2 t = e0; e0 = e1; e1 = t; }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const uint32_t mnk = __reduce_min_sync(FULL, e0.x);
    if (mnk == 0xFFFFFFFFu) { rd[k] = FLT_MAX; ri[k] = INT_MAX; continue; }  // uniform: the pool is exhausted
    unsigned who = __ballot_sync(FULL, e0.x == mnk);
    int w = __ffs(who) - 1;
    int wi = __shfl_sync(FULL, (int)e0.y, w);
    who &= who - 1;
    while (who) {  // equal distances (tie class T2): the reference's map order decides
      const int w2 = __ffs(who) - 1;
      const int i2 = __shfl_sync(FULL, (int)e0.y, w2);
      if (pcl_before(M, i2, wi)) { w = w2; wi = i2; }
      who &= who - 1;
    }
    rd[k] = __uint_as_float(mnk); ri[k] = wi;
    if (lane == w) { e0 = e1; e1 = make_uint2(0xFFFFFFFFu, (uint32_t)INT_MAX); }
  }
}
// After a reduction the pool is its own top five.
__device__ __forceinline__ int pool_reseed(uint2* pool, const float (&rd)[5], const int (&ri)[5]) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  if (lane < 5) {
    const float d = lane == 0 ? rd[0] : lane == 1 ? rd[1] : lane == 2 ? rd[2] : lane == 3 ? rd[3] : rd[4];
    const int i = lane == 0 ? ri[0] : lane == 1 ? ri[1] : lane == 2 ? ri[2] : lane == 3 ? ri[3] : ri[4];
    pool[lane] = make_uint2(__float_as_uint(d), (uint32_t)i);
  }
  __syncwarp();
  return (ri[0] != INT_MAX) + (ri[1] != INT_MAX) + (ri[2] != INT_MAX) + (ri[3] != INT_MAX) + (ri[4] != INT_MAX);
}

// Called by a whole warp with one query; returns (replicated in every lane) the five nearest map points inside the gate in
// ascending (d^2, PCL index) order; FLT_MAX / INT_MAX where fewer than five lie inside the gate.  `pool` = POOL_CAP entries of
// shared memory owned by this warp.
//
// Candidates are read 32 at a time (coalesced runs: a cell is a contiguous range, the cells of an x-row follow each other).  The few
// that fall inside the gate are appended to the pool (ballot + popc rank, one shared-memory store each); the pool is reduced to its
// best five only when it would overflow, between shells (early exit) and at the end, and every reduction tightens the gate to
// the fifth distance known so far.
template <bool SEEDED>
__device__ __forceinline__ void warp_knn5(const CellMapView& M, float gate_f, float qx, float qy, float qz, uint2* pool, float (&bd)[5], int (&bi)[5],
                                          const int* seed = nullptr) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int s = M.g.shift;
  const float inv = M.g.inv_leaf;
  int vx, vy, vz;
  voxel_of(make_float4(qx, qy, qz, 0.f), inv, vx, vy, vz);
  const int qcx = (int)(vbias(vx) >> s), qcy = (int)(vbias(vy) >> s), qcz = (int)(vbias(vz) >> s);
  const int ncmax = (1 << (21 - s)) - 1;
  const int shells = M.g.shells;
  int npool = 0;
  float gate_dyn = gate_f;  // once five neighbours are known nothing farther than the fifth can matter (ties are kept: <=)
  // the query's position inside its own cell, in voxels: [0, k)
  const float kf = (float)(1 << s);
  const float ux = fmul(qx, inv) - (float)(((qcx << s) - VOX_BIAS)), uy = fmul(qy, inv) - (float)(((qcy << s) - VOX_BIAS)),
              uz = fmul(qz, inv) - (float)(((qcz << s) - VOX_BIAS));
  const float leaf2 = M.g.leaf * M.g.leaf;
  // Seeds: five map points known to be near the query (its neighbours of the previous outer iteration: the pose moved by millimetres,
  // the map not at all).  Their largest distance bounds the fifth-nearest distance from above, so the search only has to look at the
  // cells that intersect that ball — typically 1-4 of the 27 on a dense map.  The seeds themselves are found again in those cells.
  if (SEEDED && seed) {
    const int si = lane < 5 ? __ldg(seed + lane) : 0;
    if (__all_sync(FULL, si >= 0 && si < M.n)) {
      float sd = 0.f;
      if (lane < 5) {
        const float4 p = __ldg(M.pts + si);
        const float ddx = fsub(qx, p.x), ddy = fsub(qy, p.y), ddz = fsub(qz, p.z);
        sd = fadd(fadd(fmul(ddx, ddx), fmul(ddy, ddy)), fmul(ddz, ddz));
      }
      gate_dyn = fminf(gate_dyn, __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(sd))));
    }
  }
  for (int r = 1; r <= shells; ++r) {
    // cells of Chebyshev distance exactly r (r = 1: the whole 27-cell cube): the two z faces, then the square rings of the layers between
    const int side = 2 * r + 1, face = side * side, ring = 8 * r;
    const int ncell = r == 1 ? 27 : 2 * face + (side - 2) * ring;
    for (int cb = 0; cb < ncell; cb += 32) {
      const int t = cb + lane;
      uint32_t s0 = 0, e0 = 0;
      if (t < ncell) {
        int dx, dy, dz;
        if (r == 1) {
          dx = t % 3 - 1; dy = (t / 3) % 3 - 1; dz = t / 9 - 1;
        } else if (t < 2 * face) {
          const int u = t < face ? t : t - face;
          dx = u % side - r; dy = u / side - r; dz = t < face ? -r : r;
        } else {
          const int u = t - 2 * face, v = u % ring, sd = v / (2 * r), off = v % (2 * r);
          dz = u / ring - r + 1;
          dx = sd == 0 ? -r + off : sd == 1 ? r : sd == 2 ? r - off : -r;
          dy = sd == 0 ? -r : sd == 1 ? -r + off : sd == 2 ? r : r - off;
        }
        const int cx = qcx + dx, cy = qcy + dy, cz = qcz + dz;
        // squared distance from the query to the cell's box: a cell farther than the current bound holds nothing of interest (0.1 %
        // margin for the rounding of the voxel coordinates; cells are skipped only when clearly outside)
        const float gx = dx > 0 ? (float)dx * kf - ux : dx < 0 ? ux - (float)(dx + 1) * kf : 0.f;
        const float gy = dy > 0 ? (float)dy * kf - uy : dy < 0 ? uy - (float)(dy + 1) * kf : 0.f;
        const float gz = dz > 0 ? (float)dz * kf - uz : dz < 0 ? uz - (float)(dz + 1) * kf : 0.f;
        const bool reachable = !SEEDED || (gx * gx + gy * gy + gz * gz) * leaf2 * 0.999f <= gate_dyn;  // (the unseeded search visits every cell of a shell:
                                                                                                         // measured, the test costs it more than it saves)
        if (reachable && cx >= 0 && cy >= 0 && cz >= 0 && cx <= ncmax && cy <= ncmax && cz <= ncmax) {
          const unsigned long long ck = cellkey_cells((uint32_t)cx, (uint32_t)cy, (uint32_t)cz, s);
          uint32_t h = cell_slot(ck) & M.hmask;
          for (;;) {
            const uint2 e = __ldg(M.table + h);
            if (e.x == SLOT_EMPTY) break;
            if ((key64_pt(__ldg(M.pts + e.x), M.g) >> (3 * s)) == ck) { s0 = e.x; e0 = e.y; break; }
            h = (h + 1) & M.hmask;
          }
        }
      }
      const uint32_t cnt = e0 - s0;
      if (!__any_sync(FULL, cnt != 0u)) continue;  // a batch of empty cells (outer shells of a sparse neighbourhood)
      uint32_t inc = cnt;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) { const uint32_t u = __shfl_up_sync(FULL, inc, off); if (lane >= off) inc += u; }
      const uint32_t total = __shfl_sync(FULL, inc, 31);
      const uint32_t rebase = s0 - (inc - cnt);  // map index of the range's first point minus its position in the concatenation
      for (uint32_t c0 = 0; c0 < total; c0 += 32) {
        const uint32_t c = c0 + lane;
        int l = 0, h = 31;  // owner = first lane whose inclusive count exceeds c
        // (a forward-moving owner pointer instead of this search was measured: 3 % faster on the planar 1e6-point maps, 27 % slower on
        // the uniform random cloud of the stage bench, where a batch of 32 candidates spans ~20 ranges: gpurun_out/r2q_*.json)
#pragma unroll
        for (int it = 0; it < 5; ++it) {
          const int m = (l + h) >> 1;
          const uint32_t v = __shfl_sync(FULL, inc, m);
          if (v > c) h = m; else l = m + 1;
        }
        const int idx = (int)(__shfl_sync(FULL, rebase, l) + c);
        bool cand = c < total;
        float cd = FLT_MAX;
        if (cand) {
          const float4 p = __ldg(M.pts + idx);
          const float ddx = fsub(qx, p.x), ddy = fsub(qy, p.y), ddz = fsub(qz, p.z);
          cd = fadd(fadd(fmul(ddx, ddx), fmul(ddy, ddy)), fmul(ddz, ddz));  // FLANN L2_Simple<float>
          cand = cd < gate_f && cd <= gate_dyn;  // neighbours at or beyond the gate can never be used (EM:129 / :189)
        }
        const unsigned m = __ballot_sync(FULL, cand);
        if (m == 0u) continue;
        const int cm = __popc(m);
        if (npool + cm > POOL_CAP) {  // reduce the pool to its best five (npool <= 32 + 5 afterwards)
          float rd[5]; int ri[5];
          pool_top5(M, pool, npool, rd, ri);
          npool = pool_reseed(pool, rd, ri);
          gate_dyn = fminf(gate_dyn, rd[4]);  // (rd[4] is FLT_MAX while fewer than five are known)
        }
        if (cand) pool[npool + __popc(m & lt)] = make_uint2(__float_as_uint(cd), (uint32_t)idx);
        npool += cm;
      }
    }
    if (r < shells) {  // after shell r every unseen point is farther than (r + f) cells: is the fifth best already closer?
      float rd[5]; int ri[5];
      pool_top5(M, pool, npool, rd, ri);
      const float f = fminf(fminf(fminf(ux, kf - ux), fminf(uy, kf - uy)), fminf(uz, kf - uz));  // voxels to the nearest face of the query's cell
      const float fmin_cells = fmaxf(0.f, f / kf - 1e-4f);
      const double cell = (double)M.g.leaf * (double)(1 << s) * (1.0 - 1e-5);  // a lower bound of the cell edge (inverse_leaf is rounded)
      const double reach = ((double)r + (double)fmin_cells) * cell;
      if ((double)rd[4] < reach * reach * (1.0 - 1e-5)) {
#pragma unroll
        for (int q = 0; q < 5; ++q) { bd[q] = rd[q]; bi[q] = ri[q]; }
        return;
      }
      npool = pool_reseed(pool, rd, ri);
      gate_dyn = fminf(gate_dyn, rd[4]);
    }
  }
  pool_top5(M, pool, npool, bd, bi);
}

// ------------------------------------------------------------------------------------------------
// 5-NN: EIGHT lanes per query — an A/B variant (VILF_KNN_GROUP8=1), NOT the default
// ------------------------------------------------------------------------------------------------
// The warp-per-query search above costs ~1.3 k (sparse maps) to ~3.3 k (1e6-point maps, 0.72 m cells) warp instructions per query and
// is issue bound (profiles/r2m_dense_frame_S1.txt).  This variant gives a query to a group of eight lanes (four queries per warp), as
// the hashed-grid search of k_knn.cu does: lane sl probes cells sl, sl + 8, ... of a shell and walks ITS cells' contiguous ranges
// alone, keeping a private top five in registers; the eight lists are merged at the end by five rounds of an 8-lane arg-min.  It
// executes far fewer instructions per query but was measured SLOWER on the 1e6-point maps (gpurun_out/r2n_dense*.json: 180 us per
// search launch against 110 us; 0.32 G queries/s against 1.0 G on the synthetic 1e6 / 2.6e5 stage): every lane streams its own
// 16-byte points (32 sectors per warp load instead of 4 for the coalesced batches of the warp search), the trip count of a warp is
// the longest of its 32 per-lane walks, and 104 registers leave 16 warps per SM to hide the load latency.  Kept because it is exact
// (bit-identical results, same tests) and documents the experiment.
struct Top5C {
  float d[5];
  int id[5];
};
__device__ __forceinline__ bool closer_cell(const CellMapView& M, float d, int id, float bd, int bid) {
  if (d < bd) return true;
  if (d > bd || d == FLT_MAX) return false;
  return pcl_before(M, id, bid);  // exact fp32 tie (class T2): the reference's map order decides
}
__device__ __forceinline__ void top5c_insert(const CellMapView& M, Top5C& t, float cd, int ci) {  // precondition: closer than t[4]
  t.d[4] = cd; t.id[4] = ci;
#pragma unroll
  for (int k = 4; k > 0; --k) {
    const bool sw = closer_cell(M, t.d[k], t.id[k], t.d[k - 1], t.id[k - 1]);
    const float dk = sw ? t.d[k - 1] : t.d[k], dk1 = sw ? t.d[k] : t.d[k - 1];
    const int ik = sw ? t.id[k - 1] : t.id[k], ik1 = sw ? t.id[k] : t.id[k - 1];
    t.d[k] = dk; t.d[k - 1] = dk1; t.id[k] = ik; t.id[k - 1] = ik1;
  }
}
__device__ __forceinline__ void cell_consider(const CellMapView& M, Top5C& best, float gate_f, float qx, float qy, float qz, const float4 c, int ci) {
  const float ddx = fsub(qx, c.x), ddy = fsub(qy, c.y), ddz = fsub(qz, c.z);
  const float cd = fadd(fadd(fmul(ddx, ddx), fmul(ddy, ddy)), fmul(ddz, ddz));  // FLANN L2_Simple<float>
  if (cd < gate_f && cd <= best.d[4]) {
    if (closer_cell(M, cd, ci, best.d[4], best.id[4])) top5c_insert(M, best, cd, ci);
  }
}
// [start, end) of cell (cx, cy, cz), or an empty range
__device__ __forceinline__ void cell_lookup(const CellMapView& M, int cx, int cy, int cz, int ncmax, uint32_t& s0, uint32_t& e0) {
  s0 = 0; e0 = 0;
  if (cx < 0 || cy < 0 || cz < 0 || cx > ncmax || cy > ncmax || cz > ncmax) return;
  const int s = M.g.shift;
  const unsigned long long ck = cellkey_cells((uint32_t)cx, (uint32_t)cy, (uint32_t)cz, s);
  uint32_t h = cell_slot(ck) & M.hmask;
  for (;;) {
    const uint2 e = __ldg(M.table + h);
    if (e.x == SLOT_EMPTY) return;
    if ((key64_pt(__ldg(M.pts + e.x), M.g) >> (3 * s)) == ck) { s0 = e.x; e0 = e.y; return; }
    h = (h + 1) & M.hmask;
  }
}
__device__ __forceinline__ void shell_cell(int r, int t, int& dx, int& dy, int& dz) {  // cell t of the shell of Chebyshev radius r >= 2
  const int side = 2 * r + 1, face = side * side, ring = 8 * r;
  if (t < 2 * face) {
    const int u = t < face ? t : t - face;
    dx = u % side - r; dy = u / side - r; dz = t < face ? -r : r;
  } else {
    const int u = t - 2 * face, v = u % ring, sd = v / (2 * r), off = v % (2 * r);
    dz = u / ring - r + 1;
    dx = sd == 0 ? -r + off : sd == 1 ? r : sd == 2 ? r - off : -r;
    dy = sd == 0 ? -r : sd == 1 ? -r + off : sd == 2 ? r : r - off;
  }
}

constexpr int KG = 8;  // lanes per query
// Called by all 32 lanes; the lanes of a group (lane >> 3) pass the same query and the same map.  Result replicated in the group.
__device__ __forceinline__ void group_knn5_cell(const CellMapView& M, float gate_f, float qx, float qy, float qz, bool active, float (&rd)[5], int (&ri)[5]) {
  const unsigned FULL = 0xffffffffu;
  const int sl = threadIdx.x & (KG - 1);
  Top5C best;
#pragma unroll
  for (int k = 0; k < 5; ++k) { best.d[k] = FLT_MAX; best.id[k] = INT_MAX; }
  const int s = M.g.shift;
  const int ncmax = (1 << (21 - s)) - 1;
  int qcx = 0, qcy = 0, qcz = 0;
  if (active) {
    int vx, vy, vz;
    voxel_of(make_float4(qx, qy, qz, 0.f), M.g.inv_leaf, vx, vy, vz);
    qcx = (int)(vbias(vx) >> s); qcy = (int)(vbias(vy) >> s); qcz = (int)(vbias(vz) >> s);
    // shell 1 = the 27-cell cube: lane sl owns cells sl, sl + 8, sl + 16, sl + 24; the four probes are independent
    uint32_t s0[4], e0[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = sl + KG * j;
      s0[j] = 0; e0[j] = 0;
      if (c < 27) cell_lookup(M, qcx + c % 3 - 1, qcy + (c / 3) % 3 - 1, qcz + c / 9 - 1, ncmax, s0[j], e0[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      for (uint32_t p = s0[j]; p < e0[j]; ++p) cell_consider(M, best, gate_f, qx, qy, qz, __ldg(M.pts + p), (int)p);
  }
  const int shells = M.g.shells;
  const int rmax = __reduce_max_sync(FULL, active ? shells : 1);
  if (rmax > 1) {
    float fmin_cells = 0.f;
    if (active) {
      const float k = (float)(1 << s), inv = M.g.inv_leaf;
      const float ux = fmul(qx, inv) - (float)(((qcx << s) - VOX_BIAS)), uy = fmul(qy, inv) - (float)(((qcy << s) - VOX_BIAS)),
                  uz = fmul(qz, inv) - (float)(((qcz << s) - VOX_BIAS));
      const float f = fminf(fminf(fminf(ux, k - ux), fminf(uy, k - uy)), fminf(uz, k - uz));  // voxels to the nearest face of the query's cell
      fmin_cells = fmaxf(0.f, f / k - 1e-4f);
    }
    const double cell = (double)M.g.leaf * (double)(1 << s) * (1.0 - 1e-5);  // a lower bound of the cell edge (inverse_leaf is rounded)
    for (int r = 1; r < rmax; ++r) {  // shells 1..r are done; is shell r + 1 needed?
      float ub = best.d[4];           // a lane's own fifth distance bounds the group's from above
#pragma unroll
      for (int off = 1; off < KG; off <<= 1) ub = fminf(ub, __shfl_xor_sync(FULL, ub, off));
      const double reach = ((double)r + (double)fmin_cells) * cell;
      const bool done = !active || r >= shells || (double)ub < reach * reach * (1.0 - 1e-5);
      if (__all_sync(FULL, done)) break;
      if (!done) {
        const int R = r + 1, side = 2 * R + 1;
        const int ncell = 2 * side * side + (side - 2) * 8 * R;
        for (int t = sl; t < ncell; t += KG) {
          int dx, dy, dz;
          shell_cell(R, t, dx, dy, dz);
          uint32_t a, b;
          cell_lookup(M, qcx + dx, qcy + dy, qcz + dz, ncmax, a, b);
          for (uint32_t p = a; p < b; ++p) cell_consider(M, best, gate_f, qx, qy, qz, __ldg(M.pts + p), (int)p);
        }
      }
    }
  }
#pragma unroll
  for (int round = 0; round < 5; ++round) {  // merge the eight private lists: group arg-min, the owner pops
    float wd = best.d[0];
    int wi = best.id[0];
#pragma unroll
    for (int off = 1; off < KG; off <<= 1) {
      const float od = __shfl_xor_sync(FULL, wd, off);
      const int oi = __shfl_xor_sync(FULL, wi, off);
      const bool take = active && oi != INT_MAX && (wi == INT_MAX || closer_cell(M, od, oi, wd, wi));
      wd = take ? od : wd;
      wi = take ? oi : wi;
    }
    rd[round] = wd; ri[round] = wi;
    const bool pop = (best.id[0] == wi) & (wi != INT_MAX);
#pragma unroll
    for (int k = 0; k < 4; ++k) { best.d[k] = pop ? best.d[k + 1] : best.d[k]; best.id[k] = pop ? best.id[k + 1] : best.id[k]; }
    best.d[4] = pop ? FLT_MAX : best.d[4];
    best.id[4] = pop ? INT_MAX : best.id[4];
  }
}

constexpr int KC_THREADS = 256;
#ifndef VILF_KC_MINB
#define VILF_KC_MINB 4
#endif
constexpr int KC_MINB = VILF_KC_MINB;  // resident CTAs per SM the warp search is compiled for: 4 caps it at 64 registers (16 B spilled) and was measured
                                       // against 3 (80 registers): +12 % scans/s on the configs[2] workload with four stream groups (gpurun_out/r2t_dense*.json)

__global__ void __launch_bounds__(KC_THREADS) k_knn_cell8_assoc(LaneDev* lanes, int lane0, int cur, ConfigDev cfg, const double* pose_override) {
  const int ln = lane0 + blockIdx.y;
  const LaneDev& L = lanes[ln];
  LaneVars& V = *L.v;
  const int me = V.n_map[0], ms = V.n_map[1];
  if (!(me > 10 && ms > 50)) return;  // EM:254
  const int ne = V.n_ds[0], ns = V.n_ds[1];
  const int nq = ne + ns;
  if (blockIdx.x == 0 && threadIdx.x == 0) V.opt_ran = 1;
  if ((int)blockIdx.x * 32 >= nq) return;
  double x[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) x[i] = pose_override ? pose_override[i] : V.x[i];
  CellMapView Me, Ms;
  Me.pts = L.map[0][cur]; Me.table = L.ctab[0]; Me.hmask = (uint32_t)L.cmeta[0][0]; Me.orig = L.cmeta[0][1] ? L.corig[0] : nullptr; Me.g = cfg.cg[0]; Me.n = me;
  Ms.pts = L.map[1][cur]; Ms.table = L.ctab[1]; Ms.hmask = (uint32_t)L.cmeta[1][0]; Ms.orig = L.cmeta[1][1] ? L.corig[1] : nullptr; Ms.g = cfg.cg[1]; Ms.n = ms;
  // A CTA takes 32 consecutive queries per round: warp 0 runs the fp64 transform (EM:355-363) with one query per lane, then each of
  // the 32 eight-lane groups searches one of them.
  __shared__ float4 sq[32];
  const int grp = threadIdx.x >> 3, sl = threadIdx.x & 7;
  for (int qb = blockIdx.x * 32; qb < nq; qb += gridDim.x * 32) {
    __syncthreads();
    if (threadIdx.x < 32) {
      const int myq = qb + (int)threadIdx.x;
      if (myq < nq) sq[threadIdx.x] = associate(x, myq < ne ? L.ds[0][myq] : L.ds[1][myq - ne]);
    }
    __syncthreads();
    const int q = qb + grp;
    const bool active = q < nq;
    const float4 pw = active ? sq[grp] : make_float4(0.f, 0.f, 0.f, 0.f);
    const int w = (active && q >= ne) ? 1 : 0;
    const int k = w ? q - ne : q;
    float rd[5];
    int ri[5];
    // a warp holds four groups; at the edge / surf boundary they may use different maps: run the search once per map in use
    const unsigned want1 = __ballot_sync(0xffffffffu, active && w == 1), want0 = __ballot_sync(0xffffffffu, active && w == 0);
#pragma unroll
    for (int t = 0; t < 5; ++t) { rd[t] = FLT_MAX; ri[t] = INT_MAX; }
    if (want0) {
      float d0[5]; int i0[5];
      group_knn5_cell(Me, cfg.knn_gate_f, pw.x, pw.y, pw.z, active && w == 0, d0, i0);
      if (active && w == 0) {
#pragma unroll
        for (int t = 0; t < 5; ++t) { rd[t] = d0[t]; ri[t] = i0[t]; }
      }
    }
    if (want1) {
      float d1[5]; int i1[5];
      group_knn5_cell(Ms, cfg.knn_gate_f, pw.x, pw.y, pw.z, active && w == 1, d1, i1);
      if (active && w == 1) {
#pragma unroll
        for (int t = 0; t < 5; ++t) { rd[t] = d1[t]; ri[t] = i1[t]; }
      }
    }
    if (active) {
#pragma unroll
      for (int t = 0; t < 5; ++t)
        if (sl == t) {
          L.nn_idx[w][k * 5 + t] = ri[t] == INT_MAX ? -1 : ri[t];
          L.nn_d2[w][k * 5 + t] = rd[t];
        }
    }
  }
}


template <bool SEEDED>
__global__ void __launch_bounds__(KC_THREADS, KC_MINB) k_knn_cell_assoc(LaneDev* lanes, int lane0, int cur, ConfigDev cfg, const double* pose_override) {
  const int ln = lane0 + blockIdx.y;
  const LaneDev& L = lanes[ln];
  LaneVars& V = *L.v;
  const int me = V.n_map[0], ms = V.n_map[1];
  if (!(me > 10 && ms > 50)) return;  // EM:254
  const int ne = V.n_ds[0], ns = V.n_ds[1];
  const int nq = ne + ns;
  if (blockIdx.x == 0 && threadIdx.x == 0) V.opt_ran = 1;
  if ((int)blockIdx.x * 32 >= nq) return;  // the grid is sized for a full scan; most CTAs have nothing to do
  double x[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) x[i] = pose_override ? pose_override[i] : V.x[i];
  const int lane = threadIdx.x & 31;
  const int wpb = KC_THREADS / 32;
  CellMapView Me, Ms;
  Me.pts = L.map[0][cur]; Me.table = L.ctab[0]; Me.hmask = (uint32_t)L.cmeta[0][0]; Me.orig = L.cmeta[0][1] ? L.corig[0] : nullptr; Me.g = cfg.cg[0]; Me.n = me;
  Ms.pts = L.map[1][cur]; Ms.table = L.ctab[1]; Ms.hmask = (uint32_t)L.cmeta[1][0]; Ms.orig = L.cmeta[1][1] ? L.corig[1] : nullptr; Ms.g = cfg.cg[1]; Ms.n = ms;
  // A CTA takes 32 consecutive queries: warp 0 runs the fp64 transform (EM:355-363) with one query per lane (the FP64 pipe is
  // narrow: 32 lanes repeating the same transform cost as much as 32 different ones), then every warp searches four of them.
  __shared__ float4 sq[32];
  __shared__ uint2 spool[KC_THREADS / 32][POOL_CAP];
  for (int qb = blockIdx.x * 32; qb < nq; qb += gridDim.x * 32) {
    __syncthreads();  // sq of the previous round is no longer read
    if (threadIdx.x < 32) {
      const int myq = qb + lane;
      if (myq < nq) sq[lane] = associate(x, myq < ne ? L.ds[0][myq] : L.ds[1][myq - ne]);
    }
    __syncthreads();
    const int nb = min(32, nq - qb);
    for (int j = threadIdx.x >> 5; j < nb; j += wpb) {
      const int q = qb + j;
      const float4 pw = sq[j];
      const int w = q >= ne ? 1 : 0;
      const int k = w ? q - ne : q;
      float rd[5];
      int ri[5];
      const CellMapView M = w ? Ms : Me;
      // seeded: the neighbours the previous outer iteration stored for this very feature bound the search (same map, pose moved by mm)
      warp_knn5<SEEDED>(M, cfg.knn_gate_f, pw.x, pw.y, pw.z, spool[threadIdx.x >> 5], rd, ri, SEEDED ? L.nn_idx[w] + k * 5 : nullptr);
#pragma unroll
      for (int t = 0; t < 5; ++t)
        if (lane == t) {
          L.nn_idx[w][k * 5 + t] = ri[t] == INT_MAX ? -1 : ri[t];
          L.nn_d2[w][k * 5 + t] = rd[t];
        }
    }
  }
}

void launch_knn_cell_fit(const Launch& L, LaneDev* lanes, int lane0, int nlanes, int cur, const ConfigDev& cfg, const double* pose_override, bool seeded) {
  dim3 g(KNN_G * 8, nlanes);
  static const bool group_search = getenv("VILF_KNN_GROUP8") != nullptr;  // A/B: eight lanes per query (measured slower, see above)
  if (group_search) k_knn_cell8_assoc<<<g, KC_THREADS, 0, L.st>>>(lanes, lane0, cur, cfg, pose_override);
  else if (seeded) k_knn_cell_assoc<true><<<g, KC_THREADS, 0, L.st>>>(lanes, lane0, cur, cfg, pose_override);
  else k_knn_cell_assoc<false><<<g, KC_THREADS, 0, L.st>>>(lanes, lane0, cur, cfg, pose_override);
  L.tick(K_KNN_CELL);
  launch_fit(L, lanes, lane0, nlanes, cur, cfg);
}

// nearestKSearch alone against an explicit map (vilf_knn5); indices are reported in the caller's map order.
__global__ void __launch_bounds__(KC_THREADS, KC_MINB) k_knn_cell_only(const float4* __restrict__ pts, const int* n_dev, const uint2* __restrict__ table, const int* meta,
                                                               const uint32_t* __restrict__ orig, CellGeom g, const float4* __restrict__ q, const int* nq_dev,
                                                               int* idx, float* d2, float gate_f) {
  const int nq = *nq_dev;
  const int lane = threadIdx.x & 31;
  const int wpb = KC_THREADS / 32;
  CellMapView M;
  M.pts = pts; M.table = table; M.hmask = (uint32_t)meta[0]; M.orig = meta[1] ? orig : nullptr; M.g = g; M.n = *n_dev;
  const bool empty = *n_dev == 0;
  __shared__ uint2 spool[KC_THREADS / 32][POOL_CAP];
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < nq; i += gridDim.x * wpb) {
    const float4 p = q[i];
    float rd[5];
    int ri[5];
    if (empty) {
#pragma unroll
      for (int j = 0; j < 5; ++j) { rd[j] = FLT_MAX; ri[j] = INT_MAX; }
    } else {
      warp_knn5<false>(M, gate_f, p.x, p.y, p.z, spool[threadIdx.x >> 5], rd, ri);
    }
#pragma unroll
    for (int j = 0; j < 5; ++j)
      if (lane == j) {
        idx[i * 5 + j] = ri[j] == INT_MAX ? -1 : (M.orig ? (int)M.orig[ri[j]] : ri[j]);
        d2[i * 5 + j] = rd[j];
      }
  }
}
__global__ void __launch_bounds__(KC_THREADS) k_knn_cell8_only(const float4* __restrict__ pts, const int* n_dev, const uint2* __restrict__ table, const int* meta,
                                                                const uint32_t* __restrict__ orig, CellGeom g, const float4* __restrict__ q, const int* nq_dev,
                                                                int* idx, float* d2, float gate_f) {
  const int nq = *nq_dev;
  const int sl = threadIdx.x & 7;
  const int gpb = KC_THREADS / KG;
  CellMapView M;
  M.pts = pts; M.table = table; M.hmask = (uint32_t)meta[0]; M.orig = meta[1] ? orig : nullptr; M.g = g; M.n = *n_dev;
  const bool empty = *n_dev == 0;
  const int rounds = (nq + gridDim.x * gpb - 1) / (gridDim.x * gpb);
  for (int it = 0; it < rounds; ++it) {  // whole warps stay in the loop: the search uses full-mask shuffles
    const int i = (it * gridDim.x + blockIdx.x) * gpb + (threadIdx.x >> 3);
    const bool active = i < nq && !empty;
    const float4 p = i < nq ? q[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float rd[5];
    int ri[5];
    group_knn5_cell(M, gate_f, p.x, p.y, p.z, active, rd, ri);
    if (i < nq) {
#pragma unroll
      for (int j = 0; j < 5; ++j)
        if (sl == j) {
          const bool none = empty || ri[j] == INT_MAX;
          idx[i * 5 + j] = none ? -1 : (M.orig ? (int)M.orig[ri[j]] : ri[j]);
          d2[i * 5 + j] = none ? FLT_MAX : rd[j];
        }
    }
  }
}
void launch_knn_cell_only(const Launch& L, const float4* pts, const int* n_dev, const uint2* table, const int* meta, const uint32_t* orig, CellGeom g,
                          const float4* q, const int* nq_dev, int* idx, float* d2, const ConfigDev& cfg) {
  static const bool group_search = getenv("VILF_KNN_GROUP8") != nullptr;  // A/B: eight lanes per query (measured slower)
  if (group_search) k_knn_cell8_only<<<KNN_G * 8, KC_THREADS, 0, L.st>>>(pts, n_dev, table, meta, orig, g, q, nq_dev, idx, d2, cfg.knn_gate_f);
  else k_knn_cell_only<<<KNN_G * 8, KC_THREADS, 0, L.st>>>(pts, n_dev, table, meta, orig, g, q, nq_dev, idx, d2, cfg.knn_gate_f);
  L.tick(K_KNN_CELL);
}

}  // namespace vilf
