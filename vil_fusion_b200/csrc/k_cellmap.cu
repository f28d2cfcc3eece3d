// Cell-ordered local maps: ONE spatial order for pcl::VoxelGrid (EM:347-350), pcl::CropBox (EM:335-344), createSubMap's append
// (EM:308-324) and pcl::KdTreeFLANN (EM:256-257, :128, :185).
//
// Order.  Every map point has the voxel coordinate v = floor(p * inverse_leaf) PCL's filter gives it (fp32, voxel_grid.hpp);
// with k = 2^shift voxels per search-cell edge the 63-bit key is (cell z, cell y, cell x, voxel-in-cell z, y, x).  A voxel is a
// run of equal keys, a search cell is a contiguous range, and the three cells of an x-row are adjacent in memory.
//
// Update (k_new_xform .. k_merge<EMIT>).  The previous map is already in key order and voxel-filtered; a frame adds a few
// thousand points.  So instead of re-sorting ~1e6 points (the radix path: 3-4 passes of 16 B per point plus a gather) only the
// new points are sorted, and the map update is a MERGE: merge-path tiles of 2048 elements, each staged in shared memory once,
// crop box applied on the fly, every voxel run summed sequentially in fp32 in PCL's order (old points first, then the new
// ones, in input order), output = the other map buffer.  A counting pass sizes the tiles' outputs (no atomics, no look-back
// spinning); the emitting pass writes points and the cell table.  Algorithmic bytes: 16 B read per old point + 16 B written
// per kept voxel; the second read of the old map comes from L2.
//
// Search (k_knn_cell*).  One WARP per query: lane l probes neighbour cell l of the 27 (open addressing, verified by the cell of
// the first point of the range, which is a candidate anyway); the candidate ranges are concatenated with a warp scan and read
// 32 at a time as coalesced float4; selection of the five best uses redux.sync min on the distance bits — no per-lane lists, no
// merge tree.  Finer cells than the gate radius (dense maps) are walked shell by shell with the early exit of k_knn.cu.
//
// The PCL order of the points (ascending voxel index) differs from the stored order; everything the reference's results depend
// on is order independent or handled explicitly: voxel sums run in input order inside a voxel (= PCL's stable order), exact
// distance ties are resolved by PCL rank (recomputed from the two points on the rare tie), and the API boundary
// (vilf_get_cloud, vilf_factors) converts to PCL order on demand.
#include "k_sort.cuh"
#include "k_voxel.cuh"

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
      for (;;) {
        const uint32_t st = J.table[h].x;
        if (st != SLOT_EMPTY && (key64_pt(J.dst[st], J.g) >> s3) == ck) { J.table[h].y = (uint32_t)i + 1u; break; }
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

// ------------------------------------------------------------------------------------------------
// map update, step 2: merge-path partition + sorted copy of the new points
// ------------------------------------------------------------------------------------------------
// Tile t owns merged positions [t * MERGE_TILE, (t + 1) * MERGE_TILE); part[t] = old-map elements before that diagonal
// (ties: old elements first — PCL sums a voxel in cloud order, and the old map precedes the appended points, EM:313-323).
// One warp per diagonal: 32-ary search, so ~3 rounds of two dependent loads instead of ~15.
__global__ void __launch_bounds__(256) k_merge_partition(const MergeJob* __restrict__ jobs) {
  const MergeJob& J = jobs[blockIdx.y];
  MergeVars& V = *J.mv;
  const int n_old = *J.n_map, n_new = V.n_live;
  const int n_tot = n_old + n_new;
  const int n_tiles = (n_tot + MERGE_TILE - 1) / MERGE_TILE;
  const int res = sort_passes(V.bits, J.sort.npass) & 1;
  const uint32_t* __restrict__ val = J.sort.val[res];
  // sorted copy of the live new points and their keys (coalesced loads in the merge)
  for (int j = blockIdx.x * 256 + threadIdx.x; j < n_new; j += gridDim.x * 256) {
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
        const unsigned long long kn = key64_pt(J.newpts[val[d - a - 1]], J.g);
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
    if (n_tiles > J.max_tiles) atomicOr(J.status, ST_MAP_CAPACITY);
  }
}

// ------------------------------------------------------------------------------------------------
// map update, step 3: the merge (counting pass, emitting pass)
// ------------------------------------------------------------------------------------------------
struct MergeShared {
  float4 pts[MERGE_TILE];                 // old part [0, a), new part [a, a + b); later the tile's output points
  unsigned long long keys[MERGE_TILE];    // their keys; later the output points' cell keys
  unsigned short order[MERGE_TILE];       // merged position -> staged index; later the list of cell heads
  unsigned char live[MERGE_TILE];
  uint32_t scan[MERGE_THREADS / 32];
  int carry_start; unsigned long long carry_cell; int have_prev;
  int base;
  float bb[MERGE_THREADS / 32][7];
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
  for (int w = 0; w < MERGE_THREADS / 32; ++w) { const uint32_t c = buf[w]; if (w < warp) woff += c; tot += c; }
  *total = tot;
  return woff + inc - v;
}

template <bool EMIT>
__global__ void __launch_bounds__(MERGE_THREADS, 2) k_merge(const MergeJob* __restrict__ jobs) {
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
  const int d0 = t * MERGE_TILE, d1 = min(d0 + MERGE_TILE, n_tot);
  const int b0 = d0 - a0, b1 = d1 - a1;
  const int na = a1 - a0, nb = b1 - b0, nt = na + nb;
  float lo[3], hi[3];
  merge_crop(J, lo, hi);

  if (!EMIT) {  // the counting pass also clears the cell table of the map being written
    const int H = V.hmask + 1;
    const int per = (H + n_tiles - 1) / n_tiles;
    const int hb = min(H, t * per), he = min(H, hb + per);
    for (int i = hb + tid; i < he; i += MERGE_THREADS) J.table[i] = make_uint2(SLOT_EMPTY, 0u);
  }

  // ---- stage the tile: old part, new part; bounding box of the live points (PCL guard) ----
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  int nlive = 0;
  for (int i = tid; i < nt; i += MERGE_THREADS) {
    float4 p;
    unsigned long long k;
    bool lv;
    if (i < na) {
      p = J.old_pts[a0 + i];
      k = key64_pt(p, g);
      lv = !outside(p, lo, hi);
    } else {
      p = J.nsorted[b0 + i - na];
      k = J.nkey[b0 + i - na];
      lv = true;
    }
    S.pts[i] = p; S.keys[i] = k; S.live[i] = lv ? 1 : 0;
    if (!EMIT && lv) {
      ++nlive;
      mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
      mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
      mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
    }
  }
  // key of the merged element in front of the tile (run continuation) — ~0 when there is none
  unsigned long long prev_key = ~0ull;
  if (a0 > 0) prev_key = key64_pt(J.old_pts[a0 - 1], g);
  if (b0 > 0) { const unsigned long long kn = J.nkey[b0 - 1]; prev_key = (prev_key == ~0ull || kn > prev_key) ? kn : prev_key; }
  __syncthreads();

  // ---- merged order: old i -> i + #(new < key), new j -> j + #(old <= key) ----
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
  __syncthreads();

  // ---- every run head sums its voxel: PCL's `centroid += point` in cloud order, fp32, then `/= count` ----
  constexpr int PER = MERGE_TILE / MERGE_THREADS;
  float4 outp[PER];
  unsigned long long outk[PER];
  bool has[PER];
#pragma unroll
  for (int it = 0; it < PER; ++it) {
    const int r = it * MERGE_THREADS + tid;
    has[it] = false;
    outk[it] = 0;
    outp[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= nt) continue;
    const unsigned long long k = S.keys[S.order[r]];
    const unsigned long long pk = r > 0 ? S.keys[S.order[r - 1]] : prev_key;
    if (k == pk) continue;  // continues a run: its head (in this or an earlier tile) sums it
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
    has[it] = true; outp[it] = c; outk[it] = k >> s3;
  }
  __syncthreads();  // all reads of the staged tile are done: its storage now takes the compacted outputs

  // ---- compact the outputs in merged order ----
  uint32_t run = 0;
#pragma unroll
  for (int it = 0; it < PER; ++it) {
    uint32_t tot;
    const uint32_t ex = merge_block_scan(has[it] ? 1u : 0u, S.scan, &tot);
    if (has[it]) { S.pts[run + ex] = outp[it]; S.keys[run + ex] = outk[it]; }
    run += tot;
  }
  const int count = (int)run;
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
      for (int w = 0; w < MERGE_THREADS / 32; ++w) {
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

  // ---- emitting pass: where do this tile's outputs go, and which cell run is open in front of it? ----
  {
    uint32_t part = 0;
    for (int i = tid; i < t; i += MERGE_THREADS) part += (uint32_t)J.agg[i].count;
    uint32_t tot;
    merge_block_scan(part, S.scan, &tot);
    if (tid == 0) {
      S.base = (int)tot;
      // walk back to the start of the cell run that holds the last output in front of this tile
      int have = 0, start = 0;
      unsigned long long cell = 0;
      int b = (int)tot;  // outputs before tile tt + 1
      for (int tt = t - 1; tt >= 0; --tt) {
        const TileAgg A = J.agg[tt];
        if (A.count == 0) continue;
        b -= A.count;  // outputs before tile tt
        if (!have) { have = 1; cell = A.last_cell; }
        if (A.last_run_start > 0 || A.first_cell != cell) { start = b + A.last_run_start; break; }
        // the whole tile is one run of `cell` (or its tail starts at 0): it may have begun earlier
        start = b;
        bool cont = false;
        for (int t2 = tt - 1; t2 >= 0; --t2) {
          const TileAgg B = J.agg[t2];
          if (B.count == 0) continue;
          cont = B.last_cell == cell;
          break;
        }
        if (!cont) break;
      }
      S.have_prev = have; S.carry_start = start; S.carry_cell = cell;
    }
    __syncthreads();
  }
  const int base = S.base;
  const uint32_t hmask = (uint32_t)V.hmask;
  for (int o = tid; o < count; o += MERGE_THREADS)
    if (base + o < J.cap_out) J.out_pts[base + o] = S.pts[o];
  // cell heads of this tile, compacted in order; head c closes the cell in front of it
  uint32_t nheads = 0;
  {
    uint32_t runh = 0;
    for (int ob = 0; ob < count; ob += MERGE_THREADS) {
      const int o = ob + tid;
      bool head = false;
      if (o < count) head = o > 0 ? (S.keys[o] != S.keys[o - 1]) : (!S.have_prev || S.keys[0] != S.carry_cell);
      uint32_t tot;
      const uint32_t ex = merge_block_scan(head ? 1u : 0u, S.scan, &tot);
      if (head) S.order[runh + ex] = (unsigned short)o;
      runh += tot;
    }
    nheads = runh;
    __syncthreads();
  }
  for (int c = tid; c < (int)nheads; c += MERGE_THREADS) {
    const int o = S.order[c];
    if (c > 0) {
      const int po = S.order[c - 1];
      table_insert(J.table, hmask, S.keys[po], (uint32_t)(base + po), (uint32_t)(base + o));
    } else if (S.have_prev) {
      table_insert(J.table, hmask, S.carry_cell, (uint32_t)S.carry_start, (uint32_t)(base + o));
    }
  }
  if (t == n_tiles - 1 && tid == 0) {  // the last tile closes the last cell and publishes the result
    const int total = base + count;
    if (nheads > 0) { const int po = S.order[nheads - 1]; table_insert(J.table, hmask, S.keys[po], (uint32_t)(base + po), (uint32_t)total); }
    else if (S.have_prev) table_insert(J.table, hmask, S.carry_cell, (uint32_t)S.carry_start, (uint32_t)total);
    if (total > J.cap_out) atomicOr(J.status, ST_MAP_CAPACITY);
    *J.n_map = min(total, J.cap_out);
    J.meta[0] = (int)hmask; J.meta[1] = 0;
    // PCL's "leaf size is too small" guard (voxel_grid.hpp): the reference would hand the cloud through unfiltered
    if (V.n_live_all > 0) {
      const float inv = g.inv_leaf;
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

void launch_cell_update(const Launch& L, const MergeJob* jobs_dev, const SortJob* sort_jobs_dev, int njobs, int max_tiles) {
  k_merge_reset<<<(njobs + 127) / 128, 128, 0, L.st>>>(jobs_dev, njobs);
  L.tick(K_NEW_XFORM);
  dim3 g(74, njobs);
  k_new_xform<<<g, 256, 0, L.st>>>(jobs_dev);
  L.tick(K_NEW_XFORM);
  KeyGenNew gen;
  gen.jobs = jobs_dev;
  dim3 gs(SORT_G, njobs);
  k_sort_keyhist<KeyGenNew><<<gs, SORT_THREADS, 0, L.st>>>(sort_jobs_dev, gen);
  L.tick(K_NEW_KEYHIST);
  for (int pass = 0; pass < 4; ++pass) launch_sort_scatter(L, sort_jobs_dev, njobs, pass);
  k_merge_partition<<<g, 256, 0, L.st>>>(jobs_dev);
  L.tick(K_MERGE_PART);
  dim3 gm(max_tiles, njobs);
  k_merge<false><<<gm, MERGE_THREADS, sizeof(MergeShared), L.st>>>(jobs_dev);
  L.tick(K_MERGE_COUNT);
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
};

// Is map point a before map point b in the reference's map order?  Only consulted on exact fp32 distance ties (tie class T2).
__device__ __noinline__ bool pcl_before(const CellMapView& M, int a, int b) {
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
__device__ __forceinline__ bool closer_cell(const CellMapView& M, float d, int id, float bd, int bid) {
  if (d < bd) return true;
  if (d > bd) return false;
  return pcl_before(M, id, bid);
}

// Called by a whole warp with one query; returns (replicated in every lane) the five nearest map points inside the gate in
// ascending (d^2, PCL index) order; FLT_MAX / INT_MAX where fewer than five lie inside the gate.
__device__ __forceinline__ void warp_knn5(const CellMapView& M, float gate_f, float qx, float qy, float qz, float (&bd)[5], int (&bi)[5]) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int s = M.g.shift;
  const float inv = M.g.inv_leaf;
#pragma unroll
  for (int k = 0; k < 5; ++k) { bd[k] = FLT_MAX; bi[k] = INT_MAX; }
  int vx, vy, vz;
  voxel_of(make_float4(qx, qy, qz, 0.f), inv, vx, vy, vz);
  const int qcx = (int)(vbias(vx) >> s), qcy = (int)(vbias(vy) >> s), qcz = (int)(vbias(vz) >> s);
  const int ncmax = (1 << (21 - s)) - 1;
  // distance from the query to the nearest face of its own cell, in cells (for the early exit between shells)
  float fmin_cells = 0.f;
  {
    const float k = (float)(1 << s);
    const float ux = fmul(qx, inv) - (float)(((qcx << s) - VOX_BIAS)), uy = fmul(qy, inv) - (float)(((qcy << s) - VOX_BIAS)),
                uz = fmul(qz, inv) - (float)(((qcz << s) - VOX_BIAS));
    const float f = fminf(fminf(fminf(ux, k - ux), fminf(uy, k - uy)), fminf(uz, k - uz));
    fmin_cells = fmaxf(0.f, f / k - 1e-4f);
  }
  const double cell = (double)M.g.leaf * (double)(1 << s) * (1.0 - 1e-5);  // a lower bound of the cell edge (inverse_leaf is rounded)
  const int shells = M.g.shells;
  for (int r = 1; r <= shells; ++r) {
    const int side = 2 * r + 1, ncell = side * side * side;
    for (int cb = 0; cb < ncell; cb += 32) {
      const int t = cb + lane;
      uint32_t s0 = 0, e0 = 0;
      if (t < ncell) {
        const int dx = t % side - r, dy = (t / side) % side - r, dz = t / (side * side) - r;
        const int cx = qcx + dx, cy = qcy + dy, cz = qcz + dz;
        const bool shell = r == 1 || max(max(abs(dx), abs(dy)), abs(dz)) == r;  // the interior was visited in the previous round
        if (shell && cx >= 0 && cy >= 0 && cz >= 0 && cx <= ncmax && cy <= ncmax && cz <= ncmax) {
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
      uint32_t inc = cnt;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) { const uint32_t u = __shfl_up_sync(FULL, inc, off); if (lane >= off) inc += u; }
      const uint32_t total = __shfl_sync(FULL, inc, 31);
      for (uint32_t c0 = 0; c0 < total; c0 += 32) {
        const uint32_t c = c0 + lane;
        // owner = first lane whose inclusive count exceeds c
        int l = 0, h = 31;
#pragma unroll
        for (int it = 0; it < 5; ++it) {
          const int m = (l + h) >> 1;
          const uint32_t v = __shfl_sync(FULL, inc, m);
          if (v > c) h = m; else l = m + 1;
        }
        const uint32_t so = __shfl_sync(FULL, s0, l), io = __shfl_sync(FULL, inc, l), co = __shfl_sync(FULL, cnt, l);
        const int idx = (int)(so + (c - (io - co)));
        bool cand = c < total;
        float cd = FLT_MAX;
        if (cand) {
          const float4 p = __ldg(M.pts + idx);
          const float ddx = fsub(qx, p.x), ddy = fsub(qy, p.y), ddz = fsub(qz, p.z);
          cd = fadd(fadd(fmul(ddx, ddx), fmul(ddy, ddy)), fmul(ddz, ddz));  // FLANN L2_Simple<float>
          cand = cd < gate_f;  // neighbours at or beyond the gate can never be used (EM:129 / :189)
        }
        for (;;) {  // move the best remaining candidates of this batch into the list, nearest first (at most five rounds)
          const bool better = cand && closer_cell(M, cd, idx, bd[4], bi[4]);
          if (!__any_sync(FULL, better)) break;
          const uint32_t key = better ? __float_as_uint(cd) : 0xFFFFFFFFu;
          const uint32_t mnk = __reduce_min_sync(FULL, key);
          unsigned who = __ballot_sync(FULL, better && key == mnk);
          int w = __ffs(who) - 1;
          int wi = __shfl_sync(FULL, idx, w);
          who &= who - 1;
          while (who) {  // equal distances inside one batch: the reference's map order decides
            const int w2 = __ffs(who) - 1;
            const int i2 = __shfl_sync(FULL, idx, w2);
            if (pcl_before(M, i2, wi)) { w = w2; wi = i2; }
            who &= who - 1;
          }
          const float wd = __uint_as_float(mnk);
          // insert (wd, wi): it is closer than the current fifth
          bd[4] = wd; bi[4] = wi;
#pragma unroll
          for (int k = 4; k > 0; --k) {
            const bool sw = closer_cell(M, bd[k], bi[k], bd[k - 1], bi[k - 1]);
            const float dk = sw ? bd[k - 1] : bd[k], dk1 = sw ? bd[k] : bd[k - 1];
            const int ik = sw ? bi[k - 1] : bi[k], ik1 = sw ? bi[k] : bi[k - 1];
            bd[k] = dk; bd[k - 1] = dk1; bi[k] = ik; bi[k - 1] = ik1;
          }
          if (lane == w) cand = false;
        }
      }
    }
    if (r < shells) {  // after shell r every unseen point is farther than (r + f) cells
      const double reach = ((double)r + (double)fmin_cells) * cell;
      if ((double)bd[4] < reach * reach * (1.0 - 1e-5)) break;
    }
  }
}

constexpr int KC_THREADS = 256;

__global__ void __launch_bounds__(KC_THREADS) k_knn_cell_assoc(LaneDev* lanes, int lane0, int cur, ConfigDev cfg, const double* pose_override) {
  const int ln = lane0 + blockIdx.y;
  const LaneDev& L = lanes[ln];
  LaneVars& V = *L.v;
  const int me = V.n_map[0], ms = V.n_map[1];
  if (!(me > 10 && ms > 50)) return;  // EM:254
  const int ne = V.n_ds[0], ns = V.n_ds[1];
  double x[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) x[i] = pose_override ? pose_override[i] : V.x[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) V.opt_ran = 1;
  const int lane = threadIdx.x & 31;
  const int wpb = KC_THREADS / 32;
  const int nq = ne + ns;
  CellMapView M[2];
#pragma unroll
  for (int w = 0; w < 2; ++w) {
    M[w].pts = L.map[w][cur]; M[w].table = L.ctab[w]; M[w].hmask = (uint32_t)L.cmeta[w][0];
    M[w].orig = L.cmeta[w][1] ? L.corig[w] : nullptr; M[w].g = cfg.cg[w];
  }
  for (int q = blockIdx.x * wpb + (threadIdx.x >> 5); q < nq; q += gridDim.x * wpb) {
    const int w = q >= ne ? 1 : 0;
    const int k = w ? q - ne : q;
    const float4 pw = associate(x, L.ds[w][k]);  // EM:355-363
    float rd[5];
    int ri[5];
    warp_knn5(w ? M[1] : M[0], cfg.knn_gate_f, pw.x, pw.y, pw.z, rd, ri);
#pragma unroll
    for (int j = 0; j < 5; ++j)
      if (lane == j) {
        L.nn_idx[w][k * 5 + j] = ri[j] == INT_MAX ? -1 : ri[j];
        L.nn_d2[w][k * 5 + j] = rd[j];
      }
  }
}

void launch_knn_cell_fit(const Launch& L, LaneDev* lanes, int lane0, int nlanes, int cur, const ConfigDev& cfg, const double* pose_override) {
  dim3 g(KNN_G * 2, nlanes);
  k_knn_cell_assoc<<<g, KC_THREADS, 0, L.st>>>(lanes, lane0, cur, cfg, pose_override);
  L.tick(K_KNN_CELL);
  launch_fit(L, lanes, lane0, nlanes, cur, cfg);
}

// nearestKSearch alone against an explicit map (vilf_knn5); indices are reported in the caller's map order.
__global__ void __launch_bounds__(KC_THREADS) k_knn_cell_only(const float4* __restrict__ pts, const int* n_dev, const uint2* __restrict__ table, const int* meta,
                                                               const uint32_t* __restrict__ orig, CellGeom g, const float4* __restrict__ q, const int* nq_dev,
                                                               int* idx, float* d2, float gate_f) {
  const int nq = *nq_dev;
  const int lane = threadIdx.x & 31;
  const int wpb = KC_THREADS / 32;
  CellMapView M;
  M.pts = pts; M.table = table; M.hmask = (uint32_t)meta[0]; M.orig = meta[1] ? orig : nullptr; M.g = g;
  const bool empty = *n_dev == 0;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < nq; i += gridDim.x * wpb) {
    const float4 p = q[i];
    float rd[5];
    int ri[5];
    if (empty) {
#pragma unroll
      for (int j = 0; j < 5; ++j) { rd[j] = FLT_MAX; ri[j] = INT_MAX; }
    } else {
      warp_knn5(M, gate_f, p.x, p.y, p.z, rd, ri);
    }
#pragma unroll
    for (int j = 0; j < 5; ++j)
      if (lane == j) {
        idx[i * 5 + j] = ri[j] == INT_MAX ? -1 : (M.orig ? (int)M.orig[ri[j]] : ri[j]);
        d2[i * 5 + j] = rd[j];
      }
  }
}
void launch_knn_cell_only(const Launch& L, const float4* pts, const int* n_dev, const uint2* table, const int* meta, const uint32_t* orig, CellGeom g,
                          const float4* q, const int* nq_dev, int* idx, float* d2, const ConfigDev& cfg) {
  k_knn_cell_only<<<KNN_G * 8, KC_THREADS, 0, L.st>>>(pts, n_dev, table, meta, orig, g, q, nq_dev, idx, d2, cfg.knn_gate_f);
  L.tick(K_KNN_CELL);
}

}  // namespace vilf
