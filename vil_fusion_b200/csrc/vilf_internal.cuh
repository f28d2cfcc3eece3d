// Internal declarations shared by the kernels of libvilf_cuda.so (sm_100a only).
// Data layout in HBM, per sequence ("lane"): see DESIGN.md §3.  Every kernel is launched with a fixed
// grid (x = work, y = lane or job) and reads its element counts from device memory, so that a frame is
// a static launch sequence (CUDA-graph friendly) with no host round trip between the stages.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#include <limits.h>
#include <math.h>

struct vilf_handle;

namespace vilf {

constexpr int MAX_RINGS = 128;       // ring ids are 8-bit sort keys; 255 = dropped point
constexpr int SECTORS = 6;           // FE:206
constexpr int EDGES_PER_SECTOR = 20; // FE:131
constexpr int MAX_SECTOR = 2048;     // elements per (ring, sector) the selection kernel stages in shared memory
constexpr int SORT_G = 128;          // max CTAs per radix-sort job
constexpr int SORT_THREADS = 256;
constexpr int SORT_TILE = 1024;      // 256 threads x 4 keys
constexpr int SORT_MIN_CHUNK = 2048;
constexpr int SORT_RADIX_BITS = 9;   // digit width limit; tables are [4 passes][SORT_G][512]
constexpr int SORT_RADIX = 1 << SORT_RADIX_BITS;
constexpr int VOX_G = 148;           // CTAs per voxel job for the bbox / head / centroid kernels
constexpr int GRID_G = 148;          // CTAs per hash-grid job
constexpr int KNN_G = 148;           // the kNN kernels launch 4 x KNN_G CTAs per lane (128 threads = 16 queries of 8 lanes each, grid-stride)
constexpr int FIT_G = 148;           // CTAs (128 threads) of the fit kernel per lane
#ifndef VILF_LM_THREADS
#define VILF_LM_THREADS 256
#endif
#ifndef VILF_LM_CLUSTER
#define VILF_LM_CLUSTER 8
#endif
constexpr int LM_THREADS = VILF_LM_THREADS;  // per CTA of the solve cluster
constexpr int LM_CLUSTER = VILF_LM_CLUSTER;  // CTAs (SMs) per sequence in the solve kernel (default; ConfigDev::lm_cluster selects 8 or 16 at run time)
constexpr int LM_CLUSTER_MAX = 16;
constexpr int MAX_TRACE_ROWS = 8;
constexpr int MAX_OUTER = 4;

// status bits (LaneVars::status)
constexpr int ST_SECTOR_TOO_LONG = 1;
constexpr int ST_MAP_CAPACITY = 2;
constexpr int ST_SCAN_CAPACITY = 4;
constexpr int ST_KEY_RANGE = 8;     // cell-ordered map: the cloud spans more than 2^32 (cell, voxel) keys
constexpr int ST_PCL_GUARD = 16;    // cell-ordered map: PCL's int32 guard would skip the voxel filter (legacy path handles it)
constexpr int ST_ORPHANS = 32;      // more than ORPHAN_CAP centroids crossed a voxel face in one update

// geometry of a cell-ordered map (k_cellmap.cu)
struct CellGeom {
  float inv_leaf;  // PCL's inverse_leaf_size_ = 1.0f / leaf
  float leaf;
  int shift;       // log2(voxels per cell edge)
  int shells;      // cells of Chebyshev distance <= shells around the query cell cover the gate radius
};

struct ConfigDev {
  int n_scan, n_rings, rings_total;
  double lidar_min, lidar_max, edge_threshold, knn_gate, huber, crop_half;
  float edge_leaf, surf_leaf;
  float inv_cell;  // 1 / hash-grid cell edge (a power of two >= sqrt(knn_gate))
  float knn_gate_f;  // smallest float >= knn_gate: for a float d, d < knn_gate_f <=> (double)d < knn_gate
  int max_sector;  // elements per (ring, sector) the selection kernel stages in shared memory (<= MAX_SECTOR)
  int sector_np;   // power of two >= max_sector: size of the selection kernel's sort network
  int outer_iters, lm_max_iters;
  int cap_scan, cap_map;
  int flags_no_cluster;  // VILF_FLAG_NO_CLUSTER: grid-wide multi-launch kernels everywhere
  CellGeom cg[2];        // cell-ordered edge / surf map geometry
  // ring-field / range-image extractor (k_rangeimage.cu), stage 1 when range_image != 0
  int lm_cluster;  // CTAs per solve cluster: 8, or 16 when the leaves are fine enough for tens of thousands of factors per solve
  int range_image, horizon, ri_down;
  double ri_edge_thr, ri_surf_thr;
};

// One iteration row of the trust-region trace (same columns as the oracle's LmIter).
struct LmRow {
  double iteration, step_valid, step_successful, cost, candidate_cost, model_cost_change, relative_decrease, radius, step_norm;
  double x[7];
};
struct SolveTraceDev {
  int n_edge, n_surf, termination, n_rows;
  double H0[21], g0[6], cost0, final_cost;
  LmRow rows[MAX_TRACE_ROWS];
};

// Per-lane scalars living in device memory.
struct LaneVars {
  int n_scan[2];    // points in scan buffer 0 / 1 (written by the host copy of that buffer)
  int n_edge, n_surf;
  int n_ds[2];      // voxel-filtered scan features (edge, surf)
  int n_map[2];     // local map sizes
  int n_cat[2];     // map + appended scan features = input size of the map maintenance voxel job
  int status;
  int opt_ran;
  int frames;
  int pad_;
  double x[7];      // parameter_opti (EM:383)
  double odom[12];  // globalOdom: R row-major, t (EM:387)
  double odom_last[12];
};

struct SortJob {
  const int* n;           // element count
  const int* bits;        // significant key bits (device) or null -> fixed_bits
  int fixed_bits;
  int npass;              // pass limit: 1 (ring ids) or 4; P = min(npass, ceil(bits / 9)) passes of ceil(bits / P) bits
  uint32_t* key[2];       // result in key[P & 1] / val[P & 1]
  uint32_t* val[2];
  uint2* pair[2];         // the same storage seen as (key, value) pairs: the cluster path moves a pair with one 8-byte access
  uint32_t* hist;         // [4][SORT_G][SORT_RADIX]
  uint32_t* digit_start;  // optional [257]: exclusive digit offsets of pass 0 (+ total)
};

struct VoxVars {
  int bbox[6];    // ordered-int min x,y,z / max x,y,z
  int n_valid;    // points inside the crop box (== n_in without crop)
  int min_b[3], div_b[3];
  int bits;       // significant bits of (voxel index | sentinel)
  int guard;      // PCL "leaf size too small" guard fired: output = input
  int total;      // dx*dy*dz (sentinel key of cropped-out points)
  int pad_;
  unsigned long long t[8];  // cluster path: %globaltimer at the phase boundaries (ns), written by the leader thread
};

struct VoxJob {
  const float4* in;
  const int* n_in;
  float leaf;
  int crop;                 // 0 none; 1 box = crop_center +- crop_half (EM:327-336); 2 explicit fp32 bounds
  float crop_lo[3], crop_hi[3];
  int passthrough;          // 1: no voxel filter, output = (cropped) input in order
  int emit_all;             // 1: sort by voxel index but emit EVERY point (cell-ordered map -> PCL order, grid-wide path only)
  const double* crop_center;  // 3 doubles on the device (pose translation)
  double crop_half;
  float4* out;
  int* n_out;
  int cap_out;
  int* status;              // lane status word (capacity overflow)
  VoxVars* vv;
  int* head_cnt;            // [VOX_G * 8]: heads per warp range of the sorted points (grid-wide path)
  SortJob sort;
  // cluster path only (k_cluster.cu).  Map jobs: createSubMap's append (EM:308-324) runs inside the voxel kernel:
  // in[app_n_map .. ) <- associate(app_pose, app_src[0 .. *app_n)), n_in <- min(app_cap, *app_n_map + *app_n).
  const float4* app_src;
  const int* app_n;
  const int* app_n_map;
  const double* app_pose;
  int app_cap;
  const struct GridJob* grid;  // spatial hash to build over `out` after the filter (map jobs) or null
};

struct GridJob {
  const float4* pts;
  const int* n;
  uint32_t* start;    // [hcap + 1] bucket counts, then exclusive offsets (start[H] = n)
  uint32_t* rank;     // [cap] slot of each point inside its bucket
  float4* sorted;     // [cap] x, y, z, original index (bits)
  int* hvar;          // buckets in use (power of two)
  uint32_t* partial;  // [GRID_G]
  int hcap;
  float inv_cell;     // 1 / cell edge of THIS grid: cell = cmax / 2^k, cmax = the power of two >= sqrt(knn_gate)
  int rings;          // 2^k: cells of Chebyshev distance <= rings around the query cell cover the gate radius
};

// ---- cell-ordered local map (k_cellmap.cu) ----
// A local map is stored sorted by (search cell, voxel inside the cell): cells are cubes of 2^shift voxels of the map's own
// voxel filter, so ONE order serves both pcl::VoxelGrid (a voxel is a run of equal keys; the per-frame update is a merge of
// the sorted map with the few thousand sorted new points) and the 5-NN search (a cell is a contiguous range; an open-addressing
// table maps cell -> [start, end)).
constexpr int MERGE_TILE = 1024;     // merged elements per CTA tile of the map update
constexpr int MERGE_THREADS = 256;
constexpr int ORPHAN_CAP = 256;      // centroids that rounded across a voxel face, re-inserted with the next frame's points
constexpr int VOX_BIAS = 1 << 20;    // voxel coordinates are biased into 21 unsigned bits per axis
struct MergeVars {
  int n_in;        // orphans + voxel-filtered scan features = size of the new-point sort job
  int n_orph_in;   // orphans at the head of newpts (the registered cloud starts behind them)
  int n_live;      // new points inside the crop box
  int n_old, n_tot, n_tiles;
  int vb[6];       // voxel-coordinate bounding box of the live new points
  int bits;        // significant bits of their relative sort keys
  int fb[6];       // ordered-int fp32 bounding box of every live point (PCL's "leaf size too small" guard)
  int n_live_all;
  int n_orph;      // orphans waiting for the next update
  int hmask;       // cell table size - 1 of the map being written
  int old_unique;  // the old map holds at most one point per voxel (it came out of an update; a first-frame / imported map does not)
  int pad_[2];
};
struct TileAgg {   // per merge tile, written by the counting pass
  int count;             // voxels (map points) the tile emits
  int last_run_start;    // tile-relative index of the first point of the last cell run
  unsigned long long first_cell, last_cell;
};
struct TileOut {   // per merge tile, written by the scan between the two passes
  int base;              // outputs in front of the tile
  int have_prev;         // is there any output in front of it?
  int carry_start;       // index of the first point of the cell run that is open in front of the tile
  int pad_;
  unsigned long long carry_cell;
};
struct MergeJob {
  const float4* old_pts; float4* out_pts;
  int* n_map;                 // in: points of the old map, out: points of the new one
  const float4* src; const int* n_src; const double* pose;   // voxel-filtered scan features to append at `pose` (EM:308-324); pose null: src is in the map frame
  const double* crop_center; double crop_half;  // EM:327-336; crop_center null: no crop box
  float4* newpts; float4* nsorted; unsigned long long* nkey;
  SortJob sort;               // new points: (relative key, index) pairs
  MergeVars* mv;
  uint32_t* part;             // [max_tiles + 1] old elements before every tile boundary (merge path)
  TileAgg* agg;               // [max_tiles]
  TileOut* tout;              // [max_tiles]
  uint2* table; int hcap;     // cell -> [start, end) of the map being written
  int* meta;                  // [0] cell table mask, [1] has_orig (0 after an update)
  float4* orphans;
  CellGeom g;
  int cap_out, cap_new, max_tiles;
  int* status;
};
struct CellBuildJob {         // arbitrary cloud (PCL order = input order) -> cell-ordered map + original indices + cell table
  const float4* src; const int* n;
  float4* dst; uint32_t* orig;
  uint2* table; int hcap;
  int* meta;                  // [0] mask, [1] has_orig = 1, [2..7] voxel bbox, [8] bits
  SortJob sort;
  CellGeom g;
  int* status;
};

// Per-lane device pointers.
struct LaneDev {
  LaneVars* v;
  VoxVars* vv;                       // [4]: scan edge, scan surf, map edge, map surf voxel jobs
  SolveTraceDev* trace;              // [MAX_OUTER]
  float4* scan[2]; uint16_t* ring_in[2];  // double buffered: H2D of frame t+1 overlaps the kernels of frame t
  // stage 1
  SortJob ring_sort;   // keys = ring id, values = scan index
  int2* sec_cnt;       // [MAX_RINGS*6] (edges, surfs) chosen per sector
  float4* sec_edge; int* sec_edge_src;   // [MAX_RINGS*6*20]
  float4* sec_surf; int* sec_surf_src;   // [cap_scan] indexed from the sector's first element
  float4* feat[2]; int* feat_src[2];     // edge / surf features [cap_scan]
  // stage 2: voxel-filtered scan features
  float4* ds[2];
  // maps: [which][buffer] double buffered, capacity cap_map + cap_scan
  float4* map[2][2];
  // factors (one slot per voxel-filtered feature)
  uint8_t* fvalid[2];
  double* edge_pab;    // [cap][9]
  double* surf_pnd;    // [cap][7]
  int* nn_idx[2]; float* nn_d2[2];  // [cap][5] (test hooks)
  // cell-ordered maps: search table, PCL index of every point (valid while meta[1] != 0), meta
  uint2* ctab[2]; uint32_t* corig[2]; int* cmeta[2];
  // range-image extractor: image cell -> first input point, per-ring counters, then per image point (row-major): input index,
  // column, range, curvature, marks, labels
  int* ri_owner; int* ri_info; int* ri_src; int* ri_col; float* ri_range; float* ri_curv; uint8_t* ri_picked; uint8_t* ri_label;
};

// ---- ordered-int float mapping for atomicMin/atomicMax ----
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7FFFFFFF; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

// Block-reduce a bounding box + count and merge it into a voxel job's VoxVars with 7 atomics per CTA (getMinMax3D is
// order independent).  Every thread of the CTA must call it; sm = 7 * (blockDim.x / 32) floats of shared memory.
__device__ __forceinline__ void bbox_commit(VoxVars* vv, float mn[3], float mx[3], int cnt, float* sm) {
  for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], off));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], off));
    }
    cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
  }
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();  // sm may still be in use by a previous call
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { sm[warp * 7 + a] = mn[a]; sm[warp * 7 + 3 + a] = mx[a]; }
    sm[warp * 7 + 6] = __int_as_float(cnt);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int total = 0;
    for (int w = 0; w < nw; ++w) {
#pragma unroll
      for (int a = 0; a < 3; ++a) { mn[a] = fminf(mn[a], sm[w * 7 + a]); mx[a] = fmaxf(mx[a], sm[w * 7 + 3 + a]); }
      total += __float_as_int(sm[w * 7 + 6]);
    }
    if (total > 0) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        atomicMin(&vv->bbox[a], f2ord(mn[a]));
        atomicMax(&vv->bbox[3 + a], f2ord(mx[a]));
      }
      atomicAdd(&vv->n_valid, total);
    }
  }
}

// fp32 arithmetic that must round exactly like the reference's scalar SSE code (no contraction).
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }

// ---- double-precision rigid-body helpers (Eigen 3.3.7 operation order) ----
struct D3 { double x, y, z; };
__device__ __forceinline__ D3 d3(double x, double y, double z) { D3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return d3(dadd(a.x, b.x), dadd(a.y, b.y), dadd(a.z, b.z)); }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return d3(dsub(a.x, b.x), dsub(a.y, b.y), dsub(a.z, b.z)); }
__device__ __forceinline__ D3 operator*(double s, D3 a) { return d3(dmul(s, a.x), dmul(s, a.y), dmul(s, a.z)); }
__device__ __forceinline__ double dot3(D3 a, D3 b) { return dadd(dadd(dmul(a.x, b.x), dmul(a.y, b.y)), dmul(a.z, b.z)); }
__device__ __forceinline__ D3 cross3(D3 a, D3 b) {
  return d3(dsub(dmul(a.y, b.z), dmul(a.z, b.y)), dsub(dmul(a.z, b.x), dmul(a.x, b.z)), dsub(dmul(a.x, b.y), dmul(a.y, b.x)));
}
__device__ __forceinline__ double norm3(D3 a) { return sqrt(dot3(a, a)); }
struct Q4 { double x, y, z, w; };
// Eigen QuaternionBase::_transformVector: v + w*uv + qv x uv, uv = 2 (qv x v)   (EM:358, LF:26, LF:83)
__device__ __forceinline__ D3 qrot(Q4 q, D3 v) {
  D3 qv = d3(q.x, q.y, q.z);
  D3 uv = cross3(qv, v);
  uv = uv + uv;
  return (v + q.w * uv) + cross3(qv, uv);
}
// pointAssociaToMap (EM:355-363): fp64 transform, fp32 store.
__device__ __forceinline__ float4 associate(const double* x, float4 p) {
  Q4 q; q.x = x[0]; q.y = x[1]; q.z = x[2]; q.w = x[3];
  D3 w = qrot(q, d3((double)p.x, (double)p.y, (double)p.z)) + d3(x[4], x[5], x[6]);
  return make_float4((float)w.x, (float)w.y, (float)w.z, p.w);
}

// ---- launch helpers implemented in the .cu files (all asynchronous on `st`) ----
enum KernelId {
  K_RESET = 0, K_RING_KEYHIST, K_SORT_HIST, K_SORT_SCATTER, K_SECTOR, K_COMPACT, K_VOX_BBOX, K_VOX_KEYHIST, K_VOX_HEADS, K_VOX_CENTROID,
  K_MAP_APPEND, K_MAP_INIT, K_GRID_ZERO, K_GRID_COUNT, K_GRID_SCAN_PARTIAL, K_GRID_SCAN_FINAL, K_GRID_SCATTER, K_KNN_FIT, K_KNN_ONLY,
  K_SOLVE, K_FIT, K_VOX_CLUSTER, K_GRID_CLUSTER, K_DEPTH_CLOUD, K_DEPTH_QUERY, K_RING_PARTITION,
  K_NEW_XFORM, K_NEW_KEYHIST, K_MERGE_PART, K_MERGE_COUNT, K_MERGE_EMIT, K_CELL_BUILD, K_KNN_CELL, K_SC, K_RANGE_IMAGE, K_RI_SELECT, K_COUNT
};
constexpr int PROF_PHASES = 5;      // 0 extract, 1 scan downsample, 2 association + solve, 3 map update, 4 grid build
constexpr int PROF_KSLOTS = 64;     // kernel ids per phase in a profile tag
constexpr int PROF_TAGS = PROF_PHASES * PROF_KSLOTS;
constexpr int PROF_MAX_EVENTS = 96;

// Optional per-kernel timing: one CUDA event after every launch, tagged (phase, kernel).
struct ProfSink {
  cudaEvent_t* ev;
  int* tag;
  int n, cap, phase;
};
extern int g_debug_sync;  // VILF_DEBUG_SYNC=1: synchronise after every kernel and name the one that faults (forces VILF_FLAG_NO_GRAPH)
void debug_sync_check(cudaStream_t st, int kid);
struct Launch {
  cudaStream_t st;
  int64_t* counter;  // kernels launched
  ProfSink* prof;    // null unless profiling
  void tick(int kid) const {
    ++*counter;
    if (g_debug_sync) debug_sync_check(st, kid);
    if (prof && prof->n < prof->cap) {
      cudaEventRecord(prof->ev[prof->n], st);
      prof->tag[prof->n++] = kid + PROF_KSLOTS * prof->phase;
    }
  }
};

// k_sort.cu
void launch_sort_scatter(const Launch& L, const SortJob* jobs_dev, int njobs, int pass);
// k_extract.cu
cudaError_t init_extract_kernels();  // per-device function attributes; call once per context
void launch_frame_reset(const Launch& L, LaneDev* lanes, int lane0, int nlanes, VoxVars* vv, int vv_per_lane, int predict);
void launch_extract(const Launch& L, LaneDev* lanes, const SortJob* ring_jobs, int lane0, int nlanes, int sel, const ConfigDev& cfg);
// k_rangeimage.cu
cudaError_t init_rangeimage_kernels();
void launch_extract_range_image(const Launch& L, LaneDev* lanes, int lane0, int nlanes, int sel, const ConfigDev& cfg);
// k_voxel.cu
void launch_voxel(const Launch& L, const VoxJob* jobs_dev, int njobs, const SortJob* sort_jobs_dev, bool bbox_done);
void launch_map_append(const Launch& L, LaneDev* lanes, int lane0, int nlanes, int cur, const ConfigDev& cfg);
void launch_map_init(const Launch& L, LaneDev* lanes, int lane0, int nlanes, int cur, const ConfigDev& cfg);
void launch_map_init_commit(const Launch& L, LaneDev* lanes, int lane0, int nlanes, const ConfigDev& cfg);
// k_cluster.cu
constexpr int CLUSTER_MAX_POINTS = 1 << 19;  // clouds up to this size take the one-cluster-per-cloud path
void launch_voxel_cluster(const Launch& L, const VoxJob* jobs_dev, int njobs, bool bbox_done, const ConfigDev& cfg);
void launch_grid_cluster(const Launch& L, const GridJob* jobs_dev, int njobs, const ConfigDev& cfg);
// k_knn.cu
void launch_grid_build(const Launch& L, const GridJob* jobs_dev, int njobs, const ConfigDev& cfg);
void launch_knn_fit(const Launch& L, LaneDev* lanes, const GridJob* grid_jobs, int lane0, int nlanes, int cur, const ConfigDev& cfg,
                    const double* pose_override);
void launch_knn_only(const Launch& L, const GridJob* job_dev, const float4* q, const int* nq_dev, int* idx, float* d2, const ConfigDev& cfg);
// k_cellmap.cu
cudaError_t init_cellmap_kernels();
void launch_cell_update(const Launch& L, const MergeJob* jobs_dev, const SortJob* sort_jobs_dev, int njobs, int max_tiles, bool cluster_new);
void launch_cell_build(const Launch& L, const CellBuildJob* jobs_dev, const SortJob* sort_jobs_dev, int njobs);
// seeded: the neighbour lists of the previous outer iteration of THIS frame are in nn_idx and bound the search
void launch_knn_cell_fit(const Launch& L, LaneDev* lanes, int lane0, int nlanes, int cur, const ConfigDev& cfg, const double* pose_override, bool seeded = false);
void launch_knn_cell_only(const Launch& L, const float4* pts, const int* n_dev, const uint2* table, const int* meta, const uint32_t* orig, CellGeom g,
                          const float4* q, const int* nq_dev, int* idx, float* d2, const ConfigDev& cfg);
void launch_cell_unpermute(const Launch& L, const float4* pts, const uint32_t* orig, const int* n_dev, float4* out, int cap);
void launch_fit(const Launch& L, LaneDev* lanes, int lane0, int nlanes, int cur, const ConfigDev& cfg);
// k_depth.cu
void launch_depth(const Launch& L, const float4* in, const int* n_dev, int from_scan, const double* T_dev, float4* sph, int* oidx, int* count,
                  const float* feats_dev, int m, float thr, float* depth_dev, int* nn_dev);
// k_scancontext.cu (place recognition next to the path: SCManager, Scancontext.h:42-299)
struct ScParams {
  double lidar_height, max_radius, search_ratio, dist_thres;
  int num_ring, num_sector, num_exclude_recent, num_candidates, tree_making_period;
};
cudaError_t init_sc_kernels();
// vilf_api.cu: the cloud getMapCloud(MapCloud) hands to the node (EM:371-375 = /GlobalMap) where it lies on the device, as two segments;
// `st` = the stream the odometry writes them on
int resident_scan_features(vilf_handle* h, const float4* p[2], const int* n[2], cudaStream_t* st, int* device);
void launch_sc_make(const Launch& L, const float4* p0, const int* n0_dev, const float4* p1, const int* n1_dev, const ScParams& P, int* bins, double* desc,
                    double* ringkey, double* sectorkey, float* invkey);
void launch_sc_distance(const Launch& L, const double* sc1, const double* sc2, const ScParams& P, double* out);
void launch_sc_detect(const Launch& L, const float* keys, int n_snapshot, const float* cur_key, const double* descs, int cur_index, const ScParams& P, float* dist,
                      double* res);
// k_solve.cu
cudaError_t init_solve_kernels();
void launch_solve(const Launch& L, LaneDev* lanes, int lane0, int nlanes, int outer, int finalize, const ConfigDev& cfg, int max_iters);

}  // namespace vilf
