// C ABI of libvilf_cuda.so (include/vilf.h): device context, per-sequence state, the per-frame launch
// sequence and the stage-level entry points.  No CPU fallback: every entry point needs a CUDA device.
#include "vilf_internal.cuh"
#include <nvtx3/nvToolsExt.h>
#include "../../include/vilf.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <new>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

using namespace vilf;

namespace vilf {
int g_debug_sync = 0;
void debug_sync_check(cudaStream_t st, int kid) {
  cudaError_t e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) fprintf(stderr, "[vilf] kernel id %d (%s) failed: %s\n", kid, vilf_profile_kernel_name(kid), cudaGetErrorString(e));
}
}  // namespace vilf

namespace {

constexpr int RING_SLOTS = 8;
constexpr int DEPTH_MAX_FEATURES = 8192;
constexpr int VV_PER_LANE = 4;  // scan edge, scan surf, map edge, map surf
constexpr int N_STAGE = 7;
constexpr int MAX_MARKS = PROF_MAX_EVENTS;

struct Slot {
  cudaEvent_t done = nullptr;
  cudaEvent_t stage[MAX_MARKS] = {};  // profiling: event k closes an interval attributed to stage_tag[k] (phase*64 + kernel)
  int stage_tag[MAX_MARKS] = {};
  ProfSink sink;
  LaneVars* vars_pin = nullptr;  // [nlanes]
  int64_t ticket = -1;
  int lane0 = 0, nl = 0;
  bool profiled = false, first = false;
};

struct Ctx {
  int device = 0;
  vilf_config ucfg;
  ConfigDev cfg;
  int nlanes = 0;
  cudaStream_t st = nullptr, copy_st = nullptr;
  int64_t launches = 0;
  std::vector<void*> allocs;
  std::vector<void*> pinned;
  std::vector<LaneDev> lanes_host;
  LaneDev* lanes_dev = nullptr;
  float4* scan_all[2] = {nullptr, nullptr}; uint16_t* ring_all[2] = {nullptr, nullptr};  // [buffer][lane][cap_scan]
  LaneVars* vars_dev = nullptr;
  SolveTraceDev* trace_dev = nullptr;
  VoxVars* vv_dev = nullptr;  // [nlanes * 4 + 1] (last = aux)
  SortJob* ring_jobs_dev[2] = {nullptr, nullptr};  // per scan buffer (n = &n_scan[sel])
  VoxJob* vox_scan_dev = nullptr; SortJob* vox_scan_sort_dev = nullptr;      // [nlanes*2]
  VoxJob* vox_map_dev[2] = {nullptr, nullptr}; SortJob* vox_map_sort_dev[2] = {nullptr, nullptr};  // [cur][nlanes*2]
  GridJob* grid_dev[2] = {nullptr, nullptr};                                  // [buf][nlanes*2]
  // aux (stage-level entry points)
  // steady-state frames replayed as CUDA graphs: one instantiated graph per (lane0, lanes, scan buffer, map buffer)
  struct FrameGraph { cudaGraphExec_t exec = nullptr; int launches = 0; };
  std::map<std::tuple<int, int, int, int>, FrameGraph> graphs;
  bool use_graphs = true;
  bool cluster_scan = false, cluster_map = false;  // one-cluster-per-cloud path (k_cluster.cu) for scan / map clouds
  // cell-ordered maps (k_cellmap.cu): merge update + cell table; off with VILF_FLAG_LEGACY_MAP
  bool cellmap = true;
  bool no_seed = false;  // VILF_KNN_NO_SEED=1 (A/B): the second outer iteration searches from scratch
  int max_tiles = 0;
  MergeJob* merge_dev[2] = {nullptr, nullptr}; SortJob* merge_sort_dev = nullptr;      // [cur][nlanes*2]
  CellBuildJob* build_dev[2] = {nullptr, nullptr}; SortJob* build_sort_dev = nullptr;  // [cur][nlanes*2]: first-frame map from the raw features
  std::vector<MergeJob> merge_host[2];
  std::vector<CellBuildJob> build_host[2];
  MergeVars* mv_dev = nullptr;  // [nlanes*2 + 2]
  // aux cell map (explicit maps: vilf_knn5, vilf_bench_stage, state import staging)
  float4* aux_cm_pts[2] = {nullptr, nullptr}; uint2* aux_ctab = nullptr; uint32_t* aux_corig = nullptr; int* aux_cmeta = nullptr;
  CellBuildJob* aux_build_dev = nullptr; SortJob* aux_build_sort_dev = nullptr; CellBuildJob aux_build_host;
  MergeJob* aux_merge_dev = nullptr; SortJob* aux_merge_sort_dev = nullptr; MergeJob aux_merge_host; int aux_max_tiles = 0;
  int aux_hcap = 0;
  int cap_aux = 0;
  float4* aux_in = nullptr; float4* aux_out = nullptr;
  int* aux_n = nullptr;      // [4] n_in, n_out, nq, spare
  double* aux_pose = nullptr;  // [8]
  int* aux_idx = nullptr; float* aux_d2 = nullptr;
  VoxJob* aux_vox_dev = nullptr; SortJob* aux_sort_dev = nullptr; GridJob* aux_grid_dev = nullptr;
  VoxJob aux_vox_host;
  GridJob aux_grid_host;
  float* depth_feat = nullptr; float* depth_out = nullptr; int* depth_nn = nullptr; double* depth_T = nullptr;  // [DEPTH_MAX_FEATURES] scratch
  std::vector<int> last_sel;   // scan buffer holding the most recent scan of the lane
  int* aux_head_cnt = nullptr;
  // per-lane host state
  std::vector<int> cur;        // map buffer holding the current local maps
  std::vector<int> have_map;   // localMapInited / import done
  std::vector<int> have_feat;  // features of a scan are resident
  std::vector<int> last_init;  // last map operation was the initialisation (getMapCloud semantics)
  std::vector<int64_t> frame_no;
  std::vector<vilf_handle*> handles;
  // async ring
  Slot slots[RING_SLOTS];
  int64_t next_ticket = 0;
  cudaEvent_t extract_done[2] = {nullptr, nullptr}, h2d_done[2] = {nullptr, nullptr};
  int64_t scan_sel = 0;
  // profiling
  bool profile = false;
  double kernel_ms[PROF_TAGS] = {};
  int64_t kernel_cnt[PROF_TAGS] = {};
  double frame_ms = 0;
  int64_t prof_frames = 0;
  char err[512] = {0};
};

}  // namespace

struct vilf_handle {
  Ctx* ctx;
  int lane;
};

namespace {

#define CK(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess) {                                                                              \
      snprintf(C->err, sizeof(C->err), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return VILF_ERR_CUDA;                                                                               \
    }                                                                                                     \
  } while (0)

template <class T>
cudaError_t dalloc(Ctx* C, T** p, size_t count) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 256);
  if (e != cudaSuccess) return e;
  C->allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return cudaMemset(q, 0, count * sizeof(T) + 256);
}

int fail(Ctx* C, int code, const char* msg) {
  snprintf(C->err, sizeof(C->err), "%s", msg);
  return code;
}

Launch mk(Ctx* C, ProfSink* sink = nullptr) { Launch L; L.st = C->st; L.counter = &C->launches; L.prof = sink; return L; }

// Spatial-hash build of `njobs` maps: one cluster per map for clouds of up to CLUSTER_MAX_POINTS, grid-wide otherwise.
void build_grids(Ctx* C, const Launch& L, const GridJob* jobs_dev, int njobs, bool small) {
  if (small) launch_grid_cluster(L, jobs_dev, njobs, C->cfg);
  else launch_grid_build(L, jobs_dev, njobs, C->cfg);
}

int pow2_ge(int v) { int p = 1; while (p < v) p <<= 1; return p; }

int alloc_sort(Ctx* C, SortJob& S, const int* n, const int* bits, int fixed_bits, int npass, int cap, bool digit_start) {
  S.n = n; S.bits = bits; S.fixed_bits = fixed_bits; S.npass = npass;
  for (int b = 0; b < 2; ++b) {  // one allocation per buffer: [cap] keys then [cap] values (grid-wide path) == [cap] pairs (cluster path)
    const size_t capr = ((size_t)cap + 63) / 64 * 64;
    CK(dalloc(C, &S.pair[b], capr));
    S.key[b] = reinterpret_cast<uint32_t*>(S.pair[b]);
    S.val[b] = S.key[b] + capr;
  }
  CK(dalloc(C, &S.hist, (size_t)4 * SORT_G * SORT_RADIX));
  S.digit_start = nullptr;
  if (digit_start) CK(dalloc(C, &S.digit_start, 264));
  return VILF_OK;
}

// Cell edge of a map's search grid: cmax = the power of two >= sqrt(knn_gate) (one shell of 27 cells covers the gate), halved
// while the cell still spans >= 2.5 leaf sizes, so a voxel-filtered map keeps a handful of points per cell however fine its
// leaf is (0.4 / 0.8 m leaves keep the 1 m cell; a 0.2 m leaf gets 0.5 m cells and two shells with early exit).
void grid_geometry(const Ctx* C, double leaf, GridJob& G) {
  float cell = 1.0f / C->cfg.inv_cell;
  int rings = 1;
  while (rings < 16 && (double)cell * 0.5 >= 2.5 * leaf) { cell *= 0.5f; rings *= 2; }
  G.inv_cell = 1.0f / cell;
  G.rings = rings;
}

// Search cells of a cell-ordered map: cubes of 2^shift voxels of the map's own voxel filter, walked shell by shell with early
// exit.  Cell edge = the smallest power-of-two multiple of the leaf that is >= max(2 leaves, 0.35 x gate radius):
//  - on a dense voxel-filtered surface the first shell (a box of three cells) then holds a few dozen to ~150 candidates and the fifth
//    neighbour (~1.3 leaves away) lies inside the distance that shell guarantees, so the walk stops after it;
//  - in the sparse far field of a lidar map (point spacing set by the beams, not by the leaf) cells much finer than the gate make
//    every query walk many shells of empty cells.
// Measured on the configs[2] sequence (one sequence, 3.9e4 queries per launch, warp-per-query search): leaf 0.1 m, ~1e6 map points:
// 229 us per search launch with 0.2 m cells, 95 us with 0.4 m, 180 us with 0.8 m; leaf 0.09 m, 1.13e6 points
// (gpurun_out/r2m_dense_shift*.json): 275 us with 0.18 m cells, 111 us with 0.36 m, 219 us with 0.72 m.
// Leaves at or above the gate radius use one voxel per cell.
CellGeom cell_geometry(const Ctx* C, double leaf) {
  CellGeom g;
  g.leaf = (float)leaf;
  g.inv_leaf = 1.0f / (float)leaf;  // inverse_leaf_size_ = Array4f::Ones() / leaf_size_.array()
  const double reach = std::sqrt(C->ucfg.knn_gate) * 1.001;
  int shift = 0;
  if ((double)g.leaf < reach) {
    const double target = std::max(2.0 * (double)g.leaf, 0.35 * reach);
    while ((double)(1 << shift) * (double)g.leaf < target * (1.0 - 1e-6) && shift < 12) ++shift;
  }
  if (const char* e = getenv("VILF_CELL_SHIFT")) shift = atoi(e);  // experiments only
  int shells = (int)std::ceil(reach / ((double)(1 << shift) * (double)g.leaf));
  while (shells > 6 && shift < 12) { ++shift; shells = (int)std::ceil(reach / ((double)(1 << shift) * (double)g.leaf)); }
  g.shift = shift; g.shells = shells < 1 ? 1 : shells;
  return g;
}

int alloc_merge(Ctx* C, MergeJob& M, int cap_new, int cap_map_total) {
  memset(&M, 0, sizeof(M));
  M.cap_new = cap_new;
  M.max_tiles = (cap_map_total + cap_new + MERGE_TILE - 1) / MERGE_TILE + 1;
  CK(dalloc(C, &M.newpts, (size_t)cap_new));
  CK(dalloc(C, &M.nsorted, (size_t)cap_new));
  CK(dalloc(C, &M.nkey, (size_t)cap_new));
  CK(dalloc(C, &M.part, (size_t)M.max_tiles + 2));
  CK(dalloc(C, &M.agg, (size_t)M.max_tiles + 1));
  CK(dalloc(C, &M.tout, (size_t)M.max_tiles + 1));
  CK(dalloc(C, &M.orphans, (size_t)ORPHAN_CAP));
  return VILF_OK;
}

int alloc_grid(Ctx* C, GridJob& G, const float4* pts, const int* n, int cap, double leaf) {
  G.pts = pts; G.n = n;
  grid_geometry(C, leaf, G);
  G.hcap = pow2_ge(2 * cap);
  if (G.hcap < 1024) G.hcap = 1024;
  CK(dalloc(C, &G.start, (size_t)G.hcap + 8));
  CK(dalloc(C, &G.rank, (size_t)cap));
  CK(dalloc(C, &G.sorted, (size_t)cap));
  CK(dalloc(C, &G.hvar, 4));
  CK(dalloc(C, &G.partial, (size_t)GRID_G));
  return VILF_OK;
}

int build_ctx(Ctx* C) {
  const vilf_config& u = C->ucfg;
  ConfigDev& c = C->cfg;
  c.n_scan = u.n_scan; c.n_rings = u.n_rings;
  if (u.n_scan == 0) c.rings_total = u.n_rings;
  else if (u.n_scan == 16 || u.n_scan == 32 || u.n_scan == 64) c.rings_total = u.n_scan;
  else c.rings_total = 1;
  c.lidar_min = u.lidar_min; c.lidar_max = u.lidar_max; c.edge_threshold = u.edge_threshold;
  c.knn_gate = u.knn_gate; c.huber = u.huber; c.crop_half = u.crop_half;
  c.edge_leaf = (float)u.edge_leaf; c.surf_leaf = (float)u.surf_leaf;  // setLeafSize takes floats (EM:85-86)
  float cell = 1.0f / 1024.0f;
  while ((double)cell * (double)cell < u.knn_gate) cell *= 2.0f;  // power of two >= sqrt(gate)
  c.inv_cell = 1.0f / cell;
  c.knn_gate_f = (float)u.knn_gate;
  if ((double)c.knn_gate_f < u.knn_gate) c.knn_gate_f = nextafterf(c.knn_gate_f, INFINITY);
  c.outer_iters = u.outer_iters; c.lm_max_iters = u.lm_max_iters;
  {
    int ring_pts = u.max_ring_points > 0 ? u.max_ring_points : 6 * MAX_SECTOR + 10;
    int ms = (ring_pts - 10) / SECTORS + 8;  // FE:205-214: the last sector takes the remainder (< 6 more)
    ms = (ms + 7) / 8 * 8;
    c.max_sector = ms < 64 ? 64 : (ms > MAX_SECTOR ? MAX_SECTOR : ms);
    c.sector_np = 64;
    while (c.sector_np < c.max_sector) c.sector_np <<= 1;
  }
  c.cap_scan = u.max_scan_points; c.cap_map = u.max_map_points;
  // k_solve: one 8-CTA cluster per sequence; voxel leaves below 0.25 m mean tens of thousands of factors per solve (configs[2]: 3.7e4),
  // where a 16-CTA cluster is faster (85 -> 59 us per solve, gpurun_out/r2o_dense*.json).  The choice depends on the configuration
  // only, never on the batch size: the order of the partial sums, hence the last bits of the pose, must not change with it.
  c.lm_cluster = std::min(u.edge_leaf, u.surf_leaf) < 0.25 ? 16 : 8;
  if (const char* e = getenv("VILF_LM_CLUSTER")) { const int v = atoi(e); if (v >= 1 && v <= LM_CLUSTER_MAX) c.lm_cluster = v; }  // experiments only
  c.range_image = (u.flags & VILF_FLAG_RANGE_IMAGE) ? 1 : 0;
  c.horizon = u.horizon_scan; c.ri_down = u.downsample_rate > 0 ? u.downsample_rate : 1; c.ri_edge_thr = u.ri_edge_threshold; c.ri_surf_thr = u.ri_surf_threshold;
  const int NL = C->nlanes;
  const int capS = c.cap_scan, capM = c.cap_map + c.cap_scan;

  CK(cudaSetDevice(C->device));
  CK(init_extract_kernels());
  CK(init_rangeimage_kernels());
  CK(init_solve_kernels());
  CK(cudaStreamCreateWithFlags(&C->st, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&C->copy_st, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    CK(cudaEventCreateWithFlags(&C->extract_done[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&C->h2d_done[i], cudaEventDisableTiming));
  }
  for (int s = 0; s < RING_SLOTS; ++s) {
    CK(cudaEventCreateWithFlags(&C->slots[s].done, cudaEventDisableTiming));
    for (int k = 0; k < MAX_MARKS; ++k) CK(cudaEventCreate(&C->slots[s].stage[k]));
    void* p = nullptr;
    CK(cudaMallocHost(&p, sizeof(LaneVars) * NL));
    C->pinned.push_back(p);
    C->slots[s].vars_pin = reinterpret_cast<LaneVars*>(p);
  }
  CK(dalloc(C, &C->vars_dev, (size_t)NL));
  CK(dalloc(C, &C->trace_dev, (size_t)NL * MAX_OUTER));
  CK(dalloc(C, &C->vv_dev, (size_t)NL * VV_PER_LANE + 1));
  CK(dalloc(C, &C->lanes_dev, (size_t)NL));
  for (int b = 0; b < 2; ++b) CK(dalloc(C, &C->grid_dev[b], (size_t)NL * 2));
  const bool allow_cluster = !(u.flags & VILF_FLAG_NO_CLUSTER);
  g_debug_sync = getenv("VILF_DEBUG_SYNC") != nullptr;
  C->use_graphs = !(u.flags & VILF_FLAG_NO_GRAPH) && !g_debug_sync;
  c.flags_no_cluster = (u.flags & VILF_FLAG_NO_CLUSTER) ? 1 : 0;
  C->cluster_scan = allow_cluster && capS <= CLUSTER_MAX_POINTS;
  C->cluster_map = allow_cluster && capM <= CLUSTER_MAX_POINTS;
  C->cellmap = (u.flags & VILF_FLAG_CELL_MAP) ? true : (u.flags & VILF_FLAG_LEGACY_MAP) ? false : capM > CLUSTER_MAX_POINTS;
  C->no_seed = getenv("VILF_KNN_NO_SEED") != nullptr;
  c.cg[0] = cell_geometry(C, u.edge_leaf);
  c.cg[1] = cell_geometry(C, u.surf_leaf);
  CK(init_cellmap_kernels());
  const int capN = capS + ORPHAN_CAP;  // new points of one update
  int hcapM = 1024;
  while (hcapM < 2 * (capM + capN)) hcapM <<= 1;
  CK(dalloc(C, &C->mv_dev, (size_t)NL * 2 + 2));
  for (int b = 0; b < 2; ++b) { C->merge_host[b].resize(NL * 2); C->build_host[b].resize(NL * 2); }
  std::vector<SortJob> merge_sort(NL * 2), build_sort(NL * 2);
  C->lanes_host.resize(NL);
  std::vector<SortJob> ring_jobs[2] = {std::vector<SortJob>(NL), std::vector<SortJob>(NL)};
  std::vector<VoxJob> vox_scan(NL * 2), vox_map[2] = {std::vector<VoxJob>(NL * 2), std::vector<VoxJob>(NL * 2)};
  std::vector<SortJob> vox_scan_sort(NL * 2), vox_map_sort(NL * 2);
  std::vector<GridJob> grid[2] = {std::vector<GridJob>(NL * 2), std::vector<GridJob>(NL * 2)};
  std::vector<LaneVars> vars0(NL);
  for (int b = 0; b < 2; ++b) {
    CK(dalloc(C, &C->scan_all[b], (size_t)NL * capS));
    CK(dalloc(C, &C->ring_all[b], (size_t)NL * capS));
  }
  for (int l = 0; l < NL; ++l) {
    LaneDev& L = C->lanes_host[l];
    memset(&L, 0, sizeof(L));
    L.v = C->vars_dev + l;
    L.vv = C->vv_dev + (size_t)l * VV_PER_LANE;
    L.trace = C->trace_dev + (size_t)l * MAX_OUTER;
    for (int b = 0; b < 2; ++b) {  // the lanes' scan buffers are rows of ONE array (pitch = cap_scan points): a batch whose host scans have the same pitch moves with one strided copy
      L.scan[b] = C->scan_all[b] + (size_t)l * capS;
      L.ring_in[b] = C->ring_all[b] + (size_t)l * capS;
    }
    int rc = alloc_sort(C, L.ring_sort, &L.v->n_scan[0], nullptr, 8, 1, capS, true);
    if (rc) return rc;
    for (int b = 0; b < 2; ++b) { ring_jobs[b][l] = L.ring_sort; ring_jobs[b][l].n = &L.v->n_scan[b]; }
    CK(dalloc(C, &L.sec_cnt, (size_t)MAX_RINGS * SECTORS));
    CK(dalloc(C, &L.sec_edge, (size_t)MAX_RINGS * SECTORS * EDGES_PER_SECTOR));
    CK(dalloc(C, &L.sec_edge_src, (size_t)MAX_RINGS * SECTORS * EDGES_PER_SECTOR));
    if (c.range_image) {
      CK(dalloc(C, &L.ri_owner, (size_t)c.n_rings * c.horizon));
      CK(dalloc(C, &L.ri_info, (size_t)(MAX_RINGS + 1) * 4));
      CK(dalloc(C, &L.ri_src, (size_t)capS)); CK(dalloc(C, &L.ri_col, (size_t)capS)); CK(dalloc(C, &L.ri_range, (size_t)capS)); CK(dalloc(C, &L.ri_curv, (size_t)capS));
      CK(dalloc(C, &L.ri_picked, (size_t)capS)); CK(dalloc(C, &L.ri_label, (size_t)capS));
    }
    CK(dalloc(C, &L.sec_surf, (size_t)capS));
    CK(dalloc(C, &L.sec_surf_src, (size_t)capS));
    for (int w = 0; w < 2; ++w) {
      CK(dalloc(C, &L.feat[w], (size_t)capS));
      CK(dalloc(C, &L.feat_src[w], (size_t)capS));
      CK(dalloc(C, &L.ds[w], (size_t)capS));
      for (int b = 0; b < 2; ++b) CK(dalloc(C, &L.map[w][b], (size_t)capM));
      CK(dalloc(C, &L.fvalid[w], (size_t)capS));
      CK(dalloc(C, &L.nn_idx[w], (size_t)capS * 5));
      CK(dalloc(C, &L.nn_d2[w], (size_t)capS * 5));
    }
    CK(dalloc(C, &L.edge_pab, (size_t)capS * 9));
    CK(dalloc(C, &L.surf_pnd, (size_t)capS * 7));
    // voxel jobs: scan features (EM:248-251)
    for (int w = 0; w < 2; ++w) {
      VoxJob& J = vox_scan[l * 2 + w];
      memset(&J, 0, sizeof(J));
      J.in = L.feat[w]; J.n_in = w ? &L.v->n_surf : &L.v->n_edge;
      J.leaf = w ? c.surf_leaf : c.edge_leaf;
      J.crop = 0; J.passthrough = 0; J.crop_center = nullptr; J.crop_half = 0;
      J.out = L.ds[w]; J.n_out = &L.v->n_ds[w]; J.cap_out = capS; J.status = &L.v->status;
      J.vv = C->vv_dev + l * VV_PER_LANE + w;
      CK(dalloc(C, &J.head_cnt, (size_t)VOX_G * 8));
      rc = alloc_sort(C, J.sort, J.n_in, &J.vv->bits, 0, 4, capS, false);
      if (rc) return rc;
      vox_scan_sort[l * 2 + w] = J.sort;
    }
    // voxel jobs: map maintenance (EM:327-350), one table per source buffer; sort scratch shared by both
    for (int w = 0; w < 2; ++w) {
      SortJob srt;
      VoxVars* vv = C->vv_dev + l * VV_PER_LANE + 2 + w;
      rc = alloc_sort(C, srt, &L.v->n_cat[w], &vv->bits, 0, 4, capM, false);
      if (rc) return rc;
      int* head_cnt = nullptr;
      CK(dalloc(C, &head_cnt, (size_t)VOX_G * 8));
      for (int b = 0; b < 2; ++b) {
        VoxJob& J = vox_map[b][l * 2 + w];
        memset(&J, 0, sizeof(J));
        J.in = L.map[w][b]; J.n_in = &L.v->n_cat[w];
        J.leaf = w ? c.surf_leaf : c.edge_leaf;
        J.crop = 1; J.passthrough = 0; J.crop_center = &L.v->x[4]; J.crop_half = c.crop_half;
        J.out = L.map[w][b ^ 1]; J.n_out = &L.v->n_map[w]; J.cap_out = c.cap_map; J.status = &L.v->status;
        J.vv = vv; J.head_cnt = head_cnt; J.sort = srt;
        if (C->cluster_map) {  // append + filter + grid build in one cluster kernel (k_cluster.cu)
          J.app_src = L.ds[w]; J.app_n = &L.v->n_ds[w]; J.app_n_map = &L.v->n_map[w]; J.app_pose = L.v->x; J.app_cap = capM;
          J.grid = C->grid_dev[b ^ 1] + l * 2 + w;
        }
      }
      vox_map_sort[l * 2 + w] = srt;
    }
    if (!C->cellmap) {
      for (int w = 0; w < 2; ++w) {
        GridJob G0;
        rc = alloc_grid(C, G0, L.map[w][0], &L.v->n_map[w], capM, w ? u.surf_leaf : u.edge_leaf);
        if (rc) return rc;
        grid[0][l * 2 + w] = G0;
        GridJob G1 = G0;  // the two buffers are never indexed at the same time: share the grid storage
        G1.pts = L.map[w][1];
        grid[1][l * 2 + w] = G1;
      }
    } else {
      for (int w = 0; w < 2; ++w) {
        CK(dalloc(C, &L.ctab[w], (size_t)hcapM));
        CK(dalloc(C, &L.corig[w], (size_t)capM));
        CK(dalloc(C, &L.cmeta[w], 16));
        MergeJob M;
        rc = alloc_merge(C, M, capN, capM);
        if (rc) return rc;
        M.n_map = &L.v->n_map[w];
        M.src = L.ds[w]; M.n_src = &L.v->n_ds[w]; M.pose = L.v->x;
        M.crop_center = &L.v->x[4]; M.crop_half = c.crop_half;
        M.mv = C->mv_dev + l * 2 + w;
        M.table = L.ctab[w]; M.hcap = hcapM; M.meta = L.cmeta[w];
        M.g = c.cg[w]; M.cap_out = c.cap_map; M.status = &L.v->status;
        rc = alloc_sort(C, M.sort, &M.mv->n_in, &M.mv->bits, 0, 4, capN, false);
        if (rc) return rc;
        merge_sort[l * 2 + w] = M.sort;
        C->max_tiles = M.max_tiles;
        for (int b = 0; b < 2; ++b) {
          MergeJob Mb = M;
          Mb.old_pts = L.map[w][b]; Mb.out_pts = L.map[w][b ^ 1];
          C->merge_host[b][l * 2 + w] = Mb;
          CellBuildJob B;
          memset(&B, 0, sizeof(B));
          B.src = L.feat[w]; B.n = w ? &L.v->n_surf : &L.v->n_edge;
          B.dst = L.map[w][b]; B.orig = L.corig[w]; B.table = L.ctab[w]; B.hcap = hcapM; B.meta = L.cmeta[w];
          B.sort = vox_map_sort[l * 2 + w]; B.sort.n = B.n; B.sort.bits = B.meta + 8;
          B.g = c.cg[w]; B.status = &L.v->status;
          C->build_host[b][l * 2 + w] = B;
        }
        build_sort[l * 2 + w] = C->build_host[0][l * 2 + w].sort;
      }
    }
    LaneVars& V = vars0[l];
    memset(&V, 0, sizeof(V));
    V.x[3] = 1.0;                                   // EM:383
    V.odom[0] = V.odom[4] = V.odom[8] = 1.0;        // EM:88-89
    V.odom_last[0] = V.odom_last[4] = V.odom_last[8] = 1.0;
  }
  CK(cudaMemcpy(C->vars_dev, vars0.data(), sizeof(LaneVars) * NL, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(C->lanes_dev, C->lanes_host.data(), sizeof(LaneDev) * NL, cudaMemcpyHostToDevice));
  for (int b = 0; b < 2; ++b) {
    CK(dalloc(C, &C->ring_jobs_dev[b], (size_t)NL));
    CK(cudaMemcpy(C->ring_jobs_dev[b], ring_jobs[b].data(), sizeof(SortJob) * NL, cudaMemcpyHostToDevice));
  }
  CK(dalloc(C, &C->vox_scan_dev, (size_t)NL * 2));
  CK(cudaMemcpy(C->vox_scan_dev, vox_scan.data(), sizeof(VoxJob) * NL * 2, cudaMemcpyHostToDevice));
  CK(dalloc(C, &C->vox_scan_sort_dev, (size_t)NL * 2));
  CK(cudaMemcpy(C->vox_scan_sort_dev, vox_scan_sort.data(), sizeof(SortJob) * NL * 2, cudaMemcpyHostToDevice));
  for (int b = 0; b < 2; ++b) {
    CK(dalloc(C, &C->vox_map_dev[b], (size_t)NL * 2));
    CK(cudaMemcpy(C->vox_map_dev[b], vox_map[b].data(), sizeof(VoxJob) * NL * 2, cudaMemcpyHostToDevice));
    CK(dalloc(C, &C->vox_map_sort_dev[b], (size_t)NL * 2));
    CK(cudaMemcpy(C->vox_map_sort_dev[b], vox_map_sort.data(), sizeof(SortJob) * NL * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(C->grid_dev[b], grid[b].data(), sizeof(GridJob) * NL * 2, cudaMemcpyHostToDevice));
  }
  if (C->cellmap) {
    for (int b = 0; b < 2; ++b) {
      CK(dalloc(C, &C->merge_dev[b], (size_t)NL * 2));
      CK(cudaMemcpy(C->merge_dev[b], C->merge_host[b].data(), sizeof(MergeJob) * NL * 2, cudaMemcpyHostToDevice));
      CK(dalloc(C, &C->build_dev[b], (size_t)NL * 2));
      CK(cudaMemcpy(C->build_dev[b], C->build_host[b].data(), sizeof(CellBuildJob) * NL * 2, cudaMemcpyHostToDevice));
    }
    CK(dalloc(C, &C->merge_sort_dev, (size_t)NL * 2));
    CK(cudaMemcpy(C->merge_sort_dev, merge_sort.data(), sizeof(SortJob) * NL * 2, cudaMemcpyHostToDevice));
    CK(dalloc(C, &C->build_sort_dev, (size_t)NL * 2));
    CK(cudaMemcpy(C->build_sort_dev, build_sort.data(), sizeof(SortJob) * NL * 2, cudaMemcpyHostToDevice));
  }
  // aux
  C->cap_aux = capM;
  CK(dalloc(C, &C->aux_in, (size_t)capM));
  CK(dalloc(C, &C->aux_out, (size_t)capM));
  CK(dalloc(C, &C->aux_n, 16));
  CK(dalloc(C, &C->aux_pose, 8));
  CK(dalloc(C, &C->aux_idx, (size_t)capM * 5));
  CK(dalloc(C, &C->aux_d2, (size_t)capM * 5));
  CK(dalloc(C, &C->aux_vox_dev, 1));
  CK(dalloc(C, &C->aux_sort_dev, 1));
  CK(dalloc(C, &C->aux_grid_dev, 1));
  {
    VoxJob& J = C->aux_vox_host;
    memset(&J, 0, sizeof(J));
    J.in = C->aux_in; J.n_in = C->aux_n; J.out = C->aux_out; J.n_out = C->aux_n + 1; J.cap_out = capM;
    J.status = C->aux_n + 3; J.vv = C->vv_dev + NL * VV_PER_LANE; J.crop_center = C->aux_pose;
    CK(dalloc(C, &J.head_cnt, (size_t)VOX_G * 8));
    int rc = alloc_sort(C, J.sort, J.n_in, &J.vv->bits, 0, 4, capM, false);
    if (rc) return rc;
    CK(cudaMemcpy(C->aux_sort_dev, &J.sort, sizeof(SortJob), cudaMemcpyHostToDevice));
    GridJob G;
    rc = alloc_grid(C, G, C->aux_in, C->aux_n, capM, u.edge_leaf < u.surf_leaf ? u.edge_leaf : u.surf_leaf);
    C->aux_grid_host = G;
    if (rc) return rc;
    CK(cudaMemcpy(C->aux_grid_dev, &G, sizeof(GridJob), cudaMemcpyHostToDevice));
  }
  if (C->cellmap) {
    int hc = 1024;
    while (hc < 4 * capM) hc <<= 1;
    C->aux_hcap = hc;
    for (int b = 0; b < 2; ++b) CK(dalloc(C, &C->aux_cm_pts[b], (size_t)capM));
    CK(dalloc(C, &C->aux_ctab, (size_t)hc));
    CK(dalloc(C, &C->aux_corig, (size_t)capM));
    CK(dalloc(C, &C->aux_cmeta, 16));
    CK(dalloc(C, &C->aux_build_dev, 1));
    CK(dalloc(C, &C->aux_build_sort_dev, 1));
    CK(dalloc(C, &C->aux_merge_dev, 1));
    CK(dalloc(C, &C->aux_merge_sort_dev, 1));
    CellBuildJob& B = C->aux_build_host;
    memset(&B, 0, sizeof(B));
    B.src = C->aux_in; B.n = C->aux_n; B.dst = C->aux_cm_pts[0]; B.orig = C->aux_corig; B.table = C->aux_ctab; B.hcap = hc; B.meta = C->aux_cmeta;
    B.sort = C->aux_vox_host.sort; B.sort.n = B.n; B.sort.bits = B.meta + 8;
    B.g = c.cg[u.edge_leaf < u.surf_leaf ? 0 : 1]; B.status = C->aux_n + 3;
    MergeJob& M = C->aux_merge_host;
    int rc = alloc_merge(C, M, capM, capM);
    if (rc) return rc;
    C->aux_max_tiles = M.max_tiles;
    M.old_pts = C->aux_cm_pts[0]; M.out_pts = C->aux_cm_pts[1]; M.n_map = C->aux_n + 5;
    M.src = C->aux_out; M.n_src = C->aux_n + 2; M.pose = nullptr; M.crop_center = nullptr; M.crop_half = 0;
    M.mv = C->mv_dev + NL * 2;
    M.table = C->aux_ctab; M.hcap = hc; M.meta = C->aux_cmeta; M.g = B.g; M.cap_out = capM; M.status = C->aux_n + 3;
    M.sort = C->aux_vox_host.sort; M.sort.n = &M.mv->n_in; M.sort.bits = &M.mv->bits;
  }
  CK(dalloc(C, &C->depth_feat, (size_t)DEPTH_MAX_FEATURES * 3));
  CK(dalloc(C, &C->depth_out, (size_t)DEPTH_MAX_FEATURES));
  CK(dalloc(C, &C->depth_nn, (size_t)DEPTH_MAX_FEATURES * 3));
  CK(dalloc(C, &C->depth_T, 16));
  C->last_sel.assign(NL, -1);
  C->cur.assign(NL, 0); C->have_map.assign(NL, 0); C->have_feat.assign(NL, 0); C->last_init.assign(NL, 0); C->frame_no.assign(NL, 0);
  CK(cudaDeviceSynchronize());
  return VILF_OK;
}

void destroy_ctx(Ctx* C) {
  cudaSetDevice(C->device);
  if (C->st) cudaStreamSynchronize(C->st);
  if (C->copy_st) cudaStreamSynchronize(C->copy_st);
  for (auto& kv : C->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  for (void* p : C->allocs) cudaFree(p);
  for (void* p : C->pinned) cudaFreeHost(p);
  for (int i = 0; i < 2; ++i) { if (C->extract_done[i]) cudaEventDestroy(C->extract_done[i]); if (C->h2d_done[i]) cudaEventDestroy(C->h2d_done[i]); }
  for (int s = 0; s < RING_SLOTS; ++s) {
    if (C->slots[s].done) cudaEventDestroy(C->slots[s].done);
    for (int k = 0; k < MAX_MARKS; ++k) if (C->slots[s].stage[k]) cudaEventDestroy(C->slots[s].stage[k]);
  }
  if (C->st) cudaStreamDestroy(C->st);
  if (C->copy_st) cudaStreamDestroy(C->copy_st);
  for (vilf_handle* h : C->handles) delete h;
  delete C;
}

int status_to_rc(Ctx* C, int status) {
  if (status & ST_SECTOR_TOO_LONG) return fail(C, VILF_ERR_UNSUPPORTED, "a ring has more returns than the sector kernel supports (6*2048+10)");
  if (status & ST_MAP_CAPACITY) return fail(C, VILF_ERR_CAPACITY, "local map exceeds max_map_points");
  if (status & ST_SCAN_CAPACITY) return fail(C, VILF_ERR_CAPACITY, "scan exceeds max_scan_points");
  if (status & ST_KEY_RANGE) return fail(C, VILF_ERR_UNSUPPORTED, "cloud spans more than 2^32 (cell, voxel) keys at this leaf size; use VILF_FLAG_LEGACY_MAP");
  if (status & ST_PCL_GUARD) return fail(C, VILF_ERR_UNSUPPORTED, "leaf size too small for the cropped map (PCL's int32 guard would skip the voxel filter); use VILF_FLAG_LEGACY_MAP");
  if (status & ST_ORPHANS) return fail(C, VILF_ERR_CAPACITY, "more than 256 voxel centroids crossed a voxel face in one map update");
  return VILF_OK;
}

// createSubMap (EM:298-352) for lanes [lane0, lane0+nl): append the voxel-filtered scan features at the current pose,
// crop + voxel-filter both maps into the other buffer, rebuild the search grids, flip the buffers.
void issue_submap(Ctx* C, const Launch& L, int lane0, int nl, int cur, ProfSink* sink) {
  if (C->cellmap) {
    launch_cell_update(L, C->merge_dev[cur] + lane0 * 2, C->merge_sort_dev + lane0 * 2, nl * 2, C->max_tiles,
                       !C->cfg.flags_no_cluster && C->cfg.cap_scan + ORPHAN_CAP <= CLUSTER_MAX_POINTS);
  } else if (C->cluster_map) {
    launch_voxel_cluster(L, C->vox_map_dev[cur] + lane0 * 2, nl * 2, false, C->cfg);
  } else {
    launch_map_append(L, C->lanes_dev, lane0, nl, cur, C->cfg);
    launch_voxel(L, C->vox_map_dev[cur] + lane0 * 2, nl * 2, C->vox_map_sort_dev[cur] + lane0 * 2, true);
    if (sink) sink->phase = 4;
    launch_grid_build(L, C->grid_dev[cur ^ 1] + lane0 * 2, nl * 2, C->cfg);
  }
}
void enqueue_submap(Ctx* C, const Launch& L, int lane0, int nl, ProfSink* sink) {
  const int cur = C->cur[lane0];
  issue_submap(C, L, lane0, nl, cur, sink);
  for (int l = lane0; l < lane0 + nl; ++l) C->cur[l] = cur ^ 1;
}

// Stage 1: featureExtraction::extractFeature (FE:223-232) or, with VILF_FLAG_RANGE_IMAGE, featureExtract::extractFeature (FX:96-115).
void stage1(Ctx* C, const Launch& L, int lane0, int nl, int sel) {
  if (C->cfg.range_image) launch_extract_range_image(L, C->lanes_dev, lane0, nl, sel, C->cfg);
  else launch_extract(L, C->lanes_dev, C->ring_jobs_dev[sel], lane0, nl, sel, C->cfg);
}

// The per-frame launch sequence for lanes [lane0, lane0+nl), which all share `cur`, `first` and the scan slot.
// with_extract = 0: features were uploaded by the caller (vilf_update_points / vilf_map_init_points).
void issue_frame(Ctx* C, int lane0, int nl, bool first, bool with_extract, int sel, int cur, ProfSink* sink);

int enqueue_frame(Ctx* C, int lane0, int nl, bool first, bool with_extract, int sel, Slot* S) {
  ProfSink* sink = nullptr;
  if (S && S->profiled) {
    sink = &S->sink;
    sink->ev = S->stage; sink->tag = S->stage_tag; sink->n = 0; sink->cap = MAX_MARKS; sink->phase = 0;
    cudaEventRecord(sink->ev[0], C->st);  // frame start
    sink->tag[sink->n++] = -1;
  }
  const int cur = C->cur[lane0];
  if (C->use_graphs && !sink && !first && with_extract) {
    // The steady-state frame is a fixed launch sequence whose arguments depend only on (lanes, scan buffer, map buffer): all
    // element counts live in device memory.  It is captured once per combination and replayed with one cudaGraphLaunch.
    Ctx::FrameGraph& G = C->graphs[std::make_tuple(lane0, nl, sel, cur)];
    if (!G.exec) {
      const int64_t before = C->launches;
      cudaGraph_t graph = nullptr;
      CK(cudaStreamBeginCapture(C->st, cudaStreamCaptureModeThreadLocal));
      issue_frame(C, lane0, nl, first, with_extract, sel, cur, nullptr);
      // a failure between Begin and End must not leave the stream in capture mode or an empty entry in the table
      cudaError_t ce = cudaStreamEndCapture(C->st, &graph);
      if (ce == cudaSuccess) ce = cudaGraphInstantiate(&G.exec, graph, 0);
      if (graph) cudaGraphDestroy(graph);
      const int captured = (int)(C->launches - before);
      C->launches = before;
      if (ce != cudaSuccess) {
        G.exec = nullptr;
        C->graphs.erase(std::make_tuple(lane0, nl, sel, cur));
        cudaGetLastError();
        snprintf(C->err, sizeof(C->err), "CUDA graph capture of the frame failed: %s", cudaGetErrorString(ce));
        return VILF_ERR_CUDA;
      }
      G.launches = captured;
    }
    nvtxRangePushA("vilf:frame(graph)");
    const cudaError_t ge = cudaGraphLaunch(G.exec, C->st);
    nvtxRangePop();
    CK(ge);
    C->launches += G.launches;
  } else {
    issue_frame(C, lane0, nl, first, with_extract, sel, cur, sink);
  }
  if (!first) for (int l = lane0; l < lane0 + nl; ++l) C->cur[l] = cur ^ 1;
  for (int l = lane0; l < lane0 + nl; ++l) {
    C->have_map[l] = 1; C->have_feat[l] = 1; C->last_init[l] = first ? 1 : 0; C->frame_no[l] += 1;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(C->err, sizeof(C->err), "kernel launch failed: %s", cudaGetErrorString(e));
    return VILF_ERR_CUDA;
  }
  return VILF_OK;
}

// The launch sequence of one frame (no host-side state changes: it is also what a graph capture records).
void issue_frame(Ctx* C, int lane0, int nl, bool first, bool with_extract, int sel, int cur, ProfSink* sink) {
  const Launch L = mk(C, sink);
  // NVTX range per stage (SURVEY.md §5 tracing row): visible in Nsight Systems around the launches (or, for a graph-replayed frame,
  // around its one-time capture; the replay itself is the range "vilf:frame(graph)" in enqueue_frame)
  static const char* const kStage[5] = {"vilf:extract", "vilf:scan_downsample", "vilf:associate+solve", "vilf:map_update", "vilf:search_build"};
  bool open = false;
  auto phase = [&](int p) { if (sink) sink->phase = p; if (open) nvtxRangePop(); nvtxRangePushA(kStage[p]); open = true; };
  struct Closer { bool& o; ~Closer() { if (o) nvtxRangePop(); } } closer{open};
  const ConfigDev& cfg = C->cfg;
  phase(0);
  launch_frame_reset(L, C->lanes_dev, lane0, nl, C->vv_dev, VV_PER_LANE, first ? 0 : 1);
  if (with_extract) stage1(C, L, lane0, nl, sel);
  if (first && C->cellmap) {
    phase(4);
    launch_cell_build(L, C->build_dev[cur] + lane0 * 2, C->build_sort_dev + lane0 * 2, nl * 2);
    phase(3);
    launch_map_init_commit(L, C->lanes_dev, lane0, nl, cfg);
  } else if (first) {
    phase(3);
    launch_map_init(L, C->lanes_dev, lane0, nl, cur, cfg);
    phase(4);
    build_grids(C, L, C->grid_dev[cur] + lane0 * 2, nl * 2, C->cluster_map);
  } else {
    phase(1);
    if (C->cluster_scan) launch_voxel_cluster(L, C->vox_scan_dev + lane0 * 2, nl * 2, with_extract, cfg);
    else launch_voxel(L, C->vox_scan_dev + lane0 * 2, nl * 2, C->vox_scan_sort_dev + lane0 * 2, with_extract);
    phase(2);
    for (int it = 0; it < cfg.outer_iters; ++it) {
      if (C->cellmap) launch_knn_cell_fit(L, C->lanes_dev, lane0, nl, cur, cfg, nullptr, it > 0 && !C->no_seed);
      else launch_knn_fit(L, C->lanes_dev, C->grid_dev[cur], lane0, nl, cur, cfg, nullptr);
      launch_solve(L, C->lanes_dev, lane0, nl, it, it == cfg.outer_iters - 1 ? 1 : 0, cfg, cfg.lm_max_iters);
    }
    phase(3);
    issue_submap(C, L, lane0, nl, cur, sink);
  }
}

// Pinned slabs handed out by vilf_host_alloc: the strided batch copy reads rows beyond a scan's last point (up to the longest scan of
// the batch), which is only known to be mapped memory when the whole batch lies inside one such slab.
std::mutex g_slab_mu;
std::map<const char*, size_t> g_slabs;  // base -> bytes
bool inside_one_slab(const void* first, const void* last_end) {
  std::lock_guard<std::mutex> lk(g_slab_mu);
  auto it = g_slabs.upper_bound((const char*)first);
  if (it == g_slabs.begin()) return false;
  --it;
  return (const char*)first >= it->first && (const char*)last_end <= it->first + it->second;
}

bool lanes_uniform(Ctx* C, int lane0, int nl) {
  for (int l = lane0 + 1; l < lane0 + nl; ++l)
    if (C->cur[l] != C->cur[lane0] || C->have_map[l] != C->have_map[lane0]) return false;
  return true;
}

int read_vars(Ctx* C, int lane, LaneVars* out) {
  CK(cudaMemcpyAsync(out, C->vars_dev + lane, sizeof(LaneVars), cudaMemcpyDeviceToHost, C->st));
  CK(cudaStreamSynchronize(C->st));
  return VILF_OK;
}

int submit_common(Ctx* C, int lane0, int nl, const float* const* xyzi, const int* n, const uint16_t* const* ring, bool dev_src, int64_t* ticket) {
  CK(cudaSetDevice(C->device));
  for (int i = 0; i < nl; ++i) {
    if (n[i] < 0 || !xyzi[i]) return fail(C, VILF_ERR_INVALID, "null scan or negative point count");
    if (n[i] > C->cfg.cap_scan) return fail(C, VILF_ERR_CAPACITY, "scan exceeds max_scan_points");
    if (C->cfg.n_scan == 0 && (!ring || !ring[i])) return fail(C, VILF_ERR_INVALID, "n_scan == 0 needs explicit ring ids");
  }
  if (!lanes_uniform(C, lane0, nl)) return fail(C, VILF_ERR_STATE, "sequences of a batch are not in lock-step");
  const int64_t t = C->next_ticket;
  Slot& S = C->slots[t % RING_SLOTS];
  if (S.ticket >= 0) return fail(C, VILF_ERR_STATE, "too many frames in flight: wait for earlier tickets first");
  const int sel = (int)(C->scan_sel & 1);
  C->scan_sel++;
  // H2D (or D2D) on the copy stream into scan buffer `sel`, which the extract kernels of two frames ago released
  CK(cudaStreamWaitEvent(C->copy_st, C->extract_done[sel], 0));
  // Scans of a batch that lie in one array with a row pitch of max_scan_points points (how a driver that fills fixed slots, and
  // bench.py, lay them out) move with ONE strided copy instead of one copy per scan: every cudaMemcpyAsync leaves the copy engine idle
  // for ~4 us, which at 1.8 MB per scan is 10 % of the link (tools/h2d_probe.py: 49.6 GB/s with a copy per scan, 55.2 GB/s with a copy
  // per 32 scans).  Rows are read up to the longest scan of the batch (inside the row's own pitch), so the path is taken only when the
  // whole batch lies inside ONE slab from vilf_host_alloc; the last row is copied with its own length.
  const cudaMemcpyKind kind = dev_src ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const size_t pitch = (size_t)C->cfg.cap_scan * 16;
  bool strided = nl > 2 && !dev_src, strided_ring = nl > 2 && !dev_src && ring != nullptr;  // (device-to-device: separate copies measured faster)
  int maxn = 0;
  for (int i = 0; i < nl; ++i) {
    maxn = n[i] > maxn ? n[i] : maxn;
    if ((const char*)xyzi[i] != (const char*)xyzi[0] + (size_t)i * pitch) strided = false;
    if (strided_ring && (!ring[i] || (const char*)ring[i] != (const char*)ring[0] + (size_t)i * (pitch / 8))) strided_ring = false;
    S.vars_pin[i].n_scan[sel] = n[i];
  }
  if (maxn == 0) strided = strided_ring = false;
  if (strided && !inside_one_slab(xyzi[0], (const char*)xyzi[nl - 1] + (size_t)n[nl - 1] * 16)) strided = false;
  if (strided_ring && !inside_one_slab(ring[0], (const char*)ring[nl - 1] + (size_t)n[nl - 1] * 2)) strided_ring = false;
  const int rows2d = nl - 1;
  if (strided) CK(cudaMemcpy2DAsync(C->lanes_host[lane0].scan[sel], pitch, xyzi[0], pitch, (size_t)maxn * 16, (size_t)rows2d, kind, C->copy_st));
  if (strided_ring) CK(cudaMemcpy2DAsync(C->lanes_host[lane0].ring_in[sel], pitch / 8, ring[0], pitch / 8, (size_t)maxn * 2, (size_t)rows2d, kind, C->copy_st));
  for (int i = 0; i < nl; ++i) {
    LaneDev& L = C->lanes_host[lane0 + i];
    if (n[i] > 0 && !(strided && i < rows2d)) CK(cudaMemcpyAsync(L.scan[sel], xyzi[i], (size_t)n[i] * 16, kind, C->copy_st));
    if (ring && ring[i] && n[i] > 0 && !(strided_ring && i < rows2d)) CK(cudaMemcpyAsync(L.ring_in[sel], ring[i], (size_t)n[i] * 2, kind, C->copy_st));
  }
  // the point counts of all lanes in ONE strided copy (a 4-byte column of the pinned LaneVars array -> the same column on the device)
  CK(cudaMemcpy2DAsync(&C->vars_dev[lane0].n_scan[sel], sizeof(LaneVars), &S.vars_pin[0].n_scan[sel], sizeof(LaneVars), sizeof(int), (size_t)nl,
                       cudaMemcpyHostToDevice, C->copy_st));
  CK(cudaEventRecord(C->h2d_done[sel], C->copy_st));
  CK(cudaStreamWaitEvent(C->st, C->h2d_done[sel], 0));
  S.profiled = C->profile;
  S.first = !C->have_map[lane0];
  int rc = enqueue_frame(C, lane0, nl, S.first, true, sel, &S);
  if (rc) {  // nothing of this frame was launched: give the scan buffer back so that the buffer parity and the event chain stay in step
    C->scan_sel--;
    return rc;
  }
  CK(cudaEventRecord(C->extract_done[sel], C->st));  // conservative: the whole frame (the scan is only read by stage 1)
  CK(cudaMemcpyAsync(S.vars_pin, C->vars_dev + lane0, sizeof(LaneVars) * nl, cudaMemcpyDeviceToHost, C->st));
  CK(cudaEventRecord(S.done, C->st));
  S.ticket = t; S.lane0 = lane0; S.nl = nl;
  for (int l = lane0; l < lane0 + nl; ++l) C->last_sel[l] = sel;
  C->next_ticket++;
  *ticket = t;
  return VILF_OK;
}

int wait_common(Ctx* C, int lane0, int nl, int64_t ticket, double* poses) {
  CK(cudaSetDevice(C->device));
  if (ticket < 0) return fail(C, VILF_ERR_INVALID, "bad ticket");
  Slot& S = C->slots[ticket % RING_SLOTS];
  if (S.ticket != ticket || S.lane0 != lane0 || S.nl != nl) return fail(C, VILF_ERR_INVALID, "unknown ticket");
  CK(cudaEventSynchronize(S.done));
  S.ticket = -1;
  if (S.profiled && S.sink.n > 1) {
    for (int k = 1; k < S.sink.n; ++k) {
      float ms = 0;
      const int tag = S.stage_tag[k];
      if (cudaEventElapsedTime(&ms, S.stage[k - 1], S.stage[k]) == cudaSuccess && tag >= 0 && tag < PROF_TAGS) {
        C->kernel_ms[tag] += ms;
        C->kernel_cnt[tag] += 1;
      }
    }
    float ms = 0;
    if (cudaEventElapsedTime(&ms, S.stage[0], S.stage[S.sink.n - 1]) == cudaSuccess) C->frame_ms += ms;
    C->prof_frames += 1;
  }
  int status = 0;
  for (int i = 0; i < nl; ++i) {
    if (poses) memcpy(poses + 7 * i, S.vars_pin[i].x, 7 * sizeof(double));
    status |= S.vars_pin[i].status;
  }
  return status_to_rc(C, status);
}

int host_sort_passes(int bits) { int p = (bits + SORT_RADIX_BITS - 1) / SORT_RADIX_BITS; return p < 1 ? 1 : (p > 4 ? 4 : p); }

// Cell-ordered map `w` of a lane -> the reference's map order (ascending PCL voxel index, or the input order of a map that was
// never filtered) in aux_out on the device; optionally the PCL index of every stored point.  Off the per-frame path
// (vilf_get_cloud, vilf_factors).
int map_to_pcl(Ctx* C, int lane, int w, int* n_out, std::vector<int32_t>* rank) {
  LaneDev& L = C->lanes_host[lane];
  const int cur = C->cur[lane];
  LaneVars V;
  int rc = read_vars(C, lane, &V);
  if (rc) return rc;
  const int n = V.n_map[w];
  *n_out = n;
  if (n > C->cap_aux) return fail(C, VILF_ERR_CAPACITY, "map exceeds the staging capacity");
  if (rank) rank->assign((size_t)(n > 0 ? n : 1), -1);
  if (n == 0) return VILF_OK;
  int meta[2];
  CK(cudaMemcpy(meta, L.cmeta[w], sizeof(meta), cudaMemcpyDeviceToHost));
  if (meta[1]) {  // never filtered since it was loaded: the original positions are stored
    launch_cell_unpermute(mk(C), L.map[w][cur], L.corig[w], &L.v->n_map[w], C->aux_out, C->cap_aux);
    CK(cudaGetLastError());
    if (rank) CK(cudaMemcpyAsync(rank->data(), L.corig[w], (size_t)n * 4, cudaMemcpyDeviceToHost, C->st));
    CK(cudaStreamSynchronize(C->st));
    return VILF_OK;
  }
  // a filtered map: PCL's output order is ascending voxel index -> stable sort by PCL's own index, every point emitted
  VoxJob J = C->aux_vox_host;
  J.in = L.map[w][cur]; J.n_in = &L.v->n_map[w];
  J.leaf = w ? C->cfg.surf_leaf : C->cfg.edge_leaf;
  J.crop = 0; J.passthrough = 0; J.emit_all = 1; J.sort.n = J.n_in;
  CK(cudaMemcpyAsync(C->aux_vox_dev, &J, sizeof(J), cudaMemcpyHostToDevice, C->st));
  CK(cudaMemcpyAsync(C->aux_sort_dev, &J.sort, sizeof(SortJob), cudaMemcpyHostToDevice, C->st));
  VoxVars vv;
  memset(&vv, 0, sizeof(vv));
  vv.bbox[0] = vv.bbox[1] = vv.bbox[2] = INT_MAX;
  vv.bbox[3] = vv.bbox[4] = vv.bbox[5] = INT_MIN;
  CK(cudaMemcpyAsync(J.vv, &vv, sizeof(vv), cudaMemcpyHostToDevice, C->st));
  CK(cudaStreamSynchronize(C->st));
  launch_voxel(mk(C), C->aux_vox_dev, 1, C->aux_sort_dev, false);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(&vv, J.vv, sizeof(vv), cudaMemcpyDeviceToHost, C->st));
  CK(cudaMemcpyAsync(C->aux_sort_dev, &C->aux_vox_host.sort, sizeof(SortJob), cudaMemcpyHostToDevice, C->st));
  CK(cudaStreamSynchronize(C->st));
  if (vv.guard) return fail(C, VILF_ERR_UNSUPPORTED, "PCL's int32 voxel-index guard fires on this map; its order is undefined");
  if (rank) {
    std::vector<uint32_t> val((size_t)n);
    CK(cudaMemcpy(val.data(), J.sort.val[host_sort_passes(vv.bits) & 1], (size_t)n * 4, cudaMemcpyDeviceToHost));
    for (int r = 0; r < n; ++r) (*rank)[val[r]] = r;
  }
  return VILF_OK;
}

// Arbitrary cloud already staged in aux_in / aux_n[0] -> cell-ordered map described by B.
int run_cell_build(Ctx* C, const CellBuildJob& B) {
  CK(cudaMemcpyAsync(C->aux_build_dev, &B, sizeof(B), cudaMemcpyHostToDevice, C->st));
  CK(cudaMemcpyAsync(C->aux_build_sort_dev, &B.sort, sizeof(SortJob), cudaMemcpyHostToDevice, C->st));
  CK(cudaStreamSynchronize(C->st));  // B lives on the caller's stack
  launch_cell_build(mk(C), C->aux_build_dev, C->aux_build_sort_dev, 1);
  CK(cudaGetLastError());
  return VILF_OK;
}

int upload_features(Ctx* C, int lane, const float* edge, int ne, const float* surf, int ns) {
  if (ne < 0 || ns < 0 || (ne > 0 && !edge) || (ns > 0 && !surf)) return fail(C, VILF_ERR_INVALID, "bad feature clouds");
  if (ne > C->cfg.cap_scan || ns > C->cfg.cap_scan) return fail(C, VILF_ERR_CAPACITY, "feature cloud exceeds max_scan_points");
  LaneDev& L = C->lanes_host[lane];
  if (ne) CK(cudaMemcpyAsync(L.feat[0], edge, (size_t)ne * 16, cudaMemcpyHostToDevice, C->st));
  if (ns) CK(cudaMemcpyAsync(L.feat[1], surf, (size_t)ns * 16, cudaMemcpyHostToDevice, C->st));
  int cnt[2] = {ne, ns};
  CK(cudaMemcpyAsync(&L.v->n_edge, cnt, sizeof(cnt), cudaMemcpyHostToDevice, C->st));  // n_edge, n_surf are adjacent
  CK(cudaStreamSynchronize(C->st));
  return VILF_OK;
}

int finish_sync(Ctx* C, int lane, double* pose_out) {
  LaneVars V;
  int rc = read_vars(C, lane, &V);
  if (rc) return rc;
  if (pose_out) memcpy(pose_out, V.x, 7 * sizeof(double));
  return status_to_rc(C, V.status);
}

int run_aux_voxel(Ctx* C, const float* pts, int n, float leaf, int crop, const double* center, double half, const double* mn, const double* mx,
                  int passthrough, float* out, int cap, int* n_out, int* guard) {
  CK(cudaSetDevice(C->device));
  if (n < 0 || (n > 0 && !pts) || !n_out) return fail(C, VILF_ERR_INVALID, "bad cloud");
  if (n > C->cap_aux) return fail(C, VILF_ERR_CAPACITY, "cloud exceeds max_map_points + max_scan_points");
  VoxJob J = C->aux_vox_host;
  J.leaf = leaf; J.crop = crop; J.passthrough = passthrough; J.crop_half = half;
  if (crop == 2) for (int a = 0; a < 3; ++a) { J.crop_lo[a] = (float)mn[a]; J.crop_hi[a] = (float)mx[a]; }
  CK(cudaMemcpyAsync(C->aux_vox_dev, &J, sizeof(J), cudaMemcpyHostToDevice, C->st));
  if (n) CK(cudaMemcpyAsync(C->aux_in, pts, (size_t)n * 16, cudaMemcpyHostToDevice, C->st));
  int hdr[4] = {n, 0, 0, 0};
  CK(cudaMemcpyAsync(C->aux_n, hdr, sizeof(hdr), cudaMemcpyHostToDevice, C->st));
  if (crop == 1) CK(cudaMemcpyAsync(C->aux_pose, center, 3 * sizeof(double), cudaMemcpyHostToDevice, C->st));
  VoxVars vv;
  memset(&vv, 0, sizeof(vv));
  vv.bbox[0] = vv.bbox[1] = vv.bbox[2] = INT_MAX;
  vv.bbox[3] = vv.bbox[4] = vv.bbox[5] = INT_MIN;
  CK(cudaMemcpyAsync(J.vv, &vv, sizeof(vv), cudaMemcpyHostToDevice, C->st));
  CK(cudaStreamSynchronize(C->st));  // the staged host structs above live on this stack frame
  if (!(C->ucfg.flags & VILF_FLAG_NO_CLUSTER) && n <= CLUSTER_MAX_POINTS) launch_voxel_cluster(mk(C), C->aux_vox_dev, 1, false, C->cfg);
  else launch_voxel(mk(C), C->aux_vox_dev, 1, C->aux_sort_dev, false);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(hdr, C->aux_n, sizeof(hdr), cudaMemcpyDeviceToHost, C->st));
  CK(cudaMemcpyAsync(&vv, J.vv, sizeof(vv), cudaMemcpyDeviceToHost, C->st));
  CK(cudaStreamSynchronize(C->st));
  *n_out = hdr[1];
  if (guard) *guard = vv.guard;
  if (hdr[1] > cap) return fail(C, VILF_ERR_CAPACITY, "output buffer too small");
  if (out && hdr[1] > 0) CK(cudaMemcpy(out, C->aux_out, (size_t)hdr[1] * 16, cudaMemcpyDeviceToHost));
  return VILF_OK;
}

// Upload explicit factors into the lane's factor slots and run the solve kernel (vilf_solve / vilf_normal_equations).
int run_explicit_solve(Ctx* C, int lane, const double* pose, const double* edge_pab, int ne, const double* surf_pnd, int ns, int max_iters,
                       double* pose_out, SolveTraceDev* trace_out) {
  CK(cudaSetDevice(C->device));
  if (ne < 0 || ns < 0 || ne > C->cfg.cap_scan || ns > C->cfg.cap_scan) return fail(C, VILF_ERR_CAPACITY, "too many factors");
  LaneDev& L = C->lanes_host[lane];
  LaneVars V;
  int rc = read_vars(C, lane, &V);
  if (rc) return rc;
  LaneVars W = V;
  memcpy(W.x, pose, 7 * sizeof(double));
  W.n_ds[0] = ne; W.n_ds[1] = ns; W.opt_ran = 1;
  CK(cudaMemcpyAsync(C->vars_dev + lane, &W, sizeof(W), cudaMemcpyHostToDevice, C->st));
  if (ne) {
    CK(cudaMemcpyAsync(L.edge_pab, edge_pab, (size_t)ne * 9 * sizeof(double), cudaMemcpyHostToDevice, C->st));
    CK(cudaMemsetAsync(L.fvalid[0], 1, (size_t)ne, C->st));
  }
  if (ns) {
    CK(cudaMemcpyAsync(L.surf_pnd, surf_pnd, (size_t)ns * 7 * sizeof(double), cudaMemcpyHostToDevice, C->st));
    CK(cudaMemsetAsync(L.fvalid[1], 1, (size_t)ns, C->st));
  }
  launch_solve(mk(C), C->lanes_dev, lane, 1, 0, 0, C->cfg, max_iters);
  CK(cudaGetLastError());
  LaneVars R;
  CK(cudaMemcpyAsync(&R, C->vars_dev + lane, sizeof(R), cudaMemcpyDeviceToHost, C->st));
  CK(cudaMemcpyAsync(trace_out, L.trace, sizeof(SolveTraceDev), cudaMemcpyDeviceToHost, C->st));
  CK(cudaMemcpyAsync(C->vars_dev + lane, &V, sizeof(V), cudaMemcpyHostToDevice, C->st));  // restore pose / counts
  CK(cudaStreamSynchronize(C->st));
  if (pose_out) memcpy(pose_out, R.x, 7 * sizeof(double));
  return VILF_OK;
}

}  // namespace

#define HCHECK(h)                              \
  if (!(h) || !(h)->ctx) return VILF_ERR_INVALID; \
  Ctx* C = (h)->ctx;                           \
  (void)C

extern "C" {

int vilf_default_config(vilf_config* c) {
  if (!c) return VILF_ERR_INVALID;
  c->n_scan = 64; c->n_rings = 64;
  c->lidar_min = 3.0; c->lidar_max = 90.0; c->edge_threshold = 0.1;
  c->edge_leaf = 0.4; c->surf_leaf = 0.8; c->crop_half = 100.0; c->knn_gate = 1.0; c->huber = 0.1;
  c->outer_iters = 2; c->lm_max_iters = 4;
  c->max_scan_points = 300000; c->max_map_points = 1 << 20; c->max_ring_points = 0; c->flags = 0;
  c->horizon_scan = 1800; c->downsample_rate = 1; c->ri_edge_threshold = 1.0; c->ri_surf_threshold = 0.1;  // featureExtract.hpp:85-91
  return VILF_OK;
}

int vilf_create_batch(const vilf_config* cfg, int device, int count, vilf_handle** out) {
  if (!cfg || !out || count < 1 || count > 64) return VILF_ERR_INVALID;
  if (cfg->n_scan < 0 || (cfg->n_scan == 0 && (cfg->n_rings < 1 || cfg->n_rings > MAX_RINGS))) return VILF_ERR_INVALID;
  if (cfg->outer_iters < 1 || cfg->outer_iters > MAX_OUTER || cfg->lm_max_iters < 0) return VILF_ERR_INVALID;
  if (cfg->max_scan_points < 1024 || cfg->max_map_points < 1024 || !(cfg->edge_leaf > 0) || !(cfg->surf_leaf > 0) || !(cfg->knn_gate > 0)) return VILF_ERR_INVALID;
  if ((cfg->flags & VILF_FLAG_RANGE_IMAGE) && (cfg->n_scan != 0 || cfg->horizon_scan < 64 || cfg->horizon_scan > 6138 || cfg->downsample_rate < 1)) return VILF_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return VILF_ERR_CUDA;
  Ctx* C = new (std::nothrow) Ctx();
  if (!C) return VILF_ERR_CUDA;
  C->device = device; C->ucfg = *cfg; C->nlanes = count;
  int rc = build_ctx(C);
  if (rc != VILF_OK) {
    fprintf(stderr, "vilf_create: %s\n", C->err);
    destroy_ctx(C);
    return rc;
  }
  for (int i = 0; i < count; ++i) {
    vilf_handle* h = new vilf_handle{C, i};
    C->handles.push_back(h);
    out[i] = h;
  }
  return VILF_OK;
}

int vilf_create(const vilf_config* cfg, int device, vilf_handle** out) { return vilf_create_batch(cfg, device, 1, out); }

int vilf_destroy(vilf_handle* h) {
  if (!h || !h->ctx) return VILF_ERR_INVALID;
  destroy_ctx(h->ctx);
  return VILF_OK;
}

const char* vilf_last_error(const vilf_handle* h) { return (h && h->ctx) ? h->ctx->err : "invalid handle"; }

int vilf_host_alloc(void** p, uint64_t bytes) {
  // VILF_HOST_WC=1 (experiments): write-combined staging — the host only ever writes scans into these buffers, and the copy engine's
  // reads then skip the CPU cache snoop; measured against plain pinned memory in the multi-GPU end-to-end runs
  static const bool wc = getenv("VILF_HOST_WC") != nullptr;
  if (!p) return VILF_ERR_INVALID;
  if (cudaHostAlloc(p, bytes, wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault) != cudaSuccess) return VILF_ERR_CUDA;
  std::lock_guard<std::mutex> lk(g_slab_mu);
  g_slabs[(const char*)*p] = (size_t)bytes;
  return VILF_OK;
}
int vilf_memcpy_h2d_async(void* dst_dev, const void* src_host, uint64_t bytes, void* cuda_stream) {
  return cudaMemcpyAsync(dst_dev, src_host, (size_t)bytes, cudaMemcpyHostToDevice, (cudaStream_t)cuda_stream) == cudaSuccess ? VILF_OK : VILF_ERR_CUDA;
}
int vilf_pack_pointcloud2(const uint8_t* data, int n_points, int point_step, int off_x, int off_y, int off_z, int off_intensity, float* xyzi_out) {
  // sensor_msgs/PointCloud2 -> packed x, y, z, intensity (what pcl::fromROSMsg does for PointXYZI at NODE:339-340): a strided
  // gather of four little-endian float32 fields per point; off_intensity < 0 = the message has no intensity field (0 is stored).
  if (n_points < 0 || (n_points > 0 && (!data || !xyzi_out)) || point_step < 12 || off_x < 0 || off_y < 0 || off_z < 0) return VILF_ERR_INVALID;
  if (off_x + 4 > point_step || off_y + 4 > point_step || off_z + 4 > point_step || off_intensity + 4 > point_step) return VILF_ERR_INVALID;
  for (int i = 0; i < n_points; ++i) {
    const uint8_t* p = data + (size_t)i * (size_t)point_step;
    float* o = xyzi_out + (size_t)i * 4;
    memcpy(o + 0, p + off_x, 4);
    memcpy(o + 1, p + off_y, 4);
    memcpy(o + 2, p + off_z, 4);
    if (off_intensity >= 0) memcpy(o + 3, p + off_intensity, 4);
    else o[3] = 0.0f;
  }
  return VILF_OK;
}
int vilf_unpack_pointcloud2(const float* xyzi, int n_points, int point_step, int off_x, int off_y, int off_z, int off_intensity, uint8_t* data_out) {
  // packed x, y, z, intensity -> PointCloud2 bytes (pcl::toROSMsg at NODE:439-440 is a memcpy of PointXYZI[n], whose padding
  // this zero-fills): a strided scatter of four float32 fields per point.
  if (n_points < 0 || (n_points > 0 && (!xyzi || !data_out)) || point_step < 12 || off_x < 0 || off_y < 0 || off_z < 0) return VILF_ERR_INVALID;
  if (off_x + 4 > point_step || off_y + 4 > point_step || off_z + 4 > point_step || off_intensity + 4 > point_step) return VILF_ERR_INVALID;
  for (int i = 0; i < n_points; ++i) {
    uint8_t* p = data_out + (size_t)i * (size_t)point_step;
    const float* s = xyzi + (size_t)i * 4;
    memset(p, 0, (size_t)point_step);
    memcpy(p + off_x, s + 0, 4);
    memcpy(p + off_y, s + 1, 4);
    memcpy(p + off_z, s + 2, 4);
    if (off_intensity >= 0) memcpy(p + off_intensity, s + 3, 4);
  }
  return VILF_OK;
}

int vilf_node_outputs(const double rt12[12], double last[7], double relative_out[7], double path_pose_out[7]) {
  if (!rt12 || !last) return VILF_ERR_INVALID;
  // NODE:388: Eigen::Quaterniond(Matrix3d) -- the trace / largest-diagonal branches of Eigen's rotation-matrix assignment.
  const double* R = rt12;  // row-major
  double q[4];             // x y z w
  double t = R[0] + R[4] + R[8];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t;
    q[1] = (R[2] - R[6]) * t;
    q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R[3 * k + j] - R[3 * j + k]) * t;
    q[j] = (R[3 * j + i] + R[3 * i + j]) * t;
    q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
  }
  const double tt[3] = {rt12[9], rt12[10], rt12[11]};
  if (path_pose_out) {
    for (int c = 0; c < 4; ++c) path_pose_out[c] = q[c];
    for (int c = 0; c < 3; ++c) path_pose_out[4 + c] = tt[c];
  }
  if (relative_out) {
    // NODE:400-401.  Quaternion::inverse() = conjugate / squaredNorm; product and vector rotation as Eigen evaluates them.
    const double n2 = last[0] * last[0] + last[1] * last[1] + last[2] * last[2] + last[3] * last[3];
    double a[4] = {0, 0, 0, 0};  // inverse of q_last (Eigen returns the zero quaternion for a zero input)
    if (n2 > 0.0) { a[0] = -last[0] / n2; a[1] = -last[1] / n2; a[2] = -last[2] / n2; a[3] = last[3] / n2; }
    relative_out[0] = a[3] * q[0] + a[0] * q[3] + a[1] * q[2] - a[2] * q[1];
    relative_out[1] = a[3] * q[1] + a[1] * q[3] + a[2] * q[0] - a[0] * q[2];
    relative_out[2] = a[3] * q[2] + a[2] * q[3] + a[0] * q[1] - a[1] * q[0];
    relative_out[3] = a[3] * q[3] - a[0] * q[0] - a[1] * q[1] - a[2] * q[2];
    const double v[3] = {tt[0] - last[4], tt[1] - last[5], tt[2] - last[6]};
    double uv[3] = {a[1] * v[2] - a[2] * v[1], a[2] * v[0] - a[0] * v[2], a[0] * v[1] - a[1] * v[0]};
    for (int c = 0; c < 3; ++c) uv[c] += uv[c];
    const double w[3] = {a[1] * uv[2] - a[2] * uv[1], a[2] * uv[0] - a[0] * uv[2], a[0] * uv[1] - a[1] * uv[0]};
    for (int c = 0; c < 3; ++c) relative_out[4 + c] = v[c] + a[3] * uv[c] + w[c];
  }
  for (int c = 0; c < 4; ++c) last[c] = q[c];  // NODE:445-446
  for (int c = 0; c < 3; ++c) last[4 + c] = tt[c];
  return VILF_OK;
}
int vilf_host_free(void* p) {
  {
    std::lock_guard<std::mutex> lk(g_slab_mu);
    g_slabs.erase((const char*)p);
  }
  return cudaFreeHost(p) == cudaSuccess ? VILF_OK : VILF_ERR_CUDA;
}

int vilf_submit_scan(vilf_handle* h, const float* xyzi, int n, const uint16_t* ring, int64_t* ticket) {
  HCHECK(h);
  if (!ticket) return VILF_ERR_INVALID;
  return submit_common(C, h->lane, 1, &xyzi, &n, ring ? &ring : nullptr, false, ticket);
}
int vilf_wait(vilf_handle* h, int64_t ticket, double pose_out[7]) {
  HCHECK(h);
  return wait_common(C, h->lane, 1, ticket, pose_out);
}
int vilf_process_scan(vilf_handle* h, const float* xyzi, int n, const uint16_t* ring, double pose_out[7]) {
  int64_t t = 0;
  int rc = vilf_submit_scan(h, xyzi, n, ring, &t);
  if (rc) return rc;
  return vilf_wait(h, t, pose_out);
}

static int batch_check(vilf_handle* const* hs, int count) {
  if (!hs || count < 1 || !hs[0] || !hs[0]->ctx) return VILF_ERR_INVALID;
  for (int i = 0; i < count; ++i)
    if (!hs[i] || hs[i]->ctx != hs[0]->ctx || hs[i]->lane != hs[0]->lane + i) return VILF_ERR_INVALID;
  return VILF_OK;
}
int vilf_submit_scan_batch(vilf_handle* const* hs, int count, const float* const* xyzi, const int* n, const uint16_t* const* ring, int64_t* ticket) {
  if (batch_check(hs, count) || !xyzi || !n || !ticket) return VILF_ERR_INVALID;
  return submit_common(hs[0]->ctx, hs[0]->lane, count, xyzi, n, ring, false, ticket);
}
int vilf_submit_scan_batch_dev(vilf_handle* const* hs, int count, const float* const* xyzi_dev, const int* n, const uint16_t* const* ring_dev,
                               int64_t* ticket) {
  if (batch_check(hs, count) || !xyzi_dev || !n || !ticket) return VILF_ERR_INVALID;
  return submit_common(hs[0]->ctx, hs[0]->lane, count, xyzi_dev, n, ring_dev, true, ticket);
}
int vilf_wait_batch(vilf_handle* const* hs, int count, int64_t ticket, double* poses_out) {
  if (batch_check(hs, count)) return VILF_ERR_INVALID;
  return wait_common(hs[0]->ctx, hs[0]->lane, count, ticket, poses_out);
}

int vilf_feature_extract(vilf_handle* h, const float* xyzi, int n, const uint16_t* ring, int* n_edge, int* n_surf) {
  HCHECK(h);
  CK(cudaSetDevice(C->device));
  if (n < 0 || (n > 0 && !xyzi)) return fail(C, VILF_ERR_INVALID, "bad scan");
  if (n > C->cfg.cap_scan) return fail(C, VILF_ERR_CAPACITY, "scan exceeds max_scan_points");
  if (C->cfg.n_scan == 0 && !ring) return fail(C, VILF_ERR_INVALID, "n_scan == 0 needs explicit ring ids");
  LaneDev& L = C->lanes_host[h->lane];
  CK(cudaStreamSynchronize(C->st));
  CK(cudaStreamSynchronize(C->copy_st));
  if (n) CK(cudaMemcpyAsync(L.scan[0], xyzi, (size_t)n * 16, cudaMemcpyHostToDevice, C->st));
  if (ring && n) CK(cudaMemcpyAsync(L.ring_in[0], ring, (size_t)n * 2, cudaMemcpyHostToDevice, C->st));
  CK(cudaMemcpyAsync(&L.v->n_scan[0], &n, sizeof(int), cudaMemcpyHostToDevice, C->st));
  CK(cudaMemsetAsync(&L.v->status, 0, sizeof(int), C->st));
  CK(cudaStreamSynchronize(C->st));
  stage1(C, mk(C), h->lane, 1, 0);
  CK(cudaGetLastError());
  LaneVars V;
  int rc = read_vars(C, h->lane, &V);
  if (rc) return rc;
  C->have_feat[h->lane] = 1;
  C->last_sel[h->lane] = 0;
  if (n_edge) *n_edge = V.n_edge;
  if (n_surf) *n_surf = V.n_surf;
  return status_to_rc(C, V.status);
}

int vilf_get_features(vilf_handle* h, int which, float* pts, int32_t* src, int cap, int* n) {
  HCHECK(h);
  CK(cudaSetDevice(C->device));
  if (which < 0 || which > 1 || !n) return VILF_ERR_INVALID;
  LaneVars V;
  int rc = read_vars(C, h->lane, &V);
  if (rc) return rc;
  const int cnt = which ? V.n_surf : V.n_edge;
  *n = cnt;
  if (cnt > cap && (pts || src)) return fail(C, VILF_ERR_CAPACITY, "output buffer too small");
  LaneDev& L = C->lanes_host[h->lane];
  if (pts && cnt) CK(cudaMemcpy(pts, L.feat[which], (size_t)cnt * 16, cudaMemcpyDeviceToHost));
  if (src && cnt) CK(cudaMemcpy(src, L.feat_src[which], (size_t)cnt * 4, cudaMemcpyDeviceToHost));
  return VILF_OK;
}

static int map_init_impl(Ctx* C, int lane) {
  const Launch L = mk(C);
  if (C->cellmap) {
    launch_cell_build(L, C->build_dev[C->cur[lane]] + lane * 2, C->build_sort_dev + lane * 2, 2);
    launch_map_init_commit(L, C->lanes_dev, lane, 1, C->cfg);
  } else {
    launch_map_init(L, C->lanes_dev, lane, 1, C->cur[lane], C->cfg);
    build_grids(C, L, C->grid_dev[C->cur[lane]] + lane * 2, 2, C->cluster_map);
  }
  CK(cudaGetLastError());
  C->have_map[lane] = 1; C->last_init[lane] = 1;
  return finish_sync(C, lane, nullptr);
}
int vilf_map_init(vilf_handle* h) {
  HCHECK(h);
  CK(cudaSetDevice(C->device));
  if (!C->have_feat[h->lane]) return fail(C, VILF_ERR_STATE, "no extracted features resident");
  return map_init_impl(C, h->lane);
}
int vilf_map_init_points(vilf_handle* h, const float* edge, int n_edge, const float* surf, int n_surf) {
  HCHECK(h);
  CK(cudaSetDevice(C->device));
  int rc = upload_features(C, h->lane, edge, n_edge, surf, n_surf);
  if (rc) return rc;
  C->have_feat[h->lane] = 1;
  return map_init_impl(C, h->lane);
}

static int update_impl(Ctx* C, int lane, double* pose_out) {
  if (!C->have_map[lane]) return fail(C, VILF_ERR_STATE, "local map not initialised");
  int rc = enqueue_frame(C, lane, 1, false, false, 0, nullptr);
  if (rc) return rc;
  return finish_sync(C, lane, pose_out);
}
int vilf_update(vilf_handle* h, double pose_out[7]) {
  HCHECK(h);
  CK(cudaSetDevice(C->device));
  if (!C->have_feat[h->lane]) return fail(C, VILF_ERR_STATE, "no extracted features resident");
  return update_impl(C, h->lane, pose_out);
}
int vilf_update_points(vilf_handle* h, const float* edge, int n_edge, const float* surf, int n_surf, double pose_out[7]) {
  HCHECK(h);
  CK(cudaSetDevice(C->device));
  int rc = upload_features(C, h->lane, edge, n_edge, surf, n_surf);
  if (rc) return rc;
  C->have_feat[h->lane] = 1;
  return update_impl(C, h->lane, pose_out);
}

int vilf_set_pose(vilf_handle* h, const double pose[7], int update_odom) {
  HCHECK(h);
  CK(cudaSetDevice(C->device));
  if (!pose) return VILF_ERR_INVALID;
  LaneVars V;
  int rc = read_vars(C, h->lane, &V);
  if (rc) return rc;
  memcpy(V.x, pose, 7 * sizeof(double));
  if (update_odom) {  // EM:291-293: globalOdom.linear() = q_w_c.toRotationMatrix(); translation = t_w_c (Eigen operation order)
    const double* x = V.x;
    const double tx = 2 * x[0], ty = 2 * x[1], tz = 2 * x[2];
    const double twx = tx * x[3], twy = ty * x[3], twz = tz * x[3];
    const double txx = tx * x[0], txy = ty * x[0], txz = tz * x[0];
    const double tyy = ty * x[1], tyz = tz * x[1], tzz = tz * x[2];
    double* o = V.odom;
    o[0] = 1 - (tyy + tzz); o[1] = txy - twz; o[2] = txz + twy;
    o[3] = txy + twz; o[4] = 1 - (txx + tzz); o[5] = tyz - twx;
    o[6] = txz - twy; o[7] = tyz + twx; o[8] = 1 - (txx + tyy);
    o[9] = x[4]; o[10] = x[5]; o[11] = x[6];
  }
  CK(cudaMemcpyAsync(C->vars_dev + h->lane, &V, sizeof(V), cudaMemcpyHostToDevice, C->st));
  CK(cudaStreamSynchronize(C->st));
  return VILF_OK;
}

int vilf_predict(vilf_handle* h, double pose_out[7]) {
  HCHECK(h);
  CK(cudaSetDevice(C->device));
  launch_frame_reset(mk(C), C->lanes_dev, h->lane, 1, C->vv_dev, VV_PER_LANE, 1);
  CK(cudaGetLastError());
  return finish_sync(C, h->lane, pose_out);
}

int vilf_create_submap(vilf_handle* h, const float* edge_ds, int n_edge, const float* surf_ds, int n_surf) {
  HCHECK(h);
  CK(cudaSetDevice(C->device));
  const int lane = h->lane;
  if (n_edge < 0 || n_surf < 0 || (n_edge && !edge_ds) || (n_surf && !surf_ds)) return fail(C, VILF_ERR_INVALID, "bad feature clouds");
  if (n_edge > C->cfg.cap_scan || n_surf > C->cfg.cap_scan) return fail(C, VILF_ERR_CAPACITY, "feature cloud exceeds max_scan_points");
  LaneDev& L = C->lanes_host[lane];
  if (n_edge) CK(cudaMemcpyAsync(L.ds[0], edge_ds, (size_t)n_edge * 16, cudaMemcpyHostToDevice, C->st));
  if (n_surf) CK(cudaMemcpyAsync(L.ds[1], surf_ds, (size_t)n_surf * 16, cudaMemcpyHostToDevice, C->st));
  int cnt[2] = {n_edge, n_surf};
  CK(cudaMemcpyAsync(&L.v->n_ds[0], cnt, sizeof(cnt), cudaMemcpyHostToDevice, C->st));  // n_ds[0], n_ds[1] are adjacent
  CK(cudaStreamSynchronize(C->st));
  const Launch Ln = mk(C);
  launch_frame_reset(Ln, C->lanes_dev, lane, 1, C->vv_dev, VV_PER_LANE, 0);  // bounding boxes / status of this map update
  enqueue_submap(C, Ln, lane, 1, nullptr);
  CK(cudaGetLastError());
  C->have_map[lane] = 1; C->last_init[lane] = 0;
  return finish_sync(C, lane, nullptr);
}

int vilf_get_pose(vilf_handle* h, double pose_out[7], double* rt12) {
  HCHECK(h);
  CK(cudaSetDevice(C->device));
  LaneVars V;
  int rc = read_vars(C, h->lane, &V);
  if (rc) return rc;
  if (pose_out) memcpy(pose_out, V.x, 7 * sizeof(double));
  if (rt12) memcpy(rt12, V.odom, 12 * sizeof(double));
  return VILF_OK;
}

int vilf_get_cloud(vilf_handle* h, int which, float* out, int cap, int* n) {
  HCHECK(h);
  CK(cudaSetDevice(C->device));
  if (which < 0 || which > 5 || !n) return VILF_ERR_INVALID;
  LaneVars V;
  int rc = read_vars(C, h->lane, &V);
  if (rc) return rc;
  LaneDev& L = C->lanes_host[h->lane];
  const int cur = C->cur[h->lane];
  const float4* part[2] = {nullptr, nullptr};
  int cnt[2] = {0, 0};
  if (which <= 1 && C->cellmap) {  // stored in cell order: hand it out in the reference's order
    rc = map_to_pcl(C, h->lane, which, &cnt[0], nullptr);
    if (rc) return rc;
    part[0] = C->aux_out;
  } else if (which == 4 && C->cellmap && !C->last_init[h->lane]) {  // EM:315, :323: the transformed features, behind the orphans
    MergeVars mv[2];
    CK(cudaMemcpy(mv, C->mv_dev + h->lane * 2, sizeof(mv), cudaMemcpyDeviceToHost));
    for (int w = 0; w < 2; ++w) { part[w] = C->merge_host[0][h->lane * 2 + w].newpts + mv[w].n_orph_in; cnt[w] = mv[w].n_in - mv[w].n_orph_in; }
  } else if (which <= 1) { part[0] = L.map[which][cur]; cnt[0] = V.n_map[which]; }
  else if (which <= 3) { part[0] = L.ds[which - 2]; cnt[0] = V.n_ds[which - 2]; }
  else if (C->last_init[h->lane]) {  // EM:110-114
    part[0] = L.feat[0]; cnt[0] = V.n_edge; part[1] = L.feat[1]; cnt[1] = V.n_surf;
  } else if (which == 4) {  // EM:315, :323: the world-frame features appended to the previous buffer
    for (int w = 0; w < 2; ++w) { part[w] = L.map[w][cur ^ 1] + (V.n_cat[w] - V.n_ds[w]); cnt[w] = V.n_ds[w]; }
  } else {                  // EM:304-305
    for (int w = 0; w < 2; ++w) { part[w] = L.ds[w]; cnt[w] = V.n_ds[w]; }
  }
  *n = cnt[0] + cnt[1];
  if (!out) return VILF_OK;
  if (*n > cap) return fail(C, VILF_ERR_CAPACITY, "output buffer too small");
  if (cnt[0]) CK(cudaMemcpy(out, part[0], (size_t)cnt[0] * 16, cudaMemcpyDeviceToHost));
  if (cnt[1]) CK(cudaMemcpy(out + (size_t)cnt[0] * 4, part[1], (size_t)cnt[1] * 16, cudaMemcpyDeviceToHost));
  return VILF_OK;
}

int vilf_voxel_downsample(vilf_handle* h, const float* pts, int n, float leaf, float* out, int cap, int* n_out, int* guard) {
  HCHECK(h);
  if (!(leaf > 0)) return VILF_ERR_INVALID;
  return run_aux_voxel(C, pts, n, leaf, 0, nullptr, 0, nullptr, nullptr, 0, out, cap, n_out, guard);
}
int vilf_crop_voxel_downsample(vilf_handle* h, const float* pts, int n, const double center[3], double half, float leaf, float* out, int cap, int* n_out) {
  HCHECK(h);
  if (!(leaf > 0) || !center) return VILF_ERR_INVALID;
  return run_aux_voxel(C, pts, n, leaf, 1, center, half, nullptr, nullptr, 0, out, cap, n_out, nullptr);
}
int vilf_crop_box(vilf_handle* h, const float* pts, int n, const double mn[3], const double mx[3], float* out, int cap, int* n_out) {
  HCHECK(h);
  if (!mn || !mx) return VILF_ERR_INVALID;
  return run_aux_voxel(C, pts, n, 1.0f, 2, nullptr, 0, mn, mx, 1, out, cap, n_out, nullptr);
}

int vilf_knn5(vilf_handle* h, const float* map, int m, const float* q, int nq, int32_t* idx, float* d2) {
  HCHECK(h);
  CK(cudaSetDevice(C->device));
  if (m < 0 || nq < 0 || (m > 0 && !map) || (nq > 0 && (!q || !idx || !d2))) return fail(C, VILF_ERR_INVALID, "bad arguments");
  if (m > C->cap_aux || nq > C->cap_aux) return fail(C, VILF_ERR_CAPACITY, "map or query set exceeds capacity");
  if (m) CK(cudaMemcpyAsync(C->aux_in, map, (size_t)m * 16, cudaMemcpyHostToDevice, C->st));
  if (nq) CK(cudaMemcpyAsync(C->aux_out, q, (size_t)nq * 16, cudaMemcpyHostToDevice, C->st));
  int hdr[4] = {m, 0, nq, 0};
  CK(cudaMemcpyAsync(C->aux_n, hdr, sizeof(hdr), cudaMemcpyHostToDevice, C->st));
  CK(cudaStreamSynchronize(C->st));
  const Launch L = mk(C);
  if (C->cellmap) {
    int rc = run_cell_build(C, C->aux_build_host);
    if (rc) return rc;
    launch_knn_cell_only(L, C->aux_cm_pts[0], C->aux_n, C->aux_ctab, C->aux_cmeta, C->aux_corig, C->aux_build_host.g, C->aux_out, C->aux_n + 2, C->aux_idx,
                         C->aux_d2, C->cfg);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(hdr, C->aux_n, sizeof(hdr), cudaMemcpyDeviceToHost, C->st));
    CK(cudaStreamSynchronize(C->st));
    if (hdr[3]) return status_to_rc(C, hdr[3]);
  } else {
    build_grids(C, L, C->aux_grid_dev, 1, !(C->ucfg.flags & VILF_FLAG_NO_CLUSTER) && m <= CLUSTER_MAX_POINTS);
    launch_knn_only(L, C->aux_grid_dev, C->aux_out, C->aux_n + 2, C->aux_idx, C->aux_d2, C->cfg);
  }
  CK(cudaGetLastError());
  if (nq) {
    CK(cudaMemcpyAsync(idx, C->aux_idx, (size_t)nq * 5 * 4, cudaMemcpyDeviceToHost, C->st));
    CK(cudaMemcpyAsync(d2, C->aux_d2, (size_t)nq * 5 * 4, cudaMemcpyDeviceToHost, C->st));
  }
  CK(cudaStreamSynchronize(C->st));
  return VILF_OK;
}

int vilf_bench_stage(vilf_handle* h, int stage, const float* map, int m, const float* q, int nq, float leaf, int iters, double ms_out[4]) {
  HCHECK(h);
  CK(cudaSetDevice(C->device));
  if (!map || m < 1 || iters < 1 || !ms_out || stage < 0 || stage > 2 || (stage != 1 && (!q || nq < 1)) || (stage >= 1 && !(leaf > 0)))
    return fail(C, VILF_ERR_INVALID, "bad arguments");
  if (m > C->cap_aux || nq > C->cap_aux || (stage == 2 && m + nq > C->cap_aux)) return fail(C, VILF_ERR_CAPACITY, "map or query set exceeds capacity");
  const bool small = !(C->ucfg.flags & VILF_FLAG_NO_CLUSTER) && m <= CLUSTER_MAX_POINTS;
  cudaEvent_t ev[3];
  for (int i = 0; i < 3; ++i) CK(cudaEventCreate(&ev[i]));
  CK(cudaMemcpyAsync(C->aux_in, map, (size_t)m * 16, cudaMemcpyHostToDevice, C->st));
  int hdr[8] = {m, 0, nq, 0, 0, 0, 0, 0};
  CK(cudaMemcpyAsync(C->aux_n, hdr, sizeof(hdr), cudaMemcpyHostToDevice, C->st));
  double acc[2] = {0, 0};
  const Launch L = mk(C);
  if (stage == 0 && C->cellmap) {
    CellBuildJob B = C->aux_build_host;
    if (leaf > 0) B.g = cell_geometry(C, (double)leaf);  // search cells sized for a map filtered at `leaf`
    CK(cudaMemcpyAsync(C->aux_build_dev, &B, sizeof(B), cudaMemcpyHostToDevice, C->st));
    CK(cudaMemcpyAsync(C->aux_build_sort_dev, &B.sort, sizeof(SortJob), cudaMemcpyHostToDevice, C->st));
    CK(cudaMemcpyAsync(C->aux_out, q, (size_t)nq * 16, cudaMemcpyHostToDevice, C->st));
    CK(cudaStreamSynchronize(C->st));
    for (int it = -1; it < iters; ++it) {  // iteration -1 warms up
      CK(cudaEventRecord(ev[0], C->st));
      launch_cell_build(L, C->aux_build_dev, C->aux_build_sort_dev, 1);
      CK(cudaEventRecord(ev[1], C->st));
      launch_knn_cell_only(L, C->aux_cm_pts[0], C->aux_n, C->aux_ctab, C->aux_cmeta, C->aux_corig, B.g, C->aux_out, C->aux_n + 2, C->aux_idx, C->aux_d2, C->cfg);
      CK(cudaEventRecord(ev[2], C->st));
      CK(cudaStreamSynchronize(C->st));
      CK(cudaGetLastError());
      float a = 0, b = 0;
      CK(cudaEventElapsedTime(&a, ev[0], ev[1]));
      CK(cudaEventElapsedTime(&b, ev[1], ev[2]));
      if (it >= 0) { acc[0] += a; acc[1] += b; }
    }
    ms_out[0] = acc[0] / iters; ms_out[1] = acc[1] / iters; ms_out[2] = (double)B.g.shells; ms_out[3] = (double)B.g.leaf * (double)(1 << B.g.shift);
  } else if (stage == 0) {
    GridJob Gb = C->aux_grid_host;
    if (leaf > 0) grid_geometry(C, (double)leaf, Gb);  // search-grid cell sized for a map filtered at `leaf`
    CK(cudaMemcpyAsync(C->aux_grid_dev, &Gb, sizeof(Gb), cudaMemcpyHostToDevice, C->st));
    CK(cudaStreamSynchronize(C->st));
    CK(cudaMemcpyAsync(C->aux_out, q, (size_t)nq * 16, cudaMemcpyHostToDevice, C->st));
    for (int it = -1; it < iters; ++it) {  // iteration -1 warms up
      CK(cudaEventRecord(ev[0], C->st));
      build_grids(C, L, C->aux_grid_dev, 1, small);
      CK(cudaEventRecord(ev[1], C->st));
      launch_knn_only(L, C->aux_grid_dev, C->aux_out, C->aux_n + 2, C->aux_idx, C->aux_d2, C->cfg);
      CK(cudaEventRecord(ev[2], C->st));
      CK(cudaStreamSynchronize(C->st));
      CK(cudaGetLastError());
      float a = 0, b = 0;
      CK(cudaEventElapsedTime(&a, ev[0], ev[1]));
      CK(cudaEventElapsedTime(&b, ev[1], ev[2]));
      if (it >= 0) { acc[0] += a; acc[1] += b; }
    }
    ms_out[0] = acc[0] / iters; ms_out[1] = acc[1] / iters; ms_out[2] = (double)Gb.rings; ms_out[3] = 1.0 / (double)Gb.inv_cell;
    CK(cudaMemcpy(C->aux_grid_dev, &C->aux_grid_host, sizeof(GridJob), cudaMemcpyHostToDevice));
  } else if (stage == 1 || !C->cellmap) {
    // stage 1: crop box + voxel filter of an UNSORTED cloud (radix path).  Legacy stage 2: the per-frame map maintenance of the
    // radix path = the same filter over [voxel-filtered map ..., nq appended points].
    VoxJob J = C->aux_vox_host;
    J.leaf = leaf; J.crop = 2; J.passthrough = 0;
    for (int a = 0; a < 3; ++a) { J.crop_lo[a] = -100.0f; J.crop_hi[a] = 100.0f; }  // EM:327-336 about the origin
    CK(cudaMemcpyAsync(C->aux_vox_dev, &J, sizeof(J), cudaMemcpyHostToDevice, C->st));
    VoxVars vv;
    memset(&vv, 0, sizeof(vv));
    vv.bbox[0] = vv.bbox[1] = vv.bbox[2] = INT_MAX;
    vv.bbox[3] = vv.bbox[4] = vv.bbox[5] = INT_MIN;
    int n_out = 0, n_before = m;
    if (stage == 2) {  // filter once (untimed), then time the filter of [filtered map, new points]
      CK(cudaMemcpyAsync(J.vv, &vv, sizeof(vv), cudaMemcpyHostToDevice, C->st));
      if (small) launch_voxel_cluster(L, C->aux_vox_dev, 1, false, C->cfg);
      else launch_voxel(L, C->aux_vox_dev, 1, C->aux_sort_dev, false);
      CK(cudaMemcpyAsync(hdr, C->aux_n, sizeof(hdr), cudaMemcpyDeviceToHost, C->st));
      CK(cudaStreamSynchronize(C->st));
      CK(cudaGetLastError());
      n_before = hdr[1];
      CK(cudaMemcpyAsync(C->aux_in, C->aux_out, (size_t)n_before * 16, cudaMemcpyDeviceToDevice, C->st));
      CK(cudaMemcpyAsync(C->aux_in + n_before, q, (size_t)nq * 16, cudaMemcpyHostToDevice, C->st));
      hdr[0] = n_before + nq; hdr[1] = 0;
      CK(cudaMemcpyAsync(C->aux_n, hdr, 16, cudaMemcpyHostToDevice, C->st));
      CK(cudaStreamSynchronize(C->st));
    }
    const bool small2 = !(C->ucfg.flags & VILF_FLAG_NO_CLUSTER) && hdr[0] <= CLUSTER_MAX_POINTS;
    for (int it = -1; it < iters; ++it) {
      CK(cudaMemcpyAsync(J.vv, &vv, sizeof(vv), cudaMemcpyHostToDevice, C->st));
      CK(cudaEventRecord(ev[0], C->st));
      if (small2) launch_voxel_cluster(L, C->aux_vox_dev, 1, false, C->cfg);
      else launch_voxel(L, C->aux_vox_dev, 1, C->aux_sort_dev, false);
      CK(cudaEventRecord(ev[1], C->st));
      CK(cudaMemcpyAsync(hdr, C->aux_n, 16, cudaMemcpyDeviceToHost, C->st));
      CK(cudaStreamSynchronize(C->st));
      CK(cudaGetLastError());
      float a = 0;
      CK(cudaEventElapsedTime(&a, ev[0], ev[1]));
      if (it >= 0) acc[0] += a;
      n_out = hdr[1];
    }
    ms_out[0] = acc[0] / iters; ms_out[1] = 0; ms_out[2] = stage == 2 ? (double)n_before : (double)n_out; ms_out[3] = (double)n_out;
  } else {
    // stage 2, cell-ordered map: what createSubMap costs per frame — a voxel-filtered map in cell order (built untimed from the
    // cloud by one update of an empty map) merged with nq new points: crop box + voxel filter + cell table in one update.
    MergeJob A = C->aux_merge_host;
    A.g = cell_geometry(C, (double)leaf);
    A.old_pts = C->aux_cm_pts[1]; A.out_pts = C->aux_cm_pts[0]; A.n_map = C->aux_n + 5;
    A.src = C->aux_in; A.n_src = C->aux_n;
    A.crop_center = C->aux_pose; A.crop_half = 100.0;  // EM:327-336 about the origin
    double ctr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    CK(cudaMemcpyAsync(C->aux_pose, ctr, sizeof(ctr), cudaMemcpyHostToDevice, C->st));
    MergeVars mz;
    memset(&mz, 0, sizeof(mz));
    CK(cudaMemcpyAsync(A.mv, &mz, sizeof(mz), cudaMemcpyHostToDevice, C->st));
    CK(cudaMemcpyAsync(C->aux_merge_dev, &A, sizeof(A), cudaMemcpyHostToDevice, C->st));
    CK(cudaMemcpyAsync(C->aux_merge_sort_dev, &A.sort, sizeof(SortJob), cudaMemcpyHostToDevice, C->st));
    CK(cudaMemcpyAsync(C->aux_out, q, (size_t)nq * 16, cudaMemcpyHostToDevice, C->st));
    CK(cudaStreamSynchronize(C->st));
    CK(cudaEventRecord(ev[0], C->st));
    launch_cell_update(L, C->aux_merge_dev, C->aux_merge_sort_dev, 1, C->aux_max_tiles, false);
    CK(cudaEventRecord(ev[1], C->st));
    CK(cudaMemcpyAsync(hdr, C->aux_n, sizeof(hdr), cudaMemcpyDeviceToHost, C->st));
    CK(cudaStreamSynchronize(C->st));
    CK(cudaGetLastError());
    float prep = 0;
    CK(cudaEventElapsedTime(&prep, ev[0], ev[1]));
    const int n_before = hdr[5];
    if (hdr[3]) { for (int i = 0; i < 3; ++i) cudaEventDestroy(ev[i]); return status_to_rc(C, hdr[3]); }
    MergeJob Bj = A;
    Bj.old_pts = C->aux_cm_pts[0]; Bj.out_pts = C->aux_cm_pts[1]; Bj.n_map = C->aux_n + 6;
    Bj.src = C->aux_out; Bj.n_src = C->aux_n + 2;
    CK(cudaMemcpyAsync(C->aux_merge_dev, &Bj, sizeof(Bj), cudaMemcpyHostToDevice, C->st));
    CK(cudaStreamSynchronize(C->st));
    int n_out = 0;
    for (int it = -1; it < iters; ++it) {
      CK(cudaMemcpyAsync(C->aux_n + 6, C->aux_n + 5, sizeof(int), cudaMemcpyDeviceToDevice, C->st));
      CK(cudaEventRecord(ev[0], C->st));
      launch_cell_update(L, C->aux_merge_dev, C->aux_merge_sort_dev, 1, C->aux_max_tiles,
                         !C->cfg.flags_no_cluster && nq + ORPHAN_CAP <= CLUSTER_MAX_POINTS);  // like the per-frame path: one cluster sorts the new points
      CK(cudaEventRecord(ev[1], C->st));
      CK(cudaMemcpyAsync(hdr, C->aux_n, sizeof(hdr), cudaMemcpyDeviceToHost, C->st));
      CK(cudaStreamSynchronize(C->st));
      CK(cudaGetLastError());
      float a = 0;
      CK(cudaEventElapsedTime(&a, ev[0], ev[1]));
      if (it >= 0) acc[0] += a;
      n_out = hdr[6];
    }
    ms_out[0] = acc[0] / iters; ms_out[1] = (double)prep; ms_out[2] = (double)n_before; ms_out[3] = (double)n_out;
  }
  for (int i = 0; i < 3; ++i) cudaEventDestroy(ev[i]);
  return VILF_OK;
}

int vilf_feature_depth(vilf_handle* h, const float* cloud_cam, int n, const double T_lidar_cam[16], const float* feats, int m, int num_bins,
                       float* depth_out, int32_t* nn_out, int* n_cloud) {
  HCHECK(h);
  CK(cudaSetDevice(C->device));
  const int lane = h->lane;
  if (m < 0 || m > DEPTH_MAX_FEATURES || (m > 0 && (!feats || !depth_out)) || num_bins < 1) return fail(C, VILF_ERR_INVALID, "bad arguments");
  const float4* in = nullptr;
  const int* n_dev = nullptr;
  int from_scan = 0;
  if (cloud_cam) {  // steps 4.1-4.4 on an explicit camera-frame cloud
    if (n < 0 || n > C->cap_aux) return fail(C, VILF_ERR_CAPACITY, "cloud exceeds capacity");
    if (n) CK(cudaMemcpyAsync(C->aux_in, cloud_cam, (size_t)n * 16, cudaMemcpyHostToDevice, C->st));
    CK(cudaMemcpyAsync(C->aux_n, &n, sizeof(int), cudaMemcpyHostToDevice, C->st));
    in = C->aux_in; n_dev = C->aux_n;
  } else {          // NODE:348-361 on the scan that is resident from the last extractFeature / process_scan, then 4.1-4.4
    if (!T_lidar_cam) return fail(C, VILF_ERR_INVALID, "no cloud and no extrinsic");
    if (C->last_sel[lane] < 0) return fail(C, VILF_ERR_STATE, "no scan resident");
    const int sel = C->last_sel[lane];
    CK(cudaMemcpyAsync(C->depth_T, T_lidar_cam, 16 * sizeof(double), cudaMemcpyHostToDevice, C->st));
    in = C->lanes_host[lane].scan[sel]; n_dev = &C->lanes_host[lane].v->n_scan[sel];
    from_scan = 1;
  }
  if (m) CK(cudaMemcpyAsync(C->depth_feat, feats, (size_t)m * 12, cudaMemcpyHostToDevice, C->st));
  const float bin_res = 180.0f / (float)num_bins;                                   // NODE:78
  const float thr = (float)pow(sin(bin_res / 180.0 * M_PI) * 5.0, 2);               // NODE:103
  CK(cudaStreamSynchronize(C->st));  // staged host values above live on this stack frame
  launch_depth(mk(C), in, n_dev, from_scan, C->depth_T, C->aux_out, C->aux_idx, C->aux_n + 4, C->depth_feat, m, thr, C->depth_out, C->depth_nn);
  CK(cudaGetLastError());
  int cnt = 0;
  if (m) CK(cudaMemcpyAsync(depth_out, C->depth_out, (size_t)m * 4, cudaMemcpyDeviceToHost, C->st));
  if (m && nn_out) CK(cudaMemcpyAsync(nn_out, C->depth_nn, (size_t)m * 12, cudaMemcpyDeviceToHost, C->st));
  CK(cudaMemcpyAsync(&cnt, C->aux_n + 4, sizeof(int), cudaMemcpyDeviceToHost, C->st));
  CK(cudaStreamSynchronize(C->st));
  if (n_cloud) *n_cloud = cnt;
  return VILF_OK;
}

int vilf_factors(vilf_handle* h, const double pose[7], const float* edge, int n_edge, const float* surf, int n_surf, uint8_t* edge_valid,
                 double* edge_ab, int32_t* edge_nn, float* edge_d2, uint8_t* surf_valid, double* surf_nd, int32_t* surf_nn, float* surf_d2) {
  HCHECK(h);
  CK(cudaSetDevice(C->device));
  const int lane = h->lane;
  if (!pose || n_edge < 0 || n_surf < 0) return VILF_ERR_INVALID;
  if (n_edge > C->cfg.cap_scan || n_surf > C->cfg.cap_scan) return fail(C, VILF_ERR_CAPACITY, "feature cloud exceeds max_scan_points");
  if (!C->have_map[lane]) return fail(C, VILF_ERR_STATE, "local map not initialised");
  LaneDev& L = C->lanes_host[lane];
  LaneVars V;
  int rc = read_vars(C, lane, &V);
  if (rc) return rc;
  LaneVars W = V;
  W.n_ds[0] = n_edge; W.n_ds[1] = n_surf;
  CK(cudaMemcpyAsync(C->vars_dev + lane, &W, sizeof(W), cudaMemcpyHostToDevice, C->st));
  if (n_edge) CK(cudaMemcpyAsync(L.ds[0], edge, (size_t)n_edge * 16, cudaMemcpyHostToDevice, C->st));
  if (n_surf) CK(cudaMemcpyAsync(L.ds[1], surf, (size_t)n_surf * 16, cudaMemcpyHostToDevice, C->st));
  CK(cudaMemcpyAsync(C->aux_pose, pose, 7 * sizeof(double), cudaMemcpyHostToDevice, C->st));
  for (int w = 0; w < 2; ++w) {
    const int n = w ? n_surf : n_edge;
    if (!n) continue;
    CK(cudaMemsetAsync(L.fvalid[w], 0, (size_t)n, C->st));
    CK(cudaMemsetAsync(L.nn_idx[w], 0xff, (size_t)n * 20, C->st));
    CK(cudaMemsetAsync(L.nn_d2[w], 0, (size_t)n * 20, C->st));
  }
  CK(cudaMemsetAsync(L.edge_pab, 0, (size_t)(n_edge > 0 ? n_edge : 1) * 72, C->st));
  CK(cudaMemsetAsync(L.surf_pnd, 0, (size_t)(n_surf > 0 ? n_surf : 1) * 56, C->st));
  if (C->cellmap) launch_knn_cell_fit(mk(C), C->lanes_dev, lane, 1, C->cur[lane], C->cfg, C->aux_pose);
  else launch_knn_fit(mk(C), C->lanes_dev, C->grid_dev[C->cur[lane]], lane, 1, C->cur[lane], C->cfg, C->aux_pose);
  CK(cudaGetLastError());
  std::vector<double> pab((size_t)(n_edge > 0 ? n_edge : 1) * 9), pnd((size_t)(n_surf > 0 ? n_surf : 1) * 7);
  if (n_edge) {
    CK(cudaMemcpyAsync(pab.data(), L.edge_pab, (size_t)n_edge * 72, cudaMemcpyDeviceToHost, C->st));
    if (edge_valid) CK(cudaMemcpyAsync(edge_valid, L.fvalid[0], (size_t)n_edge, cudaMemcpyDeviceToHost, C->st));
    if (edge_nn) CK(cudaMemcpyAsync(edge_nn, L.nn_idx[0], (size_t)n_edge * 20, cudaMemcpyDeviceToHost, C->st));
    if (edge_d2) CK(cudaMemcpyAsync(edge_d2, L.nn_d2[0], (size_t)n_edge * 20, cudaMemcpyDeviceToHost, C->st));
  }
  if (n_surf) {
    CK(cudaMemcpyAsync(pnd.data(), L.surf_pnd, (size_t)n_surf * 56, cudaMemcpyDeviceToHost, C->st));
    if (surf_valid) CK(cudaMemcpyAsync(surf_valid, L.fvalid[1], (size_t)n_surf, cudaMemcpyDeviceToHost, C->st));
    if (surf_nn) CK(cudaMemcpyAsync(surf_nn, L.nn_idx[1], (size_t)n_surf * 20, cudaMemcpyDeviceToHost, C->st));
    if (surf_d2) CK(cudaMemcpyAsync(surf_d2, L.nn_d2[1], (size_t)n_surf * 20, cudaMemcpyDeviceToHost, C->st));
  }
  CK(cudaMemcpyAsync(C->vars_dev + lane, &V, sizeof(V), cudaMemcpyHostToDevice, C->st));  // restore counts
  CK(cudaStreamSynchronize(C->st));
  if (C->cellmap) {  // neighbour indices refer to the cell-ordered storage: report them in the reference's map order
    for (int w = 0; w < 2; ++w) {
      int32_t* nn = w ? surf_nn : edge_nn;
      const int n = w ? n_surf : n_edge;
      if (!nn || !n) continue;
      std::vector<int32_t> rank;
      int nm = 0;
      rc = map_to_pcl(C, lane, w, &nm, &rank);
      if (rc) return rc;
      for (size_t i = 0; i < (size_t)n * 5; ++i)
        if (nn[i] >= 0 && nn[i] < nm) nn[i] = rank[nn[i]];
    }
  }
  if (edge_ab) for (int i = 0; i < n_edge; ++i) memcpy(edge_ab + (size_t)i * 6, pab.data() + (size_t)i * 9 + 3, 6 * sizeof(double));
  if (surf_nd) for (int i = 0; i < n_surf; ++i) memcpy(surf_nd + (size_t)i * 4, pnd.data() + (size_t)i * 7 + 3, 4 * sizeof(double));
  return VILF_OK;
}

int vilf_normal_equations(vilf_handle* h, const double pose[7], const double* edge_pab, int n_edge, const double* surf_pnd, int n_surf,
                          double H21[21], double g6[6], double* cost) {
  HCHECK(h);
  if (!pose) return VILF_ERR_INVALID;
  SolveTraceDev T;
  int rc = run_explicit_solve(C, h->lane, pose, edge_pab, n_edge, surf_pnd, n_surf, 0, nullptr, &T);
  if (rc) return rc;
  if (H21) memcpy(H21, T.H0, sizeof(T.H0));
  if (g6) memcpy(g6, T.g0, sizeof(T.g0));
  if (cost) *cost = T.cost0;
  return VILF_OK;
}

int vilf_solve(vilf_handle* h, double pose_inout[7], const double* edge_pab, int n_edge, const double* surf_pnd, int n_surf, int max_iters,
               double* trace, int max_rows, int* n_rows, int* termination) {
  HCHECK(h);
  if (!pose_inout || max_iters < 0) return VILF_ERR_INVALID;
  SolveTraceDev T;
  double pose_in[7];
  memcpy(pose_in, pose_inout, sizeof(pose_in));
  int rc = run_explicit_solve(C, h->lane, pose_in, edge_pab, n_edge, surf_pnd, n_surf, max_iters, pose_inout, &T);
  if (rc) return rc;
  int n = 0;
  for (int r = 0; r < T.n_rows && r < MAX_TRACE_ROWS && trace && n < max_rows; ++r, ++n) memcpy(trace + (size_t)16 * n, &T.rows[r], 16 * sizeof(double));
  if (n_rows) *n_rows = n;
  if (termination) *termination = T.termination;
  return VILF_OK;
}

int vilf_get_solves(vilf_handle* h, double* out, int max_rows, int* n_rows) {
  HCHECK(h);
  CK(cudaSetDevice(C->device));
  if (!out || !n_rows) return VILF_ERR_INVALID;
  SolveTraceDev T[MAX_OUTER];
  CK(cudaStreamSynchronize(C->st));
  CK(cudaMemcpy(T, C->lanes_host[h->lane].trace, sizeof(T), cudaMemcpyDeviceToHost));
  int n = 0;
  for (int o = 0; o < C->cfg.outer_iters && n < max_rows; ++o) {
    if (T[o].termination < 0 && T[o].n_rows == 0) continue;
    double* r = out + (size_t)8 * n++;
    r[0] = T[o].n_edge; r[1] = T[o].n_surf; r[2] = T[o].termination; r[3] = T[o].n_rows;
    r[4] = T[o].cost0; r[5] = T[o].final_cost; r[6] = r[7] = 0;
  }
  *n_rows = n;
  return VILF_OK;
}

int vilf_state_export(vilf_handle* h, double s[31]) {
  HCHECK(h);
  CK(cudaSetDevice(C->device));
  if (!s) return VILF_ERR_INVALID;
  LaneVars V;
  int rc = read_vars(C, h->lane, &V);
  if (rc) return rc;
  memcpy(s, V.x, 7 * sizeof(double));
  memcpy(s + 7, V.odom, 12 * sizeof(double));
  memcpy(s + 19, V.odom_last, 12 * sizeof(double));
  return VILF_OK;
}

int vilf_state_import(vilf_handle* h, const double s[31], const float* map_edge, int n_edge, const float* map_surf, int n_surf) {
  HCHECK(h);
  CK(cudaSetDevice(C->device));
  const int lane = h->lane;
  if (!s || n_edge < 0 || n_surf < 0 || (n_edge && !map_edge) || (n_surf && !map_surf)) return VILF_ERR_INVALID;
  const int capM = C->cfg.cap_map + C->cfg.cap_scan;
  if (n_edge > capM || n_surf > capM) return fail(C, VILF_ERR_CAPACITY, "map exceeds capacity");
  LaneDev& L = C->lanes_host[lane];
  LaneVars V;
  int rc = read_vars(C, lane, &V);
  if (rc) return rc;
  memcpy(V.x, s, 7 * sizeof(double));
  memcpy(V.odom, s + 7, 12 * sizeof(double));
  memcpy(V.odom_last, s + 19, 12 * sizeof(double));
  V.n_map[0] = n_edge; V.n_map[1] = n_surf; V.status = 0;
  const int cur = C->cur[lane];
  CK(cudaMemcpyAsync(C->vars_dev + lane, &V, sizeof(V), cudaMemcpyHostToDevice, C->st));
  if (C->cellmap) {
    for (int w = 0; w < 2; ++w) {
      const int n = w ? n_surf : n_edge;
      const float* src = w ? map_surf : map_edge;
      if (n > C->cap_aux) return fail(C, VILF_ERR_CAPACITY, "map exceeds capacity");
      if (n) CK(cudaMemcpyAsync(C->aux_in, src, (size_t)n * 16, cudaMemcpyHostToDevice, C->st));
      int hdr[4] = {n, 0, 0, 0};
      CK(cudaMemcpyAsync(C->aux_n, hdr, sizeof(hdr), cudaMemcpyHostToDevice, C->st));
      CellBuildJob B = C->build_host[cur][lane * 2 + w];
      B.src = C->aux_in; B.n = C->aux_n; B.sort.n = C->aux_n; B.status = C->aux_n + 3;
      rc = run_cell_build(C, B);
      if (rc) return rc;
      CK(cudaMemcpyAsync(hdr, C->aux_n, sizeof(hdr), cudaMemcpyDeviceToHost, C->st));
      CK(cudaStreamSynchronize(C->st));
      if (hdr[3]) return status_to_rc(C, hdr[3]);
    }
    MergeVars mz[2];
    memset(mz, 0, sizeof(mz));
    CK(cudaMemcpyAsync(C->mv_dev + lane * 2, mz, sizeof(mz), cudaMemcpyHostToDevice, C->st));  // no orphans carried over
  } else {
    if (n_edge) CK(cudaMemcpyAsync(L.map[0][cur], map_edge, (size_t)n_edge * 16, cudaMemcpyHostToDevice, C->st));
    if (n_surf) CK(cudaMemcpyAsync(L.map[1][cur], map_surf, (size_t)n_surf * 16, cudaMemcpyHostToDevice, C->st));
    build_grids(C, mk(C), C->grid_dev[cur] + lane * 2, 2, C->cluster_map);
  }
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(C->st));
  C->have_map[lane] = 1;
  return VILF_OK;
}

int vilf_profile_enable(vilf_handle* h, int on) {
  HCHECK(h);
  C->profile = on != 0;
  return VILF_OK;
}
static void prof_reset(Ctx* C) {
  memset(C->kernel_ms, 0, sizeof(C->kernel_ms));
  memset(C->kernel_cnt, 0, sizeof(C->kernel_cnt));
  C->frame_ms = 0; C->prof_frames = 0;
}
int vilf_profile_read(vilf_handle* h, double ms_out[7], int64_t* frames, int reset) {
  HCHECK(h);
  if (ms_out) {
    for (int i = 0; i < N_STAGE; ++i) ms_out[i] = 0;
    for (int tag = 0; tag < PROF_TAGS; ++tag) {
      const int ph = tag / PROF_KSLOTS, k = tag % PROF_KSLOTS;
      int stage = ph == 0 ? 0 : ph == 1 ? 1 : ph == 4 ? 2 : ph == 3 ? 5 : (k == K_SOLVE ? 4 : 3);
      ms_out[stage] += C->kernel_ms[tag];
    }
    ms_out[N_STAGE - 1] = C->frame_ms;
  }
  if (frames) *frames = C->prof_frames;
  if (reset) prof_reset(C);
  return VILF_OK;
}
int vilf_profile_read_kernels(vilf_handle* h, double* ms_out, int64_t* launches_out, int n_tags, int reset) {
  HCHECK(h);
  if (n_tags != PROF_TAGS) return VILF_ERR_INVALID;
  if (ms_out) memcpy(ms_out, C->kernel_ms, sizeof(C->kernel_ms));
  if (launches_out) memcpy(launches_out, C->kernel_cnt, sizeof(C->kernel_cnt));
  if (reset) prof_reset(C);
  return VILF_OK;
}
const char* vilf_profile_kernel_name(int kernel) {
  static const char* names[K_COUNT] = {"k_frame_reset", "k_sort_keyhist<KeyGenRing>", "k_sort_hist", "k_sort_scatter", "k_sector_select", "k_compact_features",
                                       "k_vox_bbox", "k_sort_keyhist<KeyGenVoxel>", "k_vox_heads", "k_vox_centroid", "k_map_append", "k_map_init",
                                       "k_grid_zero", "k_grid_count", "k_grid_scan_partial", "k_grid_scan_final", "k_grid_scatter", "k_knn_assoc",
                                       "k_knn_only", "k_solve", "k_fit", "k_voxel_cluster", "k_grid_cluster", "k_depth_cloud", "k_depth_query", "k_ring_partition",
                                       "k_new_xform", "k_sort_keyhist<KeyGenNew>", "k_merge_partition", "k_merge<count>", "k_merge<emit>", "k_cell_build", "k_knn_cell_assoc", "k_sc", "k_ri_*", "k_ri_select"};
  return (kernel >= 0 && kernel < K_COUNT) ? names[kernel] : "";
}
int vilf_launch_count(vilf_handle* h, int64_t* launches) {
  HCHECK(h);
  if (!launches) return VILF_ERR_INVALID;
  *launches = C->launches;
  return VILF_OK;
}
int vilf_get_stream(vilf_handle* h, void** cuda_stream) {
  HCHECK(h);
  if (!cuda_stream) return VILF_ERR_INVALID;
  *cuda_stream = (void*)C->st;
  return VILF_OK;
}
int vilf_debug_voxel_phases(vilf_handle* h, int job, int64_t out8[8]) {
  HCHECK(h);
  CK(cudaSetDevice(C->device));
  if (job < 0 || job >= VV_PER_LANE || !out8) return VILF_ERR_INVALID;
  VoxVars vv;
  CK(cudaStreamSynchronize(C->st));
  CK(cudaMemcpy(&vv, C->vv_dev + (size_t)h->lane * VV_PER_LANE + job, sizeof(vv), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 8; ++i) out8[i] = (int64_t)vv.t[i];
  return VILF_OK;
}
int vilf_get_counts(vilf_handle* h, int32_t out8[8]) {
  HCHECK(h);
  CK(cudaSetDevice(C->device));
  if (!out8) return VILF_ERR_INVALID;
  LaneVars V;
  int rc = read_vars(C, h->lane, &V);
  if (rc) return rc;
  out8[0] = V.n_edge; out8[1] = V.n_surf; out8[2] = V.n_ds[0]; out8[3] = V.n_ds[1];
  out8[4] = V.n_map[0]; out8[5] = V.n_map[1]; out8[6] = V.status; out8[7] = V.frames;
  return VILF_OK;
}

}  // extern "C"

namespace vilf {
// cloudNoRegistered as vilf_get_cloud(h, 5) hands it out (EM:110-114 after localMapInited, EM:304-305 otherwise), in place.
int resident_scan_features(vilf_handle* h, const float4* p[2], const int* n[2], cudaStream_t* st, int* device) {
  if (!h || !h->ctx) return VILF_ERR_INVALID;
  Ctx* C = h->ctx;
  const int lane = h->lane;
  if (!C->have_map[lane]) return VILF_ERR_STATE;
  LaneDev& L = C->lanes_host[lane];
  LaneVars* V = C->vars_dev + lane;
  if (C->last_init[lane]) { p[0] = L.feat[0]; n[0] = &V->n_edge; p[1] = L.feat[1]; n[1] = &V->n_surf; }
  else { p[0] = L.ds[0]; n[0] = &V->n_ds[0]; p[1] = L.ds[1]; n[1] = &V->n_ds[1]; }
  *st = C->st;
  *device = C->device;
  return VILF_OK;
}
}  // namespace vilf
