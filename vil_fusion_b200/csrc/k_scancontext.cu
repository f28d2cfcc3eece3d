// ScanContext place-recognition front end next to the odometry path: the descriptor, its ring / sector
// keys, the column-shift distance and the loop-candidate search of SCManager
// (src/global_fusion/include/Scancontext/Scancontext.h:42-299, helpers src/global_fusion/include/common.h:79-114), on the cloud
// the odometry already holds on the device (the voxel-filtered scan features of the last frame = what the node publishes as
// /GlobalMap and global_fusion turns into a key frame, poseGraphOptimization.cpp:553).
//
//   k_sc_bin       polar binning: ring = ceil(range / max_radius * R), sector = ceil(theta / 360 * S), max z per bin (:49-72).
//                  A max of floats is order independent: CTA-local bins in shared memory (atomicMax on the order-preserving
//                  integer image of the float), merged into the global bins.
//   k_sc_finish    empty bins -> 0 (:75-78), ring key = row means (:86-99), sector key = column means (:102-115)
//   k_sc_distance  distanceBtnScanContext (:163-193): sector-key alignment over all S shifts, then the cosine distance of the
//                  2 * radius + 1 shifts around it; one CTA, every sum in the oracle's sequential order (no FMA), so the result is
//                  bit-identical to the CPU restatement the tests compare with
//   k_sc_keydist + k_sc_detect   detectLoopClosureID (:210-299): exact nearest ring keys over the tree snapshot (brute force in
//                  nanoflann's metric_L2 summation order instead of the kd-tree), then the distance to each candidate
//
// pcl::IterativeClosestPoint (poseGraphOptimization.cpp:376-444) is not part of this file.
#include "vilf_internal.cuh"
#include "../../include/vilf.h"
#include <cstdio>
#include <vector>

namespace vilf {

constexpr int SC_NO_POINT = -1000;

__device__ __forceinline__ float sc_theta(float x, float y) {  // common.h:79-92 (float atan -> double product -> float)
  const double k = 180.0 / M_PI;
  if ((x >= 0) & (y >= 0)) return (float)dmul(k, (double)(float)atan((double)__fdiv_rn(y, x)));
  if ((x < 0) & (y >= 0)) return (float)dsub(180.0, dmul(k, (double)(float)atan((double)__fdiv_rn(y, -x))));
  if ((x < 0) & (y < 0)) return (float)dadd(180.0, dmul(k, (double)(float)atan((double)__fdiv_rn(y, x))));
  if ((x >= 0) & (y < 0)) return (float)dsub(360.0, dmul(k, (double)(float)atan((double)__fdiv_rn(-y, x))));
  return nanf("");
}

__global__ void k_sc_reset(int* bins, int nb) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += gridDim.x * blockDim.x) bins[i] = f2ord((float)SC_NO_POINT);
}

// Up to two cloud segments (the resident edge + surf features); n1_dev may be null.
__global__ void __launch_bounds__(256) k_sc_bin(const float4* __restrict__ p0, const int* n0_dev, const float4* __restrict__ p1, const int* n1_dev, ScParams P,
                                                 int* bins) {
  extern __shared__ int sbin[];
  const int nb = P.num_ring * P.num_sector;
  for (int i = threadIdx.x; i < nb; i += 256) sbin[i] = f2ord((float)SC_NO_POINT);
  __syncthreads();
  const int n0 = *n0_dev, n1 = n1_dev ? *n1_dev : 0;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n0 + n1; i += gridDim.x * 256) {
    const float4 p = i < n0 ? p0[i] : p1[i - n0];
    const float pz = (float)dadd((double)p.z, P.lidar_height);                       // :56 (pt.z is a float)
    const float range = __fsqrt_rn(fadd(fmul(p.x, p.x), fmul(p.y, p.y)));            // :59
    const float theta = sc_theta(p.x, p.y);                                          // :60
    if ((double)range > P.max_radius) continue;                                      // :63
    int ring = (int)ceil(dmul((double)range / P.max_radius, (double)P.num_ring));    // :66
    int sector = (int)ceil(dmul((double)theta / 360.0, (double)P.num_sector));       // :67 (NaN -> 0 here, INT_MIN on x86: both clamp to 1)
    ring = max(min(P.num_ring, ring), 1);
    sector = max(min(P.num_sector, sector), 1);
    atomicMax(&sbin[(ring - 1) * P.num_sector + (sector - 1)], f2ord(pz));           // :70-71
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += 256)
    if (sbin[i] != f2ord((float)SC_NO_POINT)) atomicMax(&bins[i], sbin[i]);
}

__global__ void __launch_bounds__(256) k_sc_finish(const int* bins, ScParams P, double* desc, double* ringkey, double* sectorkey, float* invkey) {
  extern __shared__ double sd[];
  const int R = P.num_ring, S = P.num_sector;
  for (int i = threadIdx.x; i < R * S; i += 256) {
    const float v = ord2f(bins[i]);
    const double d = v == (float)SC_NO_POINT ? 0.0 : (double)v;  // :75-78
    sd[i] = d; desc[i] = d;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < R; r += 256) {
    double s = 0;
    for (int c = 0; c < S; ++c) s = dadd(s, sd[r * S + c]);
    const double k = s / (double)S;
    ringkey[r] = k;
    invkey[r] = (float)k;  // eig2stdvec (common.h:117): the search key is the ring key rounded to float
  }
  for (int c = threadIdx.x; c < S; c += 256) {
    double s = 0;
    for (int r = 0; r < R; ++r) s = dadd(s, sd[r * S + c]);
    sectorkey[c] = s / (double)R;
  }
}

// distanceBtnScanContext by one CTA.  sm: 2 * R * S doubles (the two descriptors) + 2 * S (sector keys) + S (alignment norms)
// + (2 * radius + 1) * S similarities + flags.  Result -> out[0] = distance, out[1] = shift (as a double).
__device__ void sc_distance_cta(const double* __restrict__ g1, const double* __restrict__ g2, const ScParams& P, double* sm, double* out) {
  const int R = P.num_ring, S = P.num_sector, tid = threadIdx.x, nt = blockDim.x;
  double* a = sm;
  double* b = a + R * S;
  double* v1 = b + R * S;
  double* v2 = v1 + S;
  double* nrm = v2 + S;
  double* sim = nrm + S;                 // [nshift][S], NaN marks a sector pair that is not counted
  __shared__ int space[64];
  __shared__ int nshift_s;
  for (int i = tid; i < R * S; i += nt) { a[i] = g1[i]; b[i] = g2[i]; }
  __syncthreads();
  for (int c = tid; c < S; c += nt) {     // makeSectorkeyFromScancontext :102-115
    double s1 = 0, s2 = 0;
    for (int r = 0; r < R; ++r) { s1 = dadd(s1, a[r * S + c]); s2 = dadd(s2, b[r * S + c]); }
    v1[c] = s1 / (double)R; v2[c] = s2 / (double)R;
  }
  __syncthreads();
  for (int sh = tid; sh < S; sh += nt) {  // fastAlignUsingVkey :119-138
    double s = 0;
    for (int c = 0; c < S; ++c) { const double d = dsub(v1[c], v2[((c - sh) % S + S) % S]); s = dadd(s, dmul(d, d)); }
    nrm[sh] = sqrt(s);
  }
  __syncthreads();
  if (tid == 0) {
    int argmin = 0;
    double mn = 10000000;
    for (int sh = 0; sh < S; ++sh)
      if (nrm[sh] < mn) { argmin = sh; mn = nrm[sh]; }
    int radius = (int)round(dmul(dmul(0.5, P.search_ratio), (double)S));  // :170
    if (2 * radius + 1 > 64) radius = 31;
    int n = 0;
    space[n++] = argmin;
    for (int ii = 1; ii < radius + 1; ++ii) { space[n++] = (argmin + ii + S) % S; space[n++] = (argmin - ii + S) % S; }
    for (int i = 1; i < n; ++i) {  // std::sort ascending (:177)
      const int v = space[i];
      int j = i - 1;
      while (j >= 0 && space[j] > v) { space[j + 1] = space[j]; --j; }
      space[j + 1] = v;
    }
    nshift_s = n;
  }
  __syncthreads();
  const int nshift = nshift_s;
  for (int t = tid; t < nshift * S; t += nt) {  // distDirectSC :140-161, one (shift, column) per thread
    const int k = t / S, c = t % S;
    const int c2 = ((c - space[k]) % S + S) % S;
    double n1 = 0, n2 = 0, dot = 0;
    for (int r = 0; r < R; ++r) {
      const double x = a[r * S + c], y = b[r * S + c2];
      n1 = dadd(n1, dmul(x, x)); n2 = dadd(n2, dmul(y, y)); dot = dadd(dot, dmul(x, y));
    }
    n1 = sqrt(n1); n2 = sqrt(n2);
    sim[t] = ((n1 == 0) | (n2 == 0)) ? nan("") : dot / dmul(n1, n2);
  }
  __syncthreads();
  for (int k = tid; k < nshift; k += nt) {
    int num_eff = 0;
    double sum = 0;
    for (int c = 0; c < S; ++c) {
      const double s = sim[k * S + c];
      if (s == s) { sum = dadd(sum, s); ++num_eff; }
    }
    nrm[k] = dsub(1.0, sum / (double)num_eff);  // reuse: distance of shift k
  }
  __syncthreads();
  if (tid == 0) {
    int argmin = 0;
    double mn = 10000000;
    for (int k = 0; k < nshift; ++k)
      if (nrm[k] < mn) { argmin = space[k]; mn = nrm[k]; }
    out[0] = mn; out[1] = (double)argmin;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) k_sc_distance(const double* sc1, const double* sc2, ScParams P, double* out) {
  extern __shared__ double sm[];
  sc_distance_cta(sc1, sc2, P, sm, out);
}

// squared L2 of the current ring key to every key of the snapshot, nanoflann metric_L2 order (groups of four)
__global__ void __launch_bounds__(256) k_sc_keydist(const float* __restrict__ keys, int n_snapshot, const float* __restrict__ cur, int dim, float* dist) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n_snapshot) return;
  const float* b = keys + (size_t)i * dim;
  float result = 0.f;
  int d = 0;
  for (; d + 3 < dim; d += 4) {
    const float d0 = fsub(cur[d], b[d]), d1 = fsub(cur[d + 1], b[d + 1]), d2 = fsub(cur[d + 2], b[d + 2]), d3 = fsub(cur[d + 3], b[d + 3]);
    result = fadd(result, fadd(fadd(fadd(fmul(d0, d0), fmul(d1, d1)), fmul(d2, d2)), fmul(d3, d3)));
  }
  for (; d < dim; ++d) { const float d0 = fsub(cur[d], b[d]); result = fadd(result, fmul(d0, d0)); }
  dist[i] = result;
}

// Candidates = the num_candidates nearest keys by (distance, index); slots the search cannot fill stay 0 (Scancontext.h:247);
// then distanceBtnScanContext to each, strict-< minimum in candidate order (:259-272), threshold (:279).
// res: [0] loop id, [1] yaw difference (rad, as float bits in a double), [2] min distance, [3] nearest index.
__global__ void __launch_bounds__(256) k_sc_detect(const float* __restrict__ dist, int n_snapshot, const double* __restrict__ descs, int cur_index, ScParams P,
                                                    double* res) {
  extern __shared__ double sm[];
  __shared__ float bd[256 * 4];
  __shared__ int bi[256 * 4];
  __shared__ int cand[8];
  __shared__ double pair_out[2];
  const int tid = threadIdx.x, K = min(P.num_candidates, 4);
  float ld[4]; int li[4];
  for (int k = 0; k < 4; ++k) { ld[k] = FLT_MAX; li[k] = INT_MAX; }
  for (int i = tid; i < n_snapshot; i += 256) {  // ascending i: strict < keeps the lower index on ties
    const float d = dist[i];
    if (d < ld[K - 1]) {
      int k = K - 1;
      while (k > 0 && ld[k - 1] > d) { ld[k] = ld[k - 1]; li[k] = li[k - 1]; --k; }
      ld[k] = d; li[k] = i;
    }
  }
  for (int k = 0; k < 4; ++k) { bd[tid * 4 + k] = ld[k]; bi[tid * 4 + k] = li[k]; }
  __syncthreads();
  if (tid == 0) {
    for (int c = 0; c < K; ++c) {
      float best = FLT_MAX; int bidx = INT_MAX, bpos = -1;
      for (int t = 0; t < 256 * 4; ++t) {
        if (bi[t] == INT_MAX) continue;
        if (bd[t] < best || (bd[t] == best && bi[t] < bidx)) { best = bd[t]; bidx = bi[t]; bpos = t; }
      }
      cand[c] = bpos >= 0 ? bidx : 0;
      if (bpos >= 0) bi[bpos] = INT_MAX;
    }
  }
  __syncthreads();
  const int RS = P.num_ring * P.num_sector;
  double min_dist = 10000000;
  int nn_align = 0, nn_idx = 0;
  for (int c = 0; c < P.num_candidates && c < 8; ++c) {
    const int ci = c < K ? cand[c] : 0;
    sc_distance_cta(descs + (size_t)cur_index * RS, descs + (size_t)ci * RS, P, sm, pair_out);
    if (pair_out[0] < min_dist) { min_dist = pair_out[0]; nn_align = (int)pair_out[1]; nn_idx = ci; }
    __syncthreads();
  }
  if (tid == 0) {
    res[0] = min_dist < P.dist_thres ? (double)nn_idx : -1.0;
    const double unit = 360.0 / (double)P.num_sector;
    res[1] = (double)(float)(dmul(dmul((double)nn_align, unit), M_PI) / 180.0);  // deg2rad(nn_align * PC_UNIT_SECTORANGLE) as float (:296)
    res[2] = min_dist; res[3] = (double)nn_idx;
  }
}

static size_t sc_dist_smem(const ScParams& P) {
  const int R = P.num_ring, S = P.num_sector;
  int radius = (int)llround(0.5 * P.search_ratio * S);
  if (2 * radius + 1 > 64) radius = 31;
  return sizeof(double) * ((size_t)2 * R * S + 3 * S + (size_t)(2 * radius + 1) * S + 8);
}

cudaError_t init_sc_kernels() {
  cudaError_t e = cudaFuncSetAttribute(k_sc_distance, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_sc_detect, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
}

void launch_sc_make(const Launch& L, const float4* p0, const int* n0_dev, const float4* p1, const int* n1_dev, const ScParams& P, int* bins, double* desc,
                    double* ringkey, double* sectorkey, float* invkey) {
  const int nb = P.num_ring * P.num_sector;
  k_sc_reset<<<(nb + 255) / 256, 256, 0, L.st>>>(bins, nb);
  L.tick(K_SC);
  k_sc_bin<<<148, 256, nb * sizeof(int), L.st>>>(p0, n0_dev, p1, n1_dev, P, bins);
  L.tick(K_SC);
  k_sc_finish<<<1, 256, nb * sizeof(double), L.st>>>(bins, P, desc, ringkey, sectorkey, invkey);
  L.tick(K_SC);
}
void launch_sc_distance(const Launch& L, const double* sc1, const double* sc2, const ScParams& P, double* out) {
  k_sc_distance<<<1, 256, sc_dist_smem(P), L.st>>>(sc1, sc2, P, out);
  L.tick(K_SC);
}
void launch_sc_detect(const Launch& L, const float* keys, int n_snapshot, const float* cur_key, const double* descs, int cur_index, const ScParams& P, float* dist,
                      double* res) {
  if (n_snapshot > 0) {
    k_sc_keydist<<<(n_snapshot + 255) / 256, 256, 0, L.st>>>(keys, n_snapshot, cur_key, P.num_ring, dist);
    L.tick(K_SC);
  }
  k_sc_detect<<<1, 256, sc_dist_smem(P), L.st>>>(dist, n_snapshot, descs, cur_index, P, res);
  L.tick(K_SC);
}

}  // namespace vilf


// ------------------------------------------------------------------------------------------------
// C ABI (include/vilf.h: vilf_sc_*)
// ------------------------------------------------------------------------------------------------
struct vilf_sc {
  int device = 0;
  vilf_sc_params up;
  vilf::ScParams P;
  cudaStream_t st = nullptr;
  cudaEvent_t ev = nullptr;
  int64_t launches = 0;
  int cap = 0, cap_pts = 0;
  int n = 0;            // key frames stored (polarcontexts_.size())
  int n_snapshot = 0;   // keys the "tree" holds (polarcontext_invkeys_to_search_)
  int tree_counter = 0; // tree_making_period_conter
  int RS = 0;
  int* bins = nullptr;
  double* descs = nullptr; double* ringkeys = nullptr; double* sectorkeys = nullptr;
  float* invkeys = nullptr; float* dist = nullptr;
  double* res = nullptr;       // [4] detect result, [4..5] distance result
  double* tmp_desc = nullptr;  // [2][RS] explicit descriptors of vilf_sc_distance
  float4* pts = nullptr; int* n_dev = nullptr;
  std::vector<void*> allocs;
  char err[256] = {0};
};

namespace {
#define SCK(call)                                                                                                  \
  do {                                                                                                             \
    cudaError_t e_ = (call);                                                                                       \
    if (e_ != cudaSuccess) {                                                                                       \
      snprintf(sc->err, sizeof(sc->err), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return VILF_ERR_CUDA;                                                                                        \
    }                                                                                                              \
  } while (0)
int sc_fail(vilf_sc* sc, int code, const char* msg) { snprintf(sc->err, sizeof(sc->err), "%s", msg); return code; }
template <class T>
cudaError_t sc_alloc(vilf_sc* sc, T** p, size_t count) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 256);
  if (e != cudaSuccess) return e;
  sc->allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return cudaMemset(q, 0, count * sizeof(T) + 256);
}
vilf::Launch sc_launch(vilf_sc* sc) { vilf::Launch L; L.st = sc->st; L.counter = &sc->launches; L.prof = nullptr; return L; }
// descriptor + keys of the cloud (p0, n0) ++ (p1, n1) into slot sc->n
int sc_append(vilf_sc* sc, const float4* p0, const int* n0, const float4* p1, const int* n1) {
  const int i = sc->n;
  vilf::launch_sc_make(sc_launch(sc), p0, n0, p1, n1, sc->P, sc->bins, sc->descs + (size_t)i * sc->RS, sc->ringkeys + (size_t)i * sc->P.num_ring,
                       sc->sectorkeys + (size_t)i * sc->P.num_sector, sc->invkeys + (size_t)i * sc->P.num_ring);
  SCK(cudaGetLastError());
  sc->n = i + 1;
  return VILF_OK;
}
}  // namespace

extern "C" {

int vilf_sc_default_params(vilf_sc_params* p) {
  if (!p) return VILF_ERR_INVALID;
  p->lidar_height = 2.0; p->num_ring = 20; p->num_sector = 60; p->max_radius = 80.0;   // Scancontext.h:313-318
  p->num_exclude_recent = 30; p->num_candidates = 3;                                    // :323-324
  p->search_ratio = 0.1; p->dist_thres = 0.2; p->tree_making_period = 30;               // :327-331
  p->max_keyframes = 8192; p->max_points = 1 << 18; p->pad_ = 0;
  return VILF_OK;
}

int vilf_sc_create(const vilf_sc_params* params, int device, vilf_sc** out) {
  if (!out) return VILF_ERR_INVALID;
  *out = nullptr;
  vilf_sc_params up;
  if (params) up = *params; else vilf_sc_default_params(&up);
  if (up.num_ring < 1 || up.num_sector < 1 || up.num_ring * up.num_sector > 4096 || up.max_keyframes < 1 || up.max_points < 1 || !(up.max_radius > 0) ||
      up.num_exclude_recent < 0 || up.tree_making_period < 1 || up.num_candidates < 1)
    return VILF_ERR_INVALID;
  if (up.num_candidates > 4) return VILF_ERR_UNSUPPORTED;
  if (cudaSetDevice(device) != cudaSuccess) return VILF_ERR_CUDA;
  vilf_sc* sc = new vilf_sc();
  sc->device = device; sc->up = up;
  sc->P.lidar_height = up.lidar_height; sc->P.max_radius = up.max_radius; sc->P.search_ratio = up.search_ratio; sc->P.dist_thres = up.dist_thres;
  sc->P.num_ring = up.num_ring; sc->P.num_sector = up.num_sector; sc->P.num_exclude_recent = up.num_exclude_recent; sc->P.num_candidates = up.num_candidates;
  sc->P.tree_making_period = up.tree_making_period;
  sc->cap = up.max_keyframes; sc->cap_pts = up.max_points; sc->RS = up.num_ring * up.num_sector;
  auto boot = [&]() -> int {
    SCK(vilf::init_sc_kernels());
    SCK(cudaStreamCreateWithFlags(&sc->st, cudaStreamNonBlocking));
    SCK(cudaEventCreateWithFlags(&sc->ev, cudaEventDisableTiming));
    SCK(sc_alloc(sc, &sc->bins, (size_t)sc->RS));
    SCK(sc_alloc(sc, &sc->descs, (size_t)sc->cap * sc->RS));
    SCK(sc_alloc(sc, &sc->ringkeys, (size_t)sc->cap * up.num_ring));
    SCK(sc_alloc(sc, &sc->sectorkeys, (size_t)sc->cap * up.num_sector));
    SCK(sc_alloc(sc, &sc->invkeys, (size_t)sc->cap * up.num_ring));
    SCK(sc_alloc(sc, &sc->dist, (size_t)sc->cap));
    SCK(sc_alloc(sc, &sc->res, (size_t)8));
    SCK(sc_alloc(sc, &sc->tmp_desc, (size_t)2 * sc->RS));
    SCK(sc_alloc(sc, &sc->pts, (size_t)sc->cap_pts));
    SCK(sc_alloc(sc, &sc->n_dev, (size_t)4));
    return VILF_OK;
  };
  const int rc = boot();
  if (rc) { vilf_sc_destroy(sc); return rc; }
  *out = sc;
  return VILF_OK;
}

int vilf_sc_destroy(vilf_sc* sc) {
  if (!sc) return VILF_ERR_INVALID;
  cudaSetDevice(sc->device);
  if (sc->st) { cudaStreamSynchronize(sc->st); cudaStreamDestroy(sc->st); }
  if (sc->ev) cudaEventDestroy(sc->ev);
  for (void* p : sc->allocs) cudaFree(p);
  delete sc;
  return VILF_OK;
}

const char* vilf_sc_last_error(vilf_sc* sc) { return sc ? sc->err : "null vilf_sc"; }

int vilf_sc_size(vilf_sc* sc, int* n) {
  if (!sc || !n) return VILF_ERR_INVALID;
  *n = sc->n;
  return VILF_OK;
}

int vilf_sc_make_and_save(vilf_sc* sc, const float* xyzi, int n) {
  if (!sc || n < 0 || (n > 0 && !xyzi)) return VILF_ERR_INVALID;
  SCK(cudaSetDevice(sc->device));
  if (n > sc->cap_pts) return sc_fail(sc, VILF_ERR_CAPACITY, "cloud exceeds max_points");
  if (sc->n >= sc->cap) return sc_fail(sc, VILF_ERR_CAPACITY, "max_keyframes reached");
  if (n) SCK(cudaMemcpyAsync(sc->pts, xyzi, (size_t)n * 16, cudaMemcpyHostToDevice, sc->st));
  SCK(cudaMemcpyAsync(sc->n_dev, &n, sizeof(int), cudaMemcpyHostToDevice, sc->st));
  SCK(cudaStreamSynchronize(sc->st));  // &n lives on this stack frame
  return sc_append(sc, sc->pts, sc->n_dev, nullptr, nullptr);
}

int vilf_sc_make_and_save_resident(vilf_sc* sc, vilf_handle* h) {
  if (!sc || !h) return VILF_ERR_INVALID;
  SCK(cudaSetDevice(sc->device));
  if (sc->n >= sc->cap) return sc_fail(sc, VILF_ERR_CAPACITY, "max_keyframes reached");
  const float4* p[2]; const int* n[2];
  cudaStream_t hst = nullptr;
  int dev = -1;
  const int rc = vilf::resident_scan_features(h, p, n, &hst, &dev);
  if (rc) return sc_fail(sc, rc, "the odometry handle holds no scan features");
  if (dev != sc->device) return sc_fail(sc, VILF_ERR_INVALID, "odometry handle lives on another device");
  SCK(cudaEventRecord(sc->ev, hst));          // after everything the odometry has queued so far
  SCK(cudaStreamWaitEvent(sc->st, sc->ev, 0));
  const int r2 = sc_append(sc, p[0], n[0], p[1], n[1]);
  if (r2) return r2;
  SCK(cudaEventRecord(sc->ev, sc->st));       // the next frame must not overwrite the features while the bins are being filled
  SCK(cudaStreamWaitEvent(hst, sc->ev, 0));
  return VILF_OK;
}

int vilf_sc_get(vilf_sc* sc, int index, double* desc, double* ringkey, double* sectorkey) {
  if (!sc) return VILF_ERR_INVALID;
  SCK(cudaSetDevice(sc->device));
  if (index < 0) index += sc->n;
  if (index < 0 || index >= sc->n) return sc_fail(sc, VILF_ERR_INVALID, "no such key frame");
  if (desc) SCK(cudaMemcpyAsync(desc, sc->descs + (size_t)index * sc->RS, (size_t)sc->RS * 8, cudaMemcpyDeviceToHost, sc->st));
  if (ringkey) SCK(cudaMemcpyAsync(ringkey, sc->ringkeys + (size_t)index * sc->P.num_ring, (size_t)sc->P.num_ring * 8, cudaMemcpyDeviceToHost, sc->st));
  if (sectorkey) SCK(cudaMemcpyAsync(sectorkey, sc->sectorkeys + (size_t)index * sc->P.num_sector, (size_t)sc->P.num_sector * 8, cudaMemcpyDeviceToHost, sc->st));
  SCK(cudaStreamSynchronize(sc->st));
  return VILF_OK;
}

int vilf_sc_distance(vilf_sc* sc, const double* sc1, const double* sc2, double* dist, int* shift) {
  if (!sc || !sc1 || !sc2) return VILF_ERR_INVALID;
  SCK(cudaSetDevice(sc->device));
  SCK(cudaMemcpyAsync(sc->tmp_desc, sc1, (size_t)sc->RS * 8, cudaMemcpyHostToDevice, sc->st));
  SCK(cudaMemcpyAsync(sc->tmp_desc + sc->RS, sc2, (size_t)sc->RS * 8, cudaMemcpyHostToDevice, sc->st));
  vilf::launch_sc_distance(sc_launch(sc), sc->tmp_desc, sc->tmp_desc + sc->RS, sc->P, sc->res + 4);
  SCK(cudaGetLastError());
  double r[2];
  SCK(cudaMemcpyAsync(r, sc->res + 4, sizeof(r), cudaMemcpyDeviceToHost, sc->st));
  SCK(cudaStreamSynchronize(sc->st));
  if (dist) *dist = r[0];
  if (shift) *shift = (int)r[1];
  return VILF_OK;
}

int vilf_sc_distance_between(vilf_sc* sc, int i, int j, double* dist, int* shift) {
  if (!sc) return VILF_ERR_INVALID;
  SCK(cudaSetDevice(sc->device));
  if (i < 0) i += sc->n;
  if (j < 0) j += sc->n;
  if (i < 0 || j < 0 || i >= sc->n || j >= sc->n) return sc_fail(sc, VILF_ERR_INVALID, "no such key frame");
  vilf::launch_sc_distance(sc_launch(sc), sc->descs + (size_t)i * sc->RS, sc->descs + (size_t)j * sc->RS, sc->P, sc->res + 4);
  SCK(cudaGetLastError());
  double r[2];
  SCK(cudaMemcpyAsync(r, sc->res + 4, sizeof(r), cudaMemcpyDeviceToHost, sc->st));
  SCK(cudaStreamSynchronize(sc->st));
  if (dist) *dist = r[0];
  if (shift) *shift = (int)r[1];
  return VILF_OK;
}

int vilf_sc_detect_loop_closure(vilf_sc* sc, int* loop_id, float* yaw_diff_rad, double* min_dist, int* nn_idx) {
  if (!sc) return VILF_ERR_INVALID;
  SCK(cudaSetDevice(sc->device));
  if (loop_id) *loop_id = -1;
  if (yaw_diff_rad) *yaw_diff_rad = 0.0f;
  if (min_dist) *min_dist = 10000000;
  if (nn_idx) *nn_idx = 0;
  if (sc->n < 1) return sc_fail(sc, VILF_ERR_STATE, "no key frame stored");   // the reference would call back() on an empty vector
  if (sc->n < sc->P.num_exclude_recent + 1) return VILF_OK;                    // Scancontext.h:220-224: early return, no loop
  if (sc->tree_counter % sc->P.tree_making_period == 0) sc->n_snapshot = sc->n - sc->P.num_exclude_recent;  // :227-238
  sc->tree_counter = sc->tree_counter + 1;
  const int cur = sc->n - 1;
  vilf::launch_sc_detect(sc_launch(sc), sc->invkeys, sc->n_snapshot, sc->invkeys + (size_t)cur * sc->P.num_ring, sc->descs, cur, sc->P, sc->dist, sc->res);
  SCK(cudaGetLastError());
  double r[4];
  SCK(cudaMemcpyAsync(r, sc->res, sizeof(r), cudaMemcpyDeviceToHost, sc->st));
  SCK(cudaStreamSynchronize(sc->st));
  if (loop_id) *loop_id = (int)r[0];
  if (yaw_diff_rad) *yaw_diff_rad = (float)r[1];
  if (min_dist) *min_dist = r[2];
  if (nn_idx) *nn_idx = (int)r[3];
  return VILF_OK;
}

int vilf_sc_launch_count(vilf_sc* sc, int64_t* launches) {
  if (!sc || !launches) return VILF_ERR_INVALID;
  *launches = sc->launches;
  return VILF_OK;
}

}  // extern "C"
