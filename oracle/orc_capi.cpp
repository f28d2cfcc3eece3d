// TEST INFRASTRUCTURE — C entry points of the CPU oracle (loaded with ctypes by tests/ and by
// bench.py's CPU-baseline legs only).  NOT product code; PARITY UNPINNED, see orc_pipeline.hpp.
#include "orc_pipeline.hpp"
#include "orc_depth.hpp"
#include "orc_scancontext.hpp"
#include "orc_rangeimage.hpp"

using namespace orc;

extern "C" {

struct orc_config {
  int n_scan, n_rings;
  double lidar_min, lidar_max, edge_threshold, edge_leaf, surf_leaf, crop_half, knn_gate, huber;
  int outer_iters, lm_max_iters, voxel_order, knn_ties;
  int eig_alg, plane_alg, centroid_div, lm_solver;
};

static Config to_cfg(const orc_config* c) {
  Config k;
  k.n_scan = c->n_scan; k.n_rings = c->n_rings; k.lidar_min = c->lidar_min; k.lidar_max = c->lidar_max;
  k.edge_threshold = c->edge_threshold; k.edge_leaf = c->edge_leaf; k.surf_leaf = c->surf_leaf; k.crop_half = c->crop_half;
  k.knn_gate = c->knn_gate; k.huber = c->huber; k.outer_iters = c->outer_iters; k.lm_max_iters = c->lm_max_iters; k.voxel_order = c->voxel_order; k.knn_ties = c->knn_ties;
  k.eig_alg = c->eig_alg; k.plane_alg = c->plane_alg; k.centroid_div = c->centroid_div; k.lm_solver = c->lm_solver;
  return k;
}
static Cloud to_cloud(const float* p, int n) {
  Cloud c((size_t)std::max(n, 0));
  if (n > 0) std::memcpy(c.data(), p, (size_t)n * sizeof(P4));
  return c;
}

void orc_default_config(orc_config* c) {
  Config k;
  c->n_scan = k.n_scan; c->n_rings = k.n_rings; c->lidar_min = k.lidar_min; c->lidar_max = k.lidar_max; c->edge_threshold = k.edge_threshold;
  c->edge_leaf = k.edge_leaf; c->surf_leaf = k.surf_leaf; c->crop_half = k.crop_half; c->knn_gate = k.knn_gate; c->huber = k.huber;
  c->outer_iters = k.outer_iters; c->lm_max_iters = k.lm_max_iters; c->voxel_order = k.voxel_order; c->knn_ties = k.knn_ties;
  c->eig_alg = k.eig_alg; c->plane_alg = k.plane_alg; c->centroid_div = k.centroid_div; c->lm_solver = k.lm_solver;
}

// Stage 1.  Output buffers hold up to n points; *_src = index of the point in the input scan.
void orc_extract(const orc_config* c, const float* xyzi, int n, const uint16_t* ring, float* edge, int* edge_src, int* n_edge, float* surf,
                 int* surf_src, int* n_surf) {
  Cloud e, s;
  std::vector<int> es, ss;
  extract_features(to_cfg(c), (const P4*)xyzi, n, ring, e, es, s, ss);
  *n_edge = (int)e.size();
  *n_surf = (int)s.size();
  if (!e.empty()) { std::memcpy(edge, e.data(), e.size() * sizeof(P4)); std::memcpy(edge_src, es.data(), es.size() * sizeof(int)); }
  if (!s.empty()) { std::memcpy(surf, s.data(), s.size() * sizeof(P4)); std::memcpy(surf_src, ss.data(), ss.size() * sizeof(int)); }
}

int orc_voxel_grid(const float* pts, int n, float leaf, int order_mode, float* out, int* n_out) {
  Cloud o;
  bool ok = voxel_grid(to_cloud(pts, n), leaf, order_mode, o);
  *n_out = (int)o.size();
  if (!o.empty()) std::memcpy(out, o.data(), o.size() * sizeof(P4));
  return ok ? 1 : 0;
}

void orc_crop_box(const float* pts, int n, const double* mn, const double* mx, float* out, int* n_out) {
  Cloud o;
  crop_box(to_cloud(pts, n), mn, mx, o);
  *n_out = (int)o.size();
  if (!o.empty()) std::memcpy(out, o.data(), o.size() * sizeof(P4));
}

// exact k-NN of nq queries (stride 4 floats) against map (stride 4 floats)
void orc_knn(const float* map, int m, const float* q, int nq, int k, int* idx, float* d2) {
  Cloud c = to_cloud(map, m);
  KdTree t;
  t.build(&c);
  for (int i = 0; i < nq; ++i) t.knn(q + 4 * (size_t)i, k, idx + (size_t)k * i, d2 + (size_t)k * i);
}

// the same with tie class T2 resolved canonically: ascending (d^2, map index) instead of FLANN's first-visited-wins
void orc_knn_canonical(const float* map, int m, const float* q, int nq, int k, int* idx, float* d2) {
  Cloud c = to_cloud(map, m);
  KdTree t;
  t.set_canonical_ties(true);
  t.build(&c);
  for (int i = 0; i < nq; ++i) t.knn(q + 4 * (size_t)i, k, idx + (size_t)k * i, d2 + (size_t)k * i);
}

// Data association at `pose` (EM:117-232).  edge_ab: 6 doubles per edge point (a, b); surf_nd: 4 per surf point (n, d).
void orc_factors(const orc_config* c, const double* pose, const float* edge, int ne, const float* surf, int ns, const float* map_e, int me,
                 const float* map_s, int ms, uint8_t* edge_valid, double* edge_ab, int* edge_nn, float* edge_d2, uint8_t* surf_valid,
                 double* surf_nd, int* surf_nn, float* surf_d2) {
  Config k = to_cfg(c);
  Cloud E = to_cloud(edge, ne), S = to_cloud(surf, ns), ME = to_cloud(map_e, me), MS = to_cloud(map_s, ms);
  KdTree te, ts;
  te.set_canonical_ties(k.knn_ties == 1);
  ts.set_canonical_ties(k.knn_ties == 1);
  te.build(&ME);
  ts.build(&MS);
  Quat q{pose[0], pose[1], pose[2], pose[3]};
  V3 t{pose[4], pose[5], pose[6]};
  std::vector<EdgeFactor> ef;
  std::vector<SurfFactor> sf;
  edge_factors(k, q, t, E, ME, te, ef, edge_nn, edge_d2, edge_valid);
  surf_factors(k, q, t, S, MS, ts, sf, surf_nn, surf_d2, surf_valid);
  size_t j = 0;
  for (int i = 0; i < ne; ++i) {
    double* o = edge_ab + 6 * (size_t)i;
    if (edge_valid[i]) { const EdgeFactor& f = ef[j++]; o[0] = f.a.x; o[1] = f.a.y; o[2] = f.a.z; o[3] = f.b.x; o[4] = f.b.y; o[5] = f.b.z; }
    else for (int z = 0; z < 6; ++z) o[z] = 0;
  }
  j = 0;
  for (int i = 0; i < ns; ++i) {
    double* o = surf_nd + 4 * (size_t)i;
    if (surf_valid[i]) { const SurfFactor& f = sf[j++]; o[0] = f.n.x; o[1] = f.n.y; o[2] = f.n.z; o[3] = f.d; }
    else for (int z = 0; z < 4; ++z) o[z] = 0;
  }
}

static void unpack(const double* edge_pab, int ke, const double* surf_pnd, int ks, std::vector<EdgeFactor>& ef, std::vector<SurfFactor>& sf) {
  ef.resize(ke);
  sf.resize(ks);
  for (int i = 0; i < ke; ++i) {
    const double* p = edge_pab + 9 * (size_t)i;
    ef[i] = {V3{p[0], p[1], p[2]}, V3{p[3], p[4], p[5]}, V3{p[6], p[7], p[8]}};
  }
  for (int i = 0; i < ks; ++i) {
    const double* p = surf_pnd + 7 * (size_t)i;
    sf[i] = {V3{p[0], p[1], p[2]}, V3{p[3], p[4], p[5]}, p[6]};
  }
}

// Robustified normal equations at `pose`: H = J^T J (upper triangle, row-major, 21), g = J^T r (6), cost = 1/2 sum rho.
void orc_normal_eq(double huber, const double* pose, const double* edge_pab, int ke, const double* surf_pnd, int ks, double* H, double* g,
                   double* cost) {
  std::vector<EdgeFactor> ef;
  std::vector<SurfFactor> sf;
  unpack(edge_pab, ke, surf_pnd, ks, ef, sf);
  Problem prob{&ef, &sf, huber};
  const int m = prob.rows();
  std::vector<double> J((size_t)std::max(m, 1) * 6), r(std::max(m, 1));
  *cost = prob.evaluate(pose, r.data(), J.data(), m, g);
  int kk = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j) { double s = 0; for (int q = 0; q < m; ++q) s += J[(size_t)i * m + q] * J[(size_t)j * m + q]; H[kk++] = s; }
}

// ceres::Solve on explicit factors.  trace rows: 16 doubles = iteration, valid, successful, cost, candidate_cost, model_cost_change,
// relative_decrease, radius, step_norm, x[7].  Returns the termination code.
int orc_solve(double huber, int max_iters, double* pose, const double* edge_pab, int ke, const double* surf_pnd, int ks, double* trace,
              int max_rows, int* n_rows) {
  std::vector<EdgeFactor> ef;
  std::vector<SurfFactor> sf;
  unpack(edge_pab, ke, surf_pnd, ks, ef, sf);
  Problem prob{&ef, &sf, huber};
  SolveTrace tr;
  ceres_solve(prob, pose, max_iters, &tr);
  int n = 0;
  for (const LmIter& it : tr.iters) {
    if (n >= max_rows) break;
    double* o = trace + 16 * (size_t)n++;
    o[0] = it.iteration; o[1] = it.step_valid; o[2] = it.step_successful; o[3] = it.cost; o[4] = it.candidate_cost; o[5] = it.model_cost_change;
    o[6] = it.relative_decrease; o[7] = it.radius; o[8] = it.step_norm;
    for (int i = 0; i < 7; ++i) o[9 + i] = it.x[i];
  }
  if (n_rows) *n_rows = n;
  return tr.termination;
}

void orc_se3_plus(const double* x, const double* delta, double* out) { se3_plus(x, delta, out); }

// One edge / surf residual with its local Jacobian (LF:21-52, :79-102), for finite-difference checks.
void orc_edge_eval(const double* pose, const double* pab, double* r3, double* J18) {
  EdgeFactor f{V3{pab[0], pab[1], pab[2]}, V3{pab[3], pab[4], pab[5]}, V3{pab[6], pab[7], pab[8]}};
  edge_eval(f, Quat{pose[0], pose[1], pose[2], pose[3]}, V3{pose[4], pose[5], pose[6]}, r3, J18);
}
void orc_surf_eval(const double* pose, const double* pnd, double* r1, double* J6) {
  SurfFactor f{V3{pnd[0], pnd[1], pnd[2]}, V3{pnd[3], pnd[4], pnd[5]}, pnd[6]};
  surf_eval(f, Quat{pose[0], pose[1], pose[2], pose[3]}, V3{pose[4], pose[5], pose[6]}, r1, J6);
}
void orc_eig3(const double* C9, double* w3, double* V9) {
  M3 C, V;
  std::memcpy(C.m, C9, sizeof(C.m));
  eig3_sym(C, w3, V);
  std::memcpy(V9, V.m, sizeof(V.m));
}
// the sensitivity alternates (orc_math.hpp): Eigen 3.3.7's tridiagonal QR iteration, unpivoted Householder plane fit
void orc_eig3_alt(const double* C9, double* w3, double* V9) {
  M3 C, V;
  std::memcpy(C.m, C9, sizeof(C.m));
  eig3_sym_eigen(C, w3, V);
  std::memcpy(V9, V.m, sizeof(V.m));
}
void orc_lstsq5x3_alt(const double* A15, const double* b5, double* n3) {
  double A[5][3];
  std::memcpy(A, A15, sizeof(A));
  V3 n = lstsq5x3_colpiv(A, b5, false);
  n3[0] = n.x; n3[1] = n.y; n3[2] = n.z;
}
void orc_lstsq5x3(const double* A15, const double* b5, double* n3) {
  double A[5][3];
  std::memcpy(A, A15, sizeof(A));
  V3 n = lstsq5x3_colpiv(A, b5);
  n3[0] = n.x; n3[1] = n.y; n3[2] = n.z;
}

// ---- EstimationMapping object ----
void* orc_odom_create(const orc_config* c) { return new Odometry(to_cfg(c)); }
void orc_odom_destroy(void* h) { delete (Odometry*)h; }
void orc_odom_init_map(void* h, const float* edge, int ne, const float* surf, int ns) { ((Odometry*)h)->init_map(to_cloud(edge, ne), to_cloud(surf, ns)); }
void orc_odom_update(void* h, const float* edge, int ne, const float* surf, int ns, double* pose_out) {
  Odometry* o = (Odometry*)h;
  o->update(to_cloud(edge, ne), to_cloud(surf, ns));
  if (pose_out) std::memcpy(pose_out, o->x, sizeof(o->x));
}
// extractFeature + (frame 0 ? localMapInited : optimation_processing), the node's per-frame sequence (NODE:346-385)
void orc_odom_process_scan(void* h, const float* xyzi, int n, const uint16_t* ring, int first, double* pose_out, int* n_edge, int* n_surf) {
  Odometry* o = (Odometry*)h;
  Cloud e, s;
  std::vector<int> es, ss;
  double t0 = now_s();
  extract_features(o->cfg, (const P4*)xyzi, n, ring, e, es, s, ss);
  o->timing.extract += now_s() - t0;
  if (first) o->init_map(e, s); else o->update(e, s);
  if (pose_out) std::memcpy(pose_out, o->x, sizeof(o->x));
  if (n_edge) *n_edge = (int)e.size();
  if (n_surf) *n_surf = (int)s.size();
}
static Cloud& pick(Odometry* o, int which) {
  switch (which) {
    case 0: return o->map_edge;
    case 1: return o->map_surf;
    case 2: return o->ds_edge;
    case 3: return o->ds_surf;
    case 4: return o->registered;
    default: return o->no_registered;
  }
}
int orc_odom_cloud_size(void* h, int which) { return (int)pick((Odometry*)h, which).size(); }
void orc_odom_get_cloud(void* h, int which, float* out) {
  Cloud& c = pick((Odometry*)h, which);
  if (!c.empty()) std::memcpy(out, c.data(), c.size() * sizeof(P4));
}
void orc_odom_set_cloud(void* h, int which, const float* pts, int n) { pick((Odometry*)h, which) = to_cloud(pts, n); }
// state = pose x[7], odom (R row-major 9 + t 3), odom_last (12): 31 doubles
void orc_odom_get_state(void* h, double* s) {
  Odometry* o = (Odometry*)h;
  std::memcpy(s, o->x, 7 * sizeof(double));
  std::memcpy(s + 7, o->odom.R.m, 9 * sizeof(double));
  s[16] = o->odom.t.x; s[17] = o->odom.t.y; s[18] = o->odom.t.z;
  std::memcpy(s + 19, o->odom_last.R.m, 9 * sizeof(double));
  s[28] = o->odom_last.t.x; s[29] = o->odom_last.t.y; s[30] = o->odom_last.t.z;
}
void orc_odom_set_state(void* h, const double* s) {
  Odometry* o = (Odometry*)h;
  std::memcpy(o->x, s, 7 * sizeof(double));
  std::memcpy(o->odom.R.m, s + 7, 9 * sizeof(double));
  o->odom.t = {s[16], s[17], s[18]};
  std::memcpy(o->odom_last.R.m, s + 19, 9 * sizeof(double));
  o->odom_last.t = {s[28], s[29], s[30]};
}
// per-solve summary of the last update: rows of 8 doubles = n_edge_factors, n_surf_factors, termination, n_iters, cost0, final cost, 0, 0
int orc_odom_get_solves(void* h, double* out, int max_rows) {
  Odometry* o = (Odometry*)h;
  int n = 0;
  for (const SolveTrace& t : o->traces) {
    if (n >= max_rows) break;
    double* r = out + 8 * (size_t)n++;
    r[0] = t.n_edge; r[1] = t.n_surf; r[2] = t.termination; r[3] = (double)t.iters.size();
    r[4] = t.iters.empty() ? 0 : t.iters.front().cost; r[5] = t.iters.empty() ? 0 : t.iters.back().cost; r[6] = r[7] = 0;
  }
  return n;
}
// seconds: extract, scan DS, kd build, association (kNN + fit), solve, map update; frames
void orc_odom_get_timing(void* h, double* out7) {
  Timing& t = ((Odometry*)h)->timing;
  out7[0] = t.extract; out7[1] = t.ds; out7[2] = t.kdbuild; out7[3] = t.assoc; out7[4] = t.solve; out7[5] = t.map; out7[6] = t.frames;
}

// Depth association of visual features (feature_tracker_node.cpp:54-140, :348-361).
void orc_camera_cloud(const float* scan, int n, const double* T16, float* out, int* n_out) {
  Cloud o;
  camera_cloud((const P4*)scan, n, T16, o);
  *n_out = (int)o.size();
  if (!o.empty()) std::memcpy(out, o.data(), o.size() * sizeof(P4));
}
void orc_feature_depth(const float* cloud, int n, const float* feats, int m, int num_bins, float* depth_out, int* nn_out) {
  feature_depth((const P4*)cloud, n, feats, m, num_bins, depth_out, nn_out);
}

// Node outputs after optimation_processing (feature_tracker_node.cpp:388-401, :445-446): q_estimator from the rotation
// matrix, the relative pose published on /Odometry, then last <- current.  rt12 = row-major R, then t.
void orc_node_outputs(const double* rt12, double* last7, double* rel7, double* path7) {
  M3 R;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R.m[i][j] = rt12[3 * i + j];
  const Quat q = mat2q(R);
  const V3 t{rt12[9], rt12[10], rt12[11]};
  const Quat ql{last7[0], last7[1], last7[2], last7[3]};
  const V3 tl{last7[4], last7[5], last7[6]};
  const double n2 = ql.x * ql.x + ql.y * ql.y + ql.z * ql.z + ql.w * ql.w;  // Quaternion::inverse(): conjugate / squaredNorm
  Quat qi{0, 0, 0, 0};
  if (n2 > 0) qi = {-ql.x / n2, -ql.y / n2, -ql.z / n2, ql.w / n2};
  const Quat qr = qmul(qi, q);
  const V3 tr = qrot(qi, t - tl);
  rel7[0] = qr.x; rel7[1] = qr.y; rel7[2] = qr.z; rel7[3] = qr.w; rel7[4] = tr.x; rel7[5] = tr.y; rel7[6] = tr.z;
  path7[0] = q.x; path7[1] = q.y; path7[2] = q.z; path7[3] = q.w; path7[4] = t.x; path7[5] = t.y; path7[6] = t.z;
  last7[0] = q.x; last7[1] = q.y; last7[2] = q.z; last7[3] = q.w; last7[4] = t.x; last7[5] = t.y; last7[6] = t.z;
}


// ---- ring-field / range-image extractor (featureExtract.hpp) ----
struct orc_ri_params { int n_scan, horizon_scan, downsample_rate, pad_; double lidar_min, lidar_max, edge_threshold, surf_threshold; };
void orc_ri_default_params(orc_ri_params* p) {
  RIParams P;
  p->n_scan = P.n_scan; p->horizon_scan = P.horizon_scan; p->downsample_rate = P.downsample_rate; p->pad_ = 0;
  p->lidar_min = P.lidar_min; p->lidar_max = P.lidar_max; p->edge_threshold = P.edge_threshold; p->surf_threshold = P.surf_threshold;
}
// edge / surf: [n][4] capacity n each; dbg_* optional, capacity n each (semantic-cloud order), ring_se optional [2 * n_scan]
void orc_ri_extract(const orc_ri_params* p, const float* xyzi, const uint16_t* ring, int n, float* edge, int* edge_src, int* n_edge, float* surf, int* surf_src,
                    int* n_surf, int* dbg_src, int* dbg_col, float* dbg_range, float* dbg_curv, int* dbg_picked, int* n_sem, int* ring_se) {
  RIParams P;
  P.n_scan = p->n_scan; P.horizon_scan = p->horizon_scan; P.downsample_rate = p->downsample_rate; P.lidar_min = p->lidar_min; P.lidar_max = p->lidar_max;
  P.edge_threshold = p->edge_threshold; P.surf_threshold = p->surf_threshold;
  Cloud e, s;
  std::vector<int> es, ss;
  RIDebug D;
  ri_extract(P, (const P4*)xyzi, ring, n, e, es, s, ss, &D);
  *n_edge = (int)e.size(); *n_surf = (int)s.size();
  if (!e.empty()) { std::memcpy(edge, e.data(), e.size() * sizeof(P4)); std::memcpy(edge_src, es.data(), es.size() * sizeof(int)); }
  if (!s.empty()) { std::memcpy(surf, s.data(), s.size() * sizeof(P4)); std::memcpy(surf_src, ss.data(), ss.size() * sizeof(int)); }
  const size_t m = D.src.size();
  if (n_sem) *n_sem = (int)m;
  if (m) {
    if (dbg_src) std::memcpy(dbg_src, D.src.data(), m * sizeof(int));
    if (dbg_col) std::memcpy(dbg_col, D.col.data(), m * sizeof(int));
    if (dbg_range) std::memcpy(dbg_range, D.range.data(), m * sizeof(float));
    if (dbg_curv) std::memcpy(dbg_curv, D.curvature.data(), m * sizeof(float));
    if (dbg_picked) std::memcpy(dbg_picked, D.picked.data(), m * sizeof(int));
  }
  if (ring_se) for (int i = 0; i < P.n_scan; ++i) { ring_se[2 * i] = D.start[i]; ring_se[2 * i + 1] = D.end[i]; }
}

// ---- ScanContext (Scancontext.h) ----
struct orc_sc_params { double lidar_height; int num_ring, num_sector; double max_radius; int num_exclude_recent, num_candidates; double search_ratio, dist_thres; int tree_making_period, pad_; };
static SCParams to_sc(const orc_sc_params* p) {
  SCParams P;
  if (p) { P.lidar_height = p->lidar_height; P.num_ring = p->num_ring; P.num_sector = p->num_sector; P.max_radius = p->max_radius;
           P.num_exclude_recent = p->num_exclude_recent; P.num_candidates = p->num_candidates; P.search_ratio = p->search_ratio;
           P.dist_thres = p->dist_thres; P.tree_making_period = p->tree_making_period; }
  return P;
}
void orc_sc_default_params(orc_sc_params* p) {
  SCParams P;
  p->lidar_height = P.lidar_height; p->num_ring = P.num_ring; p->num_sector = P.num_sector; p->max_radius = P.max_radius;
  p->num_exclude_recent = P.num_exclude_recent; p->num_candidates = P.num_candidates; p->search_ratio = P.search_ratio;
  p->dist_thres = P.dist_thres; p->tree_making_period = P.tree_making_period; p->pad_ = 0;
}
void orc_sc_make(const orc_sc_params* p, const float* pts, int n, double* desc, double* ringkey, double* sectorkey) {
  SCParams P = to_sc(p);
  sc_make(P, pts, n, desc);
  if (ringkey) sc_ringkey(P, desc, ringkey);
  if (sectorkey) sc_sectorkey(P, desc, sectorkey);
}
void orc_sc_distance(const orc_sc_params* p, const double* sc1, const double* sc2, double* dist, int* shift) {
  auto r = sc_distance(to_sc(p), sc1, sc2);
  *dist = r.first; *shift = r.second;
}
void* orc_sc_create(const orc_sc_params* p) { SCManager* m = new SCManager(); m->P = to_sc(p); return m; }
void orc_sc_destroy(void* h) { delete (SCManager*)h; }
void orc_sc_add(void* h, const float* pts, int n) { ((SCManager*)h)->add(pts, n); }
void orc_sc_detect(void* h, int* loop_id, float* yaw, double* min_dist, int* nn_idx) {
  auto r = ((SCManager*)h)->detect(min_dist, nn_idx);
  *loop_id = r.first; *yaw = r.second;
}
float orc_sc_key_dist(const float* a, const float* b, int dim) { return sc_key_dist(a, b, dim); }

}  // extern "C"
