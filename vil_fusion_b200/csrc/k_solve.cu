// Stage 3/4 — the 6-DoF pose solve: ceres::Solve(TRUST_REGION, LEVENBERG_MARQUARDT, DENSE_QR,
// max_num_iterations = 4, HuberLoss(0.1), LocalSE3Parameterization) of EM:263-283 as ONE kernel per outer
// iteration: one thread-block CLUSTER of 8 CTAs per sequence.  Every CTA evaluates its share of the factors
// (LF:21-52 edge, LF:79-102 surf) and reduces its robustified normal equations (21-entry J^T J, 6-entry
// J^T r, cost) with warp shuffles + a fixed-order shared-memory tree; the leader CTA gathers the 8 partials
// over distributed shared memory (fixed order -> deterministic) and its thread 0 runs Ceres' trust-region
// bookkeeping (Jacobi scaling fixed at iteration 0, LM diagonal clamp, radius update, parameter / function /
// gradient tolerance exits, step rejection) on the 6x6 system -- accept / reject of the evaluated candidate and the
// next step in ONE serial section, with the projected-gradient norm computed beside it by a second thread, so an
// LM iteration costs two cluster barriers; the candidate pose is broadcast back through DSMEM, so the pose never
// leaves the SMs between iterations.  (The FP64 work of one evaluation, ~6 k factors
// x ~300 instructions, is what bounds this kernel: one SM took 27 us per evaluation, 8 SMs take ~4.)
//
// Ceres solves [J S; D] y = [r; 0] by QR; here the same minimiser is obtained from the normal equations
// (S J^T J S + D^2) y = S J^T r by Cholesky.  The two differ by rounding only (relative 1e-10 on a step of
// 1e-2), far inside the pose tolerance (1e-4 rad / 1e-3 m); SURVEY §8a row 9 lists the schedule reproduced.
#include "vilf_internal.cuh"
#include <cooperative_groups.h>
#ifdef VILF_LM_TIMING
#include <cstdio>
#endif

namespace vilf {

namespace {

// -DVILF_LM_TIMING: thread 0 of the first sequence's leader CTA stamps clock64() at the phase boundaries and prints the
// deltas at frame 20 (tools/dev_lmts.py).  Development only.
#ifdef VILF_LM_TIMING
__shared__ long long ts_buf[128];
__shared__ int ts_n;
#define TS() do { if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && ts_n < 128) ts_buf[ts_n++] = clock64(); } while (0)
#else
#define TS() do { } while (0)
#endif

constexpr int NACC = 30;  // 21 H + 6 g + cost + n_edge + n_surf

struct LmShared {
  double x[7], cand[7], params[7];
  double grad_spec;  // projected-gradient max-norm at the point just evaluated (adopted with the step)
  double H[21], g[6], cost;
  double acc[NACC];               // cluster-wide sums (leader only)
  double parts[LM_CLUSTER_MAX][NACC];  // leader only: the partial sums every CTA of the cluster stores here over DSMEM
  double scale[6], diag[6];
  double radius, decrease_factor, minimum_cost, x_norm, grad_max, model_cost_change, candidate_cost;
  int iteration, step_successful, reuse_diagonal, num_invalid, done, need_eval, termination, n_edge, n_surf, n_rows;
};

__device__ __forceinline__ int hidx(int i, int j) {  // upper triangle, row-major: (i <= j)
  return i * 6 - i * (i - 1) / 2 + (j - i);
}

// common.h:137-176 se(3) exponential + LocalSE3Parameterization::Plus (EM:34-49).
// The reference evaluates sin(theta/2)/theta, cos(theta/2), (1-cos theta)/theta^2 and (theta-sin theta)/theta^3 through
// sqrt, sin, cos and three divisions.  A trust-region step is a rotation of a few milliradians, so for theta < 0.05 the
// same four functions are summed as power series in theta^2 (through theta^8: truncation < 4e-18), which needs neither
// theta itself nor a transcendental -- one thread runs this between two cluster barriers, and it was most of that wait.
// (The series is the more accurate of the two for the last coefficient, which the closed form gets by cancellation.)
__device__ void se3_plus(const double* x, const double* delta, double* out) {
  const D3 omega = d3(delta[0], delta[1], delta[2]), ups = d3(delta[3], delta[4], delta[5]);
  const double t2 = dot3(omega, omega);  // theta^2
  double imag, real, ca = 0, cb = 0;
  const bool tiny = t2 < 1e-20;  // theta < 1e-10 (CM:150, :163)
  if (tiny) {
    const double t4 = dmul(t2, t2);
    imag = dadd(dsub(0.5, dmul(0.0208333, t2)), dmul(0.000260417, t4));
    real = 1.0;  // cos(theta / 2) rounds to 1 below 1e-10
  } else if (t2 < 2.5e-3) {
    imag = fma(t2, fma(t2, fma(t2, fma(t2, 1.0 / 185794560.0, -1.0 / 645120.0), 1.0 / 3840.0), -1.0 / 48.0), 0.5);
    real = fma(t2, fma(t2, fma(t2, fma(t2, 1.0 / 10321920.0, -1.0 / 46080.0), 1.0 / 384.0), -1.0 / 8.0), 1.0);
    ca = fma(t2, fma(t2, fma(t2, fma(t2, 1.0 / 3628800.0, -1.0 / 40320.0), 1.0 / 720.0), -1.0 / 24.0), 0.5);
    cb = fma(t2, fma(t2, fma(t2, fma(t2, 1.0 / 39916800.0, -1.0 / 362880.0), 1.0 / 5040.0), -1.0 / 120.0), 1.0 / 6.0);
  } else {
    const double theta = sqrt(t2);
    double sin_half, sin_t, cos_t;
    sincos(dmul(0.5, theta), &sin_half, &real);
    sincos(theta, &sin_t, &cos_t);
    imag = sin_half / theta;
    ca = dsub(1, cos_t) / t2;
    cb = dsub(theta, sin_t) / dmul(t2, theta);  // pow(theta, 3) in CM:170; the product differs by <= 1 ulp
  }
  Q4 dq; dq.x = dmul(imag, omega.x); dq.y = dmul(imag, omega.y); dq.z = dmul(imag, omega.z); dq.w = real;
  double J[3][3];
  if (tiny) {  // J = q.matrix()
    const double tx = dmul(2, dq.x), ty = dmul(2, dq.y), tz = dmul(2, dq.z);
    const double twx = dmul(tx, dq.w), twy = dmul(ty, dq.w), twz = dmul(tz, dq.w);
    const double txx = dmul(tx, dq.x), txy = dmul(ty, dq.x), txz = dmul(tz, dq.x);
    const double tyy = dmul(ty, dq.y), tyz = dmul(tz, dq.y), tzz = dmul(tz, dq.z);
    J[0][0] = dsub(1, dadd(tyy, tzz)); J[0][1] = dsub(txy, twz); J[0][2] = dadd(txz, twy);
    J[1][0] = dadd(txy, twz); J[1][1] = dsub(1, dadd(txx, tzz)); J[1][2] = dsub(tyz, twx);
    J[2][0] = dsub(txz, twy); J[2][1] = dadd(tyz, twx); J[2][2] = dsub(1, dadd(txx, tyy));
  } else {  // V = I + ca * Omega + cb * Omega^2
    const double Om[3][3] = {{0, -omega.z, omega.y}, {omega.z, 0, -omega.x}, {-omega.y, omega.x, 0}};
    double Om2[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Om2[i][j] = dadd(dadd(dmul(Om[i][0], Om[0][j]), dmul(Om[i][1], Om[1][j])), dmul(Om[i][2], Om[2][j]));
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) J[i][j] = dadd(dadd(i == j ? 1.0 : 0.0, dmul(ca, Om[i][j])), dmul(cb, Om2[i][j]));
  }
  const D3 dt = d3(dadd(dadd(dmul(J[0][0], ups.x), dmul(J[0][1], ups.y)), dmul(J[0][2], ups.z)),
                   dadd(dadd(dmul(J[1][0], ups.x), dmul(J[1][1], ups.y)), dmul(J[1][2], ups.z)),
                   dadd(dadd(dmul(J[2][0], ups.x), dmul(J[2][1], ups.y)), dmul(J[2][2], ups.z)));
  Q4 q; q.x = x[0]; q.y = x[1]; q.z = x[2]; q.w = x[3];
  // Eigen quaternion product dq * q
  const double px = dsub(dadd(dadd(dmul(dq.w, q.x), dmul(dq.x, q.w)), dmul(dq.y, q.z)), dmul(dq.z, q.y));
  const double py = dsub(dadd(dadd(dmul(dq.w, q.y), dmul(dq.y, q.w)), dmul(dq.z, q.x)), dmul(dq.x, q.z));
  const double pz = dsub(dadd(dadd(dmul(dq.w, q.z), dmul(dq.z, q.w)), dmul(dq.x, q.y)), dmul(dq.y, q.x));
  const double pw = dsub(dsub(dsub(dmul(dq.w, q.w), dmul(dq.x, q.x)), dmul(dq.y, q.y)), dmul(dq.z, q.z));
  const D3 tp = qrot(dq, d3(x[4], x[5], x[6])) + dt;
  out[0] = px; out[1] = py; out[2] = pz; out[3] = pw;
  out[4] = tp.x; out[5] = tp.y; out[6] = tp.z;
}

__device__ __forceinline__ void huber(double s, double a, double& rho0, double& sq) {
  const double b = dmul(a, a);
  if (s > b) {
    const double r = sqrt(s);
    rho0 = dsub(dmul(dmul(2, a), r), b);
    sq = sqrt(fmax(DBL_MIN, a / r));  // Corrector with rho'' <= 0: residual and Jacobian scaled by sqrt(rho')
  } else {
    rho0 = s;
    sq = 1.0;
  }
}

// acc += [J^T J (upper triangle), J^T r] of one residual row.  Z = index of a Jacobian entry that is structurally zero
// (-1: none): its products are +-0 and are skipped.  The sums are FMA-accumulated: the reference has no counterpart of
// this summation (Ceres factorises J itself), only its value matters.
template <int Z>
__device__ __forceinline__ void accumulate_row(double* acc, const double* J, double r) {
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    if (i == Z) continue;
#pragma unroll
    for (int j = i; j < 6; ++j)
      if (j != Z) acc[hidx(i, j)] = fma(J[i], J[j], acc[hidx(i, j)]);
    acc[21 + i] = fma(J[i], r, acc[21 + i]);
  }
}

// One line factor (LF:21-52) / one plane factor (LF:79-102) at pose (q, t) into the thread's partial sums.
//
// The Jacobians are the reference's products -skew(a-b) * [-skew(lp) | I] / |a-b| (LF:39-47) and n^T * [-skew(pw) | I]
// (LF:90-97) with the structural zeros of the two skew matrices and of the identity block multiplied out by hand: the
// terms left are the same roundings in the same order (x * 0 + y == y, (-x) * (-y) == x * y), 24 instead of 126 FP64
// operations for a line factor and 9 instead of 30 for a plane factor.  inv_abn = 1 / |a - b| does not depend on the
// pose (r and J multiply by it where LF:31-33, :47 divide: <= 1 ulp apart).
__device__ __forceinline__ void edge_factor(double* acc, const Q4& q, const D3& t, const D3& p, const D3& a, const D3& b, double inv_abn, double hub) {
  const D3 lp = qrot(q, p) + t;                    // LF:26
  const D3 nu = cross3(lp - a, lp - b);            // LF:28
  const D3 ab = a - b;
  double r[3] = {dmul(nu.x, inv_abn), dmul(nu.y, inv_abn), dmul(nu.z, inv_abn)};
  const double pxx = dmul(ab.x, lp.x), pyy = dmul(ab.y, lp.y), pzz = dmul(ab.z, lp.z);
  const double tx = dmul(ab.x, inv_abn), ty = dmul(ab.y, inv_abn), tz = dmul(ab.z, inv_abn);
  double J[3][6] = {
      {dmul(-dadd(pzz, pyy), inv_abn), dmul(dmul(ab.y, lp.x), inv_abn), dmul(dmul(ab.z, lp.x), inv_abn), 0.0, tz, -ty},
      {dmul(dmul(ab.x, lp.y), inv_abn), dmul(-dadd(pzz, pxx), inv_abn), dmul(dmul(ab.z, lp.y), inv_abn), -tz, 0.0, tx},
      {dmul(dmul(ab.x, lp.z), inv_abn), dmul(dmul(ab.y, lp.z), inv_abn), dmul(-dadd(pyy, pxx), inv_abn), ty, -tx, 0.0}};  // LF:47
  double rho0, sq;
  huber(dadd(dadd(dmul(r[0], r[0]), dmul(r[1], r[1])), dmul(r[2], r[2])), hub, rho0, sq);
  acc[27] += 0.5 * rho0;
  acc[28] += 1.0;
  if (sq != 1.0) {  // outliers only; x * 1.0 == x
#pragma unroll
    for (int ii = 0; ii < 3; ++ii) {
#pragma unroll
      for (int jj = 0; jj < 6; ++jj) J[ii][jj] = dmul(J[ii][jj], sq);
      r[ii] = dmul(r[ii], sq);
    }
  }
  accumulate_row<3>(acc, J[0], r[0]);
  accumulate_row<4>(acc, J[1], r[1]);
  accumulate_row<5>(acc, J[2], r[2]);
}
__device__ __forceinline__ void surf_factor(double* acc, const Q4& q, const D3& t, const D3& p, const D3& n, double d, double hub) {
  const D3 pw = qrot(q, p) + t;                 // LF:83
  double r = dadd(dot3(n, pw), d);              // LF:84
  double J[6] = {dsub(dmul(n.z, pw.y), dmul(n.y, pw.z)), dsub(dmul(n.x, pw.z), dmul(n.z, pw.x)), dsub(dmul(n.y, pw.x), dmul(n.x, pw.y)), n.x, n.y, n.z};  // LF:97
  double rho0, sq;
  huber(dmul(r, r), hub, rho0, sq);
  acc[27] += 0.5 * rho0;
  acc[29] += 1.0;
  if (sq != 1.0) {
#pragma unroll
    for (int jj = 0; jj < 6; ++jj) J[jj] = dmul(J[jj], sq);
    r = dmul(r, sq);
  }
  accumulate_row<-1>(acc, J, r);
}

// The factors do not change between the <= 5 evaluations of one solve, and only about half of the feature slots hold an
// accepted factor.  Each CTA therefore copies the accepted factors of its slots (slot i belongs to CTA (i / LM_THREADS)
// % LM_CLUSTER) ONCE into a shared-memory pool -- line records {p, a, b, 1/|a-b|} from the front, plane records {p, n, d}
// from the back, compacted in slot order by a block-wide scan, so the order of summation is a function of the input only.
// A solve whose share does not fit the pool (very dense scans) evaluates that CTA from global memory instead.
constexpr int POOL_DOUBLES = 5100;
constexpr int EDGE_REC = 10, SURF_REC = 7;
struct Stage {
  int n_edge, n_surf, staged;
};

__device__ void stage_factors(const LaneDev& L, int ne, int ns, int rank, int ncl, double* pool, int* wsum, Stage& st) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int ce = 0, cs = 0;  // CTA-uniform running counts
  bool fits = true;
  for (int base = rank * LM_THREADS; base < ne + ns; base += ncl * LM_THREADS) {
    const int i = base + tid;
    const bool is_e = i < ne, is_s = !is_e && i < ne + ns;
    double rec[9];
    bool v = false;
    if (is_e) {
      const double* f = L.edge_pab + (size_t)i * 9;
#pragma unroll
      for (int c = 0; c < 9; ++c) rec[c] = f[c];
      v = L.fvalid[0][i] != 0;
    } else if (is_s) {
      const double* f = L.surf_pnd + (size_t)(i - ne) * 7;
#pragma unroll
      for (int c = 0; c < 7; ++c) rec[c] = f[c];
      v = L.fvalid[1][i - ne] != 0;
    }
    const unsigned be = __ballot_sync(0xffffffffu, v && is_e), bs = __ballot_sync(0xffffffffu, v && is_s);
    if (lane == 0) wsum[warp] = __popc(be) | (__popc(bs) << 16);
    __syncthreads();
    int oe = ce, os = cs, te = 0, ts = 0;
#pragma unroll
    for (int w = 0; w < LM_THREADS / 32; ++w) {
      const int x = wsum[w];
      if (w < warp) { oe += x & 0xffff; os += x >> 16; }
      te += x & 0xffff; ts += x >> 16;
    }
    __syncthreads();  // wsum is rewritten in the next round
    if (EDGE_REC * (ce + te) + SURF_REC * (cs + ts) > POOL_DOUBLES) { fits = false; break; }
    const unsigned below = (1u << lane) - 1u;
    if (v && is_e) {
      double* o = pool + (size_t)EDGE_REC * (oe + __popc(be & below));
#pragma unroll
      for (int c = 0; c < 9; ++c) o[c] = rec[c];
      o[9] = 1.0 / norm3(d3(rec[3], rec[4], rec[5]) - d3(rec[6], rec[7], rec[8]));
    } else if (v && is_s) {
      double* o = pool + POOL_DOUBLES - (size_t)SURF_REC * (os + __popc(bs & below) + 1);
#pragma unroll
      for (int c = 0; c < 7; ++c) o[c] = rec[c];
    }
    ce += te; cs += ts;
  }
  if (tid == 0) { st.n_edge = ce; st.n_surf = cs; st.staged = fits ? 1 : 0; }
  __syncthreads();
}

// Robustified cost / gradient / normal matrix of every valid factor at pose x; result in S.acc (all threads sync).
// Partial sums of this CTA into lead_parts (the leader's parts[rank], over DSMEM); the caller's cluster barrier follows.
__device__ void evaluate(const LaneDev& L, int ne, int ns, const double* x, double hub, double* lead_parts, double (*wred)[NACC], const double* pool, const Stage& st, int first, int stride) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = 0;
  Q4 q; q.x = x[0]; q.y = x[1]; q.z = x[2]; q.w = x[3];
  const D3 t = d3(x[4], x[5], x[6]);
  if (st.staged) {
    // line factors to the low threads, plane factors from the high threads down: the two kinds share no warp until the
    // CTA holds more than LM_THREADS factors, and the tails of both land on different warps
    for (int j = tid; j < st.n_edge; j += LM_THREADS) {
      const double* f = pool + (size_t)EDGE_REC * j;
      edge_factor(acc, q, t, d3(f[0], f[1], f[2]), d3(f[3], f[4], f[5]), d3(f[6], f[7], f[8]), f[9], hub);
    }
    for (int j = LM_THREADS - 1 - tid; j < st.n_surf; j += LM_THREADS) {
      const double* f = pool + POOL_DOUBLES - (size_t)SURF_REC * (j + 1);
      surf_factor(acc, q, t, d3(f[0], f[1], f[2]), d3(f[3], f[4], f[5]), f[6], hub);
    }
  } else {
    for (int i = first; i < ne + ns; i += stride) {
      if (i < ne) {
        // the record is read before the validity flag is tested (every slot < ne exists): one L2 round trip, not two
        const double* f = L.edge_pab + (size_t)i * 9;
        const D3 p = d3(f[0], f[1], f[2]), a = d3(f[3], f[4], f[5]), b = d3(f[6], f[7], f[8]);
        if (!L.fvalid[0][i]) continue;
        edge_factor(acc, q, t, p, a, b, 1.0 / norm3(a - b), hub);
      } else {
        const int k = i - ne;
        const double* f = L.surf_pnd + (size_t)k * 7;
        const D3 p = d3(f[0], f[1], f[2]), n = d3(f[3], f[4], f[5]);
        const double f6 = f[6];
        if (!L.fvalid[1][k]) continue;
        surf_factor(acc, q, t, p, n, f6, hub);
      }
    }
  }
  TS();
  // Warp reduction of the 30 sums by transposition: in the round with lane distance o every lane keeps one half of its
  // current values and adds the partner's copies of that half, so 16 + 8 + 4 + 2 + 1 = 31 exchanges replace 30 x 5, and
  // lane l ends up with the warp total of sum number rev5(l) (fixed order: deterministic).
  double a32[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) a32[i] = i < NACC ? acc[i] : 0.0;
#pragma unroll
  for (int o = 16, n = 32; o > 0; o >>= 1, n >>= 1) {
    const bool hi = (lane & o) != 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j < n / 2) {
        const double send = hi ? a32[j] : a32[j + n / 2];
        const double keep = hi ? a32[j + n / 2] : a32[j];
        a32[j] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
  }
  {  // lane l holds sum index: bit 4 of l selected the half of 32, bit 3 the half of 16, ...
    const int idx = ((lane >> 4) & 1) * 16 + ((lane >> 3) & 1) * 8 + ((lane >> 2) & 1) * 4 + ((lane >> 1) & 1) * 2 + (lane & 1);
    if (idx < NACC) wred[warp][idx] = a32[0];
  }
  __syncthreads();
  if (tid < NACC) {  // this CTA's sums go straight into the leader's shared memory; the cluster barrier that follows publishes them
    double v = 0;
    for (int w = 0; w < LM_THREADS / 32; ++w) v += wred[w][tid];
    lead_parts[tid] = v;
  }
  TS();
}

__device__ double norm7(const double* v) {
  double s = 0;
  for (int i = 0; i < 7; ++i) s = dadd(s, dmul(v[i], v[i]));
  return sqrt(s);
}

__device__ void record(SolveTraceDev& T, LmShared& S, int valid, int succ, double rel, double stepn) {
  if (S.n_rows >= MAX_TRACE_ROWS) return;  // the row count lives in shared memory: no global round trip per row
  LmRow& r = T.rows[S.n_rows++];
  r.iteration = S.iteration; r.step_valid = valid; r.step_successful = succ; r.cost = S.cost; r.candidate_cost = S.candidate_cost;
  r.model_cost_change = S.model_cost_change; r.relative_decrease = rel; r.radius = S.radius; r.step_norm = stepn;
  for (int i = 0; i < 7; ++i) r.x[i] = S.x[i];
}

// Take the freshly evaluated normal equations as the linearisation at S.x (EvaluateGradientAndJacobian).
__device__ void adopt_linearisation(LmShared& S, bool first) {
  for (int i = 0; i < 21; ++i) S.H[i] = S.acc[i];
  for (int i = 0; i < 6; ++i) S.g[i] = S.acc[21 + i];
  S.cost = S.acc[27];
  if (first)
    for (int j = 0; j < 6; ++j) S.scale[j] = 1.0 / (1.0 + sqrt(S.H[hidx(j, j)]));  // Jacobi scaling, fixed at iteration 0
}

// Max-norm of the projected gradient x - Plus(x, -g) (gradient tolerance test) for the gradient in S.acc: evaluated by a
// second thread while thread 0 does the trust-region bookkeeping, adopted only if the step is.
// The value is only ever compared with 1e-10, and its quaternion part alone decides almost every time: q - dq * q has
// 2-norm |1 - dq| = 2 |sin(theta / 4)| (theta = |g_rot|, |q| = 1), so the max-norm is at least |sin(theta / 4)|.  Whenever
// that exceeds 1e-9 the function returns 1.0 ("not converged") without the full exponential, which with a gradient of
// tens of radians is the slow path of sin / cos and used to be what the leader's thread 0 waited for.
__device__ double projected_gradient_max(const double* x, const double* g) {
  const double t2 = g[0] * g[0] + g[1] * g[1] + g[2] * g[2];
  if (t2 > 1e-16 && t2 < 36.0) return 1.0;  // theta / 4 in (2.5e-9, 1.5): sin >= 2.5e-9
  if (t2 >= 36.0 && fabs(sin(0.25 * sqrt(t2))) > 1e-9) return 1.0;
  double ng[6], proj[7];
  for (int j = 0; j < 6; ++j) ng[j] = -g[j];
  se3_plus(x, ng, proj);
  double gm = 0;
  for (int i = 0; i < 7; ++i) gm = fmax(gm, fabs(x[i] - proj[i]));
  return gm;
}

// 1 / sqrt(d) from the single-precision estimate and two Newton steps (full double precision for d in float range): this
// sits six times on the serial chain of the factorisation below.
__device__ __forceinline__ double rsqrt_newton(double d) {
  if (!(d > 1e-30 && d < 1e30)) return rsqrt(d);
  double r = (double)rsqrtf((float)d);
  const double h = 0.5 * d;
  r = fma(r, fma(-h * r, r, 0.5), r);
  r = fma(r, fma(-h * r, r, 0.5), r);
  return r;
}

// (S J^T J S + D^2) y = S g by Cholesky; step = -y.  Returns false when the system is not positive definite.
// model_cost_change = -step^T (S g + S H S step / 2) = (y^T S g + y^T D^2 y) / 2 because y solves the system above.
// One thread runs this between two cluster barriers: everything is unrolled into registers and fused.
__device__ bool lm_step(const LmShared& S, double* step, double& model_cost_change) {
  double A[6][6], rhs[6], d2[6];
  const double inv_radius = 1.0 / S.radius;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int j = i; j < 6; ++j) A[j][i] = S.H[hidx(i, j)] * S.scale[i] * S.scale[j];  // lower triangle
    d2[i] = S.diag[i] * inv_radius;  // LevenbergMarquardtStrategy: lm_diagonal = sqrt(diagonal / radius), appended to J, i.e. squared here
    A[i][i] += d2[i];
    rhs[i] = S.g[i] * S.scale[i];
  }
  double inv[6];  // Cholesky in place (A becomes L), one reciprocal square root per column
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double d = A[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) d = fma(-A[j][k], A[j][k], d);
    ok = ok && d > 0;
    inv[j] = rsqrt_newton(d);
    A[j][j] = d * inv[j];
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double sacc = A[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) sacc = fma(-A[i][k], A[j][k], sacc);
      A[i][j] = sacc * inv[j];
    }
  }
  if (!ok) return false;
  double z[6], y[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double sacc = rhs[i];
#pragma unroll
    for (int k = 0; k < i; ++k) sacc = fma(-A[i][k], z[k], sacc);
    z[i] = sacc * inv[i];
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double sacc = z[i];
#pragma unroll
    for (int k = i + 1; k < 6; ++k) sacc = fma(-A[k][i], y[k], sacc);
    y[i] = sacc * inv[i];
  }
  bool finite = true;
  double m = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    step[i] = -y[i];
    finite = finite && isfinite(y[i]);
    m = fma(y[i], fma(d2[i], y[i], rhs[i]), m);
  }
  model_cost_change = 0.5 * m;
  return finite;
}

}  // namespace

// Launched with a cluster of cfg.lm_cluster CTAs along x (8, or 16 for configurations with tens of thousands of factors per solve).
__global__ void __launch_bounds__(LM_THREADS, LM_THREADS <= 256 ? 2 : 1)
k_solve(LaneDev* lanes, int lane0, int outer, int finalize, ConfigDev cfg, int max_iters) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int ncl = (int)cluster.num_blocks();
  const LaneDev& L = lanes[lane0 + blockIdx.y];
  LaneVars& V = *L.v;
  __shared__ LmShared S;
  __shared__ double wred[LM_THREADS / 32][NACC];
  __shared__ double pool[POOL_DOUBLES];
  __shared__ int wsum[LM_THREADS / 32];
  __shared__ Stage stage;
  LmShared* lead = cluster.map_shared_rank(&S, 0);  // the leader's state, visible to the whole cluster
  const int tid = threadIdx.x;
  const bool leader = rank == 0;
#ifdef VILF_LM_TIMING
  if (tid == 0) ts_n = 0;
  __syncthreads();
#endif
  const bool run = V.opt_ran != 0;  // EM:254 decided by the association kernel
  const int ne = V.n_ds[0], ns = V.n_ds[1];
  const int first = rank * LM_THREADS + tid, stride = ncl * LM_THREADS;
  SolveTraceDev& T = L.trace[outer];
  // leader: add the 8 partials in rank order
  auto gather = [&]() {
    if (leader && tid < NACC) {
      double v = 0;
      for (int r = 0; r < ncl; ++r) v += S.parts[r][tid];
      S.acc[tid] = v;
    }
    if (leader) __syncthreads();
  };
  // leader: hand the outcome of thread 0's serial section (finished? / the next candidate pose) to every CTA's own
  // shared memory, so nobody starts the next trip with a DSMEM round trip
  auto publish = [&]() {  // lanes 1 .. LM_CLUSTER-1 of the leader's warp 0, one destination CTA each
    LmShared* o = cluster.map_shared_rank(&S, tid);
    o->done = S.done;
    for (int i = 0; i < 7; ++i) o->cand[i] = S.cand[i];
  };
  // Thread 0 of the leader, between two cluster barriers: close the iteration that just ended
  // (FinalizeIterationAndCheckIfMinimizerCanContinue) and, unless the solve is over, compute the next trust-region step
  // and the candidate pose the whole cluster evaluates next (need_eval) -- or shrink the radius after an invalid step.
  auto plan_step = [&]() {
    S.need_eval = 0;
    if (S.step_successful && S.cost < S.minimum_cost) { S.minimum_cost = S.cost; for (int i = 0; i < 7; ++i) S.params[i] = S.x[i]; }
    if (S.iteration >= max_iters) { S.termination = 0; S.done = 1; return; }
    if (S.step_successful && S.grad_max <= 1e-10) { S.termination = 3; S.done = 1; return; }
    if (S.radius <= 1e-32) { S.termination = 5; S.done = 1; return; }
    ++S.iteration;
    TS();
    if (!S.reuse_diagonal)
      for (int j = 0; j < 6; ++j) S.diag[j] = fmin(fmax(S.H[hidx(j, j)] * S.scale[j] * S.scale[j], 1e-6), 1e32);
    double step[6], mcc = 0;
    TS();
    bool valid = lm_step(S, step, mcc);
    TS();
    S.reuse_diagonal = 1;
    if (valid) {
      S.model_cost_change = mcc;
      valid = mcc > 0.0;
    }
    if (!valid) {  // HandleInvalidStep
      if (++S.num_invalid >= 5) { S.termination = 5; S.done = 1; record(T, S, 0, 0, 0, 0); }
      else {
        S.radius = S.radius / S.decrease_factor; S.decrease_factor *= 2.0; S.reuse_diagonal = 1; S.step_successful = 0;
        record(T, S, 0, 0, 0, 0);
      }
    } else {
      S.num_invalid = 0;
      double delta[6];
      for (int j = 0; j < 6; ++j) delta[j] = step[j] * S.scale[j];
      se3_plus(S.x, delta, S.cand);
      TS();
      S.need_eval = 1;
    }
  };
  if (run) {
    if (leader && tid == 0) {
      for (int i = 0; i < 7; ++i) { S.x[i] = V.x[i]; S.params[i] = V.x[i]; }
      S.radius = 1e4; S.decrease_factor = 2.0; S.minimum_cost = DBL_MAX; S.reuse_diagonal = 0; S.num_invalid = 0;
      S.model_cost_change = 0; S.candidate_cost = 0; S.iteration = 0; S.step_successful = 1; S.done = 0; S.termination = 0;
      S.need_eval = 0;
      S.n_rows = 0;
    }
    TS();
    stage_factors(L, ne, ns, rank, ncl, pool, wsum, stage);
    TS();
    double xl[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) xl[i] = V.x[i];
    // One trip = one evaluation by the whole cluster (trip 0: the start pose, then the candidates) followed by the leader's
    // serial section; done / cand are written by the leader's thread 0 between cluster barriers and read right after one.
    for (int trip = 0;; ++trip) {
      if (trip > 0) {
        if (S.done) break;
#pragma unroll
        for (int i = 0; i < 7; ++i) xl[i] = S.cand[i];
      }
      evaluate(L, ne, ns, xl, cfg.huber, lead->parts[rank], wred, pool, stage, first, stride);
      cluster.sync();
      TS();
      gather();
      TS();
      if (leader) {
        if (tid == 32) S.grad_spec = projected_gradient_max(trip == 0 ? S.x : S.cand, S.acc + 21);
        int adopt = 0;
        if (tid == 0) {
          if (trip == 0) {
            S.n_edge = (int)S.acc[28]; S.n_surf = (int)S.acc[29];
            T.n_edge = S.n_edge; T.n_surf = S.n_surf;
            if (S.n_edge + S.n_surf == 0) {
              S.done = 1; S.termination = 4;
            } else {
              adopt = 1;
              S.x_norm = norm7(S.x);
              adopt_linearisation(S, true);
              for (int i = 0; i < 21; ++i) T.H0[i] = S.H[i];
              for (int i = 0; i < 6; ++i) T.g0[i] = S.g[i];
              T.cost0 = S.cost;
              record(T, S, 1, 1, 0, 0);
            }
          } else {
            S.candidate_cost = S.acc[27];
            double sn = 0;
            for (int i = 0; i < 7; ++i) sn += (S.x[i] - S.cand[i]) * (S.x[i] - S.cand[i]);
            sn = sqrt(sn);
            const double cost_change = S.cost - S.candidate_cost;
            if (sn <= 1e-8 * (S.x_norm + 1e-8)) { S.termination = 1; S.done = 1; record(T, S, 1, 0, 0, sn); }            // parameter tolerance
            else if (fabs(cost_change) <= 1e-6 * S.cost) { S.termination = 2; S.done = 1; record(T, S, 1, 0, 0, sn); }    // function tolerance
            else {
              const double rel = cost_change / S.model_cost_change;
              if (rel > 1e-3) {  // HandleSuccessfulStep
                adopt = 1;
                for (int i = 0; i < 7; ++i) S.x[i] = S.cand[i];
                S.x_norm = norm7(S.x);
                adopt_linearisation(S, false);
                S.step_successful = 1;
                { const double c = 2.0 * rel - 1.0; S.radius = S.radius / fmax(1.0 / 3.0, 1.0 - c * c * c); }
                S.radius = fmin(1e16, S.radius);
                S.decrease_factor = 2.0;
                S.reuse_diagonal = 0;
                record(T, S, 1, 1, rel, sn);
              } else {  // HandleUnsuccessfulStep
                S.step_successful = 0;
                S.radius = S.radius / S.decrease_factor; S.decrease_factor *= 2.0; S.reuse_diagonal = 1;
                record(T, S, 1, 0, rel, sn);
              }
            }
          }
        }
        TS();
        __syncthreads();  // grad_spec
        if (tid == 0 && !S.done) {
          if (adopt) S.grad_max = S.grad_spec;
          do plan_step(); while (!S.done && !S.need_eval);  // an invalid step shrinks the radius and is retried right here
        }
        if (tid < 32) {
          __syncwarp();
          if (tid >= 1 && tid < ncl) publish();
        }
        TS();
      }
      cluster.sync();
      TS();
    }
    if (leader && tid == 0) {
      for (int i = 0; i < 7; ++i) V.x[i] = S.params[i];
      T.termination = S.termination;
      T.n_rows = S.n_rows;
      T.final_cost = S.cost;
    }
  }
  if (finalize && leader && tid == 0) {  // EM:291-293: globalOdom from q_w_c / t_w_c (also when the optimisation was skipped)
    const double* x = run ? S.params : V.x;
    const double tx = dmul(2, x[0]), ty = dmul(2, x[1]), tz = dmul(2, x[2]);
    const double twx = dmul(tx, x[3]), twy = dmul(ty, x[3]), twz = dmul(tz, x[3]);
    const double txx = dmul(tx, x[0]), txy = dmul(ty, x[0]), txz = dmul(tz, x[0]);
    const double tyy = dmul(ty, x[1]), tyz = dmul(tz, x[1]), tzz = dmul(tz, x[2]);
    double* o = V.odom;
    o[0] = dsub(1, dadd(tyy, tzz)); o[1] = dsub(txy, twz); o[2] = dadd(txz, twy);
    o[3] = dadd(txy, twz); o[4] = dsub(1, dadd(txx, tzz)); o[5] = dsub(tyz, twx);
    o[6] = dsub(txz, twy); o[7] = dadd(tyz, twx); o[8] = dsub(1, dadd(txx, tyy));
    o[9] = x[4]; o[10] = x[5]; o[11] = x[6];
    V.frames += 1;
  }
#ifdef VILF_LM_TIMING
  if (leader && blockIdx.y == 0 && tid == 0 && V.frames == 20) {
    printf("LMTS outer %d n %d:", outer, ts_n);
    for (int i = 1; i < ts_n; ++i) printf(" %d", (int)(ts_buf[i] - ts_buf[i - 1]));
    printf("\n");
  }
#endif
  cluster.sync();  // no CTA may exit while the leader can still read its shared memory
}

cudaError_t init_solve_kernels() {  // clusters of 16 CTAs are beyond the portable size
  return cudaFuncSetAttribute(k_solve, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
}

void launch_solve(const Launch& L, LaneDev* lanes, int lane0, int nlanes, int outer, int finalize, const ConfigDev& cfg, int max_iters) {
  const int ncl = cfg.lm_cluster > 0 ? cfg.lm_cluster : LM_CLUSTER;
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(ncl, nlanes);
  lc.blockDim = dim3(LM_THREADS);
  lc.dynamicSmemBytes = 0;
  lc.stream = L.st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = ncl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  lc.attrs = at; lc.numAttrs = 1;
  cudaLaunchKernelEx(&lc, k_solve, lanes, lane0, outer, finalize, cfg, max_iters);
  L.tick(K_SOLVE);
}

}  // namespace vilf
