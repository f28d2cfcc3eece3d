// TEST INFRASTRUCTURE — CPU oracle for the lidar-odometry hot path.  NOT product code.
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use anything under oracle/.
//
// PARITY UNPINNED: the reference (RichExplor/VIL_Fusion) ships no tests, golden vectors or
// fixtures, and cannot be built here (needs ROS/PCL/Ceres/Eigen).  This file restates, in plain
// C++14 double/float arithmetic, the small pieces of Eigen 3.3.7 (README.md:29 of the reference)
// that the hot path relies on.  Each function cites the call site it serves.
#pragma once
#include <cmath>
#include <limits>
#include <cstring>
#include <algorithm>

namespace orc {

struct V3 { double x, y, z; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }

struct M3 { double m[3][3]; };  // row-major m[r][c]
inline M3 mul(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
  return r;
}
inline V3 mul(const M3& a, V3 v) {
  return {a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z, a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
          a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z};
}
inline M3 transpose(const M3& a) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i];
  return r;
}
inline M3 identity3() { return {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}}; }

// common.h:124-135
inline M3 skew(V3 v) { return {{{0, -v.z, v.y}, {v.z, 0, -v.x}, {-v.y, v.x, 0}}}; }

struct Quat { double x, y, z, w; };  // Eigen coefficient order (x,y,z,w) == parameter_opti[0..3], EstimationMapping.hpp:383

// Eigen QuaternionBase::operator* (quat product), used by LocalSE3Parameterization::Plus, EstimationMapping.hpp:45
inline Quat qmul(Quat a, Quat b) {
  return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
          a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
// Eigen QuaternionBase::_transformVector: v + w*uv + vec x uv with uv = 2*(vec x v).
// Call sites: EstimationMapping.hpp:358, lidarFactor.hpp:26, :83.
inline V3 qrot(Quat q, V3 v) {
  V3 qv{q.x, q.y, q.z};
  V3 uv = cross(qv, v);
  uv = uv + uv;
  return v + q.w * uv + cross(qv, uv);
}
// Eigen QuaternionBase::toRotationMatrix, EstimationMapping.hpp:292, common.h:165 (q.matrix()).
inline M3 qmat(Quat q) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  return {{{1 - (tyy + tzz), txy - twz, txz + twy}, {txy + twz, 1 - (txx + tzz), tyz - twx}, {txz - twy, tyz + twx, 1 - (txx + tyy)}}};
}
// Eigen Quaternion(Matrix3) (quaternionbase_assign_impl<...,3,3>), EstimationMapping.hpp:242.
inline Quat mat2q(const M3& a) {
  double q[4];  // x y z w
  double t = a.m[0][0] + a.m[1][1] + a.m[2][2];
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (a.m[2][1] - a.m[1][2]) * t;
    q[1] = (a.m[0][2] - a.m[2][0]) * t;
    q[2] = (a.m[1][0] - a.m[0][1]) * t;
  } else {
    int i = 0;
    if (a.m[1][1] > a.m[0][0]) i = 1;
    if (a.m[2][2] > a.m[i][i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(a.m[i][i] - a.m[j][j] - a.m[k][k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (a.m[k][j] - a.m[j][k]) * t;
    q[j] = (a.m[j][i] + a.m[i][j]) * t;
    q[k] = (a.m[k][i] + a.m[i][k]) * t;
  }
  return {q[0], q[1], q[2], q[3]};
}

// common.h:137-176  se(3) exponential used by LocalSE3Parameterization::Plus.
inline void se3_exp(const double se3[6], Quat& q, V3& t) {
  V3 omega{se3[0], se3[1], se3[2]}, upsilon{se3[3], se3[4], se3[5]};
  M3 Om = skew(omega);
  double theta = norm(omega);
  double half = 0.5 * theta;
  double imag, real = std::cos(half);
  if (theta < 1e-10) {
    double t2 = theta * theta, t4 = t2 * t2;
    imag = 0.5 - 0.0208333 * t2 + 0.000260417 * t4;
  } else {
    imag = std::sin(half) / theta;
  }
  q = {imag * omega.x, imag * omega.y, imag * omega.z, real};
  M3 J;
  if (theta < 1e-10) {
    J = qmat(q);
  } else {
    M3 Om2 = mul(Om, Om);
    double a = (1 - std::cos(theta)) / (theta * theta);
    double b = (theta - std::sin(theta)) / (std::pow(theta, 3));
    M3 I = identity3();
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) J.m[i][j] = I.m[i][j] + a * Om.m[i][j] + b * Om2.m[i][j];
  }
  t = mul(J, upsilon);
}

// LocalSE3Parameterization::Plus, EstimationMapping.hpp:34-49.  x = {qx,qy,qz,qw,tx,ty,tz}.
inline void se3_plus(const double x[7], const double delta[6], double out[7]) {
  Quat dq;
  V3 dt;
  se3_exp(delta, dq, dt);
  Quat q{x[0], x[1], x[2], x[3]};
  V3 t{x[4], x[5], x[6]};
  Quat qp = qmul(dq, q);
  V3 tp = qrot(dq, t) + dt;
  out[0] = qp.x; out[1] = qp.y; out[2] = qp.z; out[3] = qp.w;
  out[4] = tp.x; out[5] = tp.y; out[6] = tp.z;
}

// Isometry3d as (R, t).  Product / inverse follow Eigen::Transform<double,3,Isometry>
// (EstimationMapping.hpp:238): (A*B).R = A.R*B.R, (A*B).t = A.R*B.t + A.t; inv: R^T, -(R^T t).
struct Iso { M3 R; V3 t; };
inline Iso iso_identity() { return {identity3(), {0, 0, 0}}; }
inline Iso iso_mul(const Iso& a, const Iso& b) { return {mul(a.R, b.R), mul(a.R, b.t) + a.t}; }
inline Iso iso_inv(const Iso& a) {
  M3 rt = transpose(a.R);
  V3 v = mul(rt, a.t);
  return {rt, {-v.x, -v.y, -v.z}};
}

// Symmetric 3x3 eigen-decomposition standing in for Eigen::SelfAdjointEigenSolver<Matrix3d>
// (EstimationMapping.hpp:150): cyclic Jacobi to ~1e-15, eigenvalues ascending, eigenvectors in
// the columns of V (V[r][c]).  Eigen's own iteration (tridiagonal QL) is not restated: both are
// backward-stable, results agree to a few ulp of ||C||; eigenvector sign is immaterial.
inline void eig3_sym(const M3& C, double w[3], M3& V) {
  double a[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) a[i][j] = C.m[i][j];
  V = identity3();
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (off <= 1e-34 * diag || off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {  // A <- A * G
          double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {  // A <- G^T * A
          double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          double vkp = V.m[k][p], vkq = V.m[k][q];
          V.m[k][p] = c * vkp - s * vkq;
          V.m[k][q] = s * vkp + c * vkq;
        }
      }
  }
  w[0] = a[0][0]; w[1] = a[1][1]; w[2] = a[2][2];
  int idx[3] = {0, 1, 2};
  std::sort(idx, idx + 3, [&](int i, int j) { return w[i] < w[j]; });
  double ws[3] = {w[idx[0]], w[idx[1]], w[idx[2]]};
  M3 Vs;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Vs.m[r][c] = V.m[r][idx[c]];
  std::memcpy(w, ws, sizeof(ws));
  V = Vs;
}

// ALTERNATE of eig3_sym for the sensitivity study (tools/sensitivity.py; Config.eig_alg = 1): the algorithm Eigen 3.3.7's
// SelfAdjointEigenSolver<Matrix3d>::compute() actually runs (Eigen/src/Eigenvalues/SelfAdjointEigenSolver.h, Tridiagonalization.h,
// Jacobi.h — NOT under /root/reference; restated from the published source): scale by the largest |coefficient| of the lower
// triangle, the closed-form 3x3 Householder tridiagonalisation (tridiagonalization_inplace_selector<MatrixType, 3, false>),
// implicit symmetric QR steps with Wilkinson shift (tridiagonal_qr_step) until every sub-diagonal entry is negligible against its
// diagonal neighbours (2 * eps), at most 30 * n iterations, eigenvalues sorted ascending by selection, scaled back.
inline void eig3_sym_eigen(const M3& C, double w[3], M3& V) {
  double m[3][3];  // lower triangle of C
  double scale = 0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j <= i; ++j) { m[i][j] = C.m[i][j]; scale = std::max(scale, std::fabs(C.m[i][j])); }
  if (scale == 0) scale = 1;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j <= i; ++j) m[i][j] /= scale;
  double diag[3], sub[2];
  double Q[3][3];
  {
    const double tol = std::numeric_limits<double>::min();
    diag[0] = m[0][0];
    const double v1norm2 = m[2][0] * m[2][0];
    if (v1norm2 <= tol) {
      diag[1] = m[1][1]; diag[2] = m[2][2]; sub[0] = m[1][0]; sub[1] = m[2][1];
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Q[i][j] = i == j ? 1.0 : 0.0;
    } else {
      const double beta = std::sqrt(m[1][0] * m[1][0] + v1norm2);
      const double invBeta = 1.0 / beta;
      const double m01 = m[1][0] * invBeta, m02 = m[2][0] * invBeta;
      const double q = 2.0 * m01 * m[2][1] + m02 * (m[2][2] - m[1][1]);
      diag[1] = m[1][1] + m02 * q;
      diag[2] = m[2][2] - m02 * q;
      sub[0] = beta;
      sub[1] = m[2][1] - m01 * q;
      const double Qi[3][3] = {{1, 0, 0}, {0, m01, m02}, {0, m02, -m01}};
      std::memcpy(Q, Qi, sizeof(Q));
    }
  }
  const int n = 3;
  int end = n - 1, start = 0, iter = 0;
  const double considerAsZero = std::numeric_limits<double>::min();
  const double precision = 2.0 * std::numeric_limits<double>::epsilon();
  while (end > 0) {
    for (int i = start; i < end; ++i)
      if (std::fabs(sub[i]) <= (std::fabs(diag[i]) + std::fabs(diag[i + 1])) * precision || std::fabs(sub[i]) <= considerAsZero) sub[i] = 0;
    while (end > 0 && sub[end - 1] == 0.0) end--;
    if (end <= 0) break;
    iter++;
    if (iter > 30 * n) break;
    start = end - 1;
    while (start > 0 && sub[start - 1] != 0.0) start--;
    // tridiagonal_qr_step
    const double td = (diag[end - 1] - diag[end]) * 0.5;
    const double e = sub[end - 1];
    double mu = diag[end];
    if (td == 0.0) mu -= std::fabs(e);
    else {
      const double e2 = e * e;
      double h;  // numext::hypot
      {
        const double ax = std::fabs(td), ay = std::fabs(e);
        double p, qp;
        if (ax > ay) { p = ax; qp = ay / p; } else { p = ay; qp = ax / p; }
        h = p == 0.0 ? 0.0 : p * std::sqrt(1.0 + qp * qp);
      }
      if (e2 == 0.0) mu -= (e / (td + (td > 0.0 ? 1.0 : -1.0))) * (e / h);
      else mu -= e2 / (td + (td > 0.0 ? h : -h));
    }
    double x = diag[start] - mu;
    double z = sub[start];
    for (int k = start; k < end; ++k) {
      double c, s;  // JacobiRotation::makeGivens(x, z)
      if (z == 0.0) { c = x < 0.0 ? -1.0 : 1.0; s = 0.0; }
      else if (x == 0.0) { c = 0.0; s = z < 0.0 ? 1.0 : -1.0; }
      else if (std::fabs(x) > std::fabs(z)) { const double t = z / x; double u = std::sqrt(1.0 + t * t); if (x < 0.0) u = -u; c = 1.0 / u; s = -t * c; }
      else { const double t = x / z; double u = std::sqrt(1.0 + t * t); if (z < 0.0) u = -u; s = -1.0 / u; c = -t * s; }
      const double sdk = s * diag[k] + c * sub[k];
      const double dkp1 = s * sub[k] + c * diag[k + 1];
      diag[k] = c * (c * diag[k] - s * sub[k]) - s * (c * sub[k] - s * diag[k + 1]);
      diag[k + 1] = s * sdk + c * dkp1;
      sub[k] = c * sdk - s * dkp1;
      if (k > start) sub[k - 1] = c * sub[k - 1] - s * z;
      x = sub[k];
      if (k < end - 1) { z = -s * sub[k + 1]; sub[k + 1] = c * sub[k + 1]; }
      for (int i = 0; i < 3; ++i) {  // Q = Q * G: applyOnTheRight(k, k + 1, rot)
        const double xi = Q[i][k], yi = Q[i][k + 1];
        Q[i][k] = c * xi - s * yi;
        Q[i][k + 1] = s * xi + c * yi;
      }
    }
  }
  for (int i = 0; i < n - 1; ++i) {  // selection sort, ascending
    int k = i;
    for (int j = i + 1; j < n; ++j) if (diag[j] < diag[k]) k = j;
    if (k > i) { std::swap(diag[i], diag[k]); for (int r = 0; r < 3; ++r) std::swap(Q[r][i], Q[r][k]); }
  }
  for (int i = 0; i < 3; ++i) { w[i] = diag[i] * scale; for (int j = 0; j < 3; ++j) V.m[i][j] = Q[i][j]; }
}

// 5x3 least squares  min ||A n - b||  standing in for Eigen's colPivHouseholderQr().solve()
// (EstimationMapping.hpp:198): Householder QR with column pivoting on the largest remaining
// column norm (recomputed, not down-dated), then back substitution and un-pivoting.
inline V3 lstsq5x3_colpiv(const double Ain[5][3], const double bin[5], bool pivot = true) {
  double A[5][3], b[5];
  std::memcpy(A, Ain, sizeof(A));
  std::memcpy(b, bin, sizeof(b));
  int perm[3] = {0, 1, 2};
  // Eigen's rank rule (ColPivHouseholderQR::computeInPlace): pivots stop counting once the largest
  // remaining squared column norm < (eps * max initial column norm)^2 / rows * (rows - k).
  double maxcol = 0;
  for (int j = 0; j < 3; ++j) {
    double s = 0;
    for (int i = 0; i < 5; ++i) s += A[i][j] * A[i][j];
    maxcol = std::max(maxcol, std::sqrt(s));
  }
  const double thr = (maxcol * 2.220446049250313e-16) * (maxcol * 2.220446049250313e-16) / 5.0;
  int npiv = 3;
  for (int k = 0; k < 3; ++k) {
    int best = k;
    double bn = -1;
    for (int j = k; j < 3; ++j) {
      double s = 0;
      for (int i = k; i < 5; ++i) s += A[i][j] * A[i][j];
      if (s > bn) { bn = s; best = j; }
      if (!pivot) break;  // sensitivity alternate: Eigen's plain HouseholderQR (no column exchange)
    }
    if (npiv == 3 && bn < thr * (5 - k)) { npiv = k; break; }
    if (best != k) {
      for (int i = 0; i < 5; ++i) std::swap(A[i][k], A[i][best]);
      std::swap(perm[k], perm[best]);
    }
    // Householder vector for column k, rows k..4  (Eigen makeHouseholder convention)
    double tail = 0;
    for (int i = k + 1; i < 5; ++i) tail += A[i][k] * A[i][k];
    double c0 = A[k][k], beta, tau, v[5];
    if (tail <= 2.2250738585072014e-308) {
      tau = 0; beta = c0;
      for (int i = k + 1; i < 5; ++i) v[i] = 0;
    } else {
      beta = std::sqrt(c0 * c0 + tail);
      if (c0 >= 0) beta = -beta;
      for (int i = k + 1; i < 5; ++i) v[i] = A[i][k] / (c0 - beta);
      tau = (beta - c0) / beta;
    }
    A[k][k] = beta;
    for (int i = k + 1; i < 5; ++i) A[i][k] = 0;
    for (int j = k + 1; j < 3; ++j) {  // apply H = I - tau [1;v][1;v]^T to remaining columns
      double tmp = A[k][j];
      for (int i = k + 1; i < 5; ++i) tmp += v[i] * A[i][j];
      A[k][j] -= tau * tmp;
      for (int i = k + 1; i < 5; ++i) A[i][j] -= tau * v[i] * tmp;
    }
    double tmp = b[k];
    for (int i = k + 1; i < 5; ++i) tmp += v[i] * b[i];
    b[k] -= tau * tmp;
    for (int i = k + 1; i < 5; ++i) b[i] -= tau * v[i] * tmp;
  }
  double y[3] = {0, 0, 0};
  for (int k = npiv - 1; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < npiv; ++j) s -= A[k][j] * y[j];
    y[k] = s / A[k][k];
  }
  double n[3];
  for (int k = 0; k < 3; ++k) n[perm[k]] = y[k];
  return {n[0], n[1], n[2]};
}

}  // namespace orc
