// On-device Levenberg-Marquardt replacing ceres::Solve at Aloam/src/laserMapping.cpp:713-720
// and Aloam/src/laserOdometry.cpp:494-499 (HuberLoss(0.1), EigenQuaternionParameterization,
// DENSE_QR, max_num_iterations = 4, Ceres 1.14 defaults otherwise).
//
// One launch = one evaluation of all factors at the pose under test: per factor the residual
// (Aloam/src/lidarFactor.hpp:19-43, 69-90, 114-125), its analytic 3x6 / 1x6 Jacobian in the
// tangent space of q+ = [sin|d| d/|d|, cos|d|] (x) q  (d lp / d delta = -2 [R p]x,
// d lp / d t = I), the Huber corrector sqrt(rho'), and the 28 unique doubles
// {J^T J (21), J^T r (6), cost} reduced warp-shuffle -> shared -> per-block partials; the
// last block to finish sums the partials in fixed order (run-to-run deterministic, no float
// atomics) and one thread advances the trust-region state machine: Jacobi scaling, LM
// diagonal, damped 6x6 Cholesky solve in double, step-quality test, radius update and the
// function / parameter / gradient tolerances -- no host round trip inside a solve.
#include "common.cuh"
#include <float.h>
#include <stdlib.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

constexpr int NRED = 28;          // 21 + 6 + 1
constexpr int EVAL_THREADS = 256;

__device__ __forceinline__ void d_cross(const double* a, const double* b, double* o) {
  o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}

// accumulate one residual row: J (6), r, into the 27 sums.  Explicit fp64 FMAs (the TU is built with
// -fmad=false for the fp32 parity paths): the sums are only compared at 1e-5 relative (north_star), and
// the fused form halves the fp64 issue slots of the hottest loop.
__device__ __forceinline__ void d_acc_row(const double* J, double r, double* acc) {
  int k = 0;
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int b = a; b < 6; ++b) { acc[k] = fma(J[a], J[b], acc[k]); ++k; }
#pragma unroll
  for (int a = 0; a < 6; ++a) acc[21 + a] = fma(J[a], r, acc[21 + a]);
}

// rotation matrix of a (unit) quaternion, Eigen's toRotationMatrix() operation order; computed once per pass and thread
__device__ __forceinline__ void d_rot_of_q(const double* q, double* R) {
  const double tx = 2.0 * q[0], ty = 2.0 * q[1], tz = 2.0 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0], tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0 - (txx + tyy);
}

// One factor at the pose (R, t).  COST_ONLY: only the robustified cost (a candidate point whose Jacobian cannot be used
// any more: the last evaluation of a solve).  Explicit fp64 FMAs in the rotation and the sums: the normal equations are
// compared at 1e-5 relative (north_star), not bit for bit.  The edge residual divides by |a - b| ONCE (one reciprocal
// square root instead of twelve IEEE divisions -- each a ~20-instruction software sequence on the hottest path).
template <bool COST_ONLY>
__device__ __forceinline__ void d_eval_factor(const LmFactor& f, const double* R, const double* t, double* acc) {
  if (f.kind < 0) return;
  const double px = (double)f.p[0], py = (double)f.p[1], pz = (double)f.p[2];
  const double rp[3] = { fma(R[0], px, fma(R[1], py, R[2] * pz)), fma(R[3], px, fma(R[4], py, R[5] * pz)), fma(R[6], px, fma(R[7], py, R[8] * pz)) };
  const double lp[3] = { rp[0] + t[0], rp[1] + t[1], rp[2] + t[2] };
  // d lp / d delta = -2 [R p]x  (columns), d lp / d t = I
  if (f.kind == 0) {
    const double da[3] = { lp[0] - f.a[0], lp[1] - f.a[1], lp[2] - f.a[2] };
    const double db[3] = { lp[0] - f.b[0], lp[1] - f.b[1], lp[2] - f.b[2] };
    double nu[3]; d_cross(da, db, nu);
    const double de[3] = { f.a[0] - f.b[0], f.a[1] - f.b[1], f.a[2] - f.b[2] };
    const double inv = rsqrt(fma(de[0], de[0], fma(de[1], de[1], de[2] * de[2])));
    const double r[3] = { nu[0] * inv, nu[1] * inv, nu[2] * inv };
    const double s = fma(r[0], r[0], fma(r[1], r[1], r[2] * r[2]));
    double rho0 = s, sr = 1.0;
    if (s > 0.01) { const double rs = sqrt(s); rho0 = 2.0 * 0.1 * rs - 0.01; sr = sqrt(fmax(DBL_MIN, 0.1 / rs)); }
    acc[27] += 0.5 * rho0;
    if (COST_ONLY) return;
    // dr/dlp = -[de]x / |de| = M (rows m_k); J_t = M, J_rot = M (-2 [rp]x): row k = -2 (m_k x rp)
    const double w = inv * sr;
    const double e0 = de[0] * w, e1 = de[1] * w, e2 = de[2] * w;     // rows of M, scaled by the corrector: (0, e2, -e1), (-e2, 0, e0), (e1, -e0, 0)
    {
      const double J[6] = { -2.0 * fma(e2, rp[2], e1 * rp[1]), 2.0 * (e1 * rp[0]), 2.0 * (e2 * rp[0]), 0.0, e2, -e1 };
      d_acc_row(J, r[0] * sr, acc);
    }
    {
      const double J[6] = { 2.0 * (e0 * rp[1]), -2.0 * fma(e0, rp[0], e2 * rp[2]), 2.0 * (e2 * rp[1]), -e2, 0.0, e0 };
      d_acc_row(J, r[1] * sr, acc);
    }
    {
      const double J[6] = { 2.0 * (e0 * rp[2]), 2.0 * (e1 * rp[2]), -2.0 * fma(e1, rp[1], e0 * rp[0]), e1, -e0, 0.0 };
      d_acc_row(J, r[2] * sr, acc);
    }
  } else {
    double r, n[3];
    if (f.kind == 1) { n[0] = f.b[0]; n[1] = f.b[1]; n[2] = f.b[2]; r = fma(lp[0] - f.a[0], n[0], fma(lp[1] - f.a[1], n[1], (lp[2] - f.a[2]) * n[2])); }
    else { n[0] = f.a[0]; n[1] = f.a[1]; n[2] = f.a[2]; r = fma(n[0], lp[0], fma(n[1], lp[1], n[2] * lp[2])) + f.b[0]; }
    const double s = r * r;
    double rho0 = s, sr = 1.0;
    if (s > 0.01) { const double rs = sqrt(s); rho0 = 2.0 * 0.1 * rs - 0.01; sr = sqrt(fmax(DBL_MIN, 0.1 / rs)); }
    acc[27] += 0.5 * rho0;
    if (COST_ONLY) return;
    double nxv[3]; d_cross(n, rp, nxv);
    const double m2 = -2.0 * sr;
    const double J[6] = { m2 * nxv[0], m2 * nxv[1], m2 * nxv[2], n[0] * sr, n[1] * sr, n[2] * sr };
    d_acc_row(J, r * sr, acc);
  }
}

// ---- DISTORTION 1: factors with an interpolation ratio s != 1 (lidarFactor.hpp:27-34, 73-79) ---------------------------
// q_last_curr = Identity.slerp(s, q), t_last_curr = s t, lp = q_last_curr * cp + t_last_curr.  The reference differentiates
// this with Ceres Jets through Eigen's slerp (acos / sin of the quaternion's w) and multiplies by the local
// parameterisation; here the same expressions are evaluated on forward-mode dual numbers seeded with d q / d delta
// (EigenQuaternionParameterization::ComputeJacobian) and d t / d t = I, which is the same derivative by the chain rule.
// Compile-time off in the reference, so this path is written for exactness, not speed.
struct J6 { double v; double d[6]; };
__device__ __forceinline__ J6 jc(double v) { J6 r; r.v = v; for (int k = 0; k < 6; ++k) r.d[k] = 0.0; return r; }
__device__ __forceinline__ J6 operator+(const J6& a, const J6& b) { J6 r; r.v = a.v + b.v; for (int k = 0; k < 6; ++k) r.d[k] = a.d[k] + b.d[k]; return r; }
__device__ __forceinline__ J6 operator-(const J6& a, const J6& b) { J6 r; r.v = a.v - b.v; for (int k = 0; k < 6; ++k) r.d[k] = a.d[k] - b.d[k]; return r; }
__device__ __forceinline__ J6 operator-(const J6& a) { J6 r; r.v = -a.v; for (int k = 0; k < 6; ++k) r.d[k] = -a.d[k]; return r; }
__device__ __forceinline__ J6 operator*(const J6& a, const J6& b) { J6 r; r.v = a.v * b.v; for (int k = 0; k < 6; ++k) r.d[k] = a.v * b.d[k] + a.d[k] * b.v; return r; }
__device__ __forceinline__ J6 operator*(double a, const J6& b) { J6 r; r.v = a * b.v; for (int k = 0; k < 6; ++k) r.d[k] = a * b.d[k]; return r; }
__device__ __forceinline__ J6 operator/(const J6& a, const J6& b) { J6 r; const double i = 1.0 / b.v; r.v = a.v * i; for (int k = 0; k < 6; ++k) r.d[k] = (a.d[k] - r.v * b.d[k]) * i; return r; }
__device__ __forceinline__ J6 jsin(const J6& a) { J6 r; const double c = cos(a.v); r.v = sin(a.v); for (int k = 0; k < 6; ++k) r.d[k] = c * a.d[k]; return r; }
__device__ __forceinline__ J6 jacos(const J6& a) { J6 r; const double g = -1.0 / sqrt(1.0 - a.v * a.v); r.v = acos(a.v); for (int k = 0; k < 6; ++k) r.d[k] = g * a.d[k]; return r; }
__device__ __forceinline__ void jcross(const J6* a, const J6* b, J6* o) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; }

__device__ __noinline__ void d_eval_factor_ratio(const LmFactor& f, const double* q, const double* t, double s, double* acc) {
  // seeds: rows of the 4x3 local Jacobian for q = (x, y, z, w), identity for t
  J6 Q[4], T[3];
  Q[0] = jc(q[0]); Q[0].d[0] = q[3];  Q[0].d[1] = q[2];  Q[0].d[2] = -q[1];
  Q[1] = jc(q[1]); Q[1].d[0] = -q[2]; Q[1].d[1] = q[3];  Q[1].d[2] = q[0];
  Q[2] = jc(q[2]); Q[2].d[0] = q[1];  Q[2].d[1] = -q[0]; Q[2].d[2] = q[3];
  Q[3] = jc(q[3]); Q[3].d[0] = -q[0]; Q[3].d[1] = -q[1]; Q[3].d[2] = -q[2];
  for (int k = 0; k < 3; ++k) { T[k] = jc(t[k]); T[k].d[3 + k] = 1.0; }
  // Eigen 3.3 slerp from the identity
  const J6 dq = Q[3];
  const J6 absD = dq.v < 0.0 ? -dq : dq;
  J6 scale0, scale1;
  if (absD.v >= 1.0 - DBL_EPSILON) { scale0 = jc(1.0 - s); scale1 = jc(s); }
  else {
    const J6 theta = jacos(absD), sinTheta = jsin(theta);
    scale0 = jsin((1.0 - s) * theta) / sinTheta;
    scale1 = jsin(s * theta) / sinTheta;
  }
  if (dq.v < 0.0) scale1 = -scale1;
  J6 u[3] = { scale1 * Q[0], scale1 * Q[1], scale1 * Q[2] };
  const J6 w = scale0 + scale1 * Q[3];
  // QuaternionBase::_transformVector: uv = u x v; uv += uv; v + w uv + u x uv
  const J6 v[3] = { jc((double)f.p[0]), jc((double)f.p[1]), jc((double)f.p[2]) };
  J6 uv[3]; jcross(u, v, uv);
  for (int k = 0; k < 3; ++k) uv[k] = uv[k] + uv[k];
  J6 uuv[3]; jcross(u, uv, uuv);
  J6 lp[3];
  for (int k = 0; k < 3; ++k) lp[k] = ((v[k] + w * uv[k]) + uuv[k]) + s * T[k];
  J6 r[3]; int nr;
  if (f.kind == 0) {
    J6 da[3], db[3], nu[3];
    for (int k = 0; k < 3; ++k) { da[k] = lp[k] - jc(f.a[k]); db[k] = lp[k] - jc(f.b[k]); }
    jcross(da, db, nu);
    const double de[3] = { f.a[0] - f.b[0], f.a[1] - f.b[1], f.a[2] - f.b[2] };
    const double inv = 1.0 / sqrt(de[0] * de[0] + de[1] * de[1] + de[2] * de[2]);
    for (int k = 0; k < 3; ++k) r[k] = inv * nu[k];
    nr = 3;
  } else {
    r[0] = (lp[0] - jc(f.a[0])) * jc(f.b[0]) + (lp[1] - jc(f.a[1])) * jc(f.b[1]) + (lp[2] - jc(f.a[2])) * jc(f.b[2]);
    nr = 1;
  }
  double sq = 0.0;
  for (int k = 0; k < nr; ++k) sq += r[k].v * r[k].v;
  double rho0 = sq, sr = 1.0;
  if (sq > 0.01) { const double rs = sqrt(sq); rho0 = 2.0 * 0.1 * rs - 0.01; sr = sqrt(fmax(DBL_MIN, 0.1 / rs)); }
  acc[27] += 0.5 * rho0;
  for (int k = 0; k < nr; ++k) {
    double J[6];
    for (int a = 0; a < 6; ++a) J[a] = r[k].d[a] * sr;
    d_acc_row(J, r[k].v * sr, acc);
  }
}

// ---- trust-region controller (single thread) -------------------------------------------
// EigenQuaternionParameterization::Plus: q+ = [sin|d| d/|d|, cos|d|] (x) q, t+ = t + dt.  sin(|d|)/|d| and cos(|d|) are
// even functions of |d|: for |d| < 0.25 rad (every trust-region step of a registration) both are evaluated as Horner
// polynomials in |d|^2 (truncation < 1e-22) -- no square root, no division, no sincos() call on the controller's
// dependent chain; larger steps take the general path.
__device__ void d_plus(const double* x, const double* d, double* out) {
  const double n2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  double sbd, cs;
  if (n2 < 0.0625) {
    // sin(a)/a = sum (-1)^k a^2k / (2k+1)!,  cos(a) = sum (-1)^k a^2k / (2k)!,  k = 0..8
    sbd = -1.0 / 355687428096000.0;        // 17!
    sbd = fma(sbd, n2, 1.0 / 1307674368000.0);       // 15!
    sbd = fma(sbd, n2, -1.0 / 6227020800.0);         // 13!
    sbd = fma(sbd, n2, 1.0 / 39916800.0);            // 11!
    sbd = fma(sbd, n2, -1.0 / 362880.0);             // 9!
    sbd = fma(sbd, n2, 1.0 / 5040.0);
    sbd = fma(sbd, n2, -1.0 / 120.0);
    sbd = fma(sbd, n2, 1.0 / 6.0);
    sbd = fma(-sbd, n2, 1.0);
    cs = 1.0 / 20922789888000.0;           // 16!
    cs = fma(cs, n2, -1.0 / 87178291200.0);          // 14!
    cs = fma(cs, n2, 1.0 / 479001600.0);             // 12!
    cs = fma(cs, n2, -1.0 / 3628800.0);              // 10!
    cs = fma(cs, n2, 1.0 / 40320.0);
    cs = fma(cs, n2, -1.0 / 720.0);
    cs = fma(cs, n2, 1.0 / 24.0);
    cs = fma(cs, n2, -0.5);
    cs = fma(cs, n2, 1.0);
  } else {
    const double nd = sqrt(n2);
    double sn;
    sincos(nd, &sn, &cs);
    sbd = sn / nd;
  }
  const double a[4] = { sbd * d[0], sbd * d[1], sbd * d[2], cs };
  d_qmul(a, x, out);
  out[4] = x[4] + d[3]; out[5] = x[5] + d[4]; out[6] = x[6] + d[5];
}
__device__ __forceinline__ double d_norm7(const double* v) {
  double s = 0;
#pragma unroll
  for (int i = 0; i < 7; ++i) s += v[i] * v[i];
  return sqrt(s);
}
// upper-triangular packed index of (a, b), a <= b; constant-folds when a, b are unrolled loop indices
__device__ __forceinline__ constexpr int d_tri(int a, int b) { return a <= b ? a * 6 - (a * (a - 1)) / 2 + (b - a) : b * 6 - (b * (b - 1)) / 2 + (a - b); }
__device__ __forceinline__ double d_Hat(const double* H, int a, int b) { return H[d_tri(a, b)]; }
// Ceres' gradient tolerance: max |x - Plus(x, -g)| <= 1e-10.  The translation block of that difference is |g_t| itself
// (up to an ulp of x): whenever a translation gradient exceeds 2e-10 the test has failed and the rotation block -- a
// sincos of a large angle on the controller's dependent chain -- need not be formed.
__device__ double d_gradient_max_norm(const double* x, const double* g) {
  const double gt = fmax(fmax(fabs(g[3]), fabs(g[4])), fabs(g[5]));
  if (gt > 2e-10) return gt;
  double ng[6], xp[7];
#pragma unroll
  for (int i = 0; i < 6; ++i) ng[i] = -g[i];
  d_plus(x, ng, xp);
  double m = 0.0;
#pragma unroll
  for (int i = 0; i < 7; ++i) { double a = fabs(x[i] - xp[i]); if (a > m) m = a; }
  return m;
}

// solve (A) y = b for SPD 6x6 A (upper-triangular packed) via Cholesky; returns 0 on success.
// Fully unrolled (every index is a compile-time constant, the factor lives in registers) and division-free:
// one reciprocal square root per pivot instead of a sqrt and 4-5 fp64 divisions (each a ~100-cycle dependent
// software sequence on the single controller thread).  The step differs from a divide-based Cholesky in the
// last ulps only; north_star compares poses at 1e-4 and normal equations at 1e-5.
__device__ __forceinline__ int d_chol6(const double* A /*21*/, const double* b, double* y) {
  // explicit fp64 FMAs (the TU is built with -fmad=false): every s -= a * b of the three dependent chains below is one
  // instruction instead of two
  double L[21];       // lower factor, L(i,j), j <= i, at d_tri(j, i)
  double inv[6];      // 1 / L(i,i)
  int bad = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
#pragma unroll
    for (int j = 0; j <= i; ++j) {
      double s = A[d_tri(j, i)];
#pragma unroll
      for (int k = 0; k < j; ++k) s = fma(-L[d_tri(k, i)], L[d_tri(k, j)], s);
      if (i == j) { if (!(s > 0.0)) bad = 1; inv[i] = rsqrt(s); L[d_tri(i, i)] = s * inv[i]; }
      else L[d_tri(j, i)] = s * inv[j];
    }
  }
  if (bad) return 1;
  double z[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s = fma(-L[d_tri(k, i)], z[k], s);
    z[i] = s * inv[i];
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double s = z[i];
#pragma unroll
    for (int k = i + 1; k < 6; ++k) s = fma(-L[d_tri(i, k)], y[k], s);
    y[i] = s * inv[i];
  }
  return 0;
}

// Computes the next trust-region step from (H, g) at x; on success sets lm->cand and returns 1.
// Returns 0 if the solve terminated (lm->done set).
__device__ int d_compute_step(LmLmState* lm) {
  double Hs[21], gs[6], sc[6];
#pragma unroll
  for (int a = 0; a < 6; ++a) sc[a] = lm->scaling[a];
#pragma unroll
  for (int a = 0; a < 6; ++a) {
    gs[a] = lm->g[a] * sc[a];
#pragma unroll
    for (int b = a; b < 6; ++b) Hs[d_tri(a, b)] = lm->H[d_tri(a, b)] * sc[a] * sc[b];
  }
  for (;;) {
    if (lm->iteration >= lm->max_iter) { lm->termination = 0; lm->done = 1; return 0; }
    if (!(lm->radius > 1e-32)) { lm->termination = 4; lm->done = 1; return 0; }
    lm->iteration++;
    if (!lm->reuse_diagonal) {
#pragma unroll
      for (int j = 0; j < 6; ++j) { double s = Hs[d_tri(j, j)]; lm->diagonal[j] = s < 1e-6 ? 1e-6 : (s > 1e32 ? 1e32 : s); }
    }
    double A[21];
#pragma unroll
    for (int k = 0; k < 21; ++k) A[k] = Hs[k];
    const double inv_radius = lm->inv_radius;
#pragma unroll
    for (int j = 0; j < 6; ++j) A[d_tri(j, j)] += lm->diagonal[j] * inv_radius;     // D^2 = diagonal / radius
    double y[6];
    int fail = d_chol6(A, gs, y);
#pragma unroll
    for (int j = 0; j < 6; ++j) if (!isfinite(y[j])) fail = 1;
    lm->reuse_diagonal = 1;
    int valid = 0;
    double step[6];
    if (!fail) {
#pragma unroll
      for (int j = 0; j < 6; ++j) step[j] = -y[j];
      // model_cost_change = -(J s)^T (r + J s / 2) = -(s^T g + s^T H s / 2)
      double sg = 0.0, sHs = 0.0;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        sg = fma(step[a], gs[a], sg);
        double row = 0.0;
#pragma unroll
        for (int b = 0; b < 6; ++b) row = fma(Hs[d_tri(a, b)], step[b], row);
        sHs = fma(step[a], row, sHs);
      }
      lm->model_cost_change = -(sg + 0.5 * sHs);
      lm->inv_model_cost_change = 1.0 / lm->model_cost_change;      // off the chain: overlaps the candidate's Plus below
      valid = lm->model_cost_change > 0.0;
    }
    if (!valid) {
      if (++lm->num_invalid >= 5) { lm->termination = 5; lm->done = 1; return 0; }
      lm->radius *= 0.5; lm->inv_radius *= 2.0; lm->reuse_diagonal = 1;
      continue;
    }
    lm->num_invalid = 0;
    double delta[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) delta[j] = step[j] * sc[j];
    d_plus(lm->x, delta, lm->cand);
    return 1;
  }
}

__device__ void d_lm_control(LmLmState* lm, const double* red /*28*/, const LmProblem& P, int write_back) {
  const double cost_e = red[27];
  if (lm->phase == 0) {
    // IterationZero
    lm->cost = cost_e; lm->initial_cost = cost_e;
    _Pragma("unroll") for (int k = 0; k < 21; ++k) lm->H[k] = red[k];
    _Pragma("unroll") for (int k = 0; k < 6; ++k) lm->g[k] = red[21 + k];
    _Pragma("unroll") for (int j = 0; j < 6; ++j) lm->scaling[j] = 1.0 / (1.0 + sqrt(d_Hat(lm->H, j, j)));
    lm->x_norm = d_norm7(lm->x);
    lm->phase = 1;
    if (d_gradient_max_norm(lm->x, lm->g) <= 1e-10) { lm->termination = 1; lm->done = 1; }
    else d_compute_step(lm);
  } else {
    // candidate evaluated
    const double* x = lm->x; const double* c = lm->cand;
    double dn = 0.0; for (int i = 0; i < 7; ++i) dn += (x[i] - c[i]) * (x[i] - c[i]);
    const double step_norm = sqrt(dn);
    if (step_norm <= 1e-8 * (lm->x_norm + 1e-8)) { lm->termination = 2; lm->done = 1; }
    else {
      const double cost_change = lm->cost - cost_e;
      if (fabs(cost_change) <= 1e-6 * lm->cost) { lm->termination = 3; lm->done = 1; }
      else {
        const double rd = cost_change * lm->inv_model_cost_change;
        if (rd > 1e-3) {
          _Pragma("unroll") for (int i = 0; i < 7; ++i) lm->x[i] = lm->cand[i];
          lm->x_norm = d_norm7(lm->x);
          lm->cost = cost_e;
          _Pragma("unroll") for (int k = 0; k < 21; ++k) lm->H[k] = red[k];
          _Pragma("unroll") for (int k = 0; k < 6; ++k) lm->g[k] = red[21 + k];
          const double tq = 2.0 * rd - 1.0;
          double den = 1.0 - tq * tq * tq; if (den < 1.0 / 3.0) den = 1.0 / 3.0;
          lm->inv_radius = lm->inv_radius * den;      // den in [1/3, 2]
          lm->radius = lm->radius / den;
          if (lm->radius > 1e16) { lm->radius = 1e16; lm->inv_radius = 1e-16; }
          lm->decrease_factor = 2.0; lm->reuse_diagonal = 0;
          lm->num_successful++;
          if (lm->iteration < lm->max_iter && d_gradient_max_norm(lm->x, lm->g) <= 1e-10) { lm->termination = 1; lm->done = 1; }
        } else {
          lm->radius = lm->radius / lm->decrease_factor; lm->inv_radius *= lm->decrease_factor;   // a power of two: exact
          lm->decrease_factor *= 2.0; lm->reuse_diagonal = 1;
        }
        if (!lm->done) d_compute_step(lm);
      }
    }
  }
  if (lm->done && write_back) {
    _Pragma("unroll") for (int k = 0; k < 4; ++k) P.pose_q[k] = lm->x[k];
    _Pragma("unroll") for (int k = 0; k < 3; ++k) P.pose_t[k] = lm->x[4 + k];
    LmSolveSummary* S = P.summary;
    S->iterations = lm->iteration; S->num_successful = lm->num_successful; S->termination = lm->termination;
    S->num_factors = lm->nfactors; S->initial_cost = lm->initial_cost; S->final_cost = lm->cost;
  }
}

// ---- kernels -----------------------------------------------------------------------------
__global__ void k_lm_begin(LmLmState* __restrict__ lm, LmProblem P, int max_iter, int sharded) {
  // counts the factors of this association pass (corner_num / surf_num, laserMapping.cpp:620,685;
  // corner_correspondence / plane_correspondence, laserOdometry.cpp:382,480) and arms the controller
  __shared__ int ws[33];
  const int gate = P.gate ? *P.gate : 1;
  const int n0 = *P.n0, n1 = *P.n1;
  int c0 = 0, c1 = 0;
  if (gate) {
    for (int i = threadIdx.x; i < n0; i += blockDim.x) c0 += P.fac0[i].kind >= 0;
    for (int i = threadIdx.x; i < n1; i += blockDim.x) c1 += P.fac1[i].kind >= 0;
  }
  int t0, t1;
  d_block_exscan(c0, ws, &t0);
  d_block_exscan(c1, ws, &t1);
  if (threadIdx.x == 0) {
    if (P.count0) *P.count0 = t0;
    if (P.count1) *P.count1 = t1;
    for (int k = 0; k < 4; ++k) lm->x[k] = P.pose_q[k];
    for (int k = 0; k < 3; ++k) lm->x[4 + k] = P.pose_t[k];
    for (int k = 0; k < 7; ++k) lm->cand[k] = lm->x[k];
    lm->radius = 1e4; lm->inv_radius = 1e-4; lm->inv_model_cost_change = 0.0; lm->decrease_factor = 2.0; lm->reuse_diagonal = 0; lm->num_invalid = 0;
    lm->iteration = 0; lm->num_successful = 0; lm->termination = 0; lm->phase = 0; lm->max_iter = max_iter;
    lm->nfactors = t0 + t1; lm->ticket = 0; lm->initial_cost = 0.0; lm->cost = 0.0;
    lm->done = (!gate || (!sharded && (t0 + t1) == 0)) ? 1 : 0;   // sharded: the global count decides (k_lm_control)
    if (lm->done && P.summary) {   // Ceres: no residual blocks -> parameters untouched
      LmSolveSummary* S = P.summary;
      S->iterations = 0; S->num_successful = 0; S->termination = 6; S->num_factors = 0; S->initial_cost = 0.0; S->final_cost = 0.0;
    }
  }
}

// shard_ws != NULL: cube-sharded map -- the last block publishes this rank's 28 partial sums (and its factor
// counts) to the workspace instead of running the controller; the host all-reduces the workspace over the
// ranks (NCCL) and k_lm_control advances the (replicated, deterministic) controller from the reduced sums.
__global__ void __launch_bounds__(EVAL_THREADS) k_lm_eval(LmLmState* __restrict__ lm, LmProblem P,
                                                          double* __restrict__ partials, int write_back,
                                                          double* __restrict__ shard_ws) {
  if (lm->done) return;
  __shared__ double sred[EVAL_THREADS / 32][NRED];
  __shared__ bool s_last;
  const int n0 = *P.n0, n1 = *P.n1;
  const LmFactor* __restrict__ fac0 = P.fac0; const LmFactor* __restrict__ fac1 = P.fac1;
  const double* xe = lm->phase == 0 ? lm->x : lm->cand;
  const double q[4] = { xe[0], xe[1], xe[2], xe[3] };
  const double t[3] = { xe[4], xe[5], xe[6] };
  double R[9]; d_rot_of_q(q, R);
  double acc[NRED];
#pragma unroll
  for (int k = 0; k < NRED; ++k) acc[k] = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n0 + n1; i += gridDim.x * blockDim.x) {
    const LmFactor f = i < n0 ? fac0[i] : fac1[i - n0];
    if (P.frac0 && f.kind >= 0 && f.kind <= 1) d_eval_factor_ratio(f, q, t, (double)(i < n0 ? P.frac0[i] : P.frac1[i - n0]) / 0.1, acc);
    else d_eval_factor<false>(f, R, t, acc);
  }
#pragma unroll
  for (int k = 0; k < NRED; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    acc[k] = v;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NRED; ++k) sred[wid][k] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x < NRED) {
    double v = 0.0;
    for (int w = 0; w < EVAL_THREADS / 32; ++w) v += sred[w][threadIdx.x];
    partials[(size_t)blockIdx.x * 32 + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int tk = atomicAdd(&lm->ticket, 1u);
    s_last = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  __shared__ double fin[NRED];
  if (threadIdx.x < NRED) {
    double v = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) v += partials[(size_t)b * 32 + threadIdx.x];   // fixed order
    fin[threadIdx.x] = v;
  }
  __syncthreads();
  if (shard_ws) {
    if (threadIdx.x < NRED) shard_ws[threadIdx.x] = fin[threadIdx.x];
    if (threadIdx.x == 0) {
      lm->ticket = 0;
      shard_ws[28] = P.count0 ? (double)*P.count0 : 0.0;
      shard_ws[29] = P.count1 ? (double)*P.count1 : 0.0;
    }
    return;
  }
  if (threadIdx.x == 0) {
    lm->ticket = 0;
    d_lm_control(lm, fin, P, write_back);
  }
}

// controller step from the all-reduced workspace (sharded map): identical on every rank
__global__ void k_lm_control(LmLmState* __restrict__ lm, LmProblem P, const double* __restrict__ ws, int write_back) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (lm->done) return;
  if (lm->phase == 0) {          // global factor counts replace the local ones of k_lm_begin
    const int c0 = (int)ws[28], c1 = (int)ws[29];
    if (P.count0) *P.count0 = c0;
    if (P.count1) *P.count1 = c1;
    lm->nfactors = c0 + c1;
    if (c0 + c1 == 0) {
      lm->done = 1;
      if (P.summary) { LmSolveSummary* S = P.summary; S->iterations = 0; S->num_successful = 0; S->termination = 6; S->num_factors = 0; S->initial_cost = 0.0; S->final_cost = 0.0; }
      return;
    }
  }
  double red[NRED];
  for (int k = 0; k < NRED; ++k) red[k] = ws[k];
  d_lm_control(lm, red, P, write_back);
}

// ---- whole solve in ONE launch: a thread-block cluster replaces the launch-per-evaluation loop --------
// The factors of one registration (~16 k x 64 B) are a few microseconds of fp64 work; what the
// launch-per-evaluation version paid for was 6 launches per solve, each with a grid-wide "last block"
// tail and a single-thread controller behind it.  Here LMC_CLUSTER CTAs of one cluster (co-scheduled on
// one GPC) keep everything on chip: every thread evaluates its factors (grid-stride over the cluster),
// warps reduce the 30-vector {J^T J (21), J^T r (6), cost, corner count, surf count} with a halving
// butterfly (31 double shuffles instead of 150), each CTA pushes its block sums into EVERY CTA's shared
// memory through distributed shared memory, one cluster barrier, and every CTA advances its own copy
// of the deterministic trust-region controller (identical inputs -> identical state), so the next
// evaluation starts without another exchange.  Sums are taken in a fixed order: run-to-run bit-stable.
constexpr int LMC_THREADS = 256;
constexpr int LMC_CLUSTER = 16;   // max cluster size (non-portable, opt-in); 8 is used where 16 cannot be scheduled

// factor cache in shared memory: four 16-byte planes so that consecutive threads touch consecutive words
constexpr int LMC_CACHE = 8 * LMC_THREADS;      // factors per CTA (128 KB): 32 k factors over a 16-CTA cluster
__device__ __forceinline__ void d_cache_store(uint4* c, int slot, const LmFactor& f) {
  const uint4* w = reinterpret_cast<const uint4*>(&f);
#pragma unroll
  for (int k = 0; k < 4; ++k) c[k * LMC_CACHE + slot] = w[k];
}
__device__ __forceinline__ LmFactor d_cache_load(const uint4* c, int slot) {
  LmFactor f;
  uint4* w = reinterpret_cast<uint4*>(&f);
#pragma unroll
  for (int k = 0; k < 4; ++k) w[k] = c[k * LMC_CACHE + slot];
  return f;
}

// after the call lane L holds the warp-wide sum of element L (L < 32) in v[0]
__device__ __forceinline__ void d_warp_transpose_reduce32(double (&v)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const double send = upper ? v[i] : v[i + half];
      const double keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
}

// RATIO: the problem carries per-factor interpolation ratios (DISTORTION 1); a separate instantiation, so that the
// dual-number path (a call with a large stack frame) costs the common kernel neither registers nor spills
template <bool RATIO>
__global__ void __launch_bounds__(LMC_THREADS, 1)
k_lm_solve_cluster(LmLmState* __restrict__ lm_g, LmProblem P, int max_iter, int write_back, unsigned long long* __restrict__ stamps,
                   const LmShardPeers* __restrict__ peers /*NULL: map not sharded over GPUs*/, uint32_t* __restrict__ fault) {
  // Distributed shared memory may only be written once the target CTA is known to be running: every CTA arrives on the
  // cluster barrier first thing and waits for the phase behind the state set-up below, so the round trip is hidden by the
  // predecessor wait and the set-up
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  lm_pdl_enter();
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int csize = (int)cluster.num_blocks();
  __shared__ double s_part[LMC_THREADS / 32][32];
  __shared__ double s_all[2][LMC_CLUSTER][32];      // [parity][source CTA][element], written remotely
  __shared__ double s_fin[32];
  __shared__ LmLmState s_lm;
  extern __shared__ uint4 s_cache[];                // [4][LMC_CACHE]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  LM_STAMP(stamps, 0);
  // cube-sharded map (peer-memory mode): after the cluster-wide sum every rank all-gathers the 30-vectors of all ranks
  // through the exchange blocks mapped over NVLink and adds them in rank order, so the replicated controllers of all
  // GPUs see the same bits -- the collective is part of this kernel, there is no host or NCCL call inside a solve
  const unsigned long long epoch0 = peers ? peers->peer[peers->rank]->epoch : 0ull;
  unsigned int n_xchg = 0;
  const int n0 = *P.n0, n1 = *P.n1;
  const LmFactor* __restrict__ fac0 = P.fac0; const LmFactor* __restrict__ fac1 = P.fac1;
  if (threadIdx.x == 0) {
    LmLmState* lm = &s_lm;
    const int gate = P.gate ? *P.gate : 1;
    for (int k = 0; k < 4; ++k) lm->x[k] = P.pose_q[k];
    for (int k = 0; k < 3; ++k) lm->x[4 + k] = P.pose_t[k];
    for (int k = 0; k < 7; ++k) lm->cand[k] = lm->x[k];
    for (int k = 0; k < 21; ++k) lm->H[k] = 0.0;
    for (int k = 0; k < 6; ++k) { lm->g[k] = 0.0; lm->scaling[k] = 1.0; lm->diagonal[k] = 0.0; }
    lm->x_norm = 0.0; lm->model_cost_change = 0.0;
    lm->radius = 1e4; lm->inv_radius = 1e-4; lm->inv_model_cost_change = 0.0; lm->decrease_factor = 2.0; lm->reuse_diagonal = 0; lm->num_invalid = 0;
    lm->iteration = 0; lm->num_successful = 0; lm->termination = 0; lm->phase = 0; lm->max_iter = max_iter;
    lm->nfactors = 0; lm->ticket = 0; lm->initial_cost = 0.0; lm->cost = 0.0; lm->pad = 0; lm->pad2 = 0;
    lm->done = gate ? 0 : 1;
    if (!gate && crank == 0) {
      if (P.count0) *P.count0 = 0;
      if (P.count1) *P.count1 = 0;
      if (P.summary) { LmSolveSummary* S = P.summary; S->iterations = 0; S->num_successful = 0; S->termination = 6; S->num_factors = 0; S->initial_cost = 0.0; S->final_cost = 0.0; }
    }
  }
  __syncthreads();
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
  const int wb = (write_back && crank == 0) ? 1 : 0;
  LM_STAMP(stamps, 1);
  for (int it = 0; it <= max_iter; ++it) {
    if (s_lm.done) break;                       // replicated state: uniform over the whole cluster
    const double* xe = s_lm.phase == 0 ? s_lm.x : s_lm.cand;
    const double q[4] = { xe[0], xe[1], xe[2], xe[3] };
    const double t[3] = { xe[4], xe[5], xe[6] };
    double R[9]; d_rot_of_q(q, R);
    // the Jacobian of the last candidate of a solve is never used (no step follows): cost only.  s_lm.iteration is the
    // number of steps computed so far, replicated in every CTA.
    const bool cost_only = s_lm.phase != 0 && s_lm.iteration >= max_iter;
    // The factors are dealt to LMC_CLUSTER x LMC_THREADS VIRTUAL threads whatever the real cluster size is: a CTA of a
    // smaller cluster evaluates its LMC_CLUSTER / csize virtual CTAs one after the other (own accumulators, own block
    // reduction, own slice of the factor cache).  The 30-vector is therefore summed in exactly the same order for every
    // cluster size, and a solve gives the same bits alone on the GPU (16 CTAs) and inside a batch (8 CTAs).
    const int nvirt = LMC_CLUSTER / csize;
    const int vcache = LMC_CACHE / nvirt;
    for (int v = 0; v < nvirt; ++v) {
    const int vrank = crank + v * csize;
    double acc[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) acc[k] = 0.0;
    {   // stride over the virtual cluster.  The first evaluation streams the factors from global memory (next
        // factor's 64 B in flight while the current one is evaluated) and parks them in shared memory; the
        // other evaluations of the solve re-read them from there instead of paying L2 latency per factor.
      const int stride = LMC_CLUSTER * LMC_THREADS, nf = n0 + n1;
      int i = vrank * LMC_THREADS + threadIdx.x;
      int slot = threadIdx.x;
      const int sbase = v * vcache;
      LmFactor f;
      const bool first = (it == 0);
      if (i < nf) { if (first || slot >= vcache) f = i < n0 ? fac0[i] : fac1[i - n0]; else f = d_cache_load(s_cache, sbase + slot); }
      while (i < nf) {
        const int inext = i + stride, snext = slot + LMC_THREADS;
        LmFactor fn;
        if (inext < nf) { if (first || snext >= vcache) fn = inext < n0 ? fac0[inext] : fac1[inext - n0]; else fn = d_cache_load(s_cache, sbase + snext); }
        if (first && slot < vcache) d_cache_store(s_cache, sbase + slot, f);
        if (f.kind >= 0) { if (i < n0) acc[28] += 1.0; else acc[29] += 1.0; }
        if (RATIO && f.kind >= 0 && f.kind <= 1) d_eval_factor_ratio(f, q, t, (double)(i < n0 ? P.frac0[i] : P.frac1[i - n0]) / 0.1, acc);
        else if (cost_only) d_eval_factor<true>(f, R, t, acc); else d_eval_factor<false>(f, R, t, acc);
        f = fn; i = inext; slot = snext;
      }
    }
    if (v == nvirt - 1) LM_STAMP(stamps, 8 + 8 * it);
    d_warp_transpose_reduce32(acc, lane);
    s_part[wid][lane] = acc[0];
    __syncthreads();
    if (threadIdx.x < 32) {
      double sum = 0.0;
#pragma unroll
      for (int w = 0; w < LMC_THREADS / 32; ++w) sum += s_part[w][threadIdx.x];
      double* mine = &s_all[it & 1][vrank][threadIdx.x];
      for (int r = 0; r < csize; ++r) *cluster.map_shared_rank(mine, r) = sum;
    }
    __syncthreads();            // s_part is reused by the next virtual CTA
    }
    LM_STAMP(stamps, 9 + 8 * it);
    cluster.sync();
    LM_STAMP(stamps, 10 + 8 * it);
    if (threadIdx.x < 32) {
      double v = 0.0;
      for (int r = 0; r < LMC_CLUSTER; ++r) v += s_all[it & 1][r][threadIdx.x];     // fixed order over the virtual CTAs
      s_fin[threadIdx.x] = v;
    }
    __syncthreads();
    if (peers) d_shard_exchange(peers, epoch0 + (++n_xchg), s_fin, s_fin, 32, crank == 0, fault);
    if (threadIdx.x == 0) {
      LmLmState* lm = &s_lm;
      if (lm->phase == 0) {
        const int c0 = (int)s_fin[28], c1 = (int)s_fin[29];
        lm->nfactors = c0 + c1;
        if (crank == 0) { if (P.count0) *P.count0 = c0; if (P.count1) *P.count1 = c1; }
        if (c0 + c1 == 0) {                     // Ceres: no residual blocks -> parameters untouched
          lm->done = 1;
          if (wb && P.summary) { LmSolveSummary* S = P.summary; S->iterations = 0; S->num_successful = 0; S->termination = 6; S->num_factors = 0; S->initial_cost = 0.0; S->final_cost = 0.0; }
        }
      }
      if (!lm->done) d_lm_control(lm, s_fin, P, wb);
    }
    LM_STAMP(stamps, 11 + 8 * it);
    __syncthreads();
  }
  LM_STAMP(stamps, 2);
  if (crank == 0 && threadIdx.x == 0) {
    *lm_g = s_lm;     // test hooks read H, g, cost from here
    if (peers) peers->peer[peers->rank]->epoch = epoch0 + n_xchg;
  }
  cluster.sync();                                        // nobody leaves while its shared memory may still be written
}

// test hook output: H (36), g (6), cost of the factors at q_w_curr/t_w_curr
__global__ void k_lm_export_normal_eq(const LmLmState* __restrict__ lm, double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) out[a * 6 + b] = d_Hat(lm->H, a, b);
  for (int a = 0; a < 6; ++a) out[36 + a] = lm->g[a];
  out[42] = lm->cost;
  out[43] = (double)lm->nfactors;
}

// LMONO_LM_MULTILAUNCH=1 keeps the launch-per-evaluation path (the one the sharded mode builds on) for A/B runs
static bool lm_use_launch_per_eval() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("LMONO_LM_MULTILAUNCH"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

constexpr int LMC_SMEM = 4 * LMC_CACHE * 16;
// largest cluster the device schedules for the solve kernel: 16 (opt-in, non-portable) if possible, else 8
static int lm_cluster_size(lmono_ctx* ctx) {
  static int cached[64] = { 0 };
  const int d = ctx->device & 63;
  if (cached[d]) return cached[d];
  int best = 8;
  if (cudaFuncSetAttribute(k_lm_solve_cluster<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LMC_SMEM) != cudaSuccess) return -1;
  if (cudaFuncSetAttribute(k_lm_solve_cluster<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LMC_SMEM) != cudaSuccess) return -1;
  cudaFuncSetAttribute(k_lm_solve_cluster<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  if (cudaFuncSetAttribute(k_lm_solve_cluster<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(LMC_CLUSTER); cfg.blockDim = dim3(LMC_THREADS); cfg.dynamicSmemBytes = LMC_SMEM;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = LMC_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k_lm_solve_cluster<false>, &cfg) == cudaSuccess && n >= 1 &&
        cudaOccupancyMaxActiveClusters(&n, k_lm_solve_cluster<true>, &cfg) == cudaSuccess && n >= 1) best = LMC_CLUSTER;
  }
  cudaGetLastError();
  cached[d] = best;
  return best;
}
// 16 CTAs minimise the latency of one solve; with several sequences side by side 8-CTA clusters leave room for the
// other sequences' solves (each CTA holds a whole SM's register file).  LMONO_LM_CLUSTER overrides.
static int lm_cluster_pick(lmono_ctx* ctx) {
  const int best = lm_cluster_size(ctx);
  if (best < 0) return best;
  static const int env = getenv("LMONO_LM_CLUSTER") ? atoi(getenv("LMONO_LM_CLUSTER")) : 0;
  int want = env >= 1 ? env : (ctx->batch_n > 8 ? 8 : best);      // 8 GPCs: at most eight 16-CTA clusters are resident at once
  want = want < best ? want : best;
  int p2 = 1; while (p2 * 2 <= want) p2 *= 2;      // the virtual-CTA scheme needs a divisor of LMC_CLUSTER
  return p2;
}

static int eval_blocks(lmono_ctx* ctx, int n) {
  int b = lm_div_up(n, EVAL_THREADS);
  if (b < 1) b = 1;
  if (b > 2 * ctx->sm_count) b = 2 * ctx->sm_count;
  return b;
}

int lm_solve_problem(lmono_ctx* ctx, const LmProblem& P, int n_max, int max_iter, int write_back, bool shard_exchange) {
  if (!lm_use_launch_per_eval() || shard_exchange) {
    const int cs = lm_cluster_pick(ctx);
    if (cs < 0) return LMONO_E_CUDA;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs); cfg.blockDim = dim3(LMC_THREADS); cfg.dynamicSmemBytes = LMC_SMEM; cfg.stream = ctx->stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = lm_pdl_on() ? 2 : 1;
    const LmShardPeers* peers = shard_exchange ? ctx->d_shard_peers : nullptr;
    if (P.frac0) LM_CUDA(cudaLaunchKernelEx(&cfg, k_lm_solve_cluster<true>, ctx->d_lm, P, max_iter, write_back, ctx->d_stamps, peers, &ctx->d_state->fault));
    else LM_CUDA(cudaLaunchKernelEx(&cfg, k_lm_solve_cluster<false>, ctx->d_lm, P, max_iter, write_back, ctx->d_stamps, peers, &ctx->d_state->fault));
    LM_LAUNCH_CHECK();
    return LMONO_OK;
  }
  k_lm_begin<<<1, 1024, 0, ctx->stream>>>(ctx->d_lm, P, max_iter, 0);
  LM_LAUNCH_CHECK();
  const int blocks = eval_blocks(ctx, n_max);
  for (int it = 0; it <= max_iter; ++it) {
    k_lm_eval<<<blocks, EVAL_THREADS, 0, ctx->stream>>>(ctx->d_lm, P, ctx->d_partials, write_back, nullptr);
    LM_LAUNCH_CHECK();
  }
  return LMONO_OK;
}

static LmProblem map_problem(lmono_ctx* ctx, int solve_index) {
  LmMapState* st = ctx->d_state;
  LmProblem P;
  P.fac0 = ctx->d_fac[0]; P.fac1 = ctx->d_fac[1];
  P.n0 = &st->stack_n[0]; P.n1 = &st->stack_n[1];
  P.gate = &st->optimize;
  P.pose_q = st->q_w_curr; P.pose_t = st->t_w_curr;
  P.summary = &st->solve[solve_index];
  P.count0 = &st->corner_num[solve_index]; P.count1 = &st->surf_num[solve_index];
  P.frac0 = nullptr; P.frac1 = nullptr;
  return P;
}

int lm_solve_enqueue(lmono_ctx* ctx, int solve_index, int n_max_corner, int n_max_surf, int max_iter) {
  return lm_solve_problem(ctx, map_problem(ctx, solve_index), n_max_corner + n_max_surf, max_iter, 1, ctx->shard_p2p);
}

int lm_normal_eq_enqueue(lmono_ctx* ctx, int n_max_corner, int n_max_surf) {
  LM_NEED_MAP();
  // max_iter = 0: IterationZero fills H, g, cost and the controller stops immediately
  int rc = lm_solve_problem(ctx, map_problem(ctx, 0), n_max_corner + n_max_surf, 0, 0);
  if (rc) return rc;
  k_lm_export_normal_eq<<<1, 32, 0, ctx->stream>>>(ctx->d_lm, ctx->d_partials + 32 * 1024);
  LM_LAUNCH_CHECK();
  return LMONO_OK;
}

// ---- cube-sharded map: one LM evaluation = eval (local partials) -> all-reduce on the host side -> control
int lm_shard_lm_begin(lmono_ctx* ctx, int solve_index) {
  k_lm_begin<<<1, 1024, 0, ctx->stream>>>(ctx->d_lm, map_problem(ctx, solve_index), 4, 1);
  LM_LAUNCH_CHECK();
  return LMONO_OK;
}
int lm_shard_lm_eval(lmono_ctx* ctx, int solve_index) {
  const int blocks = eval_blocks(ctx, ctx->shard_nc + ctx->shard_ns);
  k_lm_eval<<<blocks, EVAL_THREADS, 0, ctx->stream>>>(ctx->d_lm, map_problem(ctx, solve_index), ctx->d_partials, 1, ctx->d_shard_ws);
  LM_LAUNCH_CHECK();
  return LMONO_OK;
}
int lm_shard_lm_control(lmono_ctx* ctx, int solve_index) {
  k_lm_control<<<1, 32, 0, ctx->stream>>>(ctx->d_lm, map_problem(ctx, solve_index), ctx->d_shard_ws, 1);
  LM_LAUNCH_CHECK();
  return LMONO_OK;
}
