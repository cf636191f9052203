/*
 * ORACLE (test infrastructure): the residuals of Aloam/src/lidarFactor.hpp and the
 * ceres::Solve the reference runs on them (Aloam/src/laserOdometry.cpp:284-291,494-499,
 * Aloam/src/laserMapping.cpp:565-572,713-720): HuberLoss(0.1),
 * EigenQuaternionParameterization on q(x,y,z,w), DENSE_QR, max_num_iterations=4, all other
 * options default.  Ceres is un-vendored (Dockerfile pins 1.12.0, the author's machine had
 * 1.14.x); this restates the published 1.14 algorithm:
 * residual_block.cc (loss "Corrector"), trust_region_minimizer.cc,
 * levenberg_marquardt_strategy.cc, dense_qr_solver.cc, local_parameterization.cc.
 * Jacobians are the exact derivatives of the expressions Ceres' autodiff differentiates
 * (Eigen's quaternion-vector product formula), then multiplied by the 4x3 local
 * parameterisation Jacobian, as Ceres does.
 */
#include "lmono_oracle.h"
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>

int lmono_cpu_householder_ls(double* A, double* b, int m, int n, double* y);

static inline void cross3(const double a[3], const double b[3], double o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

/* lp = q*p + t with Eigen's QuaternionBase::_transformVector:
 *   uv = u.cross(v); uv += uv; return v + w*uv + u.cross(uv)
 * dlp[3][4]: derivative w.r.t. (qx,qy,qz,qw). */
static void transform_point(const double q[4], const double t[3], const double p[3], double lp[3], double dlp[3][4]) {
  const double u[3] = { q[0], q[1], q[2] }; const double w = q[3];
  double uv[3]; cross3(u, p, uv);
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  double uuv[3]; cross3(u, uv, uuv);
  for (int k = 0; k < 3; ++k) lp[k] = (p[k] + w * uv[k]) + uuv[k];
  for (int k = 0; k < 3; ++k) lp[k] += t[k];
  if (dlp) {
    for (int j = 0; j < 3; ++j) {
      double e[3] = { 0, 0, 0 }; e[j] = 1.0;
      double duv[3]; cross3(e, p, duv); duv[0] *= 2.0; duv[1] *= 2.0; duv[2] *= 2.0;
      double c1[3], c2[3]; cross3(e, uv, c1); cross3(u, duv, c2);
      for (int k = 0; k < 3; ++k) dlp[k][j] = w * duv[k] + c1[k] + c2[k];
    }
    for (int k = 0; k < 3; ++k) dlp[k][3] = uv[k];
  }
}

/* Eigen 3.3 QuaternionBase::slerp(t, other) with *this = Identity (lidarFactor.hpp:29-31, laserOdometry.cpp:120):
 *   d = dot = other.w; absD = |d|; if (absD >= 1 - eps) { scale0 = 1 - t; scale1 = t; }
 *   else { theta = acos(absD); scale0 = sin((1 - t) theta) / sin(theta); scale1 = sin(t theta) / sin(theta); }
 *   if (d < 0) scale1 = -scale1;  result = scale0 * Identity + scale1 * other        (not renormalised)
 * dqs[4][4]: derivative of the result (x,y,z,w) w.r.t. other (x,y,z,w), as Ceres' Jets propagate it (the scales depend
 * on other.w only; constant in the linear branch). */
void lmono_cpu_slerp_identity(double t, const double q[4], double qs[4], double dqs[4][4]) {
  const double d = q[3], absD = fabs(d);
  double scale0, scale1, ds0 = 0.0, ds1 = 0.0;      /* d scale / d other.w */
  if (absD >= 1.0 - DBL_EPSILON) { scale0 = 1.0 - t; scale1 = t; }
  else {
    const double theta = acos(absD), sinTheta = sin(theta), cosTheta = cos(theta);
    const double a0 = (1.0 - t) * theta, a1 = t * theta;
    scale0 = sin(a0) / sinTheta; scale1 = sin(a1) / sinTheta;
    const double dtheta = -(d < 0.0 ? -1.0 : 1.0) / sqrt(1.0 - absD * absD);      /* d acos(|w|) / dw */
    ds0 = ((1.0 - t) * cos(a0) * sinTheta - sin(a0) * cosTheta) / (sinTheta * sinTheta) * dtheta;
    ds1 = (t * cos(a1) * sinTheta - sin(a1) * cosTheta) / (sinTheta * sinTheta) * dtheta;
  }
  if (d < 0.0) { scale1 = -scale1; ds1 = -ds1; }
  qs[0] = scale1 * q[0]; qs[1] = scale1 * q[1]; qs[2] = scale1 * q[2]; qs[3] = scale0 + scale1 * q[3];
  if (dqs) {
    memset(dqs, 0, 16 * sizeof(double));
    for (int k = 0; k < 3; ++k) { dqs[k][k] = scale1; dqs[k][3] = ds1 * q[k]; }
    dqs[3][3] = ds0 + ds1 * q[3] + scale1;
  }
}

/* EigenQuaternionParameterization::ComputeJacobian, x = (x,y,z,w): 4x3 */
static void local_jacobian(const double q[4], double P[4][3]) {
  P[0][0] =  q[3]; P[0][1] =  q[2]; P[0][2] = -q[1];
  P[1][0] = -q[2]; P[1][1] =  q[3]; P[1][2] =  q[0];
  P[2][0] =  q[1]; P[2][1] = -q[0]; P[2][2] =  q[3];
  P[3][0] = -q[0]; P[3][1] = -q[1]; P[3][2] = -q[2];
}

/* raw (uncorrected) residual r[nr] and Jacobian Jq[nr][4], Jt[nr][3] of one block */
static int eval_block(const o_factor* f, const double q[4], const double t[3], double r[3], double Jq[3][4], double Jt[3][3], int want_jac) {
  double lp[3], dlp[3][4];
  const double s = (f->type != O_FACTOR_PLANE_NORM && f->s != 0.0) ? f->s : 1.0;
  if (s == 1.0) transform_point(q, t, f->p, lp, want_jac ? dlp : NULL);
  else {
    /* lidarFactor.hpp:27-34,73-79: q_last_curr = Identity.slerp(s, q); t_last_curr = s * t; lp = q_last_curr * cp + t_last_curr */
    double qs[4], dqs[4][4], ts[3] = { s * t[0], s * t[1], s * t[2] }, dl[3][4];
    lmono_cpu_slerp_identity(s, q, qs, want_jac ? dqs : NULL);
    transform_point(qs, ts, f->p, lp, want_jac ? dl : NULL);
    if (want_jac)
      for (int k = 0; k < 3; ++k)
        for (int j = 0; j < 4; ++j) { double a = 0.0; for (int m = 0; m < 4; ++m) a += dl[k][m] * dqs[m][j]; dlp[k][j] = a; }
  }
  if (f->type == O_FACTOR_EDGE) {
    /* lidarFactor.hpp:35-40 */
    double da[3] = { lp[0] - f->a[0], lp[1] - f->a[1], lp[2] - f->a[2] };
    double db[3] = { lp[0] - f->b[0], lp[1] - f->b[1], lp[2] - f->b[2] };
    double nu[3]; cross3(da, db, nu);
    double de[3] = { f->a[0] - f->b[0], f->a[1] - f->b[1], f->a[2] - f->b[2] };
    double den = sqrt(de[0] * de[0] + de[1] * de[1] + de[2] * de[2]);
    for (int k = 0; k < 3; ++k) r[k] = nu[k] / den;
    if (want_jac) {
      for (int j = 0; j < 4; ++j) {
        double d[3] = { dlp[0][j], dlp[1][j], dlp[2][j] };
        double c1[3], c2[3]; cross3(d, db, c1); cross3(da, d, c2);
        for (int k = 0; k < 3; ++k) Jq[k][j] = (c1[k] + c2[k]) / den;
      }
      for (int j = 0; j < 3; ++j) {
        double d[3] = { 0, 0, 0 }; d[j] = s;            /* d lp / d t = s I */
        double c1[3], c2[3]; cross3(d, db, c1); cross3(da, d, c2);
        for (int k = 0; k < 3; ++k) Jt[k][j] = (c1[k] + c2[k]) / den;
      }
    }
    return 3;
  } else if (f->type == O_FACTOR_PLANE) {
    /* lidarFactor.hpp:87  residual = (lp - lpj).dot(ljm) */
    const double* n = f->b;
    r[0] = (lp[0] - f->a[0]) * n[0] + (lp[1] - f->a[1]) * n[1] + (lp[2] - f->a[2]) * n[2];
    if (want_jac) {
      for (int j = 0; j < 4; ++j) Jq[0][j] = dlp[0][j] * n[0] + dlp[1][j] * n[1] + dlp[2][j] * n[2];
      for (int j = 0; j < 3; ++j) Jt[0][j] = s * n[j];
    }
    return 1;
  } else {
    /* lidarFactor.hpp:123  residual = norm.dot(point_w) + negative_OA_dot_norm */
    const double* n = f->a;
    r[0] = (n[0] * lp[0] + n[1] * lp[1] + n[2] * lp[2]) + f->b[0];
    if (want_jac) {
      for (int j = 0; j < 4; ++j) Jq[0][j] = n[0] * dlp[0][j] + n[1] * dlp[1][j] + n[2] * dlp[2][j];
      for (int j = 0; j < 3; ++j) Jt[0][j] = n[j];
    }
    return 1;
  }
}

/* ceres::HuberLoss(a=0.1)::Evaluate */
static void huber(double s, double rho[3]) {
  const double a = 0.1, b = a * a;
  if (s > b) {
    const double r = sqrt(s);
    rho[0] = 2.0 * a * r - b;
    rho[1] = a / r; if (rho[1] < DBL_MIN) rho[1] = DBL_MIN;
    rho[2] = -rho[1] / (2.0 * s);
  } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
}

static int count_residuals(const o_factor* f, int nf) {
  int n = 0;
  for (int i = 0; i < nf; ++i) n += (f[i].type == O_FACTOR_EDGE) ? 3 : 1;
  return n;
}

/* Test hook (oracle/_ref builds): when set, block i of a problem is evaluated by the callback instead of eval_block --
 * the reference's own cost functors on dual numbers (refstubs/ceres/ceres.h) -- and only f[i].type (the residual
 * count) is read from the factor records.  Process-wide, not thread-safe. */
static o_block_hook g_block_hook = NULL;
static void* g_block_hook_user = NULL;
void lmono_cpu_lm_set_block_hook(o_block_hook fn, void* user) { g_block_hook = fn; g_block_hook_user = user; }

/* ProgramEvaluator::Evaluate: cost, corrected residuals r[nres], corrected local Jacobian
 * J[nres][6] (row-major), gradient g[6] = J^T r. r, J, g may be NULL. */
static void evaluate(const o_factor* f, int nf, const double x[7], double* cost, double* r, double* J, double* g) {
  const double* q = x; const double* t = x + 4;
  double P[4][3]; local_jacobian(q, P);
  double c = 0.0;
  if (g) for (int j = 0; j < 6; ++j) g[j] = 0.0;
  int row = 0;
  const int want_jac = (J != NULL) || (g != NULL);
  for (int i = 0; i < nf; ++i) {
    double rb[3], Jq[3][4], Jt[3][3];
    int nr = g_block_hook ? g_block_hook(g_block_hook_user, i, q, t, rb, Jq, Jt, want_jac) : eval_block(&f[i], q, t, rb, Jq, Jt, want_jac);
    double s = 0.0; for (int k = 0; k < nr; ++k) s += rb[k] * rb[k];
    double rho[3]; huber(s, rho);
    c += 0.5 * rho[0];
    if (r || want_jac) {
      /* Corrector: rho'' <= 0 for Huber => residual_scaling = sqrt(rho'), alpha = 0 */
      const double sr = sqrt(rho[1]);
      for (int k = 0; k < nr; ++k) {
        double Jl[6];
        if (want_jac) {
          for (int cidx = 0; cidx < 3; ++cidx) {
            double a = 0.0;
            for (int m = 0; m < 4; ++m) a += Jq[k][m] * P[m][cidx];
            Jl[cidx] = a * sr;
          }
          for (int cidx = 0; cidx < 3; ++cidx) Jl[3 + cidx] = Jt[k][cidx] * sr;
        }
        const double rc = rb[k] * sr;
        if (r) r[row + k] = rc;
        if (J) for (int m = 0; m < 6; ++m) J[(size_t)(row + k) * 6 + m] = Jl[m];
        if (g) for (int m = 0; m < 6; ++m) g[m] += Jl[m] * rc;
      }
    }
    row += nr;
  }
  *cost = c;
}

int lmono_cpu_normal_eq(const o_factor* f, int nf, const o_pose* xp, double H[36], double g[6], double* cost) {
  double x[7]; memcpy(x, xp->q, 4 * sizeof(double)); memcpy(x + 4, xp->t, 3 * sizeof(double));
  int nres = count_residuals(f, nf);
  double* J = (double*)malloc((size_t)(nres > 0 ? nres : 1) * 6 * sizeof(double));
  double* r = (double*)malloc((size_t)(nres > 0 ? nres : 1) * sizeof(double));
  evaluate(f, nf, x, cost, r, J, g);
  for (int a = 0; a < 36; ++a) H[a] = 0.0;
  for (int i = 0; i < nres; ++i)
    for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) H[a * 6 + b] += J[(size_t)i * 6 + a] * J[(size_t)i * 6 + b];
  free(J); free(r);
  return 0;
}

/* ProgramEvaluator::Plus: EigenQuaternionParameterization::Plus on q, identity on t */
static void plus(const double x[7], const double d[6], double out[7]) {
  const double nd = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (nd > 0.0) {
    const double sbd = sin(nd) / nd;
    const double aw = cos(nd), ax = sbd * d[0], ay = sbd * d[1], az = sbd * d[2];
    const double bx = x[0], by = x[1], bz = x[2], bw = x[3];
    /* Eigen quaternion product a*b */
    out[3] = aw * bw - ax * bx - ay * by - az * bz;
    out[0] = aw * bx + ax * bw + ay * bz - az * by;
    out[1] = aw * by + ay * bw + az * bx - ax * bz;
    out[2] = aw * bz + az * bw + ax * by - ay * bx;
  } else { out[0] = x[0]; out[1] = x[1]; out[2] = x[2]; out[3] = x[3]; }
  out[4] = x[4] + d[3]; out[5] = x[5] + d[4]; out[6] = x[6] + d[5];
}

static double vec_norm(const double* v, int n) { double s = 0; for (int i = 0; i < n; ++i) s += v[i] * v[i]; return sqrt(s); }

static double gradient_max_norm(const double x[7], const double g[6]) {
  double ng[6], xp[7];
  for (int i = 0; i < 6; ++i) ng[i] = -g[i];
  plus(x, ng, xp);
  double m = 0.0;
  for (int i = 0; i < 7; ++i) { double a = fabs(x[i] - xp[i]); if (a > m) m = a; }
  return m;
}

int lmono_cpu_lm_solve(const o_factor* f, int nf, o_pose* xp, int max_iter, o_solve_summary* sum) {
  o_solve_summary S; memset(&S, 0, sizeof(S));
  S.num_factors = nf;
  if (nf <= 0) { S.termination = 6; if (sum) *sum = S; return 0; }
  const int nres = count_residuals(f, nf);
  const double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
  const double min_relative_decrease = 1e-3, min_radius = 1e-32, max_radius = 1e16;
  const double min_diagonal = 1e-6, max_diagonal = 1e32;

  double x[7]; memcpy(x, xp->q, 4 * sizeof(double)); memcpy(x + 4, xp->t, 3 * sizeof(double));
  double* J = (double*)malloc((size_t)nres * 6 * sizeof(double));
  double* r = (double*)malloc((size_t)nres * sizeof(double));
  double* A = (double*)malloc((size_t)(nres + 6) * 6 * sizeof(double));
  double* rhs = (double*)malloc((size_t)(nres + 6) * sizeof(double));
  double g[6], scaling[6], diagonal[6], cost;

  /* IterationZero */
  double x_norm = vec_norm(x, 7);
  evaluate(f, nf, x, &cost, r, J, g);
  for (int j = 0; j < 6; ++j) {
    double s = 0.0; for (int i = 0; i < nres; ++i) s += J[(size_t)i * 6 + j] * J[(size_t)i * 6 + j];
    scaling[j] = 1.0 / (1.0 + sqrt(s));
  }
  for (int i = 0; i < nres; ++i) for (int j = 0; j < 6; ++j) J[(size_t)i * 6 + j] *= scaling[j];
  S.initial_cost = cost;
  double gmax = gradient_max_norm(x, g);
  double radius = 1e4, decrease_factor = 2.0;
  int reuse_diagonal = 0, num_invalid = 0;
  int iteration = 0;
  S.termination = 0;
  if (gmax <= gradient_tolerance) { S.termination = 1; goto done; }

  for (;;) {
    if (iteration >= max_iter) { S.termination = 0; break; }
    if (!(radius > min_radius)) { S.termination = 4; break; }
    ++iteration;
    /* LevenbergMarquardtStrategy::ComputeStep */
    if (!reuse_diagonal) {
      for (int j = 0; j < 6; ++j) {
        double s = 0.0; for (int i = 0; i < nres; ++i) s += J[(size_t)i * 6 + j] * J[(size_t)i * 6 + j];
        diagonal[j] = s < min_diagonal ? min_diagonal : (s > max_diagonal ? max_diagonal : s);
      }
    }
    double lm_diag[6]; for (int j = 0; j < 6; ++j) lm_diag[j] = sqrt(diagonal[j] / radius);
    /* DenseQRSolver: [J; D] y = [r; 0], step = -y */
    memcpy(A, J, (size_t)nres * 6 * sizeof(double));
    memset(A + (size_t)nres * 6, 0, 36 * sizeof(double));
    for (int j = 0; j < 6; ++j) A[(size_t)(nres + j) * 6 + j] = lm_diag[j];
    memcpy(rhs, r, (size_t)nres * sizeof(double));
    for (int j = 0; j < 6; ++j) rhs[nres + j] = 0.0;
    double step[6];
    int fail = lmono_cpu_householder_ls(A, rhs, nres + 6, 6, step);
    for (int j = 0; j < 6; ++j) if (!isfinite(step[j])) fail = 1;
    reuse_diagonal = 1;
    int step_valid = 0; double model_cost_change = 0.0;
    if (!fail) {
      for (int j = 0; j < 6; ++j) step[j] = -step[j];
      /* model_cost_change = -(J step)^T (r + J step / 2) */
      double acc = 0.0;
      for (int i = 0; i < nres; ++i) {
        double mr = 0.0; for (int j = 0; j < 6; ++j) mr += J[(size_t)i * 6 + j] * step[j];
        acc += mr * (r[i] + mr / 2.0);
      }
      model_cost_change = -acc;
      step_valid = model_cost_change > 0.0;
    }
    if (!step_valid) {
      /* HandleInvalidStep */
      if (++num_invalid >= 5) { S.termination = 5; break; }
      radius *= 0.5; reuse_diagonal = 1;
      continue;
    }
    num_invalid = 0;
    double delta[6]; for (int j = 0; j < 6; ++j) delta[j] = step[j] * scaling[j];
    double cand[7]; plus(x, delta, cand);
    double cand_cost; evaluate(f, nf, cand, &cand_cost, NULL, NULL, NULL);
    /* ParameterToleranceReached */
    double dn = 0.0; for (int i = 0; i < 7; ++i) dn += (x[i] - cand[i]) * (x[i] - cand[i]);
    double step_norm = sqrt(dn);
    if (step_norm <= parameter_tolerance * (x_norm + parameter_tolerance)) { S.termination = 2; break; }
    /* FunctionToleranceReached */
    double cost_change = cost - cand_cost;
    if (fabs(cost_change) <= function_tolerance * cost) { S.termination = 3; break; }
    double relative_decrease = cost_change / model_cost_change;
    if (relative_decrease > min_relative_decrease) {
      /* HandleSuccessfulStep */
      memcpy(x, cand, sizeof(x)); x_norm = vec_norm(x, 7);
      evaluate(f, nf, x, &cost, r, J, g);
      for (int i = 0; i < nres; ++i) for (int j = 0; j < 6; ++j) J[(size_t)i * 6 + j] *= scaling[j];
      double tq = 2.0 * relative_decrease - 1.0;
      double den = 1.0 - tq * tq * tq; if (den < 1.0 / 3.0) den = 1.0 / 3.0;
      radius = radius / den; if (radius > max_radius) radius = max_radius;
      decrease_factor = 2.0; reuse_diagonal = 0;
      ++S.num_successful;
      gmax = gradient_max_norm(x, g);
      if (iteration < max_iter && gmax <= gradient_tolerance) { S.termination = 1; break; }
    } else {
      /* HandleUnsuccessfulStep */
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = 1;
    }
  }
done:
  S.iterations = iteration;
  S.final_cost = cost;
  memcpy(xp->q, x, 4 * sizeof(double)); memcpy(xp->t, x + 4, 3 * sizeof(double));
  free(J); free(r); free(A); free(rhs);
  if (sum) *sum = S;
  return 0;
}
