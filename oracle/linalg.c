/*
 * ORACLE (test infrastructure): small dense algebra the reference obtains from Eigen 3.3
 * (un-vendored).  Restated from the published algorithms:
 *   - SelfAdjointEigenSolver<Matrix3d>::compute (scaling, 3x3 tridiagonalisation special
 *     case, implicit symmetric QR steps with Wilkinson shift, ascending sort) -- used at
 *     Aloam/src/laserMapping.cpp:605-611
 *   - ColPivHouseholderQR<Matrix<double,5,3>>::compute + solve -- :663
 *   - HouseholderQR least squares -- what ceres::DenseQRSolver does with the stacked
 *     [J; D] matrix (Ceres 1.14 dense_qr_solver.cc), used by lm.c
 * All arithmetic is IEEE double without FMA contraction (compile with -ffp-contract=off).
 */
#include "lmono_oracle.h"
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>

/* JacobiRotation<double>::makeGivens(p, q) (real case): G^T [p;q] = [r;0] */
static void make_givens(double p, double q, double* c, double* s) {
  if (q == 0.0) { *c = p < 0.0 ? -1.0 : 1.0; *s = 0.0; }
  else if (p == 0.0) { *c = 0.0; *s = q < 0.0 ? 1.0 : -1.0; }
  else if (fabs(p) > fabs(q)) {
    double t = q / p; double u = sqrt(1.0 + t * t); if (p < 0.0) u = -u;
    *c = 1.0 / u; *s = -t * (*c);
  } else {
    double t = p / q; double u = sqrt(1.0 + t * t); if (q < 0.0) u = -u;
    *s = -1.0 / u; *c = -t * (*s);
  }
}

/* internal::tridiagonal_qr_step (column-major Q, n = 3) */
static void tridiagonal_qr_step(double* diag, double* subdiag, int start, int end, double* Q /*col-major 3x3*/) {
  double td = (diag[end - 1] - diag[end]) * 0.5;
  double e = subdiag[end - 1];
  double mu = diag[end];
  if (td == 0.0) {
    mu -= fabs(e);
  } else {
    double e2 = e * e;
    double h = hypot(td, e);
    if (e2 == 0.0) mu -= (e / (td + (td > 0.0 ? 1.0 : -1.0))) * (e / h);
    else mu -= e2 / (td + (td > 0.0 ? h : -h));
  }
  double x = diag[start] - mu;
  double z = subdiag[start];
  for (int k = start; k < end; ++k) {
    double c, s;
    make_givens(x, z, &c, &s);
    double sdk = s * diag[k] + c * subdiag[k];
    double dkp1 = s * subdiag[k] + c * diag[k + 1];
    diag[k] = c * (c * diag[k] - s * subdiag[k]) - s * (c * subdiag[k] - s * diag[k + 1]);
    diag[k + 1] = s * sdk + c * dkp1;
    subdiag[k] = c * sdk - s * dkp1;
    if (k > start) subdiag[k - 1] = c * subdiag[k - 1] - s * z;
    x = subdiag[k];
    if (k < end - 1) {
      z = -s * subdiag[k + 1];
      subdiag[k + 1] = c * subdiag[k + 1];
    }
    /* Q = Q * G : q.applyOnTheRight(k, k+1, rot) */
    for (int i = 0; i < 3; ++i) {
      double xi = Q[k * 3 + i], yi = Q[(k + 1) * 3 + i];
      Q[k * 3 + i] = c * xi - s * yi;
      Q[(k + 1) * 3 + i] = s * xi + c * yi;
    }
  }
}

int lmono_cpu_eigh3(const double A[9], double w[3], double V[9]) {
  /* only the lower triangle is referenced; map to [-1,1] */
  double m00 = A[0], m10 = A[3], m11 = A[4], m20 = A[6], m21 = A[7], m22 = A[8];
  double scale = fabs(m00);
  if (fabs(m10) > scale) scale = fabs(m10);
  if (fabs(m11) > scale) scale = fabs(m11);
  if (fabs(m20) > scale) scale = fabs(m20);
  if (fabs(m21) > scale) scale = fabs(m21);
  if (fabs(m22) > scale) scale = fabs(m22);
  if (scale == 0.0) scale = 1.0;
  m00 /= scale; m10 /= scale; m11 /= scale; m20 /= scale; m21 /= scale; m22 /= scale;

  double diag[3], subdiag[2];
  double Q[9]; /* column-major */
  /* tridiagonalization_inplace_selector<MatrixType,3,false> */
  const double tol = DBL_MIN;
  diag[0] = m00;
  double v1norm2 = m20 * m20;
  if (v1norm2 <= tol) {
    diag[1] = m11; diag[2] = m22;
    subdiag[0] = m10; subdiag[1] = m21;
    Q[0] = 1; Q[1] = 0; Q[2] = 0; Q[3] = 0; Q[4] = 1; Q[5] = 0; Q[6] = 0; Q[7] = 0; Q[8] = 1;
  } else {
    double beta = sqrt(m10 * m10 + v1norm2);
    double invBeta = 1.0 / beta;
    double m01 = m10 * invBeta;
    double m02 = m20 * invBeta;
    double q = 2.0 * m01 * m21 + m02 * (m22 - m11);
    diag[1] = m11 + m02 * q;
    diag[2] = m22 - m02 * q;
    subdiag[0] = beta;
    subdiag[1] = m21 - m01 * q;
    /* mat << 1,0,0, 0,m01,m02, 0,m02,-m01  (row listing) -> column-major storage */
    Q[0] = 1; Q[1] = 0;   Q[2] = 0;
    Q[3] = 0; Q[4] = m01; Q[5] = m02;
    Q[6] = 0; Q[7] = m02; Q[8] = -m01;
  }
  /* computeFromTridiagonal_impl */
  const int n = 3;
  int end = n - 1, start = 0, iter = 0;
  const int maxIterations = 30;
  const double considerAsZero = DBL_MIN;
  const double precision = 2.0 * DBL_EPSILON;
  while (end > 0) {
    for (int i = start; i < end; ++i) {
      /* isMuchSmallerThan(|sub|, |d_i|+|d_i+1|, precision): |sub| <= (|d_i|+|d_i+1|)*precision */
      if (fabs(subdiag[i]) <= (fabs(diag[i]) + fabs(diag[i + 1])) * precision || fabs(subdiag[i]) <= considerAsZero)
        subdiag[i] = 0.0;
    }
    while (end > 0 && subdiag[end - 1] == 0.0) end--;
    if (end <= 0) break;
    iter++;
    if (iter > maxIterations * n) break;
    start = end - 1;
    while (start > 0 && subdiag[start - 1] != 0.0) start--;
    tridiagonal_qr_step(diag, subdiag, start, end, Q);
  }
  int ok = iter <= maxIterations * n;
  /* ascending selection sort with column swaps */
  for (int i = 0; i < n - 1; ++i) {
    int k = 0; double mn = diag[i];
    for (int j = 1; j < n - i; ++j) if (diag[i + j] < mn) { mn = diag[i + j]; k = j; }
    if (k > 0) {
      double tmp = diag[i]; diag[i] = diag[k + i]; diag[k + i] = tmp;
      for (int r = 0; r < 3; ++r) { double t2 = Q[i * 3 + r]; Q[i * 3 + r] = Q[(k + i) * 3 + r]; Q[(k + i) * 3 + r] = t2; }
    }
  }
  for (int i = 0; i < 3; ++i) w[i] = diag[i] * scale;
  /* row-major V[r*3+c] = Q(r,c) */
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) V[r * 3 + c] = Q[c * 3 + r];
  return ok ? 0 : 1;
}

/* makeHouseholderInPlace on v[0..len): returns tau, beta; essential part overwrites v[1..) */
static void make_householder(double* v, int len, int stride, double* tau, double* beta) {
  double tailSqNorm = 0.0;
  for (int i = 1; i < len; ++i) tailSqNorm += v[i * stride] * v[i * stride];
  double c0 = v[0];
  const double tol = DBL_MIN;
  if (tailSqNorm <= tol) {
    *tau = 0.0; *beta = c0;
    for (int i = 1; i < len; ++i) v[i * stride] = 0.0;
  } else {
    double b = sqrt(c0 * c0 + tailSqNorm);
    if (c0 >= 0.0) b = -b;
    for (int i = 1; i < len; ++i) v[i * stride] = v[i * stride] / (c0 - b);
    *tau = (b - c0) / b;
    *beta = b;
  }
}

int lmono_cpu_colpiv_qr_solve_5x3(const double Ain[15], const double bin[5], double x[3]) {
  enum { R = 5, C = 3 };
  double qr[R][C];
  for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) qr[r][c] = Ain[r * C + c];
  double hCoeffs[C]; int transp[C];
  double normsUpdated[C], normsDirect[C];
  for (int k = 0; k < C; ++k) {
    double s = 0.0; for (int r = 0; r < R; ++r) s += qr[r][k] * qr[r][k];
    normsDirect[k] = sqrt(s); normsUpdated[k] = normsDirect[k];
  }
  double maxn = normsUpdated[0]; for (int k = 1; k < C; ++k) if (normsUpdated[k] > maxn) maxn = normsUpdated[k];
  double th = maxn * DBL_EPSILON; double threshold_helper = (th * th) / (double)R;
  double norm_downdate_threshold = sqrt(DBL_EPSILON);
  int nonzero_pivots = C;
  for (int k = 0; k < C; ++k) {
    int big = k; double bn = normsUpdated[k];
    for (int j = k + 1; j < C; ++j) if (normsUpdated[j] > bn) { bn = normsUpdated[j]; big = j; }
    double biggest_sq = bn * bn;
    if (nonzero_pivots == C && biggest_sq < threshold_helper * (double)(R - k)) nonzero_pivots = k;
    transp[k] = big;
    if (k != big) {
      for (int r = 0; r < R; ++r) { double t = qr[r][k]; qr[r][k] = qr[r][big]; qr[r][big] = t; }
      double t = normsUpdated[k]; normsUpdated[k] = normsUpdated[big]; normsUpdated[big] = t;
      t = normsDirect[k]; normsDirect[k] = normsDirect[big]; normsDirect[big] = t;
    }
    double beta;
    make_householder(&qr[k][k], R - k, C, &hCoeffs[k], &beta);
    qr[k][k] = beta;
    /* applyHouseholderOnTheLeft to bottomRightCorner(R-k, C-k-1) */
    if (R - k == 1) {
      for (int j = k + 1; j < C; ++j) qr[k][j] *= (1.0 - hCoeffs[k]);
    } else if (hCoeffs[k] != 0.0) {
      for (int j = k + 1; j < C; ++j) {
        double tmp = 0.0;
        for (int r = k + 1; r < R; ++r) tmp += qr[r][k] * qr[r][j];
        tmp += qr[k][j];
        qr[k][j] -= hCoeffs[k] * tmp;
        for (int r = k + 1; r < R; ++r) qr[r][j] -= hCoeffs[k] * qr[r][k] * tmp;
      }
    }
    for (int j = k + 1; j < C; ++j) {
      if (normsUpdated[j] != 0.0) {
        double temp = fabs(qr[k][j]) / normsUpdated[j];
        temp = (1.0 + temp) * (1.0 - temp);
        temp = temp < 0.0 ? 0.0 : temp;
        double ratio = normsUpdated[j] / normsDirect[j];
        double temp2 = temp * (ratio * ratio);
        if (temp2 <= norm_downdate_threshold) {
          double s = 0.0; for (int r = k + 1; r < R; ++r) s += qr[r][j] * qr[r][j];
          normsDirect[j] = sqrt(s); normsUpdated[j] = normsDirect[j];
        } else {
          normsUpdated[j] *= sqrt(temp);
        }
      }
    }
  }
  /* permutation indices from transpositions */
  int perm[C]; for (int i = 0; i < C; ++i) perm[i] = i;
  for (int k = 0; k < C; ++k) { int t = perm[k]; perm[k] = perm[transp[k]]; perm[transp[k]] = t; }
  /* solve */
  if (nonzero_pivots == 0) { x[0] = x[1] = x[2] = 0.0; return 0; }
  double c[R]; for (int r = 0; r < R; ++r) c[r] = bin[r];
  for (int k = 0; k < nonzero_pivots; ++k) {
    /* c = H_k c, H_k = I - tau v v^T, v = [1; essential] */
    if (R - k == 1) { c[k] *= (1.0 - hCoeffs[k]); }
    else if (hCoeffs[k] != 0.0) {
      double tmp = 0.0;
      for (int r = k + 1; r < R; ++r) tmp += qr[r][k] * c[r];
      tmp += c[k];
      c[k] -= hCoeffs[k] * tmp;
      for (int r = k + 1; r < R; ++r) c[r] -= hCoeffs[k] * qr[r][k] * tmp;
    }
  }
  /* upper-triangular back substitution on the leading nonzero_pivots block */
  for (int i = nonzero_pivots - 1; i >= 0; --i) {
    double s = c[i];
    for (int j = i + 1; j < nonzero_pivots; ++j) s -= qr[i][j] * c[j];
    c[i] = s / qr[i][i];
  }
  for (int i = 0; i < nonzero_pivots; ++i) x[perm[i]] = c[i];
  for (int i = nonzero_pivots; i < C; ++i) x[perm[i]] = 0.0;
  return 0;
}

/* min || A y - b ||, A is m x n row-major (overwritten), b length m (overwritten), n <= 8.
 * Unpivoted Householder QR as Eigen::HouseholderQR::solve. */
int lmono_cpu_householder_ls(double* A, double* b, int m, int n, double* y) {
  double tau[8];
  for (int k = 0; k < n; ++k) {
    double beta;
    make_householder(&A[(size_t)k * n + k], m - k, n, &tau[k], &beta);
    A[(size_t)k * n + k] = beta;
    if (tau[k] != 0.0) {
      for (int j = k + 1; j < n; ++j) {
        double tmp = 0.0;
        for (int r = k + 1; r < m; ++r) tmp += A[(size_t)r * n + k] * A[(size_t)r * n + j];
        tmp += A[(size_t)k * n + j];
        A[(size_t)k * n + j] -= tau[k] * tmp;
        for (int r = k + 1; r < m; ++r) A[(size_t)r * n + j] -= tau[k] * A[(size_t)r * n + k] * tmp;
      }
      double tmp = 0.0;
      for (int r = k + 1; r < m; ++r) tmp += A[(size_t)r * n + k] * b[r];
      tmp += b[k];
      b[k] -= tau[k] * tmp;
      for (int r = k + 1; r < m; ++r) b[r] -= tau[k] * A[(size_t)r * n + k] * tmp;
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < n; ++j) s -= A[(size_t)i * n + j] * y[j];
    if (A[(size_t)i * n + i] == 0.0) return 1;
    y[i] = s / A[(size_t)i * n + i];
  }
  return 0;
}
