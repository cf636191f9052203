// Scan-to-map data association (Aloam/src/laserMapping.cpp:577-687) on device:
//   pointAssociateToMap -> exact 5-NN (replaces kdtree*FromMap->nearestKSearch, :582,648)
//   -> d2[4] < 1.0 gate -> 3x3 covariance + symmetric eigen-decomposition, lambda2 > 3 lambda1
//   (corner, :586-616) or 5x3 least-squares plane + 0.2 m validity (surf, :650-679)
//   -> one LmFactor per query.
//
// Exactness of the search: a correspondence is only used if its 5th neighbour has fp32
// d2 < 1.0, hence every neighbour that matters satisfies |dx|,|dy|,|dz| < 1 and its integer
// floor differs from the query's by at most 1 per axis.  With 2 m cells keyed on
// floor(x) >> 1 those floors fall in exactly 2 cells per axis: 8 cells per query, spread over
// the GROUP lanes that share a query (cells that straddle a 50 m cube border are looked up in both
// cubes, <= 27 (cube, cell) pairs).  d2 is accumulated like FLANN's L2_Simple:
// ((dx*dx) + dy*dy) + dz*dz in fp32 without FMA.  Ties are broken on the index in the
// reference's concatenation order (:533-537), so results are order-independent.
#include "common.cuh"
#include <float.h>
#include <stdlib.h>

constexpr int KNN_K = 5;
constexpr int GROUP_DEFAULT = 8;   // lanes per query (template parameter of the search kernels)


// ---- Eigen 3.3 SelfAdjointEigenSolver<Matrix3d> (same operation order as the test oracle, so the fit gates are reproducible)
__device__ __forceinline__ void d_make_givens(double p, double q, double* c, double* s) {
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

// The value of hypot() only steers the Wilkinson shift (any shift converges to the same eigen-pairs up to
// rounding); CUDA's hypot() is within 1 ulp of libm's.  Gate margins are reported by the tests.
//
// Register-resident 3x3 specialisation: Eigen's loops run over k in [start, end) with start in {0,1}, end in
// {1,2}; here they are unrolled over k = 0, 1 with guards, so that diag / subdiag / Q are only ever indexed
// with compile-time constants and live in registers (dynamic indexing put them in local memory, and with
// ~3 warps per SM nothing hid that latency).  Operation order is unchanged.
#define D3(arr, i) ((i) == 0 ? arr##0 : ((i) == 1 ? arr##1 : arr##2))

struct Eig3 { double d0, d1, d2, e0, e1; double q[9]; };    // diag, subdiag, Q (row k = eigenvector k while iterating)

__device__ __forceinline__ void d_tridiagonal_qr_step(Eig3& E, int start, int end) {
  const double dEnd = end == 2 ? E.d2 : E.d1, dEndm1 = end == 2 ? E.d1 : E.d0, eEndm1 = end == 2 ? E.e1 : E.e0;
  double td = (dEndm1 - dEnd) * 0.5;
  double e = eEndm1;
  double mu = dEnd;
  if (td == 0.0) {
    mu -= fabs(e);
  } else {
    double e2 = e * e;
    double h = hypot(td, e);
    if (e2 == 0.0) mu -= (e / (td + (td > 0.0 ? 1.0 : -1.0))) * (e / h);
    else mu -= e2 / (td + (td > 0.0 ? h : -h));
  }
  double x = (start == 0 ? E.d0 : E.d1) - mu;
  double z = start == 0 ? E.e0 : E.e1;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (k < start || k >= end) continue;
    double c, s;
    d_make_givens(x, z, &c, &s);
    double& dk = k == 0 ? E.d0 : E.d1;
    double& dk1 = k == 0 ? E.d1 : E.d2;
    double& ek = k == 0 ? E.e0 : E.e1;
    double sdk = s * dk + c * ek;
    double dkp1 = s * ek + c * dk1;
    dk = c * (c * dk - s * ek) - s * (c * ek - s * dk1);
    dk1 = s * sdk + c * dkp1;
    ek = c * sdk - s * dkp1;
    if (k > start) E.e0 = c * E.e0 - s * z;          // only k = 1 > start = 0: subdiag[k - 1] = subdiag[0]
    x = ek;
    if (k < end - 1) {                                // only k = 0, end = 2: subdiag[k + 1] = subdiag[1]
      z = -s * E.e1;
      E.e1 = c * E.e1;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double xi = E.q[k * 3 + i], yi = E.q[(k + 1) * 3 + i];
      E.q[k * 3 + i] = c * xi - s * yi;
      E.q[(k + 1) * 3 + i] = s * xi + c * yi;
    }
  }
}

// A: row-major 3x3 (lower triangle used). w ascending; Q column-major (column c = eigenvector c)
__device__ __forceinline__ void d_eigh3(const double* A, double* w, double* Q) {
  double m00 = A[0], m10 = A[3], m11 = A[4], m20 = A[6], m21 = A[7], m22 = A[8];
  double scale = fabs(m00);
  if (fabs(m10) > scale) scale = fabs(m10);
  if (fabs(m11) > scale) scale = fabs(m11);
  if (fabs(m20) > scale) scale = fabs(m20);
  if (fabs(m21) > scale) scale = fabs(m21);
  if (fabs(m22) > scale) scale = fabs(m22);
  if (scale == 0.0) scale = 1.0;
  m00 /= scale; m10 /= scale; m11 /= scale; m20 /= scale; m21 /= scale; m22 /= scale;
  Eig3 E;
  const double tol = DBL_MIN;
  E.d0 = m00;
  double v1norm2 = m20 * m20;
  if (v1norm2 <= tol) {
    E.d1 = m11; E.d2 = m22; E.e0 = m10; E.e1 = m21;
    E.q[0] = 1; E.q[1] = 0; E.q[2] = 0; E.q[3] = 0; E.q[4] = 1; E.q[5] = 0; E.q[6] = 0; E.q[7] = 0; E.q[8] = 1;
  } else {
    double beta = sqrt(m10 * m10 + v1norm2);
    double invBeta = 1.0 / beta;
    double m01 = m10 * invBeta;
    double m02 = m20 * invBeta;
    double q = 2.0 * m01 * m21 + m02 * (m22 - m11);
    E.d1 = m11 + m02 * q;
    E.d2 = m22 - m02 * q;
    E.e0 = beta;
    E.e1 = m21 - m01 * q;
    E.q[0] = 1; E.q[1] = 0;   E.q[2] = 0;
    E.q[3] = 0; E.q[4] = m01; E.q[5] = m02;
    E.q[6] = 0; E.q[7] = m02; E.q[8] = -m01;
  }
  int end = 2, start = 0, iter = 0;
  const int maxIterations = 30;
  const double precision = 2.0 * DBL_EPSILON;
  while (end > 0) {
    // for (i = start; i < end; ++i) deflation test on subdiag[i]
    if (0 >= start && 0 < end) if (fabs(E.e0) <= (fabs(E.d0) + fabs(E.d1)) * precision || fabs(E.e0) <= DBL_MIN) E.e0 = 0.0;
    if (1 >= start && 1 < end) if (fabs(E.e1) <= (fabs(E.d1) + fabs(E.d2)) * precision || fabs(E.e1) <= DBL_MIN) E.e1 = 0.0;
    while (end > 0 && (end == 2 ? E.e1 : E.e0) == 0.0) end--;
    if (end <= 0) break;
    iter++;
    if (iter > maxIterations * 3) break;
    start = end - 1;
    while (start > 0 && (start == 2 ? E.e1 : E.e0) != 0.0) start--;     // subdiag[start - 1], start in {1}: e0
    d_tridiagonal_qr_step(E, start, end);
  }
  // eigenvalues ascending (selection sort of Eigen: first minimal index, strict <), eigenvectors follow
  {
    int k = 0; double mn = E.d0;
    if (E.d1 < mn) { mn = E.d1; k = 1; }
    if (E.d2 < mn) { k = 2; }
    if (k == 1) { double t = E.d0; E.d0 = E.d1; E.d1 = t;
#pragma unroll
      for (int r = 0; r < 3; ++r) { double t2 = E.q[r]; E.q[r] = E.q[3 + r]; E.q[3 + r] = t2; } }
    else if (k == 2) { double t = E.d0; E.d0 = E.d2; E.d2 = t;
#pragma unroll
      for (int r = 0; r < 3; ++r) { double t2 = E.q[r]; E.q[r] = E.q[6 + r]; E.q[6 + r] = t2; } }
    if (E.d2 < E.d1) { double t = E.d1; E.d1 = E.d2; E.d2 = t;
#pragma unroll
      for (int r = 0; r < 3; ++r) { double t2 = E.q[3 + r]; E.q[3 + r] = E.q[6 + r]; E.q[6 + r] = t2; } }
  }
  w[0] = E.d0 * scale; w[1] = E.d1 * scale; w[2] = E.d2 * scale;
#pragma unroll
  for (int i = 0; i < 9; ++i) Q[i] = E.q[i];
}

// ---- Eigen 3.3 ColPivHouseholderQR<Matrix<double,5,3>>::solve (same operation order as the test oracle).
// Fully unrolled over the three columns; the pivot column `big` is the only dynamic index and is handled with
// conditional swaps on constant indices, so qr[15] and the norm / permutation vectors stay in registers.
__device__ __forceinline__ void d_swap_cols(double* qr, int a, int b) {      // a, b compile-time after unrolling
#pragma unroll
  for (int r = 0; r < 5; ++r) { double t = qr[r * 3 + a]; qr[r * 3 + a] = qr[r * 3 + b]; qr[r * 3 + b] = t; }
}

__device__ __forceinline__ void d_colpiv_qr_solve_5x3(const double* Ain, const double* bin, double* x) {
  constexpr int R = 5, C = 3;
  double qr[R * C];
#pragma unroll
  for (int i = 0; i < R * C; ++i) qr[i] = Ain[i];
  double hC[C]; int transp[C];
  double nU[C], nD[C];
#pragma unroll
  for (int k = 0; k < C; ++k) {
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < R; ++r) s += qr[r * C + k] * qr[r * C + k];
    nD[k] = sqrt(s); nU[k] = nD[k];
  }
  double maxn = nU[0];
#pragma unroll
  for (int k = 1; k < C; ++k) if (nU[k] > maxn) maxn = nU[k];
  double th = maxn * DBL_EPSILON; double threshold_helper = (th * th) / (double)R;
  double norm_downdate_threshold = sqrt(DBL_EPSILON);
  int nonzero_pivots = C;
#pragma unroll
  for (int k = 0; k < C; ++k) {
    int big = k; double bn = nU[k];
#pragma unroll
    for (int j = k + 1; j < C; ++j) if (nU[j] > bn) { bn = nU[j]; big = j; }
    double biggest_sq = bn * bn;
    if (nonzero_pivots == C && biggest_sq < threshold_helper * (double)(R - k)) nonzero_pivots = k;
    transp[k] = big;
#pragma unroll
    for (int j = k + 1; j < C; ++j) {
      if (big == j) {
        d_swap_cols(qr, k, j);
        double t = nU[k]; nU[k] = nU[j]; nU[j] = t;
        t = nD[k]; nD[k] = nD[j]; nD[j] = t;
      }
    }
    // householder on column k, rows k..R-1
    double beta, tau;
    {
      double tailSqNorm = 0.0;
#pragma unroll
      for (int r = k + 1; r < R; ++r) tailSqNorm += qr[r * C + k] * qr[r * C + k];
      double c0 = qr[k * C + k];
      if (tailSqNorm <= DBL_MIN) {
        tau = 0.0; beta = c0;
#pragma unroll
        for (int r = k + 1; r < R; ++r) qr[r * C + k] = 0.0;
      } else {
        double bb = sqrt(c0 * c0 + tailSqNorm);
        if (c0 >= 0.0) bb = -bb;
#pragma unroll
        for (int r = k + 1; r < R; ++r) qr[r * C + k] = qr[r * C + k] / (c0 - bb);
        tau = (bb - c0) / bb;
        beta = bb;
      }
    }
    hC[k] = tau;
    qr[k * C + k] = beta;
    if (hC[k] != 0.0) {
#pragma unroll
      for (int j = k + 1; j < C; ++j) {
        double tmp = 0.0;
#pragma unroll
        for (int r = k + 1; r < R; ++r) tmp += qr[r * C + k] * qr[r * C + j];
        tmp += qr[k * C + j];
        qr[k * C + j] -= hC[k] * tmp;
#pragma unroll
        for (int r = k + 1; r < R; ++r) qr[r * C + j] -= hC[k] * qr[r * C + k] * tmp;
      }
    }
#pragma unroll
    for (int j = k + 1; j < C; ++j) {
      if (nU[j] != 0.0) {
        double temp = fabs(qr[k * C + j]) / nU[j];
        temp = (1.0 + temp) * (1.0 - temp);
        temp = temp < 0.0 ? 0.0 : temp;
        double ratio = nU[j] / nD[j];
        double temp2 = temp * (ratio * ratio);
        if (temp2 <= norm_downdate_threshold) {
          double s = 0.0;
#pragma unroll
          for (int r = k + 1; r < R; ++r) s += qr[r * C + j] * qr[r * C + j];
          nD[j] = sqrt(s); nU[j] = nD[j];
        } else {
          nU[j] *= sqrt(temp);
        }
      }
    }
  }
  // column permutation: perm = identity with transpositions (k, transp[k]) applied in order
  int p0 = 0, p1 = 1, p2 = 2;
  { const int t = transp[0]; if (t == 1) { int u = p0; p0 = p1; p1 = u; } else if (t == 2) { int u = p0; p0 = p2; p2 = u; } }
  { const int t = transp[1]; if (t == 2) { int u = p1; p1 = p2; p2 = u; } }
  if (nonzero_pivots == 0) { x[0] = x[1] = x[2] = 0.0; return; }
  double c[R];
#pragma unroll
  for (int r = 0; r < R; ++r) c[r] = bin[r];
#pragma unroll
  for (int k = 0; k < C; ++k) {
    if (k < nonzero_pivots && hC[k] != 0.0) {
      double tmp = 0.0;
#pragma unroll
      for (int r = k + 1; r < R; ++r) tmp += qr[r * C + k] * c[r];
      tmp += c[k];
      c[k] -= hC[k] * tmp;
#pragma unroll
      for (int r = k + 1; r < R; ++r) c[r] -= hC[k] * qr[r * C + k] * tmp;
    }
  }
#pragma unroll
  for (int i = C - 1; i >= 0; --i) {
    if (i < nonzero_pivots) {
      double s = c[i];
#pragma unroll
      for (int j = i + 1; j < C; ++j) if (j < nonzero_pivots) s -= qr[i * C + j] * c[j];
      c[i] = s / qr[i * C + i];
    }
  }
  // x[perm[i]] = c[i] for i < nonzero_pivots, 0 otherwise
  const double c0 = 0 < nonzero_pivots ? c[0] : 0.0, c1 = 1 < nonzero_pivots ? c[1] : 0.0, c2 = 2 < nonzero_pivots ? c[2] : 0.0;
#pragma unroll
  for (int t = 0; t < 3; ++t) x[t] = (p0 == t) ? c0 : ((p1 == t) ? c1 : c2);
}

// ---- fits
__device__ __forceinline__ void d_fit_corner(const float4* nb, float4 ori, LmFactor* f) {
  double near[5][3], center[3] = { 0, 0, 0 };
  for (int j = 0; j < 5; ++j) {
    near[j][0] = nb[j].x; near[j][1] = nb[j].y; near[j][2] = nb[j].z;
    for (int k = 0; k < 3; ++k) center[k] = center[k] + near[j][k];
  }
  for (int k = 0; k < 3; ++k) center[k] = center[k] / 5.0;
  double cov[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
  for (int j = 0; j < 5; ++j) {
    double z[3] = { near[j][0] - center[0], near[j][1] - center[1], near[j][2] - center[2] };
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) cov[a * 3 + b] = cov[a * 3 + b] + z[a] * z[b];
  }
  double w[3], Q[9];
  d_eigh3(cov, w, Q);
  if (!(w[2] > 3 * w[1])) { f->kind = -1; return; }
  // unit_direction = eigenvectors().col(2): Q column-major, column 2 = Q[6..8]
  for (int k = 0; k < 3; ++k) { f->a[k] = 0.1 * Q[6 + k] + center[k]; f->b[k] = -0.1 * Q[6 + k] + center[k]; }
  f->p[0] = ori.x; f->p[1] = ori.y; f->p[2] = ori.z;
  f->kind = 0;
}

__device__ __forceinline__ void d_fit_surf(const float4* nb, float4 ori, LmFactor* f) {
  double A[15], B[5] = { -1, -1, -1, -1, -1 };
  for (int j = 0; j < 5; ++j) { A[j * 3 + 0] = nb[j].x; A[j * 3 + 1] = nb[j].y; A[j * 3 + 2] = nb[j].z; }
  double n[3];
  d_colpiv_qr_solve_5x3(A, B, n);
  double z2 = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
  double nn = sqrt(z2);
  double negative_OA_dot_norm = 1 / nn;
  if (z2 > 0.0) { n[0] /= nn; n[1] /= nn; n[2] /= nn; }
  for (int j = 0; j < 5; ++j) {
    if (fabs(n[0] * A[j * 3 + 0] + n[1] * A[j * 3 + 1] + n[2] * A[j * 3 + 2] + negative_OA_dot_norm) > 0.2) { f->kind = -1; return; }
  }
  f->a[0] = n[0]; f->a[1] = n[1]; f->a[2] = n[2];
  f->b[0] = negative_OA_dot_norm; f->b[1] = 0.0; f->b[2] = 0.0;
  f->p[0] = ori.x; f->p[1] = ori.y; f->p[2] = ori.z;
  f->kind = 2;
}

// ---- GROUP-lane exact 5-NN of one world-frame query against the window of one map type.
// Phase 1: lane `sub` resolves one of the (normally 8, at cube borders up to 27) candidate cells to its
// (first element, count, concatenation offset).  Phase 2: the group scans the CONCATENATION of its cells
// cooperatively -- an 8-lane prefix sum of the counts maps a flat candidate number f to (cell, offset), lane
// `sub` takes f = sub, sub + 8, ... -- so the work is balanced over the lanes whatever the cell occupancies are and
// a lane's loads are independent of each other (the per-cell scan this replaces was a dependent chain of up to
// ~15 L2 loads on the fullest cell while the lanes of empty cells idled).  Only candidates with fp32 d2 < 1.0 can
// matter (the :584,652 gate is on the 5th neighbour), the others are dropped before the top-5 insertion.
// Candidates are ranked by the 64-bit key (d2 bits << 32 | canonical index): d2 >= 0, so the float bits order like
// the value, and ties fall back to the index in the reference's concatenation order (:533-537).
// All lanes of the group call this with the same query; every lane returns the merged top-5 (key, ref),
// missing entries are key = ~0, ref = -1.
constexpr unsigned long long KNN_NOKEY = ~0ull;

__device__ __forceinline__ void d_top5_insert(unsigned long long (&tk)[KNN_K], int (&tr)[KNN_K], unsigned long long key, int ref) {
  if (key < tk[KNN_K - 1]) {
    tk[KNN_K - 1] = key; tr[KNN_K - 1] = ref;
#pragma unroll
    for (int k = KNN_K - 1; k > 0; --k) {
      if (tk[k] < tk[k - 1]) {
        const unsigned long long t = tk[k]; tk[k] = tk[k - 1]; tk[k - 1] = t;
        const int r = tr[k]; tr[k] = tr[k - 1]; tr[k - 1] = r;
      }
    }
  }
}

template <int GROUP>
__device__ __forceinline__ void d_knn5_group(const float4* __restrict__ cellpts, const uint32_t* __restrict__ cellstart, int cap,
                                             const int2* __restrict__ slot_info, int cen0, int cen1, int cen2,
                                             float qx, float qy, float qz, int sub, unsigned gmask,
                                             unsigned long long (&tk)[KNN_K], int (&tr)[KNN_K]) {
#pragma unroll
  for (int k = 0; k < KNN_K; ++k) { tk[k] = KNN_NOKEY; tr[k] = -1; }

  // per axis: the floors f-1..f+1 fall in the two cells c0 = (f-1)>>1 and c1 = c0 + 1; a cell belongs to one cube,
  // or to two when it straddles a 50 m border (at most one of c0 / c1 does): 2 or 3 (cube, local cell) pairs,
  // kept in scalar registers (no dynamically indexed arrays -> no local memory)
  constexpr int FQ_MAX = 1 << 24;     // beyond +-16.7 km an fp32 coordinate has no sub-metre neighbours left to find; keeps the integer maths in range
  const int fq[3] = { min(max((int)floorf(qx), -FQ_MAX), FQ_MAX), min(max((int)floorf(qy), -FQ_MAX), FQ_MAX), min(max((int)floorf(qz), -FQ_MAX), FQ_MAX) };
  int e0[3], e1[3], e2[3], np[3];       // packed (cube << 8 | local cell)
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int c0 = (fq[a] - 1) >> 1, c1 = c0 + 1;
    const int lo0 = d_floordiv25(c0 + 12), hi0 = d_floordiv25(c0 + 13);
    const int lo1 = d_floordiv25(c1 + 12), hi1 = d_floordiv25(c1 + 13);
    e0[a] = lo0 * 256 + (c0 - (25 * lo0 - 13));
    if (hi0 != lo0)      { e1[a] = hi0 * 256 + (c0 - (25 * hi0 - 13)); e2[a] = lo1 * 256 + (c1 - (25 * lo1 - 13)); np[a] = 3; }
    else if (hi1 != lo1) { e1[a] = lo1 * 256 + (c1 - (25 * lo1 - 13)); e2[a] = hi1 * 256 + (c1 - (25 * hi1 - 13)); np[a] = 3; }
    else                 { e1[a] = lo1 * 256 + (c1 - (25 * lo1 - 13)); e2[a] = e1[a]; np[a] = 2; }
  }
  const int ncomb = np[0] * np[1] * np[2];
  for (int cb = 0; cb < ncomb; cb += GROUP) {        // group-uniform trip count: 1, or up to 4 next to a cube border
    const int cmb = cb + sub;
    int cnt = 0, start = 0, bidx = 0;
    if (cmb < ncomb) {
      int r = cmb;                                   // mixed radix (np[0], np[1], np[2]), every radix 2 or 3: constant divisors only
      const int ix = np[0] == 2 ? (r & 1) : (r % 3); r = np[0] == 2 ? (r >> 1) : (r / 3);
      const int iy = np[1] == 2 ? (r & 1) : (r % 3); r = np[1] == 2 ? (r >> 1) : (r / 3);
      const int iz = r;
      const int ex = ix == 0 ? e0[0] : (ix == 1 ? e1[0] : e2[0]);
      const int ey = iy == 0 ? e0[1] : (iy == 1 ? e1[1] : e2[1]);
      const int ez = iz == 0 ? e0[2] : (iz == 1 ? e1[2] : e2[2]);
      const int gi = ex >> 8, ci = ex & 255, gj = ey >> 8, cj = ey & 255, gk = ez >> 8, ck = ez & 255;
      const int li = gi + cen0, lj = gj + cen1, lk = gk + cen2;
      if (li >= 0 && li < LM_GW && lj >= 0 && lj < LM_GH && lk >= 0 && lk < LM_GD) {
        const int2 inf = slot_info[d_phys_slot(gi, gj, gk)];       // {slab, concatenation offset}; slab < 0: not in the window / empty
        if (inf.x >= 0) {
          const uint32_t* cs = cellstart + (size_t)inf.x * (LM_NCELL + 1) + (ci + LM_CELLS_AXIS * (cj + LM_CELLS_AXIS * ck));
          const uint32_t b = cs[0], e = cs[1];
          cnt = (int)(e - b); start = inf.x * cap + (int)b; bidx = inf.y;
        }
      }
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < GROUP; o <<= 1) { const int t = __shfl_up_sync(gmask, incl, o, GROUP); if (sub >= o) incl += t; }
    const int total = __shfl_sync(gmask, incl, GROUP - 1, GROUP);
    start -= incl - cnt;                             // flat candidate f of this lane's cell lives at cellpts[start + f]
    int bound[GROUP > 1 ? GROUP - 1 : 1];
#pragma unroll
    for (int c = 0; c < GROUP - 1; ++c) bound[c] = __shfl_sync(gmask, incl, c, GROUP);
    for (int fb = 0; fb < total; fb += 2 * GROUP) {  // two independent candidates per lane and trip
      const int fA = fb + sub, fB = fA + GROUP;
      int cellA = 0, cellB = 0;
#pragma unroll
      for (int c = 0; c < GROUP - 1; ++c) { cellA += fA >= bound[c]; cellB += fB >= bound[c]; }
      const int sA = __shfl_sync(gmask, start, cellA, GROUP), bA = __shfl_sync(gmask, bidx, cellA, GROUP);
      const int sB = __shfl_sync(gmask, start, cellB, GROUP), bB = __shfl_sync(gmask, bidx, cellB, GROUP);
      const bool vA = fA < total, vB = fB < total;
      const int rA = vA ? sA + fA : 0, rB = vB ? sB + fB : 0;
      const float4 pA = cellpts[rA];
      const float4 pB = cellpts[rB];
      {
        const float dx = __fsub_rn(qx, pA.x), dy = __fsub_rn(qy, pA.y), dz = __fsub_rn(qz, pA.z);
        float d = __fmul_rn(dx, dx);
        d = __fadd_rn(d, __fmul_rn(dy, dy));
        d = __fadd_rn(d, __fmul_rn(dz, dz));
        if (vA && d < 1.0f)
          d_top5_insert(tk, tr, ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)(bA + __float_as_int(pA.w)), rA);
      }
      {
        const float dx = __fsub_rn(qx, pB.x), dy = __fsub_rn(qy, pB.y), dz = __fsub_rn(qz, pB.z);
        float d = __fmul_rn(dx, dx);
        d = __fadd_rn(d, __fmul_rn(dy, dy));
        d = __fadd_rn(d, __fmul_rn(dz, dz));
        if (vB && d < 1.0f)
          d_top5_insert(tk, tr, ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)(bB + __float_as_int(pB.w)), rB);
      }
    }
  }
  // merge the GROUP sorted lists: 5 rounds of (min over the lane heads, the owner pops); canonical indices are unique,
  // so exactly one lane holds the minimum
  unsigned long long ok_[KNN_K]; int or_[KNN_K];
  int head = 0;
#pragma unroll
  for (int k = 0; k < KNN_K; ++k) {
    unsigned long long hk = KNN_NOKEY; int hr = -1;
#pragma unroll
    for (int j = 0; j < KNN_K; ++j) if (head == j) { hk = tk[j]; hr = tr[j]; }
    unsigned long long mk = hk;
#pragma unroll
    for (int o = GROUP / 2; o > 0; o >>= 1) {
      const unsigned long long t = __shfl_xor_sync(gmask, mk, o);
      mk = t < mk ? t : mk;
    }
    const bool mine = hk == mk && mk != KNN_NOKEY;
    const unsigned owners = __ballot_sync(gmask, mine);
    const int mr = __shfl_sync(gmask, hr, owners ? __ffs(owners) - 1 : 0);
    if (mine) head++;
    ok_[k] = mk; or_[k] = owners ? mr : -1;
  }
#pragma unroll
  for (int k = 0; k < KNN_K; ++k) { tk[k] = ok_[k]; tr[k] = or_[k]; }
}

// ---- one-thread-per-query exact 5-NN (throughput form).  The 8-lane search above replicates the cell arithmetic in
// every lane and pays a divergent top-5 insertion per candidate trip (342 warp instructions per query); here a thread
// owns a query, walks the candidate runs itself -- the two x-adjacent cells of a cube are ONE contiguous run of the
// cell-sorted copy, so the usual query has 4 runs, not 8 cells -- and only PARKS the few candidates with d2 < 1.0
// (about one in five) in a shared-memory column; the ordered top-5 insertion runs afterwards over the parked ones
// only.  Same candidate set, same 64-bit (d2, canonical index) ranking, hence the same neighbours bit for bit.
// Per axis: c0 = (floor(q) - 1) >> 1, cube q0 = floor((c0 + 12) / 25), r0 = c0 + 12 - 25 q0 in [0, 24]; cell c0 is local
// r0 + 1 of cube q0 (and ALSO local 0 of cube q0 + 1 when r0 == 24), c1 = c0 + 1 likewise.
#ifndef LM_KNN1_THREADS
#define LM_KNN1_THREADS 128
#endif
constexpr int KNN1_THREADS = LM_KNN1_THREADS;
constexpr int KNN1_SLOTS = 12;

__device__ __forceinline__ void d_axis_split(float q, int* q0, int* r0) {
  constexpr int FQ_MAX = 1 << 24;
  const int f = min(max((int)floorf(q), -FQ_MAX), FQ_MAX);
  const int c0 = (f - 1) >> 1;
  *q0 = d_floordiv25(c0 + 12);
  *r0 = c0 + 12 - 25 * *q0;
}
// entry e (0 .. n-1, n = 2 or 3) of the (cube, local cell) pairs covering cells c0, c1 on one axis
__device__ __forceinline__ void d_axis_entry(int q0, int r0, int e, int* cube, int* cell) {
  if (r0 <= 22) { *cube = q0; *cell = r0 + 1 + e; }
  else { const int s = r0 - 23 + e; *cube = q0 + (s >= 2 ? 1 : 0); *cell = s < 2 ? 24 + s : s - 2; }   // (q0,24) (q0,25) (q0+1,0) (q0+1,1)
}
// x axis: contiguous runs (cube, first local cell, cells): one run of 2 cells, or 2 + 1 / 1 + 2 next to a cube border
__device__ __forceinline__ void d_axis_run(int q0, int r0, int e, int* cube, int* cell, int* len) {
  if (r0 <= 22) { *cube = q0; *cell = r0 + 1; *len = 2; }
  else if (r0 == 23) { *cube = q0 + e; *cell = e ? 0 : 24; *len = e ? 1 : 2; }
  else { *cube = q0 + e; *cell = e ? 0 : 25; *len = e ? 2 : 1; }
}

__device__ __forceinline__ void d_knn5_thread(const float4* __restrict__ cellpts, const uint32_t* __restrict__ cellstart, int cap,
                                              const int2* __restrict__ slot_info, int cen0, int cen1, int cen2,
                                              float qx, float qy, float qz,
                                              unsigned long long* __restrict__ s_key, int* __restrict__ s_ref,   // this thread's column, stride KNN1_THREADS
                                              unsigned long long (&tk)[KNN_K], int (&tr)[KNN_K]) {
#pragma unroll
  for (int k = 0; k < KNN_K; ++k) { tk[k] = KNN_NOKEY; tr[k] = -1; }
  int qx0, rx0, qy0, ry0, qz0, rz0;
  d_axis_split(qx, &qx0, &rx0); d_axis_split(qy, &qy0, &ry0); d_axis_split(qz, &qz0, &rz0);
  const int nrx = rx0 <= 22 ? 1 : 2, npy = ry0 <= 22 ? 2 : 3, npz = rz0 <= 22 ? 2 : 3;
  const int nyz = npy * npz, nruns = nrx * nyz;
  int npark = 0;
  for (int rb = 0; rb < nruns; rb += 4) {            // 4 runs per round (the usual query has exactly 4): their table look-ups overlap
    int start[4], cnt[4], bidx[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = rb + u;
      cnt[u] = 0; start[u] = 0; bidx[u] = 0;
      if (r < nruns) {
        const int ex = nrx == 1 ? 0 : (r >= nyz ? 1 : 0);
        const int ryz = r - ex * nyz;
        const int ez = npy == 2 ? (ryz >> 1) : (ryz / 3), ey = ryz - ez * npy;
        int gi, ci, len, gj, cj, gk, ck;
        d_axis_run(qx0, rx0, ex, &gi, &ci, &len);
        d_axis_entry(qy0, ry0, ey, &gj, &cj);
        d_axis_entry(qz0, rz0, ez, &gk, &ck);
        const int li = gi + cen0, lj = gj + cen1, lk = gk + cen2;
        if (li >= 0 && li < LM_GW && lj >= 0 && lj < LM_GH && lk >= 0 && lk < LM_GD) {
          const int2 inf = slot_info[d_phys_slot(gi, gj, gk)];
          if (inf.x >= 0) {
            const uint32_t* cs = cellstart + (size_t)inf.x * (LM_NCELL + 1) + (ci + LM_CELLS_AXIS * (cj + LM_CELLS_AXIS * ck));
            const uint32_t b = cs[0], e = cs[len];
            cnt[u] = (int)(e - b); start[u] = inf.x * cap + (int)b; bidx[u] = inf.y;
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4* __restrict__ run = cellpts + start[u];
      const int n = cnt[u], bi = bidx[u];
      for (int j = 0; j < n; j += 4) {               // four independent loads in flight per trip (the scan is a latency chain)
        float4 p[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) p[v] = run[min(j + v, n - 1)];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float dx = __fsub_rn(qx, p[v].x), dy = __fsub_rn(qy, p[v].y), dz = __fsub_rn(qz, p[v].z);
          const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
          if (j + v < n && d < 1.0f) {
            const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)(bi + __float_as_int(p[v].w));
            if (npark < KNN1_SLOTS) { s_key[npark * KNN1_THREADS] = key; s_ref[npark * KNN1_THREADS] = start[u] + j + v; ++npark; }
            else d_top5_insert(tk, tr, key, start[u] + j + v);
          }
        }
      }
    }
  }
  for (int i = 0; i < npark; ++i) d_top5_insert(tk, tr, s_key[i * KNN1_THREADS], s_ref[i * KNN1_THREADS]);
}

__global__ void __launch_bounds__(KNN1_THREADS) k_assoc_knn1(LmMapState* __restrict__ st, LmMapType M0, LmMapType M1, const int32_t* __restrict__ slot_valid_rank,
                                                             const float4* __restrict__ stack0, const float4* __restrict__ stack1, int32_t* __restrict__ nnref) {
  lm_pdl_enter();
  __shared__ unsigned long long s_key[KNN1_SLOTS * KNN1_THREADS];
  __shared__ int s_ref[KNN1_SLOTS * KNN1_THREADS];
  if (!st->optimize) return;
  const int n0 = st->stack_n[0], n1 = st->stack_n[1];
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n0 + n1) return;
  const int ty = gid < n0 ? 0 : 1;
  const int qi = ty == 0 ? gid : gid - n0;
  const float4 sel = d_associate(st->q_w_curr, st->t_w_curr, ty == 0 ? stack0[qi] : stack1[qi]);
  int32_t* out = nnref + (size_t)gid * KNN_K;
  if (st->shard_n > 1 &&
      lm_cube_owner(d_cube_coord((double)sel.x, 0), d_cube_coord((double)sel.y, 0), d_cube_coord((double)sel.z, 0), st->shard_n) != st->shard_rank) {
    out[0] = -1;
    return;
  }
  unsigned long long key[KNN_K]; int ref[KNN_K];
  d_knn5_thread(ty == 0 ? M0.cellpts : M1.cellpts, ty == 0 ? M0.cellstart : M1.cellstart, ty == 0 ? M0.cap : M1.cap,
                lm_slot_info(slot_valid_rank) + ty * LM_NSLOT, st->cen[0], st->cen[1], st->cen[2], sel.x, sel.y, sel.z,
                s_key + threadIdx.x, s_ref + threadIdx.x, key, ref);
  const bool ok = ref[KNN_K - 1] >= 0;       // every kept candidate has d2 < 1.0: the :584,652 gate is "a 5th neighbour exists"
#pragma unroll
  for (int k = 0; k < KNN_K; ++k) out[k] = ok ? ref[k] : -1;
}

// Association = two launches.  k_assoc_knn: one GROUP-lane group per query (both map types in one launch), light on
// registers so every query of a sweep is resident at once; it leaves the 5 neighbour references (or -1 when the
// d2[4] < 1.0 gate of :584,652 fails).  k_assoc_fit: one thread per query for the fp64 line / plane fit -- with the
// fits inside the search kernel all but one lane of a group idled through the fp64 tail and its registers halved occupancy.
template <int GROUP>
__device__ __forceinline__ void d_assoc_knn(LmMapState* __restrict__ st, const LmMapType& M0, const LmMapType& M1,
                                            const int32_t* __restrict__ slot_valid_rank,
                                            const float4* __restrict__ stack0, const float4* __restrict__ stack1,
                                            int32_t* __restrict__ nnref) {
  if (!st->optimize) return;
  const int n0 = st->stack_n[0], n1 = st->stack_n[1];
  const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / GROUP;
  const int sub = threadIdx.x & (GROUP - 1);
  const unsigned gmask = ((1u << GROUP) - 1u) << ((threadIdx.x & 31) & ~(GROUP - 1));
  if (gid >= n0 + n1) return;
  const int ty = gid < n0 ? 0 : 1;
  const int qi = ty == 0 ? gid : gid - n0;
  const float4 ori = ty == 0 ? stack0[qi] : stack1[qi];
  const float4 sel = d_associate(st->q_w_curr, st->t_w_curr, ori);
  int32_t* out = nnref + (size_t)gid * KNN_K;
  if (st->shard_n > 1) {   // cube-sharded map: a query belongs to the rank that owns the cube it falls in
    if (lm_cube_owner(d_cube_coord((double)sel.x, 0), d_cube_coord((double)sel.y, 0), d_cube_coord((double)sel.z, 0), st->shard_n) != st->shard_rank) {
      if (sub == 0) out[0] = -1;
      return;
    }
  }
  unsigned long long key[KNN_K]; int ref[KNN_K];
  d_knn5_group<GROUP>(ty == 0 ? M0.cellpts : M1.cellpts, ty == 0 ? M0.cellstart : M1.cellstart, ty == 0 ? M0.cap : M1.cap,
               lm_slot_info(slot_valid_rank) + ty * LM_NSLOT, st->cen[0], st->cen[1], st->cen[2],
               sel.x, sel.y, sel.z, sub, gmask, key, ref);
  // every kept candidate has d2 < 1.0, so the :584,652 gate is "a 5th neighbour exists"
  const bool ok = ref[KNN_K - 1] >= 0;
  if (GROUP >= 8) {          // lanes 0..4 store one reference each
    int mine = ref[0];
#pragma unroll
    for (int k = 1; k < KNN_K; ++k) if (sub == k) mine = ref[k];
    if (sub < KNN_K) out[sub] = ok ? mine : -1;
  } else if (sub == 0) {
#pragma unroll
    for (int k = 0; k < KNN_K; ++k) out[k] = ok ? ref[k] : -1;
  }
}
template <int GROUP>
__global__ void __launch_bounds__(256, 4) k_assoc_knn(LmMapState* __restrict__ st, LmMapType M0, LmMapType M1, const int32_t* __restrict__ slot_valid_rank,
                                                   const float4* __restrict__ stack0, const float4* __restrict__ stack1, int32_t* __restrict__ nnref) {
  lm_pdl_enter();
  d_assoc_knn<GROUP>(st, M0, M1, slot_valid_rank, stack0, stack1, nnref);
}

#ifndef LM_FIT_THREADS
#define LM_FIT_THREADS 128
#endif
__global__ void __launch_bounds__(LM_FIT_THREADS) k_assoc_fit(const LmMapState* __restrict__ st, LmMapType M0, LmMapType M1,
                                                   const float4* __restrict__ stack0, const float4* __restrict__ stack1,
                                                   const int32_t* __restrict__ nnref,
                                                   LmFactor* __restrict__ fac0, LmFactor* __restrict__ fac1) {
  lm_pdl_enter();
  if (!st->optimize) return;
  const int n0 = st->stack_n[0], n1 = st->stack_n[1];
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n0 + n1) return;
  const int ty = gid < n0 ? 0 : 1;
  const int qi = ty == 0 ? gid : gid - n0;
  LmFactor* f = (ty == 0 ? fac0 : fac1) + qi;
  const int32_t* r = nnref + (size_t)gid * KNN_K;
  if (r[0] < 0) { f->kind = -1; return; }
  const float4* __restrict__ cp = ty == 0 ? M0.cellpts : M1.cellpts;
  float4 nb[KNN_K];
#pragma unroll
  for (int k = 0; k < KNN_K; ++k) nb[k] = cp[r[k]];
  const float4 ori = ty == 0 ? stack0[qi] : stack1[qi];
  LmFactor out;
  if (ty == 0) d_fit_corner(nb, ori, &out); else d_fit_surf(nb, ori, &out);
  if (out.kind < 0) { f->kind = -1; return; }
  *f = out;
}

int lm_map_associate(lmono_ctx* ctx, int n_max_corner, int n_max_surf) {
  const int nq = n_max_corner + n_max_surf;
  if (nq <= 0) return LMONO_OK;
  // latency form (8 lanes per query) for a sequence alone on the GPU, throughput form (one thread per query, a third of
  // the instructions) when the step is one of several running side by side; LMONO_KNN_GROUP overrides (0 = thread form)
  static const int group_env = getenv("LMONO_KNN_GROUP") ? atoi(getenv("LMONO_KNN_GROUP")) : -1;
  const int group = group_env >= 0 ? group_env : (ctx->batch_n >= LM_THROUGHPUT_BATCH ? 0 : GROUP_DEFAULT);
#define KNN_ARGS ctx->d_state, ctx->map[0], ctx->map[1], ctx->d_slot_valid_rank, ctx->d_stack[0], ctx->d_stack[1], ctx->d_nnref
  if (group == 0) {
    LM_LAUNCH_PDL(k_assoc_knn1, lm_div_up(nq, KNN1_THREADS), KNN1_THREADS, 0, KNN_ARGS);
    LM_LAUNCH_CHECK();
  } else {
  if (group == 1) LM_LAUNCH_PDL(k_assoc_knn<1>, lm_div_up(nq * 1, 256), 256, 0, KNN_ARGS);
  else if (group == 2) LM_LAUNCH_PDL(k_assoc_knn<2>, lm_div_up(nq * 2, 256), 256, 0, KNN_ARGS);
  else if (group == 4) LM_LAUNCH_PDL(k_assoc_knn<4>, lm_div_up(nq * 4, 256), 256, 0, KNN_ARGS);
  else LM_LAUNCH_PDL(k_assoc_knn<8>, lm_div_up(nq * 8, 256), 256, 0, KNN_ARGS);
  LM_LAUNCH_CHECK();
  }
  LM_LAUNCH_PDL(k_assoc_fit, lm_div_up(nq, LM_FIT_THREADS), LM_FIT_THREADS, 0, ctx->d_state, ctx->map[0], ctx->map[1], ctx->d_stack[0], ctx->d_stack[1],
                                                          ctx->d_nnref, ctx->d_fac[0], ctx->d_fac[1]);
  LM_LAUNCH_CHECK();
  return LMONO_OK;
}

// test hook: world-frame queries, 5-NN output restricted to the neighbours the gate can see (d2 < 1.0); entries
// beyond that are idx = -1, d2 = +inf
constexpr int GROUP = GROUP_DEFAULT;
__global__ void __launch_bounds__(256) k_knn5_hook(const LmMapState* __restrict__ st, LmMapType M, int ty,
                                                   const int32_t* __restrict__ slot_valid_rank,
                                                   const float4* __restrict__ q, int n, int32_t* __restrict__ oidx, float* __restrict__ od2) {
  const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / GROUP;
  const int sub = threadIdx.x & (GROUP - 1);
  const unsigned gmask = ((1u << GROUP) - 1u) << ((threadIdx.x & 31) & ~(GROUP - 1));
  if (gid >= n) return;
  const float4 p = q[gid];
  unsigned long long key[KNN_K]; int ref[KNN_K];
  d_knn5_group<GROUP>(M.cellpts, M.cellstart, M.cap, lm_slot_info(slot_valid_rank) + ty * LM_NSLOT, st->cen[0], st->cen[1], st->cen[2],
               p.x, p.y, p.z, sub, gmask, key, ref);
  if (sub != 0) return;
#pragma unroll
  for (int k = 0; k < KNN_K; ++k) {
    const bool ok = ref[k] >= 0;
    oidx[gid * KNN_K + k] = ok ? (int)(unsigned)(key[k] & 0xffffffffull) : -1;
    od2[gid * KNN_K + k] = ok ? __uint_as_float((unsigned)(key[k] >> 32)) : INFINITY;
  }
}

__global__ void __launch_bounds__(KNN1_THREADS) k_knn5_hook1(const LmMapState* __restrict__ st, LmMapType M, int ty,
                                                             const int32_t* __restrict__ slot_valid_rank,
                                                             const float4* __restrict__ q, int n, int32_t* __restrict__ oidx, float* __restrict__ od2) {
  __shared__ unsigned long long s_key[KNN1_SLOTS * KNN1_THREADS];
  __shared__ int s_ref[KNN1_SLOTS * KNN1_THREADS];
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n) return;
  const float4 p = q[gid];
  unsigned long long key[KNN_K]; int ref[KNN_K];
  d_knn5_thread(M.cellpts, M.cellstart, M.cap, lm_slot_info(slot_valid_rank) + ty * LM_NSLOT, st->cen[0], st->cen[1], st->cen[2],
                p.x, p.y, p.z, s_key + threadIdx.x, s_ref + threadIdx.x, key, ref);
#pragma unroll
  for (int k = 0; k < KNN_K; ++k) {
    const bool ok = ref[k] >= 0;
    oidx[gid * KNN_K + k] = ok ? (int)(unsigned)(key[k] & 0xffffffffull) : -1;
    od2[gid * KNN_K + k] = ok ? __uint_as_float((unsigned)(key[k] >> 32)) : INFINITY;
  }
}

int lm_knn5_device(lmono_ctx* ctx, int which, const float4* d_q, int n, int32_t* d_idx, float* d_d2) {
  LM_NEED_MAP();
  if (n <= 0) return LMONO_OK;
  // same choice of form as lm_map_associate: LMONO_KNN_GROUP=0, or a ctx that runs the throughput forms
  const char* ge = getenv("LMONO_KNN_GROUP");
  if (ge ? atoi(ge) == 0 : ctx->batch_n >= LM_THROUGHPUT_BATCH) {
    k_knn5_hook1<<<lm_div_up(n, KNN1_THREADS), KNN1_THREADS, 0, ctx->stream>>>(ctx->d_state, ctx->map[which], which, ctx->d_slot_valid_rank, d_q, n, d_idx, d_d2);
    LM_LAUNCH_CHECK();
    return LMONO_OK;
  }
  const int blocks = lm_div_up(n * GROUP, 256);
  k_knn5_hook<<<blocks, 256, 0, ctx->stream>>>(ctx->d_state, ctx->map[which], which, ctx->d_slot_valid_rank, d_q, n, d_idx, d_d2);
  LM_LAUNCH_CHECK();
  return LMONO_OK;
}
