/*
 * ORACLE (test infrastructure): restatement of the main loop body of
 * Aloam/src/laserOdometry.cpp:265-568 -- scan-to-scan registration: TransformToStart
 * (:111-129, DISTORTION 0 so s = 1), 1-NN in the previous sweep's less-sharp / less-flat
 * clouds, +-NEARBY_SCAN ring searches for the 2nd / 3rd correspondence (:299-483), two
 * ceres::Solve passes of 4 iterations (:278-501), pose integration (:504-505) and the
 * swap of the "last" clouds (:554-568).
 */
#include "lmono_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now_ms(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }

struct o_odom {
  int inited;
  double para_q[4], para_t[3];       /* :97-98 q_last_curr (x,y,z,w), t_last_curr */
  double q_w_curr[4], t_w_curr[3];   /* :93-94 */
  o_pt* corner_last; int n_corner_last;
  o_pt* surf_last; int n_surf_last;
  o_kdtree* kd_corner; o_kdtree* kd_surf;
  int use_kdtree;
  int distortion;                    /* #define DISTORTION :59 */
};

static void qmul(const double a[4], const double b[4], double o[4]) {
  double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
static void qrot(const double q[4], const double v[3], double o[3]) {
  double uv[3] = { q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0] };
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  double c[3] = { q[1] * uv[2] - q[2] * uv[1], q[2] * uv[0] - q[0] * uv[2], q[0] * uv[1] - q[1] * uv[0] };
  for (int k = 0; k < 3; ++k) o[k] = (v[k] + q[3] * uv[k]) + c[k];
}

o_odom* lmono_cpu_odom_create(void) {
  o_odom* o = (o_odom*)calloc(1, sizeof(o_odom));
  o->para_q[3] = 1.0; o->q_w_curr[3] = 1.0;
  o->use_kdtree = 1;
  return o;
}
void lmono_cpu_odom_set_distortion(o_odom* o, int on) { if (o) o->distortion = on ? 1 : 0; }
void lmono_cpu_odom_destroy(o_odom* o) {
  if (!o) return;
  free(o->corner_last); free(o->surf_last);
  lmono_cpu_kdtree_free(o->kd_corner); lmono_cpu_kdtree_free(o->kd_surf);
  free(o);
}

void lmono_cpu_slerp_identity(double t, const double q[4], double qs[4], double dqs[4][4]);

/* interpolation ratio of a point, :114-118 / :375-379: (intensity - int(intensity)) is FLOAT arithmetic (float - int), the
 * division by SCAN_PERIOD = 0.1 double */
static double ratio_of(const o_pt* p, int distortion) {
  if (!distortion) return 1.0;
  const float frac = p->i - (float)(int)p->i;
  return (double)frac / 0.1;
}

/* TransformToStart :111-129.  s = 1 (DISTORTION 0): Identity.slerp(1, q) is q or -q (same rotation, bit-identical result
 * through Eigen's formula), t_point_last = 1.0 * t.  Otherwise q_point_last = Identity.slerp(s, q_last_curr) (not
 * renormalised), t_point_last = s * t_last_curr. */
static void transform_to_start_s(const double q[4], const double t[3], const o_pt* pi, o_pt* po, double s) {
  double p[3] = { pi->x, pi->y, pi->z }, r[3];
  if (s == 1.0) {
    qrot(q, p, r);
    po->x = (float)(r[0] + t[0]); po->y = (float)(r[1] + t[1]); po->z = (float)(r[2] + t[2]);
  } else {
    double qs[4];
    lmono_cpu_slerp_identity(s, q, qs, NULL);
    qrot(qs, p, r);
    po->x = (float)(r[0] + s * t[0]); po->y = (float)(r[1] + s * t[1]); po->z = (float)(r[2] + s * t[2]);
  }
  po->i = pi->i;
}
static void transform_to_start(const double q[4], const double t[3], const o_pt* pi, o_pt* po) { transform_to_start_s(q, t, pi, po, 1.0); }

static inline double sqdis(const o_pt* a, const o_pt* sel) {
  /* :322-327: float products and sums, widened on assignment */
  float d = (a->x - sel->x) * (a->x - sel->x) + (a->y - sel->y) * (a->y - sel->y) + (a->z - sel->z) * (a->z - sel->z);
  return (double)d;
}

static void nn1(const o_kdtree* kd, const o_pt* pts, int n, const o_pt* q, int32_t* idx, float* d2) {
  if (kd) lmono_cpu_kdtree_knn(kd, q, 1, 1, idx, d2);
  else lmono_cpu_knn_brute(pts, n, q, 1, 1, idx, d2);
}

/* :299-362 */
static void corner_corr(const o_pt* cl, int ncl, const o_kdtree* kd, const o_pt* sel, int* closest, int* min2) {
  const double DISTANCE_SQ_THRESHOLD = 25, NEARBY_SCAN = 2.5;
  *closest = -1; *min2 = -1;
  if (ncl <= 0) return;
  int32_t idx; float d2;
  nn1(kd, cl, ncl, sel, &idx, &d2);
  if (d2 < DISTANCE_SQ_THRESHOLD) {
    int closestPointInd = idx;
    int closestPointScanID = (int)cl[closestPointInd].i;
    double minPointSqDis2 = DISTANCE_SQ_THRESHOLD;
    int minPointInd2 = -1;
    for (int j = closestPointInd + 1; j < ncl; ++j) {
      if ((int)cl[j].i <= closestPointScanID) continue;
      if ((int)cl[j].i > (closestPointScanID + NEARBY_SCAN)) break;
      double pointSqDis = sqdis(&cl[j], sel);
      if (pointSqDis < minPointSqDis2) { minPointSqDis2 = pointSqDis; minPointInd2 = j; }
    }
    for (int j = closestPointInd - 1; j >= 0; --j) {
      if ((int)cl[j].i >= closestPointScanID) continue;
      if ((int)cl[j].i < (closestPointScanID - NEARBY_SCAN)) break;
      double pointSqDis = sqdis(&cl[j], sel);
      if (pointSqDis < minPointSqDis2) { minPointSqDis2 = pointSqDis; minPointInd2 = j; }
    }
    *closest = closestPointInd; *min2 = minPointInd2;
  }
}

/* :387-455 */
static void plane_corr(const o_pt* sl, int nsl, const o_kdtree* kd, const o_pt* sel, int* closest, int* min2, int* min3) {
  const double DISTANCE_SQ_THRESHOLD = 25, NEARBY_SCAN = 2.5;
  *closest = -1; *min2 = -1; *min3 = -1;
  if (nsl <= 0) return;
  int32_t idx; float d2;
  nn1(kd, sl, nsl, sel, &idx, &d2);
  if (d2 < DISTANCE_SQ_THRESHOLD) {
    int closestPointInd = idx;
    int closestPointScanID = (int)sl[closestPointInd].i;
    double minPointSqDis2 = DISTANCE_SQ_THRESHOLD, minPointSqDis3 = DISTANCE_SQ_THRESHOLD;
    int minPointInd2 = -1, minPointInd3 = -1;
    for (int j = closestPointInd + 1; j < nsl; ++j) {
      if ((int)sl[j].i > (closestPointScanID + NEARBY_SCAN)) break;
      double pointSqDis = sqdis(&sl[j], sel);
      if ((int)sl[j].i <= closestPointScanID && pointSqDis < minPointSqDis2) { minPointSqDis2 = pointSqDis; minPointInd2 = j; }
      else if ((int)sl[j].i > closestPointScanID && pointSqDis < minPointSqDis3) { minPointSqDis3 = pointSqDis; minPointInd3 = j; }
    }
    for (int j = closestPointInd - 1; j >= 0; --j) {
      if ((int)sl[j].i < (closestPointScanID - NEARBY_SCAN)) break;
      double pointSqDis = sqdis(&sl[j], sel);
      if ((int)sl[j].i >= closestPointScanID && pointSqDis < minPointSqDis2) { minPointSqDis2 = pointSqDis; minPointInd2 = j; }
      else if ((int)sl[j].i < closestPointScanID && pointSqDis < minPointSqDis3) { minPointSqDis3 = pointSqDis; minPointInd3 = j; }
    }
    *closest = closestPointInd; *min2 = minPointInd2; *min3 = minPointInd3;
  }
}

int lmono_cpu_odom_associate(const o_pt* sharp, int n_sharp, const o_pt* flat, int n_flat,
                             const o_pt* corner_last, int n_cl, const o_pt* surf_last, int n_sl,
                             const o_pose* last_curr, int32_t* corner_idx, int32_t* plane_idx) {
  o_kdtree* kc = lmono_cpu_kdtree_build(corner_last, n_cl);
  o_kdtree* ks = lmono_cpu_kdtree_build(surf_last, n_sl);
  for (int i = 0; i < n_sharp; ++i) {
    o_pt sel; transform_to_start(last_curr->q, last_curr->t, &sharp[i], &sel);
    int a, b; corner_corr(corner_last, n_cl, kc, &sel, &a, &b);
    corner_idx[i * 2] = a; corner_idx[i * 2 + 1] = b;
  }
  for (int i = 0; i < n_flat; ++i) {
    o_pt sel; transform_to_start(last_curr->q, last_curr->t, &flat[i], &sel);
    int a, b, c; plane_corr(surf_last, n_sl, ks, &sel, &a, &b, &c);
    plane_idx[i * 3] = a; plane_idx[i * 3 + 1] = b; plane_idx[i * 3 + 2] = c;
  }
  lmono_cpu_kdtree_free(kc); lmono_cpu_kdtree_free(ks);
  return 0;
}

int lmono_cpu_odom_step(o_odom* o, const o_pt* sharp, int n_sharp, const o_pt* less_sharp, int n_less_sharp,
                        const o_pt* flat, int n_flat, const o_pt* less_flat, int n_less_flat,
                        o_pose* last_curr, o_pose* w_curr, o_odom_report* rep) {
  o_odom_report R; memset(&R, 0, sizeof(R));
  double t_whole = now_ms();
  if (!o->inited) {
    o->inited = 1;                      /* :267-271 */
    R.inited = 0;
  } else {
    R.inited = 1;
    int cap = n_sharp + n_flat;
    o_factor* fac = (o_factor*)malloc((size_t)(cap > 0 ? cap : 1) * sizeof(o_factor));
    for (int opti = 0; opti < 2; ++opti) {     /* :278 */
      double ta = now_ms();
      int nf = 0, cc = 0, pc = 0;
      for (int i = 0; i < n_sharp; ++i) {
        const double s = ratio_of(&sharp[i], o->distortion);
        o_pt sel; transform_to_start_s(o->para_q, o->para_t, &sharp[i], &sel, s);
        int a, b; corner_corr(o->corner_last, o->n_corner_last, o->kd_corner, &sel, &a, &b);
        if (b >= 0) {                                   /* :363 */
          o_factor* f = &fac[nf++];
          f->type = O_FACTOR_EDGE; f->pad = 0; f->s = s;
          f->p[0] = sharp[i].x; f->p[1] = sharp[i].y; f->p[2] = sharp[i].z;
          f->a[0] = o->corner_last[a].x; f->a[1] = o->corner_last[a].y; f->a[2] = o->corner_last[a].z;
          f->b[0] = o->corner_last[b].x; f->b[1] = o->corner_last[b].y; f->b[2] = o->corner_last[b].z;
          cc++;
        }
      }
      for (int i = 0; i < n_flat; ++i) {
        const double s = ratio_of(&flat[i], o->distortion);
        o_pt sel; transform_to_start_s(o->para_q, o->para_t, &flat[i], &sel, s);
        int a, b, c; plane_corr(o->surf_last, o->n_surf_last, o->kd_surf, &sel, &a, &b, &c);
        if (b >= 0 && c >= 0) {                         /* :457 */
          o_factor* f = &fac[nf++];
          f->type = O_FACTOR_PLANE; f->pad = 0; f->s = s;
          f->p[0] = flat[i].x; f->p[1] = flat[i].y; f->p[2] = flat[i].z;
          const o_pt *pj = &o->surf_last[a], *pl = &o->surf_last[b], *pm = &o->surf_last[c];
          double j[3] = { pj->x, pj->y, pj->z }, l[3] = { pl->x, pl->y, pl->z }, m[3] = { pm->x, pm->y, pm->z };
          /* lidarFactor.hpp:64-65 ljm_norm = (j - l).cross(j - m); normalize() */
          double u[3] = { j[0] - l[0], j[1] - l[1], j[2] - l[2] }, v[3] = { j[0] - m[0], j[1] - m[1], j[2] - m[2] };
          double nrm[3] = { u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0] };
          double z2 = nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2];
          if (z2 > 0.0) { double nn = sqrt(z2); nrm[0] /= nn; nrm[1] /= nn; nrm[2] /= nn; }
          for (int k = 0; k < 3; ++k) { f->a[k] = j[k]; f->b[k] = nrm[k]; }
          pc++;
        }
      }
      R.corner_corr[opti] = cc; R.plane_corr[opti] = pc;
      R.ms_assoc += now_ms() - ta;
      double ts = now_ms();
      o_pose x; memcpy(x.q, o->para_q, sizeof(x.q)); memcpy(x.t, o->para_t, sizeof(x.t));
      lmono_cpu_lm_solve(fac, nf, &x, 4, &R.solve[opti]);       /* :494-499 */
      memcpy(o->para_q, x.q, sizeof(x.q)); memcpy(o->para_t, x.t, sizeof(x.t));
      R.ms_solver += now_ms() - ts;
    }
    free(fac);
    /* :504-505 */
    double tmp[3]; qrot(o->q_w_curr, o->para_t, tmp);
    for (int k = 0; k < 3; ++k) o->t_w_curr[k] = o->t_w_curr[k] + tmp[k];
    double qn[4]; qmul(o->q_w_curr, o->para_q, qn);
    memcpy(o->q_w_curr, qn, sizeof(qn));
  }
  /* :554-568 */
  free(o->corner_last); free(o->surf_last);
  o->corner_last = (o_pt*)malloc((size_t)(n_less_sharp > 0 ? n_less_sharp : 1) * sizeof(o_pt));
  o->surf_last = (o_pt*)malloc((size_t)(n_less_flat > 0 ? n_less_flat : 1) * sizeof(o_pt));
  memcpy(o->corner_last, less_sharp, (size_t)n_less_sharp * sizeof(o_pt));
  memcpy(o->surf_last, less_flat, (size_t)n_less_flat * sizeof(o_pt));
  o->n_corner_last = n_less_sharp; o->n_surf_last = n_less_flat;
  lmono_cpu_kdtree_free(o->kd_corner); lmono_cpu_kdtree_free(o->kd_surf);
  o->kd_corner = NULL; o->kd_surf = NULL;
  if (o->use_kdtree) { o->kd_corner = lmono_cpu_kdtree_build(less_sharp, n_less_sharp); o->kd_surf = lmono_cpu_kdtree_build(less_flat, n_less_flat); }
  if (last_curr) { memcpy(last_curr->q, o->para_q, sizeof(o->para_q)); memcpy(last_curr->t, o->para_t, sizeof(o->para_t)); }
  if (w_curr) { memcpy(w_curr->q, o->q_w_curr, sizeof(o->q_w_curr)); memcpy(w_curr->t, o->t_w_curr, sizeof(o->t_w_curr)); }
  R.ms_whole = now_ms() - t_whole;
  if (rep) *rep = R;
  return 0;
}
