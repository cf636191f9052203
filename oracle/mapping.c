/*
 * ORACLE (test infrastructure): restatement of Aloam/src/laserMapping.cpp process()
 * (:231-893) -- rolling 21x21x11 cube map, 5x5x3 local-map gather, VoxelGrid of the
 * incoming features, 2 x (5-NN association + line/plane fit + ceres::Solve), map insertion
 * and per-cube VoxelGrid refilter.  Every block cites the lines it follows.
 */
#include "lmono_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

enum { W = 21, Hh = 21, D = 11, NCUBE = W * Hh * D };  /* laserMapping.cpp:77-82 */

typedef struct { o_pt* p; int n, cap; } cloud;

static void cloud_push(cloud* c, const o_pt* pt) {
  if (c->n == c->cap) { c->cap = c->cap ? c->cap * 2 : 64; c->p = (o_pt*)realloc(c->p, (size_t)c->cap * sizeof(o_pt)); }
  c->p[c->n++] = *pt;
}
static void cloud_append(cloud* dst, const cloud* src) { for (int i = 0; i < src->n; ++i) cloud_push(dst, &src->p[i]); }

struct o_mapper {
  float line_res, plane_res;
  int order_mode, use_kdtree;
  int cenW, cenH, cenD;                 /* laserCloudCenWidth/Height/Depth :74-76 */
  cloud* corner[NCUBE];                 /* laserCloudCornerArray :103 */
  cloud* surf[NCUBE];                   /* laserCloudSurfArray   :104 */
  double q_wmap_wodom[4], t_wmap_wodom[3];   /* :116-117 (x,y,z,w) */
  int valid_ind[125], valid_num;        /* :85 */
  int center[3];
  cloud corner_from_map, surf_from_map; /* :96-97 */
  int frame_count;
};

static double now_ms(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }

/* Eigen quaternion helpers, q = (x,y,z,w) */
static void qmul(const double a[4], const double b[4], double o[4]) {
  double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
static void qrot(const double q[4], const double v[3], double o[3]) {
  /* QuaternionBase::_transformVector */
  double uv[3] = { q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0] };
  uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  double c[3] = { q[1] * uv[2] - q[2] * uv[1], q[2] * uv[0] - q[0] * uv[2], q[0] * uv[1] - q[1] * uv[0] };
  for (int k = 0; k < 3; ++k) o[k] = (v[k] + q[3] * uv[k]) + c[k];
}
static void qinv(const double q[4], double o[4]) {
  /* QuaternionBase::inverse: conjugate / squaredNorm */
  double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (n2 > 0.0) { o[0] = -q[0] / n2; o[1] = -q[1] / n2; o[2] = -q[2] / n2; o[3] = q[3] / n2; }
  else { o[0] = o[1] = o[2] = o[3] = 0.0; }
}

o_mapper* lmono_cpu_mapper_create(float line_res, float plane_res, int voxel_order_mode, int use_kdtree) {
  o_mapper* m = (o_mapper*)calloc(1, sizeof(o_mapper));
  m->line_res = line_res; m->plane_res = plane_res;
  m->order_mode = voxel_order_mode; m->use_kdtree = use_kdtree;
  m->cenW = 10; m->cenH = 10; m->cenD = 5;
  for (int i = 0; i < NCUBE; ++i) { m->corner[i] = (cloud*)calloc(1, sizeof(cloud)); m->surf[i] = (cloud*)calloc(1, sizeof(cloud)); }
  m->q_wmap_wodom[3] = 1.0;
  return m;
}

void lmono_cpu_mapper_destroy(o_mapper* m) {
  if (!m) return;
  for (int i = 0; i < NCUBE; ++i) { free(m->corner[i]->p); free(m->corner[i]); free(m->surf[i]->p); free(m->surf[i]); }
  free(m->corner_from_map.p); free(m->surf_from_map.p);
  free(m);
}

void lmono_cpu_mapper_get_state(const o_mapper* m, o_pose* p, int32_t cen[3]) {
  if (p) { memcpy(p->q, m->q_wmap_wodom, sizeof(p->q)); memcpy(p->t, m->t_wmap_wodom, sizeof(p->t)); }
  if (cen) { cen[0] = m->cenW; cen[1] = m->cenH; cen[2] = m->cenD; }
}
void lmono_cpu_mapper_set_state(o_mapper* m, const o_pose* p) {
  memcpy(m->q_wmap_wodom, p->q, sizeof(p->q)); memcpy(m->t_wmap_wodom, p->t, sizeof(p->t));
}

/* cube index of a world point: laserMapping.cpp:741-750 (same arithmetic as :312-321) */
static int cube_coord(double v, int cen) {
  int c = (int)((v + 25.0) / 50.0) + cen;
  if (v + 25.0 < 0) c--;
  return c;
}

static void filter_cube(o_mapper* m, cloud* c, float leaf) {
  /* :792-800  downSizeFilter.setInputCloud(cube); filter(tmp); cube = tmp */
  if (c->n == 0) return;
  o_pt* out = (o_pt*)malloc((size_t)c->n * sizeof(o_pt));
  int no = 0;
  lmono_cpu_voxel_grid(c->p, c->n, leaf, m->order_mode, out, &no);
  free(c->p); c->p = out; c->n = no; c->cap = no > 0 ? no : 0;
  if (no == 0) { free(out); c->p = NULL; c->cap = 0; }
}

/* :323-507: six while loops rotating the cube pointers. axis 0:i 1:j 2:k; dir +1 means
 * "centerCube < 3" (contents move to higher index, slot 0 is recycled and cleared). */
static void shift_cubes(o_mapper* m, int axis, int dir) {
  const int dim[3] = { W, Hh, D };
  const int stride[3] = { 1, W, W * Hh };
  int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
  for (int u = 0; u < dim[a1]; ++u) for (int v = 0; v < dim[a2]; ++v) {
    int base = u * stride[a1] + v * stride[a2];
    if (dir > 0) {
      int i = dim[axis] - 1;
      cloud* cc = m->corner[base + i * stride[axis]]; cloud* cs = m->surf[base + i * stride[axis]];
      for (; i >= 1; i--) {
        m->corner[base + i * stride[axis]] = m->corner[base + (i - 1) * stride[axis]];
        m->surf[base + i * stride[axis]] = m->surf[base + (i - 1) * stride[axis]];
      }
      m->corner[base] = cc; m->surf[base] = cs;
      cc->n = 0; cs->n = 0;
    } else {
      int i = 0;
      cloud* cc = m->corner[base]; cloud* cs = m->surf[base];
      for (; i < dim[axis] - 1; i++) {
        m->corner[base + i * stride[axis]] = m->corner[base + (i + 1) * stride[axis]];
        m->surf[base + i * stride[axis]] = m->surf[base + (i + 1) * stride[axis]];
      }
      m->corner[base + i * stride[axis]] = cc; m->surf[base + i * stride[axis]] = cs;
      cc->n = 0; cs->n = 0;
    }
  }
}

/* :312-539: centre cube, shifts, valid list, concatenation */
static void prepare_window(o_mapper* m, const double t_w_curr[3]) {
  int cI = cube_coord(t_w_curr[0], m->cenW);
  int cJ = cube_coord(t_w_curr[1], m->cenH);
  int cK = cube_coord(t_w_curr[2], m->cenD);
  while (cI < 3) { shift_cubes(m, 0, +1); cI++; m->cenW++; }
  while (cI >= W - 3) { shift_cubes(m, 0, -1); cI--; m->cenW--; }
  while (cJ < 3) { shift_cubes(m, 1, +1); cJ++; m->cenH++; }
  while (cJ >= Hh - 3) { shift_cubes(m, 1, -1); cJ--; m->cenH--; }
  while (cK < 3) { shift_cubes(m, 2, +1); cK++; m->cenD++; }
  while (cK >= D - 3) { shift_cubes(m, 2, -1); cK--; m->cenD--; }
  m->center[0] = cI; m->center[1] = cJ; m->center[2] = cK;
  m->valid_num = 0;
  for (int i = cI - 2; i <= cI + 2; i++)
    for (int j = cJ - 2; j <= cJ + 2; j++)
      for (int k = cK - 1; k <= cK + 1; k++)
        if (i >= 0 && i < W && j >= 0 && j < Hh && k >= 0 && k < D)
          m->valid_ind[m->valid_num++] = i + W * j + W * Hh * k;
  m->corner_from_map.n = 0; m->surf_from_map.n = 0;
  for (int i = 0; i < m->valid_num; i++) {
    cloud_append(&m->corner_from_map, m->corner[m->valid_ind[i]]);
    cloud_append(&m->surf_from_map, m->surf[m->valid_ind[i]]);
  }
}

static void insert_point(o_mapper* m, cloud** arr, const o_pt* pw) {
  /* :741-758 */
  int cI = cube_coord((double)pw->x, m->cenW);
  int cJ = cube_coord((double)pw->y, m->cenH);
  int cK = cube_coord((double)pw->z, m->cenD);
  if (cI >= 0 && cI < W && cJ >= 0 && cJ < Hh && cK >= 0 && cK < D)
    cloud_push(arr[cI + W * cJ + W * Hh * cK], pw);
}

int lmono_cpu_mapper_import(o_mapper* m, int which, const o_pt* pts, int n) {
  cloud** arr = which == 0 ? m->corner : m->surf;
  for (int i = 0; i < n; ++i) insert_point(m, arr, &pts[i]);
  for (int c = 0; c < NCUBE; ++c) filter_cube(m, arr[c], which == 0 ? m->line_res : m->plane_res);
  return 0;
}

int lmono_cpu_mapper_export(o_mapper* m, int which, int scope, o_pt* out, int cap) {
  cloud** arr = which == 0 ? m->corner : m->surf;
  int n = 0;
  if (scope == 0) {
    for (int i = 0; i < m->valid_num; ++i) {
      cloud* c = arr[m->valid_ind[i]];
      for (int k = 0; k < c->n; ++k) { if (n < cap) out[n] = c->p[k]; ++n; }
    }
  } else {
    for (int i = 0; i < NCUBE; ++i) {
      cloud* c = arr[i];
      for (int k = 0; k < c->n; ++k) { if (n < cap) out[n] = c->p[k]; ++n; }
    }
  }
  return n;
}

/* pointAssociateToMap :154-163 */
static void associate_to_map(const double q[4], const double t[3], const o_pt* pi, o_pt* po) {
  double pc[3] = { pi->x, pi->y, pi->z }, pw[3];
  qrot(q, pc, pw);
  po->x = (float)(pw[0] + t[0]); po->y = (float)(pw[1] + t[1]); po->z = (float)(pw[2] + t[2]);
  po->i = pi->i;
}

static void knn5(const o_mapper* m, const o_kdtree* tree, const cloud* map, const o_pt* q, int32_t idx[5], float d2[5]) {
  if (m->use_kdtree && tree) lmono_cpu_kdtree_knn(tree, q, 1, 5, idx, d2);
  else lmono_cpu_knn_brute(map->p, map->n, q, 1, 5, idx, d2);
}

/* :577-622 one corner query. returns 1 if a factor was produced */
static int corner_factor(const o_mapper* m, const o_kdtree* tree, const double q[4], const double t[3], const o_pt* ori, o_factor* f) {
  o_pt sel; associate_to_map(q, t, ori, &sel);
  if (m->corner_from_map.n < 5) return 0;
  int32_t idx[5]; float d2[5];
  knn5(m, tree, &m->corner_from_map, &sel, idx, d2);
  if (!(d2[4] < 1.0)) return 0;
  double near[5][3], center[3] = { 0, 0, 0 };
  for (int j = 0; j < 5; ++j) {
    const o_pt* p = &m->corner_from_map.p[idx[j]];
    near[j][0] = p->x; near[j][1] = p->y; near[j][2] = p->z;
    for (int k = 0; k < 3; ++k) center[k] = center[k] + near[j][k];
  }
  for (int k = 0; k < 3; ++k) center[k] = center[k] / 5.0;
  double cov[9] = { 0 };
  for (int j = 0; j < 5; ++j) {
    double z[3] = { near[j][0] - center[0], near[j][1] - center[1], near[j][2] - center[2] };
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) cov[a * 3 + b] = cov[a * 3 + b] + z[a] * z[b];
  }
  double w[3], V[9];
  lmono_cpu_eigh3(cov, w, V);
  if (!(w[2] > 3 * w[1])) return 0;
  double dir[3] = { V[0 * 3 + 2], V[1 * 3 + 2], V[2 * 3 + 2] };
  f->type = O_FACTOR_EDGE; f->s = 1.0; f->pad = 0;
  f->p[0] = ori->x; f->p[1] = ori->y; f->p[2] = ori->z;
  for (int k = 0; k < 3; ++k) { f->a[k] = 0.1 * dir[k] + center[k]; f->b[k] = -0.1 * dir[k] + center[k]; }
  return 1;
}

/* :643-687 one surf query */
static int surf_factor(const o_mapper* m, const o_kdtree* tree, const double q[4], const double t[3], const o_pt* ori, o_factor* f) {
  o_pt sel; associate_to_map(q, t, ori, &sel);
  if (m->surf_from_map.n < 5) return 0;
  int32_t idx[5]; float d2[5];
  knn5(m, tree, &m->surf_from_map, &sel, idx, d2);
  if (!(d2[4] < 1.0)) return 0;
  double A[15], B[5] = { -1, -1, -1, -1, -1 };
  for (int j = 0; j < 5; ++j) {
    const o_pt* p = &m->surf_from_map.p[idx[j]];
    A[j * 3 + 0] = p->x; A[j * 3 + 1] = p->y; A[j * 3 + 2] = p->z;
  }
  double n[3];
  lmono_cpu_colpiv_qr_solve_5x3(A, B, n);
  double z2 = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
  double nn = sqrt(z2);
  double negative_OA_dot_norm = 1 / nn;
  /* norm.normalize(): z = squaredNorm(); if (z > 0) v /= sqrt(z) */
  if (z2 > 0.0) { n[0] /= nn; n[1] /= nn; n[2] /= nn; }
  for (int j = 0; j < 5; ++j) {
    if (fabs(n[0] * A[j * 3 + 0] + n[1] * A[j * 3 + 1] + n[2] * A[j * 3 + 2] + negative_OA_dot_norm) > 0.2) return 0;
  }
  f->type = O_FACTOR_PLANE_NORM; f->s = 1.0; f->pad = 0;
  f->p[0] = ori->x; f->p[1] = ori->y; f->p[2] = ori->z;
  f->a[0] = n[0]; f->a[1] = n[1]; f->a[2] = n[2];
  f->b[0] = negative_OA_dot_norm; f->b[1] = 0; f->b[2] = 0;
  return 1;
}

static int associate(const o_mapper* m, const o_kdtree* tc, const o_kdtree* ts,
                     const o_pt* cs, int nc, const o_pt* ss, int ns,
                     const double q[4], const double t[3], o_factor* out, int cap, int32_t* n_corner, int32_t* n_surf) {
  int nf = 0, c = 0, s = 0;
  for (int i = 0; i < nc; ++i) { o_factor f; if (corner_factor(m, tc, q, t, &cs[i], &f)) { if (nf < cap) out[nf] = f; ++nf; ++c; } }
  for (int i = 0; i < ns; ++i) { o_factor f; if (surf_factor(m, ts, q, t, &ss[i], &f)) { if (nf < cap) out[nf] = f; ++nf; ++s; } }
  if (n_corner) *n_corner = c;
  if (n_surf) *n_surf = s;
  return nf;
}

int lmono_cpu_mapper_prepare_window(o_mapper* m, const double t_w_curr[3]) {
  prepare_window(m, t_w_curr);
  return m->valid_num;
}

int lmono_cpu_mapper_knn5(o_mapper* m, int which, const o_pt* queries_world, int nq, int32_t* idx, float* d2) {
  const cloud* map = which == 0 ? &m->corner_from_map : &m->surf_from_map;
  if (m->use_kdtree) {
    o_kdtree* t = lmono_cpu_kdtree_build(map->p, map->n);
    lmono_cpu_kdtree_knn(t, queries_world, nq, 5, idx, d2);
    lmono_cpu_kdtree_free(t);
  } else {
    lmono_cpu_knn_brute(map->p, map->n, queries_world, nq, 5, idx, d2);
  }
  return map->n;
}

int lmono_cpu_mapper_associate(o_mapper* m, const o_pt* cs, int nc, const o_pt* ss, int ns,
                               const o_pose* w_curr, o_factor* out, int cap, int32_t* n_corner, int32_t* n_surf) {
  prepare_window(m, w_curr->t);
  o_kdtree *tc = NULL, *ts = NULL;
  if (m->use_kdtree) { tc = lmono_cpu_kdtree_build(m->corner_from_map.p, m->corner_from_map.n); ts = lmono_cpu_kdtree_build(m->surf_from_map.p, m->surf_from_map.n); }
  int nf = associate(m, tc, ts, cs, nc, ss, ns, w_curr->q, w_curr->t, out, cap, n_corner, n_surf);
  lmono_cpu_kdtree_free(tc); lmono_cpu_kdtree_free(ts);
  return nf;
}

int lmono_cpu_map_step(o_mapper* m, const o_pt* corner_last, int nc, const o_pt* surf_last, int ns,
                       const o_pose* wodom_curr, o_pose* w_curr_out, o_map_report* rep,
                       o_pt* full_res, int nfull) {
  o_map_report R; memset(&R, 0, sizeof(R));
  double t_whole = now_ms();
  const double* q_wodom_curr = wodom_curr->q; const double* t_wodom_curr = wodom_curr->t;
  /* transformAssociateToMap :142-146 */
  double q_w_curr[4], t_w_curr[3], tmp[3];
  qmul(m->q_wmap_wodom, q_wodom_curr, q_w_curr);
  qrot(m->q_wmap_wodom, t_wodom_curr, tmp);
  for (int k = 0; k < 3; ++k) t_w_curr[k] = tmp[k] + m->t_wmap_wodom[k];

  double t0 = now_ms();
  prepare_window(m, t_w_curr);
  R.center_cube[0] = m->center[0]; R.center_cube[1] = m->center[1]; R.center_cube[2] = m->center[2];
  R.corner_from_map = m->corner_from_map.n; R.surf_from_map = m->surf_from_map.n;

  /* :542-550 */
  o_pt* cstack = (o_pt*)malloc((size_t)(nc > 0 ? nc : 1) * sizeof(o_pt));
  o_pt* sstack = (o_pt*)malloc((size_t)(ns > 0 ? ns : 1) * sizeof(o_pt));
  int ncs = 0, nss = 0;
  lmono_cpu_voxel_grid(corner_last, nc, m->line_res, m->order_mode, cstack, &ncs);
  lmono_cpu_voxel_grid(surf_last, ns, m->plane_res, m->order_mode, sstack, &nss);
  R.corner_stack = ncs; R.surf_stack = nss;
  R.ms_shift = now_ms() - t0;

  if (m->corner_from_map.n > 10 && m->surf_from_map.n > 50) {   /* :554 */
    R.optimized = 1;
    double tt = now_ms();
    o_kdtree *tc = NULL, *ts = NULL;
    if (m->use_kdtree) {   /* :558-559 */
      tc = lmono_cpu_kdtree_build(m->corner_from_map.p, m->corner_from_map.n);
      ts = lmono_cpu_kdtree_build(m->surf_from_map.p, m->surf_from_map.n);
    }
    R.ms_tree = now_ms() - tt;
    int cap = ncs + nss; o_factor* fac = (o_factor*)malloc((size_t)(cap > 0 ? cap : 1) * sizeof(o_factor));
    for (int iter = 0; iter < 2; ++iter) {  /* :562 */
      double ta = now_ms();
      int nf = associate(m, tc, ts, cstack, ncs, sstack, nss, q_w_curr, t_w_curr, fac, cap, &R.corner_num[iter], &R.surf_num[iter]);
      R.ms_assoc += now_ms() - ta;
      double tsv = now_ms();
      o_pose x; memcpy(x.q, q_w_curr, sizeof(x.q)); memcpy(x.t, t_w_curr, sizeof(x.t));
      lmono_cpu_lm_solve(fac, nf, &x, 4, &R.solve[iter]);   /* :713-720 */
      memcpy(q_w_curr, x.q, sizeof(x.q)); memcpy(t_w_curr, x.t, sizeof(x.t));
      R.ms_solver += now_ms() - tsv;
    }
    free(fac);
    lmono_cpu_kdtree_free(tc); lmono_cpu_kdtree_free(ts);
  }
  /* transformUpdate :148-152 */
  double qi[4]; qinv(q_wodom_curr, qi);
  qmul(q_w_curr, qi, m->q_wmap_wodom);
  qrot(m->q_wmap_wodom, t_wodom_curr, tmp);
  for (int k = 0; k < 3; ++k) m->t_wmap_wodom[k] = t_w_curr[k] - tmp[k];

  /* :737-783 */
  double tadd = now_ms();
  for (int i = 0; i < ncs; ++i) { o_pt pw; associate_to_map(q_w_curr, t_w_curr, &cstack[i], &pw); insert_point(m, m->corner, &pw); }
  for (int i = 0; i < nss; ++i) { o_pt pw; associate_to_map(q_w_curr, t_w_curr, &sstack[i], &pw); insert_point(m, m->surf, &pw); }
  R.ms_add = now_ms() - tadd;
  /* :788-801 */
  double tf = now_ms();
  for (int i = 0; i < m->valid_num; ++i) {
    int ind = m->valid_ind[i];
    filter_cube(m, m->corner[ind], m->line_res);
    filter_cube(m, m->surf[ind], m->plane_res);
  }
  R.ms_filter = now_ms() - tf;
  /* :838-842 */
  for (int i = 0; i < nfull; ++i) { o_pt pw; associate_to_map(q_w_curr, t_w_curr, &full_res[i], &pw); full_res[i] = pw; }
  free(cstack); free(sstack);
  m->frame_count++;
  R.cen[0] = m->cenW; R.cen[1] = m->cenH; R.cen[2] = m->cenD;
  R.ms_whole = now_ms() - t_whole;
  if (w_curr_out) { memcpy(w_curr_out->q, q_w_curr, sizeof(q_w_curr)); memcpy(w_curr_out->t, t_w_curr, sizeof(t_w_curr)); }
  if (rep) *rep = R;
  return 0;
}
