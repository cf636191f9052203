/*
 * ORACLE (test infrastructure): nearest-neighbour search as the reference performs it
 * through pcl::KdTreeFLANN<PointXYZI> (un-vendored PCL 1.8 + FLANN 1.8/1.9;
 * SURVEY.md App. B.1-B.2).  Reference call sites: Aloam/src/laserOdometry.cpp:77-78,
 * 302,390,567-568 and Aloam/src/laserMapping.cpp:107-108,558-559,582,648.
 *
 *  - lmono_cpu_knn_brute: the ground truth.  Exact k-NN on (x,y,z), distance
 *    d2 = ((dx*dx)+dy*dy)+dz*dz accumulated in fp32 (FLANN L2_Simple<float>), results
 *    ascending by (d2, index).  Equal to FLANN whenever the k+1 nearest have distinct d2.
 *  - o_kdtree: restatement of FLANN KDTreeSingleIndex (leaf_max_size 15, reorder=true,
 *    middleSplit_, searchLevel with eps=0, KNNSimpleResultSet tie behaviour) used as the
 *    realistic CPU baseline (it is what the reference pays for each sweep) and to study
 *    tie order.
 */
#include "lmono_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <math.h>

static inline float l2_simple(const float* a, const float* b) {
  /* flann::L2_Simple: result=0; for i: diff=a[i]-b[i]; result += diff*diff  (fp32) */
  float r = 0.f;
  for (int i = 0; i < 3; ++i) { float diff = a[i] - b[i]; r += diff * diff; }
  return r;
}

int lmono_cpu_knn_brute(const o_pt* pts, int n, const o_pt* queries, int nq,
                        int k, int32_t* idx, float* d2) {
  for (int q = 0; q < nq; ++q) {
    int32_t* bi = idx + (size_t)q * k; float* bd = d2 + (size_t)q * k;
    int cnt = 0;
    for (int j = 0; j < k; ++j) { bi[j] = -1; bd[j] = FLT_MAX; }
    const float qv[3] = { queries[q].x, queries[q].y, queries[q].z };
    for (int i = 0; i < n; ++i) {
      const float pv[3] = { pts[i].x, pts[i].y, pts[i].z };
      float d = l2_simple(qv, pv);
      if (cnt == k && !(d < bd[k - 1])) continue;   /* index order => ties keep the lower index */
      int pos = cnt < k ? cnt : k - 1;
      while (pos > 0 && bd[pos - 1] > d) { bd[pos] = bd[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
      bd[pos] = d; bi[pos] = i;
      if (cnt < k) ++cnt;
    }
  }
  return 0;
}

/* ---------------- FLANN KDTreeSingleIndex restatement ---------------------- */
typedef struct { float low, high; } interval;
typedef struct kd_node {
  int left, right;          /* leaf: index range into vind */
  int divfeat;
  float divlow, divhigh;
  struct kd_node *child1, *child2;
} kd_node;

struct o_kdtree {
  int n;
  float* pts;      /* n x 3, original order */
  float* data;     /* n x 3, tree order (reorder=true) */
  int* vind;
  kd_node* root;
  interval root_bbox[3];
  kd_node* pool; int pool_used, pool_cap;
};

static kd_node* new_node(o_kdtree* t) {
  if (t->pool_used == t->pool_cap) return NULL;
  kd_node* nd = &t->pool[t->pool_used++];
  memset(nd, 0, sizeof(*nd));
  return nd;
}

static void compute_minmax(const o_kdtree* t, const int* ind, int count, int dim, float* mn, float* mx) {
  *mn = t->pts[(size_t)ind[0] * 3 + dim]; *mx = *mn;
  for (int i = 1; i < count; ++i) {
    float v = t->pts[(size_t)ind[i] * 3 + dim];
    if (v < *mn) *mn = v;
    if (v > *mx) *mx = v;
  }
}

static void plane_split(const o_kdtree* t, int* ind, int count, int cutfeat, float cutval, int* lim1, int* lim2) {
  int left = 0, right = count - 1;
  for (;;) {
    while (left <= right && t->pts[(size_t)ind[left] * 3 + cutfeat] < cutval) ++left;
    while (left <= right && t->pts[(size_t)ind[right] * 3 + cutfeat] >= cutval) --right;
    if (left > right) break;
    int tmp = ind[left]; ind[left] = ind[right]; ind[right] = tmp; ++left; --right;
  }
  *lim1 = left;
  right = count - 1;
  for (;;) {
    while (left <= right && t->pts[(size_t)ind[left] * 3 + cutfeat] <= cutval) ++left;
    while (left <= right && t->pts[(size_t)ind[right] * 3 + cutfeat] > cutval) --right;
    if (left > right) break;
    int tmp = ind[left]; ind[left] = ind[right]; ind[right] = tmp; ++left; --right;
  }
  *lim2 = left;
}

static void middle_split(const o_kdtree* t, int* ind, int count, int* index, int* cutfeat, float* cutval, const interval* bbox) {
  const float EPS = 0.00001f;
  float max_span = bbox[0].high - bbox[0].low;
  for (int i = 1; i < 3; ++i) { float span = bbox[i].high - bbox[i].low; if (span > max_span) max_span = span; }
  float max_spread = -1;
  *cutfeat = 0;
  for (int i = 0; i < 3; ++i) {
    float span = bbox[i].high - bbox[i].low;
    if (span > (float)((1 - EPS) * max_span)) {
      float mn, mx;
      compute_minmax(t, ind, count, i, &mn, &mx);
      float spread = mx - mn;
      if (spread > max_spread) { *cutfeat = i; max_spread = spread; }
    }
  }
  float split_val = (bbox[*cutfeat].low + bbox[*cutfeat].high) / 2;
  float mn, mx;
  compute_minmax(t, ind, count, *cutfeat, &mn, &mx);
  if (split_val < mn) *cutval = mn;
  else if (split_val > mx) *cutval = mx;
  else *cutval = split_val;
  int lim1, lim2;
  plane_split(t, ind, count, *cutfeat, *cutval, &lim1, &lim2);
  if (lim1 > count / 2) *index = lim1;
  else if (lim2 < count / 2) *index = lim2;
  else *index = count / 2;
}

static kd_node* divide_tree(o_kdtree* t, int left, int right, interval* bbox) {
  kd_node* node = new_node(t);
  if ((right - left) <= 15) {             /* KDTreeSingleIndexParams(15) via PCL */
    node->child1 = node->child2 = NULL;
    node->left = left; node->right = right;
    for (int i = 0; i < 3; ++i) { bbox[i].low = bbox[i].high = t->pts[(size_t)t->vind[left] * 3 + i]; }
    for (int k = left + 1; k < right; ++k)
      for (int i = 0; i < 3; ++i) {
        float v = t->pts[(size_t)t->vind[k] * 3 + i];
        if (bbox[i].low > v) bbox[i].low = v;
        if (bbox[i].high < v) bbox[i].high = v;
      }
  } else {
    int idx, cutfeat; float cutval;
    middle_split(t, t->vind + left, right - left, &idx, &cutfeat, &cutval, bbox);
    node->divfeat = cutfeat;
    interval lb[3], rb[3];
    memcpy(lb, bbox, sizeof(lb)); lb[cutfeat].high = cutval;
    node->child1 = divide_tree(t, left, left + idx, lb);
    memcpy(rb, bbox, sizeof(rb)); rb[cutfeat].low = cutval;
    node->child2 = divide_tree(t, left + idx, right, rb);
    node->divlow = lb[cutfeat].high;
    node->divhigh = rb[cutfeat].low;
    for (int i = 0; i < 3; ++i) {
      bbox[i].low = lb[i].low < rb[i].low ? lb[i].low : rb[i].low;
      bbox[i].high = lb[i].high > rb[i].high ? lb[i].high : rb[i].high;
    }
  }
  return node;
}

o_kdtree* lmono_cpu_kdtree_build(const o_pt* pts, int n) {
  o_kdtree* t = (o_kdtree*)calloc(1, sizeof(o_kdtree));
  t->n = n;
  if (n <= 0) return t;
  /* KdTreeFLANN::convertCloudToArray: dense float[n x 3] of x,y,z */
  t->pts = (float*)malloc((size_t)n * 3 * sizeof(float));
  for (int i = 0; i < n; ++i) { t->pts[(size_t)i * 3] = pts[i].x; t->pts[(size_t)i * 3 + 1] = pts[i].y; t->pts[(size_t)i * 3 + 2] = pts[i].z; }
  t->vind = (int*)malloc((size_t)n * sizeof(int));
  for (int i = 0; i < n; ++i) t->vind[i] = i;
  for (int d = 0; d < 3; ++d) { t->root_bbox[d].low = t->root_bbox[d].high = t->pts[d]; }
  for (int k = 1; k < n; ++k)
    for (int d = 0; d < 3; ++d) {
      float v = t->pts[(size_t)k * 3 + d];
      if (v < t->root_bbox[d].low) t->root_bbox[d].low = v;
      if (v > t->root_bbox[d].high) t->root_bbox[d].high = v;
    }
  t->pool_cap = 2 * n + 16; t->pool = (kd_node*)malloc((size_t)t->pool_cap * sizeof(kd_node));
  interval bbox[3]; memcpy(bbox, t->root_bbox, sizeof(bbox));
  t->root = divide_tree(t, 0, n, bbox);
  /* FLANN recomputes root_bbox_ through divideTree's bbox argument */
  memcpy(t->root_bbox, bbox, sizeof(bbox));
  t->data = (float*)malloc((size_t)n * 3 * sizeof(float));
  for (int i = 0; i < n; ++i) memcpy(t->data + (size_t)i * 3, t->pts + (size_t)t->vind[i] * 3, 3 * sizeof(float));
  return t;
}

void lmono_cpu_kdtree_free(o_kdtree* t) {
  if (!t) return;
  free(t->pts); free(t->data); free(t->vind); free(t->pool); free(t);
}

typedef struct { int k, count; float worst; float* d; int32_t* i; } result_set;

static inline void rs_add(result_set* r, float dist, int index) {
  /* flann::KNNSimpleResultSet::addPoint */
  if (dist >= r->worst) return;
  if (r->count < r->k) ++r->count;
  int i;
  for (i = r->count - 1; i > 0; --i) {
    if (r->d[i - 1] > dist) { r->d[i] = r->d[i - 1]; r->i[i] = r->i[i - 1]; }
    else break;
  }
  r->d[i] = dist; r->i[i] = index;
  r->worst = r->d[r->k - 1];
}

static void search_level(const o_kdtree* t, result_set* rs, const float* vec, const kd_node* node,
                         float mindistsq, float* dists) {
  if (node->child1 == NULL && node->child2 == NULL) {
    float worst_dist = rs->worst;
    for (int i = node->left; i < node->right; ++i) {
      float dist = l2_simple(vec, t->data + (size_t)i * 3);
      if (dist < worst_dist) rs_add(rs, dist, t->vind[i]);
    }
    return;
  }
  int idx = node->divfeat;
  float val = vec[idx];
  float diff1 = val - node->divlow;
  float diff2 = val - node->divhigh;
  const kd_node *best, *other; float cut_dist;
  if ((diff1 + diff2) < 0) { best = node->child1; other = node->child2; cut_dist = (val - node->divhigh) * (val - node->divhigh); }
  else { best = node->child2; other = node->child1; cut_dist = (val - node->divlow) * (val - node->divlow); }
  search_level(t, rs, vec, best, mindistsq, dists);
  float dst = dists[idx];
  mindistsq = mindistsq + cut_dist - dst;
  dists[idx] = cut_dist;
  if (mindistsq * 1.0f <= rs->worst) search_level(t, rs, vec, other, mindistsq, dists);
  dists[idx] = dst;
}

int lmono_cpu_kdtree_knn(const o_kdtree* t, const o_pt* queries, int nq, int k, int32_t* idx, float* d2) {
  for (int q = 0; q < nq; ++q) {
    result_set rs; rs.k = k; rs.count = 0; rs.worst = FLT_MAX;
    rs.d = d2 + (size_t)q * k; rs.i = idx + (size_t)q * k;
    for (int j = 0; j < k; ++j) { rs.d[j] = FLT_MAX; rs.i[j] = -1; }
    if (t->n <= 0) continue;
    const float vec[3] = { queries[q].x, queries[q].y, queries[q].z };
    float dists[3] = { 0, 0, 0 };
    float distsq = 0.f;
    for (int i = 0; i < 3; ++i) {
      if (vec[i] < t->root_bbox[i].low) { dists[i] = (vec[i] - t->root_bbox[i].low) * (vec[i] - t->root_bbox[i].low); distsq += dists[i]; }
      if (vec[i] > t->root_bbox[i].high) { dists[i] = (vec[i] - t->root_bbox[i].high) * (vec[i] - t->root_bbox[i].high); distsq += dists[i]; }
    }
    search_level(t, &rs, vec, t->root, distsq, dists);
  }
  return 0;
}
