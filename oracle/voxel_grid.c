/*
 * ORACLE (test infrastructure): restatement of pcl::VoxelGrid<PointXYZI>::applyFilter
 * as configured by the reference (leaf set by setLeafSize, downsample_all_data=true,
 * min_points_per_voxel=0, no field filter).  PCL is an un-vendored dependency of the
 * reference (Aloam/docker/Dockerfile:4 pins PCL 1.8.0); the algorithm below follows
 * pcl/filters/impl/voxel_grid.hpp of that release.
 * Reference call sites: Aloam/src/scanRegistration.cpp:401-405,
 * Aloam/src/laserMapping.cpp:129-130,542-550,788-801,905-906.
 */
#include "lmono_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

typedef struct { uint32_t idx; uint32_t pt; } vg_pair;

static int cmp_pair_stable(const void* a, const void* b) {
  const vg_pair* pa = (const vg_pair*)a; const vg_pair* pb = (const vg_pair*)b;
  if (pa->idx != pb->idx) return pa->idx < pb->idx ? -1 : 1;
  if (pa->pt != pb->pt) return pa->pt < pb->pt ? -1 : 1;
  return 0;
}

int lmono_cpu_voxel_grid(const o_pt* in, int n, float leaf, int order_mode,
                         o_pt* out, int* n_out) {
  *n_out = 0;
  if (n <= 0) return 0;
  /* inverse_leaf_size_ = Array4f::Ones() / leaf_size_.array()  (fp32 division) */
  const float inv = 1.0f / leaf;

  /* getMinMax3D */
  float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, mx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
  for (int i = 0; i < n; ++i) {
    const float p[3] = { in[i].x, in[i].y, in[i].z };
    for (int d = 0; d < 3; ++d) { if (p[d] < mn[d]) mn[d] = p[d]; if (p[d] > mx[d]) mx[d] = p[d]; }
  }
  /* "Check that the leaf size is not too small, given the size of the data" */
  int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1;
  int64_t dy = (int64_t)((mx[1] - mn[1]) * inv) + 1;
  int64_t dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > (int64_t)INT32_MAX) {
    /* PCL warns and returns the input unchanged */
    memcpy(out, in, (size_t)n * sizeof(o_pt));
    *n_out = n;
    return 1;
  }
  int min_b[3], max_b[3], div_b[3], mul[3];
  for (int d = 0; d < 3; ++d) {
    min_b[d] = (int)floorf(mn[d] * inv);
    max_b[d] = (int)floorf(mx[d] * inv);
    div_b[d] = max_b[d] - min_b[d] + 1;
  }
  mul[0] = 1; mul[1] = div_b[0]; mul[2] = div_b[0] * div_b[1];

  vg_pair* iv = (vg_pair*)malloc((size_t)n * sizeof(vg_pair));
  for (int i = 0; i < n; ++i) {
    int ijk0 = (int)(floorf(in[i].x * inv) - (float)min_b[0]);
    int ijk1 = (int)(floorf(in[i].y * inv) - (float)min_b[1]);
    int ijk2 = (int)(floorf(in[i].z * inv) - (float)min_b[2]);
    int idx = ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2];
    iv[i].idx = (uint32_t)idx; iv[i].pt = (uint32_t)i;
  }
  if (order_mode == 1) {
    uint32_t* a = (uint32_t*)malloc((size_t)n * sizeof(uint32_t));
    uint32_t* b = (uint32_t*)malloc((size_t)n * sizeof(uint32_t));
    for (int i = 0; i < n; ++i) { a[i] = iv[i].idx; b[i] = iv[i].pt; }
    lmono_cpu_stdsort_voxel_pairs(a, b, n);
    for (int i = 0; i < n; ++i) { iv[i].idx = a[i]; iv[i].pt = b[i]; }
    free(a); free(b);
  } else {
    qsort(iv, (size_t)n, sizeof(vg_pair), cmp_pair_stable);
  }
  /* runs of equal idx -> CentroidPoint<PointXYZI>: fp32 sums of x,y,z,intensity in
   * sorted order, divided by the count converted to float. */
  int m = 0;
  int index = 0;
  while (index < n) {
    int i = index + 1;
    while (i < n && iv[i].idx == iv[index].idx) ++i;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    for (int li = index; li < i; ++li) {
      const o_pt* p = &in[iv[li].pt];
      sx += p->x; sy += p->y; sz += p->z; si += p->i;
    }
    const float cnt = (float)(i - index);
    out[m].x = sx / cnt; out[m].y = sy / cnt; out[m].z = sz / cnt; out[m].i = si / cnt;
    ++m;
    index = i;
  }
  free(iv);
  *n_out = m;
  return 0;
}
