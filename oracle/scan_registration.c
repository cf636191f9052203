/*
 * ORACLE (test infrastructure): restatement of laserCloudHandler in
 * Aloam/src/scanRegistration.cpp:114-459 (NaN / near-point removal, ring id and relative
 * time per point, ring-major reorder, 11-tap curvature, 6-sector sort per ring, sharp /
 * less-sharp / flat / less-flat picking with neighbour suppression, per-ring 0.2 m
 * VoxelGrid of the less-flat set).
 *
 * Precision notes (SURVEY.md 8a cheat-sheet):
 *  - `atan(z / sqrt(x*x+y*y))` (:166) resolves to the double libm functions on the author's
 *    toolchain (GCC 5.4: <cmath> leaves only ::atan(double)/::sqrt(double) in the global
 *    namespace).  trig_mode 0 reproduces that (default, also what the CUDA path computes);
 *    trig_mode 1 is the float-overload variant newer libstdc++ headers select.
 *  - `atan2` (:141-143, :208) is std::atan2(float, float) -> atan2f in both cases.
 *  - curvature, gap tests and range gate are fp32 without FMA, literals 0.1 / 0.05 are
 *    double (so a float curvature is widened before the compare).
 */
#include "lmono_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef struct { o_pt* p; int32_t* src; int n, cap; } ring_t;

static void ring_push(ring_t* r, const o_pt* pt, int src) {
  if (r->n == r->cap) {
    r->cap = r->cap ? r->cap * 2 : 256;
    r->p = (o_pt*)realloc(r->p, (size_t)r->cap * sizeof(o_pt));
    r->src = (int32_t*)realloc(r->src, (size_t)r->cap * sizeof(int32_t));
  }
  r->p[r->n] = *pt; r->src[r->n] = src; r->n++;
}

static const float* g_curv;
static int cmp_curv_stable(const void* a, const void* b) {
  int ia = *(const int32_t*)a, ib = *(const int32_t*)b;
  float ca = g_curv[ia], cb = g_curv[ib];
  if (ca < cb) return -1;
  if (ca > cb) return 1;
  return ia < ib ? -1 : (ia > ib ? 1 : 0);
}

static int g_trig_mode = 0;
void lmono_cpu_scan_set_trig_mode(int mode) { g_trig_mode = mode; }

int lmono_cpu_scan_register(const float* xyz_in, int n_in, int stride_floats,
                            int n_scans, float minimum_range, int voxel_order_mode, int sort_mode,
                            o_pt* full, o_pt* sharp, o_pt* less_sharp, o_pt* flat, o_pt* less_flat,
                            int32_t* labels, float* curvature, int32_t* src_index,
                            o_scan_report* rep) {
  o_scan_report R; memset(&R, 0, sizeof(R));
  R.n_in = n_in;
  if (n_scans != 16 && n_scans != 32 && n_scans != 64) return -1;
  const double scanPeriod = 0.1;   /* :60 */

  /* :136-137 removeNaNFromPointCloud + removeClosedPointCloud (order preserving) */
  float* in = (float*)malloc((size_t)(n_in > 0 ? n_in : 1) * 3 * sizeof(float));
  int32_t* in_src = (int32_t*)malloc((size_t)(n_in > 0 ? n_in : 1) * sizeof(int32_t));
  int cloudSize = 0;
  const float thres = minimum_range;
  for (int i = 0; i < n_in; ++i) {
    const float x = xyz_in[(size_t)i * stride_floats], y = xyz_in[(size_t)i * stride_floats + 1], z = xyz_in[(size_t)i * stride_floats + 2];
    if (!isfinite(x) || !isfinite(y) || !isfinite(z)) continue;
    if (x * x + y * y + z * z < thres * thres) continue;   /* :99, all fp32 */
    in[(size_t)cloudSize * 3] = x; in[(size_t)cloudSize * 3 + 1] = y; in[(size_t)cloudSize * 3 + 2] = z;
    in_src[cloudSize] = i;
    cloudSize++;
  }
  if (cloudSize == 0) { free(in); free(in_src); if (rep) *rep = R; return 0; }

  /* :141-153 */
  float startOri = -atan2f(in[1], in[0]);
  float endOri = (float)(-atan2f(in[(size_t)(cloudSize - 1) * 3 + 1], in[(size_t)(cloudSize - 1) * 3]) + 2 * M_PI);
  if (endOri - startOri > 3 * M_PI) endOri = (float)(endOri - 2 * M_PI);
  else if (endOri - startOri < M_PI) endOri = (float)(endOri + 2 * M_PI);
  R.start_ori = startOri; R.end_ori = endOri;

  int halfPassed = 0;
  int count = cloudSize;
  ring_t* rings = (ring_t*)calloc((size_t)n_scans, sizeof(ring_t));
  for (int i = 0; i < cloudSize; i++) {
    o_pt point;
    point.x = in[(size_t)i * 3]; point.y = in[(size_t)i * 3 + 1]; point.z = in[(size_t)i * 3 + 2];
    float angle;
    if (g_trig_mode == 0) angle = (float)(atan((double)point.z / sqrt((double)(point.x * point.x + point.y * point.y))) * 180 / M_PI);
    else angle = (float)(atanf(point.z / sqrtf(point.x * point.x + point.y * point.y)) * 180 / M_PI);
    int scanID = 0;
    if (n_scans == 16) {
      scanID = (int)((angle + 15) / 2 + 0.5);
      if (scanID > (n_scans - 1) || scanID < 0) { count--; continue; }
    } else if (n_scans == 32) {
      scanID = (int)((angle + 92.0 / 3.0) * 3.0 / 4.0);
      if (scanID > (n_scans - 1) || scanID < 0) { count--; continue; }
    } else {
      if (angle >= -8.83) scanID = (int)((2 - angle) * 3.0 + 0.5);
      else scanID = n_scans / 2 + (int)((-8.83 - angle) * 2.0 + 0.5);
      if (angle > 2 || angle < -24.33 || scanID > 50 || scanID < 0) { count--; continue; }   /* :195 */
    }
    float ori = -atan2f(point.y, point.x);
    if (!halfPassed) {
      if (ori < startOri - M_PI / 2) ori = (float)(ori + 2 * M_PI);
      else if (ori > startOri + M_PI * 3 / 2) ori = (float)(ori - 2 * M_PI);
      if (ori - startOri > M_PI) halfPassed = 1;
    } else {
      ori = (float)(ori + 2 * M_PI);
      if (ori < endOri - M_PI * 3 / 2) ori = (float)(ori + 2 * M_PI);
      else if (ori > endOri + M_PI / 2) ori = (float)(ori - 2 * M_PI);
    }
    float relTime = (ori - startOri) / (endOri - startOri);
    point.i = (float)(scanID + scanPeriod * relTime);
    ring_push(&rings[scanID], &point, in_src[i]);
  }
  cloudSize = count;
  R.n_kept = cloudSize;

  /* :246-252 */
  int scanStartInd[64], scanEndInd[64];
  int n = 0;
  for (int i = 0; i < n_scans; i++) {
    scanStartInd[i] = n + 5;
    for (int k = 0; k < rings[i].n; ++k) { full[n] = rings[i].p[k]; if (src_index) src_index[n] = rings[i].src[k]; n++; }
    scanEndInd[i] = n - 6;
  }
  for (int i = 0; i < 64; ++i) { R.ring_start[i] = i < n_scans ? scanStartInd[i] : 0; R.ring_end[i] = i < n_scans ? scanEndInd[i] : 0; }

  /* the reference's global scratch arrays (:66-69) are sized 400000 and neighbour writes may
   * reach +-5 beyond [5, cloudSize-5); give the scratch the same slack */
  const int N = cloudSize;
  float* cloudCurvature = (float*)calloc((size_t)N + 16, sizeof(float));
  int32_t* cloudSortInd = (int32_t*)calloc((size_t)N + 16, sizeof(int32_t));
  int32_t* cloudNeighborPicked = (int32_t*)calloc((size_t)N + 16, sizeof(int32_t));
  int32_t* cloudLabel = (int32_t*)calloc((size_t)N + 16, sizeof(int32_t));

  /* :256-266, strictly left to right in fp32 */
  for (int i = 5; i < cloudSize - 5; i++) {
    float diffX = full[i - 5].x + full[i - 4].x + full[i - 3].x + full[i - 2].x + full[i - 1].x - 10 * full[i].x + full[i + 1].x + full[i + 2].x + full[i + 3].x + full[i + 4].x + full[i + 5].x;
    float diffY = full[i - 5].y + full[i - 4].y + full[i - 3].y + full[i - 2].y + full[i - 1].y - 10 * full[i].y + full[i + 1].y + full[i + 2].y + full[i + 3].y + full[i + 4].y + full[i + 5].y;
    float diffZ = full[i - 5].z + full[i - 4].z + full[i - 3].z + full[i - 2].z + full[i - 1].z - 10 * full[i].z + full[i + 1].z + full[i + 2].z + full[i + 3].z + full[i + 4].z + full[i + 5].z;
    cloudCurvature[i] = diffX * diffX + diffY * diffY + diffZ * diffZ;
    cloudSortInd[i] = i;
    cloudNeighborPicked[i] = 0;
    cloudLabel[i] = 0;
  }

  int n_sharp = 0, n_less_sharp = 0, n_flat = 0, n_less_flat = 0;
  o_pt* lfs = (o_pt*)malloc((size_t)(N > 0 ? N : 1) * sizeof(o_pt));
  o_pt* lfs_ds = (o_pt*)malloc((size_t)(N > 0 ? N : 1) * sizeof(o_pt));
  for (int i = 0; i < n_scans; i++) {
    if (scanEndInd[i] - scanStartInd[i] < 6) continue;
    int n_lfs = 0;
    for (int j = 0; j < 6; j++) {
      int sp = scanStartInd[i] + (scanEndInd[i] - scanStartInd[i]) * j / 6;
      int ep = scanStartInd[i] + (scanEndInd[i] - scanStartInd[i]) * (j + 1) / 6 - 1;
      if (sort_mode == 1) lmono_cpu_stdsort_by_curvature(cloudSortInd + sp, ep + 1 - sp, cloudCurvature);
      else { g_curv = cloudCurvature; qsort(cloudSortInd + sp, (size_t)(ep + 1 - sp), sizeof(int32_t), cmp_curv_stable); }

      int largestPickedNum = 0;
      for (int k = ep; k >= sp; k--) {
        int ind = cloudSortInd[k];
        if (cloudNeighborPicked[ind] == 0 && cloudCurvature[ind] > 0.1) {
          largestPickedNum++;
          if (largestPickedNum <= 2) {
            cloudLabel[ind] = 2;
            sharp[n_sharp++] = full[ind];
            less_sharp[n_less_sharp++] = full[ind];
          } else if (largestPickedNum <= 20) {
            cloudLabel[ind] = 1;
            less_sharp[n_less_sharp++] = full[ind];
          } else {
            break;
          }
          cloudNeighborPicked[ind] = 1;
          for (int l = 1; l <= 5; l++) {
            float diffX = full[ind + l].x - full[ind + l - 1].x;
            float diffY = full[ind + l].y - full[ind + l - 1].y;
            float diffZ = full[ind + l].z - full[ind + l - 1].z;
            if (diffX * diffX + diffY * diffY + diffZ * diffZ > 0.05) break;
            cloudNeighborPicked[ind + l] = 1;
          }
          for (int l = -1; l >= -5; l--) {
            float diffX = full[ind + l].x - full[ind + l + 1].x;
            float diffY = full[ind + l].y - full[ind + l + 1].y;
            float diffZ = full[ind + l].z - full[ind + l + 1].z;
            if (diffX * diffX + diffY * diffY + diffZ * diffZ > 0.05) break;
            cloudNeighborPicked[ind + l] = 1;
          }
        }
      }
      int smallestPickedNum = 0;
      for (int k = sp; k <= ep; k++) {
        int ind = cloudSortInd[k];
        if (cloudNeighborPicked[ind] == 0 && cloudCurvature[ind] < 0.1) {
          cloudLabel[ind] = -1;
          flat[n_flat++] = full[ind];
          smallestPickedNum++;
          if (smallestPickedNum >= 4) break;     /* :359-362: the 4th is pushed but not marked */
          cloudNeighborPicked[ind] = 1;
          for (int l = 1; l <= 5; l++) {
            float diffX = full[ind + l].x - full[ind + l - 1].x;
            float diffY = full[ind + l].y - full[ind + l - 1].y;
            float diffZ = full[ind + l].z - full[ind + l - 1].z;
            if (diffX * diffX + diffY * diffY + diffZ * diffZ > 0.05) break;
            cloudNeighborPicked[ind + l] = 1;
          }
          for (int l = -1; l >= -5; l--) {
            float diffX = full[ind + l].x - full[ind + l + 1].x;
            float diffY = full[ind + l].y - full[ind + l + 1].y;
            float diffZ = full[ind + l].z - full[ind + l + 1].z;
            if (diffX * diffX + diffY * diffY + diffZ * diffZ > 0.05) break;
            cloudNeighborPicked[ind + l] = 1;
          }
        }
      }
      for (int k = sp; k <= ep; k++) if (cloudLabel[k] <= 0) lfs[n_lfs++] = full[k];   /* :392-398 */
    }
    int n_ds = 0;
    lmono_cpu_voxel_grid(lfs, n_lfs, 0.2f, voxel_order_mode, lfs_ds, &n_ds);   /* :401-405 */
    memcpy(less_flat + n_less_flat, lfs_ds, (size_t)n_ds * sizeof(o_pt));
    n_less_flat += n_ds;
  }
  R.n_sharp = n_sharp; R.n_less_sharp = n_less_sharp; R.n_flat = n_flat; R.n_less_flat = n_less_flat;
  if (labels) for (int i = 0; i < N; ++i) labels[i] = cloudLabel[i];
  if (curvature) for (int i = 0; i < N; ++i) curvature[i] = cloudCurvature[i];

  for (int i = 0; i < n_scans; ++i) { free(rings[i].p); free(rings[i].src); }
  free(rings); free(in); free(in_src); free(lfs); free(lfs_ds);
  free(cloudCurvature); free(cloudSortInd); free(cloudNeighborPicked); free(cloudLabel);
  if (rep) *rep = R;
  return 0;
}
