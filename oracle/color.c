/*
 * ORACLE (test infrastructure): restatement of the LiDAR -> camera colour projection of
 * mono_lidar_mapping:
 *   D1  map_build_node.cc:216-225   T = [rlc^T | -rlc^T tlc], pcl::transformPointCloud (double
 *                                   matrix, float store)
 *   D2  Map_Builder.cc:224-245      skip z < 0, PinholeCamera::spaceToPlane
 *                                   (camera_models PinholeCamera.cc:520-542, :646-662), accept
 *                                   0 < u < cols, 0 < v < rows, depth_map(int(v), int(u)) = 100 - z,
 *                                   last point in cloud order wins.  (The r = 3 HSV discs drawn
 *                                   into the debug image pro_map are visualisation only and not
 *                                   restated.)
 *   D3  Map_Builder.cc:336-403      depthFill: dilate(K), close(rect K), dilate(7x7), zero fill,
 *                                   medianBlur 5, bilateralFilter(5, 1.5, 2.0) | GaussianBlur 5x5.
 *                                   OpenCV is un-vendored; the restatement follows the scalar
 *                                   (non-SIMD) code paths of OpenCV 3.x imgproc (morph.cpp,
 *                                   smooth.cpp) and is cross-checked against cv2 in tests/.
 *   D4  Map_Builder.cc:275-322      per pixel lift (PinholeCamera.cc:450-510), 0 < d < 70 gate,
 *                                   colour fetch, |x| > 20 && y > 1.8 rejection, world transform.
 */
#include "lmono_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* PinholeCamera::distortion :646-662 */
static void distortion(const o_camera* c, double x, double y, double* dx, double* dy) {
  double mx2_u = x * x, my2_u = y * y, mxy_u = x * y;
  double rho2_u = mx2_u + my2_u;
  double rad_dist_u = c->k1 * rho2_u + c->k2 * rho2_u * rho2_u;
  *dx = x * rad_dist_u + 2.0 * c->p1 * mxy_u + c->p2 * (rho2_u + 2.0 * mx2_u);
  *dy = y * rad_dist_u + 2.0 * c->p2 * mxy_u + c->p1 * (rho2_u + 2.0 * my2_u);
}
static int no_distortion(const o_camera* c) { return c->k1 == 0.0 && c->k2 == 0.0 && c->p1 == 0.0 && c->p2 == 0.0; }

/* D1: out = (float)(T(r,0)*x + T(r,1)*y + T(r,2)*z + T(r,3)), T row-major 3x4 */
int lmono_cpu_transform_cloud(const float* in, int n, int stride_floats, const double T[12], float* out_xyz) {
  for (int i = 0; i < n; ++i) {
    const double x = in[(size_t)i * stride_floats], y = in[(size_t)i * stride_floats + 1], z = in[(size_t)i * stride_floats + 2];
    for (int r = 0; r < 3; ++r) out_xyz[(size_t)i * 3 + r] = (float)(T[r * 4 + 0] * x + T[r * 4 + 1] * y + T[r * 4 + 2] * z + T[r * 4 + 3]);
  }
  return 0;
}

/* D2 */
int lmono_cpu_project_raster(const float* pts_cam, int n, int stride_floats, const o_camera* cam, uint8_t* depth_raw) {
  const int W = cam->width, H = cam->height;
  memset(depth_raw, 0, (size_t)W * H);
  const int nod = no_distortion(cam);
  for (int i = 0; i < n; ++i) {
    const float px = pts_cam[(size_t)i * stride_floats], py = pts_cam[(size_t)i * stride_floats + 1], pz = pts_cam[(size_t)i * stride_floats + 2];
    if (pz < 0) continue;
    double X = px, Y = py, Z = pz;
    double ux = X / Z, uy = Y / Z;
    double dxp = ux, dyp = uy;
    if (!nod) { double ddx, ddy; distortion(cam, ux, uy, &ddx, &ddy); dxp = ux + ddx; dyp = uy + ddy; }
    double u = cam->fx * dxp + cam->cx, v = cam->fy * dyp + cam->cy;
    float fxy_x = (float)u, fxy_y = (float)v;                 /* cv::Point2f */
    if (fxy_x > 0 && fxy_x < W && fxy_y > 0 && fxy_y < H) {
      double depth = pz;
      /* implicit double -> uchar: truncate to int, keep the low byte (x86-64 GCC behaviour outside [0,255]) */
      int iv = (int)(100 - depth);
      depth_raw[(size_t)(int)fxy_y * W + (int)fxy_x] = (uint8_t)iv;
    }
  }
  return 0;
}

/* ---- OpenCV imgproc restatements (8UC1) ------------------------------------------------ */
static void make_kernel(int type, int ks, uint8_t* k) {
  /* cv::getStructuringElement */
  int r = ks / 2, c = ks / 2;
  double inv_r2 = r ? 1.0 / ((double)r * r) : 0;
  for (int i = 0; i < ks; ++i) {
    int j1 = 0, j2 = 0;
    if (type == 0 || (type == 1 && i == r)) j2 = ks;                 /* RECT, CROSS centre row */
    else if (type == 1) { j1 = c; j2 = c + 1; }                        /* CROSS */
    else {                                                             /* ELLIPSE */
      int dy = i - r;
      if (abs(dy) <= r) {
        int dx = (int)lrint(c * sqrt((r * r - dy * dy) * inv_r2));    /* saturate_cast<int> = cvRound */
        j1 = c - dx > 0 ? c - dx : 0;
        j2 = c + dx + 1 < ks ? c + dx + 1 : ks;
      }
    }
    for (int j = 0; j < ks; ++j) k[i * ks + j] = (j >= j1 && j < j2) ? 1 : 0;
  }
}

/* dilate (is_erode = 0) / erode with BORDER_CONSTANT = morphologyDefaultBorderValue (outside ignored) */
static void morph(const uint8_t* src, uint8_t* dst, int W, int H, const uint8_t* k, int ks, int is_erode) {
  int a = ks / 2;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      int best = is_erode ? 255 : 0;
      for (int i = 0; i < ks; ++i) {
        int yy = y + i - a; if (yy < 0 || yy >= H) continue;
        for (int j = 0; j < ks; ++j) {
          if (!k[i * ks + j]) continue;
          int xx = x + j - a; if (xx < 0 || xx >= W) continue;
          int v = src[(size_t)yy * W + xx];
          if (is_erode) { if (v < best) best = v; } else { if (v > best) best = v; }
        }
      }
      dst[(size_t)y * W + x] = (uint8_t)best;
    }
}

static int cmp_u8(const void* a, const void* b) { return (int)*(const uint8_t*)a - (int)*(const uint8_t*)b; }

static void median5(const uint8_t* src, uint8_t* dst, int W, int H) {
  /* cv::medianBlur ksize 5, BORDER_REPLICATE */
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      uint8_t v[25]; int n = 0;
      for (int i = -2; i <= 2; ++i) {
        int yy = y + i; yy = yy < 0 ? 0 : (yy >= H ? H - 1 : yy);
        for (int j = -2; j <= 2; ++j) { int xx = x + j; xx = xx < 0 ? 0 : (xx >= W ? W - 1 : xx); v[n++] = src[(size_t)yy * W + xx]; }
      }
      qsort(v, 25, 1, cmp_u8);
      dst[(size_t)y * W + x] = v[12];
    }
}

static inline int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) { if (p < 0) p = -p; else p = 2 * len - 2 - p; }
  return p;
}

static void bilateral5(const uint8_t* src, uint8_t* dst, int W, int H, double sigma_color, double sigma_space) {
  /* cv::bilateralFilter 8u, d = 5, BORDER_REFLECT_101, scalar path of bilateralFilter_8u */
  const int radius = 2;
  double gauss_color_coeff = -0.5 / (sigma_color * sigma_color);
  double gauss_space_coeff = -0.5 / (sigma_space * sigma_space);
  float color_weight[256];
  for (int i = 0; i < 256; ++i) color_weight[i] = (float)exp(i * i * gauss_color_coeff);
  float space_weight[25]; int oi[25], oj[25]; int maxk = 0;
  for (int i = -radius; i <= radius; i++)
    for (int j = -radius; j <= radius; j++) {
      double r = sqrt((double)i * i + (double)j * j);
      if (r > radius) continue;
      space_weight[maxk] = (float)exp(r * r * gauss_space_coeff);
      oi[maxk] = i; oj[maxk] = j; maxk++;
    }
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      float sum = 0, wsum = 0;
      int val0 = src[(size_t)y * W + x];
      for (int k = 0; k < maxk; ++k) {
        int yy = reflect101(y + oi[k], H), xx = reflect101(x + oj[k], W);
        int val = src[(size_t)yy * W + xx];
        float w = space_weight[k] * color_weight[abs(val - val0)];
        sum += val * w; wsum += w;
      }
      dst[(size_t)y * W + x] = (uint8_t)lrintf(sum / wsum);
    }
}

static void gaussian5(const uint8_t* src, uint8_t* dst, int W, int H) {
  /* cv::GaussianBlur 5x5, sigma 0 -> fixed kernel {1,4,6,4,1}/16 per axis, 8u fixed point, REFLECT_101 */
  static const int kw[5] = { 1, 4, 6, 4, 1 };
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      int acc = 0;
      for (int i = -2; i <= 2; ++i) {
        int yy = reflect101(y + i, H);
        for (int j = -2; j <= 2; ++j) { int xx = reflect101(x + j, W); acc += kw[i + 2] * kw[j + 2] * src[(size_t)yy * W + xx]; }
      }
      dst[(size_t)y * W + x] = (uint8_t)((acc + 128) >> 8);
    }
}

/* test hooks for the oracle/_ref builds: the OpenCV restatements one by one, so that the stand-in cv:: functions the
 * reference's depthFill calls (refstubs/opencv2/opencv.hpp) are these very routines */
void lmono_cpu_cv_kernel(int type, int ks, uint8_t* k) { make_kernel(type, ks, k); }
void lmono_cpu_cv_morph(const uint8_t* src, uint8_t* dst, int W, int H, const uint8_t* k, int ks, int is_erode) { morph(src, dst, W, H, k, ks, is_erode); }
void lmono_cpu_cv_median5(const uint8_t* src, uint8_t* dst, int W, int H) { median5(src, dst, W, H); }
void lmono_cpu_cv_bilateral5(const uint8_t* src, uint8_t* dst, int W, int H, double sigma_color, double sigma_space) { bilateral5(src, dst, W, H, sigma_color, sigma_space); }
void lmono_cpu_cv_gaussian5(const uint8_t* src, uint8_t* dst, int W, int H) { gaussian5(src, dst, W, H); }

/* D3 */
int lmono_cpu_depth_fill(const uint8_t* depth_raw, const o_camera* cam, uint8_t* out) {
  const int W = cam->width, H = cam->height; const size_t N = (size_t)W * H;
  uint8_t *a = (uint8_t*)malloc(N), *b = (uint8_t*)malloc(N), *c = (uint8_t*)malloc(N);
  uint8_t kern[31 * 31], rect[31 * 31], rect7[49];
  int ks = cam->kernel_size; if (ks < 1) ks = 1; if (ks > 31) ks = 31;
  make_kernel(cam->kernel_type, ks, kern);
  make_kernel(0, ks, rect);
  make_kernel(0, 7, rect7);
  morph(depth_raw, a, W, H, kern, ks, 0);          /* dilate_mat (:358) */
  morph(a, b, W, H, rect, ks, 0);                  /* MORPH_CLOSE = dilate ... */
  morph(b, c, W, H, rect, ks, 1);                  /* ... then erode -> hole_fill (:362) */
  morph(c, a, W, H, rect7, 7, 0);                  /* dilate 7x7 of hole_fill (:363) */
  for (size_t i = 0; i < N; ++i) if (c[i] < 0.1) c[i] = a[i];     /* :365-374 */
  median5(c, b, W, H);                             /* :391 */
  if (cam->blur_type == 0) bilateral5(b, out, W, H, 1.5, 2.0);    /* :396 */
  else gaussian5(b, out, W, H);                    /* :399 */
  free(a); free(b); free(c);
  return 0;
}

/* D4.  QT: world pose of the camera (Q, T of associateToMap). */
int lmono_cpu_lift_cloud(const uint8_t* depth, const uint8_t* bgr, const o_camera* cam, const o_pose* QT,
                         float* cloud_cam_xyz, float* cloud_world_xyz, uint8_t* cloud_rgb, int cap, int* n_out) {
  const int W = cam->width, H = cam->height;
  const double inv_K11 = 1.0 / cam->fx, inv_K13 = -cam->cx / cam->fx, inv_K22 = 1.0 / cam->fy, inv_K23 = -cam->cy / cam->fy;
  const int nod = no_distortion(cam);
  /* Eigen Quaternion::toRotationMatrix */
  const double x = QT->q[0], y = QT->q[1], z = QT->q[2], w = QT->q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z, twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  const double R[9] = { 1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy) };
  int n = 0;
  for (int j = 0; j < H; ++j)
    for (int i = 0; i < W; ++i) {
      int depth_value = 100 - depth[(size_t)j * W + i];
      if (depth_value <= 0) continue;
      if (depth_value >= 70) continue;
      double mx_d = inv_K11 * i + inv_K13, my_d = inv_K22 * j + inv_K23, mx_u, my_u;
      if (nod) { mx_u = mx_d; my_u = my_d; }
      else {
        double dx, dy; distortion(cam, mx_d, my_d, &dx, &dy);
        mx_u = mx_d - dx; my_u = my_d - dy;
        for (int it = 1; it < 8; ++it) { distortion(cam, mx_u, my_u, &dx, &dy); mx_u = mx_d - dx; my_u = my_d - dy; }
      }
      const double bz = 1.0;
      float px = (float)(depth_value * mx_u / bz), py = (float)(depth_value * my_u / bz), pz = (float)depth_value;
      if (fabsf(px) > 20 && py > 1.8) continue;       /* :305 (float abs overload) */
      if (n < cap) {
        if (cloud_cam_xyz) { cloud_cam_xyz[(size_t)n * 3] = px; cloud_cam_xyz[(size_t)n * 3 + 1] = py; cloud_cam_xyz[(size_t)n * 3 + 2] = pz; }
        for (int r = 0; r < 3; ++r)
          cloud_world_xyz[(size_t)n * 3 + r] = (float)(R[r * 3 + 0] * (double)px + R[r * 3 + 1] * (double)py + R[r * 3 + 2] * (double)pz + QT->t[r]);
        const uint8_t* c = bgr + ((size_t)j * W + i) * 3;
        cloud_rgb[(size_t)n * 3] = c[2]; cloud_rgb[(size_t)n * 3 + 1] = c[1]; cloud_rgb[(size_t)n * 3 + 2] = c[0];
      }
      ++n;
    }
  *n_out = n;
  return 0;
}
