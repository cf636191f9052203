// LiDAR -> camera colour projection of mono_lidar_mapping on device (rows D1-D4):
//   k_col_winner / k_col_raster   Map_Builder.cc:224-245 + PinholeCamera::spaceToPlane
//                                 (camera_models PinholeCamera.cc:520-542): the reference writes
//                                 depth_map(int(v), int(u)) = 100 - z point after point, so the LAST
//                                 point in cloud order wins; reproduced with an atomicMax on the point
//                                 index per pixel, then one pass that converts the winner's depth.
//   k_col_morph .. k_col_blur     depthFill, Map_Builder.cc:336-403 (OpenCV 3.x scalar semantics:
//                                 dilate / erode with the "ignore outside" border, medianBlur 5 with
//                                 replicate border, bilateralFilter(5, 1.5, 2.0) or 5x5 Gaussian with
//                                 reflect-101 border)
//   k_col_lift_*                  Map_Builder.cc:275-322 + PinholeCamera::liftProjective (:450-510):
//                                 per pixel 0 < d < 70, ray lift, colour fetch, |x| > 20 && y > 1.8
//                                 rejection, world transform; row-major order kept by a two-kernel scan.
// All pinhole arithmetic is double as in the reference, stored as float.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>
#include <math.h>

struct ColorCam { double fx, fy, cx, cy, k1, k2, p1, p2; int W, H, nod; };
struct ColorTables { float color_w[256]; float space_w[25]; int oi[25], oj[25]; int maxk; unsigned char mask[31 * 31]; int ks; };

struct ColorState {
  int W, H, cap_pts;
  float4* d_pts; int32_t* d_winner; uint8_t* d_img[4]; uint8_t* d_bgr; size_t bgr_bytes;
  ColorTables* d_tab; ColorTables h_tab;
  int32_t* d_blockcnt; int32_t* d_nout;
  float* d_proj; int n_proj;        // (u, v, z) of every point of the last frame, u = NaN where the point was not rasterised
  float* d_cam; float* d_world; uint8_t* d_rgb;
};

__device__ __forceinline__ void d_distortion(const ColorCam& c, double x, double y, double* dx, double* dy) {
  const double mx2 = x * x, my2 = y * y, mxy = x * y;
  const double rho2 = mx2 + my2;
  const double rad = c.k1 * rho2 + c.k2 * rho2 * rho2;
  *dx = x * rad + 2.0 * c.p1 * mxy + c.p2 * (rho2 + 2.0 * mx2);
  *dy = y * rad + 2.0 * c.p2 * mxy + c.p1 * (rho2 + 2.0 * my2);
}

// D1: pcl::transformPointCloud with a double 3x4 matrix (map_build_node.cc:216-225)
struct Mat34 { double m[12]; };
__global__ void __launch_bounds__(256) k_col_transform(float4* __restrict__ pts, int n, Mat34 T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  const double x = p.x, y = p.y, z = p.z;
  float4 o;
  o.x = (float)(T.m[0] * x + T.m[1] * y + T.m[2] * z + T.m[3]);
  o.y = (float)(T.m[4] * x + T.m[5] * y + T.m[6] * z + T.m[7]);
  o.z = (float)(T.m[8] * x + T.m[9] * y + T.m[10] * z + T.m[11]);
  o.w = p.w;
  pts[i] = o;
}

__global__ void __launch_bounds__(256) k_col_winner(const float4* __restrict__ pts, int n, ColorCam c, int32_t* __restrict__ winner, float* __restrict__ proj) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  proj[3 * i + 0] = __int_as_float(0x7fc00000); proj[3 * i + 1] = __int_as_float(0x7fc00000); proj[3 * i + 2] = p.z;
  if (p.z < 0) return;
  const double X = p.x, Y = p.y, Z = p.z;
  const double ux = X / Z, uy = Y / Z;
  double dxp = ux, dyp = uy;
  if (!c.nod) { double ddx, ddy; d_distortion(c, ux, uy, &ddx, &ddy); dxp = ux + ddx; dyp = uy + ddy; }
  const float fx = (float)(c.fx * dxp + c.cx), fy = (float)(c.fy * dyp + c.cy);     // cv::Point2f
  if (fx > 0 && fx < (float)c.W && fy > 0 && fy < (float)c.H) {
    atomicMax(&winner[(int)fy * c.W + (int)fx], i + 1);
    proj[3 * i + 0] = fx; proj[3 * i + 1] = fy;       // what the reference draws its r = 3 HSV disc at (Map_Builder.cc:243)
  }
}

__global__ void __launch_bounds__(256) k_col_raster(const float4* __restrict__ pts, const int32_t* __restrict__ winner, int npix, uint8_t* __restrict__ depth) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const int w = winner[i];
  uint8_t v = 0;
  if (w > 0) v = (uint8_t)(int)(100 - (double)pts[w - 1].z);    // implicit double -> uchar of the reference
  depth[i] = v;
}

__global__ void __launch_bounds__(256) k_col_morph(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int W, int H,
                                                   const ColorTables* __restrict__ tab, int use_rect, int ks, int is_erode) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const int a = ks / 2;
  int best = is_erode ? 255 : 0;
  for (int i = 0; i < ks; ++i) {
    const int yy = y + i - a; if (yy < 0 || yy >= H) continue;
    for (int j = 0; j < ks; ++j) {
      if (!use_rect && !tab->mask[i * ks + j]) continue;
      const int xx = x + j - a; if (xx < 0 || xx >= W) continue;
      const int v = src[(size_t)yy * W + xx];
      best = is_erode ? min(best, v) : max(best, v);
    }
  }
  dst[(size_t)y * W + x] = (uint8_t)best;
}

__global__ void __launch_bounds__(256) k_col_fill(uint8_t* __restrict__ hole, const uint8_t* __restrict__ dil, int npix) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npix && hole[i] == 0) hole[i] = dil[i];       // hole_fill < 0.1  <=>  == 0 for uchar (:369)
}

__global__ void __launch_bounds__(256) k_col_median5(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int W, int H) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  // 25-value median through a 256-bin-free counting approach: rank selection by comparisons
  unsigned char v[25];
  int n = 0;
#pragma unroll
  for (int i = -2; i <= 2; ++i) {
    const int yy = min(max(y + i, 0), H - 1);
#pragma unroll
    for (int j = -2; j <= 2; ++j) { const int xx = min(max(x + j, 0), W - 1); v[n++] = src[(size_t)yy * W + xx]; }
  }
  // the median is the value with exactly 12 elements ordered before it (ties by position)
  unsigned char med = 0;
#pragma unroll
  for (int a = 0; a < 25; ++a) {
    int less = 0;
#pragma unroll
    for (int b = 0; b < 25; ++b) less += (v[b] < v[a]) || (v[b] == v[a] && b < a);
    if (less == 12) med = v[a];
  }
  dst[(size_t)y * W + x] = med;
}

__device__ __forceinline__ int d_reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) { if (p < 0) p = -p; else p = 2 * len - 2 - p; }
  return p;
}

__global__ void __launch_bounds__(256) k_col_bilateral5(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int W, int H,
                                                        const ColorTables* __restrict__ tab) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  float sum = 0.f, wsum = 0.f;
  const int val0 = src[(size_t)y * W + x];
  const int maxk = tab->maxk;
  for (int k = 0; k < maxk; ++k) {
    const int yy = d_reflect101(y + tab->oi[k], H), xx = d_reflect101(x + tab->oj[k], W);
    const int val = src[(size_t)yy * W + xx];
    const float w = __fmul_rn(tab->space_w[k], tab->color_w[abs(val - val0)]);
    sum = __fadd_rn(sum, __fmul_rn((float)val, w));
    wsum = __fadd_rn(wsum, w);
  }
  dst[(size_t)y * W + x] = (uint8_t)__float2int_rn(__fdiv_rn(sum, wsum));     // cvRound
}

__global__ void __launch_bounds__(256) k_col_gaussian5(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int W, int H) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const int kw[5] = { 1, 4, 6, 4, 1 };
  int acc = 0;
#pragma unroll
  for (int i = -2; i <= 2; ++i) {
    const int yy = d_reflect101(y + i, H);
#pragma unroll
    for (int j = -2; j <= 2; ++j) { const int xx = d_reflect101(x + j, W); acc += kw[i + 2] * kw[j + 2] * src[(size_t)yy * W + xx]; }
  }
  dst[(size_t)y * W + x] = (uint8_t)((acc + 128) >> 8);
}

// D4, pass 1 (count) and pass 2 (write); pixel index e = j * W + i in row-major order
__device__ __forceinline__ bool d_lift_pixel(const ColorCam& c, const uint8_t* __restrict__ depth, int e, float* px, float* py, float* pz) {
  const int depth_value = 100 - (int)depth[e];
  if (depth_value <= 0 || depth_value >= 70) return false;
  const int j = e / c.W, i = e - j * c.W;
  const double inv_K11 = 1.0 / c.fx, inv_K13 = -c.cx / c.fx, inv_K22 = 1.0 / c.fy, inv_K23 = -c.cy / c.fy;
  const double mx_d = inv_K11 * i + inv_K13, my_d = inv_K22 * j + inv_K23;
  double mx_u = mx_d, my_u = my_d;
  if (!c.nod) {
    double dx, dy; d_distortion(c, mx_d, my_d, &dx, &dy);
    mx_u = mx_d - dx; my_u = my_d - dy;
    for (int it = 1; it < 8; ++it) { d_distortion(c, mx_u, my_u, &dx, &dy); mx_u = mx_d - dx; my_u = my_d - dy; }
  }
  *px = (float)(depth_value * mx_u / 1.0); *py = (float)(depth_value * my_u / 1.0); *pz = (float)depth_value;
  if (fabsf(*px) > 20 && (double)*py > 1.8) return false;
  return true;
}

__global__ void __launch_bounds__(256) k_col_lift_count(ColorCam c, const uint8_t* __restrict__ depth, int npix, int32_t* __restrict__ blockcnt) {
  __shared__ int ws[33];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  float px, py, pz;
  const int flag = (e < npix && d_lift_pixel(c, depth, e, &px, &py, &pz)) ? 1 : 0;
  int total;
  d_block_exscan(flag, ws, &total);
  if (threadIdx.x == 0) blockcnt[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_col_scan_blocks(int32_t* __restrict__ blockcnt, int nblocks, int32_t* __restrict__ total_out) {
  __shared__ int ws[33];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += blockDim.x) {
    const int e = base + threadIdx.x;
    const int v = e < nblocks ? blockcnt[e] : 0;
    int total;
    const int ex = d_block_exscan(v, ws, &total);
    if (e < nblocks) blockcnt[e] = s_carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) s_carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = s_carry;
}

__global__ void __launch_bounds__(256) k_col_lift_write(ColorCam c, const uint8_t* __restrict__ depth, int npix, const int32_t* __restrict__ blockoff,
                                                        const uint8_t* __restrict__ bgr, int step, Mat34 QT, int cap,
                                                        float* __restrict__ cam_xyz, float* __restrict__ world_xyz, uint8_t* __restrict__ rgb) {
  __shared__ int ws[33];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  float px = 0.f, py = 0.f, pz = 0.f;
  const int flag = (e < npix && d_lift_pixel(c, depth, e, &px, &py, &pz)) ? 1 : 0;
  int total;
  const int ex = d_block_exscan(flag, ws, &total);
  if (!flag) return;
  const int o = blockoff[blockIdx.x] + ex;
  if (o >= cap) return;
  if (cam_xyz) { cam_xyz[(size_t)o * 3] = px; cam_xyz[(size_t)o * 3 + 1] = py; cam_xyz[(size_t)o * 3 + 2] = pz; }
  const double x = px, y = py, z = pz;
#pragma unroll
  for (int r = 0; r < 3; ++r)
    world_xyz[(size_t)o * 3 + r] = (float)(QT.m[r * 4 + 0] * x + QT.m[r * 4 + 1] * y + QT.m[r * 4 + 2] * z + QT.m[r * 4 + 3]);
  const int j = e / c.W, i = e - j * c.W;
  const uint8_t* pxl = bgr + (size_t)j * step + (size_t)i * 3;
  rgb[(size_t)o * 3] = pxl[2]; rgb[(size_t)o * 3 + 1] = pxl[1]; rgb[(size_t)o * 3 + 2] = pxl[0];
}

// ---------------------------------------------------------------------------------------------
static void make_mask(int type, int ks, unsigned char* k) {          // cv::getStructuringElement
  const int r = ks / 2, c = ks / 2;
  const double inv_r2 = r ? 1.0 / ((double)r * r) : 0;
  for (int i = 0; i < ks; ++i) {
    int j1 = 0, j2 = 0;
    if (type == 0 || (type == 1 && i == r)) j2 = ks;
    else if (type == 1) { j1 = c; j2 = c + 1; }
    else {
      const int dy = i - r;
      if (abs(dy) <= r) {
        const int dx = (int)lrint(c * sqrt((r * r - dy * dy) * inv_r2));
        j1 = c - dx > 0 ? c - dx : 0;
        j2 = c + dx + 1 < ks ? c + dx + 1 : ks;
      }
    }
    for (int j = 0; j < ks; ++j) k[i * ks + j] = (j >= j1 && j < j2) ? 1 : 0;
  }
}

static int color_state(lmono_ctx* ctx, int W, int H, ColorState** out) {
  ColorState* s = (ColorState*)ctx->color_state;
  if (s && (s->W != W || s->H != H)) return LMONO_E_ARG;
  if (s) { *out = s; return LMONO_OK; }
  s = (ColorState*)calloc(1, sizeof(ColorState));
  s->W = W; s->H = H; s->cap_pts = ctx->max_sweep;
  const size_t npix = (size_t)W * H;
  LM_CUDA(cudaMalloc((void**)&s->d_pts, (size_t)s->cap_pts * sizeof(float4)));
  LM_CUDA(cudaMalloc((void**)&s->d_winner, npix * sizeof(int32_t)));
  for (int k = 0; k < 4; ++k) LM_CUDA(cudaMalloc((void**)&s->d_img[k], npix));
  s->bgr_bytes = npix * 4;
  LM_CUDA(cudaMalloc((void**)&s->d_bgr, s->bgr_bytes));
  LM_CUDA(cudaMalloc((void**)&s->d_tab, sizeof(ColorTables)));
  LM_CUDA(cudaMalloc((void**)&s->d_blockcnt, sizeof(int32_t) * (npix / 256 + 8)));
  LM_CUDA(cudaMalloc((void**)&s->d_nout, sizeof(int32_t)));
  LM_CUDA(cudaMalloc((void**)&s->d_proj, (size_t)s->cap_pts * 3 * sizeof(float)));
  LM_CUDA(cudaMalloc((void**)&s->d_cam, npix * 3 * sizeof(float)));
  LM_CUDA(cudaMalloc((void**)&s->d_world, npix * 3 * sizeof(float)));
  LM_CUDA(cudaMalloc((void**)&s->d_rgb, npix * 3));
  ctx->color_state = s;
  *out = s;
  return LMONO_OK;
}

void lm_color_free(lmono_ctx* ctx) {
  ColorState* s = (ColorState*)ctx->color_state;
  if (!s) return;
  cudaFree(s->d_pts); cudaFree(s->d_winner); for (int k = 0; k < 4; ++k) cudaFree(s->d_img[k]);
  cudaFree(s->d_bgr); cudaFree(s->d_tab); cudaFree(s->d_blockcnt); cudaFree(s->d_nout); cudaFree(s->d_proj); cudaFree(s->d_cam); cudaFree(s->d_world); cudaFree(s->d_rgb);
  free(s); ctx->color_state = nullptr;
}

extern "C" int lmono_project_color(lmono_ctx* ctx, lmono_cloud_view pts, const double* T_cam_lidar, const uint8_t* bgr, int32_t step_bytes,
                                   const lmono_pinhole* cam, const lmono_pose* Q_T, uint8_t* depth_raw, uint8_t* depth_filled,
                                   float* cloud_cam_xyz, float* cloud_world_xyz, uint8_t* cloud_rgb, int32_t capacity, int32_t* n_out) {
  if (!ctx || !cam || !bgr || !Q_T || !n_out || !cloud_world_xyz || !cloud_rgb) return LMONO_E_ARG;
  const int W = cam->width, H = cam->height;
  if (W <= 0 || H <= 0 || step_bytes < W * 3 || cam->kernel_size < 1 || cam->kernel_size > 31) return LMONO_E_ARG;
  if (pts.n > ctx->max_sweep) return LMONO_E_CAPACITY;
  ColorState* s; int rc = color_state(ctx, W, H, &s); if (rc) return rc;
  const int npix = W * H;
  if ((size_t)step_bytes * H > s->bgr_bytes) return LMONO_E_CAPACITY;
  ColorCam c; c.fx = cam->fx; c.fy = cam->fy; c.cx = cam->cx; c.cy = cam->cy; c.k1 = cam->k1; c.k2 = cam->k2; c.p1 = cam->p1; c.p2 = cam->p2;
  c.W = W; c.H = H; c.nod = (cam->k1 == 0.0 && cam->k2 == 0.0 && cam->p1 == 0.0 && cam->p2 == 0.0) ? 1 : 0;
  // tables (host libm, exactly the values OpenCV computes on the host)
  ColorTables& t = s->h_tab;
  const double gcc = -0.5 / (1.5 * 1.5), gsc = -0.5 / (2.0 * 2.0);
  for (int i = 0; i < 256; ++i) t.color_w[i] = (float)exp(i * i * gcc);
  t.maxk = 0;
  for (int i = -2; i <= 2; i++) for (int j = -2; j <= 2; j++) {
    const double r = sqrt((double)i * i + (double)j * j);
    if (r > 2) continue;
    t.space_w[t.maxk] = (float)exp(r * r * gsc); t.oi[t.maxk] = i; t.oj[t.maxk] = j; t.maxk++;
  }
  t.ks = cam->kernel_size;
  make_mask(cam->kernel_type, t.ks, t.mask);
  LM_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  LM_CUDA(cudaMemcpyAsync(s->d_tab, &t, sizeof(ColorTables), cudaMemcpyHostToDevice, ctx->stream));
  LM_CUDA(cudaMemcpyAsync(s->d_bgr, bgr, (size_t)step_bytes * H, cudaMemcpyHostToDevice, ctx->stream));
  if ((rc = lm_upload_cloud(ctx, pts, ctx->d_raw[2], s->d_pts, nullptr))) return rc;
  const int nb_pts = lm_div_up(pts.n > 0 ? pts.n : 1, 256), nb_pix = lm_div_up(npix, 256);
  LM_CUDA(cudaEventRecord(ctx->ev_k0, ctx->stream));
  lm_kmark(ctx, "begin", 0);
  if (T_cam_lidar && pts.n > 0) {
    Mat34 T; memcpy(T.m, T_cam_lidar, sizeof(T.m));
    k_col_transform<<<nb_pts, 256, 0, ctx->stream>>>(s->d_pts, pts.n, T); LM_LAUNCH_CHECK();
  }
  LM_CUDA(cudaMemsetAsync(s->d_winner, 0, (size_t)npix * sizeof(int32_t), ctx->stream));
  if (pts.n > 0) { k_col_winner<<<nb_pts, 256, 0, ctx->stream>>>(s->d_pts, pts.n, c, s->d_winner, s->d_proj); LM_LAUNCH_CHECK(); }
  s->n_proj = pts.n;
  uint8_t *raw = s->d_img[0], *a = s->d_img[1], *b = s->d_img[2], *cc = s->d_img[3];
  k_col_raster<<<nb_pix, 256, 0, ctx->stream>>>(s->d_pts, s->d_winner, npix, raw); LM_LAUNCH_CHECK();
  const dim3 g2(lm_div_up(W, 256), H);
  k_col_morph<<<g2, 256, 0, ctx->stream>>>(raw, a, W, H, s->d_tab, 0, t.ks, 0); LM_LAUNCH_CHECK();      // dilate(K)          :358
  k_col_morph<<<g2, 256, 0, ctx->stream>>>(a, b, W, H, s->d_tab, 1, t.ks, 0); LM_LAUNCH_CHECK();        // close: dilate      :362
  k_col_morph<<<g2, 256, 0, ctx->stream>>>(b, cc, W, H, s->d_tab, 1, t.ks, 1); LM_LAUNCH_CHECK();       //        erode
  k_col_morph<<<g2, 256, 0, ctx->stream>>>(cc, a, W, H, s->d_tab, 1, 7, 0); LM_LAUNCH_CHECK();          // dilate 7x7         :363
  k_col_fill<<<nb_pix, 256, 0, ctx->stream>>>(cc, a, npix); LM_LAUNCH_CHECK();                          // :365-374
  k_col_median5<<<g2, 256, 0, ctx->stream>>>(cc, b, W, H); LM_LAUNCH_CHECK();                           // :391
  if (cam->blur_type == 0) { k_col_bilateral5<<<g2, 256, 0, ctx->stream>>>(b, a, W, H, s->d_tab); LM_LAUNCH_CHECK(); }   // :396
  else { k_col_gaussian5<<<g2, 256, 0, ctx->stream>>>(b, a, W, H); LM_LAUNCH_CHECK(); }                                 // :399
  // D4
  double R[9];
  { const double x = Q_T->q[0], y = Q_T->q[1], z = Q_T->q[2], w = Q_T->q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z, twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy; R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy); }
  Mat34 QT;
  for (int r = 0; r < 3; ++r) { for (int k = 0; k < 3; ++k) QT.m[r * 4 + k] = R[r * 3 + k]; QT.m[r * 4 + 3] = Q_T->t[r]; }
  k_col_lift_count<<<nb_pix, 256, 0, ctx->stream>>>(c, a, npix, s->d_blockcnt); LM_LAUNCH_CHECK();
  k_col_scan_blocks<<<1, 1024, 0, ctx->stream>>>(s->d_blockcnt, nb_pix, s->d_nout); LM_LAUNCH_CHECK();
  k_col_lift_write<<<nb_pix, 256, 0, ctx->stream>>>(c, a, npix, s->d_blockcnt, s->d_bgr, step_bytes, QT, npix,
                                                    cloud_cam_xyz ? s->d_cam : nullptr, s->d_world, s->d_rgb); LM_LAUNCH_CHECK();
  LM_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  int n = 0;
  LM_CUDA(cudaMemcpyAsync(&n, s->d_nout, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (depth_raw) LM_CUDA(cudaMemcpyAsync(depth_raw, raw, (size_t)npix, cudaMemcpyDeviceToHost, ctx->stream));
  if (depth_filled) LM_CUDA(cudaMemcpyAsync(depth_filled, a, (size_t)npix, cudaMemcpyDeviceToHost, ctx->stream));
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  *n_out = n;
  if (n > capacity) return LMONO_E_CAPACITY;
  if (n > 0) {
    if (cloud_cam_xyz) LM_CUDA(cudaMemcpyAsync(cloud_cam_xyz, s->d_cam, (size_t)n * 12, cudaMemcpyDeviceToHost, ctx->stream));
    LM_CUDA(cudaMemcpyAsync(cloud_world_xyz, s->d_world, (size_t)n * 12, cudaMemcpyDeviceToHost, ctx->stream));
    LM_CUDA(cudaMemcpyAsync(cloud_rgb, s->d_rgb, (size_t)n * 3, cudaMemcpyDeviceToHost, ctx->stream));
    LM_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return LMONO_OK;
}

// Per-point projection of the most recent lmono_project_color call, in cloud order: uvz[3 i] = u, [3 i + 1] = v (the
// cv::Point2f of Map_Builder.cc:234, NaN where the point was behind the camera or outside the frame), [3 i + 2] = depth
// z in the camera frame.  The node draws the ~pro_map debug image from it exactly as the reference does
// (cv::circle(HSV, xy, 3, Scalar(int(clip(z, 0, 100) * 6), 255, 255), -1) in cloud order, Map_Builder.cc:240-243).
extern "C" int lmono_color_projection(lmono_ctx* ctx, float* uvz, int32_t n) {
  ColorState* s = ctx ? (ColorState*)ctx->color_state : nullptr;
  if (!s || !uvz || n < 0) return LMONO_E_ARG;
  if (n > s->n_proj) return LMONO_E_STATE;
  if (n == 0) return LMONO_OK;
  LM_CUDA(cudaMemcpyAsync(uvz, s->d_proj, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  return LMONO_OK;
}
