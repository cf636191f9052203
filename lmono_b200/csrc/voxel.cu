// pcl::VoxelGrid<PointXYZI> on device, as the reference configures it (centroid of all four
// fields, output ascending in PCL's linear voxel index).  Call sites replaced:
// Aloam/src/laserMapping.cpp:542-550 (incoming corner/surf features) and
// Aloam/src/scanRegistration.cpp:401-405 (per-ring less-flat downsample, see scanreg.cu).
// Arithmetic follows PCL 1.8 voxel_grid.hpp: bbox -> min_b/div_b, per-point
// ijk = (int)(floor(p * inv_leaf) - (float)min_b), idx = ijk . divb_mul; members summed in
// fp32 in (idx, input index) order and divided by (float)count.
#include "common.cuh"
#include <float.h>

__global__ void __launch_bounds__(1024) k_vg_bbox(const float4* __restrict__ pts, const int32_t* __restrict__ n_dev,
                                                  float inv_leaf, VgParams* __restrict__ vg) {
  __shared__ float smn[3][32], smx[3][32];
  const int n = *n_dev;
  float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, mx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float4 p = pts[i];
    mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
    mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
    mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
    if (lane == 0) { smn[d][wid] = mn[d]; smx[d][wid] = mx[d]; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = blockDim.x >> 5;
    for (int d = 0; d < 3; ++d) for (int w = 1; w < nw; ++w) { smn[d][0] = fminf(smn[d][0], smn[d][w]); smx[d][0] = fmaxf(smx[d][0], smx[d][w]); }
    vg->n = n;
    if (n <= 0) { vg->guard = 0; for (int d = 0; d < 3; ++d) { vg->min_b[d] = 0; vg->div_b[d] = 1; vg->mul[d] = 0; } return; }
    long long dd[3];
    for (int d = 0; d < 3; ++d) dd[d] = (long long)(__fmul_rn(__fsub_rn(smx[d][0], smn[d][0]), inv_leaf)) + 1;
    vg->guard = (dd[0] * dd[1] * dd[2] > (long long)INT32_MAX) ? 1 : 0;
    int maxb[3];
    for (int d = 0; d < 3; ++d) {
      vg->min_b[d] = (int)floorf(__fmul_rn(smn[d][0], inv_leaf));
      maxb[d] = (int)floorf(__fmul_rn(smx[d][0], inv_leaf));
      vg->div_b[d] = maxb[d] - vg->min_b[d] + 1;
    }
    vg->mul[0] = 1; vg->mul[1] = vg->div_b[0]; vg->mul[2] = vg->div_b[0] * vg->div_b[1];
  }
}

__global__ void __launch_bounds__(256) k_vg_keys(const float4* __restrict__ pts, float inv_leaf,
                                                 const VgParams* __restrict__ vg, unsigned long long* __restrict__ comp) {
  const int n = vg->n;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pts[i];
  int ijk0 = (int)(__fsub_rn(floorf(__fmul_rn(p.x, inv_leaf)), (float)vg->min_b[0]));
  int ijk1 = (int)(__fsub_rn(floorf(__fmul_rn(p.y, inv_leaf)), (float)vg->min_b[1]));
  int ijk2 = (int)(__fsub_rn(floorf(__fmul_rn(p.z, inv_leaf)), (float)vg->min_b[2]));
  int idx = ijk0 * vg->mul[0] + ijk1 * vg->mul[1] + ijk2 * vg->mul[2];
  comp[i] = ((unsigned long long)(uint32_t)idx << 32) | (uint32_t)i;
}

// heads of voxel runs in the sorted composite array, counted per block
__global__ void __launch_bounds__(256) k_vg_count(const unsigned long long* __restrict__ sorted,
                                                  const VgParams* __restrict__ vg, int32_t* __restrict__ blockcnt) {
  __shared__ int ws[33];
  const int n = vg->n;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int head = 0;
  if (i < n) head = (i == 0) || ((sorted[i] >> 32) != (sorted[i - 1] >> 32));
  int total;
  d_block_exscan(head, ws, &total);
  if (threadIdx.x == 0) blockcnt[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256) k_vg_write(const float4* __restrict__ pts, const unsigned long long* __restrict__ sorted,
                                                  const VgParams* __restrict__ vg, const int32_t* __restrict__ blockcnt,
                                                  float4* __restrict__ out, int32_t* __restrict__ out_n) {
  __shared__ int ws[33];
  __shared__ int s_base;
  const int n = vg->n;
  if (vg->guard) {   // PCL returns the input cloud unchanged
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = pts[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) *out_n = n;
    return;
  }
  // offset of this block = sum of the head counts of the blocks before it (few hundred at most)
  if (threadIdx.x < 32) {
    int acc = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += 32) acc += blockcnt[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) s_base = acc;
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int head = 0;
  unsigned long long me = 0;
  if (i < n) { me = sorted[i]; head = (i == 0) || ((me >> 32) != (sorted[i - 1] >> 32)); }
  int total;
  int off = d_block_exscan(head, ws, &total);
  if (head) {
    const uint32_t key = (uint32_t)(me >> 32);
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    int cnt = 0;
    for (int j = i; j < n; ++j) {
      unsigned long long c = sorted[j];
      if ((uint32_t)(c >> 32) != key) break;
      float4 p = pts[(uint32_t)c];
      sx = __fadd_rn(sx, p.x); sy = __fadd_rn(sy, p.y); sz = __fadd_rn(sz, p.z); si = __fadd_rn(si, p.w);
      ++cnt;
    }
    const float c = (float)cnt;
    out[s_base + off] = make_float4(__fdiv_rn(sx, c), __fdiv_rn(sy, c), __fdiv_rn(sz, c), __fdiv_rn(si, c));
  }
  if (i == n - 1 || (n == 0 && i == 0)) *out_n = (n == 0) ? 0 : (s_base + off + head);
}

int lm_voxel_grid_device(lmono_ctx* ctx, const float4* in, const int32_t* n_dev, int n_max, float leaf,
                         float4* out, int32_t* out_n_dev) {
  const float inv_leaf = 1.0f / leaf;
  k_vg_bbox<<<1, 1024, 0, ctx->stream>>>(in, n_dev, inv_leaf, ctx->d_vg);
  LM_LAUNCH_CHECK();
  if (n_max <= 0) {   // still define out_n
    LM_CUDA(cudaMemsetAsync(out_n_dev, 0, sizeof(int32_t), ctx->stream));
    return LMONO_OK;
  }
  const int blocks = lm_div_up(n_max, 256);
  k_vg_keys<<<blocks, 256, 0, ctx->stream>>>(in, inv_leaf, ctx->d_vg, ctx->d_sort_a);
  LM_LAUNCH_CHECK();
  int rc = lm_sort_u64(ctx, ctx->d_sort_a, ctx->d_sort_b, ctx->d_sort_c, &ctx->d_vg->n, n_max);
  if (rc) return rc;
  k_vg_count<<<blocks, 256, 0, ctx->stream>>>(ctx->d_sort_c, ctx->d_vg, ctx->d_blockcnt);
  LM_LAUNCH_CHECK();
  k_vg_write<<<blocks, 256, 0, ctx->stream>>>(in, ctx->d_sort_c, ctx->d_vg, ctx->d_blockcnt, out, out_n_dev);
  LM_LAUNCH_CHECK();
  return LMONO_OK;
}
