// pcl::VoxelGrid<PointXYZI> on device, as the reference configures it (centroid of all four
// fields, output ascending in PCL's linear voxel index).  Call sites replaced:
// Aloam/src/laserMapping.cpp:542-550 (incoming corner/surf features) and
// Aloam/src/scanRegistration.cpp:401-405 (per-ring less-flat downsample, see scanreg.cu).
// Arithmetic follows PCL 1.8 voxel_grid.hpp: per-point voxel coordinate floor(p * inv_leaf) in fp32;
// members of a voxel summed in fp32 in (voxel, input index) order and divided by (float)count.
//
// PCL sorts on idx = ijk . (1, dx, dx*dy) with ijk = floor(p * inv_leaf) - min_b: that order is the
// lexicographic (z, y, x) order of the ABSOLUTE voxel coordinates, independent of the bounding box.
// The sort key here is therefore the absolute coordinate triple (3 x 15 bits, biased) | input index
// (19 bits), which needs no bounding-box pass in front of the sort; the box is only needed for PCL's
// overflow guard ("leaf size too small": output = input) and is reduced on the side with atomics.
// Both clouds of a sweep (corner, surf) go through the same four launches:
//   k_vg_keys -> k_sort_tiles -> k_merge_ranks -> k_vg_write
#include "common.cuh"
#include <float.h>

constexpr int VG_BIAS = 16384;           // voxel coordinates in [-16384, 16383]
constexpr int VG_IDX_BITS = 19;          // input index < 524288
constexpr int VW_THREADS = 256, VW_BLOCK = VW_THREADS;      // one sorted position per thread

struct VgSegs {
  const float4* in[LM_SORT_MAXSEG]; float4* out[LM_SORT_MAXSEG];
  const float4* const* in_ind[LM_SORT_MAXSEG];      // not NULL: the input pointer is read from device memory (graph replay)
  const int32_t* n[LM_SORT_MAXSEG]; int32_t* out_n[LM_SORT_MAXSEG];
  float inv_leaf[LM_SORT_MAXSEG];
  int off[LM_SORT_MAXSEG];
  int fetch;      // 1: segment s may be fetched from host memory (LmMapState::in_src / in_stride / in_ioff [s]) into in_ptr[s]
};

__device__ __forceinline__ uint32_t d_f2ord(float f) { uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float d_ord2f(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void k_vg_reset(VgParams* vg) {
  if (threadIdx.x < LM_SORT_MAXSEG) {
    VgParams& v = vg[threadIdx.x];
    for (int d = 0; d < 3; ++d) { v.mn[d] = 0xFFFFFFFFu; v.mx[d] = 0u; }
    v.ticket = 0; v.n = 0;
  }
}

__global__ void __launch_bounds__(256) k_vg_keys(VgSegs sg, VgParams* __restrict__ vgs, unsigned long long* __restrict__ comp,
                                                 LmMapState* __restrict__ st) {
  lm_pdl_enter();
  __shared__ uint32_t smn[3][8], smx[3][8];
  const int seg = blockIdx.y;
  const int n = *sg.n[seg];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x * blockDim.x >= n) return;
  uint32_t mn[3] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu }, mx[3] = { 0u, 0u, 0u };
  if (i < n) {
    const float4* src = sg.in_ind[seg] ? *sg.in_ind[seg] : sg.in[seg];
    float4 p;
    const int stride = sg.fetch ? st->in_stride[seg] : 0;
    if (stride) {
      // fused upload: the record comes over PCIe from the caller's page-locked AoS buffer (read once), the float4 copy
      // that k_vg_write gathers from stays in device memory
      const uint8_t* rec = (const uint8_t*)st->in_src[seg] + (size_t)i * stride;
      const int ioff = st->in_ioff[seg];
      if (stride == 16 && ioff == 12 && ((uintptr_t)st->in_src[seg] & 15) == 0) p = __ldcs(reinterpret_cast<const float4*>(rec));
      else {
        const float* f = reinterpret_cast<const float*>(rec);
        p.x = __ldcs(f); p.y = __ldcs(f + 1); p.z = __ldcs(f + 2);
        p.w = ioff >= 0 ? __ldcs(reinterpret_cast<const float*>(rec + ioff)) : 0.0f;
      }
      const_cast<float4*>(src)[i] = p;
    } else p = src[i];
    const float il = sg.inv_leaf[seg];
    int v[3] = { (int)floorf(__fmul_rn(p.x, il)), (int)floorf(__fmul_rn(p.y, il)), (int)floorf(__fmul_rn(p.z, il)) };
    bool bad = false;
#pragma unroll
    for (int d = 0; d < 3; ++d) { v[d] += VG_BIAS; if (v[d] < 0 || v[d] >= 2 * VG_BIAS) { bad = true; v[d] = min(max(v[d], 0), 2 * VG_BIAS - 1); } }
    if (bad || i >= (1 << VG_IDX_BITS)) atomicOr(&st->fault, LM_FAULT_FEATURE_OVERFLOW);
    const unsigned long long key = ((unsigned long long)v[2] << 30) | ((unsigned long long)v[1] << 15) | (unsigned long long)v[0];
    comp[sg.off[seg] + i] = (key << VG_IDX_BITS) | (unsigned long long)i;
    mn[0] = mx[0] = d_f2ord(p.x); mn[1] = mx[1] = d_f2ord(p.y); mn[2] = mx[2] = d_f2ord(p.z);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = min(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = max(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
    if (lane == 0) { smn[d][wid] = mn[d]; smx[d][wid] = mx[d]; }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int d = threadIdx.x;
    uint32_t a = smn[d][0], b = smx[d][0];
    for (int w = 1; w < 8; ++w) { a = min(a, smn[d][w]); b = max(b, smx[d][w]); }
    atomicMin(&vgs[seg].mn[d], a);
    atomicMax(&vgs[seg].mx[d], b);
  }
}

// sorted composites -> one centroid per voxel run, in run order.  Each block owns VW_BLOCK consecutive
// sorted positions; its output offset = number of run heads before it, recounted from the (L2-resident,
// <= 128 KB) sorted array instead of a separate scan launch.
__global__ void __launch_bounds__(VW_THREADS) k_vg_write(VgSegs sg, VgParams* __restrict__ vgs, const unsigned long long* __restrict__ sorted_all) {
  lm_pdl_enter();
  __shared__ int ws[33];
  __shared__ int s_guard;
  __shared__ unsigned long long s_key[VW_THREADS];
  __shared__ float4 s_pt[VW_THREADS];
  const int seg = blockIdx.y;
  const int n = *sg.n[seg];
  VgParams& vg = vgs[seg];
  const float4* __restrict__ pts = sg.in_ind[seg] ? *sg.in_ind[seg] : sg.in[seg];
  float4* __restrict__ out = sg.out[seg];
  const unsigned long long* __restrict__ sorted = sorted_all + sg.off[seg];
  const int base = blockIdx.x * VW_BLOCK;
  const int nblocks_active = max(1, (n + VW_BLOCK - 1) / VW_BLOCK);
  if ((int)blockIdx.x >= nblocks_active) return;
  if (threadIdx.x == 0) {
    int guard = 0;
    if (n > 0) {
      const float il = sg.inv_leaf[seg];
      long long dd[3];
      for (int d = 0; d < 3; ++d) dd[d] = (long long)(__fmul_rn(__fsub_rn(d_ord2f(vg.mx[d]), d_ord2f(vg.mn[d])), il)) + 1;
      guard = (dd[0] * dd[1] * dd[2] > (long long)INT32_MAX) ? 1 : 0;     // PCL: "Leaf size is too small"
    }
    s_guard = guard;
  }
  __syncthreads();
  if (n == 0) {
    if (threadIdx.x == 0) *sg.out_n[seg] = 0;
  } else if (s_guard) {      // PCL returns the input cloud unchanged
    for (int i = base + threadIdx.x; i < min(n, base + VW_BLOCK); i += VW_THREADS) out[i] = pts[i];
    if (blockIdx.x == 0 && threadIdx.x == 0) *sg.out_n[seg] = n;
  } else {
    // one sorted position per thread: its composite and its point are fetched at once (two dependent round trips for the
    // whole block), a run head then sums its members out of shared memory in sorted (= input index) order; only a run
    // that crosses the block's end goes back to global memory
    const int p = base + threadIdx.x;
    const bool valid = p < n;
    const unsigned long long me = valid ? sorted[p] : ~0ULL;
    const unsigned long long key = me >> VG_IDX_BITS;
    const unsigned long long prevk = (valid && p > 0) ? (sorted[p - 1] >> VG_IDX_BITS) : ~0ULL;
    const float4 mine = valid ? pts[(uint32_t)(me & ((1u << VG_IDX_BITS) - 1u))] : make_float4(0.f, 0.f, 0.f, 0.f);
    // heads before this block (independent loads, 16 in flight)
    int before = 0;
#pragma unroll 16
    for (int i = threadIdx.x; i < base; i += VW_THREADS)
      before += (i == 0) || ((sorted[i] >> VG_IDX_BITS) != (sorted[i > 0 ? i - 1 : 0] >> VG_IDX_BITS));
    s_key[threadIdx.x] = valid ? key : ~0ULL;
    s_pt[threadIdx.x] = mine;
    int total_before;
    d_block_exscan(before, ws, &total_before);           // (its barriers also publish s_key / s_pt)
    const int head = valid && (p == 0 || key != prevk);
    int total;
    const int off = total_before + d_block_exscan(head, ws, &total);
    if (head) {
      float sx = mine.x, sy = mine.y, sz = mine.z, si = mine.w;
      int c = 1;
      int m = threadIdx.x + 1;
      for (; m < VW_THREADS && s_key[m] == key; ++m) {
        const float4 q = s_pt[m];
        sx = __fadd_rn(sx, q.x); sy = __fadd_rn(sy, q.y); sz = __fadd_rn(sz, q.z); si = __fadd_rn(si, q.w);
        ++c;
      }
      if (m == VW_THREADS) {
        for (int j = base + VW_THREADS; j < n; ++j) {
          const unsigned long long cj = sorted[j];
          if ((cj >> VG_IDX_BITS) != key) break;
          const float4 q = pts[(uint32_t)(cj & ((1u << VG_IDX_BITS) - 1u))];
          sx = __fadd_rn(sx, q.x); sy = __fadd_rn(sy, q.y); sz = __fadd_rn(sz, q.z); si = __fadd_rn(si, q.w);
          ++c;
        }
      }
      const float cf = (float)c;
      out[off] = make_float4(__fdiv_rn(sx, cf), __fdiv_rn(sy, cf), __fdiv_rn(sz, cf), __fdiv_rn(si, cf));
    }
    if (p == n - 1) *sg.out_n[seg] = off + head;          // the thread that owns the last position
  }
  // the last block to finish re-arms the bounding box for the next call
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicAdd(&vg.ticket, 1u);
    if (t == (unsigned)nblocks_active - 1) {
      for (int d = 0; d < 3; ++d) { vg.mn[d] = 0xFFFFFFFFu; vg.mx[d] = 0u; }
      vg.ticket = 0;
    }
  }
}

// VoxelGrid of up to LM_SORT_MAXSEG clouds with shared launches.  Scratch: ctx->d_sort_a/b/c (segment s at
// offset sum of n_max of the segments before it), ctx->d_vg.
int lm_voxel_grid_multi(lmono_ctx* ctx, int nseg, const float4* const* in, const int32_t* const* n_dev, const int* n_max,
                        const float* leaf, float4* const* out, int32_t* const* out_n_dev, const float4* const* const* in_ind, bool fetch) {
  if (nseg < 1 || nseg > LM_SORT_MAXSEG) return LMONO_E_ARG;
  VgSegs sg; LmSortSegs ss;
  sg.fetch = fetch ? 1 : 0;
  ss.in = ctx->d_sort_a; ss.tmp = ctx->d_sort_b; ss.out = ctx->d_sort_c;
  int off = 0, mx = 0;
  for (int s = 0; s < LM_SORT_MAXSEG; ++s) {
    const int k = s < nseg ? s : 0;
    sg.in[s] = in ? in[k] : nullptr; sg.in_ind[s] = in_ind ? in_ind[k] : nullptr; sg.out[s] = out[k]; sg.n[s] = n_dev[k]; sg.out_n[s] = out_n_dev[k]; sg.inv_leaf[s] = 1.0f / leaf[k];
    sg.off[s] = s < nseg ? off : 0; ss.off[s] = sg.off[s]; ss.n[s] = n_dev[k];
    if (s < nseg) { off += n_max[s]; mx = n_max[s] > mx ? n_max[s] : mx; }
  }
  if (mx > 0) {
    LM_LAUNCH_PDL(k_vg_keys, dim3(lm_div_up(mx, 256), nseg), 256, 0, sg, ctx->d_vg, ctx->d_sort_a, ctx->d_state);
    LM_LAUNCH_CHECK();
    int rc = lm_sort_u64_segs(ctx, ss, nseg, n_max);
    if (rc) return rc;
  }
  LM_LAUNCH_PDL(k_vg_write, dim3(max(1, lm_div_up(mx, VW_BLOCK)), nseg), VW_THREADS, 0, sg, ctx->d_vg, ctx->d_sort_c);
  LM_LAUNCH_CHECK();
  return LMONO_OK;
}

int lm_voxel_grid_device(lmono_ctx* ctx, const float4* in, const int32_t* n_dev, int n_max, float leaf,
                         float4* out, int32_t* out_n_dev) {
  return lm_voxel_grid_multi(ctx, 1, &in, &n_dev, &n_max, &leaf, &out, &out_n_dev, nullptr, false);
}

int lm_voxel_init(lmono_ctx* ctx) {
  k_vg_reset<<<1, 32, 0, ctx->stream>>>(ctx->d_vg);
  LM_LAUNCH_CHECK();
  return LMONO_OK;
}
