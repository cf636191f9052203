// scanRegistration on device: replaces laserCloudHandler, Aloam/src/scanRegistration.cpp:132-408.
//
//   k_scan_classify  :136-137,166-200  NaN / range filter, elevation -> ring id, stable in-block
//                                      rank per ring (warp match), per-block ring histogram,
//                                      first / last surviving point (for startOri / endOri)
//   k_scan_halfpass  :208-224          index of the first kept point with ori - startOri > pi
//                                      (the reference's sequential halfPassed flag)
//   k_scan_blockscan :246-252          per-ring exclusive scan of the block histograms
//   k_scan_scatter   :208-252          relTime, intensity = ring + 0.1 relTime, stable ring-major
//                                      reorder, scanStartInd / scanEndInd
//   k_scan_curvature :256-266          11-tap curvature, strictly left-to-right fp32
//   k_scan_ring      :277-405          one CTA per ring: six sector sorts (one warp each, keys in
//                                      registers), the serial greedy pick with +-5 suppression,
//                                      less-flat gather and the 0.2 m VoxelGrid of the ring
//   k_scan_compact   :304-310,356,407  ring-ordered output clouds
//
// fp32 arithmetic uses __fmul_rn/__fadd_rn in the reference's association order; comparisons
// against the double literals 0.1 / 0.05 widen the float first, as C++ does.
#include "common.cuh"
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <float.h>

constexpr int SC_THREADS = 256;
constexpr int SC_NRING = 65;            // 64 rings + bucket 64 = dropped
constexpr int SC_RING_MAX = 4096;       // points per ring handled by k_scan_ring
constexpr int SC_PICK_STRIDE = 26;      // per (ring, sector): 2 sharp + 20 less-sharp + 4 flat indices
#define SC_PI 3.14159265358979323846

struct ScanMeta {
  int32_t first_valid, last_valid, i_star, n_kept;
  int32_t out_n[4];                     // sharp, less_sharp, flat, less_flat -- directly behind n_kept: &n_kept is the 5-count array the later stages read on the device
  int32_t ring_total[SC_NRING + 3];
  int32_t ring_start[SC_NRING + 3];
  float start_ori, end_ori;
  uint32_t fault;
  int32_t n_in;                         // raw sweep size for launches that do not carry it as an argument (graph replay of the fused sweep); survives k_scan_meta_init
};

static_assert(offsetof(ScanMeta, out_n) == offsetof(ScanMeta, n_kept) + sizeof(int32_t), "lm_scan_outputs hands out &n_kept as counts[5]");

struct ScanState {
  int cap, nb_max;
  float4* d_in; int32_t* d_key; int32_t* d_block_hist; int32_t* d_block_off;
  ScanMeta* d_meta; ScanMeta* h_meta;
  float4* d_full; int32_t* d_src; float* d_curv; int32_t* d_label;
  int32_t* d_pick_idx; int32_t* d_pick_cnt;
  float4* d_lf_tmp; int32_t* d_lf_cnt;
  float4* d_out[4];
};

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int d_ring_of(float x, float y, float z, int n_scans) {
  // :166 angle = atan(z / sqrt(x*x + y*y)) * 180 / M_PI  (double libm overloads, float product-sum)
  const float r2 = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
  const float angle = (float)(atan((double)z / sqrt((double)r2)) * 180 / SC_PI);
  int scanID;
  if (n_scans == 16) {
    scanID = (int)((double)(__fdiv_rn(__fadd_rn(angle, 15.0f), 2.0f)) + 0.5);
    if (scanID > (n_scans - 1) || scanID < 0) return -1;
  } else if (n_scans == 32) {
    scanID = (int)(((double)angle + 92.0 / 3.0) * 3.0 / 4.0);
    if (scanID > (n_scans - 1) || scanID < 0) return -1;
  } else {
    if ((double)angle >= -8.83) scanID = (int)((2 - (double)angle) * 3.0 + 0.5);
    else scanID = n_scans / 2 + (int)((-8.83 - (double)angle) * 2.0 + 0.5);
    if ((double)angle > 2 || (double)angle < -24.33 || scanID > 50 || scanID < 0) return -1;
  }
  return scanID;
}

// std::atan2(float, float) of the reference's platform: glibc's atan2f / atanf up to 2.40 are the fdlibm float routines
// (argument reduction against atan(0.5), atan(1), atan(1.5), atan(inf) split in hi + lo parts, an 11-term odd
// polynomial), under 1 ulp but NOT correctly rounded.  The azimuth feeds discrete decisions (:211-233: which side of
// startOri - pi/2, endOri + pi/2 ... a point falls on), where one ulp flips relTime by a whole revolution, so the device
// evaluates the same float operations in the same order and returns the same bits (checked against glibc 2.39 on 6e8
// arguments on the host; -fmad=false and the explicit _rn intrinsics keep every product and sum separately rounded).
__device__ __forceinline__ float d_atanf_fdlibm(float x) {
  const uint32_t hx = __float_as_uint(x), ix = hx & 0x7fffffffu;
  const float hi3 = 1.5707962513e+00f, lo3 = 7.5497894159e-08f;
  int id;
  if (ix >= 0x4c000000u) { if (ix > 0x7f800000u) return __fadd_rn(x, x); return (hx >> 31) ? __fsub_rn(-hi3, lo3) : __fadd_rn(hi3, lo3); }
  if (ix < 0x3ee00000u) { if (ix < 0x31000000u) return x; id = -1; }
  else {
    x = fabsf(x);
    if (ix < 0x3f980000u) {
      if (ix < 0x3f300000u) { id = 0; x = __fdiv_rn(__fsub_rn(__fmul_rn(2.0f, x), 1.0f), __fadd_rn(2.0f, x)); }
      else { id = 1; x = __fdiv_rn(__fsub_rn(x, 1.0f), __fadd_rn(x, 1.0f)); }
    } else {
      if (ix < 0x401c0000u) { id = 2; x = __fdiv_rn(__fsub_rn(x, 1.5f), __fadd_rn(1.0f, __fmul_rn(1.5f, x))); }
      else { id = 3; x = __fdiv_rn(-1.0f, x); }
    }
  }
  const float z = __fmul_rn(x, x), w = __fmul_rn(z, z);
  float a = 1.6285819933e-02f;                        // odd terms: aT[10], [8], [6], [4], [2], [0]
  a = __fadd_rn(4.9768779427e-02f, __fmul_rn(w, a));
  a = __fadd_rn(6.6610731184e-02f, __fmul_rn(w, a));
  a = __fadd_rn(9.0908870101e-02f, __fmul_rn(w, a));
  a = __fadd_rn(1.4285714924e-01f, __fmul_rn(w, a));
  a = __fadd_rn(3.3333334327e-01f, __fmul_rn(w, a));
  const float s1 = __fmul_rn(z, a);
  float b = -3.6531571299e-02f;                       // even terms: aT[9], [7], [5], [3], [1]
  b = __fadd_rn(-5.8335702866e-02f, __fmul_rn(w, b));
  b = __fadd_rn(-7.6918758452e-02f, __fmul_rn(w, b));
  b = __fadd_rn(-1.1111110449e-01f, __fmul_rn(w, b));
  b = __fadd_rn(-2.0000000298e-01f, __fmul_rn(w, b));
  const float s2 = __fmul_rn(w, b);
  const float xs = __fmul_rn(x, __fadd_rn(s1, s2));
  if (id < 0) return __fsub_rn(x, xs);
  const float hi = id == 0 ? 4.6364760399e-01f : id == 1 ? 7.8539812565e-01f : id == 2 ? 9.8279368877e-01f : hi3;
  const float lo = id == 0 ? 5.0121582440e-09f : id == 1 ? 3.7748947079e-08f : id == 2 ? 3.4473217170e-08f : lo3;
  const float r = __fsub_rn(hi, __fsub_rn(__fsub_rn(xs, lo), x));
  return (hx >> 31) ? -r : r;
}
__device__ __forceinline__ float d_atan2f_fdlibm(float y, float x) {
  const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
  const uint32_t hx = __float_as_uint(x), hy = __float_as_uint(y), ix = hx & 0x7fffffffu, iy = hy & 0x7fffffffu;
  if (ix > 0x7f800000u || iy > 0x7f800000u) return __fadd_rn(x, y);
  if (hx == 0x3f800000u) return d_atanf_fdlibm(y);
  const int m = (int)((hy >> 31) & 1u) | (int)((hx >> 30) & 2u);          // 2 sign(x) + sign(y)
  if (iy == 0) return m < 2 ? y : (m == 2 ? __fadd_rn(pi, tiny) : __fsub_rn(-pi, tiny));
  if (ix == 0) return (hy >> 31) ? __fsub_rn(-pi_o_2, tiny) : __fadd_rn(pi_o_2, tiny);
  if (ix == 0x7f800000u) {
    if (iy == 0x7f800000u) return m == 0 ? __fadd_rn(pi_o_4, tiny) : m == 1 ? __fsub_rn(-pi_o_4, tiny) : m == 2 ? __fadd_rn(__fmul_rn(3.0f, pi_o_4), tiny) : __fsub_rn(__fmul_rn(-3.0f, pi_o_4), tiny);
    return m == 0 ? 0.0f : m == 1 ? -0.0f : m == 2 ? __fadd_rn(pi, tiny) : __fsub_rn(-pi, tiny);
  }
  if (iy == 0x7f800000u) return (hy >> 31) ? __fsub_rn(-pi_o_2, tiny) : __fadd_rn(pi_o_2, tiny);
  const int k = ((int)iy - (int)ix) >> 23;
  float z;
  if (k > 60) z = __fadd_rn(pi_o_2, __fmul_rn(0.5f, pi_lo));
  else if ((hx >> 31) && k < -60) z = 0.0f;
  else z = d_atanf_fdlibm(fabsf(__fdiv_rn(y, x)));
  if (m == 0) return z;
  if (m == 1) return __uint_as_float(__float_as_uint(z) ^ 0x80000000u);
  if (m == 2) return __fsub_rn(pi, __fsub_rn(z, pi_lo));
  return __fsub_rn(__fsub_rn(z, pi_lo), pi);
}
__device__ __forceinline__ float d_neg_atan2f(float y, float x) { return -d_atan2f_fdlibm(y, x); }

__global__ void __launch_bounds__(SC_THREADS) k_scan_classify(const float4* __restrict__ in, int n, int n_scans, float thres,
                                                              int32_t* __restrict__ key, int32_t* __restrict__ block_hist, int nb,
                                                              ScanMeta* __restrict__ meta) {
  lm_pdl_enter();
  __shared__ int counts[SC_THREADS / 32][SC_NRING];
  __shared__ int s_first[SC_THREADS / 32], s_last[SC_THREADS / 32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int e = threadIdx.x; e < (SC_THREADS / 32) * SC_NRING; e += blockDim.x) (&counts[0][0])[e] = 0;
  int ring = 64;
  bool valid0 = false;
  if (n < 0) n = meta->n_in;
  if (i < n) {
    const float4 p = in[i];
    if (isfinite(p.x) && isfinite(p.y) && isfinite(p.z)) {
      const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(p.x, p.x), __fmul_rn(p.y, p.y)), __fmul_rn(p.z, p.z));
      if (!(r2 < __fmul_rn(thres, thres))) {         // :99
        valid0 = true;
        int r = d_ring_of(p.x, p.y, p.z, n_scans);
        if (r >= 0) ring = r;
      }
    }
  }
  // first / last point surviving the NaN and range filters (cloud [0] and [cloudSize-1], :141-143)
  int fv = valid0 ? i : 0x7fffffff, lv = valid0 ? i : -1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { fv = min(fv, __shfl_xor_sync(0xffffffffu, fv, o)); lv = max(lv, __shfl_xor_sync(0xffffffffu, lv, o)); }
  if (lane == 0) { s_first[wid] = fv; s_last[wid] = lv; }
  __syncthreads();
  const unsigned peers = __match_any_sync(0xffffffffu, ring);
  const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
  if (rank_in_warp == 0) counts[wid][ring] = __popc(peers);
  __syncthreads();
  int base = 0;
  for (int w = 0; w < wid; ++w) base += counts[w][ring];
  if (i < n) key[i] = ring < 64 ? ((ring << 16) | (base + rank_in_warp)) : -1;
  if (threadIdx.x < SC_NRING) {
    int tot = 0;
    for (int w = 0; w < SC_THREADS / 32; ++w) tot += counts[w][threadIdx.x];
    block_hist[threadIdx.x * nb + blockIdx.x] = tot;
  }
  if (threadIdx.x == 0) {
    for (int w = 1; w < SC_THREADS / 32; ++w) { s_first[0] = min(s_first[0], s_first[w]); s_last[0] = max(s_last[0], s_last[w]); }
    if (s_first[0] != 0x7fffffff) atomicMin(&meta->first_valid, s_first[0]);
    if (s_last[0] >= 0) atomicMax(&meta->last_valid, s_last[0]);
  }
}

__device__ __forceinline__ void d_start_end_ori(const float4* __restrict__ in, const ScanMeta* __restrict__ meta, float* so, float* eo) {
  const float4 a = in[meta->first_valid], b = in[meta->last_valid];
  float startOri = d_neg_atan2f(a.y, a.x);
  float endOri = (float)((double)d_neg_atan2f(b.y, b.x) + 2 * SC_PI);
  if ((double)__fsub_rn(endOri, startOri) > 3 * SC_PI) endOri = (float)((double)endOri - 2 * SC_PI);
  else if ((double)__fsub_rn(endOri, startOri) < SC_PI) endOri = (float)((double)endOri + 2 * SC_PI);
  *so = startOri; *eo = endOri;
}

__global__ void __launch_bounds__(SC_THREADS) k_scan_halfpass(const float4* __restrict__ in, int n, const int32_t* __restrict__ key,
                                                              ScanMeta* __restrict__ meta) {
  lm_pdl_enter();
  __shared__ float s_so, s_eo;
  __shared__ int s_min[SC_THREADS / 32];
  if (meta->last_valid < 0) return;
  if (n < 0) n = meta->n_in;
  if (threadIdx.x == 0) d_start_end_ori(in, meta, &s_so, &s_eo);
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int cand = 0x7fffffff;
  if (i < n && key[i] >= 0) {
    const float4 p = in[i];
    float ori = d_neg_atan2f(p.y, p.x);
    const double so = (double)s_so;
    if ((double)ori < so - SC_PI / 2) ori = (float)((double)ori + 2 * SC_PI);
    else if ((double)ori > so + SC_PI * 3 / 2) ori = (float)((double)ori - 2 * SC_PI);
    if ((double)__fsub_rn(ori, s_so) > SC_PI) cand = i;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
  if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = cand;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < SC_THREADS / 32; ++w) s_min[0] = min(s_min[0], s_min[w]);
    if (s_min[0] != 0x7fffffff) atomicMin(&meta->i_star, s_min[0]);
  }
}

__global__ void __launch_bounds__(1024) k_scan_blockscan(const int32_t* __restrict__ block_hist, int32_t* __restrict__ block_off, int nb,
                                                         ScanMeta* __restrict__ meta) {
  lm_pdl_enter();
  __shared__ int ws[33];
  __shared__ int s_carry;
  const int r = blockIdx.x;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += blockDim.x) {
    const int e = base + threadIdx.x;
    int v = e < nb ? block_hist[r * nb + e] : 0;
    int total;
    int ex = d_block_exscan(v, ws, &total);
    if (e < nb) block_off[r * nb + e] = s_carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) s_carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) meta->ring_total[r] = s_carry;
}

__global__ void __launch_bounds__(SC_THREADS) k_scan_scatter(const float4* __restrict__ in, int n, int n_scans, const int32_t* __restrict__ key,
                                                             const int32_t* __restrict__ block_off, int nb, ScanMeta* __restrict__ meta,
                                                             float4* __restrict__ full, int32_t* __restrict__ src) {
  lm_pdl_enter();
  __shared__ int s_start[SC_NRING + 1];
  __shared__ float s_so, s_eo;
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int r = 0; r < 64; ++r) { s_start[r] = acc; acc += (r < n_scans) ? meta->ring_total[r] : 0; }
    s_start[64] = acc;
    s_so = 0.f; s_eo = 1.f;
    if (meta->last_valid >= 0) d_start_end_ori(in, meta, &s_so, &s_eo);
    if (blockIdx.x == 0) {
      for (int r = 0; r <= 64; ++r) meta->ring_start[r] = s_start[r];
      meta->n_kept = acc; meta->start_ori = s_so; meta->end_ori = s_eo;
    }
  }
  __syncthreads();
  if (n < 0) n = meta->n_in;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = key[i];
  if (k < 0) return;
  const int ring = k >> 16, rank = k & 0xffff;
  const float4 p = in[i];
  float ori = d_neg_atan2f(p.y, p.x);
  const double so = (double)s_so, eo = (double)s_eo;
  if (i <= meta->i_star) {                 // halfPassed still false when this point is visited
    if ((double)ori < so - SC_PI / 2) ori = (float)((double)ori + 2 * SC_PI);
    else if ((double)ori > so + SC_PI * 3 / 2) ori = (float)((double)ori - 2 * SC_PI);
  } else {
    ori = (float)((double)ori + 2 * SC_PI);
    if ((double)ori < eo - SC_PI * 3 / 2) ori = (float)((double)ori + 2 * SC_PI);
    else if ((double)ori > eo + SC_PI / 2) ori = (float)((double)ori - 2 * SC_PI);
  }
  const float relTime = __fdiv_rn(__fsub_rn(ori, s_so), __fsub_rn(s_eo, s_so));
  const int pos = s_start[ring] + block_off[ring * nb + blockIdx.x] + rank;
  full[pos] = make_float4(p.x, p.y, p.z, (float)((double)ring + 0.1 * (double)relTime));
  src[pos] = i;
}

__global__ void __launch_bounds__(SC_THREADS) k_scan_curvature(const float4* __restrict__ full, const ScanMeta* __restrict__ meta,
                                                               float* __restrict__ curv, int32_t* __restrict__ label) {
  lm_pdl_enter();
  const int N = meta->n_kept;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  label[i] = 0;
  if (i < 5 || i >= N - 5) { curv[i] = 0.f; return; }
  float dx = full[i - 5].x, dy = full[i - 5].y, dz = full[i - 5].z;
#pragma unroll
  for (int o = -4; o <= -1; ++o) { const float4 q = full[i + o]; dx = __fadd_rn(dx, q.x); dy = __fadd_rn(dy, q.y); dz = __fadd_rn(dz, q.z); }
  { const float4 q = full[i]; dx = __fsub_rn(dx, __fmul_rn(10.0f, q.x)); dy = __fsub_rn(dy, __fmul_rn(10.0f, q.y)); dz = __fsub_rn(dz, __fmul_rn(10.0f, q.z)); }
#pragma unroll
  for (int o = 1; o <= 5; ++o) { const float4 q = full[i + o]; dx = __fadd_rn(dx, q.x); dy = __fadd_rn(dy, q.y); dz = __fadd_rn(dz, q.z); }
  curv[i] = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// ---------------------------------------------------------------------------------------------
constexpr int SCR_THREADS = 1024;          // k_scan_ring: one CTA per ring

__device__ __forceinline__ bool d_gap_exceeds(const float* xyz, int a, int b) {
  // (p[a] - p[b]) squared norm > 0.05 (double literal), fp32 left to right
  const float dx = __fsub_rn(xyz[a * 3], xyz[b * 3]), dy = __fsub_rn(xyz[a * 3 + 1], xyz[b * 3 + 1]), dz = __fsub_rn(xyz[a * 3 + 2], xyz[b * 3 + 2]);
  const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  return (double)d > 0.05;
}

// Two independent block-wide bitonic networks run stage by stage side by side (a dependent chain of 66 stages for 2048
// keys is latency-bound: the second network rides in the first one's issue slots).  Same stages as d_bitonic_regs.
template <int ITEMS, int NTHREADS>
__device__ __forceinline__ void d_bitonic_regs2(unsigned long long (&va)[ITEMS], unsigned long long (&vb)[ITEMS], int t,
                                                unsigned long long* xa, unsigned long long* xb) {
  constexpr int N = ITEMS * NTHREADS;
  constexpr int LOGN = (N <= 1) ? 0 : (31 - __builtin_clz((unsigned)N));
  static_assert((1 << LOGN) == N, "ITEMS * NTHREADS must be a power of two");
#pragma unroll
  for (int lk = 1; lk <= LOGN; ++lk) {
    const int k = 1 << lk;
#pragma unroll
    for (int lj = lk - 1; lj >= 0; --lj) {
      const int j = 1 << lj;
      if (j >= 32 * ITEMS) {
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) { xa[t * ITEMS + r] = va[r]; xb[t * ITEMS + r] = vb[r]; }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
          const int i = t * ITEMS + r;
          const unsigned long long oa = xa[i ^ j], ob = xb[i ^ j];
          const bool keep_min = ((i & j) == 0) == ((i & k) == 0);
          va[r] = keep_min ? (va[r] < oa ? va[r] : oa) : (va[r] < oa ? oa : va[r]);
          vb[r] = keep_min ? (vb[r] < ob ? vb[r] : ob) : (vb[r] < ob ? ob : vb[r]);
        }
        __syncthreads();
      } else if (j >= ITEMS) {
        const int lane_x = j / ITEMS;
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
          const int i = t * ITEMS + r;
          const unsigned long long oa = __shfl_xor_sync(0xffffffffu, va[r], lane_x), ob = __shfl_xor_sync(0xffffffffu, vb[r], lane_x);
          const bool keep_min = ((i & j) == 0) == ((i & k) == 0);
          va[r] = keep_min ? (va[r] < oa ? va[r] : oa) : (va[r] < oa ? oa : va[r]);
          vb[r] = keep_min ? (vb[r] < ob ? vb[r] : ob) : (vb[r] < ob ? ob : vb[r]);
        }
      } else {
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
          if ((r & j) == 0 && (r | j) < ITEMS) {
            const int i = t * ITEMS + r;
            d_cmpx(va[r], va[r | j], (i & k) == 0);
            d_cmpx(vb[r], vb[r | j], (i & k) == 0);
          }
        }
      }
    }
  }
}

__device__ __forceinline__ void d_named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory"); }

// d_bitonic_regs for a GROUP of NTHREADS threads of the CTA (t = index inside the group): the cross-warp stages meet at the
// group's own named barrier, so several groups sort different arrays at the same time and the rest of the CTA does not wait
template <int ITEMS, int NTHREADS>
__device__ __forceinline__ void d_bitonic_regs_group(unsigned long long (&v)[ITEMS], int t, unsigned long long* xch, int bar_id) {
  constexpr int N = ITEMS * NTHREADS;
  constexpr int LOGN = (N <= 1) ? 0 : (31 - __builtin_clz((unsigned)N));
  static_assert((1 << LOGN) == N, "ITEMS * NTHREADS must be a power of two");
#pragma unroll
  for (int lk = 1; lk <= LOGN; ++lk) {
    const int k = 1 << lk;
#pragma unroll
    for (int lj = lk - 1; lj >= 0; --lj) {
      const int j = 1 << lj;
      if (j >= 32 * ITEMS) {
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) xch[t * ITEMS + r] = v[r];
        d_named_bar(bar_id, NTHREADS);
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
          const int i = t * ITEMS + r;
          const unsigned long long o = xch[i ^ j];
          const bool keep_min = ((i & j) == 0) == ((i & k) == 0);
          v[r] = keep_min ? (v[r] < o ? v[r] : o) : (v[r] < o ? o : v[r]);
        }
        d_named_bar(bar_id, NTHREADS);
      } else if (j >= ITEMS) {
        const int lane_x = j / ITEMS;
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
          const int i = t * ITEMS + r;
          const unsigned long long o = __shfl_xor_sync(0xffffffffu, v[r], lane_x);
          const bool keep_min = ((i & j) == 0) == ((i & k) == 0);
          v[r] = keep_min ? (v[r] < o ? v[r] : o) : (v[r] < o ? o : v[r]);
        }
      } else {
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
          if ((r & j) == 0 && (r | j) < ITEMS) {
            const int i = t * ITEMS + r;
            d_cmpx(v[r], v[r | j], (i & k) == 0);
          }
        }
      }
    }
  }
}

// :284-288 one sector (<= 512 points) sorted by ONE group of 128 threads, 4 keys each: a 512-key network has 45 stages, the
// 2048-key network over all six sectors 66.  Key = curvature bits (32) | local index (12), ascending.
__device__ __forceinline__ void d_group_sort_sector(unsigned long long* out, unsigned long long* xch, const float* __restrict__ curv,
                                                    int rs, int sp, int len, int t, int bar_id) {
  unsigned long long v[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int e = t * 4 + r;
    v[r] = e < len ? (((unsigned long long)__float_as_uint(curv[sp + e]) << 12) | (unsigned long long)(sp + e - rs)) : ~0ULL;
  }
  d_bitonic_regs_group<4, 128>(v, t, xch, bar_id);
#pragma unroll
  for (int r = 0; r < 4; ++r) { const int e = t * 4 + r; if (e < len) out[e] = v[r]; }
}

// :401-405 the voxel order of the ring's points (see d_block_sort_ring) by a group of 512 threads, 4 keys each, while the
// six picking warps work
__device__ __forceinline__ void d_group_sort_voxels(unsigned long long* out, const float* xyz, int rs, int S, int n, float inv, int t, int bar_id) {
  unsigned long long v[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int e = t * 4 + r;
    unsigned long long c = ~0ULL;
    if (e < n) {
      const int li = S + e - rs;
      const int vx = min(max((int)floorf(__fmul_rn(xyz[li * 3], inv)), -32768), 32767) + 32768;
      const int vy = min(max((int)floorf(__fmul_rn(xyz[li * 3 + 1], inv)), -32768), 32767) + 32768;
      const int vz = min(max((int)floorf(__fmul_rn(xyz[li * 3 + 2], inv)), -32768), 32767) + 32768;
      c = ((unsigned long long)vz << 44) | ((unsigned long long)vy << 28) | ((unsigned long long)vx << 12) | (unsigned long long)li;
    }
    v[r] = c;
  }
  d_bitonic_regs_group<4, 512>(v, t, out, bar_id);
#pragma unroll
  for (int r = 0; r < 4; ++r) out[t * 4 + r] = v[r];
}

// The two sorts of a ring, over the same n = E - S points, as one pass:
//  A  :284-288 the six sector sorts as ONE network over sector (3 bits) | curvature bits (32) | local index (12):
//     ascending (curvature, index) inside each sector, the sectors one behind the other;
//  B  :401-405 the VoxelGrid order of the ring's less-flat cloud.  PCL sorts by idx = i0 + i1 d0 + i2 d0 d1 with
//     (i0, i1, i2) = floor(p / leaf) - minb: that is the lexicographic order of the ABSOLUTE voxel coordinates (z, y, x),
//     whatever the bounding box is, so the keys can be formed before the greedy pick has decided which points belong to
//     the cloud: (vz, vy, vx) + 2^15 in 16 bits each | local index (12).  The picked (sharp / less-sharp) points are
//     dropped from the sorted sequence afterwards.
template <int ITEMS>
__device__ __forceinline__ void d_block_sort_ring(unsigned long long* xa, unsigned long long* xb, const float* __restrict__ curv,
                                                  const float* xyz, int rs, int S, int E, float inv) {
  unsigned long long va[ITEMS], vb[ITEMS];
  const int n = E - S;
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    const int e = threadIdx.x * ITEMS + r;
    unsigned long long ca = ~0ULL, cb = ~0ULL;
    if (e < n) {
      // sector j holds the offsets [n j / 6, n (j + 1) / 6)
      int j = (int)(((long long)e * 6 + 5) / n);
      while (j > 0 && (long long)n * j / 6 > e) --j;
      while (j < 5 && (long long)n * (j + 1) / 6 <= e) ++j;
      const int li = S + e - rs;
      ca = ((unsigned long long)j << 44) | ((unsigned long long)__float_as_uint(curv[S + e]) << 12) | (unsigned long long)li;
      const int vx = min(max((int)floorf(__fmul_rn(xyz[li * 3], inv)), -32768), 32767) + 32768;
      const int vy = min(max((int)floorf(__fmul_rn(xyz[li * 3 + 1], inv)), -32768), 32767) + 32768;
      const int vz = min(max((int)floorf(__fmul_rn(xyz[li * 3 + 2], inv)), -32768), 32767) + 32768;
      cb = ((unsigned long long)vz << 44) | ((unsigned long long)vy << 28) | ((unsigned long long)vx << 12) | (unsigned long long)li;
    }
    va[r] = ca; vb[r] = cb;
  }
  d_bitonic_regs2<ITEMS, SCR_THREADS>(va, vb, threadIdx.x, xa, xb);
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) { xa[threadIdx.x * ITEMS + r] = va[r]; xb[threadIdx.x * ITEMS + r] = vb[r]; }
  __syncthreads();
}

// dynamic shared memory layout of k_scan_ring
constexpr int SCR_SEC_BYTES = SC_RING_MAX * 8;                        // sector-sorted keys; later the compacted voxel-sorted keys
constexpr int SCR_VOX_BYTES = SC_RING_MAX * 8;                        // voxel-sorted keys
constexpr int SCR_XYZ_BYTES = SC_RING_MAX * 12;                       // 49152
constexpr int SCR_INT_BYTES = SC_RING_MAX * 4;                        // intensity
constexpr int SCR_PICK_BYTES = SC_RING_MAX + 32;                      // picked flags
constexpr int SCR_LABEL_BYTES = SC_RING_MAX;                          // labels (int8)
constexpr int SCR_GAP_BYTES = SC_RING_MAX;                            // gap[i] = |p[i] - p[i-1]|^2 > 0.05 (the +-5 suppression test), precomputed in parallel
constexpr int SCR_TOTAL = SCR_SEC_BYTES + SCR_VOX_BYTES + SCR_XYZ_BYTES + SCR_INT_BYTES + SCR_PICK_BYTES + SCR_LABEL_BYTES + 256 + SCR_GAP_BYTES;

// one greedy pick, warp-wide (:297-342 / :352-388).  Every lane holds one entry of the current 32-entry window of the
// sorted sector: its local index li, whether it is still pickable (`alive`, a register) and the ten gap bits around it.
// The pick is the first alive lane; its index and gap bits are broadcast, every lane clears `alive` if it falls in the
// suppressed span (the reference's two break-on-gap walks = first set gap bit), lanes 1..10 mark the span in shared
// memory for the windows that follow.  Marks are only written inside the sector [lo, hi]: what falls behind it is
// returned as a bit mask (bit b = local index hi + 1 + b), what falls before it cannot matter any more.
__device__ __forceinline__ int d_pick_one(unsigned m, int li, uint32_t gbits, bool& alive, unsigned char* picked, int lane, bool suppress,
                                          int lo, int hi, uint32_t& spill) {
  const int src = __ffs(m) - 1;
  const uint32_t pk = __shfl_sync(0xffffffffu, (uint32_t)li | (gbits << 12), src);
  const int pli = (int)(pk & 0xfffu);
  if (!suppress) { if (lane == src) alive = false; return pli; }
  const uint32_t fm = (pk >> 12) & 31u, bm = (pk >> 17) & 31u;
  const int nf = fm ? __ffs(fm) - 1 : 5, nb = bm ? __ffs(bm) - 1 : 5;
  if (li >= pli - nb && li <= pli + nf) alive = false;
  if (lane == 0) picked[pli] = 1;
  if (lane >= 1 && lane <= nf && pli + lane <= hi) picked[pli + lane] = 1;
  if (lane >= 6 && lane - 5 <= nb && pli - (lane - 5) >= lo) picked[pli - (lane - 5)] = 1;
  const int over = pli + nf - hi;
  if (over > 0) spill |= (1u << over) - 1u;
  return pli;
}

// :291-390 the greedy pick of ONE sector by one warp; `incoming` = marks the previous sector left on this sector's first
// five points.  Returns the marks this sector leaves on the next one.  Same picks in the same order as the reference's walk.
__device__ __noinline__ uint32_t d_pick_sector(int j, uint32_t incoming, bool redo, const unsigned long long* sec, int S, int n, int rs, int r,
                                               unsigned char* picked, const unsigned char* gap, signed char* label,
                                               int32_t* __restrict__ pick_idx, int32_t* __restrict__ pick_cnt, int lane) {
  const int sp = S + n * j / 6, ep = S + n * (j + 1) / 6 - 1;
  const int len = ep - sp + 1;
  const int lo = sp - rs, hi = ep - rs;
  const unsigned long long* sk = sec + (sp - S);             // the sectors lie one behind the other in the sorted array
  int* out = pick_idx + (r * 6 + j) * SC_PICK_STRIDE;
  if (redo) {
    for (int i = lo + lane; i <= hi; i += 32) { picked[i] = 0; label[i] = 0; }
    __syncwarp();
  }
  if (lane < 5 && ((incoming >> lane) & 1u) && lo + lane <= hi) picked[lo + lane] = 1;
  uint32_t spill = 0;
  int n_sharp = 0, n_ls = 0, n_flat = 0;
  int largest = 0;
  bool done = false;
  for (int k_hi = len - 1; k_hi >= 0 && !done; k_hi -= 32) {          // descending curvature
    __syncwarp();
    const int k = k_hi - lane;
    const unsigned long long c = k >= 0 ? sk[k] : 0ull;
    const bool qual = k >= 0 && (double)__uint_as_float((uint32_t)(c >> 12)) > 0.1;
    const int li = (int)(c & 0xfffull);
    const unsigned qm = __ballot_sync(0xffffffffu, qual);
    const int nvalid = qm == 0xffffffffu ? 32 : __ffs(~qm) - 1;      // sorted: nothing behind the first non-qualifying entry qualifies
    bool alive = lane < nvalid && !picked[li];
    uint32_t gbits = 0;
    if (lane < nvalid) {
#pragma unroll
      for (int l = 1; l <= 5; ++l) gbits |= (uint32_t)(gap[li + l] != 0) << (l - 1);            // :319-330
#pragma unroll
      for (int l = 1; l <= 5; ++l) gbits |= (uint32_t)(gap[li - l + 1] != 0) << (4 + l);        // :331-342
    }
    __syncwarp();                                                     // the picks below write flags other lanes have just read
    for (;;) {
      const unsigned m = __ballot_sync(0xffffffffu, alive);
      if (!m) break;
      largest++;
      if (largest > 20) { done = true; break; }
      const int pli = d_pick_one(m, li, gbits, alive, picked, lane, true, lo, hi, spill);
      if (lane == 0) {
        const int ind = pli + rs;
        if (largest <= 2) { label[pli] = 2; out[n_sharp] = ind; out[2 + n_ls] = ind; }
        else { label[pli] = 1; out[2 + n_ls] = ind; }
      }
      if (largest <= 2) n_sharp++;
      n_ls++;
    }
    if (nvalid < 32) break;
  }
  int smallest = 0;
  done = false;
  for (int k_lo = 0; k_lo < len && !done; k_lo += 32) {               // ascending curvature
    __syncwarp();
    const int k = k_lo + lane;
    const unsigned long long c = k < len ? sk[k] : 0ull;
    const bool qual = k < len && (double)__uint_as_float((uint32_t)(c >> 12)) < 0.1;
    const int li = (int)(c & 0xfffull);
    const unsigned qm = __ballot_sync(0xffffffffu, qual);
    const int nvalid = qm == 0xffffffffu ? 32 : __ffs(~qm) - 1;
    bool alive = lane < nvalid && !picked[li];
    uint32_t gbits = 0;
    if (lane < nvalid) {
#pragma unroll
      for (int l = 1; l <= 5; ++l) gbits |= (uint32_t)(gap[li + l] != 0) << (l - 1);
#pragma unroll
      for (int l = 1; l <= 5; ++l) gbits |= (uint32_t)(gap[li - l + 1] != 0) << (4 + l);
    }
    __syncwarp();
    for (;;) {
      const unsigned m = __ballot_sync(0xffffffffu, alive);
      if (!m) break;
      smallest++;
      const bool last = smallest >= 4;                                // :359-362: the 4th flat point is neither marked nor suppressing
      const int pli = d_pick_one(m, li, gbits, alive, picked, lane, !last, lo, hi, spill);
      if (lane == 0) { label[pli] = -1; out[22 + n_flat] = pli + rs; }
      n_flat++;
      if (last) { done = true; break; }
    }
    if (nvalid < 32) break;
  }
  __syncwarp();
  if (lane == 0) { pick_cnt[r * 18 + j * 3 + 0] = n_sharp; pick_cnt[r * 18 + j * 3 + 1] = n_ls; pick_cnt[r * 18 + j * 3 + 2] = n_flat; }
  return spill;
}

__global__ void __launch_bounds__(SCR_THREADS, 1) k_scan_ring(const float4* __restrict__ full, const float* __restrict__ curv,
                                                             ScanMeta* __restrict__ meta, int32_t* __restrict__ label_out,
                                                             int32_t* __restrict__ pick_idx, int32_t* __restrict__ pick_cnt,
                                                             float4* __restrict__ lf_tmp, int32_t* __restrict__ lf_cnt,
                                                             unsigned long long* __restrict__ stamps) {
  lm_pdl_enter();
#define SCR_STAMP(slot) do { if (stamps != nullptr && threadIdx.x == 0 && blockIdx.x == 32) stamps[200 + (slot)] = d_globaltimer(); } while (0)
  SCR_STAMP(0);
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned long long* sec = reinterpret_cast<unsigned long long*>(smem);
  unsigned long long* vox = reinterpret_cast<unsigned long long*>(smem + SCR_SEC_BYTES);
  float* xyz = reinterpret_cast<float*>(smem + SCR_SEC_BYTES + SCR_VOX_BYTES);
  float* inten = reinterpret_cast<float*>(smem + SCR_SEC_BYTES + SCR_VOX_BYTES + SCR_XYZ_BYTES);
  unsigned char* picked = smem + SCR_SEC_BYTES + SCR_VOX_BYTES + SCR_XYZ_BYTES + SCR_INT_BYTES;
  signed char* label = reinterpret_cast<signed char*>(picked + SCR_PICK_BYTES);
  int* ws = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(label) + SCR_LABEL_BYTES);     // [64]
  unsigned char* gap = reinterpret_cast<unsigned char*>(ws) + 256;                                  // [SC_RING_MAX]
  __shared__ int s_vg[8];
  __shared__ float s_mn[3][SCR_THREADS / 32], s_mx[3][SCR_THREADS / 32];

  const int r = blockIdx.x;
  const int rs = meta->ring_start[r], re = meta->ring_start[r + 1];
  const int L = re - rs;
  const int S = rs + 5, E = re - 6;
  if (threadIdx.x < 18) pick_cnt[r * 18 + threadIdx.x] = 0;
  if (threadIdx.x == 0) lf_cnt[r] = 0;
  if (E - S < 6) return;                                  // :279
  if (L > SC_RING_MAX) { if (threadIdx.x == 0) atomicOr(&meta->fault, 1u); return; }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int n = E - S;
  const float inv = 1.0f / 0.2f;

  for (int i0 = threadIdx.x; i0 < L; i0 += 4 * SCR_THREADS) {      // four loads in flight per thread, then the shared-memory stores
    float4 p[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int i = i0 + u * SCR_THREADS; if (i < L) p[u] = full[rs + i]; }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * SCR_THREADS;
      if (i < L) { xyz[i * 3] = p[u].x; xyz[i * 3 + 1] = p[u].y; xyz[i * 3 + 2] = p[u].z; inten[i] = p[u].w; picked[i] = 0; label[i] = 0; }
    }
  }
  __syncthreads();
  SCR_STAMP(1);
  // the greedy pick below tests |p[i] - p[i-1]|^2 > 0.05 up to ten times per pick (:319-342, 365-388): the test is a pure
  // function of two neighbours, so every thread evaluates its share once here ((a - b)^2 == (b - a)^2 exactly, one array
  // serves both walking directions)
  for (int i = threadIdx.x; i < L; i += blockDim.x) gap[i] = (i > 0 && d_gap_exceeds(xyz, i, i - 1)) ? 1 : 0;
  // Usual ring (<= 2048 points, sectors <= 512): each sector is sorted by its own group of four warps, then the six picking
  // warps and a 512-thread group that sorts the voxel keys run side by side.  Larger rings: both sorts as two block-wide
  // networks first (d_block_sort_ring), then the pick.
  const bool fast = n <= 2048 && (n + 5) / 6 <= 512;
  __shared__ uint32_t s_spill[6];
  if (fast) {
    __syncthreads();                                   // gap[] and xyz[] complete
    if (wid < 24) {
      const int g = wid >> 2;
      const int sp = S + n * g / 6, ep = S + n * (g + 1) / 6 - 1;
      d_group_sort_sector(sec + (sp - S), vox + g * 512, curv, rs, sp, ep - sp + 1, threadIdx.x - g * 128, 1 + g);
    }
    __syncthreads();
  } else {
    if (n <= 2 * SCR_THREADS) d_block_sort_ring<2>(sec, vox, curv, xyz, rs, S, E, inv);
    else d_block_sort_ring<4>(sec, vox, curv, xyz, rs, S, E, inv);
  }
  SCR_STAMP(2);

  // :291-390 greedy picking.  The reference walks the six sectors one after the other, and a pick suppresses up to five
  // neighbours on either side -- possibly across the sector border.  Marks reaching BACK into a finished sector change
  // nothing; marks reaching FORWARD change the next sector only if they hit a point that sector would have picked.
  // So six warps pick their sectors at once, each assuming no incoming marks, and a short ordered check follows: if the
  // (final) forward marks of sector j - 1 hit a point that sector j picked, sector j is picked again with those marks (about
  // one border in ten), which may in turn change what it passes on.  Same picks as the sequential walk.
  if (wid < 6) {
    const uint32_t sp_ = d_pick_sector(wid, 0u, false, sec, S, n, rs, r, picked, gap, label, pick_idx, pick_cnt, lane);
    if (lane == 0) s_spill[wid] = sp_;
    d_named_bar(8, 192);
    for (int j = 1; j < 6; ++j) {
      if (wid == j) {
        const uint32_t inc = s_spill[j - 1];
        const int lo = S + n * j / 6 - rs, hi = S + n * (j + 1) / 6 - 1 - rs;
        const bool hit = lane < 5 && ((inc >> lane) & 1u) && lo + lane <= hi && label[lo + lane] != 0;
        if (__any_sync(0xffffffffu, hit)) {
          const uint32_t sp2 = d_pick_sector(j, inc, true, sec, S, n, rs, r, picked, gap, label, pick_idx, pick_cnt, lane);
          if (lane == 0) s_spill[j] = sp2;
        }
      }
      d_named_bar(8, 192);
    }
  } else if (fast && wid >= 8 && wid < 24) {
    d_group_sort_voxels(vox, xyz, rs, S, n, inv, threadIdx.x - 256, 7);
  }
  __syncthreads();
  SCR_STAMP(3);
  for (int i = threadIdx.x; i < L; i += blockDim.x) label_out[rs + i] = (int)label[i];

  // :392-398 less-flat = every point of [S, E) with label <= 0; :401-405 VoxelGrid(0.2) of them (PCL arithmetic, see
  // voxel.cu).  The voxel-sorted sequence of all n points exists already: drop the picked ones (stable compaction).
  unsigned long long* lfs = sec;                            // the sector-sorted keys are no longer needed
  int n_lf = 0;
  float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, mx[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
  for (int base = 0; base < n; base += blockDim.x) {
    const int e = base + threadIdx.x;
    unsigned long long c = 0; int flag = 0;
    if (e < n) { c = vox[e]; flag = label[(int)(c & 0xfffull)] <= 0 ? 1 : 0; }
    int total;
    const int ex = d_block_exscan(flag, ws, &total);
    if (flag) {
      lfs[n_lf + ex] = c;
      const int li = (int)(c & 0xfffull);
#pragma unroll
      for (int d = 0; d < 3; ++d) { mn[d] = fminf(mn[d], xyz[li * 3 + d]); mx[d] = fmaxf(mx[d], xyz[li * 3 + d]); }
    }
    n_lf += total;
  }
  __syncthreads();
  SCR_STAMP(4);
  if (n_lf == 0) return;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o)); mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o)); }
    if (lane == 0) { s_mn[d][wid] = mn[d]; s_mx[d][wid] = mx[d]; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {                          // PCL's guard: (dx + 1)(dy + 1)(dz + 1) of the bounding box must fit an int
    long long dd[3];
    for (int d = 0; d < 3; ++d) {
      float a = s_mn[d][0], b = s_mx[d][0];
      for (int w = 1; w < SCR_THREADS / 32; ++w) { a = fminf(a, s_mn[d][w]); b = fmaxf(b, s_mx[d][w]); }
      dd[d] = (long long)(__fmul_rn(__fsub_rn(b, a), inv)) + 1;
    }
    s_vg[5] = (dd[0] * dd[1] * dd[2] > (long long)INT32_MAX) ? 1 : 0;
  }
  __syncthreads();
  SCR_STAMP(5);
  float4* outp = lf_tmp + rs;
  if (s_vg[5]) {                                   // PCL: leaf too small -> cloud returned unchanged (index order)
    int n_out = 0;
    for (int base = S - rs; base < E - rs; base += blockDim.x) {
      const int i = base + threadIdx.x;
      const int flag = (i < E - rs && label[i] <= 0) ? 1 : 0;
      int total;
      const int ex = d_block_exscan(flag, ws, &total);
      if (flag) outp[n_out + ex] = make_float4(xyz[i * 3], xyz[i * 3 + 1], xyz[i * 3 + 2], inten[i]);
      n_out += total;
    }
    if (threadIdx.x == 0) lf_cnt[r] = n_out;
    return;
  }
  SCR_STAMP(6);
  int n_out = 0;
  for (int base = 0; base < n_lf; base += blockDim.x) {
    const int e = base + threadIdx.x;
    int head = 0; unsigned long long me = 0;
    if (e < n_lf) { me = lfs[e]; head = (e == 0) || ((me >> 12) != (lfs[e - 1] >> 12)); }
    int total;
    const int ex = d_block_exscan(head, ws, &total);
    if (head) {
      const unsigned long long key = me >> 12;
      float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f; int cnt = 0;
      for (int m = e; m < n_lf; ++m) {
        const unsigned long long c = lfs[m];
        if ((c >> 12) != key) break;
        const int li = (int)(c & 0xfffull);
        sx = __fadd_rn(sx, xyz[li * 3]); sy = __fadd_rn(sy, xyz[li * 3 + 1]); sz = __fadd_rn(sz, xyz[li * 3 + 2]); si = __fadd_rn(si, inten[li]);
        ++cnt;
      }
      const float c = (float)cnt;
      outp[n_out + ex] = make_float4(__fdiv_rn(sx, c), __fdiv_rn(sy, c), __fdiv_rn(sz, c), __fdiv_rn(si, c));
    }
    n_out += total;
  }
  if (threadIdx.x == 0) lf_cnt[r] = n_out;
  SCR_STAMP(7);
#undef SCR_STAMP
}

__global__ void __launch_bounds__(SC_THREADS) k_scan_compact(const float4* __restrict__ full, ScanMeta* __restrict__ meta, int n_scans,
                                                             const int32_t* __restrict__ pick_idx, const int32_t* __restrict__ pick_cnt,
                                                             const float4* __restrict__ lf_tmp, const int32_t* __restrict__ lf_cnt,
                                                             float4* __restrict__ o_sharp, float4* __restrict__ o_ls,
                                                             float4* __restrict__ o_flat, float4* __restrict__ o_lf) {
  lm_pdl_enter();
  const int b = blockIdx.x;
  if (b < n_scans) {
    int off = 0;
    for (int r = 0; r < b; ++r) off += lf_cnt[r];
    const int cnt = lf_cnt[b];
    const float4* src = lf_tmp + meta->ring_start[b];
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) o_lf[off + i] = src[i];
    if (b == n_scans - 1 && threadIdx.x == 0) meta->out_n[3] = off + cnt;
    return;
  }
  // picks: (ring, sector) major, pick order inside.  Every pick CTA scans the three counts of all (ring, sector) entries
  // (a few hundred integers) to get the output offsets, then gathers ITS share of the entries, one warp per entry and one
  // lane per picked point: the <= 26 gathers of an entry are in flight together instead of one behind the other.
  __shared__ int ws[33];
  __shared__ int s_off[3][64 * 6 + 1];
  const int ne = n_scans * 6;
  int a0 = 0, a1 = 0, a2 = 0;
  for (int e0 = 0; e0 < ne; e0 += blockDim.x) {
    const int e = e0 + threadIdx.x;
    int c0 = 0, c1 = 0, c2 = 0;
    if (e < ne) { c0 = pick_cnt[e * 3]; c1 = pick_cnt[e * 3 + 1]; c2 = pick_cnt[e * 3 + 2]; }
    int t0, t1, t2;
    const int x0 = d_block_exscan(c0, ws, &t0), x1 = d_block_exscan(c1, ws, &t1), x2 = d_block_exscan(c2, ws, &t2);
    if (e < ne) { s_off[0][e] = a0 + x0; s_off[1][e] = a1 + x1; s_off[2][e] = a2 + x2; }
    a0 += t0; a1 += t1; a2 += t2;
  }
  __syncthreads();
  const int pc = b - n_scans, npc = gridDim.x - n_scans;               // this pick CTA, number of pick CTAs
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int e = pc * nw + wid; e < ne; e += npc * nw) {
    const int c0 = pick_cnt[e * 3], c1 = pick_cnt[e * 3 + 1], c2 = pick_cnt[e * 3 + 2];
    const int* idx = pick_idx + e * SC_PICK_STRIDE;
    if (lane < c0) o_sharp[s_off[0][e] + lane] = full[idx[lane]];
    if (lane < c1) o_ls[s_off[1][e] + lane] = full[idx[2 + lane]];
    if (lane < c2) o_flat[s_off[2][e] + lane] = full[idx[22 + lane]];
  }
  if (pc != 0) return;
  if (threadIdx.x == 0) { meta->out_n[0] = a0; meta->out_n[1] = a1; meta->out_n[2] = a2; }
}

__global__ void k_scan_meta_init(ScanMeta* meta) {
  lm_pdl_enter();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  meta->first_valid = 0x7fffffff; meta->last_valid = -1; meta->i_star = 0x7fffffff; meta->n_kept = 0;
  for (int r = 0; r < SC_NRING + 3; ++r) { meta->ring_total[r] = 0; meta->ring_start[r] = 0; }
  for (int k = 0; k < 4; ++k) meta->out_n[k] = 0;
  meta->start_ori = 0.f; meta->end_ori = 0.f; meta->fault = 0;
}

// ---------------------------------------------------------------------------------------------
static int scan_state(lmono_ctx* ctx, ScanState** out) {
  if (ctx->scan_state) { *out = (ScanState*)ctx->scan_state; return LMONO_OK; }
  ScanState* s = (ScanState*)calloc(1, sizeof(ScanState));
  s->cap = ctx->max_sweep; s->nb_max = lm_div_up(s->cap, SC_THREADS);
  const size_t n = (size_t)s->cap;
  LM_CUDA(cudaMalloc((void**)&s->d_in, n * sizeof(float4)));
  LM_CUDA(cudaMalloc((void**)&s->d_key, n * sizeof(int32_t)));
  LM_CUDA(cudaMalloc((void**)&s->d_block_hist, (size_t)SC_NRING * s->nb_max * sizeof(int32_t)));
  LM_CUDA(cudaMalloc((void**)&s->d_block_off, (size_t)SC_NRING * s->nb_max * sizeof(int32_t)));
  LM_CUDA(cudaMalloc((void**)&s->d_meta, sizeof(ScanMeta)));
  LM_CUDA(cudaMallocHost((void**)&s->h_meta, sizeof(ScanMeta)));
  LM_CUDA(cudaMalloc((void**)&s->d_full, n * sizeof(float4)));
  LM_CUDA(cudaMalloc((void**)&s->d_src, n * sizeof(int32_t)));
  LM_CUDA(cudaMalloc((void**)&s->d_curv, n * sizeof(float)));
  LM_CUDA(cudaMalloc((void**)&s->d_label, n * sizeof(int32_t)));
  LM_CUDA(cudaMalloc((void**)&s->d_pick_idx, (size_t)64 * 6 * SC_PICK_STRIDE * sizeof(int32_t)));
  LM_CUDA(cudaMalloc((void**)&s->d_pick_cnt, (size_t)64 * 18 * sizeof(int32_t)));
  LM_CUDA(cudaMalloc((void**)&s->d_lf_tmp, n * sizeof(float4)));
  LM_CUDA(cudaMalloc((void**)&s->d_lf_cnt, 64 * sizeof(int32_t)));
  LM_CUDA(cudaMalloc((void**)&s->d_out[0], (size_t)64 * 6 * 2 * sizeof(float4)));
  LM_CUDA(cudaMalloc((void**)&s->d_out[1], (size_t)64 * 6 * 20 * sizeof(float4)));
  LM_CUDA(cudaMalloc((void**)&s->d_out[2], (size_t)64 * 6 * 4 * sizeof(float4)));
  LM_CUDA(cudaMalloc((void**)&s->d_out[3], n * sizeof(float4)));
  LM_CUDA(cudaFuncSetAttribute(k_scan_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, SCR_TOTAL));
  ctx->scan_state = s;
  *out = s;
  return LMONO_OK;
}

void lm_scan_free(lmono_ctx* ctx) {
  ScanState* s = (ScanState*)ctx->scan_state;
  if (!s) return;
  cudaFree(s->d_in); cudaFree(s->d_key); cudaFree(s->d_block_hist); cudaFree(s->d_block_off); cudaFree(s->d_meta); cudaFreeHost(s->h_meta);
  cudaFree(s->d_full); cudaFree(s->d_src); cudaFree(s->d_curv); cudaFree(s->d_label); cudaFree(s->d_pick_idx); cudaFree(s->d_pick_cnt);
  cudaFree(s->d_lf_tmp); cudaFree(s->d_lf_cnt);
  for (int k = 0; k < 4; ++k) cudaFree(s->d_out[k]);
  free(s); ctx->scan_state = nullptr;
}

// enqueue the whole stage on a device-resident float4 input; results stay on the device
__global__ void k_scan_set_n(ScanMeta* meta, int n) {
  lm_pdl_enter(); if (threadIdx.x == 0 && blockIdx.x == 0) meta->n_in = n; }
int lm_scan_set_n(lmono_ctx* ctx, int n) {
  ScanState* s; int rc = scan_state(ctx, &s); if (rc) return rc;
  LM_LAUNCH_PDL(k_scan_set_n, 1, 32, 0, s->d_meta, n);
  LM_LAUNCH_CHECK();
  return LMONO_OK;
}
int lm_scan_input_buffer(lmono_ctx* ctx, float4** d_in) {
  ScanState* s; int rc = scan_state(ctx, &s); if (rc) return rc;
  *d_in = s->d_in;
  return LMONO_OK;
}

// n_on_device: `n` is only an upper bound that sizes the grids; the kernels read the sweep size from ScanMeta::n_in
// (lm_scan_set_n), so the launch sequence carries no per-sweep argument and can be replayed as a CUDA graph
int lm_scan_enqueue(lmono_ctx* ctx, const float4* d_in, int n, bool n_on_device) {
  ScanState* s; int rc = scan_state(ctx, &s); if (rc) return rc;
  const int n_scans = ctx->prm.scan_line;
  const int nb = lm_div_up(n > 0 ? n : 1, SC_THREADS);
  if (n_on_device) n = -1;
  LM_LAUNCH_PDL(k_scan_meta_init, 1, 32, 0, s->d_meta); LM_LAUNCH_CHECK();
  LM_LAUNCH_PDL(k_scan_classify, nb, SC_THREADS, 0, d_in, n, n_scans, ctx->prm.minimum_range, s->d_key, s->d_block_hist, nb, s->d_meta); LM_LAUNCH_CHECK();
  LM_LAUNCH_PDL(k_scan_halfpass, nb, SC_THREADS, 0, d_in, n, s->d_key, s->d_meta); LM_LAUNCH_CHECK();
  LM_LAUNCH_PDL(k_scan_blockscan, 64, 1024, 0, s->d_block_hist, s->d_block_off, nb, s->d_meta); LM_LAUNCH_CHECK();
  LM_LAUNCH_PDL(k_scan_scatter, nb, SC_THREADS, 0, d_in, n, n_scans, s->d_key, s->d_block_off, nb, s->d_meta, s->d_full, s->d_src); LM_LAUNCH_CHECK();
  LM_LAUNCH_PDL(k_scan_curvature, nb, SC_THREADS, 0, s->d_full, s->d_meta, s->d_curv, s->d_label); LM_LAUNCH_CHECK();
  LM_LAUNCH_PDL(k_scan_ring, n_scans, SCR_THREADS, SCR_TOTAL, s->d_full, s->d_curv, s->d_meta, s->d_label, s->d_pick_idx, s->d_pick_cnt, s->d_lf_tmp, s->d_lf_cnt, ctx->d_stamps); LM_LAUNCH_CHECK();
  LM_LAUNCH_PDL(k_scan_compact, n_scans + 16, SC_THREADS, 0, s->d_full, s->d_meta, n_scans, s->d_pick_idx, s->d_pick_cnt, s->d_lf_tmp, s->d_lf_cnt,
                                                              s->d_out[0], s->d_out[1], s->d_out[2], s->d_out[3]); LM_LAUNCH_CHECK();
  return LMONO_OK;
}

// device-side views of the stage outputs (for the fused pipeline and the odometry stage)
int lm_scan_outputs(lmono_ctx* ctx, const float4** full, const float4** sharp, const float4** less_sharp, const float4** flat,
                    const float4** less_flat, const int32_t** counts /* device: n_kept at [0], out_n[4] at [1..4] */) {
  ScanState* s = (ScanState*)ctx->scan_state;
  if (!s) return LMONO_E_STATE;
  if (full) *full = s->d_full;
  if (sharp) *sharp = s->d_out[0];
  if (less_sharp) *less_sharp = s->d_out[1];
  if (flat) *flat = s->d_out[2];
  if (less_flat) *less_flat = s->d_out[3];
  if (counts) *counts = &s->d_meta->n_kept;
  return LMONO_OK;
}

// fused sweep (mapping.cu: lmono_sweep_step): upload into the stage's input buffer; read the counts + report back
int lm_scan_upload(lmono_ctx* ctx, lmono_cloud_view raw, const float4** d_in) {
  ScanState* s; int rc = scan_state(ctx, &s); if (rc) return rc;
  LM_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  if ((rc = lm_upload_cloud(ctx, raw, ctx->d_raw[2], s->d_in, nullptr))) return rc;
  *d_in = s->d_in;
  return LMONO_OK;
}
static void scan_fill_report(lmono_ctx* ctx, const ScanMeta* m, int n_in, lmono_scan_report* report) {
  memset(report, 0, sizeof(*report));
  report->n_in = n_in; report->n_kept = m->n_kept;
  report->n_sharp = m->out_n[0]; report->n_less_sharp = m->out_n[1]; report->n_flat = m->out_n[2]; report->n_less_flat = m->out_n[3];
  for (int r = 0; r < 64; ++r) {
    report->ring_start[r] = r < ctx->prm.scan_line ? m->ring_start[r] + 5 : 0;
    report->ring_end[r] = r < ctx->prm.scan_line ? m->ring_start[r + 1] - 6 : 0;
  }
  report->start_ori = m->start_ori; report->end_ori = m->end_ori;
}
// sync-free sweep: enqueue the read-back of the stage's meta record; lm_scan_deliver fills the report once the stream
// (or an event behind this copy) has been waited for
int lm_scan_readback(lmono_ctx* ctx) {
  ScanState* s; int rc = scan_state(ctx, &s); if (rc) return rc;
  LM_CUDA(cudaMemcpyAsync(s->h_meta, s->d_meta, sizeof(ScanMeta), cudaMemcpyDeviceToHost, ctx->stream));
  return LMONO_OK;
}
int lm_scan_deliver(lmono_ctx* ctx, int n_in, lmono_scan_report* report) {
  ScanState* s = (ScanState*)ctx->scan_state;
  if (!s) return LMONO_E_STATE;
  if (report) scan_fill_report(ctx, s->h_meta, n_in, report);
  if (s->h_meta->fault) { fprintf(stderr, "[lmono_b200] scan_register fault bits 0x%x (ring or sector larger than the kernel limit)\n", s->h_meta->fault); return LMONO_E_DEVICE; }
  return LMONO_OK;
}
int lm_scan_fetch(lmono_ctx* ctx, int n_in, int32_t counts[5], lmono_scan_report* report) {
  ScanState* s; int rc = scan_state(ctx, &s); if (rc) return rc;
  LM_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  LM_CUDA(cudaMemcpyAsync(s->h_meta, s->d_meta, sizeof(ScanMeta), cudaMemcpyDeviceToHost, ctx->stream));
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  const ScanMeta* m = s->h_meta;
  counts[0] = m->n_kept; for (int k = 0; k < 4; ++k) counts[1 + k] = m->out_n[k];
  if (report) { scan_fill_report(ctx, m, n_in, report); float ms = 0.f; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); report->ms_gpu = ms; }
  if (m->fault) { fprintf(stderr, "[lmono_b200] scan_register fault bits 0x%x (ring or sector larger than the kernel limit)\n", m->fault); return LMONO_E_DEVICE; }
  return LMONO_OK;
}

extern "C" int lmono_scan_register(lmono_ctx* ctx, lmono_cloud_view raw, lmono_cloud_out* full, lmono_cloud_out* sharp,
                                   lmono_cloud_out* less_sharp, lmono_cloud_out* flat, lmono_cloud_out* less_flat,
                                   int32_t* labels, lmono_scan_report* report) {
  if (!ctx) return LMONO_E_ARG;
  if (raw.n > ctx->max_sweep) return LMONO_E_CAPACITY;
  ScanState* s; int rc = scan_state(ctx, &s); if (rc) return rc;
  LM_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  if ((rc = lm_upload_cloud(ctx, raw, ctx->d_raw[2], s->d_in, nullptr))) return rc;
  LM_CUDA(cudaEventRecord(ctx->ev_k0, ctx->stream));
  lm_kmark(ctx, "begin", 0);
  if ((rc = lm_scan_enqueue(ctx, s->d_in, raw.n, false))) return rc;
  LM_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  LM_CUDA(cudaMemcpyAsync(s->h_meta, s->d_meta, sizeof(ScanMeta), cudaMemcpyDeviceToHost, ctx->stream));
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  const ScanMeta* m = s->h_meta;
  if (report) {
    memset(report, 0, sizeof(*report));
    report->n_in = raw.n; report->n_kept = m->n_kept;
    report->n_sharp = m->out_n[0]; report->n_less_sharp = m->out_n[1]; report->n_flat = m->out_n[2]; report->n_less_flat = m->out_n[3];
    for (int r = 0; r < 64; ++r) {
      report->ring_start[r] = r < ctx->prm.scan_line ? m->ring_start[r] + 5 : 0;
      report->ring_end[r] = r < ctx->prm.scan_line ? m->ring_start[r + 1] - 6 : 0;
    }
    report->start_ori = m->start_ori; report->end_ori = m->end_ori;
    float ms = 0.f; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); report->ms_gpu = ms;
  }
  if (m->fault) { fprintf(stderr, "[lmono_b200] scan_register fault bits 0x%x (ring or sector larger than the kernel limit)\n", m->fault); return LMONO_E_DEVICE; }
  int worst = LMONO_OK;
  lmono_cloud_out* outs[5] = { full, sharp, less_sharp, flat, less_flat };
  const float4* srcs[5] = { s->d_full, s->d_out[0], s->d_out[1], s->d_out[2], s->d_out[3] };
  const int counts[5] = { m->n_kept, m->out_n[0], m->out_n[1], m->out_n[2], m->out_n[3] };
  for (int k = 0; k < 5; ++k) {
    if (!outs[k]) continue;
    int r2 = lm_download_cloud(ctx, srcs[k], counts[k], outs[k]);
    if (r2 && !worst) worst = r2;
  }
  if (labels && m->n_kept > 0) {
    LM_CUDA(cudaMemcpyAsync(labels, s->d_label, (size_t)m->n_kept * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    LM_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return worst;
}

// test hook: curvature and source indices of the last sweep
extern "C" int lmono_scan_debug(lmono_ctx* ctx, float* curvature, int32_t* src_index, int32_t n) {
  ScanState* s = ctx ? (ScanState*)ctx->scan_state : nullptr;
  if (!s) return LMONO_E_STATE;
  if (n <= 0) return LMONO_OK;
  if (curvature) LM_CUDA(cudaMemcpyAsync(curvature, s->d_curv, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  if (src_index) LM_CUDA(cudaMemcpyAsync(src_index, s->d_src, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  return LMONO_OK;
}
