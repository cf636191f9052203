// laserOdometry on device: replaces the main-loop body of Aloam/src/laserOdometry.cpp:265-568.
//
//   k_odom_nn1   :302,390   TransformToStart + exact nearest neighbour in the previous sweep's
//                           less-sharp / less-flat cloud.  Tiled brute force: a CTA stages 256-point
//                           tiles of the target cloud in shared memory for 64 queries x 4 lanes, target
//                           chunks across blockIdx.y, chunk results merged with a packed 64-bit
//                           atomicMin (d2 bits << 32 | index) => (d2, index) lexicographic minimum,
//                           the oracle's brute-force tie rule.  d2 as FLANN L2_Simple, fp32 no FMA.
//   k_odom_corr  :305-483   one warp per feature: the +-NEARBY_SCAN ring-window scans in the
//                           reference's visiting order (ascending j, then descending j, strict '<'),
//                           reproduced as arg-min over (d2, visit rank); builds the Edge / Plane factor.
//   LM solve     :494-499   lm.cu (shared with laserMapping)
//   k_odom_finish:504-505   t_w_curr += q_w_curr * t_last_curr ; q_w_curr *= q_last_curr
// The "last" clouds (:554-563) stay resident on the device between sweeps.
//
// Measured and rejected (round 2, profiles/odom_nn_box_r02.txt): an exact 1-NN that prunes 32-point tiles of the last cloud
// by their bounding boxes.  The clouds come ring by ring, so a tile is an arc of one ring; near the sensor those arcs
// span tens of degrees and their boxes contain most features of the neighbouring rings (bound 0), so few tiles are pruned
// and the search becomes a chain of dependent L2 round trips: 48 us per pass (131 us with per-round re-pruning) against
// 32 us for the tiled brute force, which streams the targets through shared memory at 82 % issue utilisation.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>
#include <float.h>

struct OdomDev {
  int32_t inited, do_solve;
  int32_t n_sharp, n_flat, n_less_sharp, n_less_flat;
  int32_t n_corner_last, n_surf_last;
  int32_t distortion;               // #define DISTORTION (:59), from lmono_params
  int32_t corner_corr[2], plane_corr[2];
  LmSolveSummary solve[2];
  double para_q[4], para_t[3];      // q_last_curr (x,y,z,w), t_last_curr  (laserOdometry.cpp:97-98)
  double q_w_curr[4], t_w_curr[3];  // :93-94
};

struct OdomState {
  OdomDev* d; OdomDev* h;
  float4* d_feat[4];                // sharp, less_sharp, flat, less_flat of the current sweep
  float4* d_last[2];                // laserCloudCornerLast, laserCloudSurfLast
  unsigned long long* d_best[2];    // packed 1-NN result per sharp / flat feature
  int32_t* d_corr;                  // test hook: [n_sharp*2] + [n_flat*3]
  float* d_frac[2];                 // DISTORTION 1: fractional intensity of every sharp / flat feature (factor ratio s = frac / 0.1)
  int32_t* d_ringtab;               // [2][RT_STRIDE]: per last cloud, first index with ring >= r (r = 0..65) and a "sorted by ring" flag at [66]
  int cap;
};

constexpr int RT_STRIDE = 68;
constexpr int NN_QB = 64, NN_SUB = 4, NN_TILE = 256, NN_CHUNK = 1024;     // 1024 targets per CTA: ~28 x 20 CTAs for an HDL-64 sweep instead of 7 x 20 long ones

// counts: NULL = the host knows the cloud sizes (arguments); else device counts {n_kept, sharp, less_sharp, flat, less_flat}
// as scanRegistration leaves them (fused sweep without a host round trip between the stages)
__global__ void k_odom_begin(OdomDev* o, int n_sharp, int n_ls, int n_flat, int n_lf, const int32_t* __restrict__ counts, int cap, int32_t* __restrict__ ringtab) {
  lm_pdl_enter();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  ringtab[66] = 1; ringtab[RT_STRIDE + 66] = 1;          // "sorted by ring" until k_odom_ring_table finds a descent
  for (int r = 0; r < 66; ++r) { ringtab[r] = o->n_corner_last; ringtab[RT_STRIDE + r] = o->n_surf_last; }
  if (counts) { n_sharp = min(counts[1], cap); n_ls = min(counts[2], cap); n_flat = min(counts[3], cap); n_lf = min(counts[4], cap); }   // an oversize cloud also raises LM_FAULT_FEATURE_OVERFLOW in the mapping stage
  o->n_sharp = n_sharp; o->n_less_sharp = n_ls; o->n_flat = n_flat; o->n_less_flat = n_lf;
  o->do_solve = o->inited;                       // first frame only initialises (:267-271)
  for (int k = 0; k < 2; ++k) {
    o->corner_corr[k] = 0; o->plane_corr[k] = 0;
    o->solve[k].iterations = 0; o->solve[k].num_successful = 0; o->solve[k].termination = 6; o->solve[k].num_factors = 0;
    o->solve[k].initial_cost = 0.0; o->solve[k].final_cost = 0.0;
  }
}

__global__ void __launch_bounds__(256) k_odom_best_init(unsigned long long* b0, int n0, unsigned long long* b1, int n1) {
  lm_pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n0) b0[i] = ~0ULL;
  if (i < n1) b1[i] = ~0ULL;
}

// TransformToStart (:111-129).  DISTORTION 0: s = 1, q_last_curr * p + t_last_curr in double, stored float.  DISTORTION 1:
// s = (intensity - int(intensity)) / SCAN_PERIOD -- the subtraction is float arithmetic (float - int), the division double --
// q_point_last = Identity.slerp(s, q_last_curr) (Eigen 3.3: acos / sin of w, NOT renormalised), t_point_last = s t_last_curr.
__device__ __forceinline__ void d_slerp_identity(double s, const double* q, double* qs) {
  const double d = q[3], absD = fabs(d);
  double scale0, scale1;
  if (absD >= 1.0 - DBL_EPSILON) { scale0 = 1.0 - s; scale1 = s; }
  else { const double theta = acos(absD), sinTheta = sin(theta); scale0 = sin((1.0 - s) * theta) / sinTheta; scale1 = sin(s * theta) / sinTheta; }
  if (d < 0.0) scale1 = -scale1;
  qs[0] = scale1 * q[0]; qs[1] = scale1 * q[1]; qs[2] = scale1 * q[2]; qs[3] = scale0 + scale1 * q[3];
}
__device__ __forceinline__ float d_frac_of(float intensity) { return __fsub_rn(intensity, (float)(int)intensity); }
__device__ __forceinline__ float4 d_to_start(const OdomDev* o, float4 p) {
  if (!o->distortion) return d_associate(o->para_q, o->para_t, p);
  const double s = (double)d_frac_of(p.w) / 0.1;
  double qs[4]; d_slerp_identity(s, o->para_q, qs);
  const double ts[3] = { s * o->para_t[0], s * o->para_t[1], s * o->para_t[2] };
  return d_associate(qs, ts, p);
}

// blockIdx.z: 0 = sharp vs corner_last, 1 = flat vs surf_last
__global__ void __launch_bounds__(NN_QB * NN_SUB) k_odom_nn1(const OdomDev* __restrict__ o, const float4* __restrict__ sharp,
                                                             const float4* __restrict__ flat, const float4* __restrict__ corner_last,
                                                             const float4* __restrict__ surf_last, unsigned long long* __restrict__ best0,
                                                             unsigned long long* __restrict__ best1) {
  lm_pdl_enter();
  __shared__ float4 tile[NN_TILE];
  if (!o->do_solve) return;
  const int which = blockIdx.z;
  const int nq = which == 0 ? o->n_sharp : o->n_flat;
  const int nt = which == 0 ? o->n_corner_last : o->n_surf_last;
  const float4* __restrict__ qs = which == 0 ? sharp : flat;
  const float4* __restrict__ ts = which == 0 ? corner_last : surf_last;
  unsigned long long* __restrict__ best = which == 0 ? best0 : best1;
  const int q0 = blockIdx.x * NN_QB;
  if (q0 >= nq) return;
  const int qi = q0 + (threadIdx.x / NN_SUB);
  const int sub = threadIdx.x % NN_SUB;
  float4 sel = make_float4(0.f, 0.f, 0.f, 0.f);
  if (qi < nq) sel = d_to_start(o, qs[qi]);
  // a lane visits its targets in ascending index order, so "strictly smaller d2 wins" keeps the lowest index among equal
  // distances: one float compare and two selects per pair instead of a 64-bit key build and compare; the lanes' results are
  // then combined on the packed (d2, index) key
  float bd = INFINITY; int bi = -1;
  // target chunks are dealt round-robin over gridDim.y: the grid does not depend on the (device-side) target count
  for (int c0 = blockIdx.y * NN_CHUNK; c0 < nt; c0 += gridDim.y * NN_CHUNK) {
  const int c1 = min(c0 + NN_CHUNK, nt);
  for (int tb = c0; tb < c1; tb += NN_TILE) {
    __syncthreads();
    if (tb + (int)threadIdx.x < c1) tile[threadIdx.x] = ts[tb + threadIdx.x];
    __syncthreads();
    const int tn = min(NN_TILE, c1 - tb);
    for (int k = sub; k < tn; k += NN_SUB) {
      const float4 p = tile[k];
      const float dx = __fsub_rn(sel.x, p.x), dy = __fsub_rn(sel.y, p.y), dz = __fsub_rn(sel.z, p.z);
      float d = __fmul_rn(dx, dx);
      d = __fadd_rn(d, __fmul_rn(dy, dy));
      d = __fadd_rn(d, __fmul_rn(dz, dz));
      if (d < bd) { bd = d; bi = tb + k; }
    }
  }
  }
  unsigned long long bestk = bi >= 0 ? (((unsigned long long)__float_as_uint(bd) << 32) | (uint32_t)bi) : ~0ULL;
#pragma unroll
  for (int ofs = NN_SUB / 2; ofs > 0; ofs >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, bestk, ofs);
    bestk = other < bestk ? other : bestk;
  }
  if (sub == 0 && qi < nq && bestk != ~0ULL) atomicMin(&best[qi], bestk);
}

__device__ __forceinline__ float d_sqdis(float4 a, float4 sel) {
  // (a.x - sel.x)*(a.x - sel.x) + (a.y - sel.y)*(a.y - sel.y) + (a.z - sel.z)*(a.z - sel.z), fp32 (:322-327)
  const float dx = __fsub_rn(a.x, sel.x), dy = __fsub_rn(a.y, sel.y), dz = __fsub_rn(a.z, sel.z);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

__device__ __forceinline__ unsigned long long d_warp_min_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o); v = t < v ? t : v; }
  return v;
}

// Ring tables of the two "last" clouds.  scanRegistration emits its clouds ring by ring, so the ring id (integer part of the
// intensity) is non-decreasing along a cloud and the reference's window scans -- "walk up until ring > id + 2.5" / "walk
// down until ring < id - 2.5" (:315-319,341-345) -- visit exactly the index ranges [closest + 1, first(ring >= id + 3)) and
// [first(ring >= id - 2), closest).  With the bounds known up front the scan needs no early-exit test per step and its
// loads are independent.  A cloud that is NOT monotonic in the ring id (a caller may pass anything) keeps the step-wise
// scan with the reference's break rule.  blockIdx.x: 0 corner_last, 1 surf_last.
constexpr int RT_CTAS = 16;       // CTAs per cloud for the monotonicity check (k_odom_begin arms the flags)
__global__ void __launch_bounds__(256) k_odom_ring_table(const OdomDev* __restrict__ o, const float4* __restrict__ corner_last,
                                                         const float4* __restrict__ surf_last, int32_t* __restrict__ tab_all) {
  lm_pdl_enter();
  const int c = blockIdx.y;
  const float4* __restrict__ pts = c == 0 ? corner_last : surf_last;
  const int n = c == 0 ? o->n_corner_last : o->n_surf_last;
  int32_t* tab = tab_all + c * RT_STRIDE;
  int bad = 0;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {   // ring ids outside [0, 64] (never produced by scanRegistration) also take the step-wise scan
    const int r = (int)pts[j].w;
    const int rp = j > 0 ? (int)pts[j - 1].w : -1;
    bad |= (r < 0) | (r > 64) | (rp > r);
    // j is the first index of every ring in (rp, r]: tab[ring] = first index with ring id >= ring.  Entries above the last
    // ring keep the n that k_odom_begin stored.
    if (r >= 0 && r <= 64) for (int rr = max(rp + 1, 0); rr <= r; ++rr) tab[rr] = j;
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicAnd(&tab[66], 0);
}

template <bool IS_CORNER>
__device__ __forceinline__ void d_corr_candidate(float4 p, float4 sel, int id, uint32_t rank, bool ascending, unsigned long long gate_hi,
                                                 unsigned long long& m2, unsigned long long& m3) {
  const float d = d_sqdis(p, sel);
  const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | rank;
  if ((key >> 32) < gate_hi) {
    const int ring = (int)p.w;
    if (ascending) {
      if (IS_CORNER) { if (ring > id) m2 = key < m2 ? key : m2; }
      else { if (ring <= id) m2 = key < m2 ? key : m2; else m3 = key < m3 ? key : m3; }
    } else {
      if (IS_CORNER) { if (ring < id) m2 = key < m2 ? key : m2; }
      else { if (ring >= id) m2 = key < m2 ? key : m2; else m3 = key < m3 ? key : m3; }
    }
  }
}

// one warp per feature; features [0, n_sharp) are corners, [n_sharp, n_sharp + n_flat) planes
__global__ void __launch_bounds__(256) k_odom_corr(const OdomDev* __restrict__ o, const float4* __restrict__ sharp,
                                                   const float4* __restrict__ flat, const float4* __restrict__ corner_last,
                                                   const float4* __restrict__ surf_last, const unsigned long long* __restrict__ best0,
                                                   const unsigned long long* __restrict__ best1, LmFactor* __restrict__ fac0,
                                                   LmFactor* __restrict__ fac1, int32_t* __restrict__ corr_out, const int32_t* __restrict__ ringtab,
                                                   float* __restrict__ frac0, float* __restrict__ frac1) {
  lm_pdl_enter();
  if (!o->do_solve) return;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int ns = o->n_sharp, nf = o->n_flat;
  if (warp >= ns + nf) return;
  const bool is_corner = warp < ns;
  const int qi = is_corner ? warp : warp - ns;
  const float4 ori = is_corner ? sharp[qi] : flat[qi];
  const float4* __restrict__ last = is_corner ? corner_last : surf_last;
  const int nl = is_corner ? o->n_corner_last : o->n_surf_last;
  LmFactor* f = (is_corner ? fac0 : fac1) + qi;
  int32_t* co = corr_out ? (is_corner ? corr_out + qi * 2 : corr_out + ns * 2 + qi * 3) : nullptr;
  const unsigned long long bk = (is_corner ? best0 : best1)[qi];
  const float d_nn = __uint_as_float((uint32_t)(bk >> 32));
  int closest = -1, ind2 = -1, ind3 = -1;
  if (bk != ~0ULL && (double)d_nn < 25.0) {                   // DISTANCE_SQ_THRESHOLD :305,393
    closest = (int)(uint32_t)bk;
    const float4 sel = d_to_start(o, ori);
    const int id = (int)last[closest].w;                       // closestPointScanID
    unsigned long long m2 = ~0ULL, m3 = ~0ULL;
    const unsigned long long gate = (unsigned long long)__float_as_uint(25.0f) << 32;   // candidates need d2 < 25
    const int32_t* tab = ringtab + (is_corner ? 0 : RT_STRIDE);
    const uint32_t rank0 = 1u << 30;
    if (tab[66]) {
      // ring-sorted cloud: the two windows as index ranges, four independent loads per lane and trip
      const int hi = tab[min(max(id + 3, 0), 65)], lo = tab[min(max(id - 2, 0), 65)];
      const unsigned long long ghi = gate >> 32;
      for (int j0 = closest + 1 + lane; j0 < hi; j0 += 128) {
        float4 p[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (j0 + 32 * u < hi) p[u] = last[j0 + 32 * u];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (j0 + 32 * u < hi) {
          if (is_corner) d_corr_candidate<true>(p[u], sel, id, (uint32_t)(j0 + 32 * u - closest), true, ghi, m2, m3);
          else d_corr_candidate<false>(p[u], sel, id, (uint32_t)(j0 + 32 * u - closest), true, ghi, m2, m3);
        }
      }
      for (int j0 = closest - 1 - lane; j0 >= lo; j0 -= 128) {
        float4 p[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (j0 - 32 * u >= lo) p[u] = last[j0 - 32 * u];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (j0 - 32 * u >= lo) {
          if (is_corner) d_corr_candidate<true>(p[u], sel, id, rank0 + (uint32_t)(closest - (j0 - 32 * u)), false, ghi, m2, m3);
          else d_corr_candidate<false>(p[u], sel, id, rank0 + (uint32_t)(closest - (j0 - 32 * u)), false, ghi, m2, m3);
        }
      }
    } else {
    // ascending j (:312-335, 402-427)
    bool stop = false;
    for (int base = closest + 1; base < nl && !stop; base += 32) {
      const int j = base + lane;
      bool over = false; float4 p = make_float4(0.f, 0.f, 0.f, 0.f); int ring = 0;
      if (j < nl) { p = last[j]; ring = (int)p.w; over = (double)ring > (double)id + 2.5; }
      const unsigned om = __ballot_sync(0xffffffffu, over);
      const int first_over = om ? (__ffs(om) - 1) : 32;
      if (j < nl && lane < first_over) {
        const float d = d_sqdis(p, sel);
        const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (uint32_t)(j - closest);
        if ((key >> 32) < (gate >> 32)) {
          if (is_corner) { if (ring > id) m2 = key < m2 ? key : m2; }
          else { if (ring <= id) m2 = key < m2 ? key : m2; else m3 = key < m3 ? key : m3; }
        }
      }
      stop = om != 0;
    }
    // descending j (:338-361, 430-455); visit ranks continue after the ascending pass
    stop = false;
    for (int base = closest - 1; base >= 0 && !stop; base -= 32) {
      const int j = base - lane;
      bool under = false; float4 p = make_float4(0.f, 0.f, 0.f, 0.f); int ring = 0;
      if (j >= 0) { p = last[j]; ring = (int)p.w; under = (double)ring < (double)id - 2.5; }
      const unsigned um = __ballot_sync(0xffffffffu, under);
      const int first_under = um ? (__ffs(um) - 1) : 32;
      if (j >= 0 && lane < first_under) {
        const float d = d_sqdis(p, sel);
        const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (rank0 + (uint32_t)(closest - j));
        if ((key >> 32) < (gate >> 32)) {
          if (is_corner) { if (ring < id) m2 = key < m2 ? key : m2; }
          else { if (ring >= id) m2 = key < m2 ? key : m2; else m3 = key < m3 ? key : m3; }
        }
      }
      stop = um != 0;
    }
    }
    m2 = d_warp_min_u64(m2); m3 = d_warp_min_u64(m3);
    if (m2 != ~0ULL) { const uint32_t r = (uint32_t)m2; ind2 = r >= rank0 ? closest - (int)(r - rank0) : closest + (int)r; }
    if (m3 != ~0ULL) { const uint32_t r = (uint32_t)m3; ind3 = r >= rank0 ? closest - (int)(r - rank0) : closest + (int)r; }
  }
  if (lane != 0) return;
  (is_corner ? frac0 : frac1)[qi] = d_frac_of(ori.w);
  if (co) { co[0] = closest; co[1] = ind2; if (!is_corner) co[2] = ind3; }
  if (is_corner) {
    if (ind2 < 0) { f->kind = -1; return; }                    // :363
    const float4 a = last[closest], b = last[ind2];
    f->a[0] = a.x; f->a[1] = a.y; f->a[2] = a.z;
    f->b[0] = b.x; f->b[1] = b.y; f->b[2] = b.z;
    f->p[0] = ori.x; f->p[1] = ori.y; f->p[2] = ori.z;
    f->kind = 0;
  } else {
    if (ind2 < 0 || ind3 < 0) { f->kind = -1; return; }        // :457
    const float4 pj = last[closest], pl = last[ind2], pm = last[ind3];
    const double j[3] = { pj.x, pj.y, pj.z }, l[3] = { pl.x, pl.y, pl.z }, m[3] = { pm.x, pm.y, pm.z };
    // lidarFactor.hpp:64-65
    const double u[3] = { j[0] - l[0], j[1] - l[1], j[2] - l[2] }, v[3] = { j[0] - m[0], j[1] - m[1], j[2] - m[2] };
    double n[3] = { u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0] };
    const double z2 = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
    if (z2 > 0.0) { const double nn = sqrt(z2); n[0] /= nn; n[1] /= nn; n[2] /= nn; }
    for (int k = 0; k < 3; ++k) { f->a[k] = j[k]; f->b[k] = n[k]; }
    f->p[0] = ori.x; f->p[1] = ori.y; f->p[2] = ori.z;
    f->kind = 1;
  }
}

__global__ void k_odom_finish(OdomDev* o) {
  lm_pdl_enter();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (o->do_solve) {
    double tmp[3]; d_qrot(o->q_w_curr, o->para_t, tmp);
    for (int k = 0; k < 3; ++k) o->t_w_curr[k] = o->t_w_curr[k] + tmp[k];
    double qn[4]; d_qmul(o->q_w_curr, o->para_q, qn);
    for (int k = 0; k < 4; ++k) o->q_w_curr[k] = qn[k];
  }
  o->inited = 1;
  o->n_corner_last = o->n_less_sharp; o->n_surf_last = o->n_less_flat;   // :554-563
}

// :554-563 the less-sharp / less-flat clouds of this sweep become the "last" clouds (sizes read on the device)
__global__ void __launch_bounds__(256) k_odom_keep_last(const OdomDev* __restrict__ o, const float4* __restrict__ ls, const float4* __restrict__ lf,
                                                        float4* __restrict__ last0, float4* __restrict__ last1) {
  lm_pdl_enter();
  const int n0 = o->n_less_sharp, n1 = o->n_less_flat;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n0 + n1; i += gridDim.x * blockDim.x) {
    if (i < n0) last0[i] = ls[i]; else last1[i - n0] = lf[i - n0];
  }
}

__global__ void k_odom_reset(OdomDev* o, int distortion) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  o->distortion = distortion;
  o->inited = 0; o->do_solve = 0;
  o->n_corner_last = 0; o->n_surf_last = 0;
  for (int k = 0; k < 4; ++k) { o->para_q[k] = k == 3; o->q_w_curr[k] = k == 3; }
  for (int k = 0; k < 3; ++k) { o->para_t[k] = 0.0; o->t_w_curr[k] = 0.0; }
}

static int odom_state(lmono_ctx* ctx, OdomState** out) {
  if (ctx->odom_state) { *out = (OdomState*)ctx->odom_state; return LMONO_OK; }
  OdomState* s = (OdomState*)calloc(1, sizeof(OdomState));
  s->cap = ctx->max_feat;
  const size_t n = (size_t)s->cap;
  LM_CUDA(cudaMalloc((void**)&s->d, sizeof(OdomDev)));
  LM_CUDA(cudaMallocHost((void**)&s->h, sizeof(OdomDev)));
  memset(s->h, 0, sizeof(OdomDev));      // the grid sizes of the first step are read from this mirror before any read-back
  for (int k = 0; k < 4; ++k) LM_CUDA(cudaMalloc((void**)&s->d_feat[k], n * sizeof(float4)));
  for (int k = 0; k < 2; ++k) { LM_CUDA(cudaMalloc((void**)&s->d_last[k], n * sizeof(float4))); LM_CUDA(cudaMalloc((void**)&s->d_best[k], n * sizeof(unsigned long long))); }
  LM_CUDA(cudaMalloc((void**)&s->d_corr, n * 10 * sizeof(int32_t)));
  LM_CUDA(cudaMalloc((void**)&s->d_ringtab, 2 * RT_STRIDE * sizeof(int32_t)));
  for (int k = 0; k < 2; ++k) LM_CUDA(cudaMalloc((void**)&s->d_frac[k], n * sizeof(float)));
  k_odom_reset<<<1, 32, 0, ctx->stream>>>(s->d, ctx->prm.distortion ? 1 : 0);
  LM_LAUNCH_CHECK();
  ctx->odom_state = s;
  *out = s;
  return LMONO_OK;
}

// allocate the stage's state outside any stream capture (cudaMalloc is not allowed while this thread captures)
int lm_odom_prepare(lmono_ctx* ctx) { OdomState* s; return odom_state(ctx, &s); }

void lm_odom_free(lmono_ctx* ctx) {
  OdomState* s = (OdomState*)ctx->odom_state;
  if (!s) return;
  cudaFree(s->d); cudaFreeHost(s->h);
  for (int k = 0; k < 4; ++k) cudaFree(s->d_feat[k]);
  for (int k = 0; k < 2; ++k) { cudaFree(s->d_last[k]); cudaFree(s->d_best[k]); }
  cudaFree(s->d_corr); cudaFree(s->d_ringtab); cudaFree(s->d_frac[0]); cudaFree(s->d_frac[1]);
  free(s); ctx->odom_state = nullptr;
}

// one association pass on device-resident clouds of the current sweep
static int odom_associate(lmono_ctx* ctx, OdomState* s, const float4* sharp, int n_sharp, const float4* flat, int n_flat, int nl_max0, int nl_max1, int pass) {
  const int nq = n_sharp > n_flat ? n_sharp : n_flat;
  if (nq <= 0) return LMONO_OK;
  LM_LAUNCH_PDL(k_odom_best_init, lm_div_up(nq, 256), 256, 0, s->d_best[0], n_sharp, s->d_best[1], n_flat);
  LM_LAUNCH_CHECK();
  const int nlm = nl_max0 > nl_max1 ? nl_max0 : nl_max1;
  if (nlm > 0) {
    const int gy = lm_div_up(nlm, NN_CHUNK);
    dim3 grid(lm_div_up(nq, NN_QB), gy < 48 ? gy : 48, 2);
    LM_LAUNCH_PDL(k_odom_nn1, grid, NN_QB * NN_SUB, 0, s->d, sharp, flat, s->d_last[0], s->d_last[1], s->d_best[0], s->d_best[1]);
    LM_LAUNCH_CHECK();
  }
  const int warps = n_sharp + n_flat;
  LM_LAUNCH_PDL(k_odom_corr, lm_div_up(warps * 32, 256), 256, 0, s->d, sharp, flat, s->d_last[0], s->d_last[1], s->d_best[0], s->d_best[1],
                                                                  ctx->d_fac[0], ctx->d_fac[1], s->d_corr + (size_t)pass * s->cap * 5, s->d_ringtab, s->d_frac[0], s->d_frac[1]);
  LM_LAUNCH_CHECK();
  return LMONO_OK;
}

// enqueue one odometry step on device-resident feature clouds (sizes known to the host)
// d_counts != NULL: the cloud sizes are read on the device (n_* are then upper bounds that only size the grids)
int lm_odom_enqueue(lmono_ctx* ctx, const float4* sharp, int n_sharp, const float4* less_sharp, int n_ls,
                    const float4* flat, int n_flat, const float4* less_flat, int n_lf, int prev_ls_max, int prev_lf_max,
                    const int32_t* d_counts = nullptr) {
  OdomState* s; int rc = odom_state(ctx, &s); if (rc) return rc;
  LM_LAUNCH_PDL(k_odom_begin, 1, 32, 0, s->d, n_sharp, n_ls, n_flat, n_lf, d_counts, s->cap, s->d_ringtab);
  LM_LAUNCH_CHECK();
  LM_LAUNCH_PDL(k_odom_ring_table, dim3(RT_CTAS, 2), 256, 0, s->d, s->d_last[0], s->d_last[1], s->d_ringtab);
  LM_LAUNCH_CHECK();
  for (int opti = 0; opti < 2; ++opti) {                       // :278
    if ((rc = odom_associate(ctx, s, sharp, n_sharp, flat, n_flat, prev_ls_max, prev_lf_max, opti))) return rc;
    LmProblem P;
    P.fac0 = ctx->d_fac[0]; P.fac1 = ctx->d_fac[1];
    P.n0 = &s->d->n_sharp; P.n1 = &s->d->n_flat;
    P.gate = &s->d->do_solve;
    P.pose_q = s->d->para_q; P.pose_t = s->d->para_t;
    P.summary = &s->d->solve[opti];
    P.count0 = &s->d->corner_corr[opti]; P.count1 = &s->d->plane_corr[opti];
    P.frac0 = ctx->prm.distortion ? s->d_frac[0] : nullptr; P.frac1 = ctx->prm.distortion ? s->d_frac[1] : nullptr;
    if ((rc = lm_solve_problem(ctx, P, n_sharp + n_flat, 4, 1))) return rc;
  }
  LM_LAUNCH_PDL(k_odom_finish, 1, 32, 0, s->d);
  LM_LAUNCH_CHECK();
  if (d_counts) {
    const int nb = lm_div_up(n_ls + n_lf > 0 ? n_ls + n_lf : 1, 256);
    LM_LAUNCH_PDL(k_odom_keep_last, nb < 296 ? nb : 296, 256, 0, s->d, less_sharp, less_flat, s->d_last[0], s->d_last[1]);
    LM_LAUNCH_CHECK();
    return LMONO_OK;
  }
  if (n_ls > 0) LM_CUDA(cudaMemcpyAsync(s->d_last[0], less_sharp, (size_t)n_ls * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
  if (n_lf > 0) LM_CUDA(cudaMemcpyAsync(s->d_last[1], less_flat, (size_t)n_lf * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
  return LMONO_OK;
}

// fused sweep without a host round trip after scanRegistration: every cloud size stays on the device.  Grids are sized
// from bounds the host knows: <= 2 / 20 / 4 picks per (ring, sector) (scanRegistration.cpp:297-310,352-358), the less-flat
// cloud and the previous sweep's clouds by the raw sizes of this and the previous sweep.
int lm_odom_enqueue_devcounts(lmono_ctx* ctx, const float4* sharp, const float4* less_sharp, const float4* flat, const float4* less_flat,
                              const int32_t* d_counts, int n_raw, int n_raw_prev, const double** d_pose7) {
  OdomState* s; int rc = odom_state(ctx, &s); if (rc) return rc;
  const int rings = ctx->prm.scan_line;
  const int b_sharp = rings * 6 * 2, b_ls = rings * 6 * 20, b_flat = rings * 6 * 4;
  const int b_lf = n_raw < s->cap ? n_raw : s->cap, b_lf_prev = n_raw_prev < s->cap ? n_raw_prev : s->cap;
  if (b_ls > s->cap) return LMONO_E_CAPACITY;
  if ((rc = lm_odom_enqueue(ctx, sharp, b_sharp, less_sharp, b_ls, flat, b_flat, less_flat, b_lf, b_ls, b_lf_prev > 0 ? b_lf_prev : 1, d_counts))) return rc;
  *d_pose7 = s->d->q_w_curr;
  return LMONO_OK;
}

// fused sweep (mapping.cu: lmono_sweep_step): the features are the scan stage's device buffers; the result pose
// (q_w_curr[4], t_w_curr[3], contiguous) is handed to the mapping stage as a device pointer
int lm_odom_enqueue_auto(lmono_ctx* ctx, const float4* sharp, int n_sharp, const float4* less_sharp, int n_ls,
                         const float4* flat, int n_flat, const float4* less_flat, int n_lf, const double** d_pose7) {
  OdomState* s; int rc = odom_state(ctx, &s); if (rc) return rc;
  if (n_sharp > s->cap || n_ls > s->cap || n_flat > s->cap || n_lf > s->cap) return LMONO_E_CAPACITY;
  static_assert(offsetof(OdomDev, t_w_curr) == offsetof(OdomDev, q_w_curr) + 4 * sizeof(double), "pose7 layout");
  if ((rc = lm_odom_enqueue(ctx, sharp, n_sharp, less_sharp, n_ls, flat, n_flat, less_flat, n_lf,
                            s->h->n_corner_last > 0 ? s->h->n_corner_last : s->cap, s->h->n_surf_last > 0 ? s->h->n_surf_last : s->cap))) return rc;
  *d_pose7 = s->d->q_w_curr;
  return LMONO_OK;
}
int lm_odom_readback(lmono_ctx* ctx) {
  OdomState* s; int rc = odom_state(ctx, &s); if (rc) return rc;
  LM_CUDA(cudaMemcpyAsync(s->h, s->d, sizeof(OdomDev), cudaMemcpyDeviceToHost, ctx->stream));
  return LMONO_OK;
}
static void odom_fill(const OdomDev* h, lmono_pose* last_curr, lmono_pose* w_curr, lmono_odom_report* report) {
  if (last_curr) { memcpy(last_curr->q, h->para_q, 32); memcpy(last_curr->t, h->para_t, 24); }
  if (w_curr) { memcpy(w_curr->q, h->q_w_curr, 32); memcpy(w_curr->t, h->t_w_curr, 24); }
  if (report) {
    memset(report, 0, sizeof(*report));
    report->inited = h->do_solve;
    for (int k = 0; k < 2; ++k) {
      report->corner_corr[k] = h->corner_corr[k]; report->plane_corr[k] = h->plane_corr[k];
      report->solve[k].iterations = h->solve[k].iterations; report->solve[k].num_successful = h->solve[k].num_successful;
      report->solve[k].termination = h->solve[k].termination; report->solve[k].num_factors = h->solve[k].num_factors;
      report->solve[k].initial_cost = h->solve[k].initial_cost; report->solve[k].final_cost = h->solve[k].final_cost;
    }
  }
}
// after a stream synchronisation that followed lm_odom_readback
int lm_odom_deliver(lmono_ctx* ctx, lmono_pose* last_curr, lmono_pose* w_curr, lmono_odom_report* report) {
  OdomState* s = (OdomState*)ctx->odom_state;
  if (!s) return LMONO_E_STATE;
  odom_fill(s->h, last_curr, w_curr, report);
  return LMONO_OK;
}

extern "C" int lmono_odom_step(lmono_ctx* ctx, lmono_cloud_view sharp, lmono_cloud_view less_sharp, lmono_cloud_view flat,
                               lmono_cloud_view less_flat, lmono_pose* last_curr, lmono_pose* w_curr, lmono_odom_report* report) {
  if (!ctx) return LMONO_E_ARG;
  OdomState* s; int rc = odom_state(ctx, &s); if (rc) return rc;
  if (sharp.n > s->cap || less_sharp.n > s->cap || flat.n > s->cap || less_flat.n > s->cap) return LMONO_E_CAPACITY;
  LM_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  lmono_cloud_view v[4] = { sharp, less_sharp, flat, less_flat };
  for (int k = 0; k < 4; ++k) if ((rc = lm_upload_cloud(ctx, v[k], ctx->d_raw[k < 3 ? k : 0], s->d_feat[k], nullptr))) return rc;
  LM_CUDA(cudaEventRecord(ctx->ev_k0, ctx->stream));
  lm_kmark(ctx, "begin", 0);
  // sizes of the previous sweep's clouds bound the search grids; cap is always safe
  if ((rc = lm_odom_enqueue(ctx, s->d_feat[0], sharp.n, s->d_feat[1], less_sharp.n, s->d_feat[2], flat.n, s->d_feat[3], less_flat.n,
                            s->h->n_corner_last > 0 ? s->h->n_corner_last : s->cap, s->h->n_surf_last > 0 ? s->h->n_surf_last : s->cap))) return rc;
  LM_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  LM_CUDA(cudaMemcpyAsync(s->h, s->d, sizeof(OdomDev), cudaMemcpyDeviceToHost, ctx->stream));
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  const OdomDev* h = s->h;
  if (last_curr) { memcpy(last_curr->q, h->para_q, 32); memcpy(last_curr->t, h->para_t, 24); }
  if (w_curr) { memcpy(w_curr->q, h->q_w_curr, 32); memcpy(w_curr->t, h->t_w_curr, 24); }
  if (report) {
    memset(report, 0, sizeof(*report));
    report->inited = h->do_solve;
    for (int k = 0; k < 2; ++k) {
      report->corner_corr[k] = h->corner_corr[k]; report->plane_corr[k] = h->plane_corr[k];
      report->solve[k].iterations = h->solve[k].iterations; report->solve[k].num_successful = h->solve[k].num_successful;
      report->solve[k].termination = h->solve[k].termination; report->solve[k].num_factors = h->solve[k].num_factors;
      report->solve[k].initial_cost = h->solve[k].initial_cost; report->solve[k].final_cost = h->solve[k].final_cost;
    }
    float ms = 0.f; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); report->ms_gpu = ms;
  }
  return LMONO_OK;
}

extern "C" int lmono_odom_reset(lmono_ctx* ctx) {
  if (!ctx) return LMONO_E_ARG;
  OdomState* s; int rc = odom_state(ctx, &s); if (rc) return rc;
  k_odom_reset<<<1, 32, 0, ctx->stream>>>(s->d, ctx->prm.distortion ? 1 : 0);
  LM_LAUNCH_CHECK();
  memset(s->h, 0, sizeof(OdomDev));
  return LMONO_OK;
}

// test hook: correspondences of association pass `pass` (0 or 1) of the last step:
// corner_idx [n_sharp x 2] (closest, min2), plane_idx [n_flat x 3] (closest, min2, min3)
extern "C" int lmono_odom_debug(lmono_ctx* ctx, int32_t pass, int32_t* corner_idx, int32_t n_sharp, int32_t* plane_idx, int32_t n_flat) {
  OdomState* s = ctx ? (OdomState*)ctx->odom_state : nullptr;
  if (!s || pass < 0 || pass > 1) return LMONO_E_STATE;
  const int32_t* src = s->d_corr + (size_t)pass * s->cap * 5;
  if (corner_idx && n_sharp > 0) LM_CUDA(cudaMemcpyAsync(corner_idx, src, (size_t)n_sharp * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  if (plane_idx && n_flat > 0) LM_CUDA(cudaMemcpyAsync(plane_idx, src + (size_t)n_sharp * 2, (size_t)n_flat * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  return LMONO_OK;
}
