// Device-resident rolling cube map of laserMapping (Aloam/src/laserMapping.cpp:74-104).
//
// Layout in HBM (per map type: 0 corner, 1 surf):
//   * the reference's 21x21x11 array of cube clouds becomes a ring buffer: absolute cube
//     g = logical index - laserCloudCen{Width,Height,Depth} lives in physical slot
//     (g mod 21, g mod 21, g mod 11); the pointer rotations of :323-507 reduce to updating
//     the three Cen offsets and freeing the recycled plane.
//   * each non-empty cube owns a slab from a pool: float4 XYZI points in exactly the order
//     the reference's cube cloud would have (VoxelGrid output order = ascending voxel key,
//     then appended points), ping-pong buffered so a refilter never works in place.
//   * per slab a 26^3 table of 2 m cells (key floor(x)>>1) with a cell-sorted copy of the
//     points carrying their slab position: the exact-kNN search structure (assoc.cu).
//
// Per-sweep upkeep is O(points touched), not O(map): only cubes that received points are
// re-filtered (a VoxelGrid pass over an already filtered cube is the identity, SURVEY
// App. B.3), by merging the stably sorted new points into the sorted prefix.
#include "common.cuh"

// ------------------------------------------------------------------ allocation
template <typename T> static int dev_alloc(lmono_ctx* ctx, T** p, size_t count) {
  LM_CUDA(cudaMalloc((void**)p, count * sizeof(T)));
  return LMONO_OK;
}

__global__ void k_map_reset(LmMapType M) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < LM_NSLOT; i += gridDim.x * blockDim.x) M.slot_slab[i] = -1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M.n_slabs; i += gridDim.x * blockDim.x) {
    M.free_stack[i] = M.n_slabs - 1 - i;   // pop order 0,1,2,...
    M.slab_n[i] = 0; M.slab_nsorted[i] = 0; M.slab_cur[i] = 0; M.slab_dirty[i] = 0; M.slab_unsorted[i] = 0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *M.free_top = M.n_slabs;
}

int lm_map_alloc(lmono_ctx* ctx) {
  for (int ty = 0; ty < 2; ++ty) {
    LmMapType& M = ctx->map[ty];
    M.leaf = ty == 0 ? ctx->prm.mapping_line_resolution : ctx->prm.mapping_plane_resolution;
    M.inv_leaf = 1.0f / M.leaf;
    M.cap = ty == 0 ? ctx->prm.cube_capacity_corner : ctx->prm.cube_capacity_surf;
    M.n_slabs = ty == 0 ? ctx->prm.max_cubes_corner : ctx->prm.max_cubes_surf;
    int rc;
    if ((rc = dev_alloc(ctx, &M.pts, (size_t)M.n_slabs * 2 * M.cap))) return rc;
    if ((rc = dev_alloc(ctx, &M.cellpts, (size_t)M.n_slabs * M.cap))) return rc;
    if ((rc = dev_alloc(ctx, &M.cellstart, (size_t)M.n_slabs * (LM_NCELL + 1)))) return rc;
    if ((rc = dev_alloc(ctx, &M.pkey, (size_t)M.n_slabs * 2 * M.cap))) return rc;
    if ((rc = dev_alloc(ctx, &M.cellcount, (size_t)M.n_slabs * LM_NCELL))) return rc;
    if ((rc = dev_alloc(ctx, &M.slab_unsorted, (size_t)M.n_slabs))) return rc;
    if ((rc = dev_alloc(ctx, &M.slot_slab, (size_t)LM_NSLOT))) return rc;
    if ((rc = dev_alloc(ctx, &M.slab_n, (size_t)M.n_slabs))) return rc;
    if ((rc = dev_alloc(ctx, &M.slab_nsorted, (size_t)M.n_slabs))) return rc;
    if ((rc = dev_alloc(ctx, &M.slab_cur, (size_t)M.n_slabs))) return rc;
    if ((rc = dev_alloc(ctx, &M.slab_dirty, (size_t)M.n_slabs))) return rc;
    if ((rc = dev_alloc(ctx, &M.slab_g, (size_t)M.n_slabs * 4))) return rc;
    if ((rc = dev_alloc(ctx, &M.free_stack, (size_t)M.n_slabs))) return rc;
    if ((rc = dev_alloc(ctx, &M.free_top, 1))) return rc;
    k_map_reset<<<32, 256, 0, ctx->stream>>>(M);
    LM_LAUNCH_CHECK();
  }
  return LMONO_OK;
}

void lm_map_free(lmono_ctx* ctx) {
  for (int ty = 0; ty < 2; ++ty) {
    LmMapType& M = ctx->map[ty];
    cudaFree(M.pts); cudaFree(M.cellpts); cudaFree(M.cellstart); cudaFree(M.pkey); cudaFree(M.cellcount); cudaFree(M.slab_unsorted); cudaFree(M.slot_slab);
    cudaFree(M.slab_n); cudaFree(M.slab_nsorted); cudaFree(M.slab_cur); cudaFree(M.slab_dirty);
    cudaFree(M.slab_g); cudaFree(M.free_stack); cudaFree(M.free_top);
  }
}

int lm_map_clear_device(lmono_ctx* ctx) {
  LM_NEED_MAP();
  for (int ty = 0; ty < 2; ++ty) { k_map_reset<<<32, 256, 0, ctx->stream>>>(ctx->map[ty]); LM_LAUNCH_CHECK(); }
  return LMONO_OK;
}

// ------------------------------------------------------------------ begin step: pose, window
struct PoseArg { double q[4]; double t[3]; };

__device__ __forceinline__ void d_free_slot(const LmMapType& M, int ps) {
  int sid = M.slot_slab[ps];
  if (sid >= 0) {
    M.slot_slab[ps] = -1;
    M.slab_n[sid] = 0; M.slab_nsorted[sid] = 0; M.slab_dirty[sid] = 0; M.slab_unsorted[sid] = 0;
    int pos = atomicAdd(M.free_top, 1);
    M.free_stack[pos] = sid;
  }
}

// Capacity valve for long drives (the reference's cube clouds are unbounded, the slab pools are not): give back the slabs of
// every cube farther than `keep` cubes (Chebyshev distance) from the window's centre cube.  Those cubes lie outside the
// 5x5x3 window, so the next registrations are unaffected; their points are gone from the map.
__global__ void __launch_bounds__(256) k_map_evict(const LmMapState* __restrict__ st, LmMapType M0, LmMapType M1, int keep, int32_t* __restrict__ n_freed) {
  const int keep_c = keep < 3 ? 3 : keep;            // never inside the window (+-2 cubes) or its rim
  for (int ps = blockIdx.x * blockDim.x + threadIdx.x; ps < LM_NSLOT; ps += gridDim.x * blockDim.x) {
    for (int ty = 0; ty < 2; ++ty) {
      const LmMapType& M = ty == 0 ? M0 : M1;
      const int sid = M.slot_slab[ps];
      if (sid < 0) continue;
      const int di = abs(M.slab_g[sid * 4 + 0] + st->cen[0] - st->center[0]);
      const int dj = abs(M.slab_g[sid * 4 + 1] + st->cen[1] - st->center[1]);
      const int dk = abs(M.slab_g[sid * 4 + 2] + st->cen[2] - st->center[2]);
      if (max(di, max(dj, dk)) > keep_c) { d_free_slot(M, ps); atomicAdd(n_freed, 1); }
    }
  }
}
int lm_map_evict_device(lmono_ctx* ctx, int keep, int* n_freed) {
  LM_NEED_MAP();
  int32_t* d_n = ctx->d_tmp_i32;
  LM_CUDA(cudaMemsetAsync(d_n, 0, sizeof(int32_t), ctx->stream));
  k_map_evict<<<lm_div_up(LM_NSLOT, 256), 256, 0, ctx->stream>>>(ctx->d_state, ctx->map[0], ctx->map[1], keep, d_n);
  LM_LAUNCH_CHECK();
  int n = 0;
  LM_CUDA(cudaMemcpyAsync(&n, d_n, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  if (n_freed) *n_freed = n;
  return LMONO_OK;
}

// transformAssociateToMap (:142-146), centre cube (:312-321), the six shift loops
// (:323-507), the valid list (:512-529) and the per-type offsets of the concatenation
// (:533-539), all on device so consecutive sweeps need no host round trip.
__global__ void __launch_bounds__(256) k_begin_step(LmMapState* __restrict__ st, LmMapType M0, LmMapType M1,
                                                    int32_t* __restrict__ slot_valid_rank, int32_t* __restrict__ plan, PoseArg odom,
                                                    int use_override, double ox, double oy, double oz) {
  lm_pdl_enter();
  __shared__ unsigned clear_mask[3];
  __shared__ int s_dirty_n;
  if (threadIdx.x == 0) s_dirty_n = 0;
  __shared__ int s_n[2][128];
  __shared__ int s_own[128];
  __shared__ int s_all, s_center[3], s_cen[3], s_shard[2];
  if (threadIdx.x == 0) {
    if (use_override == 2) {               // pose already stored by k_step_args (CUDA-graph replay: no per-step kernel arguments)
      for (int k = 0; k < 4; ++k) odom.q[k] = st->q_wodom_curr[k];
      for (int k = 0; k < 3; ++k) odom.t[k] = st->t_wodom_curr[k];
      use_override = 0;
    }
    double qm[4], tm[3];
    for (int k = 0; k < 4; ++k) qm[k] = st->q_wmap_wodom[k];
    for (int k = 0; k < 3; ++k) tm[k] = st->t_wmap_wodom[k];
    int cen3[3] = { st->cen[0], st->cen[1], st->cen[2] };
    s_shard[0] = st->shard_rank; s_shard[1] = st->shard_n;
    for (int k = 0; k < 4; ++k) st->q_wodom_curr[k] = odom.q[k];
    for (int k = 0; k < 3; ++k) st->t_wodom_curr[k] = odom.t[k];
    double tw[3], qw[4];
    if (use_override) {
      tw[0] = ox; tw[1] = oy; tw[2] = oz;
      qw[0] = qw[1] = qw[2] = 0.0; qw[3] = 1.0;
    } else {
      d_qmul(qm, odom.q, qw);
      double tmp[3]; d_qrot(qm, odom.t, tmp);
      for (int k = 0; k < 3; ++k) tw[k] = tmp[k] + tm[k];
    }
    for (int k = 0; k < 4; ++k) st->q_w_curr[k] = qw[k];
    for (int k = 0; k < 3; ++k) st->t_w_curr[k] = tw[k];
    const int dim[3] = { LM_GW, LM_GH, LM_GD };
    int all = 0;
    for (int a = 0; a < 3; ++a) {
      unsigned mask = 0;
      int cen = cen3[a];
      int c = d_cube_coord(tw[a], cen);
      int shifts = 0;
      while (c < 3) {                    // contents move up, logical plane dim-1 is recycled
        mask |= 1u << d_pmod(dim[a] - 1 - cen, dim[a]);
        c++; cen++;
        if (++shifts >= dim[a]) { int need = 3 - c; if (need > 0) { c += need; cen += need; } all = 1; break; }
      }
      shifts = 0;
      while (c >= dim[a] - 3) {          // contents move down, logical plane 0 is recycled
        mask |= 1u << d_pmod(-cen, dim[a]);
        c--; cen--;
        if (++shifts >= dim[a]) { int need = c - (dim[a] - 4); if (need > 0) { c -= need; cen -= need; } all = 1; break; }
      }
      st->cen[a] = cen; st->center[a] = c;
      s_cen[a] = cen; s_center[a] = c;
      clear_mask[a] = mask;
    }
    s_all = all;
    st->corner_num[0] = st->corner_num[1] = st->surf_num[0] = st->surf_num[1] = 0;
    for (int s = 0; s < 2; ++s) { st->solve[s].iterations = 0; st->solve[s].num_successful = 0; st->solve[s].termination = 6; st->solve[s].num_factors = 0; st->solve[s].initial_cost = 0.0; st->solve[s].final_cost = 0.0; }
  }
  __syncthreads();
  // free recycled planes (rare) and reset the slot -> window-rank table
  const unsigned mi = clear_mask[0], mj = clear_mask[1], mk = clear_mask[2];
  const int all = s_all;
  if (mi | mj | mk | (unsigned)all) {
    for (int ps = threadIdx.x; ps < LM_NSLOT; ps += blockDim.x) {
      int pi = ps % LM_GW, pj = (ps / LM_GW) % LM_GH, pk = ps / (LM_GW * LM_GH);
      if (all || ((mi >> pi) & 1u) || ((mj >> pj) & 1u) || ((mk >> pk) & 1u)) { d_free_slot(M0, ps); d_free_slot(M1, ps); }
    }
  }
  int2* slot_info = lm_slot_info(slot_valid_rank);
  for (int ps = threadIdx.x; ps < LM_NSLOT; ps += blockDim.x) {
    slot_valid_rank[ps] = -1;
    slot_info[ps] = make_int2(-1, 0); slot_info[LM_NSLOT + ps] = make_int2(-1, 0);
  }
  __syncthreads();
  // valid cubes in the reference's loop order (i outer, j, k inner; :512-529), one thread per cube: the clipped
  // window is a product of index ranges, so the rank of (i, j, k) in loop order is closed-form
  const int i0 = max(s_center[0] - 2, 0), i1 = min(s_center[0] + 2, LM_GW - 1);
  const int j0 = max(s_center[1] - 2, 0), j1 = min(s_center[1] + 2, LM_GH - 1);
  const int k0 = max(s_center[2] - 1, 0), k1 = min(s_center[2] + 1, LM_GD - 1);
  const int ni = max(i1 - i0 + 1, 0), nj = max(j1 - j0 + 1, 0), nk = max(k1 - k0 + 1, 0);
  const int vn = ni * nj * nk;
  if (threadIdx.x < 128) { s_n[0][threadIdx.x] = 0; s_n[1][threadIdx.x] = 0; s_own[threadIdx.x] = 0; }
  if ((int)threadIdx.x < vn) {
    const int r = threadIdx.x;
    const int gi = i0 + r / (nj * nk) - s_cen[0], gj = j0 + (r / nk) % nj - s_cen[1], gk = k0 + r % nk - s_cen[2];
    const int ps = d_phys_slot(gi, gj, gk);
    st->valid_slot[r] = ps;
    slot_valid_rank[ps] = r;
    s_own[r] = lm_cube_owner(gi, gj, gk, s_shard[1]) == s_shard[0];
    const int s0 = M0.slot_slab[ps], s1 = M1.slot_slab[ps];
    s_n[0][r] = s0 >= 0 ? M0.slab_n[s0] : 0;
    s_n[1][r] = s1 >= 0 ? M1.slab_n[s1] : 0;
    // window cubes whose search index is stale (import, points that arrived while the cube was outside the window):
    // the work list of k_index_build
    if (s0 >= 0 && M0.slab_dirty[s0]) plan[LM_PLAN_DIRTY + atomicAdd(&s_dirty_n, 1)] = r;
    if (s1 >= 0 && M1.slab_dirty[s1]) plan[LM_PLAN_DIRTY + atomicAdd(&s_dirty_n, 1)] = LM_WIN_MAX + r;
  }
  if (threadIdx.x == 0) st->valid_num = vn;
  __syncthreads();
  if (threadIdx.x == 0) plan[LM_PLAN_DIRTY_N] = s_dirty_n;
  // exclusive prefix of the cube sizes per map type (the :533-537 concatenation offsets): warp ty scans 4 entries per lane
  const int lane = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (ty < 2) {
    int v[4], own = 0, sum = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) { v[u] = s_n[ty][lane * 4 + u]; sum += v[u]; own += s_own[lane * 4 + u] ? v[u] : 0; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    int run = incl - sum;
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int r = lane * 4 + u; if (r <= vn) st->valid_off[ty][r] = run; run += v[u]; }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) own += __shfl_xor_sync(0xffffffffu, own, o);
    if (lane == 0) { st->from_map_n[ty] = total; st->shard_owned_n[ty] = own; s_n[ty][127] = total; }
  }
  __syncthreads();
  if (threadIdx.x == 0) st->optimize = (s_n[0][127] > 10 && s_n[1][127] > 50) ? 1 : 0;   // :554
  // search table of the association kernels: physical slot -> (slab id, index of the cube's first point in the
  // :533-537 concatenation), one 8-byte load per (query, cube) instead of three dependent ones
  if ((int)threadIdx.x < vn) {
    const int r = threadIdx.x;
    const int ps = st->valid_slot[r];
    slot_info[ps] = make_int2(M0.slot_slab[ps], st->valid_off[0][r]);
    slot_info[LM_NSLOT + ps] = make_int2(M1.slot_slab[ps], st->valid_off[1][r]);
  }
}

int lm_map_begin_step(lmono_ctx* ctx, const lmono_pose* wodom_curr, const double* t_override) {
  LM_NEED_MAP();
  PoseArg pa;
  if (wodom_curr) { for (int k = 0; k < 4; ++k) pa.q[k] = wodom_curr->q[k]; for (int k = 0; k < 3; ++k) pa.t[k] = wodom_curr->t[k]; }
  else { pa.q[0] = pa.q[1] = pa.q[2] = 0; pa.q[3] = 1; pa.t[0] = pa.t[1] = pa.t[2] = 0; }
  LM_LAUNCH_PDL(k_begin_step, 1, 256, 0, ctx->d_state, ctx->map[0], ctx->map[1], ctx->d_slot_valid_rank, ctx->d_rf_plan, pa,
                                           t_override ? 1 : (wodom_curr ? 0 : 2), t_override ? t_override[0] : 0.0,
                                           t_override ? t_override[1] : 0.0, t_override ? t_override[2] : 0.0);
  LM_LAUNCH_CHECK();
  return LMONO_OK;
}

// ------------------------------------------------------------------ cell index of one slab
// s_cnt: shared uint32[LM_NCELL]; ws: shared int[33].  All threads of the block call this.
__device__ void d_build_cell_index(const LmMapType& M, int sid, const float4* __restrict__ src, int n,
                                   uint32_t* s_cnt, int* ws, LmMapState* st) {
  const int* g = M.slab_g + sid * 4;
  const int g3[3] = { g[0], g[1], g[2] };
  for (int i = threadIdx.x; i < LM_NCELL; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  // four independent loads in flight per trip (a plain loop waits for each load before it issues the next: one L2 / HBM round
  // trip per 1024 points, twice over the cube)
  for (int i0 = threadIdx.x; i0 < n; i0 += 4 * blockDim.x) {
    float4 p[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int i = i0 + u * blockDim.x; if (i < n) p[u] = src[i]; }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < n) {
        int c = d_cube_cell(p[u], g3);
        if (c < 0) { atomicOr(&st->fault, LM_FAULT_CELL_RANGE); c = 0; }
        atomicAdd(&s_cnt[c], 1u);
      }
    }
  }
  __syncthreads();
  const int per = (LM_NCELL + blockDim.x - 1) / blockDim.x;
  const int b = min((int)threadIdx.x * per, LM_NCELL), e = min(b + per, LM_NCELL);
  int sum = 0;
  for (int c = b; c < e; ++c) sum += (int)s_cnt[c];
  int total;
  int run = d_block_exscan(sum, ws, &total);
  for (int c = b; c < e; ++c) { int v = (int)s_cnt[c]; s_cnt[c] = (uint32_t)run; run += v; }
  __syncthreads();
  uint32_t* cs = M.cellstart + (size_t)sid * (LM_NCELL + 1);
  for (int i = threadIdx.x; i < LM_NCELL; i += blockDim.x) cs[i] = s_cnt[i];
  if (threadIdx.x == 0) cs[LM_NCELL] = (uint32_t)n;
  __syncthreads();
  float4* cp = M.cellpts + (size_t)sid * M.cap;
  for (int i0 = threadIdx.x; i0 < n; i0 += 4 * blockDim.x) {
    float4 p[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int i = i0 + u * blockDim.x; if (i < n) p[u] = src[i]; }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      if (i < n) {
        int c = d_cube_cell(p[u], g3);
        if (c < 0) c = 0;
        const uint32_t pos = atomicAdd(&s_cnt[c], 1u);
        p[u].w = __int_as_float(i);
        cp[pos] = p[u];
      }
    }
  }
  __syncthreads();
}

constexpr int IDX_GRID = 32;
__global__ void __launch_bounds__(1024, 1) k_index_build(LmMapState* __restrict__ st, LmMapType M0, LmMapType M1, const int32_t* __restrict__ plan) {
  lm_pdl_enter();
  extern __shared__ unsigned char smem_raw[];
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(smem_raw);
  int* ws = reinterpret_cast<int*>(s_cnt + LM_NCELL);
  const int nd = plan[LM_PLAN_DIRTY_N];
  for (int w = blockIdx.x; w < nd; w += gridDim.x) {
    const int e = plan[LM_PLAN_DIRTY + w];
    const int ty = e / LM_WIN_MAX, r = e - ty * LM_WIN_MAX;
    const LmMapType& M = ty == 0 ? M0 : M1;
    const int sid = M.slot_slab[st->valid_slot[r]];
    const int n = M.slab_n[sid];
    const float4* src = M.pts + ((size_t)sid * 2 + M.slab_cur[sid]) * M.cap;
    d_build_cell_index(M, sid, src, n, s_cnt, ws, st);
    if (threadIdx.x == 0) M.slab_dirty[sid] = 0;
    __syncthreads();
  }
}

static const int kIndexSmem = LM_NCELL * 4 + 64 * 4;

int lm_map_index_build(lmono_ctx* ctx) {
  LM_LAUNCH_PDL(k_index_build, IDX_GRID, 1024, kIndexSmem, ctx->d_state, ctx->map[0], ctx->map[1], ctx->d_rf_plan);
  LM_LAUNCH_CHECK();
  return LMONO_OK;
}

// ------------------------------------------------------------------ insertion (:737-783)
// composite = type (1) | physical slot (13, 8191 = rejected) | cube-local voxel key (30) | index in stack (20):
// after the sort every cube's new points are contiguous AND already in (voxel key, arrival) order, i.e.
// the order the cube's VoxelGrid refilter needs -- no per-cube sort later
constexpr int INS_IDX_BITS = 20;
__device__ __forceinline__ uint32_t d_ins_group(unsigned long long c) { return (uint32_t)(c >> (30 + INS_IDX_BITS)); }   // type << 13 | slot
__device__ __forceinline__ uint32_t d_ins_vkey(unsigned long long c) { return (uint32_t)(c >> INS_IDX_BITS) & 0x3FFFFFFFu; }
__device__ __forceinline__ uint32_t d_ins_index(unsigned long long c) { return (uint32_t)c & ((1u << INS_IDX_BITS) - 1u); }
__global__ void __launch_bounds__(256) k_insert_prepare(LmMapState* __restrict__ st, const float4* __restrict__ stack0,
                                                        const float4* __restrict__ stack1, float4* __restrict__ world0,
                                                        float4* __restrict__ world1, unsigned long long* __restrict__ comp,
                                                        int32_t* __restrict__ n_ins, float leaf0, float inv_leaf0, float leaf1, float inv_leaf1,
                                                        int transform_update, const int32_t* __restrict__ slot_valid_rank) {
  lm_pdl_enter();
  const int n0 = st->stack_n[0], n1 = st->stack_n[1];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e == 0) *n_ins = n0 + n1;
  if (e == 0 && transform_update) d_transform_update(st);     // :734 rides along (writes q/t_wmap_wodom only; nobody here reads them)
  if (e >= n0 + n1) return;
  const int ty = e < n0 ? 0 : 1;
  const int i = ty == 0 ? e : e - n0;
  float4 pw = d_associate(st->q_w_curr, st->t_w_curr, ty == 0 ? stack0[i] : stack1[i]);
  (ty == 0 ? world0 : world1)[i] = pw;
  int cI = d_cube_coord((double)pw.x, st->cen[0]);
  int cJ = d_cube_coord((double)pw.y, st->cen[1]);
  int cK = d_cube_coord((double)pw.z, st->cen[2]);
  uint32_t ps = 8191u, vkey = 0u;
  if (cI >= 0 && cI < LM_GW && cJ >= 0 && cJ < LM_GH && cK >= 0 && cK < LM_GD &&
      d_shard_keep(pw, ty == 0 ? leaf0 : leaf1, ty == 0 ? inv_leaf0 : inv_leaf1, st->shard_rank, st->shard_n)) {
    const int g3[3] = { cI - st->cen[0], cJ - st->cen[1], cK - st->cen[2] };
    ps = (uint32_t)d_phys_slot(g3[0], g3[1], g3[2]);
    // a cube outside the 5x5x3 window is not refiltered this sweep (:788-801 walks laserCloudValidInd only): the reference
    // leaves its new points appended in ARRIVAL order, and so do we (sort key without the voxel key); k_insert_heads flags
    // the slab, so it is re-voxelised as a whole once it is inside the window
    vkey = slot_valid_rank[ps] >= 0 ? d_cube_voxel_key(pw, ty == 0 ? inv_leaf0 : inv_leaf1, g3) : 0u;
  }
  comp[e] = ((unsigned long long)(((uint32_t)ty << 13) | ps) << (30 + INS_IDX_BITS)) | ((unsigned long long)vkey << INS_IDX_BITS) | (unsigned long long)i;
}

__global__ void __launch_bounds__(256) k_insert_heads(LmMapState* __restrict__ st, LmMapType M0, LmMapType M1,
                                                      const unsigned long long* __restrict__ sorted, const int32_t* __restrict__ n_ins,
                                                      const float4* __restrict__ world0, const float4* __restrict__ world1,
                                                      int32_t* __restrict__ slot_first, int32_t* __restrict__ slot_base,
                                                      int32_t* __restrict__ slot_len, const int32_t* __restrict__ slot_valid_rank) {
  lm_pdl_enter();
  const int n = *n_ins;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const unsigned long long me = sorted[p];
  const uint32_t key = d_ins_group(me);
  if (p > 0 && d_ins_group(sorted[p - 1]) == key) return;        // not a run head
  const uint32_t ps = key & 8191u;
  if (ps == 8191u) return;                                        // outside the 21x21x11 grid: dropped (:752-754)
  const int ty = (int)(key >> 13);
  const LmMapType& M = ty == 0 ? M0 : M1;
  const int run_end = d_lower_bound_u64(sorted, n, (unsigned long long)(key + 1u) << (30 + INS_IDX_BITS));
  int len = run_end - p;
  int sid = M.slot_slab[ps];
  if (sid < 0) {
    int top = atomicSub(M.free_top, 1) - 1;
    if (top < 0) { atomicAdd(M.free_top, 1); atomicOr(&st->fault, LM_FAULT_POOL_EXHAUSTED); slot_first[ty * LM_NSLOT + ps] = p; slot_len[ty * LM_NSLOT + ps] = 0; slot_base[ty * LM_NSLOT + ps] = 0; return; }
    sid = M.free_stack[top];
    M.slot_slab[ps] = sid;
    M.slab_n[sid] = 0; M.slab_nsorted[sid] = 0; M.slab_cur[sid] = 0; M.slab_unsorted[sid] = 0;
    float4 pw = (ty == 0 ? world0 : world1)[d_ins_index(me)];
    M.slab_g[sid * 4 + 0] = d_cube_coord((double)pw.x, 0);
    M.slab_g[sid * 4 + 1] = d_cube_coord((double)pw.y, 0);
    M.slab_g[sid * 4 + 2] = d_cube_coord((double)pw.z, 0);
  }
  const int base = M.slab_n[sid];
  if (base + len > M.cap) { atomicOr(&st->fault, LM_FAULT_CUBE_OVERFLOW); len = max(0, M.cap - base); }
  slot_first[ty * LM_NSLOT + ps] = p;
  slot_base[ty * LM_NSLOT + ps] = base;
  slot_len[ty * LM_NSLOT + ps] = len;
  // the tail merge needs ONE batch of new points in voxel-key order behind a filtered prefix: a cube outside the window
  // (arrival order, see k_insert_prepare) or one that still carries an earlier unmerged batch goes through the
  // whole-slab re-voxelisation instead
  if (len > 0 && (slot_valid_rank[ps] < 0 || base > M.slab_nsorted[sid])) M.slab_unsorted[sid] = 1;
  M.slab_n[sid] = base + len;
  M.slab_dirty[sid] = 1;
}

__device__ __forceinline__ void d_insert_write(const LmMapType& M0, const LmMapType& M1, const unsigned long long* __restrict__ sorted,
                                               const int32_t* __restrict__ n_ins, const float4* __restrict__ world0,
                                               const float4* __restrict__ world1, const int32_t* __restrict__ slot_first,
                                               const int32_t* __restrict__ slot_base, const int32_t* __restrict__ slot_len) {
  const int n = *n_ins;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const unsigned long long me = sorted[p];
  const uint32_t key = d_ins_group(me);
  const uint32_t ps = key & 8191u;
  if (ps == 8191u) return;
  const int ty = (int)(key >> 13);
  const LmMapType& M = ty == 0 ? M0 : M1;
  const int sid = M.slot_slab[ps];
  if (sid < 0) return;
  const int rel = p - slot_first[ty * LM_NSLOT + ps];
  if (rel >= slot_len[ty * LM_NSLOT + ps]) return;
  const size_t buf = ((size_t)sid * 2 + M.slab_cur[sid]) * M.cap + slot_base[ty * LM_NSLOT + ps] + rel;
  M.pts[buf] = (ty == 0 ? world0 : world1)[d_ins_index(me)];
  M.pkey[buf] = d_ins_vkey(me);
}

// ------------------------------------------------------------------ refilter (:788-801)
// slab = sorted voxel-unique prefix [0,ns) + tail [ns,n) of newly inserted points, already in
// (voxel key, arrival) order (the insertion sort key carries the voxel key).
// VoxelGrid(prefix ++ tail) == merge(prefix, tail): a voxel's members are the prefix point (if any)
// followed by the tail points in arrival order, summed in fp32 in that order and divided by (float)count;
// a VoxelGrid pass over the untouched part of a filtered cube is the identity.
// The merge runs as four small kernels over (window cube, map type[, 2048-point chunk]) so that the
// ~25 k points of a busy cube are spread over a dozen SMs instead of serialised on one:
//   k_rf_tailflags per tail point: does its key run open a NEW voxel (binary search in the prefix keys)?
//   k_rf_flags   per cube: exclusive scan of those flags -> output offsets; clears the cell histogram
//   k_rf_merge   per chunk: prefix points shift by the new voxels sorting before them and absorb their tail
//                run; new voxels are centroided; every output also counts into its 2 m search cell
//   k_rf_scan    per cube: histogram -> cell starts, slab bookkeeping (ping-pong flip)
//   k_rf_scatter per chunk: cell-sorted copy for the kNN search
// A centroid that rounds across a voxel border breaks the "prefix is sorted" invariant (PCL would simply
// re-voxelise): such a slab is flagged and re-voxelised as a whole by k_refilter_whole on the next pass.
struct RfMeta { int32_t active, total_new, ns, nt, flag, cur, sid, pad; };
// chunk work list of the active cubes: entry = type << 24 | window rank << 16 | chunk; [0] of the counter array = length
static const int RF_GRID = getenv("LMONO_RF_GRID") ? atoi(getenv("LMONO_RF_GRID")) : 592;   // experiment switch

__device__ __forceinline__ bool d_rf_slab(const LmMapState* st, const LmMapType& M, int r, int* sid) {
  if (r >= st->valid_num) return false;
  const int s = M.slot_slab[st->valid_slot[r]];
  *sid = s;
  return s >= 0;
}

// per step: classify the window cubes once.  whole list = flagged slabs (re-voxelised as a whole), active list = slabs
// with a tail (their meta record is armed here: ns, nt, cur, sid).  One CTA.
__device__ __forceinline__ void d_rf_plan(LmMapState* __restrict__ st, const LmMapType& M0, const LmMapType& M1, int32_t* __restrict__ plan,
                                          RfMeta* __restrict__ meta_all, int32_t* __restrict__ work_n, int* ws /* shared int[33] */) {
  const int e = threadIdx.x;
  const int ty = e / LM_WIN_MAX, r = e - ty * LM_WIN_MAX;
  int sid = -1, n = 0, ns = 0, cur = 0;
  bool whole = false, active = false;
  if (e < 2 * LM_WIN_MAX) {
    const LmMapType& M = ty == 0 ? M0 : M1;
    if (d_rf_slab(st, M, r, &sid)) {
      n = M.slab_n[sid]; ns = M.slab_nsorted[sid]; cur = M.slab_cur[sid];
      const bool uns = M.slab_unsorted[sid] != 0;
      whole = uns && n > 0;
      active = !uns && n - ns > 0;
    }
    RfMeta* meta = meta_all + e;
    meta->active = active ? 1 : 0;
    if (active) { meta->total_new = 0; meta->ns = ns; meta->nt = n - ns; meta->flag = 0; meta->cur = cur; meta->sid = sid; }
  }
  int tot;
  const int wpos = d_block_exscan(whole ? 1 : 0, ws, &tot);
  if (whole) plan[LM_PLAN_WHOLE + wpos] = e;
  if (threadIdx.x == 0) plan[LM_PLAN_WHOLE_N] = tot;
  const int apos = d_block_exscan(active ? 1 : 0, ws, &tot);
  if (active) plan[LM_PLAN_ACTIVE + apos] = e;
  if (threadIdx.x == 0) plan[LM_PLAN_ACTIVE_N] = tot;
  if (threadIdx.x == 0) *work_n = 0;
}

__global__ void __launch_bounds__(256) k_rf_plan(LmMapState* __restrict__ st, LmMapType M0, LmMapType M1, int32_t* __restrict__ plan,
                                                 RfMeta* __restrict__ meta_all, int32_t* __restrict__ work_n) {
  lm_pdl_enter();
  __shared__ int ws[33];
  d_rf_plan(st, M0, M1, plan, meta_all, work_n, ws);
}

// k_insert_write + the refilter plan in one launch: the extra last CTA plans while the others copy the new points (the
// plan only reads what k_insert_heads finalised)
__global__ void __launch_bounds__(256) k_insert_write_plan(LmMapState* __restrict__ st, LmMapType M0, LmMapType M1, const unsigned long long* __restrict__ sorted,
                                                           const int32_t* __restrict__ n_ins, const float4* __restrict__ world0,
                                                           const float4* __restrict__ world1, const int32_t* __restrict__ slot_first,
                                                           const int32_t* __restrict__ slot_base, const int32_t* __restrict__ slot_len,
                                                           int32_t* __restrict__ plan, RfMeta* __restrict__ meta_all, int32_t* __restrict__ work_n) {
  lm_pdl_enter();
  __shared__ int ws[33];
  if (blockIdx.x == gridDim.x - 1) { d_rf_plan(st, M0, M1, plan, meta_all, work_n, ws); return; }
  d_insert_write(M0, M1, sorted, n_ins, world0, world1, slot_first, slot_base, slot_len);
}

// per cube with a tail, one CTA: tail element j opens a NEW voxel iff it is the first of its key run and the key is absent
// from the prefix (9-ary search, the lower bound is kept for k_rf_merge); exclusive scan of those flags -> output
// offsets.  A cube without a new voxel is finished here (centroids updated in place, see below); the others get their
// cell histogram cleared and their merge chunks appended to the work list.
constexpr int RF_ACT_GRID = 80;
constexpr int RF_TS_THREADS = 1024;
static_assert(LM_NCELL % 4 == 0, "vector clear of the cell histogram");
static_assert(LM_RF_CHUNK % 256 == 0, "k_rf_merge / k_rf_scatter: whole elements per thread");
__global__ void __launch_bounds__(RF_TS_THREADS) k_rf_tailscan(LmMapState* __restrict__ st, LmMapType M0, LmMapType M1, const int32_t* __restrict__ plan,
                                                               int32_t* __restrict__ nvx_all, int32_t* __restrict__ tlb_all, RfMeta* __restrict__ meta_all,
                                                               int nvx_stride, int32_t* __restrict__ work_n, int32_t* __restrict__ work, int inplace) {
  lm_pdl_enter();
  __shared__ int ws[33];
  __shared__ int s_wbase, s_mover, s_badkey;
  const int na = plan[LM_PLAN_ACTIVE_N];
  for (int a = blockIdx.x; a < na; a += gridDim.x) {
    const int e = plan[LM_PLAN_ACTIVE + a];
    const int ty = e / LM_WIN_MAX;
    const LmMapType& M = ty == 0 ? M0 : M1;
    RfMeta* meta = meta_all + e;
    const int ns = meta->ns, nt = meta->nt, sid = meta->sid;
    const uint32_t* __restrict__ pkey = M.pkey + ((size_t)sid * 2 + meta->cur) * M.cap;
    const uint32_t* __restrict__ tkey = pkey + ns;
    int32_t* __restrict__ nvx = nvx_all + (size_t)e * nvx_stride;
    int32_t* __restrict__ tlb = tlb_all + (size_t)e * nvx_stride;
    int carry = 0;
    for (int j0 = 0; j0 < nt; j0 += blockDim.x) {
      const int j = j0 + threadIdx.x;
      int nv = 0;
      if (j < nt) {
        const uint32_t key = tkey[j];
        if (j == 0 || tkey[j - 1] != key) {
          const int lb = d_lower_bound_u32_wide(pkey, ns, key);
          nv = !(lb < ns && pkey[lb] == key);
          tlb[j] = lb;
        }
      }
      int tot;
      const int ex = d_block_exscan(nv, ws, &tot);
      if (j < nt) nvx[j] = ((carry + ex) << 1) | nv;
      carry += tot;
    }
    // ---- no new voxel in this cube (the usual case once an area is mapped): VoxelGrid(prefix ++ tail) leaves every
    // point where it is and only moves the centroids of the voxels the tail hits.  They are updated IN PLACE instead of
    // rewriting the whole slab (points, keys, cell table, cell-sorted copy: ~70 B per stored point) through merge / scan /
    // scatter.  Same summation order as k_rf_merge (prefix point, then the tail members in arrival order, one division),
    // so the bits are the same.  The cell-sorted copy is patched too (the entry is looked up in its 2 m cell) unless some
    // centroid changed its cell (a 0.8 m surf voxel can straddle a cell border): then the copy stays stale, the slab
    // stays `dirty`, and k_index_build rebuilds that cube's index at the start of the next step -- beside the feature
    // VoxelGrid, off the critical path -- as it does for imported cubes.
    if (inplace && carry == 0 && nt > 0) {
      float4* pts = M.pts + ((size_t)sid * 2 + meta->cur) * M.cap;
      const int g3[3] = { M.slab_g[sid * 4], M.slab_g[sid * 4 + 1], M.slab_g[sid * 4 + 2] };
      const float il = M.inv_leaf;
      if (threadIdx.x == 0) { s_mover = 0; s_badkey = 0; }
      __syncthreads();
      // pass 1: every hit centroid is recomputed and written to the canonical buffer (that update is unconditional); a
      // centroid that left its 2 m cell or its voxel is noted.  pass 2 (only if no centroid changed cell): the entries of
      // the cell-sorted copy are patched.  A tail of up to blockDim points (the usual case) keeps its centroids in
      // registers between the passes; a longer one recomputes them from the tail (the old prefix point is gone).
      float4* cp = M.cellpts + (size_t)sid * M.cap;
      const bool one_trip = nt <= (int)blockDim.x;
      float4 pn_keep = make_float4(0.f, 0.f, 0.f, 0.f); int cell_keep = -1, lb_keep = -1;
      for (int j = threadIdx.x; j < nt; j += blockDim.x) {
        const uint32_t key = tkey[j];
        if (j > 0 && tkey[j - 1] == key) continue;           // not a run head
        const int lb = tlb[j];
        const float4 po = pts[lb];
        float sx = po.x, sy = po.y, sz = po.z, si = po.w;
        int cnt = 1;
        for (int m = j; m < nt && tkey[m] == key; ++m) {
          const float4 t = pts[ns + m];
          sx = __fadd_rn(sx, t.x); sy = __fadd_rn(sy, t.y); sz = __fadd_rn(sz, t.z); si = __fadd_rn(si, t.w);
          ++cnt;
        }
        const float c = (float)cnt;
        const float4 pn = make_float4(__fdiv_rn(sx, c), __fdiv_rn(sy, c), __fdiv_rn(sz, c), __fdiv_rn(si, c));
        const int cell = d_cube_cell(po, g3);
        if (cell < 0 || d_cube_cell(pn, g3) != cell) s_mover = 1;
        if (d_cube_voxel_key(pn, il, g3) != key) s_badkey = 1;          // the centroid left its voxel: re-voxelise the cube as a whole next time
        pts[lb] = pn;
        pn_keep = pn; cell_keep = cell; lb_keep = lb;
      }
      __syncthreads();
      if (!s_mover) {
        for (int j = threadIdx.x; j < nt; j += blockDim.x) {
          float4 pn; int cell, lb;
          if (one_trip) { if (lb_keep < 0) continue; pn = pn_keep; cell = cell_keep; lb = lb_keep; }
          else {
            const uint32_t key = tkey[j];
            if (j > 0 && tkey[j - 1] == key) continue;
            lb = tlb[j];
            pn = pts[lb];                                      // written by this thread in pass 1
            cell = d_cube_cell(pn, g3);                        // no centroid changed cell
          }
          const uint32_t* cs = M.cellstart + (size_t)sid * (LM_NCELL + 1) + cell;
          const int cb = (int)cs[0], ce = (int)cs[1];
          bool found = false;
          for (int q0 = cb; q0 < ce && !found; q0 += 4) {      // the cell holds this point exactly once
            float w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) w[u] = q0 + u < ce ? cp[q0 + u].w : __int_as_float(-1);
#pragma unroll
            for (int u = 0; u < 4; ++u) if (__float_as_int(w[u]) == lb) { cp[q0 + u] = make_float4(pn.x, pn.y, pn.z, w[u]); found = true; }
          }
          if (!found) atomicOr(&st->fault, LM_FAULT_CELL_RANGE);      // the search index did not cover the prefix (internal error)
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        M.slab_n[sid] = ns;
        M.slab_nsorted[sid] = s_badkey ? 0 : ns;
        M.slab_unsorted[sid] = s_badkey ? 1 : 0;
        M.slab_dirty[sid] = s_mover ? 1 : 0;
        meta->active = s_mover ? 3 : 2;                        // done: k_rf_scan skips it, no merge chunks (3: index rebuild pending)
        meta->total_new = ns;
      }
      __syncthreads();
      continue;
    }
    int4* cc4 = reinterpret_cast<int4*>(M.cellcount + (size_t)sid * LM_NCELL);       // LM_NCELL % 4 == 0, slabs 16 B aligned
    for (int c = threadIdx.x; c < LM_NCELL / 4; c += blockDim.x) cc4[c] = make_int4(0, 0, 0, 0);
    const int nch = (ns + nt + LM_RF_CHUNK - 1) / LM_RF_CHUNK;
    if (threadIdx.x == 0) {
      if (ns + carry > M.cap) atomicOr(&st->fault, LM_FAULT_CUBE_OVERFLOW);
      meta->total_new = carry;
      s_wbase = atomicAdd(work_n, nch);
    }
    __syncthreads();
    const int r = e - ty * LM_WIN_MAX;
    for (int c = threadIdx.x; c < nch; c += blockDim.x) work[s_wbase + c] = (ty << 24) | (r << 16) | c;
    __syncthreads();
  }
}

__device__ __forceinline__ void d_rf_emit(const LmMapType& M, int sid, int cur, int pos, float4 p, uint32_t key, const int* g3, LmMapState* st) {
  if (pos >= M.cap) return;
  const size_t o = ((size_t)sid * 2 + (cur ^ 1)) * M.cap + pos;
  M.pts[o] = p;
  M.pkey[o] = key;
  int c = d_cube_cell(p, g3);
  if (c < 0) { atomicOr(&st->fault, LM_FAULT_CELL_RANGE); c = 0; }
  atomicAdd(&M.cellcount[(size_t)sid * LM_NCELL + c], 1);
}

constexpr int RF_STAGE = 2048;        // tail keys of a cube staged in shared memory by k_rf_merge (larger tails are searched in global memory)
__global__ void __launch_bounds__(256) k_rf_merge(LmMapState* __restrict__ st, LmMapType M0, LmMapType M1,
                                                  const int32_t* __restrict__ nvx_all, const int32_t* __restrict__ tlb_all, RfMeta* __restrict__ meta_all, int nvx_stride,
                                                  const int32_t* __restrict__ work_n, const int32_t* __restrict__ work) {
  lm_pdl_enter();
 __shared__ uint32_t s_tkey[RF_STAGE];
 const int nwork = *work_n;
 for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
  const int we = work[w];
  const int ty = we >> 24, r = (we >> 16) & 0xFF;
  RfMeta* meta = meta_all + ty * LM_WIN_MAX + r;
  const int ns = meta->ns, nt = meta->nt, total_new = meta->total_new, cur = meta->cur, sid = meta->sid;
  const int e0 = (we & 0xFFFF) * LM_RF_CHUNK;
  if (e0 >= ns + nt) continue;
  const LmMapType& M = ty == 0 ? M0 : M1;
  const float4* __restrict__ src = M.pts + ((size_t)sid * 2 + cur) * M.cap;
  const uint32_t* __restrict__ pkey = M.pkey + ((size_t)sid * 2 + cur) * M.cap;
  const uint32_t* __restrict__ tkey = pkey + ns;
  const int32_t* __restrict__ nvx = nvx_all + (size_t)(ty * LM_WIN_MAX + r) * nvx_stride;
  const int32_t* __restrict__ tlb = tlb_all + (size_t)(ty * LM_WIN_MAX + r) * nvx_stride;
  const int g3[3] = { M.slab_g[sid * 4], M.slab_g[sid * 4 + 1], M.slab_g[sid * 4 + 2] };
  const float il = M.inv_leaf;
  // a chunk of prefix points looks its keys up in the cube's tail: stage the (few hundred) tail keys once, search in
  // shared memory -- the per-point binary search in global memory was a chain of ~10 dependent L2 / HBM round trips
  const bool staged = e0 < ns && nt <= RF_STAGE;
  __syncthreads();                      // the previous chunk's readers are done with s_tkey
  if (staged) for (int j = threadIdx.x; j < nt; j += blockDim.x) s_tkey[j] = tkey[j];
  __syncthreads();
  int bad = 0;
  // LM_RF_CHUNK / blockDim elements per thread.  The output buffer is the other half of the array the inputs live in, so
  // the compiler must assume every store aliases the next load: gather the inputs of all of a thread's elements first
  // (independent loads in flight together), then merge and emit.
  constexpr int PER = LM_RF_CHUNK / 256;
  const int eend = min(e0 + LM_RF_CHUNK, ns + nt);
  uint32_t key_[PER]; float4 p_[PER]; int aux_[PER];       // prefix: aux = lower bound in the tail; tail: aux = nvx flag word
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    const int e = e0 + threadIdx.x + u * 256;
    key_[u] = 0; aux_[u] = 0; p_[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e < eend) {
      key_[u] = pkey[e];              // pkey and tkey are one array: tkey[j] = pkey[ns + j]
      if (e < ns) p_[u] = src[e]; else aux_[u] = nvx[e - ns];
    }
  }
  int before_[PER];
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    const int e = e0 + threadIdx.x + u * 256;
    before_[u] = 0;
    if (e < eend && e < ns) {
      const uint32_t key = key_[u];
      int lb;
      if (staged) {
        int lo = 0, hi = nt;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_tkey[mid] < key) lo = mid + 1; else hi = mid; }
        lb = lo;
      } else {
        lb = d_lower_bound_u32(tkey, nt, key);
      }
      aux_[u] = lb;
      before_[u] = lb < nt ? (nvx[lb] >> 1) : total_new;
    } else if (e < eend && (aux_[u] & 1)) {
      before_[u] = tlb[e - ns];       // lower bound of the key in the prefix (k_rf_tailflags)
    }
  }
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    const int e = e0 + threadIdx.x + u * 256;
    if (e >= eend) continue;
    const uint32_t key = key_[u];
    if (e < ns) {
      // prefix point: shifted by the number of new voxels sorting before it; merged if the tail hits its voxel
      const int lb = aux_[u];
      float4 p = p_[u];
      const bool hit = lb < nt && (staged ? s_tkey[lb] : tkey[lb]) == key;
      if (hit) {
        float sx = p.x, sy = p.y, sz = p.z, si = p.w;
        int cnt = 1;
        for (int m = lb; m < nt && tkey[m] == key; ++m) {
          const float4 t = src[ns + m];
          sx = __fadd_rn(sx, t.x); sy = __fadd_rn(sy, t.y); sz = __fadd_rn(sz, t.z); si = __fadd_rn(si, t.w);
          ++cnt;
        }
        const float c = (float)cnt;
        p = make_float4(__fdiv_rn(sx, c), __fdiv_rn(sy, c), __fdiv_rn(sz, c), __fdiv_rn(si, c));
        bad |= d_cube_voxel_key(p, il, g3) != key;      // the centroid left its voxel
      }
      d_rf_emit(M, sid, cur, e + before_[u], p, key, g3, st);
    } else {
      const int j = e - ns;
      const int v = aux_[u];
      if (!(v & 1)) continue;
      float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
      int cnt = 0;
      for (int m = j; m < nt && tkey[m] == key; ++m) {
        const float4 t = src[ns + m];
        sx = __fadd_rn(sx, t.x); sy = __fadd_rn(sy, t.y); sz = __fadd_rn(sz, t.z); si = __fadd_rn(si, t.w);
        ++cnt;
      }
      const float c = (float)cnt;
      const float4 p = make_float4(__fdiv_rn(sx, c), __fdiv_rn(sy, c), __fdiv_rn(sz, c), __fdiv_rn(si, c));
      bad |= d_cube_voxel_key(p, il, g3) != key;
      d_rf_emit(M, sid, cur, before_[u] + (v >> 1), p, key, g3, st);
    }
  }
  if (bad) meta->flag = 1;
 }
}

constexpr int RF_SCAN_THREADS = 512;
__global__ void __launch_bounds__(RF_SCAN_THREADS) k_rf_scan(LmMapType M0, LmMapType M1, const int32_t* __restrict__ plan, RfMeta* __restrict__ meta_all,
                                                             const int32_t* __restrict__ work_n) {
  lm_pdl_enter();
  if (*work_n == 0) return;                          // every cube was updated in place: nothing was merged
  constexpr int NW = RF_SCAN_THREADS / 32;
  __shared__ int wtot[NW], wbase[NW];
  const int na = plan[LM_PLAN_ACTIVE_N];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // each warp owns a contiguous segment of the 17576 cells and walks it 32 cells at a time: all loads of the segment
  // are issued up front (values kept in registers), then scanned with a running carry
  constexpr int SEG = (LM_NCELL + NW - 1) / NW;      // 1099 cells per warp
  constexpr int ITERS = (SEG + 31) / 32;             // 35
  for (int a = blockIdx.x; a < na; a += gridDim.x) {
    const int e = plan[LM_PLAN_ACTIVE + a];
    const LmMapType& M = e < LM_WIN_MAX ? M0 : M1;
    RfMeta* meta = meta_all + e;
    if (meta->active != 1) continue;                 // updated in place by k_rf_tailscan
    const int sid = meta->sid;
    int32_t* cc = M.cellcount + (size_t)sid * LM_NCELL;
    uint32_t* cs = M.cellstart + (size_t)sid * (LM_NCELL + 1);
    const int seg0 = wid * SEG, seg1 = min(seg0 + SEG, LM_NCELL);
    int vals[ITERS];
#pragma unroll
    for (int it = 0; it < ITERS; ++it) { const int c = seg0 + it * 32 + lane; vals[it] = c < seg1 ? cc[c] : 0; }
    int carry = 0;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      int incl = vals[it];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
      const int tot = __shfl_sync(0xffffffffu, incl, 31);
      vals[it] = carry + incl - vals[it];               // exclusive within the warp segment
      carry += tot;
    }
    if (lane == 0) wtot[wid] = carry;
    __syncthreads();
    if (wid == 0) {
      const int v = lane < NW ? wtot[lane] : 0;
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
      if (lane < NW) wbase[lane] = incl - v;
    }
    __syncthreads();
    const int base = wbase[wid];
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int c = seg0 + it * 32 + lane;
      if (c < seg1) { const int v = base + vals[it]; cs[c] = (uint32_t)v; cc[c] = v; }
    }
    if (threadIdx.x == 0) {
      const int nn = min(meta->ns + meta->total_new, M.cap);
      cs[LM_NCELL] = (uint32_t)nn;
      M.slab_n[sid] = nn;
      M.slab_nsorted[sid] = meta->flag ? 0 : nn;
      M.slab_unsorted[sid] = meta->flag ? 1 : 0;
      M.slab_cur[sid] = meta->cur ^ 1;
      M.slab_dirty[sid] = 0;
      meta->total_new = nn;         // k_rf_scatter reads the new size here
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_rf_scatter(LmMapType M0, LmMapType M1, const RfMeta* __restrict__ meta_all,
                                                    const int32_t* __restrict__ work_n, const int32_t* __restrict__ work) {
  lm_pdl_enter();
 const int nwork = *work_n;
 for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
  const int we = work[w];
  const int ty = we >> 24, r = (we >> 16) & 0xFF;
  const RfMeta* meta = meta_all + ty * LM_WIN_MAX + r;
  const int nn = meta->total_new, sid = meta->sid;
  const int e0 = (we & 0xFFFF) * LM_RF_CHUNK;
  if (e0 >= nn) continue;
  const LmMapType& M = ty == 0 ? M0 : M1;
  const float4* __restrict__ src = M.pts + ((size_t)sid * 2 + (meta->cur ^ 1)) * M.cap;
  float4* __restrict__ cp = M.cellpts + (size_t)sid * M.cap;
  int32_t* cc = M.cellcount + (size_t)sid * LM_NCELL;
  const int g3[3] = { M.slab_g[sid * 4], M.slab_g[sid * 4 + 1], M.slab_g[sid * 4 + 2] };
  constexpr int PER = LM_RF_CHUNK / 256;
  const int iend = min(e0 + LM_RF_CHUNK, nn);
  float4 p_[PER]; int pos_[PER];
#pragma unroll
  for (int u = 0; u < PER; ++u) { const int i = e0 + threadIdx.x + u * 256; if (i < iend) p_[u] = src[i]; }
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    const int i = e0 + threadIdx.x + u * 256;
    if (i < iend) { int c = d_cube_cell(p_[u], g3); if (c < 0) c = 0; pos_[u] = atomicAdd(&cc[c], 1); }
  }
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    const int i = e0 + threadIdx.x + u * 256;
    if (i < iend) { float4 p = p_[u]; p.w = __int_as_float(i); cp[pos_[u]] = p; }
  }
 }
}

// Whole-slab re-voxelisation of a flagged slab (rare): one CTA per window cube and map type sorts every
// point of the slab by voxel key in shared memory and merges runs -- exactly pcl::VoxelGrid on the cube.
constexpr int RF_WHOLE_GRID = 8;
constexpr int RF_BIG_TILE = 65536;          // slabs above LM_TAIL_TILE points sort in a per-CTA global-memory scratch (slow, rare)
__global__ void __launch_bounds__(1024, 1) k_refilter_whole(LmMapState* __restrict__ st, LmMapType M0, LmMapType M1, const int32_t* __restrict__ plan,
                                                            unsigned long long* __restrict__ big_s, int32_t* __restrict__ big_nv) {
  lm_pdl_enter();
  extern __shared__ unsigned char smem_raw[];
  unsigned long long* S_sm = reinterpret_cast<unsigned long long*>(smem_raw);           // [LM_TAIL_TILE]
  int* NV_sm = reinterpret_cast<int*>(smem_raw + (size_t)LM_TAIL_TILE * 8);             // [LM_TAIL_TILE]
  int* ws = reinterpret_cast<int*>(smem_raw + (size_t)LM_TAIL_TILE * 12);               // [64]
  __shared__ int s_flag;
  const int nw = plan[LM_PLAN_WHOLE_N];
  for (int w = blockIdx.x; w < nw; w += gridDim.x) {
  const int e_ = plan[LM_PLAN_WHOLE + w];
  const LmMapType& M = e_ < LM_WIN_MAX ? M0 : M1;
  const int r = e_ < LM_WIN_MAX ? e_ : e_ - LM_WIN_MAX;
  int sid;
  if (!d_rf_slab(st, M, r, &sid)) continue;
  if (!M.slab_unsorted[sid]) continue;
  const int nt = M.slab_n[sid];
  if (nt == 0) continue;
  if (nt > RF_BIG_TILE) { if (threadIdx.x == 0) atomicOr(&st->fault, LM_FAULT_TAIL_OVERFLOW); continue; }
  // the sort keys and the run flags of a slab live in shared memory; a slab too large for that (a full surf cube that
  // has to be re-voxelised as a whole) uses this CTA's slice of a global-memory scratch: same code, ~100x slower, rare
  const bool big = nt > LM_TAIL_TILE;
  unsigned long long* S = big ? big_s + (size_t)blockIdx.x * RF_BIG_TILE : S_sm;
  int* NV = big ? big_nv + (size_t)blockIdx.x * RF_BIG_TILE : NV_sm;
  const int cur = M.slab_cur[sid];
  const float4* src = M.pts + ((size_t)sid * 2 + cur) * M.cap;
  float4* dst = M.pts + ((size_t)sid * 2 + (cur ^ 1)) * M.cap;
  uint32_t* dkey = M.pkey + ((size_t)sid * 2 + (cur ^ 1)) * M.cap;
  const int g3[3] = { M.slab_g[sid * 4], M.slab_g[sid * 4 + 1], M.slab_g[sid * 4 + 2] };
  const float il = M.inv_leaf;
  int np2 = 1; while (np2 < nt) np2 <<= 1;
  for (int j = threadIdx.x; j < np2; j += blockDim.x)
    S[j] = j < nt ? (((unsigned long long)d_cube_voxel_key(src[j], il, g3) << 32) | (uint32_t)j) : ~0ULL;
  __syncthreads();
  d_bitonic_sort(S, np2);
  const int per = (nt + blockDim.x - 1) / blockDim.x;
  const int b = min((int)threadIdx.x * per, nt), e = min(b + per, nt);
  int local = 0;
  for (int j = b; j < e; ++j) {
    const int nv = (j == 0) || ((uint32_t)(S[j - 1] >> 32) != (uint32_t)(S[j] >> 32));
    NV[j] = nv; local += nv;
  }
  int total_new;
  int run = d_block_exscan(local, ws, &total_new);
  for (int j = b; j < e; ++j) { const int nv = NV[j]; NV[j] = (run << 1) | nv; run += nv; }
  if (threadIdx.x == 0) s_flag = 0;
  __syncthreads();
  for (int j = threadIdx.x; j < nt; j += blockDim.x) {
    const int v = NV[j];
    if (!(v & 1)) continue;
    const uint32_t key = (uint32_t)(S[j] >> 32);
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    int cnt = 0;
    for (int m = j; m < nt && (uint32_t)(S[m] >> 32) == key; ++m) {
      const float4 t = src[(uint32_t)S[m]];
      sx = __fadd_rn(sx, t.x); sy = __fadd_rn(sy, t.y); sz = __fadd_rn(sz, t.z); si = __fadd_rn(si, t.w);
      ++cnt;
    }
    const float c = (float)cnt;
    const float4 p = make_float4(__fdiv_rn(sx, c), __fdiv_rn(sy, c), __fdiv_rn(sz, c), __fdiv_rn(si, c));
    if (d_cube_voxel_key(p, il, g3) != key) s_flag = 1;
    dst[v >> 1] = p; dkey[v >> 1] = key;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    M.slab_n[sid] = total_new;
    M.slab_nsorted[sid] = s_flag ? 0 : total_new;
    M.slab_unsorted[sid] = s_flag ? 1 : 0;
    M.slab_cur[sid] = cur ^ 1;
  }
  __syncthreads();
  // rebuild the search index of this cube from the new buffer (reuses the sort scratch)
  d_build_cell_index(M, sid, dst, total_new, reinterpret_cast<uint32_t*>(smem_raw), ws, st);
  if (threadIdx.x == 0) M.slab_dirty[sid] = 0;
  __syncthreads();
  }
}

static const int kRefilterSmem = LM_TAIL_TILE * 12 + 64 * 4;

int lm_map_insert_and_refilter(lmono_ctx* ctx, int n_max_corner, int n_max_surf, bool transform_update) {
  const int n_max = n_max_corner + n_max_surf;
  lm_prof_begin(ctx, LM_PROF_INSERT);
  if (n_max > 0) {
    int32_t* n_ins = ctx->d_tmp_i32;                 // [0]
    int32_t* slot_len = ctx->d_tmp_i32 + 16;         // [2*LM_NSLOT]
    const int blocks = lm_div_up(n_max, 256);
    LM_LAUNCH_PDL(k_insert_prepare, blocks, 256, 0, ctx->d_state, ctx->d_stack[0], ctx->d_stack[1], ctx->d_world[0],
                                                      ctx->d_world[1], ctx->d_sort_a, n_ins, ctx->map[0].leaf, ctx->map[0].inv_leaf,
                                                      ctx->map[1].leaf, ctx->map[1].inv_leaf, transform_update ? 1 : 0, ctx->d_slot_valid_rank);
    LM_LAUNCH_CHECK();
    int rc = lm_sort_u64(ctx, ctx->d_sort_a, ctx->d_sort_b, ctx->d_sort_c, n_ins, n_max);
    if (rc) return rc;
    LM_LAUNCH_PDL(k_insert_heads, blocks, 256, 0, ctx->d_state, ctx->map[0], ctx->map[1], ctx->d_sort_c, n_ins,
                                                    ctx->d_world[0], ctx->d_world[1], ctx->d_slot_first, ctx->d_slot_base, slot_len, ctx->d_slot_valid_rank);
    LM_LAUNCH_CHECK();
    LM_LAUNCH_PDL(k_insert_write_plan, blocks + 1, 256, 0, ctx->d_state, ctx->map[0], ctx->map[1], ctx->d_sort_c, n_ins, ctx->d_world[0],
                                                             ctx->d_world[1], ctx->d_slot_first, ctx->d_slot_base, slot_len,
                                                             ctx->d_rf_plan, (RfMeta*)ctx->d_rf_meta, ctx->d_rf_work);
    LM_LAUNCH_CHECK();
  }
  lm_prof_end(ctx);
  lm_prof_begin(ctx, LM_PROF_REFILTER);
  const int cap_max = ctx->map[0].cap > ctx->map[1].cap ? ctx->map[0].cap : ctx->map[1].cap;
  RfMeta* meta = (RfMeta*)ctx->d_rf_meta;
  int32_t* work_n = ctx->d_rf_work; int32_t* work = ctx->d_rf_work + 4;
  static const bool rf_inplace = !(getenv("LMONO_RF_INPLACE") && getenv("LMONO_RF_INPLACE")[0] == '0');      // A/B switch
  if (n_max <= 0) {
    LM_LAUNCH_PDL(k_rf_plan, 1, 256, 0, ctx->d_state, ctx->map[0], ctx->map[1], ctx->d_rf_plan, meta, work_n);
    LM_LAUNCH_CHECK();
  }
  // flagged slabs (rare) are re-voxelised as a whole beside the tail merge of the others: disjoint slabs, so inside a
  // graph capture the kernel is a parallel branch
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(ctx->stream, &cap);
  const bool fork = cap == cudaStreamCaptureStatusActive && ctx->side_stream != nullptr && !ctx->tl_on;
  cudaStream_t main_stream = ctx->stream;
  if (fork) {
    LM_CUDA(cudaEventRecord(ctx->ev_side0, main_stream));
    LM_CUDA(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_side0, 0));
    ctx->stream = ctx->side_stream;
  }
  LM_LAUNCH_PDL(k_refilter_whole, RF_WHOLE_GRID, 1024, kRefilterSmem, ctx->d_state, ctx->map[0], ctx->map[1], ctx->d_rf_plan, ctx->d_rf_big_s, ctx->d_rf_big_nv);
  if (fork) { ctx->stream = main_stream; LM_CUDA(cudaEventRecord(ctx->ev_side1, ctx->side_stream)); }
  LM_LAUNCH_CHECK();
  LM_LAUNCH_PDL(k_rf_tailscan, RF_ACT_GRID, RF_TS_THREADS, 0, ctx->d_state, ctx->map[0], ctx->map[1], ctx->d_rf_plan, ctx->d_rf_nvx, ctx->d_rf_tlb, meta, cap_max, work_n, work, rf_inplace ? 1 : 0);
  LM_LAUNCH_CHECK();
  LM_LAUNCH_PDL(k_rf_merge, RF_GRID, 256, 0, ctx->d_state, ctx->map[0], ctx->map[1], ctx->d_rf_nvx, ctx->d_rf_tlb, meta, cap_max, work_n, work);
  LM_LAUNCH_CHECK();
  LM_LAUNCH_PDL(k_rf_scan, RF_ACT_GRID, RF_SCAN_THREADS, 0, ctx->map[0], ctx->map[1], ctx->d_rf_plan, meta, work_n);
  LM_LAUNCH_CHECK();
  LM_LAUNCH_PDL(k_rf_scatter, RF_GRID, 256, 0, ctx->map[0], ctx->map[1], meta, work_n, work);
  LM_LAUNCH_CHECK();
  if (fork) LM_CUDA(cudaStreamWaitEvent(main_stream, ctx->ev_side1, 0));
  lm_prof_end(ctx);
  return LMONO_OK;
}

// study hook: the refilter records of the last step, [2][LM_WIN_MAX] x {active, new size, prefix size, tail size, flag, cur, slab, -}
extern "C" int lmono_debug_rf_meta(lmono_ctx* ctx, int32_t* out, int32_t n_ints) {
  if (ctx && !ctx->map_ready) return LMONO_E_STATE;
  if (!ctx || !out || n_ints < 0 || n_ints > (int)(sizeof(RfMeta) / 4) * 2 * LM_WIN_MAX) return LMONO_E_ARG;
  LM_CUDA(cudaMemcpyAsync(out, ctx->d_rf_meta, sizeof(int32_t) * (size_t)n_ints, cudaMemcpyDeviceToHost, ctx->stream));
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  return LMONO_OK;
}

int lm_map_configure_kernels(lmono_ctx* ctx) {
  LM_CUDA(cudaFuncSetAttribute(k_index_build, cudaFuncAttributeMaxDynamicSharedMemorySize, kIndexSmem));
  LM_CUDA(cudaFuncSetAttribute(k_refilter_whole, cudaFuncAttributeMaxDynamicSharedMemorySize, kRefilterSmem));
  const int cap_max = ctx->map[0].cap > ctx->map[1].cap ? ctx->map[0].cap : ctx->map[1].cap;
  LM_CUDA(cudaMalloc((void**)&ctx->d_rf_nvx, sizeof(int32_t) * (size_t)2 * LM_WIN_MAX * cap_max));
  LM_CUDA(cudaMalloc((void**)&ctx->d_rf_tlb, sizeof(int32_t) * (size_t)2 * LM_WIN_MAX * cap_max));
  LM_CUDA(cudaMalloc((void**)&ctx->d_rf_meta, sizeof(RfMeta) * 2 * LM_WIN_MAX));
  LM_CUDA(cudaMalloc((void**)&ctx->d_rf_work, sizeof(int32_t) * (4 + (size_t)2 * LM_WIN_MAX * (lm_div_up(cap_max, LM_RF_CHUNK) + 1))));
  LM_CUDA(cudaMemsetAsync(ctx->d_rf_meta, 0, sizeof(RfMeta) * 2 * LM_WIN_MAX, ctx->stream));
  LM_CUDA(cudaMalloc((void**)&ctx->d_rf_big_s, sizeof(unsigned long long) * (size_t)RF_WHOLE_GRID * RF_BIG_TILE));
  LM_CUDA(cudaMalloc((void**)&ctx->d_rf_big_nv, sizeof(int32_t) * (size_t)RF_WHOLE_GRID * RF_BIG_TILE));
  LM_CUDA(cudaMalloc((void**)&ctx->d_rf_plan, sizeof(int32_t) * LM_PLAN_INTS));
  LM_CUDA(cudaMemsetAsync(ctx->d_rf_plan, 0, sizeof(int32_t) * LM_PLAN_INTS, ctx->stream));

  return LMONO_OK;
}

// ------------------------------------------------------------------ export
// scope 0: window cubes in :512-537 order; scope 1: all cubes in logical linear order (:826-830)
__global__ void __launch_bounds__(1024) k_export_offsets(const LmMapState* __restrict__ st, LmMapType M, int scope,
                                                         int32_t* __restrict__ off /*[LM_NSLOT+1]*/, int32_t* __restrict__ order /*[LM_NSLOT]*/) {
  __shared__ int ws[33];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int count = scope == 0 ? st->valid_num : LM_NSLOT;
  for (int base = 0; base < count; base += blockDim.x) {
    const int e = base + threadIdx.x;
    int n = 0, ps = -1;
    if (e < count) {
      if (scope == 0) ps = st->valid_slot[e];
      else {
        int i = e % LM_GW, j = (e / LM_GW) % LM_GH, k = e / (LM_GW * LM_GH);
        ps = d_phys_slot(i - st->cen[0], j - st->cen[1], k - st->cen[2]);
      }
      int sid = M.slot_slab[ps];
      n = sid >= 0 ? M.slab_n[sid] : 0;
      order[e] = ps;
    }
    int total;
    int ex = d_block_exscan(n, ws, &total);
    if (e < count) off[e] = s_carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) s_carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) off[count] = s_carry;
}

__global__ void __launch_bounds__(256) k_export_copy(LmMapType M, const int32_t* __restrict__ off, const int32_t* __restrict__ order,
                                                     int count_slots, float4* __restrict__ out, int cap_out) {
  const int e = blockIdx.x;
  if (e >= count_slots) return;
  const int ps = order[e];
  const int sid = M.slot_slab[ps];
  if (sid < 0) return;
  const int n = M.slab_n[sid];
  const float4* src = M.pts + ((size_t)sid * 2 + M.slab_cur[sid]) * M.cap;
  const int o = off[e];
  for (int i = threadIdx.x; i < n; i += blockDim.x) if (o + i < cap_out) out[o + i] = src[i];
}

// which = 2: corner and surf interleaved cube by cube -- the order of the reference's publishers
// (laserMapping.cpp:808-816: surround += corner[ind]; surround += surf[ind]; :826-830 likewise over all 4851 cubes)
__global__ void __launch_bounds__(256) k_export_copy2(LmMapType M0, LmMapType M1, const int32_t* __restrict__ off0, const int32_t* __restrict__ off1,
                                                      const int32_t* __restrict__ order, int count_slots, float4* __restrict__ out, int cap_out) {
  const int e = blockIdx.x >> 1, ty = blockIdx.x & 1;
  if (e >= count_slots) return;
  const LmMapType& M = ty == 0 ? M0 : M1;
  const int ps = order[e];
  const int sid = M.slot_slab[ps];
  if (sid < 0) return;
  const int n = M.slab_n[sid];
  const float4* src = M.pts + ((size_t)sid * 2 + M.slab_cur[sid]) * M.cap;
  const int o = ty == 0 ? off0[e] + off1[e] : off0[e + 1] + off1[e];      // corner cubes 0..e and surf cubes 0..e-1 precede surf cube e
  for (int i = threadIdx.x; i < n; i += blockDim.x) if (o + i < cap_out) out[o + i] = src[i];
}

static int export_interleaved(lmono_ctx* ctx, int scope, int* n_total) {
  int32_t* off0 = ctx->d_export_off;
  int32_t* order = ctx->d_export_off + LM_NSLOT + 8;
  int32_t* off1 = ctx->d_export_off + 2 * LM_NSLOT + 16;
  int32_t* order1 = ctx->d_export_off + 3 * LM_NSLOT + 24;
  k_export_offsets<<<1, 1024, 0, ctx->stream>>>(ctx->d_state, ctx->map[0], scope, off0, order);
  LM_LAUNCH_CHECK();
  k_export_offsets<<<1, 1024, 0, ctx->stream>>>(ctx->d_state, ctx->map[1], scope, off1, order1);
  LM_LAUNCH_CHECK();
  int valid_num = LM_NSLOT;
  if (scope == 0) {
    LM_CUDA(cudaMemcpyAsync(&ctx->h_state->valid_num, &ctx->d_state->valid_num, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    LM_CUDA(cudaStreamSynchronize(ctx->stream));
    valid_num = ctx->h_state->valid_num;
  }
  int t0 = 0, t1 = 0;
  LM_CUDA(cudaMemcpyAsync(&t0, off0 + valid_num, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  LM_CUDA(cudaMemcpyAsync(&t1, off1 + valid_num, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  const int total = t0 + t1;
  *n_total = total;
  if (total == 0) return LMONO_OK;
  if ((size_t)total > ctx->export_cap) {
    cudaFree(ctx->d_export);
    ctx->export_cap = (size_t)total + (total >> 2) + 1024;
    LM_CUDA(cudaMalloc((void**)&ctx->d_export, ctx->export_cap * sizeof(float4)));
  }
  k_export_copy2<<<2 * valid_num, 256, 0, ctx->stream>>>(ctx->map[0], ctx->map[1], off0, off1, order, valid_num, ctx->d_export, total);
  LM_LAUNCH_CHECK();
  return LMONO_OK;
}

int lm_map_export_device(lmono_ctx* ctx, int which, int scope, int* n_total) {
  LM_NEED_MAP();
  if (which == 2) return export_interleaved(ctx, scope, n_total);
  LmMapType& M = ctx->map[which];
  int32_t* order = ctx->d_export_off + LM_NSLOT + 8;
  k_export_offsets<<<1, 1024, 0, ctx->stream>>>(ctx->d_state, M, scope, ctx->d_export_off, order);
  LM_LAUNCH_CHECK();
  int valid_num = LM_NSLOT;
  if (scope == 0) {
    LM_CUDA(cudaMemcpyAsync(&ctx->h_state->valid_num, &ctx->d_state->valid_num, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    LM_CUDA(cudaStreamSynchronize(ctx->stream));
    valid_num = ctx->h_state->valid_num;
  }
  int total = 0;
  LM_CUDA(cudaMemcpyAsync(&total, ctx->d_export_off + valid_num, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  *n_total = total;
  if (total == 0) return LMONO_OK;
  if ((size_t)total > ctx->export_cap) {
    cudaFree(ctx->d_export);
    ctx->export_cap = (size_t)total + (total >> 2) + 1024;
    LM_CUDA(cudaMalloc((void**)&ctx->d_export, ctx->export_cap * sizeof(float4)));
  }
  k_export_copy<<<valid_num, 256, 0, ctx->stream>>>(M, ctx->d_export_off, order, valid_num, ctx->d_export, total);
  LM_LAUNCH_CHECK();
  return LMONO_OK;
}

// ------------------------------------------------------------------ import into an empty map
// composite = slot(13) | voxel key(30) | index(21)
__global__ void __launch_bounds__(256) k_import_keys(LmMapState* __restrict__ st, LmMapType M, const float4* __restrict__ pts, int n,
                                                     unsigned long long* __restrict__ comp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pts[i];
  int cI = d_cube_coord((double)p.x, st->cen[0]), cJ = d_cube_coord((double)p.y, st->cen[1]), cK = d_cube_coord((double)p.z, st->cen[2]);
  unsigned long long ps = 8191ULL; uint32_t vkey = 0;
  if (cI >= 0 && cI < LM_GW && cJ >= 0 && cJ < LM_GH && cK >= 0 && cK < LM_GD && d_shard_keep(p, M.leaf, M.inv_leaf, st->shard_rank, st->shard_n)) {
    int g3[3] = { cI - st->cen[0], cJ - st->cen[1], cK - st->cen[2] };
    ps = (unsigned long long)d_phys_slot(g3[0], g3[1], g3[2]);
    vkey = d_cube_voxel_key(p, M.inv_leaf, g3);
  }
  comp[i] = (ps << 51) | ((unsigned long long)vkey << 21) | (unsigned long long)i;
}

__global__ void __launch_bounds__(256) k_import_count(const unsigned long long* __restrict__ sorted, int n, int32_t* __restrict__ blockcnt) {
  __shared__ int ws[33];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int head = 0;
  if (i < n && (sorted[i] >> 51) != 8191ULL) head = (i == 0) || ((sorted[i] >> 21) != (sorted[i - 1] >> 21));
  int total;
  d_block_exscan(head, ws, &total);
  if (threadIdx.x == 0) blockcnt[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_blocks(int32_t* __restrict__ blockcnt, int nblocks) {
  __shared__ int ws[33];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += blockDim.x) {
    const int e = base + threadIdx.x;
    int v = e < nblocks ? blockcnt[e] : 0;
    int total;
    int ex = d_block_exscan(v, ws, &total);
    if (e < nblocks) blockcnt[e] = s_carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) s_carry += total;
    __syncthreads();
  }
}

// head ranks + first head rank of each slot
__global__ void __launch_bounds__(256) k_import_ranks(const unsigned long long* __restrict__ sorted, int n, const int32_t* __restrict__ blockoff,
                                                      int32_t* __restrict__ head_rank, int32_t* __restrict__ slot_first_rank) {
  __shared__ int ws[33];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int head = 0; unsigned long long me = 0;
  if (i < n) { me = sorted[i]; if ((me >> 51) != 8191ULL) head = (i == 0) || ((me >> 21) != (sorted[i - 1] >> 21)); }
  int total;
  int ex = d_block_exscan(head, ws, &total);
  if (i < n) {
    head_rank[i] = head ? blockoff[blockIdx.x] + ex : -1;
    if (head && (i == 0 || (sorted[i - 1] >> 51) != (me >> 51))) slot_first_rank[(int)(me >> 51)] = blockoff[blockIdx.x] + ex;
  }
}

// one thread per physical slot: allocate slabs in slot order (deterministic placement)
__global__ void k_import_alloc(LmMapState* __restrict__ st, LmMapType M, const int32_t* __restrict__ slot_first_rank) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  for (int ps = 0; ps < LM_NSLOT; ++ps) {
    if (slot_first_rank[ps] < 0) continue;
    if (M.slot_slab[ps] >= 0) { st->fault |= LM_FAULT_IMPORT_NONEMPTY; continue; }
    int top = *M.free_top - 1;
    if (top < 0) { st->fault |= LM_FAULT_POOL_EXHAUSTED; continue; }
    *M.free_top = top;
    int sid = M.free_stack[top];
    M.slot_slab[ps] = sid;
    M.slab_n[sid] = 0; M.slab_nsorted[sid] = 0; M.slab_cur[sid] = 0; M.slab_dirty[sid] = 1;
    // absolute cube coordinate from the physical slot and the window offsets
    int pi = ps % LM_GW, pj = (ps / LM_GW) % LM_GH, pk = ps / (LM_GW * LM_GH);
    // logical l in [0,dim) with (l - cen) mod dim == p  =>  l = (p + cen) mod dim
    int li = d_pmod(pi + st->cen[0], LM_GW), lj = d_pmod(pj + st->cen[1], LM_GH), lk = d_pmod(pk + st->cen[2], LM_GD);
    M.slab_g[sid * 4 + 0] = li - st->cen[0]; M.slab_g[sid * 4 + 1] = lj - st->cen[1]; M.slab_g[sid * 4 + 2] = lk - st->cen[2];
  }
}

__global__ void __launch_bounds__(256) k_import_write(LmMapState* __restrict__ st, LmMapType M, const float4* __restrict__ pts,
                                                      const unsigned long long* __restrict__ sorted, int n,
                                                      const int32_t* __restrict__ head_rank, const int32_t* __restrict__ slot_first_rank) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int hr = head_rank[i];
  if (hr < 0) return;
  const unsigned long long me = sorted[i];
  const int ps = (int)(me >> 51);
  const int sid = M.slot_slab[ps];
  if (sid < 0) return;
  const unsigned long long vk = me >> 21;
  float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f; int cnt = 0;
  int j = i;
  for (; j < n && (sorted[j] >> 21) == vk; ++j) {
    float4 t = pts[(uint32_t)(sorted[j] & 0x1FFFFFULL)];
    sx = __fadd_rn(sx, t.x); sy = __fadd_rn(sy, t.y); sz = __fadd_rn(sz, t.z); si = __fadd_rn(si, t.w);
    ++cnt;
  }
  const float c = (float)cnt;
  const int pos = hr - slot_first_rank[ps];
  if (pos >= M.cap) { atomicOr(&st->fault, LM_FAULT_CUBE_OVERFLOW); return; }
  float4* dst = M.pts + ((size_t)sid * 2) * M.cap;
  dst[pos] = make_float4(__fdiv_rn(sx, c), __fdiv_rn(sy, c), __fdiv_rn(sz, c), __fdiv_rn(si, c));
  M.pkey[((size_t)sid * 2) * M.cap + pos] = (uint32_t)(vk & 0x3FFFFFFFULL);
  // last voxel of this slot fixes the slab size
  if (j >= n || (sorted[j] >> 51) != (me >> 51)) { M.slab_n[sid] = pos + 1; M.slab_nsorted[sid] = pos + 1; }
}

// after import a centroid may have crossed a voxel border: demote such slabs to "unsorted"
__global__ void __launch_bounds__(256) k_import_verify(LmMapType M) {
  __shared__ int s_flag;
  const int sid = blockIdx.x;
  if (sid >= M.n_slabs) return;
  const int n = M.slab_n[sid];
  if (n == 0 || M.slab_nsorted[sid] != n) return;
  if (threadIdx.x == 0) s_flag = 0;
  __syncthreads();
  const float4* src = M.pts + ((size_t)sid * 2 + M.slab_cur[sid]) * M.cap;
  const int g3[3] = { M.slab_g[sid * 4], M.slab_g[sid * 4 + 1], M.slab_g[sid * 4 + 2] };
  for (int i = threadIdx.x + 1; i < n; i += blockDim.x)
    if (d_cube_voxel_key(src[i], M.inv_leaf, g3) <= d_cube_voxel_key(src[i - 1], M.inv_leaf, g3)) s_flag = 1;
  __syncthreads();
  if (threadIdx.x == 0 && s_flag) { M.slab_nsorted[sid] = 0; M.slab_unsorted[sid] = 1; }
}

int lm_map_import_device(lmono_ctx* ctx, int which, const float4* d_pts, int n,
                         unsigned long long* d_a, unsigned long long* d_b, unsigned long long* d_c,
                         int32_t* d_n, int32_t* d_blockcnt, int32_t* d_head_rank) {
  LM_NEED_MAP();
  LmMapType& M = ctx->map[which];
  const int blocks = lm_div_up(n, 256);
  LM_CUDA(cudaMemcpyAsync(d_n, &n, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  k_import_keys<<<blocks, 256, 0, ctx->stream>>>(ctx->d_state, M, d_pts, n, d_a);
  LM_LAUNCH_CHECK();
  int rc = lm_sort_u64(ctx, d_a, d_b, d_c, d_n, n);
  if (rc) return rc;
  k_import_count<<<blocks, 256, 0, ctx->stream>>>(d_c, n, d_blockcnt);
  LM_LAUNCH_CHECK();
  k_scan_blocks<<<1, 1024, 0, ctx->stream>>>(d_blockcnt, blocks);
  LM_LAUNCH_CHECK();
  LM_CUDA(cudaMemsetAsync(ctx->d_slot_first, 0xFF, sizeof(int32_t) * LM_NSLOT, ctx->stream));
  k_import_ranks<<<blocks, 256, 0, ctx->stream>>>(d_c, n, d_blockcnt, d_head_rank, ctx->d_slot_first);
  LM_LAUNCH_CHECK();
  k_import_alloc<<<1, 32, 0, ctx->stream>>>(ctx->d_state, M, ctx->d_slot_first);
  LM_LAUNCH_CHECK();
  k_import_write<<<blocks, 256, 0, ctx->stream>>>(ctx->d_state, M, d_pts, d_c, n, d_head_rank, ctx->d_slot_first);
  LM_LAUNCH_CHECK();
  k_import_verify<<<M.n_slabs, 256, 0, ctx->stream>>>(M);
  LM_LAUNCH_CHECK();
  return LMONO_OK;
}
