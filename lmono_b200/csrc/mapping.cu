// C-ABI entry points of the laserMapping stage (Aloam/src/laserMapping.cpp:307-801, 838-842)
// and its test hooks.  A step is a fixed sequence of kernel launches on the ctx stream with
// no host synchronisation until the results are collected.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>

int lm_map_clear_device(lmono_ctx* ctx);
int lm_map_evict_device(lmono_ctx* ctx, int keep, int* n_freed);
int lm_map_export_device(lmono_ctx* ctx, int which, int scope, int* n_total);
int lm_map_import_device(lmono_ctx* ctx, int which, const float4* d_pts, int n,
                         unsigned long long* d_a, unsigned long long* d_b, unsigned long long* d_c,
                         int32_t* d_n, int32_t* d_blockcnt, int32_t* d_head_rank);

// transformUpdate (:148-152) + frame counter
__global__ void k_transform_update(LmMapState* st) {
  lm_pdl_enter();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  d_transform_update(st);
}

// :838-842 full-resolution sweep to the world frame
__global__ void __launch_bounds__(256) k_transform_cloud(const LmMapState* __restrict__ st, const float4* __restrict__ in, int n, float4* __restrict__ out) {
  lm_pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = d_associate(st->q_w_curr, st->t_w_curr, in[i]);
}

__global__ void k_set_counts(LmMapState* st, int n0, int n1) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { st->raw_n[0] = n0; st->raw_n[1] = n1; }
}

// every per-step scalar (feature counts, odometry pose, input pointers) enters through this one launch, so the rest
// of the step is a parameter-free kernel sequence that can be replayed as a CUDA graph
struct StepArgs { double q[4]; double t[3]; int n0, n1; int set_wmap; int slot; double wq[4]; double wt[3]; const float4* in[2];
                  const void* src[2]; int stride[2]; int ioff[2];
                  const double* pose_src;       // not NULL: q[4], t[3] of wodom_curr are read from device memory (fused sweep: the odometry stage's result)
                  const int32_t* n_src; int n_cap; };   // not NULL: feature counts read from device memory, n_src[2] corner, n_src[4] surf (scanRegistration's counts); n0 / n1 are bounds
__device__ __forceinline__ void d_apply_step_args(LmMapState* st, const StepArgs& a) {
  st->raw_n[0] = a.n0; st->raw_n[1] = a.n1;
  if (a.n_src) {
    const int c = a.n_src[2], s = a.n_src[4];
    if (c > a.n0 || s > a.n1 || c > a.n_cap || s > a.n_cap) atomicOr(&st->fault, LM_FAULT_FEATURE_OVERFLOW);
    st->raw_n[0] = min(c, min(a.n0, a.n_cap)); st->raw_n[1] = min(s, min(a.n1, a.n_cap));
  }
  st->in_ptr[0] = a.in[0]; st->in_ptr[1] = a.in[1];
  for (int k = 0; k < 2; ++k) { st->in_src[k] = a.src[k]; st->in_stride[k] = a.stride[k]; st->in_ioff[k] = a.ioff[k]; }
  st->result_slot = a.slot;
  if (a.set_wmap) {       // caller-supplied q/t_wmap_wodom (sequence batches; same effect as lmono_map_set_state before the step)
    for (int k = 0; k < 4; ++k) st->q_wmap_wodom[k] = a.wq[k];
    for (int k = 0; k < 3; ++k) st->t_wmap_wodom[k] = a.wt[k];
  }
  if (a.pose_src) {
    for (int k = 0; k < 4; ++k) st->q_wodom_curr[k] = a.pose_src[k];
    for (int k = 0; k < 3; ++k) st->t_wodom_curr[k] = a.pose_src[4 + k];
  } else {
    for (int k = 0; k < 4; ++k) st->q_wodom_curr[k] = a.q[k];
    for (int k = 0; k < 3; ++k) st->t_wodom_curr[k] = a.t[k];
  }
}
__global__ void k_step_args(LmMapState* st, StepArgs a) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  d_apply_step_args(st, a);
}
// the same for up to LM_ARGS_CHUNK sequences of a batch in one launch (one thread per sequence)
constexpr int LM_ARGS_CHUNK = 16;
struct BatchStepArgs { LmMapState* st[LM_ARGS_CHUNK]; StepArgs a[LM_ARGS_CHUNK]; int n; };
static_assert(sizeof(BatchStepArgs) < 4000, "kernel parameter space");
__global__ void k_batch_args(BatchStepArgs b) {
  if (threadIdx.x < b.n) d_apply_step_args(b.st[threadIdx.x], b.a[threadIdx.x]);
}
// last node of a pipelined step: the state (pose, counts, solver summaries, fault bits) goes to the page-locked result
// mirror the host reads after the step's event -- a plain store over PCIe instead of a copy-engine node, so that two
// steps of a ctx can be in flight with one graph (the slot is a per-step argument)
__global__ void __launch_bounds__(128) k_publish_state(const LmMapState* __restrict__ st, LmMapState* h0, LmMapState* h1) {
  lm_pdl_enter();
  const int4* __restrict__ src = reinterpret_cast<const int4*>(st);
  int4* dst = reinterpret_cast<int4*>(st->result_slot ? h1 : h0);
  for (int i = threadIdx.x; i < (int)(sizeof(LmMapState) / 16); i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
}

// :542-550 VoxelGrid of the incoming corner and surf features, both clouds through the same four launches.
// d_corner / d_surf == NULL: read the input pointers from LmMapState::in_ptr (graph replay)
static int voxel_both(lmono_ctx* ctx, const float4* d_corner, int nc, const float4* d_surf, int ns) {
  const float4* in[2] = { d_corner, d_surf };
  const float4* const* ind[2] = { d_corner ? nullptr : &ctx->d_state->in_ptr[0], d_surf ? nullptr : &ctx->d_state->in_ptr[1] };
  const int32_t* n_dev[2] = { &ctx->d_state->raw_n[0], &ctx->d_state->raw_n[1] };
  const int n_max[2] = { nc, ns };
  const float leaf[2] = { ctx->map[0].leaf, ctx->map[1].leaf };
  float4* out[2] = { ctx->d_stack[0], ctx->d_stack[1] };
  int32_t* out_n[2] = { &ctx->d_state->stack_n[0], &ctx->d_state->stack_n[1] };
  return lm_voxel_grid_multi(ctx, 2, in, n_dev, n_max, leaf, out, out_n, ind, /*fetch=*/!d_corner && !d_surf);
}

// the step body: nc / ns only size the launch grids (every kernel reads the real counts and the input pointers
// from the state)
__global__ void k_nop() {}
static int enqueue_body(lmono_ctx* ctx, int nc, int ns) {
  int rc;
  static const int n_dummy = getenv("LMONO_DUMMY_LAUNCHES") ? atoi(getenv("LMONO_DUMMY_LAUNCHES")) : 0;   // experiment: is a batch launch-rate bound?
  for (int i = 0; i < n_dummy; ++i) k_nop<<<1, 32, 0, ctx->stream>>>();
  static const int n_big = getenv("LMONO_DUMMY_BIG") ? atoi(getenv("LMONO_DUMMY_BIG")) : 0;
  for (int i = 0; i < n_big; ++i) k_nop<<<dim3(75, 2), 1024, 0, ctx->stream>>>();
  static const bool tl_env = getenv("LMONO_TIMELINE") && getenv("LMONO_TIMELINE")[0] == '1';
  if (tl_env) {
    ctx->tl_on = true; ctx->tl_n = 0;
    lm_tl(ctx, "begin", 0);
  }
  struct TlOff { lmono_ctx* c; ~TlOff() { c->tl_on = false; } } tl_off{ctx};
  // window upkeep + index build (:309-539, :558-559) and the VoxelGrid of the incoming features (:542-550) do not
  // depend on each other: inside a graph capture they become two parallel branches
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(ctx->stream, &cap);
  const bool fork = cap == cudaStreamCaptureStatusActive && ctx->side_stream != nullptr && !ctx->tl_on;
  cudaStream_t main_stream = ctx->stream;
  if (fork) {
    LM_CUDA(cudaEventRecord(ctx->ev_side0, main_stream));
    LM_CUDA(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_side0, 0));
    ctx->stream = ctx->side_stream;
  }
  lm_prof_begin(ctx, LM_PROF_WINDOW);
  rc = lm_map_begin_step(ctx, nullptr, nullptr);                            // :309-539
  lm_prof_end(ctx);
  lm_prof_begin(ctx, LM_PROF_INDEX);
  if (!rc) rc = lm_map_index_build(ctx);                                    // replaces kdtree setInputCloud :558-559
  lm_prof_end(ctx);
  if (fork) {
    ctx->stream = main_stream;
    if (!rc) LM_CUDA(cudaEventRecord(ctx->ev_side1, ctx->side_stream));
  }
  if (rc) return rc;
  // :542-550 VoxelGrid of the incoming features
  lm_prof_begin(ctx, LM_PROF_VOXEL);
  if ((rc = voxel_both(ctx, nullptr, nc, nullptr, ns))) return rc;
  lm_prof_end(ctx);
  if (fork) LM_CUDA(cudaStreamWaitEvent(main_stream, ctx->ev_side1, 0));
  if (ctx->shard_p2p && (rc = lm_shard_gate_xchg(ctx))) return rc;          // :554 on the global window content (shard.cu)
  for (int iter = 0; iter < 2; ++iter) {                                    // :562
    lm_prof_begin(ctx, LM_PROF_ASSOC);
    if ((rc = lm_map_associate(ctx, nc, ns))) return rc;                    // :577-687
    lm_prof_end(ctx);
    lm_prof_begin(ctx, LM_PROF_SOLVE);
    if ((rc = lm_solve_enqueue(ctx, iter, nc, ns, 4))) return rc;           // :713-720
    lm_prof_end(ctx);
  }
  if (nc + ns > 0) {
    if ((rc = lm_map_insert_and_refilter(ctx, nc, ns, /*transform_update=*/true))) return rc;   // :734, :737-801
  } else {
    LM_LAUNCH_PDL(k_transform_update, 1, 32, 0, ctx->d_state);            // :734
    LM_LAUNCH_CHECK();
    if ((rc = lm_map_insert_and_refilter(ctx, nc, ns, false))) return rc;
  }
  return LMONO_OK;
}

static int bucket_up(int n, int cap) {
  if (n <= 0) return 0;
  long long b = ((long long)n + LM_GRAPH_BUCKET - 1) / LM_GRAPH_BUCKET * LM_GRAPH_BUCKET;
  return (int)(b > cap ? cap : b);
}

// one step's inputs as the kernels see them: d[k] = float4 XYZI array in device memory that the step reads (for a
// fused upload: the ctx's own d_in[k], filled by k_vg_keys from src[k]); src / stride / ioff describe the caller's
// page-locked AoS buffer (stride 0 = d[k] already holds the data)
struct LmStepIn { const float4* d[2]; int n[2]; const void* src[2]; int stride[2]; int ioff[2]; const double* pose_src; const int32_t* n_src; };

static LmStepIn step_in_device(const float4* d_corner, int nc, const float4* d_surf, int ns) {
  LmStepIn in;
  memset(&in, 0, sizeof(in));
  in.d[0] = d_corner; in.d[1] = d_surf; in.n[0] = nc; in.n[1] = ns;
  return in;
}

// host cloud -> step input.  Page-locked memory (cudaHostAlloc / cudaHostRegister: the device can address it) is not
// copied here at all: k_vg_keys fetches it inside the step.  Anything else is staged with cudaMemcpyAsync (+ unpack) on
// the ctx stream.
static int resolve_input(lmono_ctx* ctx, lmono_cloud_view v, int which, LmStepIn* in) {
  if (v.n < 0 || (v.n > 0 && !v.base) || (v.n > 0 && (v.stride_bytes < 12 || (v.stride_bytes & 3)))) return LMONO_E_ARG;
  if (v.n > ctx->max_feat) return LMONO_E_CAPACITY;
  in->d[which] = ctx->d_in[which]; in->n[which] = v.n;
  in->src[which] = nullptr; in->stride[which] = 0; in->ioff[which] = 0;
  if (v.n == 0) return LMONO_OK;
  const bool layout_ok = v.intensity_offset < 0 || ((v.intensity_offset & 3) == 0 && v.intensity_offset + 4 <= v.stride_bytes);
  if (ctx->zero_copy_on && layout_ok && ((uintptr_t)v.base & 3) == 0) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, v.base) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
      in->src[which] = at.devicePointer;
      in->stride[which] = v.stride_bytes;
      in->ioff[which] = v.intensity_offset;
      return LMONO_OK;
    } else cudaGetLastError();
  }
  return lm_upload_cloud(ctx, v, ctx->d_raw[which], ctx->d_in[which], nullptr);
}

static void fill_step_args(StepArgs* a, const LmStepIn& in, const lmono_pose* wodom_curr, const lmono_pose* wmap_in, int slot) {
  memset(a, 0, sizeof(*a));
  a->pose_src = in.pose_src; a->n_src = in.n_src; a->n_cap = 1 << 30;
  if (wodom_curr) { for (int k = 0; k < 4; ++k) a->q[k] = wodom_curr->q[k]; for (int k = 0; k < 3; ++k) a->t[k] = wodom_curr->t[k]; }
  a->n0 = in.n[0]; a->n1 = in.n[1];
  for (int k = 0; k < 2; ++k) { a->in[k] = in.d[k]; a->src[k] = in.src[k]; a->stride[k] = in.stride[k]; a->ioff[k] = in.ioff[k]; }
  a->slot = slot;
  a->set_wmap = wmap_in != nullptr;
  if (wmap_in) { for (int k = 0; k < 4; ++k) a->wq[k] = wmap_in->q[k]; for (int k = 0; k < 3; ++k) a->wt[k] = wmap_in->t[k]; }
}

// enqueue the whole step
static int enqueue_step(lmono_ctx* ctx, const LmStepIn& in, const lmono_pose* wodom_curr, const lmono_pose* wmap_in = nullptr, int slot = 0) {
  LM_NEED_MAP();
  const int nc = in.n[0], ns = in.n[1];
  if (nc < 0 || ns < 0 || nc > ctx->max_feat || ns > ctx->max_feat) return LMONO_E_CAPACITY;
  int rc;
  LM_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  lm_kmark(ctx, "begin", 0);
  StepArgs a;
  fill_step_args(&a, in, wodom_curr, wmap_in, slot);
  k_step_args<<<1, 32, 0, ctx->stream>>>(ctx->d_state, a);
  LM_LAUNCH_CHECK();
  const int nc_cap = bucket_up(nc, ctx->max_feat), ns_cap = bucket_up(ns, ctx->max_feat);
  if (!ctx->graphs_on || ctx->prof_on || ctx->kmark_on) {
    if ((rc = enqueue_body(ctx, nc_cap, ns_cap))) return rc;
  } else {
    LmGraphEntry* g = nullptr;
    for (int i = 0; i < ctx->n_graphs; ++i) {
      LmGraphEntry& e = ctx->graphs[i];
      if (e.nc_cap == nc_cap && e.ns_cap == ns_cap && e.form == lm_graph_form(ctx->batch_n)) { g = &e; break; }
    }
    if (!g) {
      if (ctx->n_graphs == LM_MAX_GRAPHS) {      // cache full: drop everything
        for (int i = 0; i < ctx->n_graphs; ++i) cudaGraphExecDestroy(ctx->graphs[i].exec);
        ctx->n_graphs = 0;
      }
      const int64_t l0 = ctx->launches;
      cudaGraph_t graph = nullptr;
      LM_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
      rc = enqueue_body(ctx, nc_cap, ns_cap);
      cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
      if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
      LM_CUDA(ce);
      g = &ctx->graphs[ctx->n_graphs];
      g->nc_cap = nc_cap; g->ns_cap = ns_cap; g->form = lm_graph_form(ctx->batch_n);
      g->n_launch = (int)(ctx->launches - l0);
      ctx->launches = l0;
      ce = cudaGraphInstantiate(&g->exec, graph, 0);
      cudaGraphDestroy(graph);
      LM_CUDA(ce);
      ctx->n_graphs++;
    }
    LM_CUDA(cudaGraphLaunch(g->exec, ctx->stream));
    ctx->launches += g->n_launch;
  }
  LM_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->step_pending = true; ctx->step_timed = true;
  return LMONO_OK;
}

// ------------------------------------------------------------------ cube-sharded global map
// The map's 50 m cubes are distributed over the ranks by lm_cube_owner (each stored with a voxel-complete
// 1 m halo, d_shard_keep); queries and the pose are replicated; a query is associated on the rank that
// owns the cube it falls in, so the 5-NN is local and exact.  One registration is the same kernel
// sequence as enqueue_step, cut at the points where the ranks must exchange data: the host all-reduces
// the 35-double workspace (caller-owned device memory, e.g. a torch tensor -> torch.distributed / NCCL
// over NVLink) after lmono_shard_begin and after every lmono_shard_lm_eval.
// workspace layout: [0..20] J^T J upper triangle, [21..26] J^T r, [27] cost, [28] corner factors,
// [29] surf factors, [30] owned corner map points in the window, [31] owned surf map points, [32..34] pad.
__global__ void k_shard_config(LmMapState* st, int rank, int n) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { st->shard_rank = rank; st->shard_n = n; }
}
__global__ void k_shard_publish_counts(const LmMapState* __restrict__ st, double* __restrict__ ws) {
  if (threadIdx.x < 35) ws[threadIdx.x] = 0.0;
  __syncwarp();
  if (threadIdx.x == 0) { ws[30] = (double)st->shard_owned_n[0]; ws[31] = (double)st->shard_owned_n[1]; }
}
__global__ void k_shard_gate(LmMapState* __restrict__ st, const double* __restrict__ ws) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  // laserMapping.cpp:554 on the GLOBAL window content (halo copies are not counted twice)
  st->from_map_n[0] = (int)ws[30]; st->from_map_n[1] = (int)ws[31];
  st->optimize = (ws[30] > 10.0 && ws[31] > 50.0) ? 1 : 0;
}

extern "C" int lmono_shard_configure(lmono_ctx* ctx, int32_t rank, int32_t nranks, void* d_workspace) {
  if (!ctx || nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && !d_workspace)) return LMONO_E_ARG;
  LM_NEED_MAP();
  ctx->d_shard_ws = (double*)d_workspace;
  k_shard_config<<<1, 32, 0, ctx->stream>>>(ctx->d_state, rank, nranks);
  LM_LAUNCH_CHECK();
  return LMONO_OK;
}

extern "C" int32_t lmono_shard_owner_of_cube(int32_t gi, int32_t gj, int32_t gk, int32_t nranks) { return lm_cube_owner(gi, gj, gk, nranks); }

extern "C" int lmono_shard_begin(lmono_ctx* ctx, const void* d_corner, int32_t nc, const void* d_surf, int32_t ns,
                                 const lmono_pose* wodom_curr) {
  if (!ctx || !wodom_curr || !ctx->d_shard_ws) return LMONO_E_ARG;
  if (nc < 0 || ns < 0 || nc > ctx->max_feat || ns > ctx->max_feat) return LMONO_E_CAPACITY;
  int rc;
  ctx->shard_nc = nc; ctx->shard_ns = ns;
  LM_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  k_set_counts<<<1, 32, 0, ctx->stream>>>(ctx->d_state, nc, ns);
  LM_LAUNCH_CHECK();
  if ((rc = lm_map_begin_step(ctx, wodom_curr, nullptr))) return rc;
  if ((rc = lm_map_index_build(ctx))) return rc;
  if ((rc = voxel_both(ctx, (const float4*)d_corner, nc, (const float4*)d_surf, ns))) return rc;
  k_shard_publish_counts<<<1, 64, 0, ctx->stream>>>(ctx->d_state, ctx->d_shard_ws);
  LM_LAUNCH_CHECK();
  return LMONO_OK;
}

extern "C" int lmono_shard_gate(lmono_ctx* ctx) {
  if (!ctx || !ctx->d_shard_ws) return LMONO_E_ARG;
  k_shard_gate<<<1, 32, 0, ctx->stream>>>(ctx->d_state, ctx->d_shard_ws);
  LM_LAUNCH_CHECK();
  return LMONO_OK;
}

extern "C" int lmono_shard_associate(lmono_ctx* ctx) {
  if (!ctx || !ctx->d_shard_ws) return LMONO_E_ARG;
  return lm_map_associate(ctx, ctx->shard_nc, ctx->shard_ns);
}
extern "C" int lmono_shard_lm_begin(lmono_ctx* ctx, int32_t solve_index) {
  if (!ctx || !ctx->d_shard_ws || solve_index < 0 || solve_index > 1) return LMONO_E_ARG;
  return lm_shard_lm_begin(ctx, solve_index);
}
extern "C" int lmono_shard_lm_eval(lmono_ctx* ctx, int32_t solve_index) {
  if (!ctx || !ctx->d_shard_ws || solve_index < 0 || solve_index > 1) return LMONO_E_ARG;
  return lm_shard_lm_eval(ctx, solve_index);
}
extern "C" int lmono_shard_lm_control(lmono_ctx* ctx, int32_t solve_index) {
  if (!ctx || !ctx->d_shard_ws || solve_index < 0 || solve_index > 1) return LMONO_E_ARG;
  return lm_shard_lm_control(ctx, solve_index);
}
extern "C" int lmono_shard_end(lmono_ctx* ctx) {
  if (!ctx || !ctx->d_shard_ws) return LMONO_E_ARG;
  LM_LAUNCH_PDL(k_transform_update, 1, 32, 0, ctx->d_state);
  LM_LAUNCH_CHECK();
  int rc = lm_map_insert_and_refilter(ctx, ctx->shard_nc, ctx->shard_ns);
  if (rc) return rc;
  LM_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  ctx->step_pending = true; ctx->step_timed = true;
  return LMONO_OK;
}

static void fill_report(const LmMapState* h, lmono_map_report* r, float ms) {
  memset(r, 0, sizeof(*r));
  r->corner_from_map = h->from_map_n[0]; r->surf_from_map = h->from_map_n[1];
  r->corner_stack = h->stack_n[0]; r->surf_stack = h->stack_n[1];
  for (int k = 0; k < 2; ++k) {
    r->corner_num[k] = h->corner_num[k]; r->surf_num[k] = h->surf_num[k];
    r->solve[k].iterations = h->solve[k].iterations; r->solve[k].num_successful = h->solve[k].num_successful;
    r->solve[k].termination = h->solve[k].termination; r->solve[k].num_factors = h->solve[k].num_factors;
    r->solve[k].initial_cost = h->solve[k].initial_cost; r->solve[k].final_cost = h->solve[k].final_cost;
  }
  r->optimized = h->optimize;
  for (int k = 0; k < 3; ++k) { r->center_cube[k] = h->center[k]; r->cen[k] = h->cen[k]; }
  r->ms_gpu = ms;
}

// hand a completed state mirror to the caller
static int deliver(lmono_ctx* ctx, const LmMapState* h, float ms, lmono_pose* w_curr, lmono_pose* wmap_wodom, lmono_map_report* report) {
  if (w_curr) { memcpy(w_curr->q, h->q_w_curr, sizeof(w_curr->q)); memcpy(w_curr->t, h->t_w_curr, sizeof(w_curr->t)); }
  if (wmap_wodom) { memcpy(wmap_wodom->q, h->q_wmap_wodom, sizeof(wmap_wodom->q)); memcpy(wmap_wodom->t, h->t_wmap_wodom, sizeof(wmap_wodom->t)); }
  if (report) fill_report(h, report, ms);
  if (h->fault) {
    fprintf(stderr, "[lmono_b200] device fault bits 0x%x\n", h->fault);
    cudaMemsetAsync(&ctx->d_state->fault, 0, sizeof(uint32_t), ctx->stream);
    return LMONO_E_DEVICE;
  }
  return LMONO_OK;
}

static int collect(lmono_ctx* ctx, lmono_pose* w_curr, lmono_pose* wmap_wodom, lmono_map_report* report) {
  LM_CUDA(cudaMemcpyAsync(ctx->h_state, ctx->d_state, sizeof(LmMapState), cudaMemcpyDeviceToHost, ctx->stream));
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0.f;
  if (ctx->step_pending) { if (ctx->step_timed) cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->step_pending = false; }
  ctx->n_waited = ctx->n_submitted;       // a full synchronisation retires every pipelined step of this ctx as well
  return deliver(ctx, ctx->h_state, ms, w_curr, wmap_wodom, report);
}

extern "C" int lmono_map_step_device(lmono_ctx* ctx, const void* d_corner, int32_t nc, const void* d_surf, int32_t ns,
                                     const lmono_pose* wodom_curr) {
  if (!ctx || !wodom_curr) return LMONO_E_ARG;
  return enqueue_step(ctx, step_in_device((const float4*)d_corner, nc, (const float4*)d_surf, ns), wodom_curr);
}

// ------------------------------------------------------------------ fused sweep: scanRegistration -> laserOdometry -> laserMapping
// The reference runs the three stages as three ROS nodes that hand clouds to each other through topics
// (scanRegistration.cpp:413-441 -> laserOdometry.cpp:195-213,511-590 -> laserMapping.cpp:204-305).  A host that owns all
// three (one process per sequence, BASELINE config C-4) calls this instead: the raw sweep goes up once, the feature
// clouds and the odometry pose stay in device memory between the stages, one small read-back in the middle (the feature
// counts size the next launches) and one at the end.  Results are bit-identical to the three separate calls.
int lm_scan_upload(lmono_ctx* ctx, lmono_cloud_view raw, const float4** d_in);
int lm_scan_enqueue(lmono_ctx* ctx, const float4* d_in, int n, bool n_on_device = false);
int lm_scan_set_n(lmono_ctx* ctx, int n);
int lm_scan_input_buffer(lmono_ctx* ctx, float4** d_in);
int lm_scan_fetch(lmono_ctx* ctx, int n_in, int32_t counts[5], lmono_scan_report* report);
int lm_scan_outputs(lmono_ctx* ctx, const float4** full, const float4** sharp, const float4** less_sharp, const float4** flat,
                    const float4** less_flat, const int32_t** counts);
int lm_odom_enqueue_auto(lmono_ctx* ctx, const float4* sharp, int n_sharp, const float4* less_sharp, int n_ls,
                         const float4* flat, int n_flat, const float4* less_flat, int n_lf, const double** d_pose7);
int lm_odom_readback(lmono_ctx* ctx);
int lm_odom_prepare(lmono_ctx* ctx);
int lm_odom_deliver(lmono_ctx* ctx, lmono_pose* last_curr, lmono_pose* w_curr, lmono_odom_report* report);

extern "C" int lmono_sweep_step(lmono_ctx* ctx, lmono_cloud_view raw, lmono_pose* odom_last_curr, lmono_pose* odom_w_curr,
                                lmono_pose* map_w_curr, lmono_pose* wmap_wodom,
                                lmono_scan_report* scan_report, lmono_odom_report* odom_report, lmono_map_report* map_report) {
  if (!ctx) return LMONO_E_ARG;
  LM_NEED_MAP();
  if (raw.n > ctx->max_sweep) return LMONO_E_CAPACITY;
  int rc;
  const float4* d_in = nullptr;
  if ((rc = lm_scan_upload(ctx, raw, &d_in))) return rc;
  lm_kmark(ctx, "begin", 0);
  if ((rc = lm_scan_enqueue(ctx, d_in, raw.n))) return rc;
  int32_t counts[5];
  if ((rc = lm_scan_fetch(ctx, raw.n, counts, scan_report))) return rc;            // sync #1: n_kept, sharp, less_sharp, flat, less_flat
  const float4 *sharp, *less_sharp, *flat, *less_flat;
  if ((rc = lm_scan_outputs(ctx, nullptr, &sharp, &less_sharp, &flat, &less_flat, nullptr))) return rc;
  if (counts[2] > ctx->max_feat || counts[4] > ctx->max_feat) return LMONO_E_CAPACITY;
  const double* d_pose7 = nullptr;
  LM_CUDA(cudaEventRecord(ctx->ev_o0, ctx->stream));
  if ((rc = lm_odom_enqueue_auto(ctx, sharp, counts[1], less_sharp, counts[2], flat, counts[3], less_flat, counts[4], &d_pose7))) return rc;
  LM_CUDA(cudaEventRecord(ctx->ev_o1, ctx->stream));
  if ((rc = lm_odom_readback(ctx))) return rc;
  // laserMapping consumes /laser_cloud_corner_last = less-sharp and /laser_cloud_surf_last = less-flat of this sweep
  // (laserOdometry.cpp:554-590) with /laser_odom_to_init as the prior
  LmStepIn in = step_in_device(less_sharp, counts[2], less_flat, counts[4]);
  in.pose_src = d_pose7;
  if ((rc = enqueue_step(ctx, in, nullptr))) return rc;
  rc = collect(ctx, map_w_curr, wmap_wodom, map_report);                           // sync #2
  const int rc2 = lm_odom_deliver(ctx, odom_last_curr, odom_w_curr, odom_report);
  if (odom_report) { float ms = 0.f; if (cudaEventElapsedTime(&ms, ctx->ev_o0, ctx->ev_o1) == cudaSuccess) odom_report->ms_gpu = ms; else cudaGetLastError(); }
  return rc ? rc : rc2;
}

// ---- asynchronous fused sweep: no host round trip inside a sweep --------------------------------------------------
// lmono_sweep_step synchronises in the middle (the feature counts size the later launches).  Here every cloud size stays
// on the device: the odometry and mapping stages read scanRegistration's counts there and their grids are sized from
// bounds the host knows (picks per ring and sector, the raw size of this and the previous sweep), so a sweep is one
// run of enqueues and a host that drives several sequences can have all their sweeps in flight at once (config C-4).
int lm_scan_readback(lmono_ctx* ctx);
int lm_scan_deliver(lmono_ctx* ctx, int n_in, lmono_scan_report* report);
int lm_odom_enqueue_devcounts(lmono_ctx* ctx, const float4* sharp, const float4* less_sharp, const float4* flat, const float4* less_flat,
                              const int32_t* d_counts, int n_raw, int n_raw_prev, const double** d_pose7);

constexpr int LM_SWEEP_BUCKET = 16384;
// stages of one sweep on device-resident input, every size read on the device; n_cap bounds the raw sweep
static int sweep_enqueue_stages(lmono_ctx* ctx, const float4* d_in, int n_cap) {
  int rc;
  if ((rc = lm_scan_enqueue(ctx, d_in, n_cap, /*n_on_device=*/true))) return rc;
  const float4 *sharp, *less_sharp, *flat, *less_flat; const int32_t* d_counts;
  if ((rc = lm_scan_outputs(ctx, nullptr, &sharp, &less_sharp, &flat, &less_flat, &d_counts))) return rc;
  const double* d_pose7 = nullptr;
  if ((rc = lm_odom_enqueue_devcounts(ctx, sharp, less_sharp, flat, less_flat, d_counts, n_cap, n_cap, &d_pose7))) return rc;
  const int b_ls = ctx->prm.scan_line * 6 * 20;
  LmStepIn in = step_in_device(less_sharp, b_ls < ctx->max_feat ? b_ls : ctx->max_feat, less_flat, n_cap < ctx->max_feat ? n_cap : ctx->max_feat);
  in.pose_src = d_pose7; in.n_src = d_counts;
  StepArgs a;
  fill_step_args(&a, in, nullptr, nullptr, 0);
  a.n_cap = ctx->max_feat;
  k_step_args<<<1, 32, 0, ctx->stream>>>(ctx->d_state, a);
  LM_LAUNCH_CHECK();
  return enqueue_body(ctx, bucket_up(in.n[0], ctx->max_feat), bucket_up(in.n[1], ctx->max_feat));
}

static int sweep_graph_launch(lmono_ctx* ctx, const float4* d_in, int n) {
  const int n_cap = lm_div_up(n > 0 ? n : 1, LM_SWEEP_BUCKET) * LM_SWEEP_BUCKET;
  const int form = lm_graph_form(ctx->batch_n);
  int gi = -1;
  for (int i = 0; i < ctx->n_sweep_graphs; ++i) if (ctx->sweep_graphs[i].n_cap == n_cap && ctx->sweep_graphs[i].form == form) gi = i;
  if (gi < 0) {
    if (ctx->n_sweep_graphs == 8) { for (int i = 0; i < 8; ++i) cudaGraphExecDestroy(ctx->sweep_graphs[i].exec); ctx->n_sweep_graphs = 0; }
    const int64_t l0 = ctx->launches;
    cudaGraph_t graph = nullptr;
    LM_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = sweep_enqueue_stages(ctx, d_in, n_cap);
    const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    LM_CUDA(ce);
    gi = ctx->n_sweep_graphs;
    ctx->sweep_graphs[gi].n_cap = n_cap; ctx->sweep_graphs[gi].form = form;
    ctx->sweep_graphs[gi].n_launch = (int)(ctx->launches - l0);
    ctx->launches = l0;
    const cudaError_t ci = cudaGraphInstantiate(&ctx->sweep_graphs[gi].exec, graph, 0);
    cudaGraphDestroy(graph);
    LM_CUDA(ci);
    ctx->n_sweep_graphs++;
  }
  LM_CUDA(cudaGraphLaunch(ctx->sweep_graphs[gi].exec, ctx->stream));
  ctx->launches += ctx->sweep_graphs[gi].n_launch;
  LM_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  LM_CUDA(cudaEventRecord(ctx->ev_o0, ctx->stream)); LM_CUDA(cudaEventRecord(ctx->ev_o1, ctx->stream));      // stages are not timed separately inside the graph
  ctx->step_pending = true; ctx->step_timed = true;
  return LMONO_OK;
}

extern "C" int lmono_sweep_submit(lmono_ctx* ctx, lmono_cloud_view raw, int own_stream) {
  if (!ctx) return LMONO_E_ARG;
  LM_NEED_MAP();
  if (raw.n > ctx->max_sweep) return LMONO_E_CAPACITY;
  if (ctx->sweep_outstanding) return LMONO_E_STATE;
  if (!ctx->ev_sweep) LM_CUDA(cudaEventCreateWithFlags(&ctx->ev_sweep, cudaEventDisableTiming));
  cudaStream_t caller = ctx->stream;
  if (own_stream && !ctx->own_stream) {
    // the ctx shares its stream with other sequences: run the sweep on a private stream, ordered after the caller's
    if (!ctx->sweep_stream) LM_CUDA(cudaStreamCreateWithFlags(&ctx->sweep_stream, cudaStreamNonBlocking));
    LM_CUDA(cudaEventRecord(ctx->ev_sync, caller));
    LM_CUDA(cudaStreamWaitEvent(ctx->sweep_stream, ctx->ev_sync, 0));
    ctx->stream = ctx->sweep_stream;
  }
  struct Restore { lmono_ctx* c; cudaStream_t s; ~Restore() { c->stream = s; } } restore{ctx, caller};
  int rc;
  const float4* d_in = nullptr;
  if ((rc = lm_scan_upload(ctx, raw, &d_in))) return rc;                       // records ev0
  if (ctx->graphs_on && !ctx->prof_on && !ctx->kmark_on && !ctx->shard_p2p) {
    // the three stages as ONE graph per raw-size bucket: nothing in it depends on the host (sweep size, feature counts and
    // the odometry pose are read from device memory), so a sweep costs the host an upload, one argument launch, one
    // graph launch and the read-backs
    if ((rc = lm_odom_prepare(ctx))) return rc;
    if ((rc = lm_scan_set_n(ctx, raw.n))) return rc;
    rc = sweep_graph_launch(ctx, d_in, raw.n);
    if (rc) return rc;
    if ((rc = lm_odom_readback(ctx))) return rc;
    if ((rc = lm_scan_readback(ctx))) return rc;
  } else {
  lm_kmark(ctx, "begin", 0);
  if ((rc = lm_scan_enqueue(ctx, d_in, raw.n))) return rc;
  LM_CUDA(cudaEventRecord(ctx->ev_k0, ctx->stream));
  const float4 *sharp, *less_sharp, *flat, *less_flat; const int32_t* d_counts;
  if ((rc = lm_scan_outputs(ctx, nullptr, &sharp, &less_sharp, &flat, &less_flat, &d_counts))) return rc;
  const double* d_pose7 = nullptr;
  LM_CUDA(cudaEventRecord(ctx->ev_o0, ctx->stream));
  if ((rc = lm_odom_enqueue_devcounts(ctx, sharp, less_sharp, flat, less_flat, d_counts, raw.n, ctx->sweep_prev_n > 0 ? ctx->sweep_prev_n : raw.n, &d_pose7))) return rc;
  LM_CUDA(cudaEventRecord(ctx->ev_o1, ctx->stream));
  if ((rc = lm_odom_readback(ctx))) return rc;
  if ((rc = lm_scan_readback(ctx))) return rc;
  const int b_ls = ctx->prm.scan_line * 6 * 20;
  LmStepIn in = step_in_device(less_sharp, b_ls < ctx->max_feat ? b_ls : ctx->max_feat, less_flat, raw.n < ctx->max_feat ? raw.n : ctx->max_feat);
  in.pose_src = d_pose7; in.n_src = d_counts;
  if ((rc = enqueue_step(ctx, in, nullptr))) return rc;                          // records ev0 (again) .. ev1 around the mapping stage
  }
  LM_CUDA(cudaMemcpyAsync(ctx->h_state, ctx->d_state, sizeof(LmMapState), cudaMemcpyDeviceToHost, ctx->stream));
  LM_CUDA(cudaEventRecord(ctx->ev_sweep, ctx->stream));
  // no join back into the caller's stream here (it would chain the sweeps of ctxs that share that stream one behind the
  // other): lmono_sweep_wait blocks the host until the sweep is done, so whatever is enqueued afterwards sees the new map
  ctx->sweep_n_in = raw.n; ctx->sweep_prev_n = raw.n; ctx->sweep_outstanding = true;
  return LMONO_OK;
}

extern "C" int lmono_sweep_wait(lmono_ctx* ctx, lmono_pose* odom_last_curr, lmono_pose* odom_w_curr, lmono_pose* map_w_curr, lmono_pose* wmap_wodom,
                                lmono_scan_report* scan_report, lmono_odom_report* odom_report, lmono_map_report* map_report) {
  if (!ctx) return LMONO_E_ARG;
  if (!ctx->sweep_outstanding) return LMONO_E_STATE;
  LM_CUDA(cudaEventSynchronize(ctx->ev_sweep));
  ctx->sweep_outstanding = false;
  float ms_map = 0.f, ms_scan = 0.f, ms_odom = 0.f;
  if (cudaEventElapsedTime(&ms_map, ctx->ev0, ctx->ev1) != cudaSuccess) cudaGetLastError();
  if (cudaEventElapsedTime(&ms_odom, ctx->ev_o0, ctx->ev_o1) != cudaSuccess) cudaGetLastError();
  ctx->step_pending = false;
  const int rc_scan = lm_scan_deliver(ctx, ctx->sweep_n_in, scan_report);
  if (scan_report) scan_report->ms_gpu = ms_scan;       // the stage start event is reused by the mapping stage: not timed in this path
  const int rc_odom = lm_odom_deliver(ctx, odom_last_curr, odom_w_curr, odom_report);
  if (odom_report) odom_report->ms_gpu = ms_odom;
  const int rc_map = deliver(ctx, ctx->h_state, ms_map, map_w_curr, wmap_wodom, map_report);
  return rc_scan ? rc_scan : (rc_map ? rc_map : rc_odom);
}

// n independent sequences, one fused sweep each (BASELINE config C-4): all sweeps are enqueued before the first result is
// waited for, every ctx on its own stream, so the stages of different sequences overlap on the device
extern "C" int lmono_sweep_step_batch(lmono_ctx* const* ctxs, int32_t n, const lmono_cloud_view* raws, lmono_pose* odom_w_curr, lmono_pose* map_w_curr,
                                      lmono_scan_report* scan_reports, lmono_odom_report* odom_reports, lmono_map_report* map_reports) {
  if (!ctxs || n < 0 || !raws) return LMONO_E_ARG;
  for (int i = 0; i < n; ++i) if (!ctxs[i]) return LMONO_E_ARG;
  int first = LMONO_OK, n_sub = 0;
  for (int i = 0; i < n; ++i) {
    const int rc = lmono_sweep_submit(ctxs[i], raws[i], 1);
    if (rc) { first = rc; break; }
    ++n_sub;
  }
  for (int i = 0; i < n_sub; ++i) {
    const int rc = lmono_sweep_wait(ctxs[i], nullptr, odom_w_curr ? &odom_w_curr[i] : nullptr, map_w_curr ? &map_w_curr[i] : nullptr, nullptr,
                                    scan_reports ? &scan_reports[i] : nullptr, odom_reports ? &odom_reports[i] : nullptr, map_reports ? &map_reports[i] : nullptr);
    if (rc && !first) first = rc;
  }
  return first;
}

extern "C" int lmono_map_collect(lmono_ctx* ctx, lmono_pose* w_curr, lmono_pose* wmap_wodom, lmono_map_report* report) {
  if (!ctx) return LMONO_E_ARG;
  return collect(ctx, w_curr, wmap_wodom, report);
}

static int step_inputs(lmono_ctx* ctx, lmono_cloud_view corner_last, lmono_cloud_view surf_last, LmStepIn* in) {
  memset(in, 0, sizeof(*in));
  int rc;
  if ((rc = resolve_input(ctx, corner_last, 0, in))) return rc;
  return resolve_input(ctx, surf_last, 1, in);
}

extern "C" int lmono_map_step(lmono_ctx* ctx, lmono_cloud_view corner_last, lmono_cloud_view surf_last,
                              const lmono_pose* wodom_curr, lmono_pose* w_curr, lmono_pose* wmap_wodom,
                              lmono_map_report* report, lmono_cloud_view full_res, lmono_cloud_out* registered) {
  if (!ctx || !wodom_curr) return LMONO_E_ARG;
  if (corner_last.n > ctx->max_feat || surf_last.n > ctx->max_feat || full_res.n > ctx->max_sweep) return LMONO_E_CAPACITY;
  int rc;
  LmStepIn in;
  if ((rc = step_inputs(ctx, corner_last, surf_last, &in))) return rc;
  if ((rc = enqueue_step(ctx, in, wodom_curr))) return rc;
  const bool want_full = registered && full_res.n > 0;
  if (want_full) {
    float4* d_full_in = (float4*)ctx->d_raw[0];   // raw staging is free again once the step is enqueued
    if ((rc = lm_upload_cloud(ctx, full_res, ctx->d_raw[2], d_full_in, nullptr))) return rc;
    LM_LAUNCH_PDL(k_transform_cloud, lm_div_up(full_res.n, 256), 256, 0, ctx->d_state, d_full_in, full_res.n, ctx->d_full);
    LM_LAUNCH_CHECK();
  }
  rc = collect(ctx, w_curr, wmap_wodom, report);
  if (want_full) { int rc2 = lm_download_cloud(ctx, ctx->d_full, full_res.n, registered); if (!rc) rc = rc2; }
  else if (registered) registered->n_out = 0;
  return rc;
}

// ---- sequence batches (config C-4): n independent ctxs driven from one host thread, overlapping on the device.
// The steps of all n sequences are captured as parallel branches of ONE CUDA graph (fork / join inside the graph,
// every branch ends by publishing its state into the ctx's page-locked result mirror), cached in ctxs[0].  A batch step
// then costs the host one k_batch_args launch (poses, counts, input pointers / host source descriptors and the result
// slot of every sequence as kernel parameters) and one cudaGraphLaunch on the origin stream.  Host clouds in
// page-locked memory are not copied by the host at all: the first kernel of a branch fetches them (resolve_input).
extern "C" int lmono_map_step_async(lmono_ctx* ctx, lmono_cloud_view corner_last, lmono_cloud_view surf_last, const lmono_pose* wodom_curr) {
  if (!ctx || !wodom_curr) return LMONO_E_ARG;
  LmStepIn in;
  int rc = step_inputs(ctx, corner_last, surf_last, &in);
  if (rc) return rc;
  return enqueue_step(ctx, in, wodom_curr, nullptr);
}

static int batch_lazy_init(lmono_ctx* ctx /*leader*/) {
  if (ctx->bgraphs) return LMONO_OK;
  ctx->cap_streams = (cudaStream_t*)calloc(LM_BATCH_MAX, sizeof(cudaStream_t));
  ctx->bgraphs = (LmBatchGraph*)calloc(LM_MAX_BGRAPHS, sizeof(LmBatchGraph));
  if (!ctx->cap_streams || !ctx->bgraphs) return LMONO_E_ARG;
  ctx->n_bgraphs = 0;
  return LMONO_OK;
}

void lm_batch_free(lmono_ctx* ctx) {
  if (ctx->bgraphs) { for (int i = 0; i < ctx->n_bgraphs; ++i) cudaGraphExecDestroy(ctx->bgraphs[i].exec); free(ctx->bgraphs); ctx->bgraphs = nullptr; }
  if (ctx->cap_streams) { for (int i = 0; i < LM_BATCH_MAX; ++i) if (ctx->cap_streams[i]) cudaStreamDestroy(ctx->cap_streams[i]); free(ctx->cap_streams); ctx->cap_streams = nullptr; }
}

static int publish_state(lmono_ctx* ctx) {
  LM_LAUNCH_PDL(k_publish_state, 1, 128, 0, ctx->d_state, ctx->h_ring[0], ctx->h_ring[1]);
  LM_LAUNCH_CHECK();
  return LMONO_OK;
}

// capture the n step bodies as parallel branches: origin -> fork -> {body_i ; publish state_i} -> join -> origin
static int batch_capture(lmono_ctx* lead, lmono_ctx* const* ctxs, int n, const int* nc_cap, const int* ns_cap, cudaStream_t origin, LmBatchGraph* g) {
  lmono_ctx* ctx = lead;
  for (int i = 0; i < n; ++i)
    if (!lead->cap_streams[i]) LM_CUDA(cudaStreamCreateWithFlags(&lead->cap_streams[i], cudaStreamNonBlocking));
  cudaGraph_t graph = nullptr;
  LM_CUDA(cudaStreamBeginCapture(origin, cudaStreamCaptureModeThreadLocal));
  int rc = LMONO_OK;
  cudaError_t ce = cudaEventRecord(lead->ev_fork, origin);
  for (int i = 0; i < n && !rc && ce == cudaSuccess; ++i) {
    lmono_ctx* c = ctxs[i];
    cudaStream_t saved = c->stream;
    const int64_t l0 = c->launches;
    c->stream = lead->cap_streams[i];
    c->batch_n = n;
    ce = cudaStreamWaitEvent(c->stream, lead->ev_fork, 0);
    if (ce == cudaSuccess) rc = enqueue_body(c, nc_cap[i], ns_cap[i]);
    if (!rc && ce == cudaSuccess) rc = publish_state(c);
    if (!rc && ce == cudaSuccess) ce = cudaEventRecord(c->ev_join, c->stream);
    if (!rc && ce == cudaSuccess) ce = cudaStreamWaitEvent(origin, c->ev_join, 0);
    g->n_launch[i] = (int)(c->launches - l0);
    c->launches = l0;
    c->stream = saved;
    c->batch_n = c->batch_hint;
  }
  cudaError_t ce2 = cudaStreamEndCapture(origin, &graph);
  if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
  LM_CUDA(ce); LM_CUDA(ce2);
  ce = cudaGraphInstantiate(&g->exec, graph, 0);
  cudaGraphDestroy(graph);
  LM_CUDA(ce);
  g->n = n;
  for (int i = 0; i < n; ++i) { g->ctxs[i] = ctxs[i]; g->uids[i] = ctxs[i]->uid; g->nc_cap[i] = nc_cap[i]; g->ns_cap[i] = ns_cap[i]; }
  return LMONO_OK;
}

// in[i]: the inputs of sequence i.  origin: the stream the batch is ordered on.  Every ctx's step is a pipelined
// submission: result slot = n_submitted & 1, completion event ev_res[slot] (at most two may be outstanding per ctx).
static int batch_enqueue(lmono_ctx* const* ctxs, int n, const LmStepIn* in, const lmono_pose* wodom_curr, const lmono_pose* wmap_in,
                         cudaStream_t origin) {
  lmono_ctx* lead = ctxs[0];
  lmono_ctx* ctx = lead;
  int rc;
  if ((rc = batch_lazy_init(lead))) return rc;
  // launch grids of every branch are sized for the largest sequence of the batch (kernels read the real counts from
  // the state): one graph per (ctx set, bucket of the maximum), not one per combination of per-sequence buckets
  int mc = 0, ms = 0;
  for (int i = 0; i < n; ++i) {
    lmono_ctx* c = ctxs[i];
    if (in[i].n[0] < 0 || in[i].n[1] < 0 || in[i].n[0] > c->max_feat || in[i].n[1] > c->max_feat) return LMONO_E_CAPACITY;
    if (c->n_submitted - c->n_waited >= 2) return LMONO_E_ARG;      // both result mirrors of this ctx are still unread
    mc = in[i].n[0] > mc ? in[i].n[0] : mc; ms = in[i].n[1] > ms ? in[i].n[1] : ms;
  }
  int nc_cap[LM_BATCH_MAX], ns_cap[LM_BATCH_MAX];
  for (int i = 0; i < n; ++i) {
    lmono_ctx* c = ctxs[i];
    nc_cap[i] = bucket_up(mc, c->max_feat); ns_cap[i] = bucket_up(ms, c->max_feat);
    // order the batch after whatever the ctx has in flight on its own stream (staged uploads, earlier single steps)
    if (c->stream != origin) { LM_CUDA(cudaEventRecord(c->ev_sync, c->stream)); LM_CUDA(cudaStreamWaitEvent(origin, c->ev_sync, 0)); }
  }
  LM_CUDA(cudaEventRecord(lead->ev0, origin));
  for (int base = 0; base < n; base += LM_ARGS_CHUNK) {
    BatchStepArgs b;
    b.n = n - base < LM_ARGS_CHUNK ? n - base : LM_ARGS_CHUNK;
    for (int k = 0; k < LM_ARGS_CHUNK; ++k) {
      const int i = base + (k < b.n ? k : 0);
      b.st[k] = ctxs[i]->d_state;
      fill_step_args(&b.a[k], in[i], &wodom_curr[i], wmap_in ? &wmap_in[i] : nullptr, (int)(ctxs[i]->n_submitted & 1u));
    }
    k_batch_args<<<1, 32, 0, origin>>>(b);
    LM_LAUNCH_CHECK();
  }
  LmBatchGraph* g = nullptr;
  for (int e = 0; e < lead->n_bgraphs && !g; ++e) {
    LmBatchGraph& q = lead->bgraphs[e];
    if (q.n != n) continue;
    bool same = true;
    for (int i = 0; i < n && same; ++i) same = q.ctxs[i] == ctxs[i] && q.uids[i] == ctxs[i]->uid && q.nc_cap[i] == nc_cap[i] && q.ns_cap[i] == ns_cap[i];
    if (same) g = &q;
  }
  if (!g) {
    if (lead->n_bgraphs == LM_MAX_BGRAPHS) {
      for (int e = 0; e < lead->n_bgraphs; ++e) cudaGraphExecDestroy(lead->bgraphs[e].exec);
      lead->n_bgraphs = 0;
    }
    g = &lead->bgraphs[lead->n_bgraphs];
    if ((rc = batch_capture(lead, ctxs, n, nc_cap, ns_cap, origin, g))) return rc;
    lead->n_bgraphs++;
  }
  LM_CUDA(cudaGraphLaunch(g->exec, origin));
  LM_CUDA(cudaEventRecord(lead->ev1, origin));
  bool any_other = false;
  for (int i = 0; i < n; ++i) {
    lmono_ctx* c = ctxs[i];
    c->launches += g->n_launch[i];
    c->step_pending = true; c->step_timed = (c == lead);
    LM_CUDA(cudaEventRecord(c->ev_res[c->n_submitted & 1u], origin));
    c->n_submitted++;
    any_other |= c->stream != origin;
  }
  if (any_other) {      // later work on a ctx's own stream is ordered after the batch
    LM_CUDA(cudaEventRecord(lead->ev_done, origin));
    for (int i = 0; i < n; ++i) if (ctxs[i]->stream != origin) LM_CUDA(cudaStreamWaitEvent(ctxs[i]->stream, lead->ev_done, 0));
  }
  return LMONO_OK;
}

static bool batch_graph_ok(lmono_ctx* const* ctxs, int n) {
  if (n > LM_BATCH_MAX) return false;
  for (int i = 0; i < n; ++i) {
    if (!ctxs[i]->graphs_on || ctxs[i]->prof_on || ctxs[i]->kmark_on || ctxs[i]->device != ctxs[0]->device) return false;
    for (int j = 0; j < i; ++j) if (ctxs[j] == ctxs[i]) return false;
  }
  return true;
}

// plain launches (LMONO_NO_GRAPH, profiler or kernel marks on, > LM_BATCH_MAX sequences): every ctx on its own stream,
// fork / join with events; the same pipelined-submission bookkeeping as batch_enqueue
static int batch_enqueue_plain(lmono_ctx* const* ctxs, int n, const LmStepIn* in, const lmono_pose* wodom_curr, const lmono_pose* wmap_in,
                               cudaStream_t js) {
  for (int i = 0; i < n; ++i) if (ctxs[i]->n_submitted - ctxs[i]->n_waited >= 2) return LMONO_E_ARG;
  if (js) {
    lmono_ctx* ctx = ctxs[0];
    LM_CUDA(cudaEventRecord(ctx->ev_fork, js));
  }
  for (int i = 0; i < n; ++i) {
    lmono_ctx* ctx = ctxs[i];
    if (js && ctx->stream != js) LM_CUDA(cudaStreamWaitEvent(ctx->stream, ctxs[0]->ev_fork, 0));
    int rc;
    const int slot = (int)(ctx->n_submitted & 1u);
    ctx->batch_n = n;
    rc = enqueue_step(ctx, in[i], &wodom_curr[i], wmap_in ? &wmap_in[i] : nullptr, slot);
    ctx->batch_n = ctx->batch_hint;
    if (rc) return rc;
    if ((rc = publish_state(ctx))) return rc;
    LM_CUDA(cudaEventRecord(ctx->ev_res[slot], ctx->stream));
    ctx->n_submitted++;
    if (js && ctx->stream != js) { LM_CUDA(cudaEventRecord(ctx->ev_join, ctx->stream)); LM_CUDA(cudaStreamWaitEvent(js, ctx->ev_join, 0)); }
  }
  return LMONO_OK;
}

extern "C" int lmono_map_step_device_batch(lmono_ctx* const* ctxs, int32_t n, const void* const* d_corner, const int32_t* n_corner,
                                           const void* const* d_surf, const int32_t* n_surf, const lmono_pose* wodom_curr,
                                           const lmono_pose* wmap_wodom_in, void* join_stream) {
  if (!ctxs || n < 0 || !d_corner || !d_surf || !n_corner || !n_surf || !wodom_curr) return LMONO_E_ARG;
  if (n == 0) return LMONO_OK;
  if (n > LM_BATCH_MAX) return LMONO_E_CAPACITY;
  for (int i = 0; i < n; ++i) if (!ctxs[i]) return LMONO_E_ARG;
  cudaStream_t js = (cudaStream_t)join_stream;
  LmStepIn in[LM_BATCH_MAX];
  for (int i = 0; i < n; ++i) {
    in[i] = step_in_device((const float4*)d_corner[i], n_corner[i], (const float4*)d_surf[i], n_surf[i]);
    // device-resident steps are retired by lmono_map_collect / lmono_sync, not by lmono_map_wait_batch: keep the
    // pipelined-submission window open
    ctxs[i]->n_waited = ctxs[i]->n_submitted;
  }
  if (batch_graph_ok(ctxs, n)) return batch_enqueue(ctxs, n, in, wodom_curr, wmap_wodom_in, js ? js : ctxs[0]->stream);
  return batch_enqueue_plain(ctxs, n, in, wodom_curr, wmap_wodom_in, js);
}

extern "C" int lmono_map_submit_batch(lmono_ctx* const* ctxs, int32_t n, const lmono_cloud_view* corner_last, const lmono_cloud_view* surf_last,
                                      const lmono_pose* wodom_curr, const lmono_pose* wmap_wodom_in) {
  if (!ctxs || n < 0 || !corner_last || !surf_last || !wodom_curr) return LMONO_E_ARG;
  if (n == 0) return LMONO_OK;
  if (n > LM_BATCH_MAX) return LMONO_E_CAPACITY;
  for (int i = 0; i < n; ++i) if (!ctxs[i]) return LMONO_E_ARG;
  LmStepIn in[LM_BATCH_MAX];
  for (int i = 0; i < n; ++i) {     // page-locked clouds: nothing to do here; others: staged on the ctx's own stream
    int rc = step_inputs(ctxs[i], corner_last[i], surf_last[i], &in[i]);
    if (rc) return rc;
  }
  if (batch_graph_ok(ctxs, n)) return batch_enqueue(ctxs, n, in, wodom_curr, wmap_wodom_in, ctxs[0]->stream);
  return batch_enqueue_plain(ctxs, n, in, wodom_curr, wmap_wodom_in, nullptr);
}

extern "C" int lmono_map_wait_batch(lmono_ctx* const* ctxs, int32_t n, lmono_pose* w_curr, lmono_pose* wmap_wodom, lmono_map_report* reports) {
  if (!ctxs || n < 0) return LMONO_E_ARG;
  for (int i = 0; i < n; ++i) if (!ctxs[i] || ctxs[i]->n_waited == ctxs[i]->n_submitted) return LMONO_E_ARG;
  int first = LMONO_OK;
  for (int i = 0; i < n; ++i) {
    lmono_ctx* ctx = ctxs[i];
    const int slot = (int)(ctx->n_waited & 1u);
    cudaError_t e = cudaEventSynchronize(ctx->ev_res[slot]);
    if (e != cudaSuccess) { ctx->last_cuda_error = (int)e; if (!first) first = LMONO_E_CUDA; continue; }
    const bool last = ctx->n_submitted - ctx->n_waited == 1;
    float ms = 0.f;
    if (last && ctx->step_pending) { if (ctx->step_timed) cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->step_pending = false; }
    ctx->n_waited++;
    int rc = deliver(ctx, ctx->h_ring[slot], ms, w_curr ? &w_curr[i] : nullptr, wmap_wodom ? &wmap_wodom[i] : nullptr, reports ? &reports[i] : nullptr);
    if (rc && !first) first = rc;
  }
  return first;
}

extern "C" int lmono_map_step_batch(lmono_ctx* const* ctxs, int32_t n, const lmono_cloud_view* corner_last, const lmono_cloud_view* surf_last,
                                    const lmono_pose* wodom_curr, const lmono_pose* wmap_wodom_in,
                                    lmono_pose* w_curr, lmono_pose* wmap_wodom, lmono_map_report* reports) {
  if (n == 0) return LMONO_OK;
  // a synchronous step retires anything still in flight first, so that the result returned is this step's
  if (ctxs && n > 0)
    for (int i = 0; i < n; ++i) {
      lmono_ctx* ctx = ctxs[i];
      if (!ctx) continue;
      for (; ctx->n_waited != ctx->n_submitted; ctx->n_waited++) LM_CUDA(cudaEventSynchronize(ctx->ev_res[ctx->n_waited & 1u]));
    }
  int rc = lmono_map_submit_batch(ctxs, n, corner_last, surf_last, wodom_curr, wmap_wodom_in);
  if (rc) return rc;
  return lmono_map_wait_batch(ctxs, n, w_curr, wmap_wodom, reports);
}

extern "C" int lmono_map_get_state(lmono_ctx* ctx, lmono_pose* wmap_wodom, int32_t cen[3]) {
  if (!ctx) return LMONO_E_ARG;
  LM_CUDA(cudaMemcpyAsync(ctx->h_state, ctx->d_state, sizeof(LmMapState), cudaMemcpyDeviceToHost, ctx->stream));
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  if (wmap_wodom) { memcpy(wmap_wodom->q, ctx->h_state->q_wmap_wodom, 32); memcpy(wmap_wodom->t, ctx->h_state->t_wmap_wodom, 24); }
  if (cen) for (int k = 0; k < 3; ++k) cen[k] = ctx->h_state->cen[k];
  return LMONO_OK;
}

__global__ void k_set_wmap(LmMapState* st, double qx, double qy, double qz, double qw, double tx, double ty, double tz) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  st->q_wmap_wodom[0] = qx; st->q_wmap_wodom[1] = qy; st->q_wmap_wodom[2] = qz; st->q_wmap_wodom[3] = qw;
  st->t_wmap_wodom[0] = tx; st->t_wmap_wodom[1] = ty; st->t_wmap_wodom[2] = tz;
}

// enqueue-only (ordered on the ctx stream, no host synchronisation)
extern "C" int lmono_map_set_state(lmono_ctx* ctx, const lmono_pose* p) {
  if (!ctx || !p) return LMONO_E_ARG;
  k_set_wmap<<<1, 32, 0, ctx->stream>>>(ctx->d_state, p->q[0], p->q[1], p->q[2], p->q[3], p->t[0], p->t[1], p->t[2]);
  LM_LAUNCH_CHECK();
  return LMONO_OK;
}

extern "C" int32_t lmono_map_result_bytes(void) { return (int32_t)sizeof(LmMapState); }

extern "C" int lmono_map_evict(lmono_ctx* ctx, int32_t keep_cubes, int32_t* n_freed) {
  if (!ctx) return LMONO_E_ARG;
  if (ctx->step_pending) return LMONO_E_STATE;
  int n = 0;
  const int rc = lm_map_evict_device(ctx, keep_cubes, &n);
  if (n_freed) *n_freed = n;
  return rc;
}

extern "C" int lmono_map_clear(lmono_ctx* ctx) {
  if (!ctx) return LMONO_E_ARG;
  return lm_map_clear_device(ctx);
}

extern "C" int lmono_map_export(lmono_ctx* ctx, int which, int scope, lmono_cloud_out* out) {
  if (!ctx || !out || which < 0 || which > 2 || scope < 0 || scope > 1) return LMONO_E_ARG;
  int total = 0;
  int rc = lm_map_export_device(ctx, which, scope, &total);
  if (rc) return rc;
  return lm_download_cloud(ctx, ctx->d_export, total, out);
}

extern "C" int lmono_map_import(lmono_ctx* ctx, int which, lmono_cloud_view pts) {
  if (!ctx || which < 0 || which > 1) return LMONO_E_ARG;
  if (pts.n <= 0) return LMONO_OK;
  if (pts.n >= (1 << 21)) return LMONO_E_CAPACITY;
  const size_t n = (size_t)pts.n;
  float4* d_pts = nullptr; uint8_t* d_rawtmp = nullptr;
  unsigned long long *a = nullptr, *b = nullptr, *c = nullptr; int32_t* ints = nullptr;
  LM_CUDA(cudaMalloc((void**)&d_pts, n * sizeof(float4)));
  LM_CUDA(cudaMalloc((void**)&d_rawtmp, n * (size_t)pts.stride_bytes));
  LM_CUDA(cudaMalloc((void**)&a, n * 8)); LM_CUDA(cudaMalloc((void**)&b, n * 8)); LM_CUDA(cudaMalloc((void**)&c, n * 8));
  LM_CUDA(cudaMalloc((void**)&ints, sizeof(int32_t) * (n + n / 256 + 64)));
  size_t saved = ctx->raw_bytes; ctx->raw_bytes = n * (size_t)pts.stride_bytes;
  int rc = lm_upload_cloud(ctx, pts, d_rawtmp, d_pts, nullptr);
  ctx->raw_bytes = saved;
  if (!rc) rc = lm_map_import_device(ctx, which, d_pts, pts.n, a, b, c, ints, ints + 16, ints + 16 + (n / 256 + 32));
  cudaStreamSynchronize(ctx->stream);
  cudaFree(d_pts); cudaFree(d_rawtmp); cudaFree(a); cudaFree(b); cudaFree(c); cudaFree(ints);
  if (rc) return rc;
  uint32_t bits = 0;
  lmono_last_fault(ctx, &bits);
  if (bits) { fprintf(stderr, "[lmono_b200] import fault bits 0x%x\n", bits); return LMONO_E_DEVICE; }
  return LMONO_OK;
}

extern "C" int lmono_map_prepare_window(lmono_ctx* ctx, const double t_w_curr[3]) {
  if (!ctx || !t_w_curr) return LMONO_E_ARG;
  int rc = lm_map_begin_step(ctx, nullptr, t_w_curr);
  if (rc) return rc;
  if ((rc = lm_map_index_build(ctx))) return rc;
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  return LMONO_OK;
}

extern "C" int lmono_knn5_device(lmono_ctx* ctx, int which, const void* d_q, int32_t n, void* d_idx, void* d_d2) {
  if (!ctx || which < 0 || which > 1) return LMONO_E_ARG;
  return lm_knn5_device(ctx, which, (const float4*)d_q, n, (int32_t*)d_idx, (float*)d_d2);
}

extern "C" int lmono_knn5(lmono_ctx* ctx, int which, lmono_cloud_view q, int32_t* idx, float* d2) {
  if (!ctx || which < 0 || which > 1 || !idx || !d2) return LMONO_E_ARG;
  if (q.n <= 0) return LMONO_OK;
  if (q.n > ctx->max_feat) return LMONO_E_CAPACITY;
  int rc = lm_upload_cloud(ctx, q, ctx->d_raw[0], ctx->d_in[0], nullptr);
  if (rc) return rc;
  int32_t* d_idx = (int32_t*)ctx->d_sort_a; float* d_d2 = (float*)ctx->d_sort_b;
  if ((rc = lm_knn5_device(ctx, which, ctx->d_in[0], q.n, d_idx, d_d2))) return rc;
  LM_CUDA(cudaMemcpyAsync(idx, d_idx, (size_t)q.n * 5 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  LM_CUDA(cudaMemcpyAsync(d2, d_d2, (size_t)q.n * 5 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  return LMONO_OK;
}

__global__ void k_set_pose_and_counts(LmMapState* st, double qx, double qy, double qz, double qw, double tx, double ty, double tz, int n0, int n1) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  st->q_w_curr[0] = qx; st->q_w_curr[1] = qy; st->q_w_curr[2] = qz; st->q_w_curr[3] = qw;
  st->t_w_curr[0] = tx; st->t_w_curr[1] = ty; st->t_w_curr[2] = tz;
  st->stack_n[0] = n0; st->stack_n[1] = n1;
  st->optimize = 1;
}

extern "C" int lmono_map_normal_eq(lmono_ctx* ctx, lmono_cloud_view corner_stack, lmono_cloud_view surf_stack,
                                   const lmono_pose* w, double H[36], double g[6], double* cost,
                                   int32_t* n_corner, int32_t* n_surf) {
  if (!ctx || !w || !H || !g || !cost) return LMONO_E_ARG;
  if (corner_stack.n > ctx->max_feat || surf_stack.n > ctx->max_feat) return LMONO_E_CAPACITY;
  int rc;
  if ((rc = lm_map_begin_step(ctx, nullptr, w->t))) return rc;
  if ((rc = lm_map_index_build(ctx))) return rc;
  if ((rc = lm_upload_cloud(ctx, corner_stack, ctx->d_raw[0], ctx->d_stack[0], nullptr))) return rc;
  if ((rc = lm_upload_cloud(ctx, surf_stack, ctx->d_raw[1], ctx->d_stack[1], nullptr))) return rc;
  k_set_pose_and_counts<<<1, 32, 0, ctx->stream>>>(ctx->d_state, w->q[0], w->q[1], w->q[2], w->q[3], w->t[0], w->t[1], w->t[2],
                                                   corner_stack.n, surf_stack.n);
  LM_LAUNCH_CHECK();
  if ((rc = lm_map_associate(ctx, corner_stack.n, surf_stack.n))) return rc;
  if ((rc = lm_normal_eq_enqueue(ctx, corner_stack.n, surf_stack.n))) return rc;
  double out[44];
  LM_CUDA(cudaMemcpyAsync(out, ctx->d_partials + 32 * 1024, sizeof(out), cudaMemcpyDeviceToHost, ctx->stream));
  LM_CUDA(cudaMemcpyAsync(ctx->h_state, ctx->d_state, sizeof(LmMapState), cudaMemcpyDeviceToHost, ctx->stream));
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  memcpy(H, out, 36 * sizeof(double)); memcpy(g, out + 36, 6 * sizeof(double)); *cost = out[42];
  if (n_corner) *n_corner = ctx->h_state->corner_num[0];
  if (n_surf) *n_surf = ctx->h_state->surf_num[0];
  return LMONO_OK;
}

extern "C" int lmono_voxel_grid(lmono_ctx* ctx, lmono_cloud_view in, float leaf, lmono_cloud_out* out) {
  if (!ctx || !out || !(leaf > 0)) return LMONO_E_ARG;
  if (in.n > ctx->max_feat) return LMONO_E_CAPACITY;
  int rc = lm_upload_cloud(ctx, in, ctx->d_raw[0], ctx->d_in[0], &ctx->d_state->raw_n[0]);
  if (rc) return rc;
  if ((rc = lm_voxel_grid_device(ctx, ctx->d_in[0], &ctx->d_state->raw_n[0], in.n, leaf, ctx->d_stack[0], &ctx->d_state->stack_n[0]))) return rc;
  int n_out = 0;
  LM_CUDA(cudaMemcpyAsync(&n_out, &ctx->d_state->stack_n[0], sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  return lm_download_cloud(ctx, ctx->d_stack[0], n_out, out);
}

// ---- per-phase CUDA-event profiler (bench.py) --------------------------------------------
extern "C" int lmono_profile_enable(lmono_ctx* ctx, int on) {
  if (!ctx) return LMONO_E_ARG;
  if (on && !ctx->prof_ev[0][0]) {
    for (int i = 0; i < LM_PROF_MAX_EVENTS; ++i) { LM_CUDA(cudaEventCreate(&ctx->prof_ev[i][0])); LM_CUDA(cudaEventCreate(&ctx->prof_ev[i][1])); }
  }
  ctx->prof_on = on != 0; ctx->prof_n = 0;
  return LMONO_OK;
}

extern "C" int lmono_profile_read(lmono_ctx* ctx, float* ms /*[LM_PROF_NTAGS]*/, int32_t* counts /*[LM_PROF_NTAGS]*/) {
  if (!ctx || !ms || !counts) return LMONO_E_ARG;
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int t = 0; t < LM_PROF_NTAGS; ++t) { ms[t] = 0.f; counts[t] = 0; }
  for (int i = 0; i < ctx->prof_n; ++i) {
    float e = 0.f;
    if (cudaEventElapsedTime(&e, ctx->prof_ev[i][0], ctx->prof_ev[i][1]) == cudaSuccess) { ms[ctx->prof_tag[i]] += e; counts[ctx->prof_tag[i]]++; }
  }
  ctx->prof_n = 0;
  return LMONO_OK;
}

// ---- per-launch marks: a CUDA event after every kernel launch of the (non-graph) step, keyed by launch site ----
// LMONO_TIMELINE=1: "<file>:<line> <globaltimer ns>" per stamp of the last step of this ctx (lines in launch order; the
// stamp after a launch runs when that kernel has finished)
// How many sequences the caller runs side by side on this GPU OUTSIDE lmono_map_*_batch calls (e.g. one host thread and
// stream per sequence): n >= 4 selects the throughput forms of the kernels for this ctx's single steps too.
extern "C" int lmono_set_concurrency_hint(lmono_ctx* ctx, int32_t n) {
  if (!ctx || n < 1) return LMONO_E_ARG;
  ctx->batch_hint = n; ctx->batch_n = n;
  return LMONO_OK;
}

extern "C" int lmono_timeline_dump(lmono_ctx* ctx, char* buf, int32_t cap) {
  if (!ctx || !buf || cap <= 0) return LMONO_E_ARG;
  buf[0] = 0;
  if (!ctx->d_tl || ctx->tl_n <= 0) return LMONO_OK;
  unsigned long long h[LM_TL_MAX];
  LM_CUDA(cudaMemcpy(h, ctx->d_tl, sizeof(unsigned long long) * ctx->tl_n, cudaMemcpyDeviceToHost));
  int off = 0;
  for (int i = 0; i < ctx->tl_n; ++i) {
    const char* f = ctx->tl_file[i]; const char* sl = strrchr(f, '/');
    off += snprintf(buf + off, off < cap ? cap - off : 0, "%s:%d %llu\n", sl ? sl + 1 : f, ctx->tl_line[i], h[i]);
    if (off >= cap) break;
  }
  return LMONO_OK;
}

extern "C" int lmono_kmarks_enable(lmono_ctx* ctx, int on) {
  if (!ctx) return LMONO_E_ARG;
  if (on && !ctx->kmark_ev) {
    ctx->kmark_ev = (cudaEvent_t*)calloc(LM_KMARK_MAX, sizeof(cudaEvent_t));
    ctx->kmark_file = (const char**)calloc(LM_KMARK_MAX, sizeof(char*));
    ctx->kmark_line = (int*)calloc(LM_KMARK_MAX, sizeof(int));
    for (int i = 0; i < LM_KMARK_MAX; ++i) LM_CUDA(cudaEventCreate(&ctx->kmark_ev[i]));
  }
  ctx->kmark_on = on != 0; ctx->kmark_n = 0;
  return LMONO_OK;
}

// text dump "file:line count total_ms" per launch site (time from the previous mark to this one)
extern "C" int lmono_kmarks_dump(lmono_ctx* ctx, char* buf, int32_t cap) {
  if (!ctx || !buf || cap < 64) return LMONO_E_ARG;
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  struct Site { const char* f; int l; int n; double ms; };
  static Site sites[512];
  int ns = 0;
  for (int i = 1; i < ctx->kmark_n; ++i) {
    if (ctx->kmark_line[i] == 0) continue;      // "begin" marks only delimit steps
    float e = 0.f;
    if (cudaEventElapsedTime(&e, ctx->kmark_ev[i - 1], ctx->kmark_ev[i]) != cudaSuccess) continue;
    int k = 0;
    for (; k < ns; ++k) if (sites[k].f == ctx->kmark_file[i] && sites[k].l == ctx->kmark_line[i]) break;
    if (k == ns) { if (ns == 512) continue; sites[ns].f = ctx->kmark_file[i]; sites[ns].l = ctx->kmark_line[i]; sites[ns].n = 0; sites[ns].ms = 0; ns++; }
    sites[k].n++; sites[k].ms += e;
  }
  int off = 0;
  for (int k = 0; k < ns && off < cap - 160; ++k) {
    const char* f = strrchr(sites[k].f, '/'); f = f ? f + 1 : sites[k].f;
    off += snprintf(buf + off, (size_t)(cap - off), "%s:%d %d %.6f\n", f, sites[k].l, sites[k].n, sites[k].ms);
  }
  buf[off] = 0;
  ctx->kmark_n = 0;
  return LMONO_OK;
}
