// Context lifecycle, parameter defaults, host<->device cloud marshalling.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>

int lm_map_configure_kernels(lmono_ctx* ctx);
void lm_scan_free(lmono_ctx* ctx);
void lm_odom_free(lmono_ctx* ctx);
void lm_color_free(lmono_ctx* ctx);
void lm_batch_free(lmono_ctx* ctx);

extern "C" void lmono_default_params(lmono_params* p) {
  memset(p, 0, sizeof(*p));
  // Aloam/launch/aloam_velodyne_HDL_64.launch:3-13
  p->scan_line = 64; p->minimum_range = 5.0f;
  p->mapping_line_resolution = 0.4f; p->mapping_plane_resolution = 0.8f;
  p->mapping_skip_frame = 1;
  p->max_sweep_points = 262144; p->max_feature_points = 131072;
  p->cube_capacity_corner = 32768; p->cube_capacity_surf = 49152;
  p->max_cubes_corner = 768; p->max_cubes_surf = 768;
  p->image_width = 1241; p->image_height = 376;
}

extern "C" const char* lmono_strerror(int code) {
  switch (code) {
    case LMONO_OK: return "ok";
    case LMONO_E_ARG: return "bad argument";
    case LMONO_E_CAPACITY: return "capacity too small";
    case LMONO_E_CUDA: return "CUDA error";
    case LMONO_E_STATE: return "invalid state for this call";
    case LMONO_E_DEVICE: return "device-side fault flag raised";
    default: return "unknown error";
  }
}

__global__ void k_state_init(LmMapState* st) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  st->cen[0] = 10; st->cen[1] = 10; st->cen[2] = 5;      // laserMapping.cpp:74-76
  st->center[0] = st->center[1] = st->center[2] = 0;
  st->valid_num = 0; st->from_map_n[0] = st->from_map_n[1] = 0; st->optimize = 0;
  st->stack_n[0] = st->stack_n[1] = 0; st->raw_n[0] = st->raw_n[1] = 0;
  st->frame_count = 0; st->fault = 0;
  st->shard_rank = 0; st->shard_n = 1; st->shard_owned_n[0] = st->shard_owned_n[1] = 0;
  for (int k = 0; k < 4; ++k) { st->q_wmap_wodom[k] = k == 3; st->q_wodom_curr[k] = k == 3; st->q_w_curr[k] = k == 3; }
  for (int k = 0; k < 3; ++k) { st->t_wmap_wodom[k] = 0; st->t_wodom_curr[k] = 0; st->t_w_curr[k] = 0; }
}

static int create_impl(lmono_ctx* ctx, void* stream);

// Failures after the ctx exists are routed through lmono_destroy (every member is calloc-zeroed, cudaFree(NULL) and the
// NULL-guarded destroys are no-ops), so a failed creation of the Nth sequence of a batch leaks nothing.
extern "C" int lmono_create(int device, const lmono_params* params, void* stream, lmono_ctx** out) {
  if (!out) return LMONO_E_ARG;
  *out = nullptr;
  lmono_ctx* ctx = (lmono_ctx*)calloc(1, sizeof(lmono_ctx));
  if (!ctx) return LMONO_E_ARG;
  lmono_params def; lmono_default_params(&def);
  ctx->prm = params ? *params : def;
  lmono_params& P = ctx->prm;
  if (P.scan_line == 0) P.scan_line = def.scan_line;
  if (P.mapping_line_resolution <= 0) P.mapping_line_resolution = def.mapping_line_resolution;
  if (P.mapping_plane_resolution <= 0) P.mapping_plane_resolution = def.mapping_plane_resolution;
  if (P.mapping_skip_frame <= 0) P.mapping_skip_frame = 1;
  if (P.max_sweep_points <= 0) P.max_sweep_points = def.max_sweep_points;
  if (P.max_feature_points <= 0) P.max_feature_points = def.max_feature_points;
  if (P.cube_capacity_corner <= 0) P.cube_capacity_corner = def.cube_capacity_corner;
  if (P.cube_capacity_surf <= 0) P.cube_capacity_surf = def.cube_capacity_surf;
  if (P.max_cubes_corner <= 0) P.max_cubes_corner = def.max_cubes_corner;
  if (P.max_cubes_surf <= 0) P.max_cubes_surf = def.max_cubes_surf;
  if (P.image_width <= 0) P.image_width = def.image_width;
  if (P.image_height <= 0) P.image_height = def.image_height;
  // every argument check comes before the first device allocation
  if (P.mapping_line_resolution < 0.06f || P.mapping_plane_resolution < 0.06f) { free(ctx); return LMONO_E_ARG; }
  if (P.scan_line != 16 && P.scan_line != 32 && P.scan_line != 64) { free(ctx); return LMONO_E_ARG; }
  if (P.max_feature_points > (1 << 19)) { fprintf(stderr, "[lmono_b200] max_feature_points must be <= 524288\n"); free(ctx); return LMONO_E_ARG; }
  if (P.max_cubes_corner > 8190 || P.max_cubes_surf > 8190) { free(ctx); return LMONO_E_ARG; }
  static unsigned long long next_uid = 0;
  ctx->uid = __atomic_add_fetch(&next_uid, 1ull, __ATOMIC_RELAXED);
  ctx->device = device;
  ctx->batch_n = 1; ctx->batch_hint = 1;
  { const char* ng = getenv("LMONO_NO_GRAPH"); ctx->graphs_on = !(ng && ng[0] == '1'); }
  ctx->max_feat = P.max_feature_points; ctx->max_sweep = P.max_sweep_points;
  int rc = create_impl(ctx, stream);
  if (rc) { lmono_destroy(ctx); return rc; }
  *out = ctx;
  return LMONO_OK;
}

static int create_impl(lmono_ctx* ctx, void* stream) {
  const int device = ctx->device;
  lmono_params& P = ctx->prm; (void)P;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { fprintf(stderr, "[lmono_b200] cudaSetDevice(%d): %s\n", device, cudaGetErrorString(e)); return LMONO_E_CUDA; }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return LMONO_E_CUDA;
  ctx->sm_count = prop.multiProcessorCount;
  if (stream) { ctx->stream = (cudaStream_t)stream; ctx->own_stream = false; }
  else { LM_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
  LM_CUDA(cudaEventCreate(&ctx->ev0)); LM_CUDA(cudaEventCreate(&ctx->ev1)); LM_CUDA(cudaEventCreate(&ctx->ev_k0)); LM_CUDA(cudaEventCreate(&ctx->ev_o0)); LM_CUDA(cudaEventCreate(&ctx->ev_o1));
  LM_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  LM_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
  LM_CUDA(cudaEventCreateWithFlags(&ctx->ev_sync, cudaEventDisableTiming));
  LM_CUDA(cudaEventCreateWithFlags(&ctx->ev_done, cudaEventDisableTiming));
  LM_CUDA(cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
  LM_CUDA(cudaEventCreateWithFlags(&ctx->ev_side0, cudaEventDisableTiming));
  LM_CUDA(cudaEventCreateWithFlags(&ctx->ev_side1, cudaEventDisableTiming));

  LM_CUDA(cudaMalloc((void**)&ctx->d_state, sizeof(LmMapState)));
  LM_CUDA(cudaMallocHost((void**)&ctx->h_state, sizeof(LmMapState)));
  for (int i = 0; i < 2; ++i) {
    LM_CUDA(cudaHostAlloc((void**)&ctx->h_ring[i], sizeof(LmMapState), cudaHostAllocMapped | cudaHostAllocPortable));
    memset(ctx->h_ring[i], 0, sizeof(LmMapState));
    LM_CUDA(cudaEventCreateWithFlags(&ctx->ev_res[i], cudaEventDisableTiming));
  }
  ctx->n_submitted = ctx->n_waited = 0;
  { const char* nz = getenv("LMONO_NO_ZEROCOPY"); ctx->zero_copy_on = !(nz && nz[0] == '1'); }
  LM_CUDA(cudaMalloc((void**)&ctx->d_lm, sizeof(LmLmState)));
  LM_CUDA(cudaMemsetAsync(ctx->d_lm, 0, sizeof(LmLmState), ctx->stream));
  LM_CUDA(cudaMalloc((void**)&ctx->d_slot_valid_rank, sizeof(int32_t) * LM_SLOT_TABLE_INTS));
  LM_CUDA(cudaMemsetAsync(ctx->d_slot_valid_rank, 0xFF, sizeof(int32_t) * LM_SLOT_TABLE_INTS, ctx->stream));
  LM_CUDA(cudaMalloc((void**)&ctx->d_partials, sizeof(double) * 32 * (1024 + 8)));
  LM_CUDA(cudaMalloc((void**)&ctx->d_stamps, sizeof(unsigned long long) * 4096));
  LM_CUDA(cudaMemsetAsync(ctx->d_stamps, 0, sizeof(unsigned long long) * 4096, ctx->stream));
  LM_CUDA(cudaMalloc((void**)&ctx->d_tl, sizeof(unsigned long long) * LM_TL_MAX));
  const size_t nf = (size_t)ctx->max_feat;
  ctx->raw_bytes = (size_t)ctx->max_sweep * 32;
  for (int i = 0; i < 3; ++i) LM_CUDA(cudaMalloc((void**)&ctx->d_raw[i], ctx->raw_bytes));
  for (int i = 0; i < 2; ++i) {
    LM_CUDA(cudaMalloc((void**)&ctx->d_in[i], nf * sizeof(float4)));
    LM_CUDA(cudaMalloc((void**)&ctx->d_stack[i], nf * sizeof(float4)));
    LM_CUDA(cudaMalloc((void**)&ctx->d_world[i], nf * sizeof(float4)));
    LM_CUDA(cudaMalloc((void**)&ctx->d_fac[i], nf * sizeof(LmFactor)));
  }
  LM_CUDA(cudaMalloc((void**)&ctx->d_nnref, sizeof(int32_t) * 5 * 2 * nf));
  const size_t nsort = (size_t)(2 * ctx->max_feat > ctx->max_sweep ? 2 * ctx->max_feat : ctx->max_sweep);
  LM_CUDA(cudaMalloc((void**)&ctx->d_sort_a, nsort * sizeof(unsigned long long)));
  LM_CUDA(cudaMalloc((void**)&ctx->d_sort_b, nsort * sizeof(unsigned long long)));
  LM_CUDA(cudaMalloc((void**)&ctx->d_sort_c, nsort * sizeof(unsigned long long)));
  LM_CUDA(cudaMalloc((void**)&ctx->d_blockcnt, sizeof(int32_t) * (nsort / 256 + 64)));
  LM_CUDA(cudaMalloc((void**)&ctx->d_sort_done, sizeof(int32_t) * 8));
  LM_CUDA(cudaMalloc((void**)&ctx->d_tmp_i32, sizeof(int32_t) * (2 * LM_NSLOT + 64 + nsort)));
  LM_CUDA(cudaMalloc((void**)&ctx->d_vg, sizeof(VgParams) * 4));
  { int rcv = lm_voxel_init(ctx); if (rcv) return rcv; }
  LM_CUDA(cudaMalloc((void**)&ctx->d_full, (size_t)ctx->max_sweep * sizeof(float4)));
  LM_CUDA(cudaMalloc((void**)&ctx->d_slot_first, sizeof(int32_t) * 2 * LM_NSLOT));
  LM_CUDA(cudaMalloc((void**)&ctx->d_slot_base, sizeof(int32_t) * 2 * LM_NSLOT));
  LM_CUDA(cudaMalloc((void**)&ctx->d_export_off, sizeof(int32_t) * (4 * LM_NSLOT + 64)));
  ctx->export_cap = 1 << 20;
  LM_CUDA(cudaMalloc((void**)&ctx->d_export, ctx->export_cap * sizeof(float4)));
  k_state_init<<<1, 32, 0, ctx->stream>>>(ctx->d_state);
  LM_LAUNCH_CHECK();
  int rc = LMONO_OK;
  if (ctx->prm.stages == 0 || (ctx->prm.stages & LMONO_STAGE_MAPPING)) {     // a scan / odometry / colour-only ctx does not pay for the cube map
    rc = lm_map_alloc(ctx);
    if (rc) return rc;
    rc = lm_map_configure_kernels(ctx);
    if (rc) return rc;
    ctx->map_ready = true;
  }
  if ((rc = lm_sort_configure(ctx))) return rc;
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  return LMONO_OK;
}

extern "C" void lmono_destroy(lmono_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < ctx->n_graphs; ++i) cudaGraphExecDestroy(ctx->graphs[i].exec);
  for (int i = 0; i < ctx->n_sweep_graphs; ++i) cudaGraphExecDestroy(ctx->sweep_graphs[i].exec);
  lm_batch_free(ctx);
  lm_shard_free(ctx);
  lm_scan_free(ctx);
  lm_odom_free(ctx);
  lm_color_free(ctx);
  lm_map_free(ctx);
  cudaFree(ctx->d_state); cudaFreeHost(ctx->h_state); for (int i = 0; i < 2; ++i) { cudaFreeHost(ctx->h_ring[i]); if (ctx->ev_res[i]) cudaEventDestroy(ctx->ev_res[i]); } cudaFree(ctx->d_lm); cudaFree(ctx->d_slot_valid_rank); cudaFree(ctx->d_partials); cudaFree(ctx->d_stamps); cudaFree(ctx->d_tl); cudaFree(ctx->d_nnref); cudaFree(ctx->d_rf_nvx); cudaFree(ctx->d_rf_tlb); cudaFree(ctx->d_rf_work); cudaFree(ctx->d_rf_meta); cudaFree(ctx->d_rf_plan); cudaFree(ctx->d_rf_big_s); cudaFree(ctx->d_rf_big_nv);
  for (int i = 0; i < 3; ++i) cudaFree(ctx->d_raw[i]);
  for (int i = 0; i < 2; ++i) { cudaFree(ctx->d_in[i]); cudaFree(ctx->d_stack[i]); cudaFree(ctx->d_world[i]); cudaFree(ctx->d_fac[i]); }
  cudaFree(ctx->d_sort_a); cudaFree(ctx->d_sort_b); cudaFree(ctx->d_sort_c); cudaFree(ctx->d_blockcnt); cudaFree(ctx->d_sort_done); cudaFree(ctx->d_tmp_i32);
  cudaFree(ctx->d_vg); cudaFree(ctx->d_full); cudaFree(ctx->d_slot_first); cudaFree(ctx->d_slot_base); cudaFree(ctx->d_export_off); cudaFree(ctx->d_export);
  cudaEvent_t evs[] = { ctx->ev0, ctx->ev1, ctx->ev_k0, ctx->ev_o0, ctx->ev_o1, ctx->ev_fork, ctx->ev_join, ctx->ev_sync, ctx->ev_done, ctx->ev_side0, ctx->ev_side1 };
  for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
  if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
  if (ctx->sweep_stream) cudaStreamDestroy(ctx->sweep_stream);
  if (ctx->ev_sweep) cudaEventDestroy(ctx->ev_sweep);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  cudaGetLastError();
  free(ctx);
}

extern "C" int lmono_sync(lmono_ctx* ctx) {
  if (!ctx) return LMONO_E_ARG;
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  return LMONO_OK;
}

extern "C" int lmono_debug_stamps(lmono_ctx* ctx, uint64_t* out, int32_t n) {
  if (!ctx || !out || n < 0 || n > 4096) return LMONO_E_ARG;
  LM_CUDA(cudaMemcpyAsync(out, ctx->d_stamps, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  return LMONO_OK;
}

// device time of the most recent lmono_scan_register / lmono_odom_step / lmono_project_color call: whole call on the
// stream (uploads + kernels) and kernels only (inputs resident in HBM)
extern "C" int lmono_stage_times(lmono_ctx* ctx, float* ms_with_uploads, float* ms_kernels) {
  if (!ctx) return LMONO_E_ARG;
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  float a = 0.f, b = 0.f;
  if (cudaEventElapsedTime(&a, ctx->ev0, ctx->ev1) != cudaSuccess) { cudaGetLastError(); a = 0.f; }
  if (cudaEventElapsedTime(&b, ctx->ev_k0, ctx->ev1) != cudaSuccess) { cudaGetLastError(); b = 0.f; }
  if (ms_with_uploads) *ms_with_uploads = a;
  if (ms_kernels) *ms_kernels = b;
  return LMONO_OK;
}

extern "C" int64_t lmono_launch_count(const lmono_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int lmono_last_fault(lmono_ctx* ctx, uint32_t* bits) {
  if (!ctx || !bits) return LMONO_E_ARG;
  LM_CUDA(cudaMemcpyAsync(&ctx->h_state->fault, &ctx->d_state->fault, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  *bits = ctx->h_state->fault;
  LM_CUDA(cudaMemsetAsync(&ctx->d_state->fault, 0, sizeof(uint32_t), ctx->stream));
  return LMONO_OK;
}

// ---- cloud marshalling ---------------------------------------------------------------
__global__ void __launch_bounds__(256) k_unpack_cloud(const uint8_t* __restrict__ raw, int n, int stride, int ioff, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = reinterpret_cast<const float*>(raw + (size_t)i * stride);
  float4 v; v.x = p[0]; v.y = p[1]; v.z = p[2];
  v.w = ioff >= 0 ? *reinterpret_cast<const float*>(raw + (size_t)i * stride + ioff) : 0.0f;
  out[i] = v;
}

int lm_upload_cloud(lmono_ctx* ctx, lmono_cloud_view v, uint8_t* d_raw, float4* d_out, int32_t* d_n) {
  if (v.n < 0 || (v.n > 0 && !v.base) || (v.n > 0 && (v.stride_bytes < 12 || (v.stride_bytes & 3)))) return LMONO_E_ARG;
  if (d_n) LM_CUDA(cudaMemcpyAsync(d_n, &v.n, sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  if (v.n == 0) return LMONO_OK;
  if (v.stride_bytes == 16 && v.intensity_offset == 12) {
    LM_CUDA(cudaMemcpyAsync(d_out, v.base, (size_t)v.n * 16, cudaMemcpyHostToDevice, ctx->stream));
    return LMONO_OK;
  }
  if ((size_t)v.n * v.stride_bytes > ctx->raw_bytes) return LMONO_E_CAPACITY;
  LM_CUDA(cudaMemcpyAsync(d_raw, v.base, (size_t)v.n * v.stride_bytes, cudaMemcpyHostToDevice, ctx->stream));
  k_unpack_cloud<<<lm_div_up(v.n, 256), 256, 0, ctx->stream>>>(d_raw, v.n, v.stride_bytes, v.intensity_offset, d_out);
  LM_LAUNCH_CHECK();
  return LMONO_OK;
}

// synchronous: copies n packed float4 points into the caller's strided buffer
int lm_download_cloud(lmono_ctx* ctx, const float4* d_src, int n, lmono_cloud_out* out) {
  out->n_out = n;
  if (n > out->capacity) return LMONO_E_CAPACITY;
  if (n == 0) return LMONO_OK;
  if (!out->base || out->stride_bytes < 12) return LMONO_E_ARG;
  if (out->stride_bytes == 16 && out->intensity_offset == 12) {
    LM_CUDA(cudaMemcpyAsync(out->base, d_src, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    LM_CUDA(cudaStreamSynchronize(ctx->stream));
    return LMONO_OK;
  }
  float4* tmp = (float4*)malloc((size_t)n * sizeof(float4));
  if (!tmp) return LMONO_E_CAPACITY;
  cudaError_t e = cudaMemcpyAsync(tmp, d_src, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) { free(tmp); ctx->last_cuda_error = (int)e; return LMONO_E_CUDA; }
  for (int i = 0; i < n; ++i) {
    uint8_t* rec = (uint8_t*)out->base + (size_t)i * out->stride_bytes;
    memcpy(rec, &tmp[i], 12);
    if (out->intensity_offset >= 0) memcpy(rec + out->intensity_offset, &tmp[i].w, 4);
  }
  free(tmp);
  return LMONO_OK;
}
