// Cube-sharded global map, peer-memory mode (BASELINE config C-5).
//
// The NCCL mode (mapping.cu: lmono_shard_begin ... lmono_shard_end) cuts a registration into kernels around 11
// host-issued all-reduces of 35 doubles.  Here the exchange is part of the kernels: every rank owns an exchange block
// (LmShardXchg, common.cuh) in its HBM that is mapped into all other ranks -- cudaIpc handles between the
// one-process-per-GPU ranks, NVLink 5 / NVSwitch underneath, or plain device pointers when the ranks are contexts of
// one process -- and a kernel that needs the sum over the ranks stores its partial into every peer's block, raises a
// flag with release semantics and polls its own block (d_shard_exchange).  A sharded registration is then the SAME
// kernel sequence as an unsharded one (enqueue_step in mapping.cu, replayed as one CUDA graph): k_shard_gate_xchg makes
// the laserMapping.cpp:554 gate global, and k_lm_solve_cluster (lm.cu) all-gathers the 30-vector
// {J^T J, J^T r, cost, counts} after every evaluation and sums it in rank order, so all ranks advance identical
// trust-region controllers.  No host call, no NCCL launch and no stream synchronisation inside a registration.
#include "common.cuh"
#include <string.h>

__global__ void k_shard_config2(LmMapState* st, int rank, int n) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { st->shard_rank = rank; st->shard_n = n; }
}

// laserMapping.cpp:554 on the GLOBAL window content: each rank contributes the points of the window cubes it OWNS
// (halo copies are not counted twice); one CTA
__global__ void __launch_bounds__(64) k_shard_gate_xchg(LmMapState* __restrict__ st, const LmShardPeers* __restrict__ peers) {
  __shared__ double s_io[LM_XCHG_DOUBLES];
  if (threadIdx.x < LM_XCHG_DOUBLES) s_io[threadIdx.x] = 0.0;
  __syncthreads();
  if (threadIdx.x == 0) { s_io[30] = (double)st->shard_owned_n[0]; s_io[31] = (double)st->shard_owned_n[1]; }
  __syncthreads();
  LmShardXchg* mine = peers->peer[peers->rank];
  const unsigned long long epoch = mine->epoch + 1ull;
  d_shard_exchange(peers, epoch, s_io, s_io, 32, true, &st->fault);
  if (threadIdx.x == 0) {
    st->from_map_n[0] = (int)s_io[30]; st->from_map_n[1] = (int)s_io[31];
    st->optimize = (s_io[30] > 10.0 && s_io[31] > 50.0) ? 1 : 0;
    mine->epoch = epoch;
  }
}

int lm_shard_gate_xchg(lmono_ctx* ctx) {
  k_shard_gate_xchg<<<1, 64, 0, ctx->stream>>>(ctx->d_state, ctx->d_shard_peers);
  LM_LAUNCH_CHECK();
  return LMONO_OK;
}

void lm_shard_free(lmono_ctx* ctx) {
  for (int r = 0; r < LM_SHARD_MAX; ++r) if (ctx->xchg_opened[r]) { cudaIpcCloseMemHandle(ctx->xchg_opened[r]); ctx->xchg_opened[r] = nullptr; }
  cudaFree(ctx->d_xchg); ctx->d_xchg = nullptr;
  cudaFree(ctx->d_shard_peers); ctx->d_shard_peers = nullptr;
  ctx->shard_p2p = false;
}

// Allocates this rank's exchange block (zeroed) and returns its cudaIpcMemHandle_t (64 bytes) for the peers, and the
// local device pointer (for peers that live in the same process).
extern "C" int lmono_shard_xchg_create(lmono_ctx* ctx, void* ipc_handle_out /*[64]*/, void** local_ptr_out) {
  if (!ctx) return LMONO_E_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "include/lmono.h documents a 64-byte handle");
  LM_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->d_xchg) {
    // a dedicated allocation: cudaIpcGetMemHandle exports the whole underlying allocation
    LM_CUDA(cudaMalloc((void**)&ctx->d_xchg, sizeof(LmShardXchg)));
    LM_CUDA(cudaMalloc((void**)&ctx->d_shard_peers, sizeof(LmShardPeers)));
  }
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  LM_CUDA(cudaMemset(ctx->d_xchg, 0, sizeof(LmShardXchg)));
  LM_CUDA(cudaDeviceSynchronize());
  if (ipc_handle_out) {
    cudaIpcMemHandle_t h;
    LM_CUDA(cudaIpcGetMemHandle(&h, ctx->d_xchg));
    memcpy(ipc_handle_out, &h, sizeof(h));
  }
  if (local_ptr_out) *local_ptr_out = ctx->d_xchg;
  return LMONO_OK;
}

// Maps the peers' exchange blocks and switches the ctx to the peer-memory mode.  ipc_handles: [nranks][64] as returned
// by lmono_shard_xchg_create on every rank (gathered by the host over any transport); same_process_ptrs (may be NULL):
// entry r != NULL is rank r's local pointer and is used instead of its handle (ranks that are contexts of this process
// on this device -- cudaIpcOpenMemHandle refuses handles of the calling process).
extern "C" int lmono_shard_xchg_open(lmono_ctx* ctx, int32_t rank, int32_t nranks, const void* ipc_handles, void* const* same_process_ptrs) {
  if (!ctx || nranks < 1 || nranks > LM_SHARD_MAX || rank < 0 || rank >= nranks || !ctx->d_xchg) return LMONO_E_ARG;
  if (nranks > 1 && !ipc_handles && !same_process_ptrs) return LMONO_E_ARG;
  LM_CUDA(cudaSetDevice(ctx->device));
  LmShardPeers P;
  memset(&P, 0, sizeof(P));
  P.rank = rank; P.n = nranks;
  for (int r = 0; r < nranks; ++r) {
    if (r == rank) { P.peer[r] = ctx->d_xchg; continue; }
    if (same_process_ptrs && same_process_ptrs[r]) { P.peer[r] = (LmShardXchg*)same_process_ptrs[r]; continue; }
    if (!ipc_handles) return LMONO_E_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)ipc_handles + (size_t)r * sizeof(h), sizeof(h));
    void* p = nullptr;
    LM_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->xchg_opened[r] = p;
    P.peer[r] = (LmShardXchg*)p;
  }
  LM_CUDA(cudaMemcpyAsync(ctx->d_shard_peers, &P, sizeof(P), cudaMemcpyHostToDevice, ctx->stream));
  k_shard_config2<<<1, 32, 0, ctx->stream>>>(ctx->d_state, rank, nranks);
  LM_LAUNCH_CHECK();
  LM_CUDA(cudaStreamSynchronize(ctx->stream));      // P lives on this stack frame
  // graphs captured before the switch do not contain the exchange kernels
  for (int i = 0; i < ctx->n_graphs; ++i) cudaGraphExecDestroy(ctx->graphs[i].exec);
  ctx->n_graphs = 0;
  ctx->shard_p2p = nranks > 1;
  return LMONO_OK;
}

// exchanges completed, nanoseconds the publishing CTAs spent posting + waiting for the slowest rank, timeouts (fault bit)
extern "C" int lmono_shard_xchg_stats(lmono_ctx* ctx, uint64_t out[3], int32_t reset) {
  if (!ctx || !out || !ctx->d_xchg) return LMONO_E_ARG;
  LmShardXchg h;
  LM_CUDA(cudaStreamSynchronize(ctx->stream));
  LM_CUDA(cudaMemcpy(&h.epoch, &ctx->d_xchg->epoch, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  out[0] = h.epoch; out[1] = h.wait_ns; out[2] = h.n_xchg;
  if (reset) LM_CUDA(cudaMemset(&ctx->d_xchg->wait_ns, 0, 2 * sizeof(unsigned long long)));
  return LMONO_OK;
}
