// lmono_b200 internal header: device data layout, shared device helpers, host ctx.
// Target: sm_100a (B200).  Compiled with -fmad=false so that every fp32/fp64 product-sum
// is rounded exactly like the reference's x86-64 (no FMA) arithmetic -- feature labels,
// voxel keys, kNN distances and fit gates are bit-exact against the oracle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "../../include/lmono.h"

// ---------------------------------------------------------------- constants
// rolling cube grid of the reference (Aloam/src/laserMapping.cpp:74-82)
constexpr int LM_GW = 21, LM_GH = 21, LM_GD = 11;
constexpr int LM_NSLOT = LM_GW * LM_GH * LM_GD;   // 4851
constexpr int LM_MAX_VALID = 125;                 // :85 (<=75 used)
// per-cube search grid: 2 m cells keyed on floor(x)>>1; a 50 m cube (+1 floor of slack on
// each side) spans 26 cells per axis.
constexpr int LM_CELLS_AXIS = 26;
constexpr int LM_NCELL = LM_CELLS_AXIS * LM_CELLS_AXIS * LM_CELLS_AXIS;   // 17576
constexpr int LM_SORT_TILE = 1024;                // elements per CTA in the global tile sort (1024: twice the CTAs of 2048, each less than half the network)
constexpr int LM_SORT_MAXSEG = 2;                 // independent segments sorted by one pair of launches
constexpr int LM_TAIL_TILE = 16384;               // max points of a slab the whole-slab fallback refilter can re-voxelise (smem sort)
constexpr int LM_RF_CHUNK = 1024;                 // points per CTA in the chunked refilter kernels
constexpr int LM_WIN_MAX = 75;                    // cubes of the 5x5x3 window
// d_rf_plan layout: counts, then three lists of (type * LM_WIN_MAX + window rank).  The per-step kernels that only a few
// cubes need (index build, whole-slab re-voxelisation, tail merge) run on small grids that walk these lists instead of
// one (mostly idle) CTA per window cube: with many sequences per GPU the idle CTAs of all of them queue for SM slots.
constexpr int LM_PLAN_DIRTY_N = 0, LM_PLAN_WHOLE_N = 1, LM_PLAN_ACTIVE_N = 2;
constexpr int LM_PLAN_DIRTY = 16, LM_PLAN_WHOLE = 16 + 2 * LM_WIN_MAX, LM_PLAN_ACTIVE = 16 + 4 * LM_WIN_MAX, LM_PLAN_INTS = 16 + 6 * LM_WIN_MAX;

// device fault bits (LmMapState::fault)
enum : unsigned {
  LM_FAULT_CUBE_OVERFLOW = 1u << 0,   // a cube slab exceeded its capacity
  LM_FAULT_POOL_EXHAUSTED = 1u << 1,  // no free cube slab
  LM_FAULT_TAIL_OVERFLOW = 1u << 2,   // unsorted tail larger than LM_TAIL_TILE
  LM_FAULT_CELL_RANGE = 1u << 3,      // point outside its cube's cell table (internal error)
  LM_FAULT_FEATURE_OVERFLOW = 1u << 4,
  LM_FAULT_IMPORT_NONEMPTY = 1u << 5,
  LM_FAULT_SHARD_TIMEOUT = 1u << 6,   // a peer rank did not post its partial sums within LM_XCHG_TIMEOUT_NS
};

// ---------------------------------------------------------------- cube-sharded map: peer-memory exchange block
// One per rank, in that rank's HBM, mapped into every peer.  An exchange with sequence number `epoch` (1, 2, ...; the
// same on every rank because all ranks run the same kernel sequence) uses slot epoch & 1: rank s stores its partial
// sums into payload[slot][s] of EVERY rank (stores over NVLink), fences, then stores `epoch` into flag[slot][s] of every
// rank with release semantics; a rank acquires its OWN flag[slot][0..n) (local polling, no NVLink reads) and sums the
// payloads in rank order -> every rank gets the same bits.  Two slots suffice because a rank posts epoch e+1 only after
// all its CTAs consumed epoch e (cluster barrier / kernel boundary between consecutive exchanges).
constexpr int LM_SHARD_MAX = 16;
constexpr int LM_XCHG_DOUBLES = 40;
constexpr unsigned long long LM_XCHG_TIMEOUT_NS = 2000000000ull;   // a dead peer raises a fault instead of hanging the GPU
struct LmShardXchg {
  unsigned long long flag[2][LM_SHARD_MAX];
  double payload[2][LM_SHARD_MAX][LM_XCHG_DOUBLES];
  unsigned long long epoch;          // exchanges completed by this rank (advanced by the last kernel that exchanged)
  unsigned long long wait_ns;        // time the publisher spent waiting for the slowest peer (statistics)
  unsigned long long n_xchg;
};
struct LmShardPeers { LmShardXchg* peer[LM_SHARD_MAX]; int32_t rank, n; };

// ---------------------------------------------------------------- device structs
struct LmFactor {           // 64 B, one per query (kind < 0: no factor)
  double a[3];              // EDGE: point_a | PLANE: point_j | PLANE_NORM: unit normal
  double b[3];              // EDGE: point_b | PLANE: unit normal ljm | PLANE_NORM: b[0] = negative_OA_dot_norm
  float p[3];               // curr_point (sensor frame), stored as the float it is
  int32_t kind;             // -1 none, 0 edge, 1 plane, 2 plane_norm
};
static_assert(sizeof(LmFactor) == 64, "LmFactor must be 64 bytes");

struct LmSolveSummary { int32_t iterations, num_successful, termination, num_factors; double initial_cost, final_cost; };

struct LmLmState {          // Levenberg-Marquardt controller state (one solve at a time)
  double x[7];              // accepted parameters q(xyzw), t
  double cand[7];           // candidate being evaluated
  double x_norm, cost;
  double H[21], g[6];       // upper-triangular J^T J and J^T r at x (unscaled tangent space)
  double scaling[6], diagonal[6];
  double radius, decrease_factor, model_cost_change;
  int32_t reuse_diagonal, num_invalid;
  int32_t iteration, num_successful, termination, done, phase, max_iter;
  int32_t nfactors, pad;
  double initial_cost;
  // scratch for the grid reduction
  unsigned int ticket;
  unsigned int pad2;
  // reciprocals kept beside radius / model_cost_change so that the controller's dependent chain (step quality -> radius ->
  // damped normal equations -> Cholesky -> candidate) carries one fp64 division less per pass each
  double inv_radius, inv_model_cost_change;
};

struct LmProblem {          // one LM solve: factor arrays, pose in/out, report slots (device pointers)
  const LmFactor* fac0; const LmFactor* fac1;
  const int32_t* n0; const int32_t* n1;   // factor-slot counts (invalid slots have kind < 0)
  const int32_t* gate;                    // solve only if *gate != 0 (NULL = always)
  double* pose_q; double* pose_t;         // q (x,y,z,w), t: initial value in, result out
  LmSolveSummary* summary;
  int32_t* count0; int32_t* count1;       // valid factors per array
  // DISTORTION 1 (laserOdometry.cpp:59): fractional part of the intensity of every factor's point (its interpolation ratio is
  // s = frac / SCAN_PERIOD, lidarFactor.hpp:14,60); NULL = every factor has s = 1
  const float* frac0; const float* frac1;
};

struct alignas(16) LmMapState {
  int32_t cen[3];                        // laserCloudCenWidth/Height/Depth
  int32_t center[3];                     // centerCubeI/J/K after the shift
  int32_t valid_num;
  int32_t valid_slot[LM_MAX_VALID + 3];  // physical slot of each window cube, reference order
  int32_t valid_off[2][LM_MAX_VALID + 3];// exclusive prefix of cube sizes per map type
  int32_t from_map_n[2];
  int32_t optimize;
  int32_t stack_n[2];                    // voxel-filtered feature counts (corner, surf)
  int32_t raw_n[2];
  int32_t frame_count;
  uint32_t fault;
  int32_t corner_num[2], surf_num[2];
  LmSolveSummary solve[2];
  double q_wmap_wodom[4], t_wmap_wodom[3];
  double q_wodom_curr[4], t_wodom_curr[3];
  double q_w_curr[4], t_w_curr[3];
  // cube-sharded global map (shard.cu): shard_n <= 1 means "not sharded, every cube is mine"
  int32_t shard_rank, shard_n;
  int32_t shard_owned_n[2];              // points of OWNED window cubes (summed over ranks for the :554 gate)
  // device pointers of this step's corner / surf features (written by k_step_args / k_batch_args), so that the
  // captured kernel sequence does not depend on where the caller keeps its inputs
  const float4* in_ptr[2];
  // fused upload: in_stride != 0 -> this step's features are fetched by k_vg_keys straight from page-locked host
  // memory (in_src = device alias of the caller's AoS buffer, record stride / intensity offset in bytes) into in_ptr
  const void* in_src[2];
  int32_t in_stride[2], in_ioff[2];
  int32_t result_slot;                   // which pinned result mirror (lmono_ctx::h_ring) k_publish_state writes
  int32_t pad_[3];
};
static_assert(sizeof(LmMapState) % 16 == 0, "k_publish_state copies 16-byte words");

struct LmMapType {           // one per map (0 corner, 1 surf); device pointers, passed by value
  float leaf, inv_leaf;
  int32_t cap;               // points per cube slab
  int32_t n_slabs;
  float4* pts;               // [n_slabs][2][cap]  canonical (VoxelGrid) order, ping-pong
  float4* cellpts;           // [n_slabs][cap]     cell-sorted copy, .w = bits of slab position
  uint32_t* cellstart;       // [n_slabs][LM_NCELL+1]
  uint32_t* pkey;            // [n_slabs][2][cap]  cube-local voxel key of every point of pts (same ping-pong buffer)
  int32_t* cellcount;        // [n_slabs][LM_NCELL] refilter scratch: cell histogram, then scatter cursors
  int32_t* slab_unsorted;    // [n_slabs] 1 = a centroid crossed a voxel border: the slab must be re-voxelised as a whole
  int32_t* slot_slab;        // [LM_NSLOT] slab id or -1
  int32_t* slab_n;           // [n_slabs] points in slab (sorted prefix + tail)
  int32_t* slab_nsorted;     // [n_slabs] length of the sorted, voxel-unique prefix
  int32_t* slab_cur;         // [n_slabs] current ping-pong buffer
  int32_t* slab_dirty;       // [n_slabs] 1 = cell index stale
  int32_t* slab_g;           // [n_slabs][4] absolute cube coordinate (gi,gj,gk) of the slab
  int32_t* free_stack;       // [n_slabs]
  int32_t* free_top;         // [1]
};

struct VgParams {            // per-cloud VoxelGrid side state (voxel.cu): bounding box for PCL's overflow guard
  uint32_t mn[3], mx[3];     // order-preserving encodings of the float min / max (atomicMin / atomicMax)
  uint32_t ticket;           // blocks of k_vg_write that finished (the last one re-arms mn / mx)
  int32_t n;
};

// profiler phases (lmono_profile_read order)
enum { LM_PROF_WINDOW = 0, LM_PROF_INDEX, LM_PROF_VOXEL, LM_PROF_ASSOC, LM_PROF_SOLVE, LM_PROF_INSERT, LM_PROF_REFILTER,
       LM_PROF_MISC, LM_PROF_NTAGS };
constexpr int LM_PROF_MAX_EVENTS = 2048;

// ---------------------------------------------------------------- host ctx
// One captured CUDA graph per launch-grid capacity bucket of a laserMapping step (inputs are read through
// LmMapState::in_ptr, so the graph does not depend on the caller's buffers).
struct LmGraphEntry { int nc_cap, ns_cap; int form; cudaGraphExec_t exec; int n_launch; };   // form: every host-side choice baked in at capture (lm_graph_form)
constexpr int LM_MAX_GRAPHS = 64;
// Sequence batches: ONE graph holds the steps of all n sequences as parallel branches (fork / join inside the
// graph), so a batch step costs the host one small argument launch + one cudaGraphLaunch.  Cached in ctxs[0].
constexpr int LM_BATCH_MAX = 64;
constexpr int LM_MAX_BGRAPHS = 16;
struct LmBatchGraph { int n; lmono_ctx* ctxs[LM_BATCH_MAX]; unsigned long long uids[LM_BATCH_MAX];   // uids: a ctx address can be reused after lmono_destroy, its uid never is
                       int nc_cap[LM_BATCH_MAX], ns_cap[LM_BATCH_MAX]; cudaGraphExec_t exec; int n_launch[LM_BATCH_MAX]; };
constexpr int LM_GRAPH_BUCKET = 2048;     // launch grids are sized for counts rounded up to this

constexpr int LM_TL_MAX = 64;
constexpr int LM_THROUGHPUT_BATCH = 4;
// kernel-form class of a step enqueued with `batch_n` sequences sharing the GPU: 0 latency forms, 1 throughput forms with the
// full LM cluster, 2 throughput forms with 8-CTA LM clusters (lm_cluster_pick).  Part of every graph cache key.
static inline int lm_graph_form(int batch_n) { return batch_n >= LM_THROUGHPUT_BATCH ? (batch_n > 8 ? 2 : 1) : 0; }
struct lmono_ctx {
  unsigned long long uid;       // process-wide unique, never reused (keys of graph caches held by OTHER ctxs)
  int device;
  lmono_params prm;
  cudaStream_t stream;
  bool own_stream;
  cudaStream_t side_stream; cudaEvent_t ev_side0, ev_side1;   // fork / join of the independent head of a captured step (window upkeep || feature VoxelGrid)
  int last_cuda_error;
  int64_t launches;
  cudaEvent_t ev0, ev1;
  // asynchronous fused sweep (lmono_sweep_submit / _wait): its own stream when the ctx shares the caller's stream with other sequences
  cudaStream_t sweep_stream; cudaEvent_t ev_sweep; int sweep_n_in, sweep_prev_n; bool sweep_outstanding;
  // the whole sweep (scanRegistration + laserOdometry + laserMapping) as ONE CUDA graph per raw-size bucket
  struct { int n_cap; int form; cudaGraphExec_t exec; int n_launch; } sweep_graphs[8]; int n_sweep_graphs;
  cudaEvent_t ev_o0, ev_o1;     // odometry stage of a fused sweep (lmono_sweep_step)
  cudaEvent_t ev_k0;            // after the uploads of a stage call: ev_k0 .. ev1 = the stage's kernels with inputs resident in HBM (lmono_stage_times)
  cudaEvent_t ev_fork;          // fork point of a sequence batch (lmono_map_step_device_batch)
  cudaEvent_t ev_join;          // end of this ctx's branch of a batch (capture join, or the legacy multi-stream join)
  cudaEvent_t ev_sync;          // orders a batch graph after earlier work on this ctx's own stream
  cudaEvent_t ev_done;          // (batch leader) the batch graph has finished on the origin stream
  cudaStream_t* cap_streams;    // (batch leader) [LM_BATCH_MAX] branch streams, used only while capturing a batch graph
  LmBatchGraph* bgraphs; int n_bgraphs;      // (batch leader) [LM_MAX_BGRAPHS]
  bool step_timed;              // ev0 / ev1 bracket the pending step (false for steps that ran inside a batch graph)
  int batch_hint;               // lmono_set_concurrency_hint: what batch_n falls back to outside a batch call (default 1)
  int batch_n;                  // sequences sharing the GPU in the step being enqueued (1 = alone): >= LM_THROUGHPUT_BATCH picks the
                                // throughput forms of the kernels (one-thread-per-query kNN, 8-CTA LM clusters) over the latency forms

  LmMapType map[2];
  bool map_ready;               // the slab pools exist (lmono_params::stages includes LMONO_STAGE_MAPPING)
  LmMapState* d_state;
  LmMapState* h_state;          // pinned mirror
  // pipelined host API (lmono_map_submit_batch / _wait_batch): up to two steps of a ctx may be in flight; the step
  // graph publishes the state into h_ring[slot] (mapped page-locked memory) and ev_res[slot] marks its completion
  LmMapState* h_ring[2]; cudaEvent_t ev_res[2];
  uint32_t n_submitted, n_waited;
  bool zero_copy_on;            // LMONO_NO_ZEROCOPY=1: always stage host inputs through cudaMemcpyAsync
  LmLmState* d_lm;
  int32_t* d_slot_valid_rank;   // [LM_NSLOT]
  double* d_partials;           // reduction partials [max_blocks][32]

  // feature staging
  uint8_t* d_raw[3]; size_t raw_bytes;     // raw AoS uploads (corner, surf, full)
  float4* d_in[2];                          // unpacked XYZI inputs
  float4* d_stack[2];                       // voxel-filtered features
  float4* d_world[2];                       // features in the world frame (insertion)
  LmFactor* d_fac[2];
  int32_t* d_nnref;                         // [2 * max_feat][5] neighbour references of the association pass (-1 = gate failed)
  unsigned long long* d_sort_a; unsigned long long* d_sort_b; unsigned long long* d_sort_c;
  int32_t* d_blockcnt;                      // block counts for 2-kernel scans
  int32_t* d_sort_done;                     // [LM_SORT_MAXSEG] sort.cu: the segment was merged by the single shared-memory pass
  int32_t* d_tmp_i32;                       // misc int scratch [max_feature_points*2]
  VgParams* d_vg;
  float4* d_full;                           // full-res sweep
  int32_t* d_slot_first; int32_t* d_slot_base; // insertion run tables [LM_NSLOT]
  int32_t* d_rf_nvx;                        // refilter scratch [2][75][max cap]: (new voxels before j) << 1 | (j starts a new voxel)
  int32_t* d_rf_tlb;                        // refilter scratch [2][75][max cap]: lower bound of a tail run head's key in the prefix
  int32_t* d_rf_work;                       // [4 + 2*75*chunks]: chunk work list of the cubes being refiltered
  int32_t* d_rf_meta;                       // [2][75][8]: active, total_new, ns, nt, unsorted flag, cur
  unsigned long long* d_rf_big_s; int32_t* d_rf_big_nv;   // global-memory sort scratch of k_refilter_whole for slabs above LM_TAIL_TILE points
  int32_t* d_rf_plan;                       // compact per-step work lists (LM_PLAN_*): dirty cubes, cubes to re-voxelise whole, cubes with a tail
  float4* d_export; size_t export_cap;      // export / import staging
  int32_t* d_export_off;                    // [LM_NSLOT+1]
  int max_feat, max_sweep;
  int sm_count;
  bool step_pending;
  // cube-sharded mode (shard.cu): caller-owned device workspace the host all-reduces between kernels
  double* d_shard_ws; int shard_nc, shard_ns;
  // peer-memory mode of the sharded map (shard.cu): every rank's exchange block is mapped into every other rank
  // (cudaIpc over NVLink, or plain pointers inside one process); the LM solve kernel all-gathers its partial sums itself
  bool shard_p2p; LmShardXchg* d_xchg; LmShardPeers* d_shard_peers; void* xchg_opened[LM_SHARD_MAX];
  // CUDA-graph replay of the laserMapping step (mapping.cu); LMONO_NO_GRAPH=1 disables it
  bool graphs_on; int n_graphs; LmGraphEntry graphs[LM_MAX_GRAPHS];
  // in-kernel %globaltimer stamps (ns) for latency studies (lmono_debug_stamps); 256 slots, written by thread 0 of CTA 0
  unsigned long long* d_stamps;
  // later stages (scan registration / odometry / colour) attach their own state
  void* scan_state; void* odom_state; void* color_state;
  // optional per-phase CUDA-event profiler (bench.py roofline numbers)
  bool prof_on; int prof_n; int prof_tag[LM_PROF_MAX_EVENTS]; cudaEvent_t prof_ev[LM_PROF_MAX_EVENTS][2];
  // per-launch marks (lmono_kmarks_*): one CUDA event after every kernel launch, keyed by the launch site
  bool kmark_on; int kmark_n; cudaEvent_t* kmark_ev; const char** kmark_file; int* kmark_line;
  // graph-compatible timeline (LMONO_TIMELINE=1, lmono_timeline_dump): a one-thread %globaltimer stamp kernel after every
  // launch, captured into the step / batch graphs, so concurrent sequences can be laid on one time axis
  bool tl_on; int tl_n; unsigned long long* d_tl; const char* tl_file[LM_TL_MAX]; int tl_line[LM_TL_MAX];
};

static inline void lm_prof_begin(lmono_ctx* ctx, int tag) {
  if (!ctx->prof_on || ctx->prof_n >= LM_PROF_MAX_EVENTS) return;
  ctx->prof_tag[ctx->prof_n] = tag;
  cudaEventRecord(ctx->prof_ev[ctx->prof_n][0], ctx->stream);
}
static inline void lm_prof_end(lmono_ctx* ctx) {
  if (!ctx->prof_on || ctx->prof_n >= LM_PROF_MAX_EVENTS) return;
  cudaEventRecord(ctx->prof_ev[ctx->prof_n][1], ctx->stream);
  ctx->prof_n++;
}

#define LM_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { ctx->last_cuda_error = (int)_e; \
  fprintf(stderr, "[lmono_b200] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e), __FILE__, __LINE__, cudaGetErrorString(_e)); \
  return LMONO_E_CUDA; } } while (0)
constexpr int LM_KMARK_MAX = 8192;
#ifdef __CUDACC__
// launched like the step kernels (programmatic dependent launch, see lm_pdl_enter below): the stamp is taken once the
// preceding kernel has completed, and the following kernel may already be resident
static __global__ void k_tl_stamp(unsigned long long* p) {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); *p = t;
}
static inline bool lm_pdl_on();
static inline void lm_tl(lmono_ctx* ctx, const char* file, int line) {
  if (!ctx->tl_on || !ctx->d_tl || ctx->tl_n >= LM_TL_MAX) return;
  ctx->tl_file[ctx->tl_n] = file; ctx->tl_line[ctx->tl_n] = line;
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1); cfg.blockDim = dim3(1); cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = lm_pdl_on() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, k_tl_stamp, ctx->d_tl + ctx->tl_n);
  }
  ctx->tl_n++;
}
#else
static inline void lm_tl(lmono_ctx*, const char*, int) {}
#endif
static inline void lm_kmark(lmono_ctx* ctx, const char* file, int line) {
  if (!ctx->kmark_on || ctx->kmark_n >= LM_KMARK_MAX) return;
  ctx->kmark_file[ctx->kmark_n] = file; ctx->kmark_line[ctx->kmark_n] = line;
  cudaEventRecord(ctx->kmark_ev[ctx->kmark_n], ctx->stream);
  ctx->kmark_n++;
}
#define LM_LAUNCH_CHECK() do { ctx->launches++; lm_kmark(ctx, __FILE__, __LINE__); lm_tl(ctx, __FILE__, __LINE__); cudaError_t _e = cudaGetLastError(); if (_e != cudaSuccess) { ctx->last_cuda_error = (int)_e; \
  fprintf(stderr, "[lmono_b200] launch error %s at %s:%d\n", cudaGetErrorName(_e), __FILE__, __LINE__); return LMONO_E_CUDA; } } while (0)

static inline int lm_div_up(int a, int b) { return (a + b - 1) / b; }
#define LM_NEED_MAP() do { if (!ctx->map_ready) return LMONO_E_STATE; } while (0)

// ---------------------------------------------------------------- programmatic dependent launch
// A registration is a chain of ~20 small dependent kernels.  Kernels on the step path are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and start with griddepcontrol.wait (lm_pdl_enter): the successor
// grid may be scheduled as soon as every CTA of its predecessor has exited, instead of after the predecessor's
// completion has been processed, and the wait orders it behind the predecessor's memory flush.  Completion is
// transitive (a grid cannot complete before its own wait returns), hence every kernel launched through LM_LAUNCH_PDL
// MUST call lm_pdl_enter() first thing in every thread (tests/test_abi.py checks the sources).  The instruction is a
// no-op for a kernel launched without the attribute.  LMONO_PDL=0 launches everything with plain stream order.
//
// Measured on B200 (profiles/pdl_variants_r02.log, C-3 workload, one sequence alone / 8 sequences per GPU):
//   plain launches                                  243.4 us / 403 us per step
//   LM_PDL_MODE 2  wait only (the default)          240.5 us / 392 us
//   LM_PDL_MODE 1  wait, then launch_dependents     246.0 us / 508 us   the early-resident successor CTAs hold registers and
//                                                                        shared memory that the other sequences' running kernels need
//   LM_PDL_MODE 0  launch_dependents, then wait     WRONG RESULTS: a successor that became resident before its predecessor
//                                                   had written a buffer later read that buffer through stale L1 / read-only
//                                                   (LDG.NC, const __restrict__) lines -- griddepcontrol.wait orders the
//                                                   grids but does not invalidate them;
//   LM_PDL_MODE 3  as 0 + fence.acq_rel.gpu         correct, 259.9 us / 541 us
// i.e. the chain is bound by what happens inside its kernels (dependent L2 / HBM round trips after the L2 flush), not by
// the launch gaps between them; early release of the successor only takes SM resources away from the other sequences.
static inline bool lm_pdl_on() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("LMONO_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}
#ifdef __CUDACC__
#ifndef LM_PDL_MODE
#define LM_PDL_MODE 2
#endif
__device__ __forceinline__ void lm_pdl_enter() {
#if LM_PDL_MODE == 0
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
#elif LM_PDL_MODE == 1
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;");
#elif LM_PDL_MODE == 2
  asm volatile("griddepcontrol.wait;" ::: "memory");
#elif LM_PDL_MODE == 3
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
}
template <typename... P, typename... A>
static inline void lm_launch_pdl(cudaStream_t stream, void (*kern)(P...), dim3 grid, dim3 block, size_t smem, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = lm_pdl_on() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, static_cast<P>(args)...);     // a failure is picked up by LM_LAUNCH_CHECK (cudaGetLastError)
}
#define LM_LAUNCH_PDL(kern, grid, block, smem, ...) lm_launch_pdl(ctx->stream, kern, dim3(grid), dim3(block), (size_t)(smem), __VA_ARGS__)
#endif

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long d_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define LM_STAMP(buf, slot) do { if ((buf) != nullptr && threadIdx.x == 0 && blockIdx.x == 0) (buf)[(slot)] = d_globaltimer(); } while (0)
// a mod m in [0, m) for |a| < m * 2^22 (cube coordinates, window offsets): one unsigned remainder by a constant, no sign fix-up
__device__ __forceinline__ int d_pmod(int a, int m) { return (int)((unsigned)(a + m * (1 << 22)) % (unsigned)m); }
__device__ __forceinline__ int d_floordiv(int a, int b) { int q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) --q; return q; }
// floor(a / 25) for |a| < 2^25 (2 m cell coordinates): unsigned division by a constant
__device__ __forceinline__ int d_floordiv25(int a) { return (int)((unsigned)(a + 25 * (1 << 21)) / 25u) - (1 << 21); }

// cube coordinate of a world value: (int)((v + 25.0) / 50.0) + cen, minus one if v + 25.0 < 0
// (Aloam/src/laserMapping.cpp:312-321, 741-750)
__device__ __forceinline__ int d_cube_coord(double v, int cen) {
  int c = (int)((v + 25.0) / 50.0) + cen;
  if (v + 25.0 < 0) c--;
  return c;
}
// the slot tables share one allocation: [0, LM_NSLOT) window rank of a physical slot (-1 = outside the window), then
// int2 {slab id or -1, concatenation offset} per map type (k_begin_step)
constexpr int LM_SLOT_TABLE_INTS = (LM_NSLOT + 1) + 4 * LM_NSLOT;
__host__ __device__ __forceinline__ int2* lm_slot_info(int32_t* slot_valid_rank) { return reinterpret_cast<int2*>(slot_valid_rank + LM_NSLOT + 1); }
__host__ __device__ __forceinline__ const int2* lm_slot_info(const int32_t* slot_valid_rank) { return reinterpret_cast<const int2*>(slot_valid_rank + LM_NSLOT + 1); }
__device__ __forceinline__ int d_phys_slot(int gi, int gj, int gk) {
  return d_pmod(gi, LM_GW) + LM_GW * d_pmod(gj, LM_GH) + LM_GW * LM_GH * d_pmod(gk, LM_GD);
}

// Eigen quaternion algebra, q = (x,y,z,w), double, evaluated in Eigen's operation order
__device__ __forceinline__ void d_qmul(const double* a, const double* b, double* o) {
  double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
  double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
__device__ __forceinline__ void d_qrot(const double* q, const double* v, double* o) {
  double uv0 = q[1] * v[2] - q[2] * v[1], uv1 = q[2] * v[0] - q[0] * v[2], uv2 = q[0] * v[1] - q[1] * v[0];
  uv0 += uv0; uv1 += uv1; uv2 += uv2;
  double c0 = q[1] * uv2 - q[2] * uv1, c1 = q[2] * uv0 - q[0] * uv2, c2 = q[0] * uv1 - q[1] * uv0;
  o[0] = (v[0] + q[3] * uv0) + c0; o[1] = (v[1] + q[3] * uv1) + c1; o[2] = (v[2] + q[3] * uv2) + c2;
}
__device__ __forceinline__ void d_qinv(const double* q, double* o) {
  double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (n2 > 0.0) { o[0] = -q[0] / n2; o[1] = -q[1] / n2; o[2] = -q[2] / n2; o[3] = q[3] / n2; }
  else { o[0] = o[1] = o[2] = o[3] = 0.0; }
}
// pointAssociateToMap (laserMapping.cpp:154-163): double rotate + translate, stored as float
__device__ __forceinline__ float4 d_associate(const double* q, const double* t, float4 pi) {
  double v[3] = { (double)pi.x, (double)pi.y, (double)pi.z }, o[3];
  d_qrot(q, v, o);
  float4 r;
  r.x = (float)(o[0] + t[0]); r.y = (float)(o[1] + t[1]); r.z = (float)(o[2] + t[2]); r.w = pi.w;
  return r;
}

// transformUpdate (laserMapping.cpp:148-152) + frame counter; one thread
__device__ __forceinline__ void d_transform_update(LmMapState* st) {
  double qi[4]; d_qinv(st->q_wodom_curr, qi);
  double qn[4]; d_qmul(st->q_w_curr, qi, qn);
  for (int k = 0; k < 4; ++k) st->q_wmap_wodom[k] = qn[k];
  double tmp[3]; d_qrot(qn, st->t_wodom_curr, tmp);
  for (int k = 0; k < 3; ++k) st->t_wmap_wodom[k] = st->t_w_curr[k] - tmp[k];
  st->frame_count++;
}

// voxel coordinate of PCL VoxelGrid: floor(x * inverse_leaf) in fp32
__device__ __forceinline__ int d_voxel_coord(float x, float inv_leaf) { return (int)floorf(__fmul_rn(x, inv_leaf)); }

// cube-local voxel key (z-major, then y, then x -- the order PCL's linear index induces).
// 10 bits per axis; origin = voxel of the cube's lower corner minus 2 (slack).
__device__ __forceinline__ uint32_t d_cube_voxel_key(float4 p, float inv_leaf, const int* g) {
  int ox = (int)floorf((float)(50 * g[0] - 25) * inv_leaf) - 2;
  int oy = (int)floorf((float)(50 * g[1] - 25) * inv_leaf) - 2;
  int oz = (int)floorf((float)(50 * g[2] - 25) * inv_leaf) - 2;
  int lx = d_voxel_coord(p.x, inv_leaf) - ox, ly = d_voxel_coord(p.y, inv_leaf) - oy, lz = d_voxel_coord(p.z, inv_leaf) - oz;
  lx = min(max(lx, 0), 1023); ly = min(max(ly, 0), 1023); lz = min(max(lz, 0), 1023);
  return ((uint32_t)lz << 20) | ((uint32_t)ly << 10) | (uint32_t)lx;
}

// cube-local 2 m cell index of a point; returns -1 if outside the table (never for points
// stored in the cube)
__device__ __forceinline__ int d_cube_cell(float4 p, const int* g) {
  int cx = ((int)floorf(p.x) >> 1) - (25 * g[0] - 13);
  int cy = ((int)floorf(p.y) >> 1) - (25 * g[1] - 13);
  int cz = ((int)floorf(p.z) >> 1) - (25 * g[2] - 13);
  if ((unsigned)cx >= (unsigned)LM_CELLS_AXIS || (unsigned)cy >= (unsigned)LM_CELLS_AXIS || (unsigned)cz >= (unsigned)LM_CELLS_AXIS) return -1;
  return cx + LM_CELLS_AXIS * (cy + LM_CELLS_AXIS * cz);
}

// ---- cube sharding: owner rank of an absolute cube coordinate, and the rule that decides
// whether this rank stores a map point: it owns the point's cube, or the point's VoxelGrid voxel
// reaches into the LM_SHARD_HALO shell around an adjacent cube it owns.  Accepted neighbours satisfy
// fp32 d2 < 1.0 (laserMapping.cpp:584,652), so a 1 m shell (taken voxel-complete, with slack) makes
// every owned query's 5-NN local and exact; voxel-complete halos make the per-cube VoxelGrid
// refilter of a halo copy reproduce the owner's centroids bit for bit.
#define LM_SHARD_HALO 1.25
// Cyclic: owner = (gi + 3 gj + 5 gk) mod n.  A 5x5x3 search window (and its ground layer, where most points are) then
// spreads almost evenly over the ranks -- 1.07x (1.28x) the mean on the busiest of 8 ranks; a hash gave 1.5x (1.9x) --
// and the busiest rank bounds every registration.  Not contiguous blocks: a window must not land on one GPU.
__host__ __device__ __forceinline__ int lm_cube_owner(int gi, int gj, int gk, int nranks) {
  if (nranks <= 1) return 0;
  const int v = (gi + 3 * gj + 5 * gk) % nranks;
  return v < 0 ? v + nranks : v;
}
__device__ __forceinline__ bool d_shard_keep(float4 p, float leaf, float inv_leaf, int rank, int nranks) {
  if (nranks <= 1) return true;
  const float pc[3] = { p.x, p.y, p.z };
  int g[3]; bool lo[3], hi[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    g[a] = d_cube_coord((double)pc[a], 0);
    const double v = (double)floorf(__fmul_rn(pc[a], inv_leaf));
    const double vlo = v * (double)leaf, vhi = (v + 1.0) * (double)leaf;
    lo[a] = vlo < (50.0 * g[a] - 25.0) + LM_SHARD_HALO;
    hi[a] = vhi > (50.0 * g[a] + 25.0) - LM_SHARD_HALO;
  }
  if (lm_cube_owner(g[0], g[1], g[2], nranks) == rank) return true;
  for (int dz = -1; dz <= 1; ++dz) {
    if ((dz < 0 && !lo[2]) || (dz > 0 && !hi[2])) continue;
    for (int dy = -1; dy <= 1; ++dy) {
      if ((dy < 0 && !lo[1]) || (dy > 0 && !hi[1])) continue;
      for (int dx = -1; dx <= 1; ++dx) {
        if ((dx < 0 && !lo[0]) || (dx > 0 && !hi[0])) continue;
        if ((dx | dy | dz) == 0) continue;
        if (lm_cube_owner(g[0] + dx, g[1] + dy, g[2] + dz, nranks) == rank) return true;
      }
    }
  }
  return false;
}

// ---- peer-memory all-gather + fixed-order sum of `nvals` doubles (cube-sharded map).  Called by ALL threads of a CTA
// (blockDim.x >= max(nvals, n)); `in` / `out` are shared-memory arrays (may alias), `publisher` is CTA-uniform: exactly
// one CTA per rank and exchange publishes, any number of CTAs may consume.
__device__ __forceinline__ void d_st_release_sys_u64(unsigned long long* p, unsigned long long v) { asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long d_ld_acquire_sys_u64(const unsigned long long* p) { unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void d_st_relaxed_sys_f64(double* p, double v) { asm volatile("st.relaxed.sys.global.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ double d_ld_relaxed_sys_f64(const double* p) { double v; asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void d_shard_exchange(const LmShardPeers* __restrict__ peers, unsigned long long epoch, const double* in, double* out,
                                                 int nvals, bool publisher, uint32_t* fault) {
  const int rank = peers->rank, n = peers->n;
  const int par = (int)(epoch & 1ull);
  LmShardXchg* mine = peers->peer[rank];
  const unsigned long long t_enter = d_globaltimer();
  if (publisher) {
    if ((int)threadIdx.x < nvals) {
      const double v = in[threadIdx.x];
      for (int r = 0; r < n; ++r) d_st_relaxed_sys_f64(&peers->peer[r]->payload[par][rank][threadIdx.x], v);
      __threadfence_system();
    }
    __syncthreads();
    if ((int)threadIdx.x < n) d_st_release_sys_u64(&peers->peer[threadIdx.x]->flag[par][rank], epoch);
  }
  if ((int)threadIdx.x < n) {
    const unsigned long long t0 = d_globaltimer();
    while (d_ld_acquire_sys_u64(&mine->flag[par][threadIdx.x]) < epoch) {
      if (d_globaltimer() - t0 > LM_XCHG_TIMEOUT_NS) { atomicOr(fault, LM_FAULT_SHARD_TIMEOUT); break; }
    }
  }
  __syncthreads();
  if (publisher && threadIdx.x == 0) { mine->wait_ns += d_globaltimer() - t_enter; mine->n_xchg += 1; }   // statistics: post + wait for the slowest rank
  if ((int)threadIdx.x < nvals) {
    double s = 0.0;
    for (int r = 0; r < n; ++r) s += d_ld_relaxed_sys_f64(&mine->payload[par][r][threadIdx.x]);
    out[threadIdx.x] = s;
  }
  __syncthreads();
}

// block-wide exclusive scan of one int per thread (blockDim.x <= 1024, multiple of 32).
// warp_sums: shared int[33]. Returns the exclusive prefix; *total = block sum.
__device__ __forceinline__ int d_block_exscan(int v, int* warp_sums, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int s = lane < nw ? warp_sums[lane] : 0;
    int si = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, si, o); if (lane >= o) si += t; }
    if (lane < nw) warp_sums[lane] = si - s;      // exclusive warp offsets
    if (lane == 31) warp_sums[32] = si;            // total
  }
  __syncthreads();
  int res = warp_sums[wid] + incl - v;
  *total = warp_sums[32];
  __syncthreads();
  return res;
}

// in-place ascending bitonic sort of n_pow2 64-bit keys in shared memory by the whole block
__device__ __forceinline__ void d_bitonic_sort(unsigned long long* s, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < (n_pow2 >> 1); i += blockDim.x) {
        int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));   // index with bit j cleared
        int hi = lo | j;
        bool up = ((lo & k) == 0);
        unsigned long long a = s[lo], b = s[hi];
        if ((a > b) == up) { s[lo] = b; s[hi] = a; }
      }
      __syncthreads();
    }
  }
}

// ---- register-resident bitonic sort ---------------------------------------------------
// Sorts N = ITEMS * nthreads 64-bit keys ascending; thread t of the group holds elements
// [t*ITEMS, (t+1)*ITEMS).  Stages whose partner distance j is < ITEMS run in registers,
// ITEMS <= j < 32*ITEMS through warp shuffles, larger j through shared memory (2 barriers
// per stage, only log2(N/(32*ITEMS)) * (..+1)/2 of them).  `nthreads` threads with contiguous
// ids starting at `tid0 = 0` must call it together; if nthreads > 32, `xch` must hold N keys
// and the whole block must participate (uses __syncthreads).
__device__ __forceinline__ void d_cmpx(unsigned long long& a, unsigned long long& b, bool up) {
  // a at the lower index, b at the higher
  if ((a > b) == up) { unsigned long long t = a; a = b; b = t; }
}

template <int ITEMS, int NTHREADS>
__device__ __forceinline__ void d_bitonic_regs(unsigned long long (&v)[ITEMS], int t, unsigned long long* xch) {
  // every loop has a compile-time trip count and is fully unrolled: v[] must be indexed with constants
  // only, otherwise it is demoted to local memory and the network runs ~5x slower
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
        // cross-warp: exchange through shared memory
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) xch[t * ITEMS + r] = v[r];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
          const int i = t * ITEMS + r;
          const unsigned long long o = xch[i ^ j];
          const bool up = (i & k) == 0;
          const bool lower = (i & j) == 0;
          const unsigned long long mn = v[r] < o ? v[r] : o, mx = v[r] < o ? o : v[r];
          v[r] = (lower == up) ? mn : mx;
        }
        __syncthreads();
      } else if (j >= ITEMS) {
        const int lane_x = j / ITEMS;
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
          const int i = t * ITEMS + r;
          const unsigned long long o = __shfl_xor_sync(0xffffffffu, v[r], lane_x);
          const bool up = (i & k) == 0;
          const bool lower = (i & j) == 0;
          const unsigned long long mn = v[r] < o ? v[r] : o, mx = v[r] < o ? o : v[r];
          v[r] = (lower == up) ? mn : mx;
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

// first index in sorted[0..n) with sorted[i] >= key
__device__ __forceinline__ int d_lower_bound_u64(const unsigned long long* sorted, int n, unsigned long long key) {
  int lo = 0, hi = n;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (sorted[mid] < key) lo = mid + 1; else hi = mid; }
  return lo;
}
__device__ __forceinline__ int d_lower_bound_u32(const uint32_t* sorted, int n, uint32_t key) {
  int lo = 0, hi = n;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (sorted[mid] < key) lo = mid + 1; else hi = mid; }
  return lo;
}
// the same result with eight independent probes per step (9-ary search): ~log9(n) dependent memory round trips instead
// of log2(n) -- for searches whose cost is the latency of the chain, not the number of loads
__device__ __forceinline__ int d_lower_bound_u32_wide(const uint32_t* __restrict__ sorted, int n, uint32_t key) {
  int lo = 0, hi = n;                       // the answer lies in [lo, hi]
  while (hi - lo > 8) {
    const int step = (hi - lo) / 9;         // >= 1
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = sorted[lo + (k + 1) * step];
    int c = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) c += v[k] < key ? 1 : 0;
    const int nlo = c > 0 ? lo + c * step + 1 : lo;
    const int nhi = c < 8 ? lo + (c + 1) * step : hi;
    lo = nlo; hi = nhi;
  }
  uint32_t v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = lo + k < hi ? sorted[lo + k] : 0xFFFFFFFFu;
  int c = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) c += (lo + k < hi && v[k] < key) ? 1 : 0;
  return lo + c;
}
#endif  // __CUDACC__

// ---------------------------------------------------------------- cross-TU host functions
// sort.cu: segment s sorts *n[s] keys in[off[s] ...] -> out[off[s] ...] (tmp[off[s] ...] is scratch)
struct LmSortSegs { const unsigned long long* in; unsigned long long* tmp; unsigned long long* out; int off[LM_SORT_MAXSEG]; const int32_t* n[LM_SORT_MAXSEG]; };
int lm_sort_u64_segs(lmono_ctx* ctx, const LmSortSegs& sg, int nseg, const int* n_max);
int lm_sort_configure(lmono_ctx* ctx);
// sorts n (device int *n_dev, <= n_max) unique 64-bit keys ascending: in -> out (tmp is scratch)
int lm_sort_u64(lmono_ctx* ctx, const unsigned long long* in, unsigned long long* tmp, unsigned long long* out,
                const int32_t* n_dev, int n_max);
// voxel.cu: VoxelGrid of `in` (n_dev points, <= n_max) -> out, *out_n_dev; _multi runs up to LM_SORT_MAXSEG clouds through shared launches
int lm_voxel_init(lmono_ctx* ctx);
int lm_voxel_grid_multi(lmono_ctx* ctx, int nseg, const float4* const* in, const int32_t* const* n_dev, const int* n_max,
                        const float* leaf, float4* const* out, int32_t* const* out_n_dev,
                        const float4* const* const* in_ind = nullptr,    // in_ind[k] != NULL: read the input pointer from device memory
                        bool fetch = false);   // honour LmMapState::in_src / in_stride (fused upload from page-locked host memory)
int lm_voxel_grid_device(lmono_ctx* ctx, const float4* in, const int32_t* n_dev, int n_max, float leaf,
                         float4* out, int32_t* out_n_dev);
// mapstore.cu
int lm_map_alloc(lmono_ctx* ctx);
void lm_map_free(lmono_ctx* ctx);
// pose source: wodom_curr (by value), else t_override (test hooks), else the q/t_wodom_curr already in the state
int lm_map_begin_step(lmono_ctx* ctx, const lmono_pose* wodom_curr, const double* t_override);
int lm_map_index_build(lmono_ctx* ctx);
int lm_map_insert_and_refilter(lmono_ctx* ctx, int n_max_corner, int n_max_surf, bool transform_update = false);
// assoc.cu
int lm_map_associate(lmono_ctx* ctx, int n_max_corner, int n_max_surf);
int lm_knn5_device(lmono_ctx* ctx, int which, const float4* d_q, int n, int32_t* d_idx, float* d_d2);
// lm.cu
int lm_solve_enqueue(lmono_ctx* ctx, int solve_index, int n_max_corner, int n_max_surf, int max_iter);
int lm_normal_eq_enqueue(lmono_ctx* ctx, int n_max_corner, int n_max_surf);
int lm_solve_problem(lmono_ctx* ctx, const LmProblem& P, int n_max, int max_iter, int write_back, bool shard_exchange = false);
// shard.cu: peer-memory gate exchange (owned window counts -> the global :554 gate), one launch
int lm_shard_gate_xchg(lmono_ctx* ctx);
void lm_shard_free(lmono_ctx* ctx);
// sharded LM pieces: begin / partial evaluation into the workspace / controller from the reduced workspace
int lm_shard_lm_begin(lmono_ctx* ctx, int solve_index);
int lm_shard_lm_eval(lmono_ctx* ctx, int solve_index);
int lm_shard_lm_control(lmono_ctx* ctx, int solve_index);
// ctx.cu helpers
int lm_upload_cloud(lmono_ctx* ctx, lmono_cloud_view v, uint8_t* d_raw, float4* d_out, int32_t* d_n /*may be null*/);
int lm_download_cloud(lmono_ctx* ctx, const float4* d_src, int n, lmono_cloud_out* out);
