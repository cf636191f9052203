/*
 * lmono.h -- C ABI of liblmono_b200.so: the B200-native (sm_100a CUDA) implementation of
 * LMONO-Fusion's LiDAR registration hot path (A-LOAM scanRegistration -> laserOdometry ->
 * laserMapping, plus the mono_lidar_mapping LiDAR->camera colour projection).
 *
 * The reference (bobocode/lmono) exposes NO plugin/FFI interface for this path: the hot
 * loops are inlined in the ROS node mains.  Each entry point below
 * therefore replaces a block of a reference node's callback, cited as file:line relative
 * to the reference tree; nodes/ holds the patched node sources that call them and
 * INTEGRATION.md shows the binding.  Plain C: opaque handle, POD structs, caller-owned
 * host buffers, no exceptions, no torch types.
 *
 * Conventions: return 0 = OK, negative = error (lmono_strerror).  "Not enough map points"
 * and "<10 correspondences" are statuses in the report, not errors (they mirror
 * Aloam/src/laserMapping.cpp:730-733 and Aloam/src/laserOdometry.cpp:488-491).  One ctx is
 * single-caller (one process() thread per node, laserMapping.cpp:934); several ctxs (one
 * per sequence / GPU) are independent.  There is no CPU fallback: every entry point fails
 * with LMONO_E_CUDA if the device is unusable.
 */
#ifndef LMONO_H
#define LMONO_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LMONO_ABI_VERSION 1

enum {
  LMONO_OK = 0,
  LMONO_E_ARG = -1,          /* bad argument */
  LMONO_E_CAPACITY = -2,     /* an output buffer or internal capacity is too small */
  LMONO_E_CUDA = -3,         /* CUDA runtime error (latched in the ctx) */
  LMONO_E_STATE = -4,        /* call not valid in the current state */
  LMONO_E_DEVICE = -5        /* device-side fault flag raised by a kernel (see lmono_last_fault) */
};

typedef struct lmono_ctx lmono_ctx;

/* Host point cloud, array of structs.  x,y,z are floats at byte offsets 0,4,8 of each
 * record; intensity is a float at intensity_offset (or absent if < 0).  pcl::PointXYZI is
 * {stride 32, intensity_offset 16} (Aloam/include/aloam_velodyne/common.h:43);
 * pcl::PointXYZ is {16, -1}; KITTI .bin (Aloam/src/kittiHelper.cpp:25-35) is {16, 12}. */
typedef struct {
  const void* base;
  int32_t n;
  int32_t stride_bytes;
  int32_t intensity_offset;
} lmono_cloud_view;

/* Caller-allocated output cloud in the same layout; callee fills n_out.  If capacity is
 * too small the call returns LMONO_E_CAPACITY and n_out holds the required size. */
typedef struct {
  void* base;
  int32_t capacity;
  int32_t stride_bytes;
  int32_t intensity_offset;
  int32_t n_out;
} lmono_cloud_out;

typedef struct { double q[4]; /* x,y,z,w */ double t[3]; } lmono_pose;

/* Parameters = the reference's ROS params / compile-time constants, same names. */
typedef struct {
  int32_t scan_line;                 /* scanRegistration.cpp:466  (16/32/64) */
  float   minimum_range;             /* scanRegistration.cpp:468 */
  float   mapping_line_resolution;   /* laserMapping.cpp:902 */
  float   mapping_plane_resolution;  /* laserMapping.cpp:903 */
  int32_t mapping_skip_frame;        /* laserOdometry.cpp:191: cadence at which the ODOMETRY NODE forwards sweeps to laserMapping
                                      * (nodes/laserOdometry_b200.cpp applies it); lmono_sweep_step* map every sweep they are given */
  /* capacities (0 = default) */
  int32_t max_sweep_points;          /* raw points per sweep            (default 262144) */
  int32_t max_feature_points;        /* points per feature cloud        (default 131072) */
  int32_t cube_capacity_corner;      /* points per 50 m cube, corner map (default 32768) */
  int32_t cube_capacity_surf;        /* points per 50 m cube, surf map   (default 49152) */
  int32_t max_cubes_corner;          /* non-empty cubes held at once     (default 768) */
  int32_t max_cubes_surf;            /*                                  (default 768) */
  int32_t image_width, image_height; /* colour projection raster (default 1241 x 376) */
  int32_t distortion;                /* #define DISTORTION of laserOdometry.cpp:59 (default 0, as compiled in the reference): 1 = per-point
                                        interpolation ratio s = (intensity - int(intensity)) / 0.1 in TransformToStart and the factors */
  int32_t stages;                    /* what this ctx will be asked to run: 0 = everything (default); else a mask of LMONO_STAGE_*.  Only a ctx
                                        with LMONO_STAGE_MAPPING allocates the cube map's slab pools (GBs); its map / sweep / shard calls
                                        return LMONO_E_STATE otherwise.  A scanRegistration- or laserOdometry-only node passes its own bit. */
  int32_t reserved[6];
} lmono_params;
#define LMONO_STAGE_SCAN     1
#define LMONO_STAGE_ODOMETRY 2
#define LMONO_STAGE_MAPPING  4
#define LMONO_STAGE_COLOUR   8

void lmono_default_params(lmono_params* p);

typedef struct {
  int32_t iterations, num_successful, termination, num_factors;
  double initial_cost, final_cost;
} lmono_solve_summary;     /* termination: 0 max-iter 1 gradient 2 parameter 3 function 4 radius 5 failure 6 none */

/* laserMapping report: the numbers the reference prints (laserMapping.cpp:552-553,707-708). */
typedef struct {
  int32_t corner_from_map, surf_from_map;   /* :538-539 */
  int32_t corner_stack, surf_stack;         /* :545,550 */
  int32_t corner_num[2], surf_num[2];       /* :620,685 per outer iteration */
  int32_t optimized;                        /* :554 gate (0 => "Map corner and surf num are not enough") */
  int32_t center_cube[3];
  int32_t cen[3];                           /* laserCloudCenWidth/Height/Depth after the shift */
  lmono_solve_summary solve[2];
  float ms_gpu;                             /* device time of the step (CUDA events) */
} lmono_map_report;

typedef struct {
  int32_t inited;                           /* 0 on the first frame (laserOdometry.cpp:267-271) */
  int32_t corner_corr[2], plane_corr[2];    /* :382,480 */
  lmono_solve_summary solve[2];
  float ms_gpu;
} lmono_odom_report;

typedef struct {
  int32_t n_in, n_kept;
  int32_t n_sharp, n_less_sharp, n_flat, n_less_flat;
  int32_t ring_start[64], ring_end[64];     /* scanStartInd / scanEndInd (scanRegistration.cpp:249-251) */
  float start_ori, end_ori;
  float ms_gpu;
} lmono_scan_report;

typedef struct {
  double fx, fy, cx, cy, k1, k2, p1, p2;    /* camera_models PinholeCamera parameters */
  int32_t width, height;
  int32_t kernel_type;                       /* 0 FULL 1 CROSS 2 ELLIPSE (map_build_node.cc:278) */
  int32_t kernel_size;
  int32_t blur_type;                         /* 0 bilateral 1 gaussian */
} lmono_pinhole;

/* ------------------------------------------------------------------ lifecycle */
/* stream: a cudaStream_t (as void*) the ctx should enqueue on, or NULL to create its own. */
int  lmono_create(int device, const lmono_params* params, void* stream, lmono_ctx** out);
void lmono_destroy(lmono_ctx* ctx);
const char* lmono_strerror(int code);
/* device-side fault bits raised since the last call (0 = none); clears them. */
#define LMONO_FAULT_CUBE_OVERFLOW    (1u << 0)  /* a cube's slab is full: new points of that cube were dropped */
#define LMONO_FAULT_POOL_EXHAUSTED   (1u << 1)  /* no free slab for a newly touched cube (lmono_map_evict frees far ones) */
#define LMONO_FAULT_TAIL_OVERFLOW    (1u << 2)  /* a cube to re-voxelise as a whole exceeds 65 536 points */
#define LMONO_FAULT_CELL_RANGE       (1u << 3)  /* internal: a point outside its cube's search table */
#define LMONO_FAULT_FEATURE_OVERFLOW (1u << 4)  /* a feature cloud exceeded max_feature_points */
#define LMONO_FAULT_IMPORT_NONEMPTY  (1u << 5)  /* lmono_map_import into a cube that already holds points */
#define LMONO_FAULT_SHARD_TIMEOUT    (1u << 6)  /* cube-sharded map, peer-memory mode: a rank did not answer */
int  lmono_last_fault(lmono_ctx* ctx, uint32_t* bits);
int  lmono_sync(lmono_ctx* ctx);
/* Device time (CUDA events on the ctx stream) of the most recent lmono_scan_register / lmono_odom_step /
 * lmono_project_color call: uploads + kernels, and kernels only (inputs resident in HBM). */
int  lmono_stage_times(lmono_ctx* ctx, float* ms_with_uploads /*may be NULL*/, float* ms_kernels /*may be NULL*/);
/* number of kernels this ctx has launched so far (bench.py reports it as gpu_launches). */
int64_t lmono_launch_count(const lmono_ctx* ctx);

/* ------------------------------------------------------------------ L3: laserMapping */
/* Replaces Aloam/src/laserMapping.cpp:307-801 (+ :838-842 when full_res is given):
 * transformAssociateToMap, cube-window shift, local-map gather, VoxelGrid of the incoming
 * features, 2 x (5-NN + line/plane fit + LM solve), transformUpdate, map insertion and
 * per-cube VoxelGrid refilter.  corner_last / surf_last are the odometry node's
 * /laser_cloud_corner_last and /laser_cloud_surf_last clouds (sensor frame). */
int lmono_map_step(lmono_ctx* ctx, lmono_cloud_view corner_last, lmono_cloud_view surf_last,
                   const lmono_pose* wodom_curr, lmono_pose* w_curr /*out*/,
                   lmono_pose* wmap_wodom /*out, may be NULL*/, lmono_map_report* report /*may be NULL*/,
                   lmono_cloud_view full_res /*n=0 to skip*/, lmono_cloud_out* registered /*may be NULL*/);

/* Device-resident variant: inputs are packed XYZI float4 arrays already in HBM; the call
 * only enqueues work on the ctx stream.  lmono_map_collect() synchronises and returns the
 * results of the most recent step. */
int lmono_map_step_device(lmono_ctx* ctx, const void* d_corner_xyzi, int32_t n_corner,
                          const void* d_surf_xyzi, int32_t n_surf, const lmono_pose* wodom_curr);
int lmono_map_collect(lmono_ctx* ctx, lmono_pose* w_curr, lmono_pose* wmap_wodom, lmono_map_report* report);

/* Sequence batches (BASELINE config C-4: independent sequences per GPU).  One ctx per sequence; a batch call
 * drives n of them from one host thread so that their registrations overlap on the device (every kernel of a
 * step is small next to 148 SMs).  The reference runs one laserMapping process per sequence
 * (Aloam/src/laserMapping.cpp:931-934); this is the same independence with one process per GPU.
 * The steps of all n sequences run as parallel branches of ONE CUDA graph (cached in ctxs[0], max 64 sequences):
 * a batch step costs the host one small argument launch + one graph launch.  The ctxs may share one stream
 * (cheapest: nothing else to order) or own their streams (the batch is then ordered after the work already
 * enqueued on each, and later work on each is ordered after the batch, with events).
 * Host clouds in page-locked memory (cudaHostAlloc / cudaHostRegister) are never copied by the host: the first
 * kernel of the step reads them over PCIe while it computes the voxel keys (fused upload, any record stride).  Such
 * buffers must stay valid and unmodified until the step has been waited for.  Pageable clouds are staged with
 * cudaMemcpyAsync as before.  LMONO_NO_ZEROCOPY=1 forces staging.
 *   lmono_map_step_async        lmono_map_step without the final wait: enqueues the registration;
 *                               lmono_map_collect() waits for this ctx and returns the result.
 *   lmono_map_step_batch        = lmono_map_submit_batch + lmono_map_wait_batch.  Arrays have n entries;
 *                               wmap_wodom_in (may be NULL) replaces q/t_wmap_wodom of ctx i before its step.
 *   lmono_map_submit_batch      enqueue-only: the batch graph on the stream of ctxs[0]; every branch ends by storing
 *                               the sequence's state into one of the ctx's two page-locked result mirrors.  Up to TWO
 *                               submissions per ctx may be outstanding (LMONO_E_ARG otherwise), so the host can
 *                               prepare and submit sweep k+1 while sweep k runs (the mapping input of sweep k+1 does
 *                               not depend on the mapping result of sweep k: laserMapping.cpp:235-305 only pairs
 *                               odometry outputs).
 *   lmono_map_wait_batch        waits for the OLDEST outstanding submission of every listed ctx and returns its
 *                               results (same outputs as lmono_map_step_batch).
 *   lmono_map_step_device_batch device-resident inputs, enqueue-only, on join_stream (NULL: the stream of ctxs[0]):
 *                               ordered after the work already enqueued there, and work enqueued there afterwards is
 *                               ordered after the whole batch (no host synchronisation).
 *                               lmono_map_collect() returns the state of a sequence. */
int lmono_map_step_async(lmono_ctx* ctx, lmono_cloud_view corner_last, lmono_cloud_view surf_last, const lmono_pose* wodom_curr);
int lmono_map_step_batch(lmono_ctx* const* ctxs, int32_t n, const lmono_cloud_view* corner_last, const lmono_cloud_view* surf_last,
                         const lmono_pose* wodom_curr, const lmono_pose* wmap_wodom_in /*may be NULL*/,
                         lmono_pose* w_curr /*out, may be NULL*/, lmono_pose* wmap_wodom /*out, may be NULL*/,
                         lmono_map_report* reports /*out, may be NULL*/);
int lmono_map_submit_batch(lmono_ctx* const* ctxs, int32_t n, const lmono_cloud_view* corner_last, const lmono_cloud_view* surf_last,
                           const lmono_pose* wodom_curr, const lmono_pose* wmap_wodom_in /*may be NULL*/);
int lmono_map_wait_batch(lmono_ctx* const* ctxs, int32_t n, lmono_pose* w_curr /*out, may be NULL*/,
                         lmono_pose* wmap_wodom /*out, may be NULL*/, lmono_map_report* reports /*out, may be NULL*/);
int lmono_map_step_device_batch(lmono_ctx* const* ctxs, int32_t n, const void* const* d_corner_xyzi, const int32_t* n_corner,
                                const void* const* d_surf_xyzi, const int32_t* n_surf, const lmono_pose* wodom_curr,
                                const lmono_pose* wmap_wodom_in /*may be NULL*/, void* join_stream /*may be NULL*/);

/* q_wmap_wodom / t_wmap_wodom (laserMapping.cpp:116-117) */
int lmono_map_get_state(lmono_ctx* ctx, lmono_pose* wmap_wodom, int32_t cen[3]);
int lmono_map_set_state(lmono_ctx* ctx, const lmono_pose* wmap_wodom);   /* enqueue-only */
/* bytes lmono_map_collect / lmono_map_step read back from the device per step (pose + report) */
int32_t lmono_map_result_bytes(void);

/* Map exchange (the publishers at laserMapping.cpp:806-836 and checkpoint/restore).
 * which: 0 corner, 1 surf, 2 both, interleaved cube by cube (corner cube, then surf cube) exactly as the reference's
 * /laser_cloud_surround and /laser_cloud_map messages are assembled (:808-816, :826-830).  scope: 0 = the <=75 cubes
 * of the current window in the order of :512-537, 1 = all 4851 cubes in the order of :826-830. */
int lmono_map_export(lmono_ctx* ctx, int which, int scope, lmono_cloud_out* out);
/* Load world-frame points into an EMPTY map: every point goes to its cube
 * (laserMapping.cpp:741-758 arithmetic) and every cube is VoxelGrid-filtered (:788-801). */
int lmono_map_import(lmono_ctx* ctx, int which, lmono_cloud_view pts_world);
int lmono_map_clear(lmono_ctx* ctx);
/* Capacity valve for long drives: the reference's cube clouds (laserMapping.cpp:104) grow without bound, the slab pools do
 * not.  Frees the slabs of every cube more than keep_cubes (>= 3) cubes away from the window's centre cube in any axis --
 * cubes the 5x5x3 window cannot reach without the sensor travelling (keep_cubes - 2) x 50 m first; their points leave the
 * map.  Call it between steps, e.g. when a step returned LMONO_E_DEVICE with LMONO_FAULT_POOL_EXHAUSTED in
 * lmono_last_fault (nodes/laserMapping_b200.cpp does).  n_freed (may be NULL): slabs returned to the pools. */
int lmono_map_evict(lmono_ctx* ctx, int32_t keep_cubes, int32_t* n_freed);

/* ------------------------------------------------------------------ cube-sharded global map (multi-GPU)
 * Extension beyond the reference (its map is one process's 21x21x11 cube array, laserMapping.cpp:74-104):
 * the cubes are distributed over `nranks` contexts (one per GPU) by a hash of the absolute cube
 * coordinate, each stored with a 1 m halo (accepted neighbours have d2 < 1.0, laserMapping.cpp:584,652,
 * so every query's 5-NN is local to the rank owning the query's cube).  One registration =
 *   lmono_shard_begin -> [all-reduce ws] -> lmono_shard_gate ->
 *   2 x { lmono_shard_associate, lmono_shard_lm_begin, 5 x { lmono_shard_lm_eval -> [all-reduce ws] ->
 *         lmono_shard_lm_control } } -> lmono_shard_end -> lmono_map_collect
 * where [all-reduce ws] is a SUM all-reduce of the 35 doubles of d_workspace over the ranks, issued by
 * the host on the ctx stream (torch.distributed / ncclAllReduce; lmono_b200/shard.py is the host side).
 * d_workspace: caller-owned device memory, >= 35 doubles: [0..20] J^T J upper triangle, [21..26] J^T r,
 * [27] cost, [28..29] corner / surf factor counts, [30..31] owned map points in the window, pad.
 * lmono_map_import and the insertion at the end of every registration keep only the points this
 * rank owns or holds as halo.  All calls are enqueue-only. */
int lmono_shard_configure(lmono_ctx* ctx, int32_t rank, int32_t nranks, void* d_workspace);
int lmono_shard_begin(lmono_ctx* ctx, const void* d_corner_xyzi, int32_t n_corner, const void* d_surf_xyzi, int32_t n_surf,
                      const lmono_pose* wodom_curr);
int lmono_shard_gate(lmono_ctx* ctx);
int lmono_shard_associate(lmono_ctx* ctx);
int lmono_shard_lm_begin(lmono_ctx* ctx, int32_t solve_index);
int lmono_shard_lm_eval(lmono_ctx* ctx, int32_t solve_index);
int lmono_shard_lm_control(lmono_ctx* ctx, int32_t solve_index);
int lmono_shard_end(lmono_ctx* ctx);
/* Peer-memory mode of the sharded map: the all-reduce is done by the kernels themselves over NVLink peer memory, so a
 * sharded registration is ONE enqueue (lmono_map_step_device / lmono_map_step, replayed as one CUDA graph) with no host
 * call, NCCL launch or synchronisation inside it.  Setup, once per rank (one rank = one GPU = one ctx, <= 16 ranks):
 *   lmono_shard_xchg_create  allocates this rank's exchange block; returns its cudaIpcMemHandle_t (64 bytes) for ranks
 *                            in other processes and its device pointer for ranks that are contexts of this process;
 *   [the host gathers the handles of all ranks over any transport: torch.distributed, MPI, a file]
 *   lmono_shard_xchg_open    maps the peers' blocks (ipc_handles: [nranks][64]; same_process_ptrs[r] != NULL replaces
 *                            handle r) and switches the ctx to this mode; also sets rank / nranks like
 *                            lmono_shard_configure.  All ranks must then issue the same sequence of registrations.
 * A rank whose peer does not answer within 2 s raises fault bit 6 (LMONO_E_DEVICE from lmono_map_collect) instead of
 * hanging.  lmono_shard_xchg_stats: {exchanges completed, ns spent posting + waiting for the slowest rank, exchanges
 * counted in the ns figure}. */
int lmono_shard_xchg_create(lmono_ctx* ctx, void* ipc_handle_out /*[64], may be NULL*/, void** local_ptr_out /*may be NULL*/);
int lmono_shard_xchg_open(lmono_ctx* ctx, int32_t rank, int32_t nranks, const void* ipc_handles /*[nranks][64], may be NULL*/,
                          void* const* same_process_ptrs /*[nranks], may be NULL*/);
int lmono_shard_xchg_stats(lmono_ctx* ctx, uint64_t out[3], int32_t reset);
/* owner rank of the cube containing a world point / of an absolute cube coordinate (cube 0 is centred on the origin) */
int32_t lmono_shard_owner_of_cube(int32_t gi, int32_t gj, int32_t gk, int32_t nranks);

/* ------------------------------------------------------------------ test / bench hooks */
/* Positions the cube window for a pose translation (laserMapping.cpp:312-539). */
int lmono_map_prepare_window(lmono_ctx* ctx, const double t_w_curr[3]);
/* 5-NN of world-frame queries against the current window (replaces the
 * kdtree*FromMap->nearestKSearch calls at laserMapping.cpp:582,648).  idx are positions in
 * the :533-537 concatenation; a query whose 5th neighbour is not within d2 < 1.0 gets
 * idx = -1 / d2 = +inf in the slots that could not be proven. */
int lmono_knn5(lmono_ctx* ctx, int which, lmono_cloud_view queries_world, int32_t* idx, float* d2);
int lmono_knn5_device(lmono_ctx* ctx, int which, const void* d_queries_xyzi, int32_t n, void* d_idx, void* d_d2);
/* One association pass (laserMapping.cpp:577-687) at pose w_curr followed by one
 * evaluation of the 6x6 normal equations (H = J^T J, g = J^T r, cost) of the resulting
 * factors, Huber(0.1) corrected, in the tangent space (rotation, translation). */
int lmono_map_normal_eq(lmono_ctx* ctx, lmono_cloud_view corner_stack, lmono_cloud_view surf_stack,
                        const lmono_pose* w_curr, double H[36], double g[6], double* cost,
                        int32_t* n_corner, int32_t* n_surf);
/* pcl::VoxelGrid<PointXYZI> as configured by the reference (canonical index-order sums). */
int lmono_voxel_grid(lmono_ctx* ctx, lmono_cloud_view in, float leaf, lmono_cloud_out* out);

/* Per-launch device timing: with marks enabled every kernel launch of lmono_map_step* is followed by a CUDA
 * event (the step then runs as plain launches, not as a graph replay).  lmono_kmarks_dump writes one line
 * "source.cu:line count total_ms" per launch site; bench.py --kernels maps the sites to kernel names. */
int lmono_kmarks_enable(lmono_ctx* ctx, int on);
/* Kernel forms.  A step that runs alone on the GPU uses the latency forms of the kernels (8 lanes per kNN query, 16-CTA
 * LM clusters); steps of >= 4 sequences enqueued together (lmono_map_*_batch) use the throughput forms (one thread per
 * kNN query, 8-CTA clusters above 8 sequences).  Both give the same bits.  A caller that overlaps sequences itself
 * (one host thread / stream per ctx) states the concurrency here. */
int lmono_set_concurrency_hint(lmono_ctx* ctx, int32_t n_sequences_on_this_gpu);
/* With LMONO_TIMELINE=1 in the environment every launch of a mapping step is followed by a one-thread %globaltimer
 * stamp kernel (also inside the step / batch graphs); lmono_timeline_dump writes "<file>:<line> <ns>" per stamp of the
 * last step, so the kernels of concurrent sequences can be laid on one time axis (profiles/batch_timeline.py). */
int lmono_timeline_dump(lmono_ctx* ctx, char* buf, int32_t cap);
int lmono_kmarks_dump(lmono_ctx* ctx, char* buf, int32_t cap);

/* Latency study hook: %globaltimer stamps (ns) written by instrumented kernels (slot map in DESIGN.md):
 * [0..63] the LM solve kernel of the most recent solve: 0 start, 1 armed, then per evaluation e (8 slots from
 * 8 + 8 e): factors evaluated, warp+block reduced, cluster exchanged, controller done;
 * [200..207] k_scan_ring of ring 32: start, ring staged, sectors sorted, greedy pick done, less-flat list, voxel bounds,
 * voxel keys sorted, centroids written. */
int lmono_debug_stamps(lmono_ctx* ctx, uint64_t* out /*[n]*/, int32_t n /*<= 4096*/);
/* Study hook: the per-cube refilter records of the last mapping step, [2 map types][75 window cubes] x 8 ints
 * {has a tail, size after the merge, filtered prefix before it, tail points, re-voxelise flag, buffer, slab, -}. */
int lmono_debug_rf_meta(lmono_ctx* ctx, int32_t* out, int32_t n_ints /*<= 1200*/);

/* Per-phase device timing (CUDA events on the ctx stream) of lmono_map_step, used by bench.py
 * for the roofline numbers.  Phases: 0 window shift, 1 cell-index build, 2 VoxelGrid of the
 * features, 3 association (5-NN + fits), 4 LM solve, 5 insertion, 6 cube refilter, 7 misc. */
#define LMONO_PROFILE_PHASES 8
int lmono_profile_enable(lmono_ctx* ctx, int on);
int lmono_profile_read(lmono_ctx* ctx, float* ms /*[8]*/, int32_t* counts /*[8]*/);

/* ------------------------------------------------------------------ L1: scanRegistration */
/* Replaces Aloam/src/scanRegistration.cpp:132-408. */
int lmono_scan_register(lmono_ctx* ctx, lmono_cloud_view raw, lmono_cloud_out* full,
                        lmono_cloud_out* sharp, lmono_cloud_out* less_sharp,
                        lmono_cloud_out* flat, lmono_cloud_out* less_flat,
                        int32_t* labels /*may be NULL; one per point of `full`*/,
                        lmono_scan_report* report);

/* test hook: curvature (scanRegistration.cpp:262) and index into the raw input of the first n
 * points of the last sweep's ring-sorted cloud */
/* Fused sweep: the three A-LOAM stages of one sweep in one call (a host that owns scanRegistration, laserOdometry and
 * laserMapping of a sequence; replaces the topic hand-offs scanRegistration.cpp:413-441 -> laserOdometry.cpp:511-590 ->
 * laserMapping.cpp:204-305).  The raw sweep is uploaded once; feature clouds and the odometry pose stay in device
 * memory between the stages.  Outputs (all optional): q/t_last_curr and q/t_w_curr of the odometry, q/t_w_curr and
 * q/t_wmap_wodom of the mapping, the three stage reports.  Bit-identical to lmono_scan_register -> lmono_odom_step ->
 * lmono_map_step(less_sharp, less_flat, odometry pose). */
int lmono_sweep_step(lmono_ctx* ctx, lmono_cloud_view raw, lmono_pose* odom_last_curr, lmono_pose* odom_w_curr,
                     lmono_pose* map_w_curr, lmono_pose* wmap_wodom,
                     lmono_scan_report* scan_report, lmono_odom_report* odom_report, lmono_map_report* map_report);
/* The same without any host synchronisation inside the sweep: lmono_sweep_submit enqueues upload, the three stages and the
 * result read-backs (the feature counts stay on the device; grids are sized from bounds), lmono_sweep_wait blocks until the
 * sweep is done and returns what lmono_sweep_step returns (bit-identical).  One sweep per ctx may be outstanding and no other
 * call on the ctx may come between the two.  own_stream != 0: a ctx that was created on a caller-supplied stream shared
 * with other sequences runs the sweep on a private stream (ordered after what the caller's stream holds at submit time;
 * later work is ordered by lmono_sweep_wait, which blocks the host), so that the sweeps of several ctxs overlap.  A page-locked `raw` buffer must stay valid until
 * the wait.  lmono_sweep_step_batch = submit for every ctx, then wait for every ctx (BASELINE config C-4: n independent
 * sequences per GPU, fused L1 -> L2 -> L3); array arguments have n entries and may be NULL. */
int lmono_sweep_submit(lmono_ctx* ctx, lmono_cloud_view raw, int own_stream);
int lmono_sweep_wait(lmono_ctx* ctx, lmono_pose* odom_last_curr, lmono_pose* odom_w_curr, lmono_pose* map_w_curr, lmono_pose* wmap_wodom,
                     lmono_scan_report* scan_report, lmono_odom_report* odom_report, lmono_map_report* map_report);
int lmono_sweep_step_batch(lmono_ctx* const* ctxs, int32_t n, const lmono_cloud_view* raws, lmono_pose* odom_w_curr, lmono_pose* map_w_curr,
                           lmono_scan_report* scan_reports, lmono_odom_report* odom_reports, lmono_map_report* map_reports);
int lmono_scan_debug(lmono_ctx* ctx, float* curvature, int32_t* src_index, int32_t n);

/* ------------------------------------------------------------------ L2: laserOdometry */
/* Replaces Aloam/src/laserOdometry.cpp:265-568. */
int lmono_odom_step(lmono_ctx* ctx, lmono_cloud_view sharp, lmono_cloud_view less_sharp,
                    lmono_cloud_view flat, lmono_cloud_view less_flat,
                    lmono_pose* last_curr /*out*/, lmono_pose* w_curr /*out*/, lmono_odom_report* report);
int lmono_odom_reset(lmono_ctx* ctx);
/* test hook: correspondences of association pass 0 / 1 of the last step; corner_idx is
 * [n_sharp x 2] (closestPointInd, minPointInd2), plane_idx [n_flat x 3] (+ minPointInd3); -1 = none */
int lmono_odom_debug(lmono_ctx* ctx, int32_t pass, int32_t* corner_idx, int32_t n_sharp, int32_t* plane_idx, int32_t n_flat);

/* ------------------------------------------------------------------ L6: colour projection */
/* Replaces mono_lidar_mapping/src/map_build_node.cc:216-225 (LiDAR -> camera extrinsic transform,
 * when T_cam_lidar is given: the node's 3x4 row-major `transformation` = [rlc^T | -rlc^T tlc]),
 * src/map_builder/Map_Builder.cc:224-245 (8-bit inverse-depth raster, last point wins), :336-403
 * (depthFill) and :275-322 (per-pixel lift, colour fetch, world transform by Q_T).
 * Outputs: depth_raw / depth_filled are width*height bytes; the clouds are packed xyz floats and
 * rgb bytes in row-major pixel order (the order of rgb_cloud / w_cloud in the reference). */
int lmono_project_color(lmono_ctx* ctx, lmono_cloud_view pts, const double* T_cam_lidar /*[12] or NULL*/,
                        const uint8_t* bgr, int32_t step_bytes,
                        const lmono_pinhole* cam, const lmono_pose* Q_T,
                        uint8_t* depth_raw /*may be NULL*/, uint8_t* depth_filled /*may be NULL*/,
                        float* cloud_cam_xyz /*may be NULL*/, float* cloud_world_xyz, uint8_t* cloud_rgb,
                        int32_t capacity, int32_t* n_out);

/* Per-point projection of the most recent lmono_project_color call, cloud order: u, v (the cv::Point2f of
 * Map_Builder.cc:234; NaN where the point was not rasterised) and camera-frame depth z.  The ~pro_map debug image
 * (Map_Builder.cc:240-265: r = 3 HSV discs drawn in cloud order) is drawn from it on the host with the reference's own
 * cv::circle call (nodes/map_build_node_b200.cpp). */
int lmono_color_projection(lmono_ctx* ctx, float* uvz /*[n][3]*/, int32_t n);

#ifdef __cplusplus
}
#endif
#endif
