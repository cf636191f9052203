// Device sort of unique 64-bit keys whose counts live in device memory; up to LM_SORT_MAXSEG
// independent segments per call (e.g. the corner and the surf VoxelGrid of one sweep share launches).
// Two launches, no host synchronisation:
//   k_sort_tiles : each CTA bitonic-sorts one 1024-key tile held in registers (4 keys/thread x 256
//                  threads: 3 in-thread + 5 warp-shuffle partner distances, only 6 of the 66 stages go
//                  through shared memory).  Small tiles on many SMs: the network is issue-bound, so
//                  16 k keys on 8 SMs finish in a fraction of the time of 4 SMs x 4096.
//   k_merge_ranks: one pass merges groups of 16 sorted runs: every key finds its rank = own position +
//                  sum over the other runs of its group of lower_bound(run, key) (keys are unique) and
//                  scatters.  The searches over different runs are independent: four run interleaved
//                  per thread so their (L1/L2-resident) loads overlap instead of forming one long
//                  dependent chain.  <= 32 k keys need one pass, 2 M keys (map import) three.
// Used for VoxelGrid keys (PCL sorts cloud_point_index_idx, voxel_grid.hpp), cube insertion
// order and map import.  Keys are (sort key << k | original index) composites, so the result
// equals a STABLE sort by key -- the canonical order DESIGN.md defines in place of libstdc++'s
// unstable std::sort.
#include "common.cuh"

constexpr int ST_THREADS = 256;
constexpr int ST_ITEMS = LM_SORT_TILE / ST_THREADS;
static_assert(LM_SORT_TILE == 1024 && ST_ITEMS == 4, "tile geometry");

__global__ void __launch_bounds__(ST_THREADS) k_sort_tiles(LmSortSegs sg, int dst_is_tmp) {
  lm_pdl_enter();
  __shared__ unsigned long long s[LM_SORT_TILE];
  const int seg = blockIdx.y;
  const int n = *sg.n[seg];
  const int base = blockIdx.x * LM_SORT_TILE;
  if (base >= n) return;
  const unsigned long long* __restrict__ in = sg.in + sg.off[seg];
  // coalesced load through shared memory into the blocked register arrangement
  for (int i = threadIdx.x; i < LM_SORT_TILE; i += ST_THREADS) s[i] = (base + i < n) ? in[base + i] : ~0ULL;
  __syncthreads();
  unsigned long long v[ST_ITEMS];
#pragma unroll
  for (int r = 0; r < ST_ITEMS; ++r) v[r] = s[threadIdx.x * ST_ITEMS + r];
  __syncthreads();
  d_bitonic_regs<ST_ITEMS, ST_THREADS>(v, threadIdx.x, s);
#pragma unroll
  for (int r = 0; r < ST_ITEMS; ++r) s[threadIdx.x * ST_ITEMS + r] = v[r];
  __syncthreads();
  unsigned long long* dst = (dst_is_tmp ? sg.tmp : sg.out) + sg.off[seg];
  for (int i = threadIdx.x; i < LM_SORT_TILE; i += ST_THREADS) if (base + i < n) dst[base + i] = s[i];
}

// One merge pass: sorted runs of `run` keys (a power of two) are merged in groups of LM_MERGE_GROUP into
// runs of LM_MERGE_GROUP * run keys.  Every key ranks itself inside the other runs of its group with
// branch-free binary searches, four of them interleaved so their loads overlap.
constexpr int LM_MERGE_GROUP = 16;

__global__ void __launch_bounds__(256) k_merge_ranks(LmSortSegs sg, int run, int src_is_tmp, const int32_t* __restrict__ done, int pass) {
  lm_pdl_enter();
  const int seg = blockIdx.y;
  const int n = *sg.n[seg];
  const unsigned long long* __restrict__ src = (src_is_tmp ? sg.tmp : sg.out) + sg.off[seg];
  unsigned long long* __restrict__ dst = (src_is_tmp ? sg.out : sg.tmp) + sg.off[seg];
  if (done && done[seg]) {
    // the segment turned out small enough for the single shared-memory pass (k_merge_ranks_smem below), which left the
    // sorted keys where pass 0 would have: pass 0 has nothing to do, the later passes only carry them to the next buffer
    if (pass > 0) for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) dst[e] = src[e];
    return;
  }
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const unsigned long long key = src[e];
    const int my_run = e / run;
    const int g0 = (my_run / LM_MERGE_GROUP) * LM_MERGE_GROUP;        // first run of my group
    int rank = g0 * run + (e - my_run * run);
#pragma unroll
    for (int b = 0; b < LM_MERGE_GROUP; b += 4) {
      int pos[4] = { 0, 0, 0, 0 }, tn[4];
      const unsigned long long* a[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int t = g0 + b + u;
        const long long start = (long long)t * run;
        const bool on = t != my_run && start < n;
        a[u] = src + (on ? start : 0);
        tn[u] = on ? (int)min((long long)run, (long long)n - start) : 0;
      }
      for (int s = run; s > 0; s >>= 1) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int probe = pos[u] + s;
          if (probe <= tn[u] && a[u][probe - 1] < key) pos[u] = probe;
        }
      }
      rank += pos[0] + pos[1] + pos[2] + pos[3];
    }
    dst[rank] = key;
  }
}

// Single-pass variant for n <= LM_MERGE_SMEM_RUNS runs (every per-sweep sort): a dependent chain of ~40 L2
// accesses per key is what the global-memory searches cost (~0.15 us each on B200), so each CTA first copies
// ALL runs of its segment into shared memory (<= 12 x 16 KB, coalesced, one L2 pass) and ranks its 1024 keys
// against them there (~30-cycle accesses).
constexpr int LM_MERGE_SMEM_RUNS = 28;            // 28 x 8 KB = 224 KB of the 227 KB a CTA may have
constexpr int MS_THREADS = 1024;

// `done` != NULL: the launch grids were sized for a bound far above the real count (the fused sweep only knows the raw
// sweep size): the kernel decides on the device whether the segment fits this path and tells the global-memory passes
// that follow (k_merge_ranks) through done[seg].
__global__ void __launch_bounds__(MS_THREADS, 1) k_merge_ranks_smem(LmSortSegs sg, int src_is_tmp, int32_t* __restrict__ done) {
  lm_pdl_enter();
  extern __shared__ unsigned long long s_runs[];        // [nruns][LM_SORT_TILE]
  const int seg = blockIdx.y;
  const int n = *sg.n[seg];
  const int e0 = blockIdx.x * MS_THREADS;
  if (done) {
    const bool fits = n <= LM_MERGE_SMEM_RUNS * LM_SORT_TILE;
    if (blockIdx.x == 0 && threadIdx.x == 0) done[seg] = fits ? 1 : 0;
    if (!fits) return;
  }
  if (e0 >= n) return;
  const unsigned long long* __restrict__ src = (src_is_tmp ? sg.tmp : sg.out) + sg.off[seg];
  unsigned long long* __restrict__ dst = (src_is_tmp ? sg.out : sg.tmp) + sg.off[seg];
  const int nruns = (n + LM_SORT_TILE - 1) / LM_SORT_TILE;
  // stage every run of the segment: eight independent loads in flight per thread (a plain copy loop waits for each
  // load before it issues the next one: ~16 dependent L2 round trips for a 16 k-key segment)
  const int total = nruns * LM_SORT_TILE;
  for (int i0 = threadIdx.x; i0 < total; i0 += 8 * MS_THREADS) {
    unsigned long long v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int i = i0 + u * MS_THREADS; v[u] = i < n ? src[i] : ~0ULL; }
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int i = i0 + u * MS_THREADS; if (i < total) s_runs[i] = v[u]; }
  }
  __syncthreads();
  const int e = e0 + threadIdx.x;
  if (e >= n) return;
  const unsigned long long key = s_runs[e];
  const int my_run = e / LM_SORT_TILE;
  int rank = e - my_run * LM_SORT_TILE;
  // four runs at a time, their binary searches interleaved step by step (independent shared-memory loads in flight
  // instead of one chain of ~30-cycle accesses per run); runs are padded with ~0ULL: no length checks
  for (int t0 = 0; t0 < nruns; t0 += 4) {
    const unsigned long long* a[4];
    int pos[4] = { 0, 0, 0, 0 };
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = s_runs + min(t0 + u, nruns - 1) * LM_SORT_TILE;
#pragma unroll
    for (int s = LM_SORT_TILE / 2; s > 0; s >>= 1) {
#pragma unroll
      for (int u = 0; u < 4; ++u) if (a[u][pos[u] + s - 1] < key) pos[u] += s;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      pos[u] += (a[u][pos[u]] < key);                              // 10 halvings + the last element
      if (t0 + u < nruns && t0 + u != my_run) rank += pos[u];
    }
  }
  dst[rank] = key;
}

static int merge_passes(int n_max) {
  int p = 0;
  for (long long run = LM_SORT_TILE; run < n_max; run *= LM_MERGE_GROUP) ++p;
  return p;
}

int lm_sort_u64_segs(lmono_ctx* ctx, const LmSortSegs& sg, int nseg, const int* n_max) {
  int mx = 0;
  for (int s = 0; s < nseg; ++s) mx = n_max[s] > mx ? n_max[s] : mx;
  if (mx <= 0 || nseg <= 0) return LMONO_OK;
  const int ntiles = lm_div_up(mx, LM_SORT_TILE);
  if (ntiles > 1 && ntiles <= LM_MERGE_SMEM_RUNS) {
    LM_LAUNCH_PDL(k_sort_tiles, dim3(ntiles, nseg), ST_THREADS, 0, sg, 1);
    LM_LAUNCH_CHECK();
    LM_LAUNCH_PDL(k_merge_ranks_smem, dim3(lm_div_up(mx, MS_THREADS), nseg), MS_THREADS, (size_t)ntiles * LM_SORT_TILE * 8, sg, 1, (int32_t*)nullptr);
    LM_LAUNCH_CHECK();
    return LMONO_OK;
  }
  const int passes = merge_passes(mx);
  // buffers alternate tmp <-> out per pass and the last pass must land in `out`
  int cur_is_tmp = (passes & 1) ? 1 : 0;
  LM_LAUNCH_PDL(k_sort_tiles, dim3(ntiles, nseg), ST_THREADS, 0, sg, cur_is_tmp);
  LM_LAUNCH_CHECK();
  // per-sweep sorts whose grids are sized for a loose bound: try the single shared-memory pass first (decided on the
  // device from the real count); the global passes then skip / copy.  The map import (millions of keys) goes straight
  // to the global passes.
  int32_t* done = nullptr;
  if (passes >= 1 && mx <= 262144) {
    done = ctx->d_sort_done;
    LM_LAUNCH_PDL(k_merge_ranks_smem, dim3(LM_MERGE_SMEM_RUNS, nseg), MS_THREADS, (size_t)LM_MERGE_SMEM_RUNS * LM_SORT_TILE * 8, sg, cur_is_tmp, done);
    LM_LAUNCH_CHECK();
  }
  long long run = LM_SORT_TILE;
  for (int p = 0; p < passes; ++p, run *= LM_MERGE_GROUP) {
    LM_LAUNCH_PDL(k_merge_ranks, dim3(lm_div_up(mx, 256), nseg), 256, 0, sg, (int)run, cur_is_tmp, (const int32_t*)done, p);
    LM_LAUNCH_CHECK();
    cur_is_tmp ^= 1;
  }
  return LMONO_OK;
}

int lm_sort_u64(lmono_ctx* ctx, const unsigned long long* in, unsigned long long* tmp, unsigned long long* out,
                const int32_t* n_dev, int n_max) {
  LmSortSegs sg;
  sg.in = in; sg.tmp = tmp; sg.out = out;
  for (int s = 0; s < LM_SORT_MAXSEG; ++s) { sg.off[s] = 0; sg.n[s] = n_dev; }
  return lm_sort_u64_segs(ctx, sg, 1, &n_max);
}

int lm_sort_configure(lmono_ctx* ctx) {
  LM_CUDA(cudaFuncSetAttribute(k_merge_ranks_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, LM_MERGE_SMEM_RUNS * LM_SORT_TILE * 8));
  return LMONO_OK;
}
