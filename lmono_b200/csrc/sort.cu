// Device sort of unique 64-bit keys whose count lives in device memory.
// Two launches, no host synchronisation:
//   k_sort_tiles : each CTA bitonic-sorts one 4096-key tile held in registers (8 keys/thread;
//                  in-thread, warp-shuffle and only 10 shared-memory stages of 78)
//   k_merge_ranks: every key finds its global rank = own position + sum over the other
//                  tiles of lower_bound(tile, key)  (keys are unique), and scatters.
// Used for VoxelGrid keys (PCL sorts cloud_point_index_idx, voxel_grid.hpp), cube
// insertion order and map import.  Keys are (sort key << 32 | original index) composites,
// so the result equals a STABLE sort by key -- the canonical order DESIGN.md defines in
// place of libstdc++'s unstable std::sort.
#include "common.cuh"

__global__ void __launch_bounds__(512) k_sort_tiles(const unsigned long long* __restrict__ in,
                                                    unsigned long long* __restrict__ tmp,
                                                    unsigned long long* __restrict__ out,
                                                    const int32_t* __restrict__ n_dev) {
  __shared__ unsigned long long s[LM_SORT_TILE];
  const int n = *n_dev;
  const int base = blockIdx.x * LM_SORT_TILE;
  if (base >= n) return;
  constexpr int ITEMS = LM_SORT_TILE / 512;
  // coalesced load through shared memory into the blocked register arrangement
  for (int i = threadIdx.x; i < LM_SORT_TILE; i += blockDim.x) s[i] = (base + i < n) ? in[base + i] : ~0ULL;
  __syncthreads();
  unsigned long long v[ITEMS];
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) v[r] = s[threadIdx.x * ITEMS + r];
  __syncthreads();
  d_bitonic_regs<ITEMS>(v, threadIdx.x, 512, s);
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) s[threadIdx.x * ITEMS + r] = v[r];
  __syncthreads();
  unsigned long long* dst = (n <= LM_SORT_TILE) ? out : tmp;
  for (int i = threadIdx.x; i < LM_SORT_TILE; i += blockDim.x) if (base + i < n) dst[base + i] = s[i];
}

__global__ void __launch_bounds__(256) k_merge_ranks(const unsigned long long* __restrict__ tmp,
                                                     unsigned long long* __restrict__ out,
                                                     const int32_t* __restrict__ n_dev) {
  const int n = *n_dev;
  if (n <= LM_SORT_TILE) return;   // single tile already written to out
  const int ntiles = (n + LM_SORT_TILE - 1) / LM_SORT_TILE;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const unsigned long long key = tmp[e];
    const int my_tile = e / LM_SORT_TILE;
    int rank = e - my_tile * LM_SORT_TILE;
    for (int t = 0; t < ntiles; ++t) {
      if (t == my_tile) continue;
      const int tb = t * LM_SORT_TILE;
      const int tn = min(LM_SORT_TILE, n - tb);
      rank += d_lower_bound_u64(tmp + tb, tn, key);
    }
    out[rank] = key;
  }
}

int lm_sort_u64(lmono_ctx* ctx, const unsigned long long* in, unsigned long long* tmp, unsigned long long* out,
                const int32_t* n_dev, int n_max) {
  if (n_max <= 0) return LMONO_OK;
  const int ntiles = lm_div_up(n_max, LM_SORT_TILE);
  k_sort_tiles<<<ntiles, 512, 0, ctx->stream>>>(in, tmp, out, n_dev);
  LM_LAUNCH_CHECK();
  if (ntiles > 1) {
    int blocks = lm_div_up(n_max, 256);
    k_merge_ranks<<<blocks, 256, 0, ctx->stream>>>(tmp, out, n_dev);
    LM_LAUNCH_CHECK();
  }
  return LMONO_OK;
}
