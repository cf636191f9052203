// ORACLE (test infrastructure).  The reference relies on two UNSTABLE libstdc++
// std::sort calls whose tie order is an introsort artefact:
//   * pcl::VoxelGrid sorts cloud_point_index_idx by idx only (PCL 1.8 voxel_grid.hpp;
//     call sites Aloam/src/laserMapping.cpp:542-550,788-801, scanRegistration.cpp:401-405)
//   * Aloam/src/scanRegistration.cpp:71,288  std::sort(cloudSortInd+sp, cloudSortInd+ep+1, comp)
// These shims call the real std::sort of this toolchain with the same element type and
// comparator so the oracle can reproduce that order ("order_mode 1" / "sort_mode 1").
#include <algorithm>
#include <vector>
#include <cstdint>
#include "lmono_oracle.h"

namespace {
struct cloud_point_index_idx {
  unsigned int idx;
  unsigned int cloud_point_index;
  bool operator<(const cloud_point_index_idx& p) const { return idx < p.idx; }
};
const float* g_curv = nullptr;
bool comp(int i, int j) { return g_curv[i] < g_curv[j]; }
}  // namespace

extern "C" void lmono_cpu_stdsort_voxel_pairs(uint32_t* idx, uint32_t* pt, int n) {
  std::vector<cloud_point_index_idx> v(static_cast<size_t>(n));
  for (int i = 0; i < n; ++i) { v[i].idx = idx[i]; v[i].cloud_point_index = pt[i]; }
  std::sort(v.begin(), v.end(), std::less<cloud_point_index_idx>());
  for (int i = 0; i < n; ++i) { idx[i] = v[i].idx; pt[i] = v[i].cloud_point_index; }
}

extern "C" void lmono_cpu_stdsort_by_curvature(int32_t* ind, int n, const float* curvature) {
  g_curv = curvature;
  std::sort(ind, ind + n, comp);
  g_curv = nullptr;
}
