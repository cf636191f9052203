// ROS-free replay harness: runs the three A-LOAM stages through the lmono C ABI on a directory of
// KITTI-layout sweeps (float32 x, y, z, intensity per point, the format Aloam/src/kittiHelper.cpp:25-35,
// 140-151 reads and publishes on /velodyne_points), in the order the ROS graph would
// (ascanRegistration -> alaserOdometry -> alaserMapping with mapping_skip_frame = 1), and prints one
// line per sweep:  index  odom(qx qy qz qw tx ty tz)  mapped(qx qy qz qw tx ty tz)  counts.
// Build: make -C nodes   (needs only g++ and lmono_b200/csrc/liblmono_b200.so).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "lmono.h"

static bool read_bin(const std::string& path, std::vector<float>& buf) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  const long bytes = ftell(f);
  fseek(f, 0, SEEK_SET);
  buf.resize(static_cast<size_t>(bytes) / sizeof(float));
  const size_t got = fread(buf.data(), sizeof(float), buf.size(), f);
  fclose(f);
  return got == buf.size();
}

#define CHECK(expr) do { int _rc = (expr); if (_rc != LMONO_OK) { fprintf(stderr, "%s -> %d (%s)\n", #expr, _rc, lmono_strerror(_rc)); return 2; } } while (0)

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s <dir with 000000.bin ...> <n_sweeps> [scan_line=64] [minimum_range=5] [fused]\n", argv[0]); return 1; }
  const bool fused = argc > 5 && strcmp(argv[5], "fused") == 0;     // one lmono_sweep_step per sweep instead of the three stage calls
  const std::string dir = argv[1];
  const int n_sweeps = atoi(argv[2]);
  lmono_params prm;
  lmono_default_params(&prm);
  if (argc > 3) prm.scan_line = atoi(argv[3]);
  if (argc > 4) prm.minimum_range = static_cast<float>(atof(argv[4]));
  lmono_ctx* ctx = nullptr;
  CHECK(lmono_create(0, &prm, nullptr, &ctx));

  std::vector<float> raw, full, sharp(64 * 12 * 4), less_sharp(64 * 120 * 4), flat(64 * 24 * 4), less_flat, registered;
  for (int k = 0; k < n_sweeps; ++k) {
    char name[64]; snprintf(name, sizeof(name), "/%06d.bin", k);
    if (!read_bin(dir + name, raw)) { fprintf(stderr, "cannot read %s%s\n", dir.c_str(), name); return 1; }
    const int n = static_cast<int>(raw.size() / 4);
    full.resize(static_cast<size_t>(n) * 4 + 4); less_flat.resize(static_cast<size_t>(n) * 4 + 4); registered.resize(static_cast<size_t>(n) * 4 + 4);
    lmono_cloud_view v_raw = { raw.data(), n, 16, 12 };
    if (fused) {
      lmono_pose last_curr, odom, mapped, wm; lmono_scan_report srep; lmono_odom_report orep; lmono_map_report mrep;
      CHECK(lmono_sweep_step(ctx, v_raw, &last_curr, &odom, &mapped, &wm, &srep, &orep, &mrep));
      printf("%d  %.9f %.9f %.9f %.9f %.6f %.6f %.6f  %.9f %.9f %.9f %.9f %.6f %.6f %.6f  kept %d sharp %d flat %d corr %d %d map %d %d opt %d\n",
             k, odom.q[0], odom.q[1], odom.q[2], odom.q[3], odom.t[0], odom.t[1], odom.t[2],
             mapped.q[0], mapped.q[1], mapped.q[2], mapped.q[3], mapped.t[0], mapped.t[1], mapped.t[2],
             srep.n_kept, srep.n_sharp, srep.n_flat, orep.corner_corr[1], orep.plane_corr[1],
             mrep.corner_from_map, mrep.surf_from_map, mrep.optimized);
      continue;
    }
    lmono_cloud_out o_full = { full.data(), n, 16, 12, 0 }, o_sharp = { sharp.data(), 64 * 12, 16, 12, 0 },
                    o_ls = { less_sharp.data(), 64 * 120, 16, 12, 0 }, o_flat = { flat.data(), 64 * 24, 16, 12, 0 },
                    o_lf = { less_flat.data(), n, 16, 12, 0 };
    lmono_scan_report srep;
    CHECK(lmono_scan_register(ctx, v_raw, &o_full, &o_sharp, &o_ls, &o_flat, &o_lf, nullptr, &srep));
    lmono_cloud_view v_sharp = { sharp.data(), o_sharp.n_out, 16, 12 }, v_ls = { less_sharp.data(), o_ls.n_out, 16, 12 },
                     v_flat = { flat.data(), o_flat.n_out, 16, 12 }, v_lf = { less_flat.data(), o_lf.n_out, 16, 12 },
                     v_full = { full.data(), o_full.n_out, 16, 12 };
    lmono_pose last_curr, odom; lmono_odom_report orep;
    CHECK(lmono_odom_step(ctx, v_sharp, v_ls, v_flat, v_lf, &last_curr, &odom, &orep));
    // /laser_cloud_corner_last = this sweep's less-sharp, /laser_cloud_surf_last = less-flat (laserOdometry.cpp:554-590)
    lmono_pose mapped, wm; lmono_map_report mrep;
    lmono_cloud_out o_reg = { registered.data(), n, 16, 12, 0 };
    CHECK(lmono_map_step(ctx, v_ls, v_lf, &odom, &mapped, &wm, &mrep, v_full, &o_reg));
    printf("%d  %.9f %.9f %.9f %.9f %.6f %.6f %.6f  %.9f %.9f %.9f %.9f %.6f %.6f %.6f  kept %d sharp %d flat %d corr %d %d map %d %d opt %d\n",
           k, odom.q[0], odom.q[1], odom.q[2], odom.q[3], odom.t[0], odom.t[1], odom.t[2],
           mapped.q[0], mapped.q[1], mapped.q[2], mapped.q[3], mapped.t[0], mapped.t[1], mapped.t[2],
           srep.n_kept, srep.n_sharp, srep.n_flat, orep.corner_corr[1], orep.plane_corr[1],
           mrep.corner_from_map, mrep.surf_from_map, mrep.optimized);
  }
  lmono_destroy(ctx);
  return 0;
}
