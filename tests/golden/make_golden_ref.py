#!/usr/bin/env python
"""Generates tests/golden/ref_nodes.npz: small seeded inputs together with the outputs of THE REFERENCE'S OWN CODE --
the scanRegistration / laserOdometry / laserMapping nodes and the colour mapper compiled from /root/reference into
oracle/_ref (`make -C oracle ref`; library stand-ins underneath, see oracle/refstubs/README.md and DESIGN.md section 2).

Unlike oracle/_ref itself these vectors are committed, so `tests/test_golden.py` can hold the oracle (-m "not gpu") and
the CUDA library (-m gpu) against reference-generated outputs on any machine, with no reference tree and no prebuilt
library at hand.  Needs /root/reference (run in the build container):  python tests/golden/make_golden_ref.py
Every array is stored (inputs included) so the fixtures do not depend on numpy's RNG streams."""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle_lib as O  # noqa: E402
import scenario  # noqa: E402
from lmono_b200 import synth  # noqa: E402


def main():
    for name in ("scanreg", "odom", "mapping", "color"):
        assert O.ref_lib(name) is not None, f"oracle/_ref/libref_{name}.so missing: this script needs the reference tree"
    out = {}
    w = synth.make_world()

    # --- scanRegistration node: one 64-ring and one 16-ring sweep
    for tag, n_scans, min_range, n_az, s in (("s64", 64, 5.0, 500, 7.0), ("s16", 16, 0.3, 900, 9.0)):
        rng = np.random.default_rng(1000 + n_scans)
        q, t = synth.loop_pose(w, s)
        raw = np.ascontiguousarray(synth.raycast_sweep(w, q, t, n_scans, n_az, rng)[:, :3], np.float32)
        raw[3, 2] = np.nan
        r = O.ref_scan_register(raw, n_scans, min_range)
        out[f"{tag}_raw"] = raw
        out[f"{tag}_cfg"] = np.array([n_scans, min_range], np.float64)
        for k in ("full", "labels", "curvature", "sharp", "less_sharp", "flat", "less_flat"):
            out[f"{tag}_{k}"] = r[k].astype(np.int8) if k == "labels" else r[k]

    # --- laserOdometry node: six consecutive 16-ring sweeps (features from the reference's own scanRegistration)
    ref_od = O.RefOdometry()
    rng = np.random.default_rng(77)
    poses = []
    for k in range(6):
        q, t = synth.loop_pose(w, 1.0 * k)
        raw = synth.raycast_sweep(w, q, t, 16, 900, rng)
        f = O.ref_scan_register(raw, 16, 0.3)
        for name in ("sharp", "less_sharp", "flat", "less_flat"):
            out[f"od{k}_{name}"] = f[name]
        (lq, lt), (wq, wt), cnt = ref_od.step(f["sharp"], f["less_sharp"], f["flat"], f["less_flat"], f["full"])
        poses.append(np.concatenate([lq, lt, wq, wt, cnt.astype(np.float64)]))
    out["od_poses"] = np.array(poses)                      # q_last_curr t_last_curr q_w_curr t_w_curr corner_corr plane_corr

    # --- laserMapping node: six sweeps from an empty map
    ref_map = O.RefMapper()
    mp = []
    for k, (c, s, qg, tg, qo, to) in enumerate(scenario.sweeps(6, n_corner=700, n_surf=4000)):
        full = np.concatenate([c, s])[:1500]
        rq, rt, (wq, wt), cen, rfull = ref_map.step(c, s, qo, to, full)
        out[f"mp{k}_corner"] = c.astype(np.float32)
        out[f"mp{k}_surf"] = s.astype(np.float32)
        out[f"mp{k}_full"] = full.astype(np.float32)
        out[f"mp{k}_registered"] = rfull
        mp.append(np.concatenate([qo, to, rq, rt, wq, wt, np.array(cen, np.float64)]))
    out["mp_poses"] = np.array(mp)                         # q_odom t_odom q_w_curr t_w_curr q_wmap_wodom t_wmap_wodom cen
    out["mp_map_corner"] = ref_map.export(0)
    out["mp_map_surf"] = ref_map.export(1)

    # --- colour mapper: one frame, small image
    cam = O.make_camera(width=320, height=96, cx=160.0, cy=48.0, fx=180.0, fy=180.0, kernel_type=2, kernel_size=5, blur_type=0)
    rng = np.random.default_rng(5)
    n = 9000
    pts = np.zeros((n, 3), np.float32)
    pts[:, 2] = rng.uniform(-5, 120, n)
    pts[:, 0] = rng.uniform(-0.9, 0.9, n) * pts[:, 2]
    pts[:, 1] = rng.uniform(-0.27, 0.27, n) * pts[:, 2]
    img = rng.integers(0, 256, (cam.height, cam.width, 3), dtype=np.uint8)
    q = np.array([0.02, -0.05, 0.2, 0.97])
    q /= np.linalg.norm(q)
    t = np.array([4.0, -2.0, 0.7])
    raw, filled, cc, cw, rgb = O.ref_color_frame(pts, img, cam, q, t)
    out.update(col_pts=pts, col_img=img, col_cam=np.array([cam.fx, cam.fy, cam.cx, cam.cy, cam.k1, cam.k2, cam.p1, cam.p2, cam.width, cam.height,
                                                             cam.kernel_type, cam.kernel_size, cam.blur_type], np.float64),
               col_pose=np.concatenate([q, t]), col_raw=raw, col_filled=filled, col_cloud_cam=cc, col_cloud_world=cw, col_rgb=rgb)

    path = os.path.join(HERE, "ref_nodes.npz")
    np.savez_compressed(path, **out)
    print(f"ref_nodes.npz  {os.path.getsize(path) / 1024:.1f} KiB, {len(out)} arrays")


if __name__ == "__main__":
    main()
