#!/usr/bin/env python
"""Generates tests/golden/*.npz: small seeded inputs together with the outputs of the CPU oracle
(oracle/*.c, the restatement of the reference's A-LOAM / map-builder arithmetic).

The reference ships no tests or fixtures for this path; the oracle itself is pinned against the
reference's own sources compiled into oracle/_ref (tests/test_oracle_vs_ref.py, make_golden_ref.py).  These
vectors guard the oracle AND the CUDA path against silent drift:
`tests/test_golden.py` re-runs the oracle (-m "not gpu") and the C-ABI CUDA library (-m gpu) on
the stored inputs and compares with the stored outputs.  Regenerate only when the oracle is
deliberately changed:  python tests/golden/make_golden.py
Every array is stored (inputs included) so the fixtures do not depend on numpy's RNG streams.
"""
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
from lmono_b200 import synth  # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}.npz  {os.path.getsize(path) / 1024:.1f} KiB  " + ", ".join(f"{k}{list(v.shape)}" for k, v in arrays.items()))


def rep_counts(rep):
    return np.array([rep.corner_from_map, rep.surf_from_map, rep.corner_stack, rep.surf_stack,
                     rep.corner_num[0], rep.corner_num[1], rep.surf_num[0], rep.surf_num[1], rep.optimized,
                     rep.solve[0].iterations, rep.solve[1].iterations, rep.solve[0].num_successful,
                     rep.solve[1].num_successful], np.int32)


def gen_primitives():
    rng = np.random.default_rng(101)
    # pcl::VoxelGrid (laserMapping.cpp:542-550 call sites): clustered points so that voxels have several members
    base = rng.uniform(-20, 20, (900, 3))
    pts = (base[rng.integers(0, 900, 4000)] + rng.normal(0, 0.15, (4000, 3))).astype(np.float32)
    pts = np.concatenate([pts, rng.uniform(0, 64, (4000, 1)).astype(np.float32)], axis=1)
    vg04 = O.voxel_grid(pts, 0.4)
    vg08 = O.voxel_grid(pts, 0.8)
    # FLANN-equivalent exact 5-NN (laserMapping.cpp:582,648)
    cloud = np.concatenate([rng.uniform(-15, 15, (20000, 3)), np.zeros((20000, 1))], axis=1).astype(np.float32)
    qs = np.concatenate([rng.uniform(-15, 15, (300, 3)), np.zeros((300, 1))], axis=1).astype(np.float32)
    idx, d2 = O.knn_brute(cloud, qs, 5)
    idx_t, d2_t = O.knn_kdtree(cloud, qs, 5)
    assert np.array_equal(idx, idx_t) and np.array_equal(d2, d2_t)
    # Eigen SelfAdjointEigenSolver<Matrix3d> / colPivHouseholderQr 5x3 (laserMapping.cpp:605-611,663)
    covs, ws, Vs, As, xs = [], [], [], [], []
    for _ in range(16):
        P = rng.normal(0, 1, (5, 3)) * rng.uniform(0.01, 2.0, 3)
        Z = P - P.mean(0)
        cov = Z.T @ Z
        w, V, rc = O.eigh3(cov)
        assert rc == 0
        covs.append(cov); ws.append(w); Vs.append(V)
        A = rng.normal(0, 1, (5, 3)) + rng.uniform(-5, 5, 3)
        As.append(A); xs.append(O.colpiv_solve_5x3(A, -np.ones(5)))
    save("primitives", vg_in=pts, vg_out_04=vg04, vg_out_08=vg08, knn_cloud=cloud, knn_queries=qs, knn_idx=idx, knn_d2=d2,
         eig_cov=np.array(covs), eig_w=np.array(ws), eig_V=np.array(Vs), qr_A=np.array(As), qr_x=np.array(xs))


def gen_scan():
    w = synth.make_world()
    out = {}
    # n_az is deliberately NOT a multiple of 4: on a regular azimuth grid a point exactly 90 deg from the last
    # point sits on the `ori > endOri + pi/2` branch of the unwrap (scanRegistration.cpp:226-233), where the last
    # ulp of atan2f (libm build vs device) decides a 2*pi jump of relTime -- the hazard DESIGN.md documents
    for n_scans, min_range, n_az in ((64, 5.0, 301), (32, 0.3, 301), (16, 0.3, 301)):
        q, t = synth.loop_pose(w, 5.0)
        raw = synth.raycast_sweep(w, q, t, n_scans, n_az, np.random.default_rng(40 + n_scans)).astype(np.float32)
        r = O.scan_register(raw, n_scans, min_range)
        rep = r["report"]
        k = f"s{n_scans}_"
        out[k + "raw"] = raw
        out[k + "labels"] = r["labels"].astype(np.int8)
        out[k + "src_index"] = r["src_index"]
        out[k + "curvature"] = r["curvature"]
        out[k + "ring_start"] = np.array(list(rep.ring_start), np.int32)
        out[k + "ring_end"] = np.array(list(rep.ring_end), np.int32)
        out[k + "counts"] = np.array([rep.n_in, rep.n_kept, rep.n_sharp, rep.n_less_sharp, rep.n_flat, rep.n_less_flat], np.int32)
        out[k + "less_flat"] = r["less_flat"]
        out[k + "full_intensity"] = r["full"][:, 3]
    save("scan_registration", **out)


def gen_odometry():
    w = synth.make_world()
    rng = np.random.default_rng(77)
    od = O.Odometry()
    out = {}
    feats = []
    for k in range(3):
        q, t = synth.loop_pose(w, 0.8 * k)
        raw = synth.raycast_sweep(w, q, t, 64, 451, rng).astype(np.float32)
        r = O.scan_register(raw, 64, 5.0)
        feats.append(r)
        (lq, lt), (wq, wt), rep = od.step(r["sharp"], r["less_sharp"], r["flat"], r["less_flat"])
        for name in ("sharp", "less_sharp", "flat", "less_flat"):
            out[f"f{k}_{name}"] = r[name]
        out[f"f{k}_last_curr"] = np.concatenate([lq, lt])
        out[f"f{k}_w_curr"] = np.concatenate([wq, wt])
        out[f"f{k}_corr"] = np.array([rep.inited, rep.corner_corr[0], rep.corner_corr[1], rep.plane_corr[0], rep.plane_corr[1]], np.int32)
    # association pass at the identity guess between frames 0 -> 1 (laserOdometry.cpp:299-483)
    ci, pi = O.odom_associate(feats[1]["sharp"], feats[1]["flat"], feats[0]["less_sharp"], feats[0]["less_flat"], (0, 0, 0, 1), (0, 0, 0))
    out["assoc01_corner"] = ci
    out["assoc01_plane"] = pi
    save("odometry", **out)


def gen_mapping():
    w = synth.make_world()
    _, t0 = synth.loop_pose(w, 0.0)
    cm, sm = synth.sample_map(w, t0, half_xy=40.0, n_surf=60_000, n_corner=15_000)
    cm = cm.astype(np.float32); sm = sm.astype(np.float32)
    om = O.Mapper()
    om.import_points(0, cm); om.import_points(1, sm)
    out = {"map_corner_in": cm, "map_surf_in": sm}
    rng = np.random.default_rng(9)
    # 5-NN in the window index space (:533-537) + one association / normal-equation evaluation
    q, t = synth.loop_pose(w, 0.0)
    c, s = synth.sample_sweep_features(w, q, t, rng, 800, 3000)
    c = c.astype(np.float32); s = s.astype(np.float32)
    om.prepare_window(t)
    cs, ss = O.voxel_grid(c, 0.4), O.voxel_grid(s, 0.8)
    from lmono_b200.synth import quat_to_rot
    R = quat_to_rot(q)
    qw = ss.copy(); qw[:, :3] = (ss[:, :3].astype(np.float64) @ R.T + t).astype(np.float32)
    idx, d2 = om.knn5(1, qw)
    out["knn_queries_world"] = qw; out["knn_idx"] = idx; out["knn_d2"] = d2
    fac, nc, ns = om.associate(cs, ss, q, t)
    H, g, cost = O.normal_eq(fac, q, t)
    out["ne_pose"] = np.concatenate([q, t]); out["ne_H"] = H; out["ne_g"] = g
    out["ne_cost_nc_ns"] = np.array([cost, nc, ns])
    out["ne_corner_stack"] = cs; out["ne_surf_stack"] = ss
    # three full registrations incl. map update (:307-801)
    for k in range(3):
        q, t = synth.loop_pose(w, 1.0 * k)
        c, s = synth.sample_sweep_features(w, q, t, rng, 800, 3000)
        qp, tp = synth.perturb_pose(q, t, rng, 0.2, 1.0)
        c = c.astype(np.float32); s = s.astype(np.float32)
        mq, mt, rep, _ = om.step(c, s, qp, tp)
        out[f"k{k}_corner"] = c; out[f"k{k}_surf"] = s
        out[f"k{k}_odom"] = np.concatenate([qp, tp]); out[f"k{k}_w_curr"] = np.concatenate([mq, mt])
        out[f"k{k}_counts"] = rep_counts(rep)
    out["map_corner_out"] = om.export(0, 1)
    out["map_surf_out"] = om.export(1, 1)
    save("mapping", **out)


def gen_color():
    rng = np.random.default_rng(5)
    W, H = 310, 94            # quarter-size KITTI raster keeps the fixture small; same code path
    cam = O.make_camera(fx=179.714, fy=179.714, cx=151.798, cy=46.304, width=W, height=H)
    n = 6000
    z = rng.uniform(2.0, 60.0, n)
    x = rng.uniform(-1, 1, n) * z * 0.9
    y = rng.uniform(-0.3, 0.3, n) * z
    pts_cam = np.stack([x, y, z], 1).astype(np.float32)
    bgr = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    raw = O.project_raster(pts_cam, cam)
    fill = O.depth_fill(raw, cam)
    q = np.array([0.01, -0.02, 0.03, 1.0]); q /= np.linalg.norm(q)
    t = np.array([1.0, 2.0, 3.0])
    cc, cw, rgb = O.lift_cloud(fill, bgr, cam, q, t)
    save("color", pts_cam=pts_cam, bgr=bgr, cam=np.array([179.714, 179.714, 151.798, 46.304, W, H]), pose=np.concatenate([q, t]),
         depth_raw=raw, depth_filled=fill, cloud_cam=cc, cloud_world=cw, rgb=rgb)


if __name__ == "__main__":
    gen_primitives()
    gen_scan()
    gen_odometry()
    gen_mapping()
    gen_color()
