#!/usr/bin/env python
"""Wall-clock time of every stage of the pipeline through the host API (host buffers in, host buffers out) on
synthetic HDL-64 sweeps: scanRegistration -> laserOdometry -> laserMapping (+ colour projection of the sweep)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from lmono_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
w = synth.make_world()
rng = np.random.default_rng(2)
raws = []
for k in range(n):
    q, t = synth.loop_pose(w, 1.0 * k)
    raws.append(synth.raycast_sweep(w, q, t, 64, 1875, rng))
ctx = api.Context(device=0)
cam = api.Pinhole(718.856, 718.856, 607.1928, 185.2157, 0, 0, 0, 0, 1241, 376, 0, 5, 0)
img = np.random.default_rng(5).integers(0, 255, (376, 1241, 3), dtype=np.uint8)
T = {"scan": [], "odom": [], "map": [], "color": []}
G = {"scan": [], "odom": [], "map": []}
for k, raw in enumerate(raws):
    t0 = time.perf_counter()
    r = ctx.scan_register(raw)
    t1 = time.perf_counter()
    (_, _), (oq, ot), orep = ctx.odom_step(r["sharp"], r["less_sharp"], r["flat"], r["less_flat"])
    t2 = time.perf_counter()
    mq, mt, mrep, _ = ctx.map_step(r["less_sharp"], r["less_flat"], oq, ot)
    t3 = time.perf_counter()
    pts = raw[:, :3].copy(); pts = pts[:, [1, 2, 0]] * np.array([-1, -1, 1], np.float32)      # LiDAR -> camera axes
    ctx.project_color(pts, img, cam, mq, mt, want_cam=False)
    t4 = time.perf_counter()
    if k >= 5:
        G["scan"].append(r["report"].ms_gpu); G["odom"].append(orep.ms_gpu); G["map"].append(mrep.ms_gpu if hasattr(mrep, "ms_gpu") else float("nan"))
        T["scan"].append(t1 - t0); T["odom"].append(t2 - t1); T["map"].append(t3 - t2); T["color"].append(t4 - t3)
ctx2 = api.Context(device=0)
F = []
for k, raw in enumerate(raws):
    t0 = time.perf_counter()
    ctx2.sweep_step(raw)
    if k >= 5:
        F.append(time.perf_counter() - t0)
print(f"fused  {1e3 * float(np.mean(F)):8.3f} ms per sweep (min {1e3 * min(F):.3f}) -> {1.0 / float(np.mean(F)):.0f} sweeps/s for one sequence (lmono_sweep_step)")
tot = 0.0
for k, v in T.items():
    ms = 1e3 * float(np.mean(v))
    tot += ms if k != "color" else 0.0
    print(f"{k:6s} {ms:8.3f} ms per sweep (min {1e3 * min(v):.3f})" + (f"   device {float(np.mean(G[k])):.3f} ms" if k in G else ""))
print(f"scan+odom+map {tot:.3f} ms per sweep -> {1e3 / tot:.0f} sweeps/s for one sequence, points per sweep {len(raws[0])}, "
      f"features sharp/less_sharp/flat/less_flat = {len(r['sharp'])}/{len(r['less_sharp'])}/{len(r['flat'])}/{len(r['less_flat'])}")
