#!/usr/bin/env python
"""One pass over every stage between cudaProfilerStart/Stop, plain launches (every kernel its own ncu result):
  * two scan-to-map registrations of the bench workload C-3 (latency forms, then throughput forms),
  * one fused HDL-64 sweep (scanRegistration -> laserOdometry -> laserMapping) against the imported C-3 map,
  * one HDL-32 scan-to-scan odometry step (C-2),
  * one colour frame (120 k points -> 1241 x 376).
The command the round-2 ncu captures under profiles/ were taken from:
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python profiles/one_sweep.py
  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof python profiles/one_sweep.py
"""
import os
import sys

os.environ["LMONO_NO_GRAPH"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import bench
from lmono_b200 import api, synth

city, cm, sm, sweeps = bench.make_workload(0, n_sweeps=8)
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
ctx = api.Context(device=0, stream=st.cuda_stream, cube_capacity_corner=32768, cube_capacity_surf=65536)
ctx.map_import(0, cm); ctx.map_import(1, sm)
d = [(torch.from_numpy(c).to(dev), torch.from_numpy(s).to(dev)) for (c, s, *_r) in sweeps]


def step(i):
    c, s = d[i % len(d)]
    ctx.map_set_state([0, 0, 0, 1], [0, 0, 0])
    ctx.map_step_device(c.data_ptr(), c.shape[0], s.data_ptr(), s.shape[0], sweeps[i % len(d)][4], sweeps[i % len(d)][5])


rng = np.random.default_rng(2)
poses = [synth.city_pose(city, 0.5 * k) for k in range(6)]
raws = [np.ascontiguousarray(synth.raycast_sweep_torch(city, q, t, 64, 1875, rng, device=dev), np.float32) for (q, t) in poses]
wld = synth.make_world(seed=20261018)
raws32 = [np.ascontiguousarray(synth.raycast_sweep_torch(wld, *synth.loop_pose(wld, 1.0 * k), 32, 1875, rng, device=dev), np.float32) for k in range(4)]
c32 = api.Context(device=0, stream=st.cuda_stream, scan_line=32, minimum_range=0.3, max_cubes_corner=8, max_cubes_surf=8,
                  cube_capacity_corner=1024, cube_capacity_surf=1024)
f32 = [c32.scan_register(r) for r in raws32]
cam = api.Pinhole(718.856, 718.856, 607.1928, 185.2157, 0.0, 0.0, 0.0, 0.0, 1241, 376, 0, 5, 0)
T = np.array([[0.0, -1.0, 0.0, 0.0], [0.0, 0.0, -1.0, -0.08], [1.0, 0.0, 0.0, -0.27]])
img = rng.integers(0, 256, (376, 1241, 3), dtype=np.uint8)

for i in range(4):
    step(i)
ctx.map_collect()
ctx.map_set_state(*poses[0])
for k in range(5):
    ctx.sweep_step(raws[k])
for k in range(3):
    c32.odom_step(f32[k]["sharp"], f32[k]["less_sharp"], f32[k]["flat"], f32[k]["less_flat"])
ctx.project_color(raws[0], img, cam, [0, 0, 0, 1], [0, 0, 0], T_cam_lidar=T)
torch.cuda.synchronize()
torch.cuda.profiler.start()
step(4)
ctx.map_collect()
ctx.set_concurrency_hint(8)
step(5)
ctx.map_collect()
ctx.set_concurrency_hint(1)
ctx.sweep_step(raws[5])
c32.odom_step(f32[3]["sharp"], f32[3]["less_sharp"], f32[3]["flat"], f32[3]["less_flat"])
ctx.project_color(raws[0], img, cam, [0, 0, 0, 1], [0, 0, 0], T_cam_lidar=T)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled: 2 registrations, 1 fused sweep, 1 HDL-32 odometry step, 1 colour frame")
ctx.close(); c32.close()
