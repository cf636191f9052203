#!/usr/bin/env python
"""Two scan-to-map registrations of the bench workload (C-3) between cudaProfilerStart/Stop, first in the latency
forms of the kernels, then in the throughput forms (lmono_set_concurrency_hint) -- the command the ncu captures in
profiles/ were taken from:
  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof python profiles/one_step.py
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python profiles/one_step.py
"""
import os
import sys

os.environ["LMONO_NO_GRAPH"] = "1"          # plain launches: every kernel is its own ncu result
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from lmono_b200 import api

_, cm, sm, sweeps = bench.make_workload(0, n_sweeps=8)
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
ctx = api.Context(device=0, stream=st.cuda_stream)
ctx.map_import(0, cm); ctx.map_import(1, sm)
d = [(torch.from_numpy(c).to(dev), torch.from_numpy(s).to(dev)) for (c, s, *_r) in sweeps]


def step(i):
    c, s = d[i % len(d)]
    ctx.map_set_state([0, 0, 0, 1], [0, 0, 0])
    ctx.map_step_device(c.data_ptr(), c.shape[0], s.data_ptr(), s.shape[0], sweeps[i % len(d)][4], sweeps[i % len(d)][5])


for i in range(4):
    step(i)
ctx.map_collect()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step(4)
ctx.map_collect()
ctx.set_concurrency_hint(8)
step(5)
ctx.map_collect()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
q, t, rep = ctx.map_collect()
print("profiled 2 registrations; factors", list(rep.corner_num), list(rep.surf_num))
ctx.close()
