#!/usr/bin/env python
"""Whole-GPU counters of the BATCHED step (S independent sequences, one CUDA graph with S parallel branches).

ncu's default kernel replay serialises kernels, so it cannot say what bounds a batch whose kernels overlap; range
replay measures everything between cudaProfilerStart/Stop as one region with the kernels running concurrently:
  ncu --replay-mode app-range --metrics <list> python profiles/batch_range.py [S] [steps]
Without ncu the script just prints the CUDA-event time of the same region."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from lmono_b200 import api

S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
_, cm, sm, sweeps = bench.make_workload(0, n_sweeps=8)
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
ctxs = []
for s_ in range(S):
    c_ = api.Context(device=0, stream=st.cuda_stream)
    c_.map_import(0, cm); c_.map_import(1, sm); c_.sync()
    ctxs.append(c_)
batch = api.SequenceBatch(ctxs)
d = [(torch.from_numpy(c).to(dev), torch.from_numpy(s).to(dev)) for (c, s, *_r) in sweeps]
nsw = len(sweeps)
ident = ([0, 0, 0, 1], [0, 0, 0])
bargs = []
for i in range(nsw):
    ks = [(i + 3 * s_) % nsw for s_ in range(S)]
    a = api.BatchArgs(S)
    a.set_odom([(sweeps[k][4], sweeps[k][5]) for k in ks]).set_wmap_in([ident] * S)
    a.set_device_inputs([d[k][0].data_ptr() for k in ks], [d[k][0].shape[0] for k in ks],
                        [d[k][1].data_ptr() for k in ks], [d[k][1].shape[0] for k in ks])
    bargs.append(a)
for i in range(4):
    batch.step_device(join_stream=st.cuda_stream, args=bargs[i % nsw])
batch.collect()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record(st)
for i in range(steps):
    batch.step_device(join_stream=st.cuda_stream, args=bargs[(4 + i) % nsw])
e1.record(st)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
ms = e0.elapsed_time(e1)
print(f"S={S} steps={steps}: {ms:.3f} ms -> {1e3 * S * steps / ms:.0f} registrations/s, {1e3 * ms / (S * steps):.1f} us per registration")
batch.close()
