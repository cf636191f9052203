#!/usr/bin/env python
"""Throughput of S sequences on one GPU driven as G independent groups (one SequenceBatch + stream per group) instead
of one lockstep batch: the groups drift out of phase, so the narrow phases of one (LM solve, sorts) overlap the wide
phases of another (kNN, refilter).   python profiles/batch_groups.py S G [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from lmono_b200 import api

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
G = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 100
_, cm, sm, sweeps = bench.make_workload(0, n_sweeps=8)
dev = torch.device("cuda", 0)
main = torch.cuda.Stream(device=dev); torch.cuda.set_stream(main)
d = [(torch.from_numpy(c).to(dev), torch.from_numpy(s).to(dev)) for (c, s, *_r) in sweeps]
nsw = len(sweeps)
ident = ([0, 0, 0, 1], [0, 0, 0])
groups = []
per = S // G
for g in range(G):
    st = torch.cuda.Stream(device=dev)
    ctxs = []
    for s_ in range(per):
        c_ = api.Context(device=0, stream=st.cuda_stream)
        c_.set_concurrency_hint(S)
        c_.map_import(0, cm); c_.map_import(1, sm); c_.sync()
        ctxs.append(c_)
    b = api.SequenceBatch(ctxs)
    bargs = []
    for i in range(nsw):
        ks = [(i + 3 * (g * per + s_)) % nsw for s_ in range(per)]
        a = api.BatchArgs(per)
        a.set_odom([(sweeps[k][4], sweeps[k][5]) for k in ks]).set_wmap_in([ident] * per)
        a.set_device_inputs([d[k][0].data_ptr() for k in ks], [d[k][0].shape[0] for k in ks],
                            [d[k][1].data_ptr() for k in ks], [d[k][1].shape[0] for k in ks])
        bargs.append(a)
    groups.append((st, b, bargs))
for i in range(5):
    for st, b, bargs in groups:
        b.step_device(join_stream=st.cuda_stream, args=bargs[i % nsw])
for st, b, bargs in groups:
    b.collect()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(main)
for st, b, bargs in groups:
    st.wait_stream(main)
t0 = time.perf_counter()
for i in range(steps):
    for st, b, bargs in groups:
        b.step_device(join_stream=st.cuda_stream, args=bargs[(5 + i) % nsw])
host = time.perf_counter() - t0
for st, b, bargs in groups:
    main.wait_stream(st)
e1.record(main)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
res = [b.collect() for st, b, bargs in groups]
print(f"S={S} G={G} steps={steps}: {ms:.2f} ms -> {1e3 * S * steps / ms:.0f} registrations/s ({1e3 * ms / steps:.1f} us per round of {S}), host enqueue {1e3 * host / steps:.3f} ms/round")
