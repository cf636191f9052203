#!/usr/bin/env python
"""Host cost of the asynchronous fused sweep: seconds spent inside lmono_sweep_submit (enqueue of ~45 launches) per
sequence, versus the time until all sweeps of a batch step have completed."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from lmono_b200 import api, synth
S, NSW = 8, 10
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
seqs = []
for s_ in range(S):
    w = synth.make_world(seed=100 + s_); rng = np.random.default_rng(s_)
    seqs.append([torch.from_numpy(np.ascontiguousarray(synth.raycast_sweep_torch(w, *synth.loop_pose(w, 1.0 * k), 64, 1875, rng, device=dev), np.float32)).pin_memory() for k in range(NSW)])
ctxs = [api.Context(device=0, stream=st.cuda_stream) for _ in range(S)]
sub, tot = [], []
for k in range(NSW):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s_ in range(S):
        ctxs[s_].sweep_submit(seqs[s_][k].numpy())
    t1 = time.perf_counter()
    for s_ in range(S):
        ctxs[s_].sweep_wait()
    t2 = time.perf_counter()
    if k >= 3:
        sub.append((t1 - t0) / S); tot.append(t2 - t0)
print(f"submit: {1e6 * np.mean(sub):.0f} us per sweep on the host; batch step of {S}: {1e3 * np.mean(tot):.3f} ms ({S / np.mean(tot):.0f} sweeps/s); "
      f"host enqueue share {100 * np.mean(sub) * S / np.mean(tot):.0f} %")
