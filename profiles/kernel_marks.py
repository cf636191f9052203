#!/usr/bin/env python
"""In-situ per-kernel device times of one scan-to-map registration on the bench workload (C-3).

Every kernel launch of the step is followed by a CUDA event (lmono_kmarks_*); a GPU-side sleep in front of
each step lets the host enqueue the whole step before the device starts it, so the deltas are device
durations in the real kernel order and cache state (L2 flushed before each step like bench.py), not host
launch latency.  Usage: python profiles/kernel_marks.py [steps]"""
import ctypes as C
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from lmono_b200 import api

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 24
_, cm, sm, sweeps = bench.make_workload(0, n_sweeps=8)
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
ctx = api.Context(device=0, stream=st.cuda_stream)
ctx.map_import(0, cm); ctx.map_import(1, sm)
d = [(torch.from_numpy(c).to(dev), torch.from_numpy(s).to(dev)) for (c, s, *_r) in sweeps]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def step(i):
    c, s = d[i % len(d)]
    ctx.map_set_state([0, 0, 0, 1], [0, 0, 0])
    ctx.map_step_device(c.data_ptr(), c.shape[0], s.data_ptr(), s.shape[0], sweeps[i % len(d)][4], sweeps[i % len(d)][5])

for i in range(5):
    step(i)
ctx.map_collect()
ctx._chk(ctx.L.lmono_kmarks_enable(ctx._h, 1), "kmarks")
for i in range(steps):
    flush.fill_(i & 255)
    torch.cuda._sleep(4_000_000)          # ~2 ms: the host gets ahead of the device
    step(5 + i)
buf = C.create_string_buffer(1 << 16)
ctx._chk(ctx.L.lmono_kmarks_dump(ctx._h, buf, len(buf)), "dump")
ctx.map_collect()
rows = []
for line in buf.value.decode().splitlines():
    site, n, ms = line.split()
    f, l = site.split(":")
    src = open(os.path.join(ROOT, "lmono_b200", "csrc", f)).read().splitlines()
    name = "?"
    for k in range(int(l) - 1, max(int(l) - 8, -1), -1):
        m = re.search(r"(k_\w+)(?:<\w+>)?\s*(?:<<<|,)", src[k]) if k < len(src) else None
        if m:
            name = m.group(1); break
    rows.append((name, site, int(n), float(ms)))
tot = sum(r[3] for r in rows)
print(f"{'kernel':26s} {'site':18s} {'n/step':>6s} {'us/launch':>10s} {'us/step':>9s} {'share':>6s}")
for name, site, n, ms in sorted(rows, key=lambda r: -r[3]):
    print(f"{name:26s} {site:18s} {n / steps:6.1f} {1e3 * ms / n:10.2f} {1e3 * ms / steps:9.2f} {100 * ms / tot:5.1f}%")
print(f"sum of kernel deltas {1e3 * tot / steps:.1f} us/step over {sum(r[2] for r in rows) / steps:.0f} launches/step")
