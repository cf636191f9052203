#!/usr/bin/env python
"""Timeline of a BATCHED step: which kernels of the S concurrent sequences stretch under contention.

LMONO_TIMELINE=1 makes every launch of a step be followed by a one-thread %globaltimer stamp kernel (inside the batch
graph); this script runs S sequences alone (S = 1) and together, and prints per launch site the mean duration (stamp
to stamp) alone vs. in the batch.   LMONO_TIMELINE=1 python profiles/batch_timeline.py [S] [steps]"""
import ctypes as C
import os
import re
import sys
from collections import defaultdict

os.environ.setdefault("LMONO_TIMELINE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from lmono_b200 import api

S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
_, cm, sm, sweeps = bench.make_workload(0, n_sweeps=8)
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
L = api.lib()
L.lmono_timeline_dump.argtypes = [C.c_void_p, C.c_char_p, C.c_int32]
d = [(torch.from_numpy(c).to(dev), torch.from_numpy(s).to(dev)) for (c, s, *_r) in sweeps]
nsw = len(sweeps)
ident = ([0, 0, 0, 1], [0, 0, 0])
SRC = {}
WARM = int(os.environ.get("LMONO_TIMELINE_WARM", "64"))


def site_name(site):
    f, l = site.split(":")
    if f == "begin":
        return "begin"
    if f not in SRC:
        SRC[f] = open(os.path.join(ROOT, "lmono_b200", "csrc", f)).read().splitlines()
    src = SRC[f]
    for k in range(int(l) - 1, max(int(l) - 8, -1), -1):
        m = re.search(r"(k_\w+)(?:<\w+>)?\s*(?:<<<|,)", src[k]) if k < len(src) else None
        if m:
            return "k_assoc_knn" if m.group(1) == "k_assoc_knn1" else m.group(1)     # latency / throughput form share a row
    return site


def run(n):
    ctxs = []
    for s_ in range(n):
        c_ = api.Context(device=0, stream=st.cuda_stream)
        c_.map_import(0, cm); c_.map_import(1, sm); c_.sync()
        ctxs.append(c_)
    batch = api.SequenceBatch(ctxs)
    bargs = []
    for i in range(nsw):
        ks = [(i + 3 * s_) % nsw for s_ in range(n)]
        a = api.BatchArgs(n)
        a.set_odom([(sweeps[k][4], sweeps[k][5]) for k in ks]).set_wmap_in([ident] * n)
        a.set_device_inputs([d[k][0].data_ptr() for k in ks], [d[k][0].shape[0] for k in ks],
                            [d[k][1].data_ptr() for k in ks], [d[k][1].shape[0] for k in ks])
        bargs.append(a)
    # warm-up long enough for the maps to stop gaining voxels (the same sweeps are registered again and again): the timeline
    # then shows the steady state the bench measures (refilter in place); LMONO_TIMELINE_WARM=4 shows the first steps after
    # an import instead, with the merge path of the refilter at work
    for i in range(WARM):
        batch.step_device(join_stream=st.cuda_stream, args=bargs[i % nsw])
    batch.collect()
    dur = defaultdict(list)       # (ordinal, kernel) -> [us]
    spans = []
    for i in range(steps):
        batch.step_device(join_stream=st.cuda_stream, args=bargs[(WARM + i) % nsw])
        torch.cuda.synchronize()
        t0s, t1s = [], []
        for c_ in ctxs:
            buf = C.create_string_buffer(1 << 14)
            L.lmono_timeline_dump(c_._h, buf, len(buf))
            rows = [l.split() for l in buf.value.decode().splitlines()]
            ts = [int(r[1]) for r in rows]
            t0s.append(ts[0]); t1s.append(ts[-1])
            for k in range(1, len(rows)):
                dur[(k, site_name(rows[k][0]))].append((ts[k] - ts[k - 1]) * 1e-3)
        spans.append((max(t1s) - min(t0s)) * 1e-3)
    batch.collect()
    batch.close()
    return {k: sum(v) / len(v) for k, v in dur.items()}, sum(spans) / len(spans)


alone, span1 = run(1)
multi, spanS = run(S)
print(f"step span: alone {span1:.1f} us, S={S}: {spanS:.1f} us ({spanS / S:.1f} us per registration)")
print(f"{'#':>3s} {'kernel':24s} {'alone us':>9s} {'S=' + str(S) + ' us':>9s} {'x':>6s}")
ta = tm = 0.0
for k in sorted(alone):
    a, m = alone[k], multi.get(k, float('nan'))
    ta += a; tm += m
    print(f"{k[0]:3d} {k[1]:24s} {a:9.2f} {m:9.2f} {m / a:6.2f}")
print(f"    {'sum':24s} {ta:9.2f} {tm:9.2f} {tm / ta:6.2f}")
