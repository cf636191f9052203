#!/usr/bin/env python
"""One-minute performance snapshot on the bench workload (C-3), for iterating on kernels:
  * latency of one registration alone on the GPU (graph replay, L2 flushed), registrations/s of the 8-sequence batch;
  * CUDA-event time per kernel of the un-graphed step, latency and throughput forms (lmono_kmarks_*);
  * %globaltimer stamps inside the last LM solve (evaluation / reduce / cluster barrier / controller per pass);
  * device ms of the three stages of a fused sweep on raw HDL-64 sweeps.
Usage: python profiles/quick.py [tag]   ->  gpurun_out/quick_<tag>.txt"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import bench
from lmono_b200 import api, synth

tag = sys.argv[1] if len(sys.argv) > 1 else "x"
parts = set(sys.argv[2].split(",")) if len(sys.argv) > 2 else {"single", "batch", "marks", "lm", "sweep"}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", f"quick_{tag}.txt"), "w")


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s); out.write(s + "\n"); out.flush()


_, cm, sm, sweeps = bench.make_workload(0, n_sweeps=12)
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
d = [(torch.from_numpy(c).to(dev), torch.from_numpy(s).to(dev)) for (c, s, *_r) in sweeps]
ident = ([0, 0, 0, 1], [0, 0, 0])
S = 8
ctxs = []
for s_ in range(S if "batch" in parts else 1):
    c_ = api.Context(device=0, stream=st.cuda_stream)
    c_.map_import(0, cm); c_.map_import(1, sm)
    ctxs.append(c_)
ctx = ctxs[0]


def step(i):
    k = i % len(d)
    ctx.map_set_state(*ident)
    ctx.map_step_device(d[k][0].data_ptr(), d[k][0].shape[0], d[k][1].data_ptr(), d[k][1].shape[0], sweeps[k][4], sweeps[k][5])


if "single" in parts:
    for i in range(5):
        step(i)
    ctx.map_collect()
    N = 100
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(N)]; e1 = [torch.cuda.Event(enable_timing=True) for _ in range(N)]
    for i in range(N):
        flush.fill_(i & 255); e0[i].record(st); step(5 + i); e1[i].record(st)
    torch.cuda.synchronize()
    ms = np.array([a.elapsed_time(b) for a, b in zip(e0, e1)])
    q, t, rep = ctx.map_collect()
    P(f"single sequence: {1e3 * ms.mean():.1f} us per registration (median {1e3 * np.median(ms):.1f}), launches/step {ctx.launch_count()}")
    if hasattr(ctx.L, "lmono_debug_rf_meta"):
        mo = (C.c_int32 * (2 * 75 * 8))()
        ctx.L.lmono_debug_rf_meta(ctx._h, mo, 2 * 75 * 8)
        m_ = np.array(mo[:], dtype=np.int64).reshape(150, 8)
        act = m_[m_[:, 0] >= 1]
        newv = act[:, 1] - act[:, 2]
        P(f"   refilter of the last step: {len(act)} cubes with new points: {int((act[:, 0] == 1).sum())} merged (new voxels), {int((act[:, 0] == 2).sum())} updated in place, "
          f"{int((act[:, 0] == 3).sum())} in place + index rebuild next step; {int(act[:, 2].sum())} stored points in them, {int(act[:, 3].sum())} new points, {int(newv.sum())} new voxels")

if "batch" in parts:
    batch = api.SequenceBatch(ctxs)
    nsw = len(sweeps)
    bargs = []
    for i in range(nsw):
        ks = [(i + 3 * s_) % nsw for s_ in range(S)]
        a_ = api.BatchArgs(S)
        a_.set_odom([(sweeps[k][4], sweeps[k][5]) for k in ks]).set_wmap_in([ident] * S)
        a_.set_device_inputs([d[k][0].data_ptr() for k in ks], [d[k][0].shape[0] for k in ks], [d[k][1].data_ptr() for k in ks], [d[k][1].shape[0] for k in ks])
        bargs.append(a_)
    for i in range(5):
        batch.step_device(join_stream=st.cuda_stream, args=bargs[i % nsw])
    batch.collect()
    N = 100
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(N)]; e1 = [torch.cuda.Event(enable_timing=True) for _ in range(N)]
    for i in range(N):
        flush.fill_(i & 255); e0[i].record(st); batch.step_device(join_stream=st.cuda_stream, args=bargs[(5 + i) % nsw]); e1[i].record(st)
    torch.cuda.synchronize()
    ms = np.array([a.elapsed_time(b) for a, b in zip(e0, e1)])
    res = batch.collect()
    err = max(float(np.linalg.norm(r[1] - sweeps[((5 + N - 1) + 3 * s_) % nsw][3])) for s_, r in enumerate(res))
    P(f"batch of {S}: {1e3 * ms.mean():.1f} us per step = {S / ms.mean() * 1e3:.0f} registrations/s (worst registration error {err:.4f} m)")

if "marks" in parts:
    for form, hint in (("latency", 1), ("throughput", 8)):
        ctx.set_concurrency_hint(hint)
        for i in range(3):
            step(i)
        ctx.map_collect()
        ctx.kernel_marks_enable(True)
        n = 24
        for i in range(n):
            flush.fill_(1); torch.cuda._sleep(4_000_000); step(5 + i)
        m = ctx.kernel_marks(); ctx.kernel_marks_enable(False)
        m.pop("k_set_wmap", None)
        tot = sum(v[1] for v in m.values())
        P(f"--- per-kernel CUDA-event times, {form} forms: sum {1e3 * tot / n:.1f} us/step over {sum(v[0] for v in m.values()) / n:.0f} launches")
        for k, v in sorted(m.items(), key=lambda kv: -kv[1][1]):
            P(f"   {k:24s} {v[0] / n:4.1f} x {1e3 * v[1] / v[0]:7.2f} us = {1e3 * v[1] / n:7.2f} us/step")
    ctx.set_concurrency_hint(1)

if "lm" in parts:
    for i in range(3):
        step(i)
    ctx.map_collect()
    o = (C.c_uint64 * 64)()
    ctx.L.lmono_debug_stamps(ctx._h, o, 64)
    a = np.array(o[:], dtype=np.int64)
    P(f"--- LM solve kernel (last solve): total {a[2] - a[0]} ns, arm {a[1] - a[0]} ns")
    for e in range(5):
        b = a[8 + 8 * e:12 + 8 * e]
        prev = a[1] if e == 0 else a[11 + 8 * (e - 1)]
        if b[0] > 0 and b[0] >= prev:
            P(f"   pass {e}: eval {b[0] - prev} reduce {b[1] - b[0]} cluster {b[2] - b[1]} control {b[3] - b[2]} ns")

if "sweep" in parts:
    wld = synth.make_world()
    rng = np.random.default_rng(2)
    raws = [np.ascontiguousarray(synth.raycast_sweep_torch(wld, *synth.loop_pose(wld, 1.0 * k), 64, 1875, rng, device=dev), np.float32) for k in range(16)]
    pctx = api.Context(device=0, stream=st.cuda_stream)
    ms = []
    wall = []
    for k, raw in enumerate(raws):
        t0 = time.perf_counter()
        o = pctx.sweep_step(raw)
        wall.append(time.perf_counter() - t0)
        ms.append((o[3].ms_gpu, o[4].ms_gpu, o[5].ms_gpu))
    P("   mapping ms per sweep:", " ".join(f"{m_[2]:.3f}" for m_ in ms))
    ms = np.array(ms[4:])
    P(f"--- fused sweep (map grown from empty): scanRegistration {1e3 * ms[:, 0].mean():.0f} us, odometry {1e3 * ms[:, 1].mean():.0f} us, "
      f"mapping {1e3 * ms[:, 2].mean():.0f} us device; wall {1e3 * np.mean(wall[4:]):.3f} ms per sweep")
    o = (C.c_uint64 * 256)()
    pctx.L.lmono_debug_stamps(pctx._h, o, 256)
    a = np.array(o[200:208], dtype=np.int64)
    P("   k_scan_ring (ring 32) ns: load %d, sector sorts %d, greedy pick %d, labels + less-flat list %d, voxel bounds %d, voxel sort %d, centroids %d; total %d"
      % tuple(list(np.diff(a)) + [a[7] - a[0]]))
    pctx.kernel_marks_enable(True)
    for raw in raws[:8]:
        pctx.sweep_step(raw)
    m = pctx.kernel_marks(); pctx.kernel_marks_enable(False)
    for k, v in sorted(m.items(), key=lambda kv: -kv[1][1]):
        P(f"   {k:24s} {v[0] / 8:4.1f} x {1e3 * v[1] / v[0]:7.2f} us = {1e3 * v[1] / 8:7.2f} us/sweep")
    pctx.close()
for c_ in ctxs:
    c_.close()
