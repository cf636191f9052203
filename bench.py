#!/usr/bin/env python
"""bench.py -- scan-to-map registrations/s on the C-3 workload of SURVEY.md section 8(d):
an HDL-64-sized feature sweep (~4 k corner + ~12 k surf queries after the 0.4 / 0.8 m voxel
filter) registered against a ~1 M-point local map in the reference's 21x21x11 cube structure,
one full pass of laserMapping's process() per step (window upkeep, VoxelGrid of the features,
2 x (5-NN association + line/plane fits + LM solve), map insertion and cube refilter).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Our arm: `value` = device-resident inputs, CUDA-event timed, L2 flushed between steps;
`e2e` = the same registrations through the public host API (lmono_map_step) with pinned host
buffers, H2D of the features and D2H of the pose inside the timed region.  The reference arm
times the CPU oracle (a restatement of the reference's PCL/FLANN/Ceres path; the reference
itself cannot be built here) on the host cores.  Multi-GPU: one independent sequence per rank
(weak scaling, no data-path collective).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from lmono_b200 import synth  # noqa: E402

METRIC = "scan-to-map registrations/s"
UNIT = "registrations/s"
WORKLOAD = "C-3 HDL-64 scan-to-map: ~4k corner + ~12k surf queries vs ~1M-pt 21x21x11 cube local map (0.4/0.8 m voxels), incl. map update"
N_DISTINCT_SWEEPS = 24
RAW_CORNER, RAW_SURF = 6800, 16000


# ----------------------------------------------------------------------------- workload
def voxel_dedupe(pts, leaf):
    """keep the first sample of every global voxel (host-side thinning of the raw samples so
    that a map import stays below 2^21 points; the import itself still runs the VoxelGrid)."""
    inv = np.float32(1.0) / np.float32(leaf)
    ijk = np.floor(pts[:, :3] * inv).astype(np.int64)
    key = (ijk[:, 2] * 2_000_003 + ijk[:, 1]) * 2_000_003 + ijk[:, 0]
    _, first = np.unique(key, return_index=True)
    return pts[np.sort(first)]


def make_workload(rank=0, n_sweeps=N_DISTINCT_SWEEPS):
    city = synth.make_city(seed=7, pole_pitch=3.7, street_radius=18.0)   # sensor stays inside the centre cube
    c = (0.0, 0.0, 0.0)
    cm, sm = synth.sample_map(city, c, half_xy=125.0, n_surf=7_000_000, n_corner=2_500_000, seed=7)
    rng = np.random.default_rng(70)
    roofs = synth.to_xyzi(synth.sample_box_roofs(city, c, 180.0, 1_500_000, rng) + rng.normal(0, 0.01, (1_500_000, 3)))
    roofs = roofs[(np.abs(roofs[:, 0]) < 125.0) & (np.abs(roofs[:, 1]) < 125.0)]
    sm = np.concatenate([sm, roofs])
    cm = voxel_dedupe(cm, 0.4)
    sm = voxel_dedupe(sm, 0.8)
    rng = np.random.default_rng(11 + 1000 * rank)
    sweeps = []
    for k in range(n_sweeps):
        q, t = synth.city_pose(city, 1.0 * k + 7.0 * rank)
        co, su = synth.sample_sweep_features(city, q, t, rng, RAW_CORNER, RAW_SURF, max_range=60.0)
        qp, tp = synth.perturb_pose(q, t, rng, 0.2, 1.0)        # U(+-0.2 m, +-1 deg), SURVEY 8d C-3
        sweeps.append((co, su, q, t, qp, tp))
    return city, cm, sm, sweeps


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- reference arm
def _oracle_worker(args):
    rank, steps, warmup = args
    import oracle_lib as O
    _, cm, sm, sweeps = make_workload(rank)
    m = O.Mapper(0.4, 0.8, 0, 1)
    m.import_points(0, cm)
    m.import_points(1, sm)
    phases = {"ms_tree": 0.0, "ms_assoc": 0.0, "ms_solver": 0.0, "ms_filter": 0.0, "ms_add": 0.0, "ms_shift": 0.0}
    for i in range(warmup):
        c, s, q, t, qp, tp = sweeps[i % len(sweeps)]
        m.set_state([0, 0, 0, 1], [0, 0, 0])
        m.step(c, s, qp, tp)
    t0 = time.perf_counter()
    for i in range(steps):
        c, s, q, t, qp, tp = sweeps[(warmup + i) % len(sweeps)]
        m.set_state([0, 0, 0, 1], [0, 0, 0])
        _, _, rep, _ = m.step(c, s, qp, tp)
        for k in phases:
            phases[k] += getattr(rep, k)
    dt = time.perf_counter() - t0
    nmap = len(m.export(0)) + len(m.export(1))
    return dt, phases, nmap


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 16))
    with mp.get_context("spawn").Pool(procs) as pool:
        res = pool.map(_oracle_worker, [(r, args.steps, args.warmup) for r in range(procs)])
    wall = max(r[0] for r in res)
    value = procs * args.steps / wall
    ph = {k: sum(r[1][k] for r in res) / (procs * args.steps) for k in res[0][1]}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps / procs * procs, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "map_points": res[0][2], "parallelism": f"{procs} independent sequences on {procs} host processes"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port",
                         "sample": f"{args.steps} registrations per process x {procs} processes (oracle restatement of the PCL/FLANN/Ceres path: "
                                   "per-sweep KD-tree rebuild, 5-NN, fits, Ceres-style LM, per-cube VoxelGrid refilter)",
                         "ms_per_registration_phases": {k: round(v, 3) for k, v in ph.items()}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from lmono_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    S = max(1, args.sequences)
    _, cm, sm, sweeps = make_workload(rank)
    nsw = len(sweeps)
    # one real (non-default) stream for torch, NCCL, the CUDA events and every sequence ctx: a batch step is ONE CUDA
    # graph launched on it, whose S parallel branches (one registration per sequence) fork and join inside the graph.
    main = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(main)
    assert main.cuda_stream != 0
    seq_streams = [main] * S
    ctxs = []
    for s_ in range(S):
        c_ = api.Context(device=local, stream=seq_streams[s_].cuda_stream)
        c_.map_import(0, cm)
        c_.map_import(1, sm)
        c_.sync()
        ctxs.append(c_)
    ctx = ctxs[0]
    batch = api.SequenceBatch(ctxs)

    # device-resident copies of the sweeps (value leg) and pinned host copies (e2e leg)
    d_sweeps = [(torch.from_numpy(c).to(dev), torch.from_numpy(s).to(dev)) for (c, s, *_rest) in sweeps]
    h_sweeps = [(torch.from_numpy(c).pin_memory(), torch.from_numpy(s).pin_memory()) for (c, s, *_rest) in sweeps]
    h_np = [(a.numpy(), b.numpy()) for a, b in h_sweeps]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    ident = ([0, 0, 0, 1], [0, 0, 0])

    # sequence s registers sweep (i + 3 s) mod 24 at batch step i; every registration starts from its own
    # U(+-0.2 m, +-1 deg) perturbation (wmap_wodom reset to identity).  One argument set per i mod 24.
    def sweep_of(i, s_):
        return (i + 3 * s_) % nsw

    bargs = []
    for i in range(nsw):
        ks = [sweep_of(i, s_) for s_ in range(S)]
        a_ = api.BatchArgs(S)
        a_.set_odom([(sweeps[k][4], sweeps[k][5]) for k in ks]).set_wmap_in([ident] * S)
        a_.set_device_inputs([d_sweeps[k][0].data_ptr() for k in ks], [d_sweeps[k][0].shape[0] for k in ks],
                             [d_sweeps[k][1].data_ptr() for k in ks], [d_sweeps[k][1].shape[0] for k in ks])
        a_.set_host_inputs([h_np[k][0] for k in ks], [h_np[k][1] for k in ks])
        bargs.append(a_)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t_ = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return float(t_.item())

    def check_converged(res, i_last):
        worst = 0.0
        for s_, (q_, t_, rep_) in enumerate(res):
            tg = sweeps[sweep_of(i_last, s_)][3]
            err = float(np.linalg.norm(t_ - tg))
            assert rep_.optimized == 1 and err < 0.1, (s_, rep_.optimized, err)
            worst = max(worst, err)
        return worst

    W = max(args.warmup, 3)
    for i in range(W):
        batch.step_device(join_stream=main.cuda_stream, args=bargs[i % nsw])
    batch.collect()

    # ---- value leg: S sequences per GPU, device-resident inputs.  Per batch step: L2 flush on the main stream, event,
    # fork -> one registration per sequence, overlapping on the device -> join, event.  value = registrations / sum of the
    # event-bracketed times (the flush is outside the brackets).
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    n_launch0 = sum(c_.launch_count() for c_ in ctxs)
    t_host0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        ev0[i].record(main)
        batch.step_device(join_stream=main.cuda_stream, args=bargs[(W + i) % nsw])
        ev1[i].record(main)
    host_enqueue_s = time.perf_counter() - t_host0
    barrier()
    n_launch = sum(c_.launch_count() for c_ in ctxs) - n_launch0
    clocks = sampler.stop()
    res = batch.collect()
    step_ms = [a.elapsed_time(b) for a, b in zip(ev0, ev1)]
    total_ms_max = allmax(float(sum(step_ms)))
    value = world * S * args.steps / (total_ms_max * 1e-3)
    reg_err = check_converged(res, W + args.steps - 1)     # the work inside the timed region converged

    # ---- e2e leg: public host API, page-locked host inputs.  Every step moves the features of all S sequences host ->
    # device (fetched over PCIe by the first kernel of each registration) and the pose + report of every sequence device
    # -> host, inside the timed region.  Pipelined as a streaming consumer would (lmono_map_submit_batch /
    # lmono_map_wait_batch): sweep k+1 is submitted while sweep k runs, every result is read before sweep k+2 goes in.
    for i in range(3):
        batch.step(args=bargs[i % nsw])
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    h2d = 0
    e0.record(main)
    t0 = time.perf_counter()
    batch.submit(args=bargs[W % nsw])
    for i in range(args.steps):
        a_ = bargs[(W + i) % nsw]
        if i + 1 < args.steps:
            batch.submit(args=bargs[(W + i + 1) % nsw])
        res = batch.wait()
        h2d += sum(a_.cv[s_].n + a_.sv[s_].n for s_ in range(S)) * 16
    e1.record(main)
    barrier()
    e2e_wall = time.perf_counter() - t0
    check_converged(res, W + args.steps - 1)
    e2e_ms = allmax(max(e0.elapsed_time(e1), e2e_wall * 1e3))
    e2e_value = world * S * args.steps / (e2e_ms * 1e-3)
    d2h = int(ctx.L.lmono_map_result_bytes()) * S * args.steps      # pose + report read back per registration
    # the same without pipelining: lmono_map_step_batch (submit + wait per step)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        res = batch.step(args=bargs[(W + i) % nsw])
    barrier()
    e2e_sync_ms = allmax((time.perf_counter() - t0) * 1e3)
    e2e_sync_value = world * S * args.steps / (e2e_sync_ms * 1e-3)

    # ---- the same value leg with twice the sequences per GPU (how far the GPU is from saturated at S)
    more = None
    if args.more_sequences > S:
        S2 = args.more_sequences
        ctxs2 = list(ctxs)
        for s_ in range(S, S2):
            c_ = api.Context(device=local, stream=main.cuda_stream)
            c_.map_import(0, cm)
            c_.map_import(1, sm)
            c_.sync()
            ctxs2.append(c_)
        batch2 = api.SequenceBatch(ctxs2)
        bargs2 = []
        for i in range(nsw):
            ks = [sweep_of(i, s_) for s_ in range(S2)]
            a_ = api.BatchArgs(S2)
            a_.set_odom([(sweeps[k][4], sweeps[k][5]) for k in ks]).set_wmap_in([ident] * S2)
            a_.set_device_inputs([d_sweeps[k][0].data_ptr() for k in ks], [d_sweeps[k][0].shape[0] for k in ks],
                                 [d_sweeps[k][1].data_ptr() for k in ks], [d_sweeps[k][1].shape[0] for k in ks])
            bargs2.append(a_)
        for i in range(W):
            batch2.step_device(join_stream=main.cuda_stream, args=bargs2[i % nsw])
        batch2.collect()
        barrier()
        m0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        m1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        for i in range(args.steps):
            flush.fill_(i & 0xFF)
            m0[i].record(main)
            batch2.step_device(join_stream=main.cuda_stream, args=bargs2[(W + i) % nsw])
            m1[i].record(main)
        barrier()
        batch2.collect()
        ms2 = allmax(float(sum(a.elapsed_time(b) for a, b in zip(m0, m1))))
        more = {"sequences_per_gpu": S2, "value": world * S2 * args.steps / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / args.steps,
                "note": "same measurement as `value` with more independent sequences per GPU"}
        for c_ in ctxs2[S:]:
            c_.close()

    # ---- single sequence alone on the GPU (latency of one registration; the round-1 headline): device-resident inputs,
    # CUDA events per step, L2 flushed between steps
    def step_single(i):
        k = i % nsw
        ctx.map_set_state(*ident)
        ctx.map_step_device(d_sweeps[k][0].data_ptr(), d_sweeps[k][0].shape[0], d_sweeps[k][1].data_ptr(), d_sweeps[k][1].shape[0],
                            sweeps[k][4], sweeps[k][5])

    for i in range(3):
        step_single(i)
    barrier()
    sv0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    sv1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        sv0[i].record(main)
        step_single(W + i)
        sv1[i].record(main)
    barrier()
    single_ms = allmax(float(sum(a.elapsed_time(b) for a, b in zip(sv0, sv1)))) / args.steps
    # the same through the host API (lmono_map_step: upload, step, synchronous read-back)
    t0 = time.perf_counter()
    for i in range(args.steps):
        k = (W + i) % nsw
        ctx.map_set_state(*ident)
        ctx.map_step(h_np[k][0], h_np[k][1], sweeps[k][4], sweeps[k][5])
    single_e2e_ms = allmax((time.perf_counter() - t0) * 1e3) / args.steps

    # ---- per-kernel device times in situ: a CUDA event after every launch of the step (plain launches, same kernel
    # order and cache state as the timed leg: L2 flushed before each step); a GPU-side sleep in front of each step lets
    # the host enqueue the whole step before the device starts it, so the deltas are device durations, not launch latency
    nprof = min(args.steps, 24)

    def marks_leg():
        ctx.kernel_marks_enable(True)
        for i in range(nprof):
            flush.fill_(1)
            torch.cuda._sleep(4_000_000)
            step_single(W + i)
        m_ = ctx.kernel_marks()
        ctx.kernel_marks_enable(False)
        m_.pop("k_set_wmap", None)          # first mark of a step: its delta contains the flush + sleep
        return m_

    marks = marks_leg()                      # latency forms (a sequence alone on the GPU)
    ctx.set_concurrency_hint(max(S, 4))      # throughput forms, as the batched legs above ran them
    marks_tp = marks_leg()
    ctx.set_concurrency_hint(1)
    q_last, t_last, rep = ctx.map_collect()
    kern_us = {k: 1e3 * v[1] / nprof for k, v in marks.items()}                    # us per step
    kern_us_tp = {k: 1e3 * v[1] / nprof for k, v in marks_tp.items()}
    kern_launch_ms = {k: v[1] / max(v[0], 1) for k, v in marks.items()}            # ms per launch
    for k, v in marks_tp.items():
        kern_launch_ms.setdefault(k, v[1] / max(v[0], 1))                          # kernels only the throughput form launches
    nq = rep.corner_stack + rep.surf_stack
    nmap = rep.corner_from_map + rep.surf_from_map
    nfac = rep.corner_num[1] + rep.surf_num[1]
    evals = sum(s_.iterations + 1 for s_ in rep.solve)
    ncu = {}
    ncu_file = "ncu_r01_j_full_metrics.json"
    try:
        for l in json.load(open(os.path.join(ROOT, "profiles", ncu_file)))["launches"]:
            ncu.setdefault(l["kernel"], []).append(l)
    except Exception:
        pass

    def roof(kernel, alg_bytes, what):
        ms = kern_launch_ms.get(kernel)
        if not ms:
            return None
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        r = {"kernel": kernel, "what": what, "algorithmic_bytes_per_launch": float(alg_bytes), "avg_launch_ms": ms,
             "achieved": gbs, "frac": gbs / hbm_peak}
        if kernel in ncu:
            r["traffic"] = float(np.mean([l["dram_bytes_total"] for l in ncu[kernel]]))
            r["l2_hit_pct_ncu"] = float(np.mean([l["lts__t_sector_hit_rate.pct"] for l in ncu[kernel]]))
        return r

    nraw = RAW_CORNER + RAW_SURF
    others = [roof("k_lm_solve_cluster", 64.0 * nfac * max(evals, 1) / 2.0, "one LM solve: E evaluations x F factors x 64 B (SURVEY 8d); fp64-latency bound, factors stay in shared memory after the first pass"),
              roof("k_assoc_fit", 160.0 * nq, "line / plane fit: 5 neighbours + query + factor record per query"),
              roof("k_sort_tiles", 16.0 * nraw, "register bitonic tile sort: 8 B key read + write"),
              roof("k_merge_ranks_smem", 16.0 * nraw, "rank merge of the sorted tiles in shared memory"),
              roof("k_vg_write", 8.0 * nraw + 16.0 * nraw + 16.0 * nq, "VoxelGrid centroids of both feature clouds")]
    knn8 = roof("k_assoc_knn", 116.0 * nq, "latency form of the exact 5-NN (8 lanes per query), used when a sequence runs alone on the GPU")
    if knn8:
        others.insert(0, knn8)
    knn = roof("k_assoc_knn1", 116.0 * nq, "exact 5-NN: 16 B query + 5 x 16 B neighbours + 5 x 4 B indices per query (SURVEY 8d)") or {}
    roofline = {
        "bound": "hbm", "kernel": "k_assoc_knn1 (exact 5-NN of every feature against the cube map, throughput form = the one the batched step runs; one launch per outer iteration)",
        "achieved": knn.get("achieved"), "peak": hbm_peak, "unit": "GB/s", "peak_source": peak_src,
        "frac": knn.get("frac"), "traffic": knn.get("traffic"),
        "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full capture committed as profiles/" + ncu_file + " (cold caches per ncu replay)",
        "algorithmic_bytes_per_launch": knn.get("algorithmic_bytes_per_launch"), "avg_launch_ms": knn.get("avg_launch_ms"),
        "timing": "one sequence alone on the GPU, un-graphed step, CUDA event after every launch",
        "queries_per_launch": int(nq), "knn_queries_per_s": nq / (knn["avg_launch_ms"] * 1e-3) if knn.get("avg_launch_ms") else None,
        "knn_queries_per_s_batched": 2.0 * nq * value,
        "l2_hit_pct_ncu": knn.get("l2_hit_pct_ncu"),
        "kernel_us_per_step": {k: round(v, 2) for k, v in sorted(kern_us.items(), key=lambda kv: -kv[1])},
        "kernel_us_per_step_throughput_forms": {k: round(v, 2) for k, v in sorted(kern_us_tp.items(), key=lambda kv: -kv[1])},
        "kernel_us_note": "CUDA-event deltas between consecutive launches of the un-graphed step of ONE sequence: each includes ~3 us of event + launch gap, "
                          "so the sum exceeds ms_per_step (graph replay); profiles/ holds the ncu launch list of the same command and the "
                          "timeline of the batched step (profiles/batch_timeline_r01.txt)",
        "other_kernels": [r for r in others if r],
        "note": "the map (~16 MB + 16 MB index) fits the 126 MB L2 and one registration moves ~30-60 MB algorithmically (5-10 us of HBM time): "
                "the step is bound by dependent L2 / HBM round trips, fp64 latency and ~27 launches, not by HBM bandwidth (DESIGN.md section 4); "
                "whole-GPU counters of the batched step: profiles/batch_range_r01.csv",
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "map_points": int(nmap), "queries_per_sweep": int(nq),
                   "raw_features_per_sweep": [RAW_CORNER, RAW_SURF], "sequences_per_gpu": S,
                   "registrations_per_step": S * world,
                   "l2": "flushed between steps (256 MiB write)",
                   "parallelism": f"{S} independent sequences per GPU (one ctx each, BASELINE config C-4); one registration of every sequence per step = one CUDA graph with {S} parallel branches, no collective"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps,
                "pipeline": "lmono_map_submit_batch / lmono_map_wait_batch, two steps in flight; every result read on the host",
                "unpipelined_value": e2e_sync_value},
        "single_sequence": {"value": 1e3 / single_ms * world, "unit": UNIT, "ms_per_registration": single_ms,
                            "e2e_value": 1e3 / single_e2e_ms * world, "e2e_ms_per_registration": single_e2e_ms,
                            "note": "one sequence alone on each GPU: latency of one registration (graph replay, L2 flushed between steps)"},
        "more_sequences": more,
        "host_enqueue_ms_per_step": 1e3 * host_enqueue_s / args.steps,
        "gpu_launches": int(n_launch),
        "clocks": clocks,
        "roofline": roofline,
        "registration_error_m": reg_err,
    }

    # ---- the whole A-LOAM chain of one sequence (BASELINE config C-4 "fused L1 -> L2 -> L3"): raw 64-ring sweeps of
    # ~113 k points through lmono_sweep_step (scanRegistration -> laserOdometry -> laserMapping, host sweep in, poses out)
    if rank == 0 and not args.no_pipeline:
        try:
            wld = synth.make_world()
            rng_p = np.random.default_rng(2)
            raws = []
            for k in range(30):
                q_, t_ = synth.loop_pose(wld, 1.0 * k)
                raws.append(np.ascontiguousarray(synth.raycast_sweep(wld, q_, t_, 64, 1875, rng_p), np.float32))
            pctx = api.Context(device=local, stream=main.cuda_stream)
            tp_ = []
            for k, raw in enumerate(raws):
                t0 = time.perf_counter()
                out_ = pctx.sweep_step(raw)
                if k >= 6:
                    tp_.append(time.perf_counter() - t0)
            truth = float(np.linalg.norm(synth.loop_pose(wld, 29.0)[1] - synth.loop_pose(wld, 0.0)[1]))
            drift = float(np.linalg.norm(out_[2][1])) - truth      # mapped translation of the last sweep vs the true chord
            line["fused_sweep"] = {"sweeps_per_s": 1.0 / float(np.mean(tp_)), "ms_per_sweep": 1e3 * float(np.mean(tp_)),
                                   "points_per_sweep": int(len(raws[0])), "sequences": 1, "travelled_minus_truth_m": drift,
                                   "what": "lmono_sweep_step: raw HDL-64 sweep (pageable host memory) -> scanRegistration -> laserOdometry -> "
                                           "laserMapping -> poses on the host; one sequence, synchronous calls, wall clock"}
            pctx.close()
        except Exception as e:                                      # noqa: BLE001  (never lose the headline line to the extra leg)
            line["fused_sweep"] = {"error": str(e)}

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the same workload
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle_lib as O
        m = O.Mapper(0.4, 0.8, 0, 1)
        m.import_points(0, cm)
        m.import_points(1, sm)
        ncpu = args.cpu_steps
        t0 = time.perf_counter()
        for i in range(ncpu):
            c, s, q, t, qp, tp = sweeps[i % len(sweeps)]
            m.set_state([0, 0, 0, 1], [0, 0, 0])
            m.step(c, s, qp, tp)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": ncpu / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"{ncpu} registrations of the same workload on 1 host thread (the reference nodes are single-threaded)"}
    if rank == 0:
        emit(line)                                   # before any teardown: a crash while freeing must not eat the result
    try:
        batch.close()
        if world > 1:
            dist.destroy_process_group()
    except Exception as e:                           # noqa: BLE001
        print(f"[bench] teardown: {e}", file=sys.stderr)


_REAL_STDOUT = None


def emit(line):
    """the ONE JSON line of this run, on the process's original stdout"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    # stdout carries exactly one JSON line: whatever libraries print while we run (NCCL's version banner, ...) goes to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sequences", type=int, default=8, help="independent sequences per GPU (one ctx + stream each)")
    ap.add_argument("--more-sequences", type=int, default=16, help="extra value leg with this many sequences per GPU (0 = skip)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-pipeline", action="store_true", help="skip the fused-sweep (scanRegistration -> odometry -> mapping) leg")
    ap.add_argument("--cpu-steps", type=int, default=60)
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 40:
            args.steps = 40          # bounded sample: ~1 s of CPU per registration
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
