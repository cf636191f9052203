#!/usr/bin/env python
"""Where does the host side of a batch step go?  Per S (sequences per GPU): host time of the enqueue call with an idle
stream, device time of the step (events), and the synchronous host-API step (upload + step + read-back).
  python profiles/host_costs.py [S ...]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402

import bench  # noqa: E402
from lmono_b200 import api  # noqa: E402


def main():
    Ss = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    _, cm, sm, sweeps = bench.make_workload(0, 8)
    nsw = len(sweeps)
    main_s = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(main_s)
    d_sweeps = [(torch.from_numpy(c).to(dev), torch.from_numpy(s).to(dev)) for (c, s, *_r) in sweeps]
    h_sweeps = [(torch.from_numpy(c).pin_memory(), torch.from_numpy(s).pin_memory()) for (c, s, *_r) in sweeps]
    h_np = [(a.numpy(), b.numpy()) for a, b in h_sweeps]
    ident = ([0, 0, 0, 1], [0, 0, 0])
    out = []
    ctxs = []
    for S in Ss:
        while len(ctxs) < S:
            c_ = api.Context(device=0, stream=main_s.cuda_stream)
            c_.map_import(0, cm)
            c_.map_import(1, sm)
            c_.sync()
            ctxs.append(c_)
        batch = api.SequenceBatch(ctxs[:S])
        bargs = []
        for i in range(nsw):
            ks = [(i + 3 * s_) % nsw for s_ in range(S)]
            a_ = api.BatchArgs(S)
            a_.set_odom([(sweeps[k][4], sweeps[k][5]) for k in ks]).set_wmap_in([ident] * S)
            a_.set_device_inputs([d_sweeps[k][0].data_ptr() for k in ks], [d_sweeps[k][0].shape[0] for k in ks],
                                 [d_sweeps[k][1].data_ptr() for k in ks], [d_sweeps[k][1].shape[0] for k in ks])
            a_.set_host_inputs([h_np[k][0] for k in ks], [h_np[k][1] for k in ks])
            bargs.append(a_)
        for i in range(5):
            batch.step_device(join_stream=main_s.cuda_stream, args=bargs[i % nsw])
        torch.cuda.synchronize()
        n = 50
        t_enq = t_tot = 0.0
        dev_ms = 0.0
        for i in range(n):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0.record(main_s)
            batch.step_device(join_stream=main_s.cuda_stream, args=bargs[i % nsw])
            e1.record(main_s)
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            t_enq += t1 - t0
            t_tot += t2 - t0
            dev_ms += e0.elapsed_time(e1)
        for i in range(5):
            batch.step(args=bargs[i % nsw])
        t0 = time.perf_counter()
        for i in range(n):
            batch.step(args=bargs[i % nsw])
        t_step = time.perf_counter() - t0
        # raw ctypes call without the python result list
        a = bargs[0]
        t0 = time.perf_counter()
        for i in range(n):
            a = bargs[i % nsw]
            batch.L.lmono_map_step_batch(batch._h, batch.n, a.cv, a.sv, a.odom, a.wmap_in, batch._w, batch._wm, batch._rep)
        t_raw = time.perf_counter() - t0
        out.append({"S": S, "enqueue_host_us": 1e6 * t_enq / n, "enqueue_to_done_us": 1e6 * t_tot / n, "device_us": 1e3 * dev_ms / n,
                    "host_api_step_us": 1e6 * t_step / n, "host_api_raw_call_us": 1e6 * t_raw / n,
                    "regs_per_s_device": S / (dev_ms / n * 1e-3), "regs_per_s_host_api": S / (t_step / n)})
        print(json.dumps(out[-1]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "host_costs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
