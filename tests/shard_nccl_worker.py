"""torchrun worker of tests/test_gpu_shard.py::test_two_gpu_nccl_sharded_registration and of the C-5
measurement (`python -m torch.distributed.run --nproc-per-node N tests/shard_nccl_worker.py [--bench]`):
N ranks hold a cube-sharded map, register the same replicated sweeps with NCCL all-reduces of the
35-double workspace, and every rank checks the result against an unsharded ctx on its own GPU."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import scenario  # noqa: E402
from lmono_b200 import api, shard  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    # one real (non-default) stream shared by torch / NCCL and both contexts: a 0 handle would make each ctx
    # create its own stream and the all-reduces would not be ordered with the kernels
    ts = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(ts)
    stream = ts.cuda_stream
    cm, sm = scenario.small_map(half_xy=80.0, n_surf=250_000, n_corner=60_000)
    results = {}
    for mode in ("nccl", "p2p"):
        ctx = api.Context(device=local, stream=stream)
        m = shard.ShardedMapper.on_gpu(ctx, dev) if mode == "nccl" else shard.PeerMemoryMapper.connect(ctx)
        kept = (m.import_global(0, cm), m.import_global(1, sm))
        ref = api.Context(device=local, stream=stream)
        ref.map_import(0, cm); ref.map_import(1, sm)
        worst = 0.0
        times = []
        for (c, s, q, t, qp, tp) in scenario.sweeps(6, seed=8, n_corner=1500, n_surf=8000, ds=9.0):
            dc, ds_ = torch.from_numpy(c).to(dev), torch.from_numpy(s).to(dev)
            rq, rt, rrep, _ = ref.map_step(c, s, qp, tp)
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            m.step(dc.data_ptr(), len(c), ds_.data_ptr(), len(s), qp, tp)
            gq, gt, grep = m.collect()
            times.append(time.perf_counter() - t0)
            assert list(grep.corner_num) == list(rrep.corner_num) and list(grep.surf_num) == list(rrep.surf_num), (mode, list(grep.corner_num), list(rrep.corner_num))
            assert [x.iterations for x in grep.solve] == [x.iterations for x in rrep.solve], mode
            assert np.linalg.norm(gt - rt) <= 1e-4 and 2 * np.arccos(min(1.0, abs(float(np.dot(gq, rq))))) <= 1e-4
            worst = max(worst, float(np.linalg.norm(gt - rt)))
            # all ranks hold the same pose bit for bit
            pose = torch.tensor(np.concatenate([gq, gt]), device=dev)
            lo, hi = pose.clone(), pose.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            assert torch.equal(lo, hi), mode
        for which in (0, 1):
            full, part = ref.map_export(which, 1), ctx.map_export(which, 1)
            a = part[shard.owner_of_points(part, world) == rank]
            b = full[shard.owner_of_points(full, world) == rank]
            assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), mode
        results[mode] = {"kept": kept, "worst_dt_m": worst, "ms_per_registration_wall": 1e3 * float(np.median(times)),
                         "host_allreduces": getattr(m, "n_allreduce", 0),
                         "device_exchanges": ctx.shard_xchg_stats()["epoch"] if mode == "p2p" else 0}
        torch.cuda.synchronize(); dist.barrier()
        ctx.close(); ref.close()
    print(json.dumps({"rank": rank, "world": world, "of": (len(cm), len(sm)), **results}))
    dist.barrier()
    if rank == 0:
        print("SHARD_NCCL_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
