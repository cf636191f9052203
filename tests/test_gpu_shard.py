"""GPU tests of the cube-sharded global map through the C ABI.

(1) On ONE GPU: R contexts configured as ranks 0..R-1 of R are driven in lockstep in one process, the
    all-reduce replaced by summing their workspaces -- exercises every sharded CUDA path
    (ownership in k_associate, halo-routed import / insertion, partial LM evaluation, replicated
    controller) and checks it against the unsharded CUDA path: same factor counts, normal equations
    within 1e-5 relative, poses within 1e-4, owned cubes bit-identical.
(2) With >= 2 GPUs: the same registration under torchrun with NCCL all-reduces."""
import os
import subprocess
import sys

import numpy as np
import pytest

import scenario
from lmono_b200 import shard

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lockstep(mappers, args):
    import torch
    gens = [m.step_phases(*a) for m, a in zip(mappers, args)]
    while True:
        alive = [next(g, None) for g in gens]
        if alive[0] is None:
            assert all(a is None for a in alive)
            return
        torch.cuda.synchronize()
        tot = sum(m.e.ws[:shard.WS_DOUBLES] for m in mappers)
        for m in mappers:
            m.e.ws[:shard.WS_DOUBLES] = tot
        torch.cuda.synchronize()


@pytest.mark.parametrize("nranks", [2, 3])
def test_sharded_registration_matches_unsharded_on_one_gpu(gpu_ctx_factory, nranks):
    import torch
    dev = torch.device("cuda", 0)
    cm, sm = scenario.small_map(half_xy=80.0, n_surf=250_000, n_corner=60_000)
    ts = torch.cuda.Stream(device=dev)       # real stream shared by torch and all contexts (handle 0 = "create your own")
    torch.cuda.set_stream(ts)
    stream = ts.cuda_stream
    from lmono_b200 import api
    ref = api.Context(device=0, stream=stream)
    ref.map_import(0, cm); ref.map_import(1, sm)
    ctxs = [api.Context(device=0, stream=stream) for _ in range(nranks)]
    mappers = [shard.ShardedMapper.on_gpu(c, dev, rank=r, nranks=nranks) for r, c in enumerate(ctxs)]
    kept = [(m.import_global(0, cm), m.import_global(1, sm)) for m in mappers]
    assert all(k[1] < len(sm) for k in kept)
    try:
        worst = 0.0
        for (c, s, q, t, qp, tp) in scenario.sweeps(4, seed=8, n_corner=1500, n_surf=8000, ds=9.0):
            dc, ds_ = torch.from_numpy(c).to(dev), torch.from_numpy(s).to(dev)
            rq, rt, rrep, _ = ref.map_step(c, s, qp, tp)
            _lockstep(mappers, [(dc.data_ptr(), len(c), ds_.data_ptr(), len(s), qp, tp)] * nranks)
            outs = [m.collect() for m in mappers]
            for (gq, gt, grep) in outs:
                assert list(grep.corner_num) == list(rrep.corner_num) and list(grep.surf_num) == list(rrep.surf_num)
                assert (grep.corner_from_map, grep.surf_from_map, grep.optimized) == (rrep.corner_from_map, rrep.surf_from_map, rrep.optimized)
                assert np.linalg.norm(gt - rt) <= 1e-4 and 2 * np.arccos(min(1.0, abs(float(np.dot(gq, rq))))) <= 1e-4
                assert [s_.iterations for s_ in grep.solve] == [s_.iterations for s_ in rrep.solve]
                worst = max(worst, float(np.linalg.norm(gt - rt)))
            assert all(np.array_equal(outs[0][1], o[1]) and np.array_equal(outs[0][0], o[0]) for o in outs)   # replicated controller
        print(f"sharded x{nranks} vs unsharded: worst |dt| = {worst:.2e} m; kept per rank {kept} of {(len(cm), len(sm))}")
        # after 4 registrations with insertion + refilter: every rank's OWNED cubes equal the unsharded map bit for bit
        for which in (0, 1):
            full = ref.map_export(which, 1)
            own_full = shard.owner_of_points(full, nranks)
            n_union = 0
            for r, c in enumerate(ctxs):
                part = c.map_export(which, 1)
                own = shard.owner_of_points(part, nranks) == r
                a, b = part[own], full[own_full == r]
                assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), (which, r)
                n_union += len(a)
            assert n_union == len(full)
    finally:
        torch.cuda.synchronize()
        torch.cuda.set_stream(torch.cuda.default_stream(dev))
        ref.close()
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("nranks", [2, 3])
def test_peer_memory_mode_matches_unsharded_on_one_gpu(gpu_ctx_factory, nranks):
    """Peer-memory mode (csrc/shard.cu): the ranks are contexts of this process, each on its own stream, their exchange
    blocks plain device pointers.  A registration is ONE enqueue per rank (the step graph with the gate exchange and the
    LM cluster kernel that all-gathers its sums itself); the kernels of the ranks run concurrently on the GPU and meet
    in the exchanges.  Against the unsharded ctx: same counts and LM trace, poses within 1e-4, owned cubes bit-identical."""
    import torch
    from lmono_b200 import api
    dev = torch.device("cuda", 0)
    cm, sm = scenario.small_map(half_xy=80.0, n_surf=250_000, n_corner=60_000)
    ref = api.Context(device=0)
    ref.map_import(0, cm); ref.map_import(1, sm)
    ctxs = [api.Context(device=0) for _ in range(nranks)]          # stream=NULL: every ctx owns a non-blocking stream
    mappers = shard.PeerMemoryMapper.connect_local(ctxs)
    kept = [(m.import_global(0, cm), m.import_global(1, sm)) for m in mappers]
    assert all(k[1] < len(sm) for k in kept)
    try:
        worst = 0.0
        for k, (c, s, q, t, qp, tp) in enumerate(scenario.sweeps(5, seed=8, n_corner=1500, n_surf=8000, ds=9.0)):
            dc, ds_ = torch.from_numpy(c).to(dev), torch.from_numpy(s).to(dev)
            torch.cuda.synchronize()
            rq, rt, rrep, _ = ref.map_step(c, s, qp, tp)
            for m in mappers:
                m.step(dc.data_ptr(), len(c), ds_.data_ptr(), len(s), qp, tp)       # enqueue-only
            outs = [m.collect() for m in mappers]
            for (gq, gt, grep) in outs:
                assert list(grep.corner_num) == list(rrep.corner_num) and list(grep.surf_num) == list(rrep.surf_num), k
                assert (grep.corner_from_map, grep.surf_from_map, grep.optimized) == (rrep.corner_from_map, rrep.surf_from_map, rrep.optimized)
                assert np.linalg.norm(gt - rt) <= 1e-4 and 2 * np.arccos(min(1.0, abs(float(np.dot(gq, rq))))) <= 1e-4
                assert [s_.iterations for s_ in grep.solve] == [s_.iterations for s_ in rrep.solve]
                assert [s_.termination for s_ in grep.solve] == [s_.termination for s_ in rrep.solve]
                worst = max(worst, float(np.linalg.norm(gt - rt)))
            assert all(np.array_equal(outs[0][1], o[1]) and np.array_equal(outs[0][0], o[0]) for o in outs)   # replicated controllers, same bits
        st = [c.shard_xchg_stats() for c in ctxs]
        assert all(x["epoch"] == st[0]["epoch"] and x["epoch"] >= 5 * 3 for x in st), st
        print(f"peer-memory x{nranks} vs unsharded: worst |dt| = {worst:.2e} m; exchanges {st[0]['epoch']}, "
              f"{st[0]['wait_ns'] / max(st[0]['exchanges'], 1) / 1e3:.1f} us per exchange (ranks share ONE GPU here)")
        for which in (0, 1):
            full = ref.map_export(which, 1)
            own_full = shard.owner_of_points(full, nranks)
            n_union = 0
            for r, c in enumerate(ctxs):
                part = c.map_export(which, 1)
                own = shard.owner_of_points(part, nranks) == r
                a, b = part[own], full[own_full == r]
                assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), (which, r)
                n_union += len(a)
            assert n_union == len(full)
    finally:
        torch.cuda.synchronize()
        for c in ctxs:
            c.close()
        ref.close()


def test_device_import_applies_the_same_rule_as_the_host_prefilter(gpu_ctx_factory):
    """lmono_map_import on a sharded ctx keeps owner + halo points itself: importing the whole cloud
    gives the same map as importing the numpy-prefiltered subset (host rule == device rule)."""
    import torch
    from lmono_b200 import api
    dev = torch.device("cuda", 0)
    cm, sm = scenario.small_map(half_xy=80.0, n_surf=250_000, n_corner=60_000)
    a, b = api.Context(device=0), api.Context(device=0)
    try:
        ma = shard.ShardedMapper.on_gpu(a, dev, rank=1, nranks=4)
        shard.ShardedMapper.on_gpu(b, dev, rank=1, nranks=4)
        ma.import_global(1, sm)
        for chunk in shard.cube_chunks(sm):
            b.map_import(1, np.ascontiguousarray(chunk))
        ea, eb = a.map_export(1, 1), b.map_export(1, 1)
        assert 0 < len(ea)
        assert ea.shape == eb.shape and np.array_equal(ea.view(np.uint32), eb.view(np.uint32))
    finally:
        a.close(); b.close()


def test_two_gpu_nccl_sharded_registration():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731",
                          os.path.join(ROOT, "tests", "shard_nccl_worker.py")], capture_output=True, text=True, timeout=600)
    print(out.stdout[-3000:], out.stderr[-3000:])
    assert out.returncode == 0
    assert "SHARD_NCCL_OK" in out.stdout
