"""End-to-end checks on the GPU box: (1) the three stages chained through the Python binding agree
with the oracle chain sweep by sweep; (2) the ROS-free C++ harness nodes/replay_kitti (the host
code path a catkin node would use) produces the same poses from KITTI-layout .bin files."""
import os
import subprocess

import numpy as np
import pytest

from lmono_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _raw_sweeps(n, seed=2):
    w = synth.make_world()
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        q, t = synth.loop_pose(w, 1.0 * k)
        out.append((synth.raycast_sweep(w, q, t, 64, 1875, rng), q, t))
    return out


def _chain_gpu(ctx, raw):
    r = ctx.scan_register(raw)
    (_, _), (oq, ot), orep = ctx.odom_step(r["sharp"], r["less_sharp"], r["flat"], r["less_flat"])
    mq, mt, mrep, _ = ctx.map_step(r["less_sharp"], r["less_flat"], oq, ot)
    return r, (oq, ot), (mq, mt), orep, mrep


def test_three_stage_chain_matches_oracle(gpu_ctx_factory, oracle):
    ctx = gpu_ctx_factory()
    od, om = oracle.Odometry(), oracle.Mapper()
    worst = 0.0
    for k, (raw, q, t) in enumerate(_raw_sweeps(8)):
        r, (oq, ot), (mq, mt), orep, mrep = _chain_gpu(ctx, raw)
        rr = oracle.scan_register(raw, 64, 5.0)
        assert np.array_equal(r["labels"], rr["labels"])
        (_, _), (roq, rot), _ = od.step(rr["sharp"], rr["less_sharp"], rr["flat"], rr["less_flat"])
        rmq, rmt, rrep, _ = om.step(rr["less_sharp"], rr["less_flat"], roq, rot)
        assert np.linalg.norm(ot - rot) <= 1e-4 and np.linalg.norm(mt - rmt) <= 1e-4, (k, ot, rot, mt, rmt)
        assert abs(abs(float(np.dot(mq, rmq))) - 1.0) <= 1e-8
        assert (mrep.corner_from_map, mrep.surf_from_map, mrep.optimized) == (rrep.corner_from_map, rrep.surf_from_map, rrep.optimized)
        worst = max(worst, float(np.linalg.norm(mt - rmt)))
    # the map starts empty: the first sweeps only fill it, later ones are optimised against it
    assert mrep.optimized == 1
    # driving 1 m per sweep along x: the mapped pose must follow
    assert abs(np.linalg.norm(mt) - 7.0) < 0.3
    print("worst mapped-pose deviation vs oracle over 8 sweeps: %.3e m" % worst)


def test_fused_sweep_equals_three_calls(gpu_ctx_factory):
    """lmono_sweep_step (features and odometry pose stay in device memory) against the three separate calls: same bits"""
    a, b = gpu_ctx_factory(), gpu_ctx_factory()
    for k, (raw, q, t) in enumerate(_raw_sweeps(8, seed=4)):
        r, (oq, ot), (mq, mt), orep, mrep = _chain_gpu(a, raw)
        (lq, lt), (foq, fot), (fmq, fmt), srep, forep, fmrep = b.sweep_step(raw)
        assert (srep.n_kept, srep.n_sharp, srep.n_less_sharp, srep.n_flat, srep.n_less_flat) == \
               (len(r["full"]), len(r["sharp"]), len(r["less_sharp"]), len(r["flat"]), len(r["less_flat"]))
        assert np.array_equal(foq, oq) and np.array_equal(fot, ot), k
        assert np.array_equal(fmq, mq) and np.array_equal(fmt, mt), k
        assert list(forep.corner_corr) == list(orep.corner_corr) and list(forep.plane_corr) == list(orep.plane_corr)
        assert list(fmrep.corner_num) == list(mrep.corner_num) and list(fmrep.surf_num) == list(mrep.surf_num)
        assert (fmrep.corner_from_map, fmrep.surf_from_map, fmrep.optimized) == (mrep.corner_from_map, mrep.surf_from_map, mrep.optimized)
    for w in (0, 1):
        assert np.array_equal(a.map_export(w, 1).view(np.uint32), b.map_export(w, 1).view(np.uint32))
    assert fmrep.optimized == 1


def test_async_sweep_and_sweep_batch_equal_the_synchronous_sweep(gpu_ctx_factory):
    """lmono_sweep_submit / _wait (no host round trip inside a sweep: the feature counts stay on the device, grids are sized
    from bounds) and lmono_sweep_step_batch (4 sequences on one shared caller stream, each sweep on a private stream) must
    give exactly what lmono_sweep_step gives, sequence by sequence and bit for bit."""
    import torch
    from lmono_b200 import api
    NS, NK = 4, 6
    seqs = [_raw_sweeps(NK, seed=20 + s) for s in range(NS)]
    ref = []
    for s in range(NS):
        c = gpu_ctx_factory()
        ref.append([c.sweep_step(raw) for (raw, _, _) in seqs[s]])
        ref[-1].append([c.map_export(w, 1) for w in (0, 1)])
        c.close()
    # (1) one sequence through submit / wait
    c = gpu_ctx_factory()
    for k, (raw, _, _) in enumerate(seqs[0]):
        raw = np.ascontiguousarray(raw, np.float32)
        c.sweep_submit(raw)
        got = c.sweep_wait()
        exp = ref[0][k]
        for a in range(3):
            assert np.array_equal(got[a][0], exp[a][0]) and np.array_equal(got[a][1], exp[a][1]), (k, a)
        assert (got[3].n_kept, got[3].n_sharp, got[3].n_less_sharp, got[3].n_flat, got[3].n_less_flat) == \
               (exp[3].n_kept, exp[3].n_sharp, exp[3].n_less_sharp, exp[3].n_flat, exp[3].n_less_flat)
        assert list(got[4].corner_corr) == list(exp[4].corner_corr) and list(got[4].plane_corr) == list(exp[4].plane_corr)
        assert list(got[5].corner_num) == list(exp[5].corner_num) and list(got[5].surf_num) == list(exp[5].surf_num)
    for w in (0, 1):
        assert np.array_equal(c.map_export(w, 1).view(np.uint32), ref[0][NK][w].view(np.uint32))
    # (2) four sequences per call, all created on ONE caller stream
    st = torch.cuda.Stream()
    ctxs = [gpu_ctx_factory(stream=st.cuda_stream) for _ in range(NS)]
    batch = api.SweepBatch(ctxs)
    for k in range(NK):
        raws = [torch.from_numpy(np.ascontiguousarray(seqs[s][k][0], np.float32)).pin_memory() for s in range(NS)]
        res = batch.step([r.numpy() for r in raws])
        for s in range(NS):
            (oq, ot), (mq, mt), srep, orep, mrep = res[s]
            exp = ref[s][k]
            assert np.array_equal(oq, exp[1][0]) and np.array_equal(ot, exp[1][1]), (s, k)
            assert np.array_equal(mq, exp[2][0]) and np.array_equal(mt, exp[2][1]), (s, k)
            assert list(mrep.corner_num) == list(exp[5].corner_num) and list(mrep.surf_num) == list(exp[5].surf_num)
    for s in range(NS):
        for w in (0, 1):
            assert np.array_equal(ctxs[s].map_export(w, 1).view(np.uint32), ref[s][NK][w].view(np.uint32)), (s, w)


def test_cpp_replay_harness_matches_python_binding(gpu_ctx_factory, tmp_path):
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "nodes")], check=True)
    sweeps = _raw_sweeps(4, seed=6)
    for k, (raw, _, _) in enumerate(sweeps):
        raw.astype(np.float32).tofile(tmp_path / f"{k:06d}.bin")
    exe = os.path.join(ROOT, "nodes", "replay_kitti")
    out = subprocess.run([exe, str(tmp_path), "4"], check=True, capture_output=True, text=True).stdout
    lines = [l.split() for l in out.strip().splitlines()]
    assert len(lines) == 4
    # the same sweeps through one lmono_sweep_step each: identical lines
    out_f = subprocess.run([exe, str(tmp_path), "4", "64", "5", "fused"], check=True, capture_output=True, text=True).stdout
    assert out_f == out
    ctx = gpu_ctx_factory()
    for k, (raw, _, _) in enumerate(sweeps):
        _, (oq, ot), (mq, mt), _, _ = _chain_gpu(ctx, raw)
        vals = np.array([float(x) for x in lines[k][1:15]])
        assert np.allclose(vals[4:7], ot, atol=2e-6) and np.allclose(vals[11:14], mt, atol=2e-6)
        assert np.allclose(vals[0:4], oq, atol=2e-9) and np.allclose(vals[7:11], mq, atol=2e-9)


@pytest.mark.gpu
def test_stage_only_context_allocates_no_map():
    """lmono_params::stages: a scanRegistration / laserOdometry-only ctx (what those two nodes create) does not allocate the cube
    map's slab pools and refuses the mapping calls; its own stage gives the same results as a full ctx."""
    import torch
    from lmono_b200 import api, synth
    wld = synth.make_world()
    rng = np.random.default_rng(5)
    raw = np.ascontiguousarray(synth.raycast_sweep_torch(wld, *synth.loop_pose(wld, 3.0), 64, 900, rng, device=torch.device("cuda", 0)), np.float32)
    free0 = torch.cuda.mem_get_info(0)[0]
    small = api.Context(device=0, stages=1 | 2)
    used_small = free0 - torch.cuda.mem_get_info(0)[0]
    full = api.Context(device=0)
    used_full = free0 - torch.cuda.mem_get_info(0)[0] - used_small
    try:
        assert used_small < 0.5e9 < 2e9 < used_full, (used_small, used_full)
        a, b = small.scan_register(raw), full.scan_register(raw)
        for k in ("full", "sharp", "less_sharp", "flat", "less_flat"):
            assert a[k].shape == b[k].shape and (a[k].view(np.uint32) == b[k].view(np.uint32)).all(), k
        c, s = a["less_sharp"], a["less_flat"]
        with pytest.raises(api.LmonoError):
            small.map_step(c, s, [0, 0, 0, 1], [0, 0, 0])
        with pytest.raises(api.LmonoError):
            small.map_import(0, c)
    finally:
        small.close(); full.close()
