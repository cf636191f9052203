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
