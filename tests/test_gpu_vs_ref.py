"""GPU parity tests against the REFERENCE'S OWN SOURCES (no oracle in between): the CUDA library through the C ABI beside
the reference's scanRegistration / laserOdometry / laserMapping nodes as compiled from /root/reference into oracle/_ref
(`make -C oracle ref`; the prebuilt files travel to the GPU box with the snapshot, the tests skip where there are none).
The nodes run on this box's host cores through their own callbacks and main loops; PCL / FLANN / Eigen / Ceres underneath
them are the stand-ins of oracle/refstubs.  Bars: labels, ring order, curvature and feature picks
bit for bit; poses within 1e-4 m / 1e-4 rad per sweep (north_star); map cubes point for point (coordinates to fp32
summation order: the device sums a voxel's members in input order, libstdc++'s introsort leaves them in another)."""
import numpy as np
import pytest

import oracle_lib
import scenario
from lmono_b200 import synth

pytestmark = pytest.mark.gpu


def _need(name):
    if oracle_lib.ref_lib(name) is None:
        pytest.skip(f"oracle/_ref/libref_{name}.so not built")


@pytest.mark.parametrize("n_scans,min_range,n_az", [(64, 5.0, 1875), (32, 0.3, 1875), (16, 0.3, 1800)])
def test_scan_register_equals_reference_node(gpu_ctx_factory, n_scans, min_range, n_az):
    _need("scanreg")
    ctx = gpu_ctx_factory(scan_line=n_scans, minimum_range=min_range)
    w = synth.make_world()
    for seed in (11, 12):
        q, t = synth.loop_pose(w, 3.0 + seed)
        raw = synth.raycast_sweep(w, q, t, n_scans, n_az, np.random.default_rng(seed))
        raw[5, 1] = np.nan
        got = ctx.scan_register(raw, want_debug=True)
        ref = oracle_lib.ref_scan_register(raw, n_scans, min_range)
        assert got["full"].shape == ref["full"].shape and len(ref["full"]) > 10000
        assert np.array_equal(got["full"][:, :3].view(np.uint32), ref["full"][:, :3].view(np.uint32))
        assert np.array_equal(np.floor(got["full"][:, 3]), np.floor(ref["full"][:, 3]))                 # ring ids
        assert np.abs(got["full"][:, 3] - ref["full"][:, 3]).max() <= 4e-6                              # 0.1 relTime: atan2f of libm vs the device's
        assert np.array_equal(got["curvature"][5:-5].view(np.uint32), ref["curvature"][5:-5].view(np.uint32))
        assert np.array_equal(got["labels"], ref["labels"])
        for k in ("sharp", "less_sharp", "flat"):
            assert got[k].shape == ref[k].shape and len(ref[k]) > 0, k
            assert np.array_equal(got[k][:, :3].view(np.uint32), ref[k][:, :3].view(np.uint32)), k
        assert got["less_flat"].shape == ref["less_flat"].shape
        assert np.abs(got["less_flat"] - ref["less_flat"]).max() <= 3e-5


def test_odometry_equals_reference_node(gpu_ctx_factory):
    _need("odom")
    ctx = gpu_ctx_factory(scan_line=32, minimum_range=0.3, max_cubes_corner=8, max_cubes_surf=8, cube_capacity_corner=1024, cube_capacity_surf=1024)
    ref = oracle_lib.RefOdometry()
    w = synth.make_world()
    rng = np.random.default_rng(9)
    worst = 0.0
    for k in range(8):
        q, t = synth.loop_pose(w, 1.0 * k)
        raw = synth.raycast_sweep(w, q, t, 32, 1875, rng)
        f = ctx.scan_register(raw)
        (glq, glt), (gwq, gwt), grep = ctx.odom_step(f["sharp"], f["less_sharp"], f["flat"], f["less_flat"])
        (rlq, rlt), (rwq, rwt), cnt = ref.step(f["sharp"], f["less_sharp"], f["flat"], f["less_flat"], f["full"])
        assert [int(cnt[0]), int(cnt[1])] == [grep.corner_corr[1], grep.plane_corr[1]], k
        d = max(np.abs(glq - rlq).max(), np.abs(glt - rlt).max(), np.abs(gwq - rwq).max(), np.abs(gwt - rwt).max())
        worst = max(worst, float(d))
        assert d <= 1e-4, (k, d)
    assert cnt[0] > 100 and cnt[1] > 300 and np.linalg.norm(gwt) > 5.0
    print(f"odometry vs reference node: worst pose difference {worst:.2e}")


def test_mapping_equals_reference_node(gpu_ctx_factory):
    _need("mapping")
    ctx = gpu_ctx_factory()
    ref = oracle_lib.RefMapper()
    worst = 0.0
    n_opt = 0
    for k, (c, s, qg, tg, qo, to) in enumerate(scenario.sweeps(10, n_corner=1500, n_surf=9000)):
        full = np.concatenate([c, s])[:6000]
        gq, gt, grep, gfull = ctx.map_step(c, s, qo, to, full_res=full)
        rq, rt, _, rcen, rfull = ref.step(c, s, qo, to, full)
        assert list(grep.cen) == rcen, k
        d = max(np.abs(gq - rq).max(), np.abs(gt - rt).max())
        worst = max(worst, float(d))
        assert d <= 1e-4, (k, d)
        assert np.abs(gfull - rfull).max() <= 1e-4, k
        for which in (0, 1):
            a, b = ctx.map_export(which, 1), ref.export(which)
            assert a.shape == b.shape, (k, which, a.shape, b.shape)                 # same voxels occupied in every cube, in the same order
            assert np.abs(a - b).max() <= 1e-4, (k, which)
        n_opt += int(grep.optimized)
    assert n_opt >= 8 and ctx.last_fault() == 0
    print(f"mapping vs reference node: worst pose difference {worst:.2e}")
