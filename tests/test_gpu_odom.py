"""GPU parity tests of laserOdometry (rows B1-B6): correspondences bit-exact, LM traces and
poses against the oracle restatement of Aloam/src/laserOdometry.cpp, on consecutive ray-cast
sweeps whose features come from the oracle's scanRegistration."""
import numpy as np
import pytest

from lmono_b200 import synth

pytestmark = pytest.mark.gpu


def _features(oracle, n_scans, min_range, n_sweeps, seed=1):
    w = synth.make_world()
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n_sweeps):
        q, t = synth.loop_pose(w, 1.0 * k)
        raw = synth.raycast_sweep(w, q, t, n_scans, 1875, rng)
        r = oracle.scan_register(raw, n_scans, min_range)
        out.append((r, q, t))
    return out


def rot_angle(qa, qb):
    return 2.0 * np.arccos(min(1.0, abs(float(np.dot(qa, qb)))))


@pytest.mark.parametrize("n_scans,min_range", [(64, 5.0), (32, 0.3)])
def test_odometry_sequence(gpu_ctx_factory, oracle, n_scans, min_range):
    ctx = gpu_ctx_factory(scan_line=n_scans, minimum_range=min_range)
    feats = _features(oracle, n_scans, min_range, 6)
    od = oracle.Odometry()
    prev = None
    for k, (r, q, t) in enumerate(feats):
        (glq, glt), (gwq, gwt), grep = ctx.odom_step(r["sharp"], r["less_sharp"], r["flat"], r["less_flat"])
        # the association passes of this step, checked against the oracle run on the same state
        if prev is not None:
            for which_pass, pose in ((0, prev_pose),):
                ci, pi = ctx.odom_debug(which_pass, len(r["sharp"]), len(r["flat"]))
                rci, rpi = oracle.odom_associate(r["sharp"], r["flat"], prev["less_sharp"], prev["less_flat"], pose[0], pose[1])
                assert np.array_equal(ci, rci), (k, np.argwhere(ci != rci)[:5])
                assert np.array_equal(pi, rpi), (k, np.argwhere(pi != rpi)[:5])
        (rlq, rlt), (rwq, rwt), rrep = od.step(r["sharp"], r["less_sharp"], r["flat"], r["less_flat"])
        assert grep.inited == rrep.inited
        assert list(grep.corner_corr) == list(rrep.corner_corr), (k, list(grep.corner_corr), list(rrep.corner_corr))
        assert list(grep.plane_corr) == list(rrep.plane_corr), (k, list(grep.plane_corr), list(rrep.plane_corr))
        for it in range(2 if rrep.inited else 0):
            assert grep.solve[it].iterations == rrep.solve[it].iterations
            assert grep.solve[it].termination == rrep.solve[it].termination
            assert abs(grep.solve[it].final_cost - rrep.solve[it].final_cost) <= 1e-7 * max(1.0, rrep.solve[it].final_cost)
        assert np.linalg.norm(glt - rlt) <= 1e-4 and rot_angle(glq, rlq) <= 1e-4
        assert np.linalg.norm(gwt - rwt) <= 1e-4 and rot_angle(gwq, rwq) <= 1e-4
        if k >= 3:   # after the warm start has settled the odometry should track the 1 m/sweep motion
            assert abs(np.linalg.norm(glt) - 1.0) < 0.1, glt
        prev = r
        prev_pose = (glq, glt)
    print(f"{n_scans} rings: last t_last_curr {glt}, |dt| vs oracle {np.linalg.norm(glt - rlt):.2e}, gpu {grep.ms_gpu:.3f} ms")


def test_odometry_first_frame_and_reset(gpu_ctx_factory, oracle):
    ctx = gpu_ctx_factory()
    feats = _features(oracle, 64, 5.0, 2, seed=9)
    r = feats[0][0]
    (lq, lt), (wq, wt), rep = ctx.odom_step(r["sharp"], r["less_sharp"], r["flat"], r["less_flat"])
    assert rep.inited == 0 and np.allclose(lq, [0, 0, 0, 1]) and np.allclose(lt, 0) and np.allclose(wt, 0)
    r = feats[1][0]
    _, _, rep = ctx.odom_step(r["sharp"], r["less_sharp"], r["flat"], r["less_flat"])
    assert rep.inited == 1 and rep.corner_corr[0] > 100
    ctx.odom_reset()
    _, _, rep = ctx.odom_step(r["sharp"], r["less_sharp"], r["flat"], r["less_flat"])
    assert rep.inited == 0
    # empty feature clouds must not crash and must not solve
    e = np.zeros((0, 4), np.float32)
    _, _, rep = ctx.odom_step(e, e, e, e)
    assert rep.corner_corr[0] == 0 and rep.plane_corr[0] == 0


def test_odometry_with_motion_distortion(gpu_ctx_factory, oracle):
    """#define DISTORTION 1 (laserOdometry.cpp:59, compile-time off in the reference): every point is interpolated to the
    sweep start with s = (intensity - int(intensity)) / SCAN_PERIOD through Eigen's slerp (:111-129) and the factors carry
    that s (:374-381, 472-479; lidarFactor.hpp:27-34, 73-79).  The oracle differentiates slerp analytically, the device
    evaluates the same expressions on dual numbers: counts, LM trace and poses must agree, and differ from the s = 1 run."""
    feats = _features(oracle, 64, 5.0, 6, seed=3)
    ctx = gpu_ctx_factory(distortion=1)
    plain = gpu_ctx_factory()
    od = oracle.Odometry()
    od.set_distortion(True)
    differs = 0.0
    for k, (r, q, t) in enumerate(feats):
        assert np.any(r["sharp"][:, 3] != np.floor(r["sharp"][:, 3]))          # relTime fractions are there
        (glq, glt), (gwq, gwt), grep = ctx.odom_step(r["sharp"], r["less_sharp"], r["flat"], r["less_flat"])
        (rlq, rlt), (rwq, rwt), rrep = od.step(r["sharp"], r["less_sharp"], r["flat"], r["less_flat"])
        (plq, plt), _, _ = plain.odom_step(r["sharp"], r["less_sharp"], r["flat"], r["less_flat"])
        assert grep.inited == rrep.inited
        assert list(grep.corner_corr) == list(rrep.corner_corr), (k, list(grep.corner_corr), list(rrep.corner_corr))
        assert list(grep.plane_corr) == list(rrep.plane_corr), (k, list(grep.plane_corr), list(rrep.plane_corr))
        for it in range(2 if rrep.inited else 0):
            assert grep.solve[it].iterations == rrep.solve[it].iterations, (k, it)
            assert grep.solve[it].termination == rrep.solve[it].termination, (k, it)
            assert abs(grep.solve[it].final_cost - rrep.solve[it].final_cost) <= 1e-7 * max(1.0, rrep.solve[it].final_cost)
        assert np.linalg.norm(glt - rlt) <= 1e-4 and rot_angle(glq, rlq) <= 1e-4, (k, glt, rlt)
        assert np.linalg.norm(gwt - rwt) <= 1e-4 and rot_angle(gwq, rwq) <= 1e-4
        differs = max(differs, float(np.linalg.norm(glt - plt)))
    assert differs > 1e-3, differs          # the interpolation changes the estimate: the path is really taken
    print(f"DISTORTION 1: last t_last_curr {glt} (s = 1 run: {plt}), |dt| vs oracle {np.linalg.norm(glt - rlt):.2e}")
