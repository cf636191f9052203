"""CPU test (-m "not gpu"): one whole laserOdometry step of the oracle (two passes of correspondence search + ceres::Solve,
then the pose composition of laserOdometry.cpp:278-505) against the independent Python pieces composed the same way:
test_oracle_odom_python.py_associate -> LidarEdgeFactor / LidarPlaneFactor (lidarFactor.hpp:12-104) ->
test_oracle_lm_python.py_lm (numeric Jacobians) -> t_w += q_w * t; q_w *= q."""
import numpy as np

from lmono_b200 import synth
from test_oracle_lm_python import py_lm, qrot
from test_oracle_odom_python import py_associate


def qmul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])


def py_odom_step(oracle, state, sharp, less_sharp, flat, less_flat):
    if state["inited"]:
        q, t = state["q_lc"], state["t_lc"]
        counts = []
        for _ in range(2):                                                       # :278 opti_counter
            ci, pi = py_associate(sharp, flat, state["corner_last"], state["surf_last"], q, t)
            fac = []
            cl, sl = state["corner_last"][:, :3].astype(np.float64), state["surf_last"][:, :3].astype(np.float64)
            for i, (a, b) in enumerate(ci):
                if b >= 0:
                    fac.append((0, sharp[i, :3].astype(np.float64), cl[a], cl[b]))
            nc = len(fac)
            for i, (a, b, c) in enumerate(pi):
                if b >= 0 and c >= 0:
                    n = np.cross(sl[a] - sl[b], sl[a] - sl[c])                 # lidarFactor.hpp:64-65
                    fac.append((1, flat[i, :3].astype(np.float64), sl[a], n / np.linalg.norm(n)))
            counts.append((nc, len(fac) - nc))
            f = np.zeros(len(fac), oracle.FACTOR_DTYPE)
            for k, (ty, p, a, b) in enumerate(fac):
                f["type"][k] = ty; f["p"][k] = p; f["a"][k] = a; f["b"][k] = b
            q, t, _ = py_lm(f, q, t, 4)
        state["q_lc"], state["t_lc"] = q, t
        state["t_w"] = state["t_w"] + qrot(state["q_w"], t)                     # :503-504
        state["q_w"] = qmul(state["q_w"], q)
        state["counts"] = counts
    state["inited"] = True                                                       # :267-271
    state["corner_last"], state["surf_last"] = less_sharp, less_flat             # :554-566
    return state


def test_oracle_odometry_step_equals_python_composition(oracle):
    w = synth.make_world()
    rng = np.random.default_rng(6)
    od = oracle.Odometry()
    st = dict(inited=False, q_lc=np.array([0.0, 0.0, 0.0, 1.0]), t_lc=np.zeros(3), q_w=np.array([0.0, 0.0, 0.0, 1.0]), t_w=np.zeros(3))
    for k in range(3):
        q, t = synth.loop_pose(w, 1.0 * k)
        r = oracle.scan_register(synth.raycast_sweep(w, q, t, 64, 500, rng), 64, 5.0)
        sharp, flat = r["sharp"][:100], r["flat"][:160]                          # Python loops + numeric Jacobians: keep it small
        (lq, lt), (wq, wt), rep = od.step(sharp, r["less_sharp"], flat, r["less_flat"])
        st = py_odom_step(oracle, st, sharp, r["less_sharp"], flat, r["less_flat"])
        if k == 0:
            assert rep.inited == 0
            continue
        assert [(rep.corner_corr[i], rep.plane_corr[i]) for i in range(2)] == st["counts"], k
        assert st["counts"][1][0] > 40 and st["counts"][1][1] > 80
        assert np.abs(lq - st["q_lc"]).max() <= 1e-7 and np.abs(lt - st["t_lc"]).max() <= 1e-7, (k, lt, st["t_lc"])
        assert np.abs(wq - st["q_w"]).max() <= 1e-7 and np.abs(wt - st["t_w"]).max() <= 1e-7, k
    assert abs(np.linalg.norm(st["t_w"]) - 2.0) < 0.2                             # 1 m per sweep along the road
