"""Downstream check of the pose parity bar: the fusion estimator consumes /aft_mapped_to_init through
LASERFactor (mono_lidar_mapping/include/factor/LaserFactor.h:24-62, built at src/image_process/Estimator.cc:1167-1180):
for consecutive frames i, j it stores delta_ij = R_i^T R_j and delta_pij = R_i^T (P_j - P_i) of the LiDAR poses as the
measurement and weights the residual [Q_i^-1 (P_j - P_i) - delta_pij ; 2 vec(delta_ij^-1 (Q_i^-1 Q_j))] by
sqrt_info = laser_w * factor_weight * I (Estimator.cc:95; 2 * 1000 in config/kitti_config_00.yaml:69-70).

The test runs the whole A-LOAM chain on the GPU and through the oracle, builds the factor's measurement from the GPU
poses and evaluates its weighted residual at the oracle's poses: what the estimator would see change if the reference
nodes were swapped for the drop-in.  The north_star bar (1e-4 m / 1e-4 rad per sweep) corresponds to a weighted residual
of 0.2 (a fifth of one sigma); the measured difference is many orders below that."""
import numpy as np
import pytest

from lmono_b200 import synth

pytestmark = pytest.mark.gpu
SQRT_INFO = 2.0 * 1000.0          # laser_w * factor_weight, config/kitti_config_00.yaml:69-70


def _qmul(a, b):
    return synth.quat_mul(a, b)


def _qinv(q):
    return np.array([-q[0], -q[1], -q[2], q[3]]) / float(np.dot(q, q))


def laser_factor_residual(meas_i, meas_j, at_i, at_j):
    """LaserFactor.h:31-36 (measurement from poses meas_*) and :45-62 (residual at poses at_*); poses are (q xyzw, t)"""
    (qi, pi), (qj, pj) = meas_i, meas_j
    Ri, Rj = synth.quat_to_rot(qi), synth.quat_to_rot(qj)
    delta_q = _qmul(_qinv(qi), qj)                      # Quaterniond(R_i^T R_j)
    delta_p = Ri.T @ (pj - pi)
    (Qi, Pi), (Qj, Pj) = at_i, at_j
    r_p = synth.quat_to_rot(_qinv(Qi)) @ (Pj - Pi) - delta_p
    r_q = 2.0 * _qmul(_qinv(delta_q), _qmul(_qinv(Qi), Qj))[:3]
    return SQRT_INFO * np.concatenate([r_p, r_q])


def test_laser_factor_sees_the_same_relative_poses(gpu_ctx_factory, oracle):
    w = synth.make_world()
    rng = np.random.default_rng(31)
    ctx = gpu_ctx_factory()
    od, om = oracle.Odometry(), oracle.Mapper()
    gpu, ref = [], []
    for k in range(14):
        q, t = synth.loop_pose(w, 1.0 * k)
        raw = synth.raycast_sweep(w, q, t, 64, 1875, rng)
        o = ctx.sweep_step(np.ascontiguousarray(raw, np.float32))
        gpu.append(o[2])                               # mapped pose = /aft_mapped_to_init
        r = oracle.scan_register(raw, 64, 5.0)
        _, (wq, wt), _ = od.step(r["sharp"], r["less_sharp"], r["flat"], r["less_flat"])
        mq, mt, _, _ = om.step(r["less_sharp"], r["less_flat"], wq, wt)
        ref.append((mq, mt))
    worst = 0.0
    for i in range(len(gpu) - 1):
        res = laser_factor_residual(gpu[i], gpu[i + 1], ref[i], ref[i + 1])
        worst = max(worst, float(np.abs(res).max()))
        # sanity: the factor is sensitive at the level of the bar -- a 1e-4 m shift of P_j shows up as 0.2
        shifted = (ref[i + 1][0], ref[i + 1][1] + np.array([1e-4, 0.0, 0.0]))
        assert 0.15 < np.abs(laser_factor_residual(gpu[i], gpu[i + 1], ref[i], shifted)).max() < 0.25
    print(f"largest weighted LASERFactor residual between the GPU and the oracle pose streams: {worst:.3e} (bar: 0.2 = 1e-4 m x sqrt_info 2000)")
    assert worst < 1e-3
