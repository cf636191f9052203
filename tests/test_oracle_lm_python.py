"""CPU test (-m "not gpu"): the oracle's Ceres-style solve (C: analytic Jacobians of Eigen's q * v formula times the 4x3
local-parameterisation Jacobian, loss corrector, Levenberg-Marquardt trust region) against an independent Python
restatement that differentiates NUMERICALLY: residuals of Aloam/src/lidarFactor.hpp evaluated at Plus(x, +-h e_k)
(EigenQuaternionParameterization), HuberLoss(0.1) corrector, and the trust-region loop of Ceres 1.14
(trust_region_minimizer.cc / levenberg_marquardt_strategy.cc, options of laserMapping.cpp:713-720).  Same iteration
trace, same termination, final pose to 1e-8."""
import numpy as np

from test_oracle_primitives import _plus, _random_factors


def qrot(q, p):
    u, w = q[:3], q[3]
    uv = 2.0 * np.cross(u, p)
    return p + w * uv + np.cross(u, uv)


def residuals(f, q, t):
    """corrected residual vector (sqrt(rho') r per block) and the robust cost"""
    out, cost = [], 0.0
    for k in range(len(f)):
        lp = qrot(q, f["p"][k]) + t
        ty = f["type"][k]
        if ty == 0:                                              # LidarEdgeFactor (:12-55)
            a, b = f["a"][k], f["b"][k]
            r = np.cross(lp - a, lp - b) / np.linalg.norm(a - b)
        elif ty == 1:                                            # LidarPlaneFactor (:57-104): (lp - j) . n
            r = np.array([(lp - f["a"][k]) @ f["b"][k]])
        else:                                                    # LidarPlaneNormFactor (:106-138): n . lp + d
            r = np.array([f["a"][k] @ lp + f["b"][k][0]])
        s = float(r @ r)
        if s > 0.01:                                             # HuberLoss(0.1): rho = 2 a sqrt(s) - a^2, rho' = a / sqrt(s)
            cost += 0.5 * (0.2 * np.sqrt(s) - 0.01)
            r = r * np.sqrt(0.1 / np.sqrt(s))
        else:
            cost += 0.5 * s
        out.append(r)
    return np.concatenate(out), cost


def evaluate(f, q, t, h=1e-6):
    r, cost = residuals(f, q, t)
    # the corrector scales J by the same sqrt(rho') as r (rho'' <= 0), and rho' depends on the point of evaluation only
    # weakly: differentiate the UNcorrected residual direction by keeping the weights of the centre point
    J = np.zeros((len(r), 6))
    w = []
    for k in range(len(f)):
        lp = qrot(q, f["p"][k]) + t
        ty = f["type"][k]
        if ty == 0:
            rr = np.cross(lp - f["a"][k], lp - f["b"][k]) / np.linalg.norm(f["a"][k] - f["b"][k])
        elif ty == 1:
            rr = np.array([(lp - f["a"][k]) @ f["b"][k]])
        else:
            rr = np.array([f["a"][k] @ lp + f["b"][k][0]])
        s = float(rr @ rr)
        w += [np.sqrt(0.1 / np.sqrt(s)) if s > 0.01 else 1.0] * len(rr)
    w = np.array(w)

    def raw(qq, tt):
        o = []
        for k in range(len(f)):
            lp = qrot(qq, f["p"][k]) + tt
            ty = f["type"][k]
            if ty == 0:
                o.append(np.cross(lp - f["a"][k], lp - f["b"][k]) / np.linalg.norm(f["a"][k] - f["b"][k]))
            elif ty == 1:
                o.append(np.array([(lp - f["a"][k]) @ f["b"][k]]))
            else:
                o.append(np.array([f["a"][k] @ lp + f["b"][k][0]]))
        return np.concatenate(o)

    for j in range(6):
        d = np.zeros(6)
        d[j] = h
        qp, tp = _plus(q, t, d)
        qm, tm = _plus(q, t, -d)
        J[:, j] = w * (raw(qp, tp) - raw(qm, tm)) / (2 * h)
    return r, J, cost


def py_lm(f, q, t, max_iter=4):
    r, J, cost = evaluate(f, q, t)
    scale = 1.0 / (1.0 + np.sqrt((J * J).sum(0)))               # jacobi_scaling, computed once
    initial = cost
    radius, dec = 1e4, 2.0
    it = nsucc = 0
    term = 0
    reuse = False
    diag = None
    x = np.r_[q, t]
    trace = [cost]
    while True:
        Js = J * scale
        g = Js.T @ r
        if it >= max_iter:
            term = 0
            break
        it += 1
        if not reuse:
            diag = np.clip((Js * Js).sum(0), 1e-6, 1e32)
        A = Js.T @ Js + np.diag(diag / radius)
        d = np.linalg.solve(A, -g)
        model = -(d @ g + 0.5 * d @ (Js.T @ Js) @ d)
        assert model > 0
        qc, tc = _plus(x[:4], x[4:], d * scale)
        xc = np.r_[qc, tc]
        if np.linalg.norm(xc - x) <= 1e-8 * (np.linalg.norm(x) + 1e-8):
            term = 2
            break
        rc, Jc, cc = evaluate(f, qc, tc)
        if abs(cost - cc) <= 1e-6 * cost:
            term = 3
            break
        rho = (cost - cc) / model
        if rho > 1e-3:
            x, r, J, cost = xc, rc, Jc, cc
            radius = min(radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3), 1e16)
            dec, reuse = 2.0, False
            nsucc += 1
            trace.append(cost)
        else:
            radius /= dec
            dec *= 2.0
            reuse = True
    return x[:4], x[4:], dict(iterations=it, num_successful=nsucc, termination=term, initial_cost=initial, final_cost=cost)


def test_oracle_lm_equals_numeric_python_restatement(oracle):
    rng = np.random.default_rng(17)
    for trial in range(3):
        f = _random_factors(oracle, rng, 150)
        q0 = np.array([0.01, -0.006, 0.008, 1.0]) * np.r_[rng.uniform(0.5, 1.5, 3), 1.0]
        q0 /= np.linalg.norm(q0)
        t0 = rng.uniform(-0.08, 0.08, 3)
        q, t, s = oracle.lm_solve(f, q0, t0, 4)
        pq, pt, ps = py_lm(f, q0, t0, 4)
        assert (s.iterations, s.num_successful, s.termination) == (ps["iterations"], ps["num_successful"], ps["termination"]), trial
        assert abs(s.initial_cost - ps["initial_cost"]) <= 1e-12 * ps["initial_cost"]
        assert abs(s.final_cost - ps["final_cost"]) <= 1e-7 * ps["final_cost"]
        assert np.abs(q - pq).max() <= 1e-8 and np.abs(t - pt).max() <= 1e-8, (trial, q - pq, t - pt)
