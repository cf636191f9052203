"""CPU test (-m "not gpu"): the oracle's laserOdometry correspondence search (C, KD-tree 1-NN + ring-window scans) against an
independent Python restatement of Aloam/src/laserOdometry.cpp:111-129, 299-483 (numpy float32 brute-force 1-NN, plain
Python ring-window loops).  Every (closest, min2[, min3]) index triple must agree."""
import numpy as np

from lmono_b200 import synth

f32 = np.float32
DIST2 = 25.0           # DISTANCE_SQ_THRESHOLD (:65)
NEARBY = 2.5           # NEARBY_SCAN (:64)


def to_start(q, t, pts):
    """TransformToStart (:111-129) with DISTORTION 0: s = 1, q_point_last = q_last_curr; Eigen's q * v in double, stored float."""
    qv, w = np.asarray(q[:3], np.float64), float(q[3])
    out = np.zeros((len(pts), 3), np.float32)
    for i, p in enumerate(pts[:, :3].astype(np.float64)):
        uv = 2.0 * np.cross(qv, p)
        out[i] = (p + w * uv + np.cross(qv, uv) + np.asarray(t, np.float64)).astype(np.float32)
    return out


def sqdis32(a, sel):
    """float products and sums, left to right (:322-327)"""
    d = a - sel
    return f32(f32(f32(d[0] * d[0]) + f32(d[1] * d[1])) + f32(d[2] * d[2]))


def nn1(targets, sel):
    d = targets[:, :3] - sel[None, :]
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]            # float32 throughout, FLANN L2_Simple order
    j = int(np.argmin(d2))                                                       # first minimum = lowest index on ties
    return j, d2[j]


def py_associate(sharp, flat, corner_last, surf_last, q, t):
    ring_c = corner_last[:, 3].astype(np.int64)                                   # int(intensity): truncation (values >= 0)
    ring_s = surf_last[:, 3].astype(np.int64)
    ci = np.full((len(sharp), 2), -1, np.int32)
    pi = np.full((len(flat), 3), -1, np.int32)
    for i, sel in enumerate(to_start(q, t, sharp)):                               # :299-362
        c, d2 = nn1(corner_last, sel)
        if not float(d2) < DIST2:
            continue
        rid = ring_c[c]
        best, m2 = DIST2, -1
        for j in range(c + 1, len(corner_last)):
            if ring_c[j] <= rid:
                continue
            if ring_c[j] > rid + NEARBY:
                break
            d = float(sqdis32(corner_last[j, :3], sel))
            if d < best:
                best, m2 = d, j
        for j in range(c - 1, -1, -1):
            if ring_c[j] >= rid:
                continue
            if ring_c[j] < rid - NEARBY:
                break
            d = float(sqdis32(corner_last[j, :3], sel))
            if d < best:
                best, m2 = d, j
        ci[i] = (c, m2)
    for i, sel in enumerate(to_start(q, t, flat)):                                # :387-455
        c, d2 = nn1(surf_last, sel)
        if not float(d2) < DIST2:
            continue
        rid = ring_s[c]
        b2 = b3 = DIST2
        m2 = m3 = -1
        for j in range(c + 1, len(surf_last)):
            if ring_s[j] > rid + NEARBY:
                break
            d = float(sqdis32(surf_last[j, :3], sel))
            if ring_s[j] <= rid and d < b2:
                b2, m2 = d, j
            elif ring_s[j] > rid and d < b3:
                b3, m3 = d, j
        for j in range(c - 1, -1, -1):
            if ring_s[j] < rid - NEARBY:
                break
            d = float(sqdis32(surf_last[j, :3], sel))
            if ring_s[j] >= rid and d < b2:
                b2, m2 = d, j
            elif ring_s[j] < rid and d < b3:
                b3, m3 = d, j
        pi[i] = (c, m2, m3)
    return ci, pi


def test_oracle_odometry_correspondences_equal_python_restatement(oracle):
    w = synth.make_world()
    rng = np.random.default_rng(5)
    regs = []
    for k in range(2):
        q, t = synth.loop_pose(w, 1.0 * k)
        regs.append(oracle.scan_register(synth.raycast_sweep(w, q, t, 64, 600, rng), 64, 5.0))
    prev, cur = regs
    q_lc = np.array([0.0, 0.0, np.sin(0.004), np.cos(0.004)])                     # a small guess, like para_q / para_t mid-solve
    t_lc = np.array([0.9, 0.02, -0.01])
    sharp, flat = cur["sharp"][:120], cur["flat"][:200]                            # Python loops: keep it to a few hundred features
    ci, pi = oracle.odom_associate(sharp, flat, prev["less_sharp"], prev["less_flat"], q_lc, t_lc)
    pci, ppi = py_associate(sharp, flat, prev["less_sharp"], prev["less_flat"], q_lc, t_lc)
    assert (ci[:, 1] >= 0).sum() > 60 and ((pi[:, 1] >= 0) & (pi[:, 2] >= 0)).sum() > 100      # the scenario does associate
    assert np.array_equal(ci, pci)
    assert np.array_equal(pi, ppi)
