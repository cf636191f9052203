"""CPU test (-m "not gpu"): the oracle's scan-to-map association (C: KD-tree 5-NN, hand-written 3x3 eigen solver and 5x3
column-pivoting QR) against an independent Python restatement of Aloam/src/laserMapping.cpp:154-163, 577-687 built on
numpy (float32 brute-force 5-NN, numpy.linalg.eigh, numpy.linalg.lstsq).  The accepted factor sets must be identical
(every gate -- d2[4] < 1, lambda2 > 3 lambda1, |n.p + d| <= 0.2 -- falls the same way) and the factor parameters agree
to 1e-9 (eigenvector sign aside: a <-> b)."""
import numpy as np

import scenario

f32 = np.float32


def associate_to_map(q, t, pts):
    """pointAssociateToMap (:154-163): Eigen q * v in double + t, stored float"""
    qv, w = np.asarray(q[:3], np.float64), float(q[3])
    out = np.zeros((len(pts), 3), np.float32)
    for i, p in enumerate(pts[:, :3].astype(np.float64)):
        uv = 2.0 * np.cross(qv, p)
        out[i] = (p + w * uv + np.cross(qv, uv) + np.asarray(t, np.float64)).astype(np.float32)
    return out


def knn5(mp, sel):
    d = mp[:, :3] - sel[None, :]
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]                 # float32, FLANN L2_Simple order
    idx = np.lexsort((np.arange(len(d2)), d2))[:5]                                    # (d2, index) ascending
    return idx, d2[idx]


def py_associate(cmap, smap, cstack, sstack, q, t):
    out = []
    for ori, sel in zip(cstack, associate_to_map(q, t, cstack)):                      # :577-622
        idx, d2 = knn5(cmap, sel)
        if not float(d2[4]) < 1.0:
            continue
        nb = cmap[idx, :3].astype(np.float64)
        center = nb.sum(0) / 5.0
        z = nb - center
        w, v = np.linalg.eigh(z.T @ z)                                                # ascending eigenvalues
        if w[2] > 3 * w[1]:
            out.append((0, ori[:3].astype(np.float64), center + 0.1 * v[:, 2], center - 0.1 * v[:, 2]))
    nc = len(out)
    for ori, sel in zip(sstack, associate_to_map(q, t, sstack)):                      # :643-687
        idx, d2 = knn5(smap, sel)
        if not float(d2[4]) < 1.0:
            continue
        A = smap[idx, :3].astype(np.float64)
        n = np.linalg.lstsq(A, -np.ones(5), rcond=None)[0]
        d = 1.0 / np.linalg.norm(n)
        n = n / np.linalg.norm(n)
        if np.all(np.abs(A @ n + d) <= 0.2):
            out.append((2, ori[:3].astype(np.float64), n, np.array([d, 0.0, 0.0])))
    return out, nc, len(out) - nc


def test_oracle_map_association_equals_python_restatement(oracle):
    cm, sm = scenario.small_map(half_xy=40.0, n_surf=60_000, n_corner=15_000)
    om = oracle.Mapper()
    om.import_points(0, cm)
    om.import_points(1, sm)
    c, s, q, t, qp, tp = scenario.sweeps(1, seed=21, n_corner=400, n_surf=1500)[0]
    cs, ss = oracle.voxel_grid(c, 0.4, 0)[:150], oracle.voxel_grid(s, 0.8, 0)[:400]   # a few hundred queries: Python loops
    om.prepare_window(tp)
    fac, nc, ns = om.associate(cs, ss, qp, tp)
    ref, rnc, rns = py_associate(om.export(0, 0), om.export(1, 0), cs, ss, qp, tp)
    assert (nc, ns) == (rnc, rns) and nc > 20 and ns > 80
    assert len(fac) == len(ref)
    for f, (ty, p, a, b) in zip(fac, ref):
        assert (f["type"] == 0) == (ty == 0)
        assert np.array_equal(f["p"], p)
        if ty == 0:       # line through the 5 neighbours: a / b = centre +- 0.1 * principal direction (sign free)
            same = np.abs(f["a"] - a).max() <= 1e-9 and np.abs(f["b"] - b).max() <= 1e-9
            swapped = np.abs(f["a"] - b).max() <= 1e-9 and np.abs(f["b"] - a).max() <= 1e-9
            assert same or swapped
        else:             # plane: unit normal and offset
            assert np.abs(f["a"] - a).max() <= 1e-9 and abs(f["b"][0] - b[0]) <= 1e-9
