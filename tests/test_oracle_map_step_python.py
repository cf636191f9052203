"""CPU test (-m "not gpu"): whole laserMapping passes of the oracle (Aloam/src/laserMapping.cpp:142-152, 307-801) against a
Python composition of independent pieces: cube bookkeeping restated here (centre cube, 5x5x3 valid list in i/j/k order,
insertion by cube index, per-cube VoxelGrid refilter of the valid cubes), numpy VoxelGrid, the association of
test_oracle_assoc_python and the numeric-Jacobian LM of test_oracle_lm_python."""
import numpy as np

import scenario
from test_oracle_assoc_python import associate_to_map, py_associate
from test_oracle_lm_python import py_lm, qrot
from test_oracle_odom_step_python import qmul
from test_oracle_primitives import np_voxel_grid

W, H, D = 21, 21, 11                                    # :77-79


def cube_coord(v, cen):                                 # :312-321 / :741-750
    c = int((float(v) + 25.0) / 50.0) + cen
    if float(v) + 25.0 < 0:
        c -= 1
    return c


class PyMapper:
    def __init__(self):
        self.cubes = [dict(), dict()]                   # linear cube index -> float32 [n, 4]
        self.cen = (10, 10, 5)                          # :74-76
        self.q_wm, self.t_wm = np.array([0.0, 0.0, 0.0, 1.0]), np.zeros(3)
        self.leaf = (0.4, 0.8)

    def insert(self, which, pts):
        for p in pts:
            i, j, k = (cube_coord(p[a], self.cen[a]) for a in range(3))
            if 0 <= i < W and 0 <= j < H and 0 <= k < D:
                self.cubes[which].setdefault(i + W * j + W * H * k, []).append(p)

    def filter(self, which, idx):
        if idx in self.cubes[which] and len(self.cubes[which][idx]) > 0:
            self.cubes[which][idx] = list(np_voxel_grid(np.array(self.cubes[which][idx], np.float32), self.leaf[which]))

    def import_points(self, which, pts):
        self.insert(which, pts)
        for idx in list(self.cubes[which]):
            self.filter(which, idx)

    def export_all(self, which):
        out = [np.array(self.cubes[which][i], np.float32) for i in sorted(self.cubes[which]) if len(self.cubes[which][i])]
        return np.concatenate(out) if out else np.zeros((0, 4), np.float32)

    def step(self, oracle, corner_last, surf_last, q_odom, t_odom):
        q_w = qmul(self.q_wm, q_odom)                                            # :142-146
        t_w = qrot(self.q_wm, t_odom) + self.t_wm
        c = [cube_coord(t_w[a], self.cen[a]) for a in range(3)]
        assert 3 <= c[0] < W - 3 and 3 <= c[1] < H - 3 and 3 <= c[2] < D - 3     # no shift in this scenario (:323-507)
        valid = [i + W * j + W * H * k
                 for i in range(c[0] - 2, c[0] + 3) for j in range(c[1] - 2, c[1] + 3) for k in range(c[2] - 1, c[2] + 2)
                 if 0 <= i < W and 0 <= j < H and 0 <= k < D]                    # :512-529
        maps = []
        for which in (0, 1):
            parts = [np.array(self.cubes[which][v], np.float32) for v in valid if len(self.cubes[which].get(v, []))]
            maps.append(np.concatenate(parts) if parts else np.zeros((0, 4), np.float32))
        cs = np_voxel_grid(corner_last, 0.4) if len(corner_last) else corner_last  # :542-550
        ss = np_voxel_grid(surf_last, 0.8) if len(surf_last) else surf_last
        counts = []
        if len(maps[0]) > 10 and len(maps[1]) > 50:                              # :554
            for _ in range(2):                                                   # :562
                fac, nc, ns = py_associate(maps[0], maps[1], cs, ss, q_w, t_w)
                counts.append((nc, ns))
                f = np.zeros(len(fac), oracle.FACTOR_DTYPE)
                for k, (ty, p, a, b) in enumerate(fac):
                    f["type"][k] = ty; f["p"][k] = p; f["a"][k] = a; f["b"][k] = b
                q_w, t_w, _ = py_lm(f, q_w, t_w, 4)                              # :713-720
        qi = np.r_[-q_odom[:3], q_odom[3]] / float(q_odom @ q_odom)              # :148-152
        self.q_wm = qmul(q_w, qi)
        self.t_wm = t_w - qrot(self.q_wm, t_odom)
        for which, st in ((0, cs), (1, ss)):                                     # :737-783
            if len(st):
                pw = np.concatenate([associate_to_map(q_w, t_w, st), st[:, 3:4]], 1).astype(np.float32)
                self.insert(which, pw)
        for v in valid:                                                          # :788-801
            self.filter(0, v)
            self.filter(1, v)
        return q_w, t_w, counts, (len(maps[0]), len(maps[1]))


def test_oracle_mapping_steps_equal_python_composition(oracle):
    cm, sm = scenario.small_map(half_xy=40.0, n_surf=60_000, n_corner=15_000)
    om, pm = oracle.Mapper(), PyMapper()
    for which, pts in ((0, cm), (1, sm)):
        om.import_points(which, pts)
        pm.import_points(which, pts)
        assert np.array_equal(om.export(which, 1), pm.export_all(which))
    for k, (c, s, q, t, qp, tp) in enumerate(scenario.sweeps(2, seed=21, n_corner=400, n_surf=1500, dt=0.1, drot=0.5)):
        c, s = c[:200], s[:600]                                                  # Python loops: a few hundred queries
        oq, ot, rep, _ = om.step(c, s, qp, tp)
        pq, pt, counts, nmap = pm.step(oracle, c, s, np.asarray(qp, np.float64), np.asarray(tp, np.float64))
        assert (rep.corner_from_map, rep.surf_from_map) == nmap and rep.optimized == 1
        assert [(rep.corner_num[i], rep.surf_num[i]) for i in range(2)] == counts, k
        assert counts[1][0] > 20 and counts[1][1] > 100
        assert np.abs(oq - pq).max() <= 1e-7 and np.abs(ot - pt).max() <= 1e-7, (k, ot, pt)
        for which in (0, 1):
            a, b = om.export(which, 1), pm.export_all(which)
            assert a.shape == b.shape and np.allclose(a, b, rtol=0, atol=2e-5), (k, which)


def py_shift_window(cubes, cen, t_w):
    """:312-507 restated on a dict {(i, j, k): points}: returns the new dict, the new cen and the centre cube."""
    dims = (W, H, D)
    cen = list(cen)
    c = [cube_coord(t_w[a], cen[a]) for a in range(3)]
    for a in range(3):
        while c[a] < 3:                                 # contents move up one cube, the top plane is recycled empty
            cubes = {tuple(k[b] + (1 if b == a else 0) for b in range(3)): v for k, v in cubes.items() if k[a] + 1 < dims[a]}
            c[a] += 1
            cen[a] += 1
        while c[a] >= dims[a] - 3:                      # contents move down one cube, plane 0 is recycled empty
            cubes = {tuple(k[b] - (1 if b == a else 0) for b in range(3)): v for k, v in cubes.items() if k[a] - 1 >= 0}
            c[a] -= 1
            cen[a] -= 1
    return cubes, tuple(cen), tuple(c)


def test_oracle_window_shift_equals_python_restatement(oracle):
    rng = np.random.default_rng(31)
    pts = np.zeros((6000, 4), np.float32)
    pts[:, :3] = rng.uniform(-180, 180, (6000, 3)) * np.array([1.0, 1.0, 0.5], np.float32)
    pts[:, 3] = np.arange(6000) % 50
    om = oracle.Mapper()
    om.import_points(1, pts)
    cubes, cen = {}, (10, 10, 5)
    for p in pts:                                        # import: by cube index, then VoxelGrid per cube
        key = tuple(cube_coord(p[a], cen[a]) for a in range(3))
        if all(0 <= key[a] < (W, H, D)[a] for a in range(3)):
            cubes.setdefault(key, []).append(p)
    cubes = {k: list(np_voxel_grid(np.array(v, np.float32), 0.8)) for k, v in cubes.items()}

    def flat(cs):
        keys = sorted(cs, key=lambda k: k[0] + W * k[1] + W * H * k[2])
        return np.concatenate([np.array(cs[k], np.float32) for k in keys]) if keys else np.zeros((0, 4), np.float32)

    assert np.array_equal(om.export(1, 1), flat(cubes))
    for t_w in ([420.0, 0.0, 0.0], [470.0, -380.0, 0.0], [470.0, -460.0, 130.0], [-300.0, 100.0, -160.0], [0.0, 0.0, 0.0],
                [760.0, 0.0, 0.0], [0.0, 0.0, 0.0]):
        om.prepare_window(t_w)
        cubes, cen, centre = py_shift_window(cubes, cen, t_w)
        _, _, ocen = om.get_state()
        assert tuple(ocen) == cen, (t_w, ocen, cen)
        assert np.array_equal(om.export(1, 1), flat(cubes)), t_w
        valid = [(i, j, k) for i in range(centre[0] - 2, centre[0] + 3) for j in range(centre[1] - 2, centre[1] + 3)
                 for k in range(centre[2] - 1, centre[2] + 2) if 0 <= i < W and 0 <= j < H and 0 <= k < D]
        win = [np.array(cubes[v], np.float32) for v in valid if v in cubes]
        win = np.concatenate(win) if win else np.zeros((0, 4), np.float32)
        assert np.array_equal(om.export(1, 0), win), t_w            # the :512-537 concatenation = index space of the kNN
    assert len(flat(cubes)) < len(pts)                               # some planes were recycled on the way
