"""CPU tests (-m "not gpu") of the oracle's stage restatements: scanRegistration, laserOdometry
and laserMapping behave like A-LOAM on synthetic sweeps (sanity + invariants; the reference
has no tests of its own to pin against)."""
import numpy as np
import pytest

import scenario
from lmono_b200 import synth


@pytest.fixture(scope="module")
def sweeps64(oracle):
    w = synth.make_world()
    rng = np.random.default_rng(1)
    out = []
    for k in range(4):
        q, t = synth.loop_pose(w, 1.0 * k)
        raw = synth.raycast_sweep(w, q, t, 64, 1875, rng)
        out.append((raw, oracle.scan_register(raw, 64, 5.0), q, t))
    return out


def test_scan_registration_structure(sweeps64):
    raw, r, _, _ = sweeps64[0]
    rep = r["report"]
    assert rep.n_in == len(raw) and 0 < rep.n_kept < rep.n_in
    rings = np.floor(r["full"][:, 3]).astype(int)
    assert np.all(np.diff(rings) >= 0)                     # ring-major
    assert rings.max() <= 50                               # this fork keeps rings 0..50 (scanRegistration.cpp:195)
    assert rep.n_sharp <= 51 * 12 and rep.n_flat <= 51 * 24 and rep.n_less_sharp <= 51 * 120
    lab = r["labels"]
    assert set(np.unique(lab)) <= {-1, 0, 1, 2}
    assert (lab == 2).sum() == rep.n_sharp and (lab >= 1).sum() == rep.n_less_sharp and (lab == -1).sum() == rep.n_flat
    # source indices are a stable partition of the kept input points
    src = r["src_index"]
    for ring in np.unique(rings)[:5]:
        assert np.all(np.diff(src[rings == ring]) > 0)
    assert np.array_equal(raw[src, :3], r["full"][:, :3])
    # near points were removed
    assert np.all(np.linalg.norm(r["full"][:, :3], axis=1) >= 5.0 - 1e-4)


def test_scan_registration_sort_modes_agree_without_ties(oracle, sweeps64):
    raw = sweeps64[1][0]
    a = oracle.scan_register(raw, 64, 5.0, sort_mode=0)
    b = oracle.scan_register(raw, 64, 5.0, sort_mode=1)
    c = a["curvature"][5:-5]
    if len(np.unique(c)) == len(c):                         # no exact curvature ties: std::sort order is unique
        assert np.array_equal(a["labels"], b["labels"])
    else:
        assert (a["labels"] != b["labels"]).mean() < 1e-3


def test_odometry_tracks_motion(oracle, sweeps64):
    od = oracle.Odometry()
    for k, (_, r, q, t) in enumerate(sweeps64):
        (lq, lt), (wq, wt), rep = od.step(r["sharp"], r["less_sharp"], r["flat"], r["less_flat"])
        if k == 0:
            assert rep.inited == 0
        else:
            assert rep.corner_corr[1] > 300 and rep.plane_corr[1] > 800
    assert abs(np.linalg.norm(lt) - 1.0) < 0.1


def test_mapping_registers_against_imported_map(oracle):
    cm, sm = scenario.small_map(half_xy=60.0, n_surf=150_000, n_corner=40_000)
    m = oracle.Mapper()
    m.import_points(0, cm)
    m.import_points(1, sm)
    for (c, s, q, t, qp, tp) in scenario.sweeps(3, seed=5, n_corner=1500, n_surf=8000):
        m.set_state([0, 0, 0, 1], [0, 0, 0])
        eq, et, rep, _ = m.step(c, s, qp, tp)
        assert rep.optimized == 1
        assert np.linalg.norm(et - t) < 0.03 < np.linalg.norm(tp - t) + 0.03


def test_mapping_window_shift_keeps_points(oracle):
    """Drive far enough that the 21x21x11 window shifts; cubes that stay inside keep their points."""
    m = oracle.Mapper()
    rng = np.random.default_rng(0)
    pts = np.zeros((5000, 4), np.float32)
    pts[:, :3] = rng.uniform(-40, 40, (5000, 3))
    m.import_points(1, pts)
    before = m.export(1, 1)
    n_valid = m.prepare_window([400.0, 0.0, 0.0])       # centre cube 8 -> still inside, no shift yet
    _, _, cen0 = m.get_state()
    m.prepare_window([460.0, 0.0, 0.0])                 # centre cube would be 19 >= 18 -> shift
    _, _, cen1 = m.get_state()
    assert cen1[0] < cen0[0]
    after = m.export(1, 1)
    assert len(after) == len(before)
    assert np.array_equal(np.sort(after.view(np.uint32), axis=0), np.sort(before.view(np.uint32), axis=0))
    m.prepare_window([2000.0, 0.0, 0.0])                # far away: everything scrolled out
    assert len(m.export(1, 1)) == 0


def test_mapping_cubes_outside_the_window_keep_arrival_order_until_they_enter_it(oracle):
    """laserMapping.cpp:737-801: every feature is pushed into its cube, but only the cubes of laserCloudValidInd are
    re-filtered.  A cube outside the 5x5x3 window therefore keeps its new points appended in arrival order, over as many
    sweeps as it takes, and is voxel-filtered (prefix ++ appended, summed in that order) the first time the window
    covers it.  Independent restatement: numpy VoxelGrid of test_oracle_primitives."""
    from test_oracle_primitives import np_voxel_grid
    rng = np.random.default_rng(9)
    e = np.zeros((0, 4), np.float32)

    def batch(n):
        p = np.zeros((n, 4), np.float32)
        p[:, 0] = rng.uniform(130.0, 145.0, n)          # cube +3 in x: outside the window around the origin
        p[:, 1:3] = rng.uniform(-10.0, 10.0, (n, 2))
        p[:, 3] = rng.uniform(0, 50, n)
        return p

    m = oracle.Mapper()
    ident = ([0, 0, 0, 1], [0, 0, 0])
    s1, s2 = batch(3000), batch(2500)
    _, _, rep, _ = m.step(e, s1, *ident)
    assert rep.optimized == 0                             # empty map: the sweep only fills it, pose = prior
    v1 = np_voxel_grid(s1, 0.8)
    assert np.array_equal(m.export(1, 1), v1)             # appended in the order of the (voxel-filtered) sweep, not re-filtered
    m.set_state(*ident)
    m.step(e, s2, *ident)
    v2 = np_voxel_grid(s2, 0.8)
    assert np.array_equal(m.export(1, 1), np.concatenate([v1, v2]))
    m.set_state(*ident)
    _, _, rep, _ = m.step(e, e, [0, 0, 0, 1], [150.0, 0.0, 0.0])      # the window now covers the cube
    assert rep.surf_from_map == len(v1) + len(v2)         # the search sees prefix ++ appended ...
    assert np.array_equal(m.export(1, 1), np_voxel_grid(np.concatenate([v1, v2]), 0.8))    # ... and the refilter merges them
