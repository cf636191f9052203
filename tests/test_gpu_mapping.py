"""GPU parity tests of the scan-to-map path: every check calls the CUDA library through its
C ABI (lmono_b200.api -> liblmono_b200.so) and compares with the CPU oracle on the same
seeded inputs.  Bars (BASELINE.json north_star): voxel centroids, kNN indices / distances and
map contents bit-exact; 6x6 normal equations <= 1e-5 relative; poses <= 1e-4 m / 1e-4 rad."""
import numpy as np
import pytest

import scenario

pytestmark = pytest.mark.gpu


def rot_angle(qa, qb):
    d = abs(float(np.dot(qa, qb)))
    return 2.0 * np.arccos(min(1.0, d))


@pytest.mark.parametrize("leaf,n,span", [(0.4, 6000, 60.0), (0.8, 50000, 70.0), (0.2, 20000, 15.0), (0.8, 1, 1.0), (0.4, 0, 1.0)])
def test_voxel_grid_bit_exact(gpu_ctx_factory, oracle, leaf, n, span):
    ctx = gpu_ctx_factory()
    rng = np.random.default_rng(n + 1)
    pts = np.zeros((n, 4), np.float32)
    pts[:, :3] = rng.uniform(-span, span, (n, 3))
    pts[:, 2] *= 0.2
    pts[:, 3] = rng.uniform(0, 64, n)
    got = ctx.voxel_grid(pts, leaf)
    ref = oracle.voxel_grid(pts, leaf, 0)
    assert got.shape == ref.shape
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_voxel_grid_duplicates_and_boundaries(gpu_ctx_factory, oracle):
    """many points per voxel, points exactly on voxel borders, negative coordinates"""
    ctx = gpu_ctx_factory()
    rng = np.random.default_rng(5)
    base = rng.integers(-20, 20, (4000, 3)).astype(np.float32) * np.float32(0.4)
    jitter = rng.choice([0.0, 0.0, 0.1, 0.39], (4000, 3)).astype(np.float32)
    pts = np.zeros((4000, 4), np.float32)
    pts[:, :3] = base + jitter
    pts[:, 3] = np.arange(4000) % 7
    got = ctx.voxel_grid(pts, 0.4)
    ref = oracle.voxel_grid(pts, 0.4, 0)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


@pytest.fixture(scope="module")
def loaded(gpu_ctx_factory, oracle):
    cm, sm = scenario.small_map()
    ctx = gpu_ctx_factory()
    ctx.map_import(0, cm)
    ctx.map_import(1, sm)
    om = oracle.Mapper()
    om.import_points(0, cm)
    om.import_points(1, sm)
    return ctx, om


def test_import_export_bit_exact(loaded):
    ctx, om = loaded
    for which in (0, 1):
        got = ctx.map_export(which, 1)
        ref = om.export(which, 1)
        assert got.shape == ref.shape, (which, got.shape, ref.shape)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


@pytest.mark.parametrize("form", ["latency", "throughput"])
def test_knn5_bit_exact(loaded, form):
    """both forms of the search kernel (8 lanes per query / one thread per query) against the oracle's brute force"""
    ctx, om = loaded
    ctx.set_concurrency_hint(8 if form == "throughput" else 1)
    w = scenario.world()
    from lmono_b200 import synth
    q0, t0 = synth.loop_pose(w, 0.0)
    rng = np.random.default_rng(11)
    ctx.map_prepare_window(t0)
    om.prepare_window(t0)
    for which, n in ((0, 3000), (1, 12000)):
        ref_map = om.export(which, 0)
        got_map = ctx.map_export(which, 0)
        assert np.array_equal(ref_map.view(np.uint32), got_map.view(np.uint32))
        # queries: map points + noise (dense hits) and uniform points (many rejects)
        pick = rng.integers(0, len(ref_map), n)
        q = ref_map[pick].copy()
        q[:, :3] += rng.normal(0, 0.3, (n, 3)).astype(np.float32)
        q[: n // 10, :3] = (t0 + rng.uniform(-60, 60, (n // 10, 3))).astype(np.float32)
        gi, gd = ctx.knn5(which, q)
        ri, rd = om.knn5(which, q)
        accept = rd[:, 4] < 1.0
        assert accept.sum() > n // 4
        assert np.array_equal(gi[accept], ri[accept])
        assert np.array_equal(gd[accept].view(np.uint32), rd[accept].view(np.uint32))
        # rejected queries: the GPU must not report a 5th neighbour inside the gate
        rej = ~accept
        assert np.all(~(gd[rej, 4] < 1.0))
        # ties among the 6 nearest would make FLANN order-dependent: report, do not hide
        ties = (np.diff(rd[accept], axis=1) == 0).any(axis=1).sum()
        print(f"which={which} accepted={accept.sum()} rejected={rej.sum()} exact-distance ties in top-5={ties}")
    ctx.set_concurrency_hint(1)


@pytest.mark.parametrize("form", ["latency", "throughput"])
def test_knn5_dense_map_at_cube_corner(gpu_ctx_factory, oracle, form):
    """Worst case for the cell search: a dense 3-D lattice (one point per 0.4 m voxel: ~65 points within 1 m of a query,
    far more than the thread-per-query kernel parks in shared memory) around the corner (25, 25, 25) shared by 8 cubes,
    so that queries see cells straddling cube borders on all three axes (up to 27 cube/cell pairs, 18 runs)."""
    rng = np.random.default_rng(77)
    g = np.arange(19.0, 31.0, 0.4, dtype=np.float32) + np.float32(0.2)
    xx, yy, zz = np.meshgrid(g, g, g, indexing="ij")
    pts = np.stack([xx.ravel(), yy.ravel(), zz.ravel(), np.zeros(xx.size, np.float32)], 1).astype(np.float32)
    pts[:, :3] += rng.uniform(-0.15, 0.15, (len(pts), 3)).astype(np.float32)
    ctx = gpu_ctx_factory()
    om = oracle.Mapper()
    for which in (0, 1):
        ctx.map_import(which, pts)
        om.import_points(which, pts)
    ctx.set_concurrency_hint(8 if form == "throughput" else 1)
    t0 = np.array([25.0, 25.0, 25.0])
    ctx.map_prepare_window(t0)
    om.prepare_window(t0)
    n = 6000
    q = np.zeros((n, 4), np.float32)
    q[:, :3] = rng.uniform(21.0, 29.0, (n, 3)).astype(np.float32)
    q[: n // 3, :3] = (25.0 + rng.uniform(-1.2, 1.2, (n // 3, 3))).astype(np.float32)       # around the shared corner
    q[n // 3: n // 2, 0] = np.float32(25.0)                                                   # exactly on a cube border
    q[n // 2: 2 * n // 3, 1] = np.float32(24.99999)
    for which in (0, 1):
        assert np.array_equal(ctx.map_export(which, 0).view(np.uint32), om.export(which, 0).view(np.uint32))
        gi, gd = ctx.knn5(which, q)
        ri, rd = om.knn5(which, q)
        accept = rd[:, 4] < 1.0
        assert accept.sum() > n // 2
        assert np.array_equal(gi[accept], ri[accept])
        assert np.array_equal(gd[accept].view(np.uint32), rd[accept].view(np.uint32))
        assert np.all(~(gd[~accept, 4] < 1.0))


def test_knn_brute_matches_kdtree_here(loaded, oracle):
    _, om = loaded
    ref_map = om.export(1, 0)
    rng = np.random.default_rng(2)
    q = ref_map[rng.integers(0, len(ref_map), 500)].copy()
    q[:, :3] += rng.normal(0, 0.2, (500, 3)).astype(np.float32)
    ib, db = oracle.knn_brute(ref_map, q)
    ik, dk = oracle.knn_kdtree(ref_map, q)
    assert np.array_equal(db, dk)
    assert np.array_equal(ib, ik)


def test_normal_equations(loaded, oracle):
    ctx, om = loaded
    c, s, q, t, qp, tp = scenario.sweeps(1, seed=21)[0]
    cs = oracle.voxel_grid(c, 0.4, 0)
    ss = oracle.voxel_grid(s, 0.8, 0)
    H, g, cost, nc, ns = ctx.map_normal_eq(cs, ss, qp, tp)
    fac, rnc, rns = om.associate(cs, ss, qp, tp)
    Hr, gr, costr = oracle.normal_eq(fac, qp, tp)
    assert (nc, ns) == (rnc, rns)
    assert nc > 100 and ns > 1000
    scale = np.abs(Hr).max()
    assert np.abs(H - Hr).max() <= 1e-5 * scale, np.abs(H - Hr).max() / scale
    assert np.abs(g - gr).max() <= 1e-5 * np.abs(gr).max()
    assert abs(cost - costr) <= 1e-9 * costr
    print("H rel err", np.abs(H - Hr).max() / scale, "g rel err", np.abs(g - gr).max() / np.abs(gr).max())


def test_map_step_sequence(gpu_ctx_factory, oracle):
    """30 sweeps through lmono_map_step vs the oracle's process(): poses, counts, LM traces
    and the final map contents."""
    cm, sm = scenario.small_map()
    ctx = gpu_ctx_factory()
    ctx.map_import(0, cm)
    ctx.map_import(1, sm)
    om = oracle.Mapper()
    om.import_points(0, cm)
    om.import_points(1, sm)
    worst_t = worst_r = 0.0
    for k, (c, s, q, t, qp, tp) in enumerate(scenario.sweeps(30, seed=7, dt=0.1, drot=0.5)):
        gq, gt, grep, _ = ctx.map_step(c, s, qp, tp)
        rq, rt, rrep, _ = om.step(c, s, qp, tp)
        assert (grep.corner_from_map, grep.surf_from_map) == (rrep.corner_from_map, rrep.surf_from_map), k
        assert (grep.corner_stack, grep.surf_stack) == (rrep.corner_stack, rrep.surf_stack), k
        assert list(grep.corner_num) == list(rrep.corner_num), (k, list(grep.corner_num), list(rrep.corner_num))
        assert list(grep.surf_num) == list(rrep.surf_num), (k, list(grep.surf_num), list(rrep.surf_num))
        for it in range(2):
            assert grep.solve[it].iterations == rrep.solve[it].iterations
            assert grep.solve[it].termination == rrep.solve[it].termination
            assert abs(grep.solve[it].final_cost - rrep.solve[it].final_cost) <= 1e-7 * max(1.0, rrep.solve[it].final_cost)
        dt = float(np.linalg.norm(gt - rt))
        dr = rot_angle(gq, rq)
        worst_t, worst_r = max(worst_t, dt), max(worst_r, dr)
        assert dt <= 1e-4 and dr <= 1e-4, (k, dt, dr)
        # the registration itself must be good (not just equal to the oracle)
        assert np.linalg.norm(gt - t) < 0.05
    print("worst pose deviation vs oracle: %.3e m %.3e rad" % (worst_t, worst_r))
    for which in (0, 1):
        got = ctx.map_export(which, 1)
        ref = om.export(which, 1)
        assert got.shape == ref.shape
        same = (got.view(np.uint32) == ref.view(np.uint32)).all(axis=1).mean()
        print(f"map {which}: {len(got)} points, bit-identical rows {same:.6f}")
        # Rows that differ: the device pose equals the oracle's to ~1e-16 (different rounding inside the LM solve), so a
        # feature's world coordinate can land one fp32 ulp away; it then moves the centroid of the voxel it was merged
        # into by a few ulps.  Print how many rows and how far, and bound both.
        diff = ~(got.view(np.uint32) == ref.view(np.uint32)).all(axis=1)
        if diff.any():
            a32, b32 = got[diff, :3].view(np.int32).astype(np.int64), ref[diff, :3].view(np.int32).astype(np.int64)
            print(f"   {int(diff.sum())} rows differ: max |d| {np.abs(got[diff, :3] - ref[diff, :3]).max():.2e} m, max {int(np.abs(a32 - b32).max())} fp32 ulps in xyz")
        assert np.allclose(got, ref, rtol=0, atol=2e-5)
        assert same > 0.999


def test_full_res_registration(loaded, oracle):
    ctx, om = loaded
    c, s, q, t, qp, tp = scenario.sweeps(1, seed=33)[0]
    full = np.concatenate([c, s])[:5000]
    st = ctx.map_get_state()
    ctx.map_set_state([0, 0, 0, 1], [0, 0, 0])
    gq, gt, grep, greg = ctx.map_step(c, s, qp, tp, full_res=full)
    assert greg.shape == full.shape
    # same transform applied on host in double, stored as float
    from lmono_b200 import synth
    R = synth.quat_to_rot(gq)
    ref = (full[:, :3].astype(np.float64) @ R.T + gt).astype(np.float32)
    assert np.abs(greg[:, :3] - ref).max() < 2e-5
    assert np.array_equal(greg[:, 3], full[:, 3])


@pytest.mark.parametrize("form", ["latency", "throughput"])
def test_map_step_empty_and_degenerate_inputs(gpu_ctx_factory, oracle, form):
    """Edge cases of one laserMapping pass against the oracle, step by step on the same state: nothing at all, a first
    sweep into an empty map (laserMapping.cpp:554 gate fails: no optimisation, the sweep only fills the map), a sweep
    without corner features, a sweep without any feature, and a sweep whose features find no neighbour within 1 m
    (:584,652 gates reject every query: Ceres gets zero residual blocks and leaves the pose at the prior)."""
    e = np.zeros((0, 4), np.float32)
    sw = scenario.sweeps(3, seed=3)
    ctx = gpu_ctx_factory()
    ctx.set_concurrency_hint(8 if form == "throughput" else 1)
    om = oracle.Mapper()

    def both(c, s, q, t, tag):
        gq, gt, grep, _ = ctx.map_step(c, s, q, t)
        rq, rt, rrep, _ = om.step(c, s, q, t)
        assert (grep.optimized, grep.corner_from_map, grep.surf_from_map, grep.corner_stack, grep.surf_stack) == \
               (rrep.optimized, rrep.corner_from_map, rrep.surf_from_map, rrep.corner_stack, rrep.surf_stack), tag
        assert list(grep.corner_num) == list(rrep.corner_num) and list(grep.surf_num) == list(rrep.surf_num), tag
        assert np.linalg.norm(gt - rt) <= 1e-4 and rot_angle(gq, rq) <= 1e-4, (tag, gt, rt)
        return grep

    ident = (np.array([0.0, 0.0, 0.0, 1.0]), np.zeros(3))
    rep = both(e, e, *ident, "nothing")
    assert rep.optimized == 0
    c, s, q, t, qp, tp = sw[0]
    rep = both(c, s, q, t, "first sweep into an empty map")
    assert rep.optimized == 0 and rep.corner_stack > 0
    for which in (0, 1):
        assert np.array_equal(ctx.map_export(which, 1).view(np.uint32), om.export(which, 1).view(np.uint32)), which
    # the rest against a loaded map (fresh state on both sides)
    cm, sm = scenario.small_map()
    ctx = gpu_ctx_factory()
    ctx.set_concurrency_hint(8 if form == "throughput" else 1)
    om = oracle.Mapper()
    for which, pts in ((0, cm), (1, sm)):
        ctx.map_import(which, pts)
        om.import_points(which, pts)
    c, s, q, t, qp, tp = sw[1]
    rep = both(e, s, qp, tp, "no corner features")
    assert rep.optimized == 1 and list(rep.corner_num) == [0, 0] and rep.surf_num[1] > 1000
    rep = both(e, e, qp, tp, "no features")
    assert list(rep.surf_num) == [0, 0]
    cf, sf = c.copy(), s.copy()
    cf[:, 0] += np.float32(300.0)
    sf[:, 0] += np.float32(300.0)
    rep = both(cf, sf, qp, tp, "no neighbour within the gate")
    assert rep.optimized == 1 and list(rep.corner_num) == [0, 0] and list(rep.surf_num) == [0, 0]
    for which in (0, 1):
        assert np.array_equal(ctx.map_export(which, 1).view(np.uint32), om.export(which, 1).view(np.uint32)), which
    # those points sit, unfiltered and in arrival order, in cubes outside the window.  Move the window over them: the cubes
    # are indexed for the search and re-voxelised (the reference's VoxelGrid over prefix + appended points)
    from lmono_b200 import synth
    t_far = tp + synth.quat_to_rot(qp) @ np.array([250.0, 0.0, 0.0])       # the sweep was displaced along the sensor's x axis
    rep = both(e, e, qp, t_far, "window moves over the appended cubes")
    assert rep.corner_from_map > 0 and rep.surf_from_map > 0
    for which in (0, 1):
        assert np.array_equal(ctx.map_export(which, 0).view(np.uint32), om.export(which, 0).view(np.uint32)), which
        assert np.array_equal(ctx.map_export(which, 1).view(np.uint32), om.export(which, 1).view(np.uint32)), which
    # a second displaced sweep lands on the now filtered cubes (tail merge on top of the re-voxelised prefix)
    # a second displaced sweep (the window is back at the start): appended, in arrival order, behind the filtered prefix
    rep = both(cf, sf, qp, tp, "second sweep into the same cubes")
    for which in (0, 1):
        assert np.array_equal(ctx.map_export(which, 1).view(np.uint32), om.export(which, 1).view(np.uint32)), which
    # and back over them: VoxelGrid over (filtered prefix ++ appended points)
    rep = both(e, e, qp, t_far, "window over the cubes again")
    for which in (0, 1):
        assert np.array_equal(ctx.map_export(which, 0).view(np.uint32), om.export(which, 0).view(np.uint32)), which
        assert np.array_equal(ctx.map_export(which, 1).view(np.uint32), om.export(which, 1).view(np.uint32)), which


def test_whole_slab_revoxelisation_of_a_full_cube(gpu_ctx_factory, oracle):
    """A cube that already holds ~27 k filtered surf points (more than the shared-memory sort of the whole-slab path
    takes) collects a sweep while it is outside the window; when the window reaches it, it is re-voxelised as a whole
    through the global-memory scratch: pcl::VoxelGrid over prefix ++ appended points, bit for bit."""
    from lmono_b200 import synth
    e = np.zeros((0, 4), np.float32)
    c, s, q, t, qp, tp = scenario.sweeps(3, seed=3)[1]
    R = synth.quat_to_rot(qp)
    far = tp + R @ np.array([300.0, 0.0, 0.0])
    centre = np.round(far / 50.0) * 50.0
    g = np.arange(-12.0, 12.0, 0.8, dtype=np.float32) + np.float32(0.4)
    xx, yy, zz = np.meshgrid(g, g, g, indexing="ij")
    blob = np.stack([xx.ravel(), yy.ravel(), zz.ravel(), np.zeros(xx.size, np.float32)], 1).astype(np.float32)
    blob[:, :3] += centre.astype(np.float32)
    assert len(blob) > 16384
    cm, sm = scenario.small_map()
    sm2 = np.concatenate([sm, blob])
    ctx = gpu_ctx_factory()
    om = oracle.Mapper()
    for which, pts in ((0, cm), (1, sm2)):
        ctx.map_import(which, pts)
        om.import_points(which, pts)
    cf, sf = c.copy(), s.copy()
    cf[:, 0] += np.float32(300.0)
    sf[:, 0] += np.float32(300.0)
    for (a, b, tt, tag) in ((cf, sf, tp, "sweep into cubes outside the window"), (e, e, tp + R @ np.array([250.0, 0.0, 0.0]), "window over them")):
        gq, gt, grep, _ = ctx.map_step(a, b, qp, tt)
        rq, rt, rrep, _ = om.step(a, b, qp, tt)
        assert (grep.optimized, grep.corner_from_map, grep.surf_from_map) == (rrep.optimized, rrep.corner_from_map, rrep.surf_from_map), tag
        assert np.linalg.norm(gt - rt) <= 1e-4, tag
    assert rrep.surf_from_map > 16384
    for which in (0, 1):
        assert np.array_equal(ctx.map_export(which, 0).view(np.uint32), om.export(which, 0).view(np.uint32)), which
        assert np.array_equal(ctx.map_export(which, 1).view(np.uint32), om.export(which, 1).view(np.uint32)), which


def test_export_interleaved_like_the_reference_publishers(gpu_ctx_factory):
    """lmono_map_export(which = 2): corner cube, surf cube, corner cube, ... in the cube order of the publishers
    (laserMapping.cpp:808-816 over laserCloudSurroundInd, :826-830 over all 4851 cubes)."""
    rng = np.random.default_rng(9)
    def cloud(n, tag):
        p = np.zeros((n, 4), np.float32)
        p[:, :3] = rng.uniform([-180, -180, -40], [180, 180, 40], (n, 3))
        p[:, 3] = tag                       # the VoxelGrid centroid of equal intensities keeps the tag
        return p
    ctx = gpu_ctx_factory()
    ctx.map_import(0, cloud(30_000, 1.0))
    ctx.map_import(1, cloud(80_000, 2.0))
    ctx.map_prepare_window(np.array([30.0, -20.0, 5.0]))
    q, t, cen = ctx.map_get_state()
    for scope in (0, 1):
        c, s, both = ctx.map_export(0, scope), ctx.map_export(1, scope), ctx.map_export(2, scope)
        assert len(both) == len(c) + len(s) and len(c) > 0 and len(s) > 0
        assert np.array_equal(both[both[:, 3] == 1.0].view(np.uint32), c.view(np.uint32))
        assert np.array_equal(both[both[:, 3] == 2.0].view(np.uint32), s.view(np.uint32))
        g = np.floor((both[:, :3].astype(np.float64) + 25.0) / 50.0).astype(np.int64) + np.array(cen)
        if scope == 1:
            cube = g[:, 0] + 21 * g[:, 1] + 441 * g[:, 2]            # :826-830 linear index
        else:
            cube = (g[:, 0] * 64 + g[:, 1]) * 64 + g[:, 2]            # :512-529 loop order: i outer, j, k inner
        key = cube * 4 + both[:, 3].astype(np.int64)
        assert np.all(np.diff(key) >= 0), scope                      # grouped by cube, corner before surf inside a cube
        assert len(np.unique(cube)) > 5


@pytest.mark.gpu
def test_pool_exhaustion_is_reported_and_eviction_recovers(gpu_ctx_factory):
    """The reference's cube clouds are unbounded, the slab pools are not: with a pool of 12 slabs per map a drive that
    touches more cubes raises LMONO_FAULT_POOL_EXHAUSTED (pose still valid, new points of unplaced cubes dropped);
    lmono_map_evict gives the far cubes' slabs back and the next registrations place their points again."""
    from lmono_b200 import api
    ctx = gpu_ctx_factory(max_cubes_corner=12, max_cubes_surf=12)
    rng = np.random.default_rng(3)

    def blob(n=4000):                            # sensor-frame features around the sensor
        p = np.zeros((n, 4), np.float32)
        p[:, 0] = rng.uniform(-20, 20, n); p[:, 1] = rng.uniform(-20, 20, n); p[:, 2] = rng.uniform(-2, 2, n)
        return p

    # walk along x in 50 m steps: each step touches new cubes; nothing is ever optimised (empty map), points are inserted
    faulted = False
    for k in range(40):
        x = 50.0 * k
        try:
            ctx.map_step(blob(500), blob(), [0, 0, 0, 1], [x, 0, 0])
        except api.LmonoError as e:
            assert e.code == -5, e
            faulted = True
            break
    assert faulted
    freed = ctx.map_evict(4)
    assert freed > 0
    for k2 in range(k + 1, k + 6):              # after the eviction new cubes can be placed again
        x = 50.0 * k2
        ctx.map_step(blob(500), blob(), [0, 0, 0, 1], [x, 0, 0])
    n_near = len(ctx.map_export(1, 0))
    assert n_near > 0
