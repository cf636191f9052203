"""CPU tests (-m "not gpu"): the oracle against the REFERENCE'S OWN SOURCES compiled in this container.

`make -C oracle ref` compiles translation units of /root/reference where they lie (a driver #includes them; nothing is
copied) against the functional stand-ins of oracle/refstubs/ for ROS / PCL / Ceres / Eigen, into oracle/_ref/*.so.
Everything between the library calls is reference code built by this toolchain: these tests are what pins the oracle.
They run wherever oracle/_ref exists (built here; the files travel to the GPU box with the snapshot) and skip otherwise."""
import numpy as np
import pytest

import oracle_lib
from lmono_b200 import synth


def _need(name):
    if oracle_lib.ref_lib(name) is None:
        pytest.skip(f"oracle/_ref/libref_{name}.so not built (no reference tree here)")


def _bits(a):
    return a.view(np.uint32) if a.dtype == np.float32 else a


# scanRegistration.cpp:113-459.  The node's std::sort and libm calls are this toolchain's, so the oracle is asked for the
# same ones (sort_mode 1 = libstdc++ std::sort on the curvature comparator, voxel_order_mode 1 = std::sort on the voxel
# index alone inside VoxelGrid); `atan` / `sqrt` at :166 resolve to the double overloads here, the oracle's default.
@pytest.mark.parametrize("n_scans,min_range,azimuths,seed", [(16, 0.3, 900, 3), (32, 0.3, 1875, 4), (64, 5.0, 1875, 5), (64, 5.0, 1875, 6)])
def test_scan_registration_oracle_equals_reference_source(oracle, n_scans, min_range, azimuths, seed):
    _need("scanreg")
    w = synth.make_world()
    rng = np.random.default_rng(seed)
    q, t = synth.loop_pose(w, 2.0 * seed)
    raw = synth.raycast_sweep(w, q, t, n_scans, azimuths, rng)
    raw[7, 0] = np.nan                                                     # removeNaNFromPointCloud :136
    raw[11, :3] = 0.01                                                     # removeClosedPointCloud :137
    ref = oracle_lib.ref_scan_register(raw, n_scans, min_range)
    ora = oracle.scan_register(raw, n_scans, min_range, voxel_order_mode=1, sort_mode=1)
    assert len(ref["full"]) > 10000
    for k in ("full", "curvature", "labels", "sharp", "less_sharp", "flat", "less_flat"):
        assert ref[k].shape == ora[k].shape, k
        assert np.array_equal(_bits(ref[k]), _bits(ora[k])), k             # every coordinate, ring id, relTime, curvature bit and label
    assert len(ref["sharp"]) > 0 and len(ref["flat"]) > 0 and len(ref["less_flat"]) > 1000
    # the canonical orders the CUDA path implements (stable sort; voxel members summed in input order) differ from the
    # toolchain's introsort only where the reference is not well defined: same labels, same picks, centroids within fp32 summation order
    can = oracle.scan_register(raw, n_scans, min_range)
    for k in ("full", "curvature", "labels", "sharp", "less_sharp", "flat"):
        assert np.array_equal(_bits(ref[k]), _bits(can[k])), k
    assert ref["less_flat"].shape == can["less_flat"].shape
    assert np.abs(ref["less_flat"] - can["less_flat"]).max() <= 3e-5              # a few fp32 ulps at 50-100 m (summation order of the voxel members)


# lidarFactor.hpp: LidarEdgeFactor / LidarPlaneNormFactor evaluated on dual numbers (the autodiff stand-in) inside a problem
# built the way the reference builds it, against the oracle's analytic residual blocks: normal equations, iteration
# trace, termination and final pose.  (LidarPlaneFactor takes three points, not a stored normal: it is covered by the
# odometry sequence below.)
def test_factors_and_solve_oracle_equals_reference_functors(oracle):
    _need("odom")
    from test_oracle_primitives import _random_factors
    rng = np.random.default_rng(17)
    for trial in range(4):
        f = _random_factors(oracle, rng, 200)
        f = f[f["type"] != 1]
        q0 = np.array([0.01, -0.006, 0.008, 1.0]) * np.r_[rng.uniform(0.5, 1.5, 3), 1.0]
        q0 /= np.linalg.norm(q0)
        t0 = rng.uniform(-0.08, 0.08, 3)
        H, g, c = oracle.normal_eq(f, q0, t0)
        rH, rg, rc = oracle_lib.ref_normal_eq(f, q0, t0)
        assert np.abs(H - rH).max() <= 1e-13 * np.abs(H).max() and np.abs(g - rg).max() <= 1e-13 * np.abs(g).max() and abs(c - rc) <= 1e-13 * c
        q, t, s = oracle.lm_solve(f, q0, t0, 4)
        rq, rt, rs = oracle_lib.ref_lm_solve(f, q0, t0, 4)
        assert (s.iterations, s.num_successful, s.termination) == (rs.iterations, rs.num_successful, rs.termination), trial
        assert np.abs(q - rq).max() <= 1e-14 and np.abs(t - rt).max() <= 1e-14
        assert abs(s.final_cost - rs.final_cost) <= 1e-12 * s.final_cost


# the same with interpolation ratios s in (0, 1] on the edge factors: lidarFactor.hpp:27-34, q_last_curr = Identity.slerp(s, q),
# t_last_curr = s t -- the factor arithmetic of `#define DISTORTION 1`, evaluated by the reference functor on dual numbers
def test_interpolated_edge_factors_oracle_equals_reference_functors(oracle):
    _need("odom")
    from test_oracle_primitives import _random_factors
    rng = np.random.default_rng(23)
    for trial in range(3):
        f = _random_factors(oracle, rng, 200)
        f = f[f["type"] == 0]
        f["s"] = rng.uniform(0.05, 1.0, len(f))
        q0 = np.array([0.03, -0.02, 0.025, 1.0]) * np.r_[rng.uniform(0.5, 1.5, 3), 1.0]
        q0 /= np.linalg.norm(q0)
        t0 = rng.uniform(-0.2, 0.2, 3)
        H, g, c = oracle.normal_eq(f, q0, t0)
        rH, rg, rc = oracle_lib.ref_normal_eq(f, q0, t0)
        assert np.abs(H - rH).max() <= 1e-13 * np.abs(H).max() and np.abs(g - rg).max() <= 1e-13 * np.abs(g).max() and abs(c - rc) <= 1e-13 * c
        q, t, s_ = oracle.lm_solve(f, q0, t0, 4)
        rq, rt, rs = oracle_lib.ref_lm_solve(f, q0, t0, 4)
        assert (s_.iterations, s_.num_successful, s_.termination) == (rs.iterations, rs.num_successful, rs.termination), trial
        assert np.abs(q - rq).max() <= 1e-14 and np.abs(t - rt).max() <= 1e-14


# laserOdometry.cpp:220-598, the node's own loop over consecutive HDL-32 sweeps: TransformToStart, the two
# correspondence searches with their ring-window scans, factor construction (LidarEdgeFactor / LidarPlaneFactor), two
# solves per sweep, pose accumulation, buffer swap
def test_odometry_sequence_oracle_equals_reference_source(oracle):
    _need("odom")
    w = synth.make_world()
    rng = np.random.default_rng(5)
    ref = oracle_lib.RefOdometry()
    od = oracle.Odometry()
    moved = 0.0
    for k in range(8):
        q, t = synth.loop_pose(w, 1.0 * k)
        raw = synth.raycast_sweep(w, q, t, 32, 900, rng)
        f = oracle.scan_register(raw, 32, 0.3)
        (lq, lt), (wq, wt), rep = od.step(f["sharp"], f["less_sharp"], f["flat"], f["less_flat"])
        (rlq, rlt), (rwq, rwt), cnt = ref.step(f["sharp"], f["less_sharp"], f["flat"], f["less_flat"], f["full"])
        assert [int(cnt[0]), int(cnt[1])] == [rep.corner_corr[1], rep.plane_corr[1]], k         # correspondences of the last pass
        if k:
            assert cnt[0] > 100 and cnt[1] > 300
        assert np.abs(lq - rlq).max() <= 1e-14 and np.abs(lt - rlt).max() <= 1e-13, k
        assert np.abs(wq - rwq).max() <= 1e-14 and np.abs(wt - rwt).max() <= 1e-13, k
        moved = float(np.linalg.norm(wt))
    assert moved > 5.0
    od.close()


# laserMapping.cpp:232-903, the node's own process() over consecutive sweeps from an empty map: transformAssociateToMap,
# window bookkeeping, the 5-NN association with the eigenvalue / plane-distance gates, LidarEdgeFactor /
# LidarPlaneNormFactor construction, two solves, transformUpdate, insertion, per-cube VoxelGrid refilter, the registered cloud
def test_mapping_sequence_oracle_equals_reference_source(oracle):
    _need("mapping")
    import scenario
    ref = oracle_lib.RefMapper()
    om = oracle.Mapper(order_mode=1, use_kdtree=1)          # the stand-in VoxelGrid sorts with this toolchain's std::sort, like PCL would
    n_opt = 0
    for k, (c, s, qg, tg, qo, to) in enumerate(scenario.sweeps(10, n_corner=1500, n_surf=9000)):
        full = np.concatenate([c, s])[:6000]
        rq, rt, (rwq, rwt), rcen, rfull = ref.step(c, s, qo, to, full)
        q, t, rep, ofull = om.step(c, s, qo, to, full)
        wq, wt, cen = om.get_state()
        assert rcen == cen, k
        assert np.abs(q - rq).max() <= 1e-13 and np.abs(t - rt).max() <= 1e-12, (k, q - rq, t - rt)
        assert np.abs(wq - rwq).max() <= 1e-13 and np.abs(wt - rwt).max() <= 1e-11, k
        assert np.array_equal(rfull.view(np.uint32), ofull.view(np.uint32)), k                    # /velodyne_cloud_registered
        for which in (0, 1):
            a, b = ref.export(which), om.export(which, 1)
            assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), (k, which)   # every cube, bit for bit
        n_opt += int(rep.solve[0].iterations > 0)
    assert n_opt >= 8
    om.close()


def _sparse(seed, n, lo, hi):
    rng = np.random.default_rng(seed)
    p = np.zeros((n, 4), np.float32)
    p[:, :3] = rng.uniform(lo, hi, (n, 3))
    return p


# laserMapping.cpp:323-507: the six window-shift loops.  The sensor hops out to +-600 m and back along x and y, down to
# -360 m and up to +360 m along z, then along a diagonal; every hop inserts a sparse sweep.  cen, both maps and the pose
# are compared after every hop.
def test_mapping_window_shifts_oracle_equals_reference_source(oracle):
    _need("mapping")
    ref = oracle_lib.RefMapper()
    om = oracle.Mapper(order_mode=1, use_kdtree=1)
    legs = []
    for axis, reach in ((0, 600.0), (1, 600.0), (2, 360.0)):
        for sgn in (1, -1):
            out = [sgn * 40.0 * k for k in range(1, int(reach / 40) + 1)]
            for v in out + out[-2::-1] + [0.0]:
                t = np.zeros(3)
                t[axis] = v
                legs.append(t)
    legs += [np.array([55.0 * k, -47.0 * k, 31.0 * k]) for k in list(range(1, 10)) + list(range(8, -1, -1))]
    q = np.array([0.0, 0.0, 0.0, 1.0])
    cens = set()
    for h, t in enumerate(legs):
        c = _sparse(1000 + h, 300, [-70, -70, -30], [70, 70, 30])
        s = _sparse(2000 + h, 900, [-70, -70, -30], [70, 70, 30])
        rq, rt, _, rcen, _ = ref.step(c, s, q, t)
        oq, ot, rep, _ = om.step(c, s, q, t)
        _, _, cen = om.get_state()
        assert rcen == cen, h
        cens.add(tuple(cen))
        assert np.array_equal(rq, oq) and np.array_equal(rt, ot), h
        for which in (0, 1):
            a, b = ref.export(which), om.export(which, 1)
            assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), (h, which)
    lo = np.min(np.array(list(cens)), axis=0)
    hi = np.max(np.array(list(cens)), axis=0)
    assert (lo < [10, 10, 5]).all() and (hi > [10, 10, 5]).all()       # the window moved in both directions on every axis
    om.close()


# Map_Builder.cc:213-334 MapBuilder::associateToMap (with depthFill :336-403 and Point3DTo2D :405-416): the projection raster
# with its implicit conversions and last-writer rule, depthFill's own sequence of morphology / blur calls for the three
# kernel types and both blur types, the per-pixel lift (depth window, |x| > 20 && y > 1.8 rule), the world transform --
# with and without lens distortion
@pytest.mark.parametrize("kernel_type,kernel_size,blur_type,dist", [(0, 5, 0, False), (1, 5, 1, False), (2, 7, 0, False), (0, 5, 1, True)])
def test_colour_mapper_oracle_equals_reference_source(oracle, kernel_type, kernel_size, blur_type, dist):
    _need("color")
    kw = dict(k1=-0.05, k2=0.01, p1=0.001, p2=-0.0005) if dist else {}
    cam = oracle.make_camera(kernel_type=kernel_type, kernel_size=kernel_size, blur_type=blur_type, **kw)
    rng = np.random.default_rng(12 + kernel_type)
    n = 120_000
    pts = np.zeros((n, 3), np.float32)
    pts[:, 2] = rng.uniform(-5, 130, n)                                       # behind the camera, and beyond 100 m (the uchar wraps)
    pts[:, 0] = rng.uniform(-1, 1, n) * pts[:, 2]
    pts[:, 1] = rng.uniform(-0.35, 0.35, n) * pts[:, 2]
    pts[:5, 2] = 0.0
    img = rng.integers(0, 255, (cam.height, cam.width, 3)).astype(np.uint8)
    q = np.array([0.1, -0.2, 0.05, 0.97])
    q /= np.linalg.norm(q)
    t = np.array([3.0, -1.0, 0.5])
    raw, filled, cc, cw, rgb = oracle_lib.ref_color_frame(pts, img, cam, q, t)
    oraw = oracle.project_raster(pts, cam)
    assert (raw > 0).sum() > 50000 and np.array_equal(raw, oraw)
    ofill = oracle.depth_fill(oraw, cam)
    assert np.array_equal(filled, ofill)
    occ, ocw, orgb = oracle.lift_cloud(ofill, img, cam, q, t)
    assert cc.shape == occ.shape and len(cc) > 10000
    assert np.array_equal(cc.view(np.uint32), occ.view(np.uint32)) and np.array_equal(cw.view(np.uint32), ocw.view(np.uint32)) and np.array_equal(rgb, orgb)


def _rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def colour_node_case(seed=3, n=120_000):
    """lidar-frame cloud, KITTI-like lidar-to-camera extrinsic, image, camera pose; T = [rlc^T | -rlc^T tlc] (map_build_node.cc:216-221)"""
    rng = np.random.default_rng(seed)
    pc = np.zeros((n, 3), np.float32)
    pc[:, 2] = rng.uniform(-5, 110, n)
    pc[:, 0] = rng.uniform(-1, 1, n) * pc[:, 2] * 0.95
    pc[:, 1] = rng.uniform(-0.3, 0.3, n) * pc[:, 2]
    q_lc = np.array([0.5, -0.5, 0.5, 0.5]) + np.array([0.01, -0.02, 0.015, 0.0])
    q_lc /= np.linalg.norm(q_lc)
    t_lc = np.array([0.27, -0.08, -0.05])
    rlc = _rot(q_lc)
    pl = ((rlc @ pc.astype(np.float64).T).T + t_lc).astype(np.float32)
    img = rng.integers(0, 255, (376, 1241, 3)).astype(np.uint8)
    q = np.array([0.01, -0.02, 0.3, 0.95])
    q /= np.linalg.norm(q)
    t = np.array([10.0, -3.0, 1.5])
    T = np.concatenate([rlc.T, (-rlc.T @ t_lc)[:, None]], axis=1)
    return pl, img, q_lc, t_lc, q, t, T


# map_build_node.cc:73-238: the colour-map NODE -- extrinsicHandler, the three buffers and their synchronisation in
# process(), T = [rlc^T | -rlc^T tlc], the cloud transform, then MapBuilder::associateToMap (rows D1-D4 in one pass)
def test_colour_node_oracle_equals_reference_source(oracle):
    _need("color")
    cam = oracle.make_camera()
    pl, img, q_lc, t_lc, q, t, T = colour_node_case()
    raw, filled, cc, cw, rgb = oracle_lib.ref_mapnode_frame(pl, img, cam, q_lc, t_lc, q, t)
    oraw = oracle.project_raster(oracle.transform_cloud(pl, T), cam)
    assert (raw > 0).sum() > 50000 and np.array_equal(raw, oraw)
    ofill = oracle.depth_fill(oraw, cam)
    assert np.array_equal(filled, ofill)
    occ, ocw, orgb = oracle.lift_cloud(ofill, img, cam, q, t)
    assert cc.shape == occ.shape and len(cc) > 100000
    assert np.array_equal(cc.view(np.uint32), occ.view(np.uint32)) and np.array_equal(cw.view(np.uint32), ocw.view(np.uint32)) and np.array_equal(rgb, orgb)


# ---------------------------------------------------------------------------------------------------- edge cases
def _edge_sweeps():
    w = synth.make_world()
    rng = np.random.default_rng(1)
    q, t = synth.loop_pose(w, 4.0)
    raw = synth.raycast_sweep(w, q, t, 64, 1875, rng)
    az = np.arctan2(raw[:, 1], raw[:, 0])
    el = np.degrees(np.arctan2(raw[:, 2], np.hypot(raw[:, 0], raw[:, 1])))
    keep = np.ones(len(raw), bool)
    keep[np.where((el > -3) & (el < -1))[0][5:]] = False
    raw16 = synth.raycast_sweep(w, q, t, 16, 1800, rng)
    lifted = raw16.copy()
    lifted[::50, 2] += 30.0
    raw32 = synth.raycast_sweep(w, q, t, 32, 1875, rng)
    return {
        "partial azimuth (start / end orientation fix-ups)": (raw[(az > -1.0) & (az < 1.2)], 64, 5.0),
        "every 7th point (wide gaps)": (raw[::7], 64, 5.0),
        "rings with fewer than six points (:279)": (raw[keep], 64, 5.0),
        "reversed point order": (raw[::-1].copy(), 64, 5.0),
        "elevations outside the ring rule (:169-200)": (lifted, 16, 0.3),
        "sweep starting mid-revolution (halfPassed)": (np.roll(raw32, 20000, axis=0), 32, 0.3),
        "duplicated points (zero gaps, curvature ties)": (np.concatenate([raw16[:3000], raw16[:3000]]), 16, 0.3),
    }


@pytest.mark.parametrize("case", list(_edge_sweeps().keys()) if oracle_lib.ref_lib("scanreg") is not None else [])
def test_scan_registration_edge_cases_oracle_equals_reference_source(oracle, case):
    raw, n_scans, min_range = _edge_sweeps()[case]
    ref = oracle_lib.ref_scan_register(raw, n_scans, min_range)
    ora = oracle.scan_register(raw, n_scans, min_range, voxel_order_mode=1, sort_mode=1)
    assert len(ref["full"]) > 1000
    for k in ("full", "labels", "sharp", "less_sharp", "flat", "less_flat"):
        assert ref[k].shape == ora[k].shape and np.array_equal(_bits(ref[k]), _bits(ora[k])), (case, k)
    assert np.array_equal(_bits(ref["curvature"][5:-5]), _bits(ora["curvature"][5:-5]))


def test_odometry_few_and_no_features_oracle_equals_reference_source(oracle):
    _need("odom")
    w = synth.make_world()
    rng = np.random.default_rng(2)
    ref = oracle_lib.RefOdometry()
    od = oracle.Odometry()
    for k in range(5):
        q, t = synth.loop_pose(w, 1.0 * k)
        f = oracle.scan_register(synth.raycast_sweep(w, q, t, 16, 900, rng), 16, 0.3)
        ns, nf = (len(f["sharp"]), len(f["flat"])) if k < 2 else ((6, 9) if k < 4 else (0, 0))       # "less correspondence!" (:487-490), then an empty problem
        a = (f["sharp"][:ns], f["less_sharp"], f["flat"][:nf], f["less_flat"])
        (lq, lt), (wq, wt), rep = od.step(*a)
        (rlq, rlt), (rwq, rwt), cnt = ref.step(*a, f["full"])
        assert [int(cnt[0]), int(cnt[1])] == [rep.corner_corr[1], rep.plane_corr[1]], k
        assert np.abs(np.r_[lq, lt, wq, wt] - np.r_[rlq, rlt, rwq, rwt]).max() <= 1e-12, k
    od.close()


def test_mapping_sparse_map_and_outside_grid_oracle_equals_reference_source(oracle):
    """:554 gate (fewer than 10 corner / 50 surf map points: no optimisation), then a pose 3 km away and 400 m down: every
    point of the sweep falls outside the 21 x 21 x 11 grid (:752-757) and the centre index runs far out of [3, dim - 4]"""
    _need("mapping")
    import scenario
    rm = oracle_lib.RefMapper()
    om = oracle.Mapper(order_mode=1, use_kdtree=1)
    opt = []
    for k, (c, s, qg, tg, qo, to) in enumerate(scenario.sweeps(6, n_corner=1500, n_surf=9000)):
        if k < 2:
            c, s = c[:8], s[:40]
        if k == 4:
            to = to + np.array([3000.0, 0.0, 0.0])
        if k == 5:
            to = to + np.array([0.0, 0.0, -400.0])
        rq, rt, _, rcen, _ = rm.step(c, s, qo, to)
        q, t, rep, _ = om.step(c, s, qo, to)
        assert rcen == om.get_state()[2], k
        assert np.abs(np.r_[q, t] - np.r_[rq, rt]).max() <= 1e-11, k
        for which in (0, 1):
            a, b = rm.export(which), om.export(which, 1)
            assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), (k, which)
        opt.append(int(rep.optimized))
    assert opt == [0, 0, 1, 1, 0, 0]
    om.close()
