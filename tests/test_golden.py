"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py).

The reference has no tests / fixtures for this path.  Two kinds of stored vectors: seeded inputs +
ORACLE outputs (make_golden.py: every stage and primitive), and seeded inputs + outputs of the REFERENCE'S OWN CODE
compiled from /root/reference (ref_nodes.npz, make_golden_ref.py; second half of this file).  The CPU half
(-m "not gpu") checks that the oracle still reproduces them; the GPU half (-m gpu) checks the CUDA library through
the C ABI against the same stored outputs -- no oracle involved on that side."""
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name + ".npz"))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


# ----------------------------------------------------------------------------- oracle vs golden (CPU)
def test_oracle_primitives_golden(oracle):
    g = load("primitives")
    assert np.array_equal(bits(oracle.voxel_grid(g["vg_in"], 0.4)), bits(g["vg_out_04"]))
    assert np.array_equal(bits(oracle.voxel_grid(g["vg_in"], 0.8)), bits(g["vg_out_08"]))
    for fn in (oracle.knn_brute, oracle.knn_kdtree):
        idx, d2 = fn(g["knn_cloud"], g["knn_queries"], 5)
        assert np.array_equal(idx, g["knn_idx"]) and np.array_equal(bits(d2), bits(g["knn_d2"]))
    for cov, w, V, A, x in zip(g["eig_cov"], g["eig_w"], g["eig_V"], g["qr_A"], g["qr_x"]):
        w2, V2, rc = oracle.eigh3(cov)
        assert rc == 0 and np.array_equal(w2, w) and np.array_equal(V2, V)
        assert np.array_equal(oracle.colpiv_solve_5x3(A, -np.ones(5)), x)
        # and the stored answers are right: independent numpy check
        assert np.allclose(np.linalg.eigvalsh(cov), w, rtol=1e-10, atol=1e-14)
        assert np.allclose(np.linalg.lstsq(A, -np.ones(5), rcond=None)[0], x, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("n_scans,min_range", [(64, 5.0), (32, 0.3), (16, 0.3)])
def test_oracle_scan_golden(oracle, n_scans, min_range):
    g = load("scan_registration")
    k = f"s{n_scans}_"
    r = oracle.scan_register(g[k + "raw"], n_scans, min_range)
    _check_scan(r, g, k)


def _check_scan(r, g, k, gpu=False):
    rep = r["report"]
    assert [rep.n_in, rep.n_kept, rep.n_sharp, rep.n_less_sharp, rep.n_flat, rep.n_less_flat] == list(g[k + "counts"])
    assert np.array_equal(np.array(list(rep.ring_start)), g[k + "ring_start"])
    assert np.array_equal(np.array(list(rep.ring_end)), g[k + "ring_end"])
    assert np.array_equal(r["labels"].astype(np.int8), g[k + "labels"])
    assert np.array_equal(r["src_index"], g[k + "src_index"])
    assert np.array_equal(bits(r["curvature"]), bits(g[k + "curvature"]))
    assert np.array_equal(bits(r["less_flat"][:, :3]), bits(g[k + "less_flat"][:, :3]))
    # relTime goes through atan2f: the integer ring part is exact, the fraction is compared at 4e-6 (+ ulps after averaging)
    assert np.array_equal(np.floor(r["full"][:, 3]), np.floor(g[k + "full_intensity"]))
    tol = 4e-6 if gpu else 0.0
    dI = np.abs(r["full"][:, 3] - g[k + "full_intensity"])
    assert dI.max() <= tol, (dI.max(), int(dI.argmax()), r["full"][dI.argmax()], g[k + "full_intensity"][dI.argmax()], rep.start_ori, rep.end_ori)
    assert np.all(np.abs(r["less_flat"][:, 3] - g[k + "less_flat"][:, 3]) <= tol + (2.0 * np.spacing(np.abs(g[k + "less_flat"][:, 3])) if gpu else 0.0))


def test_oracle_odometry_golden(oracle):
    g = load("odometry")
    od = oracle.Odometry()
    for k in range(3):
        (lq, lt), (wq, wt), rep = od.step(g[f"f{k}_sharp"], g[f"f{k}_less_sharp"], g[f"f{k}_flat"], g[f"f{k}_less_flat"])
        assert [rep.inited, rep.corner_corr[0], rep.corner_corr[1], rep.plane_corr[0], rep.plane_corr[1]] == list(g[f"f{k}_corr"])
        assert np.array_equal(np.concatenate([lq, lt]), g[f"f{k}_last_curr"])
        assert np.array_equal(np.concatenate([wq, wt]), g[f"f{k}_w_curr"])
    ci, pi = oracle.odom_associate(g["f1_sharp"], g["f1_flat"], g["f0_less_sharp"], g["f0_less_flat"], (0, 0, 0, 1), (0, 0, 0))
    assert np.array_equal(ci, g["assoc01_corner"]) and np.array_equal(pi, g["assoc01_plane"])


def _counts(rep):
    return [rep.corner_from_map, rep.surf_from_map, rep.corner_stack, rep.surf_stack, rep.corner_num[0], rep.corner_num[1],
            rep.surf_num[0], rep.surf_num[1], rep.optimized, rep.solve[0].iterations, rep.solve[1].iterations,
            rep.solve[0].num_successful, rep.solve[1].num_successful]


def test_oracle_mapping_golden(oracle):
    g = load("mapping")
    om = oracle.Mapper()
    om.import_points(0, g["map_corner_in"]); om.import_points(1, g["map_surf_in"])
    om.prepare_window(g["ne_pose"][4:])
    idx, d2 = om.knn5(1, g["knn_queries_world"])
    assert np.array_equal(idx, g["knn_idx"]) and np.array_equal(bits(d2), bits(g["knn_d2"]))
    fac, nc, ns = om.associate(g["ne_corner_stack"], g["ne_surf_stack"], g["ne_pose"][:4], g["ne_pose"][4:])
    H, gg, cost = oracle.normal_eq(fac, g["ne_pose"][:4], g["ne_pose"][4:])
    assert np.array_equal(H, g["ne_H"]) and np.array_equal(gg, g["ne_g"]) and [cost, nc, ns] == list(g["ne_cost_nc_ns"])
    for k in range(3):
        o = g[f"k{k}_odom"]
        q, t, rep, _ = om.step(g[f"k{k}_corner"], g[f"k{k}_surf"], o[:4], o[4:])
        assert _counts(rep) == list(g[f"k{k}_counts"])
        assert np.array_equal(np.concatenate([q, t]), g[f"k{k}_w_curr"])
    assert np.array_equal(bits(om.export(0, 1)), bits(g["map_corner_out"]))
    assert np.array_equal(bits(om.export(1, 1)), bits(g["map_surf_out"]))


def _cam(oracle_or_api, g, cls):
    fx, fy, cx, cy, W, H = g["cam"]
    return cls(fx, fy, cx, cy, 0.0, 0.0, 0.0, 0.0, int(W), int(H), 0, 5, 0)


def test_oracle_color_golden(oracle):
    g = load("color")
    cam = _cam(oracle, g, oracle.Camera)
    raw = oracle.project_raster(g["pts_cam"], cam)
    assert np.array_equal(raw, g["depth_raw"])
    fill = oracle.depth_fill(raw, cam)
    assert np.array_equal(fill, g["depth_filled"])
    cc, cw, rgb = oracle.lift_cloud(fill, g["bgr"], cam, g["pose"][:4], g["pose"][4:])
    assert np.array_equal(bits(cc), bits(g["cloud_cam"])) and np.array_equal(bits(cw), bits(g["cloud_world"]))
    assert np.array_equal(rgb, g["rgb"])


# ----------------------------------------------------------------------------- CUDA path vs golden (GPU)
@pytest.mark.gpu
def test_gpu_voxel_grid_golden(gpu_ctx_factory):
    g = load("primitives")
    ctx = gpu_ctx_factory()
    assert np.array_equal(bits(ctx.voxel_grid(g["vg_in"], 0.4)), bits(g["vg_out_04"]))
    assert np.array_equal(bits(ctx.voxel_grid(g["vg_in"], 0.8)), bits(g["vg_out_08"]))


@pytest.mark.gpu
@pytest.mark.parametrize("n_scans,min_range", [(64, 5.0), (32, 0.3), (16, 0.3)])
def test_gpu_scan_golden(gpu_ctx_factory, n_scans, min_range):
    g = load("scan_registration")
    ctx = gpu_ctx_factory(scan_line=n_scans, minimum_range=min_range)
    _check_scan(ctx.scan_register(g[f"s{n_scans}_raw"], want_debug=True), g, f"s{n_scans}_", gpu=True)


@pytest.mark.gpu
def test_gpu_odometry_golden(gpu_ctx_factory):
    g = load("odometry")
    ctx = gpu_ctx_factory()
    for k in range(3):
        (lq, lt), (wq, wt), rep = ctx.odom_step(g[f"f{k}_sharp"], g[f"f{k}_less_sharp"], g[f"f{k}_flat"], g[f"f{k}_less_flat"])
        assert [rep.inited, rep.corner_corr[0], rep.corner_corr[1], rep.plane_corr[0], rep.plane_corr[1]] == list(g[f"f{k}_corr"])
        assert np.linalg.norm(lt - g[f"f{k}_last_curr"][4:]) <= 1e-4 and np.linalg.norm(wt - g[f"f{k}_w_curr"][4:]) <= 1e-4
        assert 2 * np.arccos(min(1.0, abs(float(np.dot(wq, g[f"f{k}_w_curr"][:4]))))) <= 1e-4
        if k == 1:
            ci, pi = ctx.odom_debug(0, len(g["f1_sharp"]), len(g["f1_flat"]))
            assert np.array_equal(ci, g["assoc01_corner"]) and np.array_equal(pi, g["assoc01_plane"])


@pytest.mark.gpu
def test_gpu_mapping_golden(gpu_ctx_factory):
    g = load("mapping")
    ctx = gpu_ctx_factory()
    ctx.map_import(0, g["map_corner_in"]); ctx.map_import(1, g["map_surf_in"])
    ctx.map_prepare_window(g["ne_pose"][4:])
    idx, d2 = ctx.knn5(1, g["knn_queries_world"])
    ok = g["knn_d2"][:, 4] < 1.0                       # contract of the grid search: exact wherever the d2[4] < 1 gate passes
    assert ok.sum() > 1000
    assert np.array_equal(idx[ok], g["knn_idx"][ok]) and np.array_equal(bits(d2[ok]), bits(g["knn_d2"][ok]))
    H, gg, cost, nc, ns = ctx.map_normal_eq(g["ne_corner_stack"], g["ne_surf_stack"], g["ne_pose"][:4], g["ne_pose"][4:])
    assert [nc, ns] == list(g["ne_cost_nc_ns"][1:])
    scale = np.abs(g["ne_H"]).max()
    assert np.abs(H - g["ne_H"]).max() <= 1e-5 * scale and np.abs(gg - g["ne_g"]).max() <= 1e-5 * np.abs(g["ne_g"]).max()
    assert abs(cost - g["ne_cost_nc_ns"][0]) <= 1e-5 * g["ne_cost_nc_ns"][0]
    for k in range(3):
        o = g[f"k{k}_odom"]
        q, t, rep, _ = ctx.map_step(g[f"k{k}_corner"], g[f"k{k}_surf"], o[:4], o[4:])
        assert _counts(rep) == list(g[f"k{k}_counts"])
        assert np.linalg.norm(t - g[f"k{k}_w_curr"][4:]) <= 1e-4
        assert 2 * np.arccos(min(1.0, abs(float(np.dot(q, g[f"k{k}_w_curr"][:4]))))) <= 1e-4
    for which, key in ((0, "map_corner_out"), (1, "map_surf_out")):
        got = ctx.map_export(which, 1)
        assert got.shape == g[key].shape and np.allclose(got, g[key], rtol=0, atol=2e-5)


@pytest.mark.gpu
def test_gpu_color_golden(gpu_ctx_factory):
    from lmono_b200 import api
    g = load("color")
    fx, fy, cx, cy, W, H = g["cam"]
    ctx = gpu_ctx_factory(image_width=int(W), image_height=int(H))
    cam = api.Pinhole(fx, fy, cx, cy, 0.0, 0.0, 0.0, 0.0, int(W), int(H), 0, 5, 0)
    r = ctx.project_color(g["pts_cam"], g["bgr"], cam, g["pose"][:4], g["pose"][4:])
    assert np.array_equal(r["depth_raw"], g["depth_raw"]) and np.array_equal(r["depth"], g["depth_filled"])
    assert np.array_equal(bits(r["cloud_cam"]), bits(g["cloud_cam"])) and np.array_equal(bits(r["cloud_world"]), bits(g["cloud_world"]))
    assert np.array_equal(r["rgb"], g["rgb"])


# ============================================================================= vectors generated by the REFERENCE's own code
# tests/golden/ref_nodes.npz (tests/golden/make_golden_ref.py): outputs of the reference's scanRegistration / laserOdometry /
# laserMapping nodes and colour mapper as compiled from /root/reference into oracle/_ref.  Committed, so these checks need
# neither the reference tree nor the prebuilt libraries.
def _ang(q0, q1):
    return 2 * np.arccos(min(1.0, abs(float(np.dot(q0, q1)))))


def _check_scan_ref(got, g, tag, exact_intensity):
    assert got["full"].shape == g[f"{tag}_full"].shape
    assert np.array_equal(bits(got["full"][:, :3]), bits(g[f"{tag}_full"][:, :3]))
    assert np.array_equal(np.floor(got["full"][:, 3]), np.floor(g[f"{tag}_full"][:, 3]))
    assert np.abs(got["full"][:, 3] - g[f"{tag}_full"][:, 3]).max() <= (0.0 if exact_intensity else 4e-6)
    assert np.array_equal(bits(got["curvature"][5:-5]), bits(g[f"{tag}_curvature"][5:-5]))
    assert np.array_equal(np.asarray(got["labels"], np.int8), g[f"{tag}_labels"])
    for k in ("sharp", "less_sharp", "flat"):
        assert got[k].shape == g[f"{tag}_{k}"].shape and len(got[k]) > 0, k
        assert np.array_equal(bits(got[k][:, :3]), bits(g[f"{tag}_{k}"][:, :3])), k
    assert got["less_flat"].shape == g[f"{tag}_less_flat"].shape
    assert np.abs(got["less_flat"] - g[f"{tag}_less_flat"]).max() <= 3e-5          # voxel members summed in another order (introsort vs input order)


@pytest.mark.parametrize("tag", ["s64", "s16"])
def test_oracle_scan_reference_golden(oracle, tag):
    g = load("ref_nodes")
    n_scans, min_range = int(g[f"{tag}_cfg"][0]), float(g[f"{tag}_cfg"][1])
    _check_scan_ref(oracle.scan_register(g[f"{tag}_raw"], n_scans, min_range), g, tag, exact_intensity=True)
    r = oracle.scan_register(g[f"{tag}_raw"], n_scans, min_range, voxel_order_mode=1, sort_mode=1)        # the toolchain's own sorts: everything bit for bit
    assert np.array_equal(bits(r["less_flat"]), bits(g[f"{tag}_less_flat"]))


def test_oracle_odometry_reference_golden(oracle):
    g = load("ref_nodes")
    od = oracle.Odometry()
    for k in range(6):
        (lq, lt), (wq, wt), rep = od.step(*(g[f"od{k}_{n}"] for n in ("sharp", "less_sharp", "flat", "less_flat")))
        p = g["od_poses"][k]
        assert [rep.corner_corr[1], rep.plane_corr[1]] == [int(p[14]), int(p[15])], k
        assert np.abs(np.concatenate([lq, lt, wq, wt]) - p[:14]).max() <= 1e-12, k
    assert np.linalg.norm(g["od_poses"][5][11:14]) > 3.0
    od.close()


def test_oracle_mapping_reference_golden(oracle):
    g = load("ref_nodes")
    om = oracle.Mapper(order_mode=1, use_kdtree=1)
    for k in range(6):
        p = g["mp_poses"][k]
        q, t, rep, reg = om.step(g[f"mp{k}_corner"], g[f"mp{k}_surf"], p[:4], p[4:7], g[f"mp{k}_full"])
        wq, wt, cen = om.get_state()
        assert np.abs(np.concatenate([q, t]) - p[7:14]).max() <= 1e-11 and np.abs(np.concatenate([wq, wt]) - p[14:21]).max() <= 1e-10, k
        assert cen == [int(v) for v in p[21:24]]
        assert np.array_equal(bits(reg), bits(g[f"mp{k}_registered"])), k
    for which, key in ((0, "mp_map_corner"), (1, "mp_map_surf")):
        got = om.export(which, 1)
        assert got.shape == g[key].shape and np.array_equal(bits(got), bits(g[key])), key
    om.close()


def _ref_cam(g):
    c = g["col_cam"]
    return c, dict(fx=c[0], fy=c[1], cx=c[2], cy=c[3], k1=c[4], k2=c[5], p1=c[6], p2=c[7], width=int(c[8]), height=int(c[9]),
                   kernel_type=int(c[10]), kernel_size=int(c[11]), blur_type=int(c[12]))


def test_oracle_color_reference_golden(oracle):
    g = load("ref_nodes")
    _, kw = _ref_cam(g)
    cam = oracle.make_camera(**kw)
    raw = oracle.project_raster(g["col_pts"], cam)
    assert np.array_equal(raw, g["col_raw"])
    fill = oracle.depth_fill(raw, cam)
    assert np.array_equal(fill, g["col_filled"])
    cc, cw, rgb = oracle.lift_cloud(fill, g["col_img"], cam, g["col_pose"][:4], g["col_pose"][4:])
    assert len(cc) > 1000 and np.array_equal(bits(cc), bits(g["col_cloud_cam"])) and np.array_equal(bits(cw), bits(g["col_cloud_world"]))
    assert np.array_equal(rgb, g["col_rgb"])


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["s64", "s16"])
def test_gpu_scan_reference_golden(gpu_ctx_factory, tag):
    g = load("ref_nodes")
    ctx = gpu_ctx_factory(scan_line=int(g[f"{tag}_cfg"][0]), minimum_range=float(g[f"{tag}_cfg"][1]))
    _check_scan_ref(ctx.scan_register(g[f"{tag}_raw"], want_debug=True), g, tag, exact_intensity=False)


@pytest.mark.gpu
def test_gpu_odometry_reference_golden(gpu_ctx_factory):
    g = load("ref_nodes")
    ctx = gpu_ctx_factory(scan_line=16, minimum_range=0.3, max_cubes_corner=8, max_cubes_surf=8, cube_capacity_corner=1024, cube_capacity_surf=1024)
    for k in range(6):
        (lq, lt), (wq, wt), rep = ctx.odom_step(*(g[f"od{k}_{n}"] for n in ("sharp", "less_sharp", "flat", "less_flat")))
        p = g["od_poses"][k]
        assert [rep.corner_corr[1], rep.plane_corr[1]] == [int(p[14]), int(p[15])], k
        assert np.linalg.norm(lt - p[4:7]) <= 1e-4 and np.linalg.norm(wt - p[11:14]) <= 1e-4 and _ang(lq, p[0:4]) <= 1e-4 and _ang(wq, p[7:11]) <= 1e-4, k


@pytest.mark.gpu
def test_gpu_mapping_reference_golden(gpu_ctx_factory):
    g = load("ref_nodes")
    ctx = gpu_ctx_factory()
    for k in range(6):
        p = g["mp_poses"][k]
        q, t, rep, reg = ctx.map_step(g[f"mp{k}_corner"], g[f"mp{k}_surf"], p[:4], p[4:7], full_res=g[f"mp{k}_full"])
        assert np.linalg.norm(t - p[11:14]) <= 1e-4 and _ang(q, p[7:11]) <= 1e-4, k
        assert list(rep.cen) == [int(v) for v in p[21:24]]
        assert np.abs(reg - g[f"mp{k}_registered"]).max() <= 1e-4, k
    for which, key in ((0, "mp_map_corner"), (1, "mp_map_surf")):
        got = ctx.map_export(which, 1)
        assert got.shape == g[key].shape and np.abs(got - g[key]).max() <= 1e-4, key


@pytest.mark.gpu
def test_gpu_color_reference_golden(gpu_ctx_factory):
    from lmono_b200 import api
    g = load("ref_nodes")
    c, kw = _ref_cam(g)
    ctx = gpu_ctx_factory(image_width=kw["width"], image_height=kw["height"])
    cam = api.Pinhole(c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], kw["width"], kw["height"], kw["kernel_type"], kw["kernel_size"], kw["blur_type"])
    r = ctx.project_color(g["col_pts"], g["col_img"], cam, g["col_pose"][:4], g["col_pose"][4:])
    assert np.array_equal(r["depth_raw"], g["col_raw"]) and np.array_equal(r["depth"], g["col_filled"])
    assert np.array_equal(bits(r["cloud_cam"]), bits(g["col_cloud_cam"])) and np.array_equal(bits(r["cloud_world"]), bits(g["col_cloud_world"]))
    assert np.array_equal(r["rgb"], g["col_rgb"])
