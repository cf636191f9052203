"""GPU parity tests of the colour projection (rows D1-D4) against the oracle restatement of
map_build_node.cc:216-225 and Map_Builder.cc:213-416; the depthFill stages are additionally
cross-checked against cv2 in test_oracle_color.py."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _scene(n=120_000, seed=5):
    rng = np.random.default_rng(seed)
    pts = np.zeros((n, 3), np.float32)
    pts[:, 2] = rng.uniform(-5, 110, n)          # some behind the camera, some beyond 100 m
    pts[:, 0] = rng.uniform(-1, 1, n) * pts[:, 2] * 0.95
    pts[:, 1] = rng.uniform(-0.3, 0.3, n) * pts[:, 2]
    img = rng.integers(0, 256, (376, 1241, 3), dtype=np.uint8)
    return pts, img


@pytest.mark.parametrize("kernel_type,blur_type,dist", [(0, 0, False), (1, 1, False), (2, 0, True)])
def test_project_color_matches_oracle(gpu_ctx_factory, oracle, kernel_type, blur_type, dist):
    from lmono_b200 import api
    ctx = gpu_ctx_factory()
    pts, img = _scene()
    kw = dict(k1=-0.05, k2=0.01, p1=0.001, p2=-0.002) if dist else {}
    ocam = oracle.make_camera(kernel_type=kernel_type, blur_type=blur_type, **kw)
    gcam = api.Pinhole(ocam.fx, ocam.fy, ocam.cx, ocam.cy, ocam.k1, ocam.k2, ocam.p1, ocam.p2, ocam.width, ocam.height,
                       kernel_type, 5, blur_type)
    q = np.array([0.01, -0.02, 0.3, 0.95]); q /= np.linalg.norm(q)
    t = np.array([10.0, -3.0, 1.5])
    got = ctx.project_color(pts, img, gcam, q, t)
    raw = oracle.project_raster(pts, ocam)
    assert np.array_equal(got["depth_raw"], raw)
    fill = oracle.depth_fill(raw, ocam)
    assert np.array_equal(got["depth"], fill)
    cc, cw, rgb = oracle.lift_cloud(fill, img, ocam, q, t)
    assert got["cloud_cam"].shape == cc.shape and len(cc) > 10000
    assert np.array_equal(got["cloud_cam"].view(np.uint32), cc.view(np.uint32))
    assert np.array_equal(got["cloud_world"].view(np.uint32), cw.view(np.uint32))
    assert np.array_equal(got["rgb"], rgb)


def test_extrinsic_transform_and_empty(gpu_ctx_factory, oracle):
    from lmono_b200 import api
    ctx = gpu_ctx_factory()
    pts, img = _scene(20_000, 9)
    ocam = oracle.make_camera()
    gcam = api.Pinhole(ocam.fx, ocam.fy, ocam.cx, ocam.cy, 0, 0, 0, 0, ocam.width, ocam.height, 0, 5, 0)
    # map_build_node.cc:216-225: T = [rlc^T | -rlc^T tlc]
    a = 0.02
    rlc = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
    tlc = np.array([0.1, -0.2, 0.05])
    T = np.concatenate([rlc.T, (-rlc.T @ tlc)[:, None]], axis=1)
    got = ctx.project_color(pts, img, gcam, [0, 0, 0, 1], [0, 0, 0], T_cam_lidar=T)
    pc = oracle.transform_cloud(pts, T)
    raw = oracle.project_raster(pc, ocam)
    assert np.array_equal(got["depth_raw"], raw)
    got = ctx.project_color(pts[:0], img, gcam, [0, 0, 0, 1], [0, 0, 0])
    assert got["depth_raw"].max() == 0 and len(got["cloud_world"]) == 0


def test_projection_for_the_pro_map_discs(gpu_ctx_factory, oracle):
    """lmono_color_projection: the cv::Point2f and depth of every point, in cloud order (what the node draws the r = 3 HSV
    discs of ~pro_map at, Map_Builder.cc:234-243).  Writing 100 - z at (int(v), int(u)) point after point on the host must
    rebuild the oracle's raster (last point wins), with and without lens distortion."""
    from lmono_b200 import api
    ctx = gpu_ctx_factory()
    pts, img = _scene(60_000, 12)
    for kw in ({}, dict(k1=-0.05, k2=0.01, p1=0.001, p2=-0.002)):
        ocam = oracle.make_camera(**kw)
        gcam = api.Pinhole(ocam.fx, ocam.fy, ocam.cx, ocam.cy, ocam.k1, ocam.k2, ocam.p1, ocam.p2, ocam.width, ocam.height, 0, 5, 0)
        ctx.project_color(pts, img, gcam, [0, 0, 0, 1], [0, 0, 0])
        uvz = ctx.color_projection(len(pts))
        assert np.array_equal(uvz[:, 2], pts[:, 2])
        ok = ~np.isnan(uvz[:, 0])
        assert 1000 < ok.sum() < len(pts) and not np.any(ok & (pts[:, 2] < 0))
        raster = np.zeros((ocam.height, ocam.width), np.uint8)
        u, v = uvz[ok, 0].astype(np.int64), uvz[ok, 1].astype(np.int64)
        val = (100.0 - uvz[ok, 2].astype(np.float64)).astype(np.int64).astype(np.uint8)       # double -> int -> low byte, as the reference's implicit conversion
        raster[v, u] = val                                                                     # numpy assigns in order: the last point wins
        assert np.array_equal(raster, oracle.project_raster(pts, ocam))
