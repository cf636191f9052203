"""CPU tests: the oracle's depthFill (Map_Builder.cc:336-403) against cv2.  dilate / close /
median / 5x5 Gaussian are exact integer operations and must agree bit for bit.  cv2's wheel
routes bilateralFilter through IPP, which truncates where OpenCV 3.x's own code (the reference's
dependency, restated by the oracle) calls cvRound -- so there the bar is: cv2 == floor, oracle ==
round of the same weighted mean, i.e. 0 <= oracle - cv2 <= 1."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def _raw(oracle, cam, seed=5, n=120_000):
    rng = np.random.default_rng(seed)
    pts = np.zeros((n, 3), np.float32)
    pts[:, 2] = rng.uniform(-5, 80, n)
    pts[:, 0] = rng.uniform(-1, 1, n) * pts[:, 2] * 0.9
    pts[:, 1] = rng.uniform(-0.3, 0.3, n) * pts[:, 2]
    return pts, oracle.project_raster(pts, cam)


def _cv_fill(raw, kt, bt):
    K = cv2.getStructuringElement([cv2.MORPH_RECT, cv2.MORPH_CROSS, cv2.MORPH_ELLIPSE][kt], (5, 5))
    d1 = cv2.dilate(raw, K)
    hf = cv2.morphologyEx(d1, cv2.MORPH_CLOSE, cv2.getStructuringElement(cv2.MORPH_RECT, (5, 5)))
    d2 = cv2.dilate(hf, cv2.getStructuringElement(cv2.MORPH_RECT, (7, 7)))
    hf = np.where(hf < 0.1, d2, hf)
    med = cv2.medianBlur(hf, 5)
    return cv2.bilateralFilter(med, 5, 1.5, 2.0) if bt == 0 else cv2.GaussianBlur(med, (5, 5), 0)


@pytest.mark.parametrize("kt", [0, 1, 2])
def test_depth_fill_gaussian_path_bit_exact_vs_cv2(oracle, kt):
    cam = oracle.make_camera(kernel_type=kt, blur_type=1)
    _, raw = _raw(oracle, cam)
    assert np.array_equal(oracle.depth_fill(raw, cam), _cv_fill(raw, kt, 1))


def test_depth_fill_bilateral_vs_cv2_rounding_only(oracle):
    cam = oracle.make_camera(kernel_type=0, blur_type=0)
    _, raw = _raw(oracle, cam)
    d = oracle.depth_fill(raw, cam).astype(int) - _cv_fill(raw, 0, 0).astype(int)
    assert d.min() >= 0 and d.max() <= 1


def test_raster_last_writer_wins_and_bounds(oracle):
    cam = oracle.make_camera()
    # two points on the same pixel: the later one wins; a point behind the camera is skipped
    pts = np.array([[0.0, 0.0, 10.0], [0.0, 0.0, 30.0], [0.0, 0.0, -1.0], [1e6, 0.0, 10.0]], np.float32)
    raw = oracle.project_raster(pts, cam)
    v, u = int(cam.cy), int(cam.cx)
    assert raw[v, u] == 70 and (raw > 0).sum() == 1


def test_lift_inverts_projection(oracle):
    cam = oracle.make_camera()
    depth = np.zeros((cam.height, cam.width), np.uint8)
    depth[100, 200] = 60          # d = 40 m
    depth[50, 700] = 100          # d = 0 -> skipped
    depth[10, 10] = 20            # d = 80 >= 70 -> skipped
    img = np.zeros((cam.height, cam.width, 3), np.uint8)
    img[100, 200] = (1, 2, 3)
    cc, cw, rgb = oracle.lift_cloud(depth, img, cam, [0, 0, 0, 1], [1, 2, 3])
    assert len(cc) == 1 and tuple(rgb[0]) == (3, 2, 1)
    x, y, z = cc[0]
    assert z == 40 and abs(x - 40 * (200 - cam.cx) / cam.fx) < 1e-4 and abs(y - 40 * (100 - cam.cy) / cam.fy) < 1e-4
    assert np.allclose(cw[0], cc[0] + [1, 2, 3], atol=1e-5)
