"""CPU test (-m "not gpu"): the oracle's projection raster and depth-image lift (C) against an independent Python
restatement of mono_lidar_mapping/src/map_builder/Map_Builder.cc:224-245, 275-322, 405-416 and
camera_models/src/camera_models/PinholeCamera.cc:450-542 (zero distortion, the KITTI configuration)."""
import numpy as np


def py_raster(pts, cam):
    out = np.zeros((cam.height, cam.width), np.uint8)
    for x, y, z in pts[:, :3].astype(np.float32):
        if z < 0:                                                               # :226-229
            continue
        with np.errstate(divide="ignore", invalid="ignore"):                    # z == 0: inf / nan fail the bounds test, as in C++
            u = np.float32(cam.fx * (np.float64(x) / np.float64(z)) + cam.cx)   # spaceToPlane in double -> cv::Point2f
            v = np.float32(cam.fy * (np.float64(y) / np.float64(z)) + cam.cy)
        if u > 0 and u < cam.width and v > 0 and v < cam.height:                # :233
            out[int(v), int(u)] = int(100.0 - float(z)) & 0xFF                  # :238 double -> uchar (x86-64: truncate, low byte); last writer wins
    return out


def rot(q):
    x, y, z, w = q                                                              # Eigen::Quaterniond::toRotationMatrix
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def py_lift(depth, bgr, cam, q, t):
    cc, cw, rgb = [], [], []
    R = rot(np.asarray(q, np.float64))
    for j in range(cam.height):                                                 # :275-322, row-major
        for i in range(cam.width):
            d = 100 - int(depth[j, i])
            if d <= 0 or d >= 70:
                continue
            mx = (1.0 / cam.fx) * i + (-cam.cx / cam.fx)                        # liftProjective, m_noDistortion
            my = (1.0 / cam.fy) * j + (-cam.cy / cam.fy)
            p = np.array([d * mx / 1.0, d * my / 1.0, d], np.float32)
            if abs(float(p[0])) > 20 and float(p[1]) > 1.8:                     # :305 (float abs, see DESIGN.md hazards); floats widened to the double literals
                continue
            cc.append(p)
            cw.append((R @ p.astype(np.float64) + np.asarray(t, np.float64)).astype(np.float32))
            rgb.append(bgr[j, i][::-1])
    return np.array(cc, np.float32).reshape(-1, 3), np.array(cw, np.float32).reshape(-1, 3), np.array(rgb, np.uint8).reshape(-1, 3)


def test_oracle_raster_equals_python_restatement(oracle):
    cam = oracle.make_camera()
    rng = np.random.default_rng(12)
    n = 30_000
    pts = np.zeros((n, 3), np.float32)
    pts[:, 2] = rng.uniform(-5, 130, n)                                         # behind the camera, and beyond 100 m (the uchar wraps)
    pts[:, 0] = rng.uniform(-1, 1, n) * pts[:, 2]
    pts[:, 1] = rng.uniform(-0.35, 0.35, n) * pts[:, 2]
    pts[:5, 2] = 0.0
    got = oracle.project_raster(pts, cam)
    assert (got > 0).sum() > 5000
    assert np.array_equal(got, py_raster(pts, cam))


def test_oracle_lift_equals_python_restatement(oracle):
    cam = oracle.make_camera(width=160, height=48, cx=80.0, cy=24.0, fx=90.0, fy=90.0)   # small image: Python loops over every pixel
    rng = np.random.default_rng(13)
    depth = rng.integers(0, 140, (cam.height, cam.width)).astype(np.uint8)
    img = rng.integers(0, 255, (cam.height, cam.width, 3)).astype(np.uint8)
    q = np.array([0.1, -0.2, 0.05, 0.97]); q /= np.linalg.norm(q)
    t = np.array([3.0, -1.0, 0.5])
    cc, cw, rgb = oracle.lift_cloud(depth, img, cam, q, t)
    pc, pw, prgb = py_lift(depth, img, cam, q, t)
    assert len(cc) == len(pc) > 1000
    assert np.array_equal(cc, pc) and np.array_equal(rgb, prgb)
    assert np.abs(cw - pw).max() <= 2e-5
    assert (np.abs(pc[:, 0]) > 20).sum() > 50                                    # the :305 filter is exercised
