"""CPU test (-m "not gpu"): the oracle's scanRegistration (C) against an independent line-by-line Python restatement of
laserCloudHandler (Aloam/src/scanRegistration.cpp:132-408) with numpy float32 scalars.  Two restatements written
separately from the same reference text agreeing on every ring id, curvature bit, label and picked feature is what
pins stage A in the absence of reference tests."""
import math

import numpy as np
import pytest

from lmono_b200 import synth
from test_oracle_primitives import np_voxel_grid

f32 = np.float32
PI = math.pi


def py_scan_registration(raw, n_scans, minimum_range, scan_period=0.1):
    # :136-137 removeNaN + removeClosedPointCloud (:85-112): x*x + y*y + z*z < thres*thres in float, order preserving
    pts = []
    th2 = f32(minimum_range) * f32(minimum_range)
    for p in raw[:, :3].astype(np.float32):
        if not np.all(np.isfinite(p)):
            continue
        if f32(f32(p[0] * p[0]) + f32(p[1] * p[1])) + f32(p[2] * p[2]) < th2:
            continue
        pts.append(p)
    n = len(pts)
    # :141-153 (std::atan2(float, float) -> float; "+ 2 * M_PI" in double, stored to float)
    start_ori = f32(-f32(math.atan2(float(pts[0][1]), float(pts[0][0]))))
    end_ori = f32(float(f32(-f32(math.atan2(float(pts[-1][1]), float(pts[-1][0]))))) + 2 * PI)
    if float(f32(end_ori - start_ori)) > 3 * PI:
        end_ori = f32(float(end_ori) - 2 * PI)
    elif float(f32(end_ori - start_ori)) < PI:
        end_ori = f32(float(end_ori) + 2 * PI)
    half = False
    rings = [[] for _ in range(n_scans)]
    src = [[] for _ in range(n_scans)]
    kept_src = 0
    k_in = -1
    for p in pts:
        k_in += 1
        x, y, z = p
        # :166 atan / sqrt resolve to the double C functions on the author's toolchain (DESIGN.md hazard table)
        r2 = f32(f32(x * x) + f32(y * y))
        angle = f32(math.atan(float(z) / math.sqrt(float(r2))) * 180 / PI)
        if n_scans == 16:
            sid = int(float(f32(f32(angle + f32(15)) / f32(2))) + 0.5)       # :171
            if sid > n_scans - 1 or sid < 0:
                continue
        elif n_scans == 32:
            sid = int((float(angle) + 92.0 / 3.0) * 3.0 / 4.0)              # :180
            if sid > n_scans - 1 or sid < 0:
                continue
        else:
            if float(angle) >= -8.83:
                sid = int((2 - float(angle)) * 3.0 + 0.5)
            else:
                sid = n_scans // 2 + int((-8.83 - float(angle)) * 2.0 + 0.5)
            if float(angle) > 2 or float(angle) < -24.33 or sid > 50 or sid < 0:
                continue
        ori = f32(-f32(math.atan2(float(y), float(x))))                     # :208
        if not half:
            if float(ori) < float(start_ori) - PI / 2:
                ori = f32(float(ori) + 2 * PI)
            elif float(ori) > float(start_ori) + PI * 3 / 2:
                ori = f32(float(ori) - 2 * PI)
            if float(f32(ori - start_ori)) > PI:
                half = True
        else:
            ori = f32(float(ori) + 2 * PI)
            if float(ori) < float(end_ori) - PI * 3 / 2:
                ori = f32(float(ori) + 2 * PI)
            elif float(ori) > float(end_ori) + PI / 2:
                ori = f32(float(ori) - 2 * PI)
        rel = f32(f32(ori - start_ori) / f32(end_ori - start_ori))          # :238
        inten = f32(sid + scan_period * float(rel))                          # :239 (double product / sum -> float)
        rings[sid].append((x, y, z, inten))
        src[sid].append(k_in)
    cloud = np.array([q for r in rings for q in r], np.float32).reshape(-1, 4)
    N = len(cloud)
    start_ind, end_ind = [], []
    pos = 0
    for r in rings:                                                          # :246-252
        start_ind.append(pos + 5)
        pos += len(r)
        end_ind.append(pos - 6)
    curv = np.zeros(N, np.float32)
    X = cloud[:, :3]
    for i in range(5, N - 5):                                                # :256-266 strictly left to right in float
        d = np.zeros(3, np.float32)
        for a in range(3):
            s = f32(X[i - 5, a] + X[i - 4, a])
            s = f32(s + X[i - 3, a]); s = f32(s + X[i - 2, a]); s = f32(s + X[i - 1, a])
            s = f32(s - f32(f32(10) * X[i, a]))
            s = f32(s + X[i + 1, a]); s = f32(s + X[i + 2, a]); s = f32(s + X[i + 3, a]); s = f32(s + X[i + 4, a]); s = f32(s + X[i + 5, a])
            d[a] = s
        curv[i] = f32(f32(f32(d[0] * d[0]) + f32(d[1] * d[1])) + f32(d[2] * d[2]))
    picked = np.zeros(N + 16, np.int32)
    label = np.zeros(N, np.int32)

    def gap(a, b):                                                           # :319-342
        dx, dy, dz = f32(X[a, 0] - X[b, 0]), f32(X[a, 1] - X[b, 1]), f32(X[a, 2] - X[b, 2])
        return float(f32(f32(f32(dx * dx) + f32(dy * dy)) + f32(dz * dz))) > 0.05

    def suppress(ind):
        for l in range(1, 6):
            if gap(ind + l, ind + l - 1):
                break
            picked[ind + l] = 1
        for l in range(-1, -6, -1):
            if gap(ind + l, ind + l + 1):
                break
            picked[ind + l] = 1

    sharp, less_sharp, flat, less_flat = [], [], [], []
    for i in range(n_scans):
        if end_ind[i] - start_ind[i] < 6:                                    # :279
            continue
        lf_scan = []
        for j in range(6):
            sp = start_ind[i] + (end_ind[i] - start_ind[i]) * j // 6
            ep = start_ind[i] + (end_ind[i] - start_ind[i]) * (j + 1) // 6 - 1
            order = sorted(range(sp, ep + 1), key=lambda k: (curv[k], k))    # std::sort by curvature; canonical tie order
            largest = 0
            for ind in reversed(order):
                if picked[ind] == 0 and float(curv[ind]) > 0.1:
                    largest += 1
                    if largest <= 2:
                        label[ind] = 2; sharp.append(cloud[ind]); less_sharp.append(cloud[ind])
                    elif largest <= 20:
                        label[ind] = 1; less_sharp.append(cloud[ind])
                    else:
                        break
                    picked[ind] = 1
                    suppress(ind)
            smallest = 0
            for ind in order:
                if picked[ind] == 0 and float(curv[ind]) < 0.1:
                    label[ind] = -1; flat.append(cloud[ind])
                    smallest += 1
                    if smallest >= 4:                                        # :359-362: pushed, but neither marked nor suppressed
                        break
                    picked[ind] = 1
                    suppress(ind)
            for k in range(sp, ep + 1):                                      # :392-398
                if label[k] <= 0:
                    lf_scan.append(cloud[k])
        if lf_scan:
            less_flat.append(np_voxel_grid(np.array(lf_scan, np.float32), 0.2))   # :401-405
    cat = lambda l: np.array(l, np.float32).reshape(-1, 4)
    return {"full": cloud, "curvature": curv, "labels": label, "sharp": cat(sharp), "less_sharp": cat(less_sharp), "flat": cat(flat),
            "less_flat": np.concatenate(less_flat) if less_flat else np.zeros((0, 4), np.float32)}


@pytest.mark.parametrize("n_scans,min_range", [(16, 0.3), (32, 0.3), (64, 5.0)])
def test_oracle_scan_registration_equals_python_restatement(oracle, n_scans, min_range):
    w = synth.make_world()
    rng = np.random.default_rng(3)
    q, t = synth.loop_pose(w, 2.0)
    raw = synth.raycast_sweep(w, q, t, n_scans, 360, rng)          # 360 azimuth steps per ring keep the Python loops short
    raw[7, 0] = np.nan                                              # removeNaNFromPointCloud
    r = oracle.scan_register(raw, n_scans, min_range)
    p = py_scan_registration(raw, n_scans, min_range)
    assert len(p["full"]) == len(r["full"]) > 1000
    assert np.array_equal(p["full"][:, :3], r["full"][:, :3])
    assert np.array_equal(np.floor(p["full"][:, 3]), np.floor(r["full"][:, 3]))                  # ring ids
    assert np.abs(p["full"][:, 3] - r["full"][:, 3]).max() <= 4e-6                               # 0.1 * relTime (atan2f is not correctly rounded everywhere)
    assert np.array_equal(p["curvature"][5:-5].view(np.uint32), r["curvature"][5:-5].view(np.uint32))
    assert np.array_equal(p["labels"], r["labels"])
    for k in ("sharp", "less_sharp", "flat"):
        assert p[k].shape == r[k].shape and len(p[k]) > 0, k
        assert np.array_equal(p[k][:, :3], r[k][:, :3]), k
    assert p["less_flat"].shape == r["less_flat"].shape
    assert np.array_equal(p["less_flat"][:, :3], r["less_flat"][:, :3])
    assert np.abs(p["less_flat"][:, 3] - r["less_flat"][:, 3]).max() <= 8e-6
