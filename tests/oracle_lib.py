"""ctypes loader for the CPU oracle (oracle/liblmono_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never by lmono_b200/."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "liblmono_oracle.so")


class Pose(C.Structure):
    _fields_ = [("q", C.c_double * 4), ("t", C.c_double * 3)]

    @staticmethod
    def make(q=(0, 0, 0, 1), t=(0, 0, 0)):
        p = Pose()
        p.q[:] = [float(v) for v in q]
        p.t[:] = [float(v) for v in t]
        return p

    def as_np(self):
        return np.array(self.q[:]), np.array(self.t[:])


class Factor(C.Structure):
    _fields_ = [("type", C.c_int32), ("pad", C.c_int32), ("p", C.c_double * 3),
                ("a", C.c_double * 3), ("b", C.c_double * 3), ("s", C.c_double)]


FACTOR_DTYPE = np.dtype([("type", "<i4"), ("pad", "<i4"), ("p", "<f8", 3), ("a", "<f8", 3), ("b", "<f8", 3), ("s", "<f8")])


class SolveSummary(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("num_successful", C.c_int32), ("termination", C.c_int32),
                ("num_factors", C.c_int32), ("initial_cost", C.c_double), ("final_cost", C.c_double)]


class MapReport(C.Structure):
    _fields_ = [("corner_from_map", C.c_int32), ("surf_from_map", C.c_int32),
                ("corner_stack", C.c_int32), ("surf_stack", C.c_int32),
                ("corner_num", C.c_int32 * 2), ("surf_num", C.c_int32 * 2),
                ("optimized", C.c_int32), ("center_cube", C.c_int32 * 3), ("cen", C.c_int32 * 3),
                ("solve", SolveSummary * 2),
                ("ms_shift", C.c_double), ("ms_tree", C.c_double), ("ms_assoc", C.c_double),
                ("ms_solver", C.c_double), ("ms_add", C.c_double), ("ms_filter", C.c_double),
                ("ms_whole", C.c_double)]


class ScanReport(C.Structure):
    _fields_ = [("n_in", C.c_int32), ("n_kept", C.c_int32), ("n_sharp", C.c_int32),
                ("n_less_sharp", C.c_int32), ("n_flat", C.c_int32), ("n_less_flat", C.c_int32),
                ("ring_start", C.c_int32 * 64), ("ring_end", C.c_int32 * 64),
                ("start_ori", C.c_float), ("end_ori", C.c_float), ("min_margin_ok", C.c_int32)]


class OdomReport(C.Structure):
    _fields_ = [("inited", C.c_int32), ("corner_corr", C.c_int32 * 2), ("plane_corr", C.c_int32 * 2),
                ("solve", SolveSummary * 2),
                ("ms_assoc", C.c_double), ("ms_solver", C.c_double), ("ms_whole", C.c_double)]


class Camera(C.Structure):
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("k1", C.c_double), ("k2", C.c_double), ("p1", C.c_double), ("p2", C.c_double),
                ("width", C.c_int32), ("height", C.c_int32), ("kernel_type", C.c_int32),
                ("kernel_size", C.c_int32), ("blur_type", C.c_int32)]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc/g++ (seconds)."""
    d = os.path.join(_ROOT, "oracle")
    srcs = [os.path.join(d, f) for f in os.listdir(d) if f.endswith((".c", ".cpp", ".h", "Makefile"))]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.run(["make", "-s", "-C", d], check=True)
    return _SO


REF_ROOT = "/root/reference"          # exists in the build container only, never on the GPU box


def build_ref() -> str:
    """oracle/_ref: the reference's own translation units compiled against oracle/refstubs (only where the reference
    tree exists; elsewhere the prebuilt files that travelled with the snapshot are used as they are)."""
    d = os.path.join(_ROOT, "oracle")
    if os.path.isdir(os.path.join(REF_ROOT, "Aloam", "src")):
        build()                                    # the stand-ins link against the oracle
        try:
            subprocess.run(["make", "-s", "-C", d, "ref", f"REF={REF_ROOT}"], check=True)
        except (subprocess.CalledProcessError, OSError) as e:     # checker infrastructure: a failed build skips the tests that need it
            print(f"[oracle_lib] oracle/_ref not (re)built: {e}", file=sys.stderr)
    return os.path.join(d, "_ref")


_ref_libs = {}


def ref_lib(name):
    """ctypes handle of oracle/_ref/libref_<name>.so, or None when it was never built (no reference tree, no prebuilt file)."""
    if name not in _ref_libs:
        path = os.path.join(build_ref(), f"libref_{name}.so")
        _ref_libs[name] = None
        if os.path.exists(path):
            try:
                lib()                              # oracle/liblmono_oracle.so, which the stand-ins call into
                _ref_libs[name] = C.CDLL(path)
            except OSError as e:
                print(f"[oracle_lib] {path} does not load: {e}", file=sys.stderr)
    return _ref_libs[name]


def ref_scan_register(raw, n_scans=64, minimum_range=5.0):
    """laserCloudHandler of the reference (Aloam/src/scanRegistration.cpp:113-459) as compiled from /root/reference:
    what the node published on its five topics, plus its cloudLabel / cloudCurvature arrays."""
    R = ref_lib("scanreg")
    raw = np.ascontiguousarray(raw, np.float32)
    n = len(raw)
    names = ("full", "sharp", "less_sharp", "flat", "less_flat")
    bufs = [np.zeros((max(n, 1), 4), np.float32) for _ in names]
    counts = np.zeros(5, np.int32)
    labels = np.zeros(max(n, 1), np.int32)
    curv = np.zeros(max(n, 1), np.float32)
    R.ref_scan_register.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double] + [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 3
    rc = R.ref_scan_register(_p(raw), n, raw.shape[1], n_scans, float(minimum_range), *[_p(b) for b in bufs], n, _p(counts), _p(labels), _p(curv))
    assert rc == 0, rc
    out = {k: bufs[i][: counts[i]] for i, k in enumerate(names)}
    out["labels"] = labels[: counts[0]]
    out["curvature"] = curv[: counts[0]]
    return out


class RefOdometry:
    """The reference's laserOdometry node (Aloam/src/laserOdometry.cpp) compiled from /root/reference: one step() per
    sweep through the node's own main loop.  Process-wide state, like the node's globals."""

    def __init__(self):
        self.R = ref_lib("odom")
        lib()                                      # the stand-ins call the oracle's kd-tree and minimiser
        self.R.ref_odom_reset()
        self.k = 0

    def step(self, sharp, less_sharp, flat, less_flat, full=None):
        a = [_f32(x).reshape(-1, 4) for x in (sharp, less_sharp, flat, less_flat, full if full is not None else np.zeros((0, 4), np.float32))]
        out = np.zeros(14)
        cnt = np.zeros(2, np.int32)
        args = []
        for x in a:
            args += [_p(x), len(x)]
        rc = self.R.ref_odom_step(*args, C.c_double(0.1 * self.k), _p(out), _p(cnt))
        assert rc == 0, rc
        self.k += 1
        return (out[7:11].copy(), out[11:14].copy()), (out[0:4].copy(), out[4:7].copy()), cnt


def ref_lm_solve(factors, q, t, max_iter=4):
    """ceres::Solve on a problem built like the reference builds it, with the cost functors of lidarFactor.hpp as compiled
    from /root/reference evaluated on dual numbers; the minimiser loop is oracle/lm.c (factor types 0 and 2)."""
    R = ref_lib("odom")
    lib()
    f = np.ascontiguousarray(factors, FACTOR_DTYPE)
    pose = Pose.make(q, t)
    s = SolveSummary()
    rc = R.ref_lm_solve(_p(f), len(f), C.byref(pose), max_iter, C.byref(s))
    assert rc == 0, rc
    qo, to = pose.as_np()
    return qo, to, s


def ref_normal_eq(factors, q, t):
    R = ref_lib("odom")
    lib()
    f = np.ascontiguousarray(factors, FACTOR_DTYPE)
    H = np.zeros(36)
    g = np.zeros(6)
    cost = C.c_double(0)
    pose = Pose.make(q, t)
    rc = R.ref_normal_eq(_p(f), len(f), C.byref(pose), _p(H), _p(g), C.byref(cost))
    assert rc == 0, rc
    return H.reshape(6, 6), g, cost.value


class RefMapper:
    """The reference's laserMapping node (Aloam/src/laserMapping.cpp) compiled from /root/reference: one step() per sweep
    through the node's callbacks and process().  Process-wide state, like the node's globals."""

    def __init__(self, line_res=0.4, plane_res=0.8):
        self.R = ref_lib("mapping")
        lib()
        assert self.R.ref_mapping_reset(C.c_double(line_res), C.c_double(plane_res)) == 0
        self.k = 0

    def step(self, corner_last, surf_last, q_odom, t_odom, full_res=None):
        c = _f32(corner_last).reshape(-1, 4)
        s = _f32(surf_last).reshape(-1, 4)
        fr = None if full_res is None else _f32(full_res).reshape(-1, 4).copy()
        out = np.zeros(14)
        info = np.zeros(4, np.int32)
        q = np.ascontiguousarray(q_odom, np.float64)
        t = np.ascontiguousarray(t_odom, np.float64)
        rc = self.R.ref_mapping_step(_p(c), len(c), _p(s), len(s), _p(fr) if fr is not None else None, 0 if fr is None else len(fr),
                                     _p(q), _p(t), C.c_double(0.1 * self.k), _p(out), _p(info))
        assert rc == 0, rc
        self.k += 1
        return out[0:4].copy(), out[4:7].copy(), (out[7:11].copy(), out[11:14].copy()), [int(v) for v in info[:3]], fr

    def import_points(self, which, pts):
        pts = _f32(pts).reshape(-1, 4)
        self.R.ref_mapping_import(which, _p(pts), len(pts))

    def set_state(self, q, t):
        q = np.ascontiguousarray(q, np.float64)
        t = np.ascontiguousarray(t, np.float64)
        self.R.ref_mapping_set_state(_p(q), _p(t))

    def export(self, which):
        n = self.R.ref_mapping_export(which, None, 0)
        o = np.zeros((max(n, 1), 4), np.float32)
        self.R.ref_mapping_export(which, _p(o), n)
        return o[:n]


def ref_color_frame(pts_cam, bgr, cam, q, t):
    """MapBuilder::associateToMap of the reference (mono_lidar_mapping/src/map_builder/Map_Builder.cc:213-334) as compiled
    from /root/reference: raster before depthFill, depth image after it, lifted cloud (camera frame, world frame, r g b)."""
    R = ref_lib("color")
    lib()
    pts = _f32(pts_cam)
    img = np.ascontiguousarray(bgr, np.uint8)
    npix = cam.width * cam.height
    raw = np.zeros((cam.height, cam.width), np.uint8)
    filled = np.zeros((cam.height, cam.width), np.uint8)
    cc = np.zeros((npix, 3), np.float32)
    cw = np.zeros((npix, 3), np.float32)
    rgb = np.zeros((npix, 3), np.uint8)
    n = C.c_int(0)
    pose = Pose.make(q, t)
    rc = R.ref_color_frame(_p(pts), len(pts), pts.shape[1], _p(img), C.byref(cam), C.byref(pose), _p(raw), _p(filled), _p(cc), _p(cw), _p(rgb), npix, C.byref(n))
    assert rc == 0, rc
    return raw, filled, cc[: n.value], cw[: n.value], rgb[: n.value]


def ref_mapnode_frame(pts_lidar, bgr, cam, q_lc, t_lc, q, t, stamp=1.0):
    """The reference's colour-map NODE (mono_lidar_mapping/src/map_build_node.cc) as compiled from /root/reference: the
    lidar-to-camera extrinsic, the lidar-frame cloud, the image and the camera pose go through the node's own handlers and
    process() (:73-238), which transforms the cloud and calls MapBuilder::associateToMap.  Same outputs as ref_color_frame."""
    R = ref_lib("color")
    lib()
    pts = _f32(pts_lidar)
    img = np.ascontiguousarray(bgr, np.uint8)
    npix = cam.width * cam.height
    raw = np.zeros((cam.height, cam.width), np.uint8)
    filled = np.zeros((cam.height, cam.width), np.uint8)
    cc = np.zeros((npix, 3), np.float32)
    cw = np.zeros((npix, 3), np.float32)
    rgb = np.zeros((npix, 3), np.uint8)
    n = C.c_int(0)
    ext = Pose.make(q_lc, t_lc)
    pose = Pose.make(q, t)
    rc = R.ref_mapnode_frame(_p(pts), len(pts), pts.shape[1], _p(img), C.byref(cam), C.byref(ext), C.byref(pose), C.c_double(stamp),
                             _p(raw), _p(filled), _p(cc), _p(cw), _p(rgb), npix, C.byref(n))
    assert rc == 0, rc
    return raw, filled, cc[: n.value], cw[: n.value], rgb[: n.value]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.lmono_cpu_mapper_create.restype = C.c_void_p
        _lib.lmono_cpu_kdtree_build.restype = C.c_void_p
        if hasattr(_lib, "lmono_cpu_odom_create"):
            _lib.lmono_cpu_odom_create.restype = C.c_void_p
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def voxel_grid(pts, leaf, order_mode=0):
    pts = _f32(pts).reshape(-1, 4)
    out = np.zeros_like(pts)
    n = C.c_int(0)
    lib().lmono_cpu_voxel_grid(_p(pts), len(pts), C.c_float(leaf), order_mode, _p(out), C.byref(n))
    return out[: n.value].copy()


def knn_brute(pts, queries, k=5):
    pts = _f32(pts).reshape(-1, 4)
    queries = _f32(queries).reshape(-1, 4)
    idx = np.zeros((len(queries), k), np.int32)
    d2 = np.zeros((len(queries), k), np.float32)
    lib().lmono_cpu_knn_brute(_p(pts), len(pts), _p(queries), len(queries), k, _p(idx), _p(d2))
    return idx, d2


def knn_kdtree(pts, queries, k=5):
    pts = _f32(pts).reshape(-1, 4)
    queries = _f32(queries).reshape(-1, 4)
    idx = np.zeros((len(queries), k), np.int32)
    d2 = np.zeros((len(queries), k), np.float32)
    L = lib()
    t = C.c_void_p(L.lmono_cpu_kdtree_build(_p(pts), len(pts)))
    L.lmono_cpu_kdtree_knn(t, _p(queries), len(queries), k, _p(idx), _p(d2))
    L.lmono_cpu_kdtree_free(t)
    return idx, d2


def eigh3(A):
    A = np.ascontiguousarray(A, np.float64).reshape(9)
    w = np.zeros(3)
    V = np.zeros(9)
    rc = lib().lmono_cpu_eigh3(_p(A), _p(w), _p(V))
    return w, V.reshape(3, 3), rc


def colpiv_solve_5x3(A, b):
    A = np.ascontiguousarray(A, np.float64).reshape(15)
    b = np.ascontiguousarray(b, np.float64).reshape(5)
    x = np.zeros(3)
    lib().lmono_cpu_colpiv_qr_solve_5x3(_p(A), _p(b), _p(x))
    return x


def normal_eq(factors, q, t):
    f = np.ascontiguousarray(factors, FACTOR_DTYPE)
    H = np.zeros(36)
    g = np.zeros(6)
    cost = C.c_double(0)
    pose = Pose.make(q, t)
    lib().lmono_cpu_normal_eq(_p(f), len(f), C.byref(pose), _p(H), _p(g), C.byref(cost))
    return H.reshape(6, 6), g, cost.value


def lm_solve(factors, q, t, max_iter=4):
    f = np.ascontiguousarray(factors, FACTOR_DTYPE)
    pose = Pose.make(q, t)
    s = SolveSummary()
    lib().lmono_cpu_lm_solve(_p(f), len(f), C.byref(pose), max_iter, C.byref(s))
    qo, to = pose.as_np()
    return qo, to, s


class Mapper:
    """Oracle laserMapping state (Aloam/src/laserMapping.cpp globals)."""

    def __init__(self, line_res=0.4, plane_res=0.8, order_mode=0, use_kdtree=1):
        self.L = lib()
        self.h = C.c_void_p(self.L.lmono_cpu_mapper_create(C.c_float(line_res), C.c_float(plane_res), order_mode, use_kdtree))

    def close(self):
        if self.h:
            self.L.lmono_cpu_mapper_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def import_points(self, which, pts):
        pts = _f32(pts).reshape(-1, 4)
        self.L.lmono_cpu_mapper_import(self.h, which, _p(pts), len(pts))

    def export(self, which, scope=1):
        n = self.L.lmono_cpu_mapper_export(self.h, which, scope, None, 0)
        out = np.zeros((max(n, 1), 4), np.float32)
        self.L.lmono_cpu_mapper_export(self.h, which, scope, _p(out), n)
        return out[:n]

    def get_state(self):
        p = Pose()
        cen = (C.c_int32 * 3)()
        self.L.lmono_cpu_mapper_get_state(self.h, C.byref(p), cen)
        q, t = p.as_np()
        return q, t, list(cen)

    def set_state(self, q, t):
        p = Pose.make(q, t)
        self.L.lmono_cpu_mapper_set_state(self.h, C.byref(p))

    def prepare_window(self, t_w_curr):
        t = (C.c_double * 3)(*[float(v) for v in t_w_curr])
        return self.L.lmono_cpu_mapper_prepare_window(self.h, t)

    def knn5(self, which, queries_world):
        qw = _f32(queries_world).reshape(-1, 4)
        idx = np.zeros((len(qw), 5), np.int32)
        d2 = np.zeros((len(qw), 5), np.float32)
        self.L.lmono_cpu_mapper_knn5(self.h, which, _p(qw), len(qw), _p(idx), _p(d2))
        return idx, d2

    def associate(self, corner_stack, surf_stack, q, t):
        cs = _f32(corner_stack).reshape(-1, 4)
        ss = _f32(surf_stack).reshape(-1, 4)
        cap = len(cs) + len(ss)
        out = np.zeros(max(cap, 1), FACTOR_DTYPE)
        nc = C.c_int32(0)
        ns = C.c_int32(0)
        pose = Pose.make(q, t)
        nf = self.L.lmono_cpu_mapper_associate(self.h, _p(cs), len(cs), _p(ss), len(ss), C.byref(pose),
                                               _p(out), cap, C.byref(nc), C.byref(ns))
        return out[:nf], nc.value, ns.value

    def step(self, corner_last, surf_last, q_odom, t_odom, full_res=None):
        cl = _f32(corner_last).reshape(-1, 4)
        sl = _f32(surf_last).reshape(-1, 4)
        odom = Pose.make(q_odom, t_odom)
        w = Pose()
        rep = MapReport()
        fr = None
        nfull = 0
        if full_res is not None:
            fr = _f32(full_res).reshape(-1, 4).copy()
            nfull = len(fr)
        self.L.lmono_cpu_map_step(self.h, _p(cl), len(cl), _p(sl), len(sl), C.byref(odom), C.byref(w),
                                  C.byref(rep), _p(fr) if fr is not None else None, nfull)
        q, t = w.as_np()
        return q, t, rep, fr


def scan_register(raw, n_scans=64, minimum_range=5.0, voxel_order_mode=0, sort_mode=0, trig_mode=0):
    """raw: float32 [n, >=3].  Returns dict with full/sharp/less_sharp/flat/less_flat clouds,
    labels, curvature, src_index, report."""
    raw = np.ascontiguousarray(raw, np.float32)
    n = len(raw)
    L = lib()
    L.lmono_cpu_scan_set_trig_mode(trig_mode)
    cap = max(n, 1)
    bufs = {k: np.zeros((cap, 4), np.float32) for k in ("full", "sharp", "less_sharp", "flat", "less_flat")}
    labels = np.zeros(cap, np.int32)
    curv = np.zeros(cap, np.float32)
    src = np.zeros(cap, np.int32)
    rep = ScanReport()
    rc = L.lmono_cpu_scan_register(_p(raw), n, raw.shape[1], n_scans, C.c_float(minimum_range), voxel_order_mode, sort_mode,
                                   _p(bufs["full"]), _p(bufs["sharp"]), _p(bufs["less_sharp"]), _p(bufs["flat"]),
                                   _p(bufs["less_flat"]), _p(labels), _p(curv), _p(src), C.byref(rep))
    assert rc == 0
    out = {"full": bufs["full"][: rep.n_kept], "sharp": bufs["sharp"][: rep.n_sharp],
           "less_sharp": bufs["less_sharp"][: rep.n_less_sharp], "flat": bufs["flat"][: rep.n_flat],
           "less_flat": bufs["less_flat"][: rep.n_less_flat], "labels": labels[: rep.n_kept],
           "curvature": curv[: rep.n_kept], "src_index": src[: rep.n_kept], "report": rep}
    return out


class Odometry:
    """Oracle laserOdometry state (Aloam/src/laserOdometry.cpp globals)."""

    def __init__(self):
        self.L = lib()
        self.h = C.c_void_p(self.L.lmono_cpu_odom_create())

    def close(self):
        if self.h:
            self.L.lmono_cpu_odom_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_distortion(self, on=True):
        """#define DISTORTION 1 of Aloam/src/laserOdometry.cpp:59"""
        self.L.lmono_cpu_odom_set_distortion(self.h, 1 if on else 0)

    def step(self, sharp, less_sharp, flat, less_flat):
        a, b, c, d = (_f32(x).reshape(-1, 4) for x in (sharp, less_sharp, flat, less_flat))
        lc = Pose()
        wc = Pose()
        rep = OdomReport()
        self.L.lmono_cpu_odom_step(self.h, _p(a), len(a), _p(b), len(b), _p(c), len(c), _p(d), len(d),
                                   C.byref(lc), C.byref(wc), C.byref(rep))
        return lc.as_np(), wc.as_np(), rep


def odom_associate(sharp, flat, corner_last, surf_last, q, t):
    a, c, cl, sl = (_f32(x).reshape(-1, 4) for x in (sharp, flat, corner_last, surf_last))
    ci = np.zeros((len(a), 2), np.int32)
    pi = np.zeros((len(c), 3), np.int32)
    pose = Pose.make(q, t)
    lib().lmono_cpu_odom_associate(_p(a), len(a), _p(c), len(c), _p(cl), len(cl), _p(sl), len(sl), C.byref(pose), _p(ci), _p(pi))
    return ci, pi


def make_camera(fx=718.856, fy=718.856, cx=607.1928, cy=185.2157, k1=0.0, k2=0.0, p1=0.0, p2=0.0,
                width=1241, height=376, kernel_type=0, kernel_size=5, blur_type=0):
    """mono_lidar_mapping/config/kitti00_cam.yaml + kitti_map_config_00.yaml defaults."""
    return Camera(fx, fy, cx, cy, k1, k2, p1, p2, width, height, kernel_type, kernel_size, blur_type)


def transform_cloud(pts, T):
    pts = _f32(pts)
    T = np.ascontiguousarray(T, np.float64).reshape(12)
    out = np.zeros((len(pts), 3), np.float32)
    lib().lmono_cpu_transform_cloud(_p(pts), len(pts), pts.shape[1], _p(T), _p(out))
    return out


def project_raster(pts_cam, cam):
    pts = _f32(pts_cam)
    out = np.zeros((cam.height, cam.width), np.uint8)
    lib().lmono_cpu_project_raster(_p(pts), len(pts), pts.shape[1], C.byref(cam), _p(out))
    return out


def depth_fill(depth_raw, cam):
    d = np.ascontiguousarray(depth_raw, np.uint8)
    out = np.zeros_like(d)
    lib().lmono_cpu_depth_fill(_p(d), C.byref(cam), _p(out))
    return out


def lift_cloud(depth, bgr, cam, q, t):
    d = np.ascontiguousarray(depth, np.uint8)
    img = np.ascontiguousarray(bgr, np.uint8)
    cap = cam.width * cam.height
    cc = np.zeros((cap, 3), np.float32)
    cw = np.zeros((cap, 3), np.float32)
    rgb = np.zeros((cap, 3), np.uint8)
    n = C.c_int(0)
    pose = Pose.make(q, t)
    lib().lmono_cpu_lift_cloud(_p(d), _p(img), C.byref(cam), C.byref(pose), _p(cc), _p(cw), _p(rgb), cap, C.byref(n))
    return cc[: n.value], cw[: n.value], rgb[: n.value]
