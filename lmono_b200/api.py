"""Host-side binding of liblmono_b200.so (include/lmono.h) over ctypes.

The reference's host code is C++ inside ROS nodes (nodes/ holds the patched node sources
that call the same C ABI); this module is the Python face used by tests/, bench.py and
__graft_entry__.py.  Method names follow the reference stages:
``scan_register`` (Aloam/src/scanRegistration.cpp laserCloudHandler), ``odom_step``
(Aloam/src/laserOdometry.cpp main loop body), ``map_step`` (Aloam/src/laserMapping.cpp
process()) and ``project_color`` (mono_lidar_mapping Map_Builder::associateToMap).

There is no CPU fallback: loading fails loudly when the CUDA library has not been built,
and every call raises LmonoError when the device is unusable.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_SO = os.environ.get("LMONO_SO") or os.path.join(_CSRC, "liblmono_b200.so")      # LMONO_SO: an experiment build of the same sources


class LmonoError(RuntimeError):
    def __init__(self, code, what=""):
        self.code = code
        msg = ""
        try:
            msg = lib().lmono_strerror(code).decode()
        except Exception:
            pass
        super().__init__(f"lmono error {code} ({msg}) {what}")


class CloudView(C.Structure):
    _fields_ = [("base", C.c_void_p), ("n", C.c_int32), ("stride_bytes", C.c_int32), ("intensity_offset", C.c_int32)]


class CloudOut(C.Structure):
    _fields_ = [("base", C.c_void_p), ("capacity", C.c_int32), ("stride_bytes", C.c_int32),
                ("intensity_offset", C.c_int32), ("n_out", C.c_int32)]


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


class Params(C.Structure):
    _fields_ = [("scan_line", C.c_int32), ("minimum_range", C.c_float),
                ("mapping_line_resolution", C.c_float), ("mapping_plane_resolution", C.c_float),
                ("mapping_skip_frame", C.c_int32),
                ("max_sweep_points", C.c_int32), ("max_feature_points", C.c_int32),
                ("cube_capacity_corner", C.c_int32), ("cube_capacity_surf", C.c_int32),
                ("max_cubes_corner", C.c_int32), ("max_cubes_surf", C.c_int32),
                ("image_width", C.c_int32), ("image_height", C.c_int32), ("distortion", C.c_int32),
                ("stages", C.c_int32), ("reserved", C.c_int32 * 6)]


class SolveSummary(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("num_successful", C.c_int32), ("termination", C.c_int32),
                ("num_factors", C.c_int32), ("initial_cost", C.c_double), ("final_cost", C.c_double)]


class MapReport(C.Structure):
    _fields_ = [("corner_from_map", C.c_int32), ("surf_from_map", C.c_int32),
                ("corner_stack", C.c_int32), ("surf_stack", C.c_int32),
                ("corner_num", C.c_int32 * 2), ("surf_num", C.c_int32 * 2),
                ("optimized", C.c_int32), ("center_cube", C.c_int32 * 3), ("cen", C.c_int32 * 3),
                ("solve", SolveSummary * 2), ("ms_gpu", C.c_float)]


class OdomReport(C.Structure):
    _fields_ = [("inited", C.c_int32), ("corner_corr", C.c_int32 * 2), ("plane_corr", C.c_int32 * 2),
                ("solve", SolveSummary * 2), ("ms_gpu", C.c_float)]


class ScanReport(C.Structure):
    _fields_ = [("n_in", C.c_int32), ("n_kept", C.c_int32), ("n_sharp", C.c_int32), ("n_less_sharp", C.c_int32),
                ("n_flat", C.c_int32), ("n_less_flat", C.c_int32),
                ("ring_start", C.c_int32 * 64), ("ring_end", C.c_int32 * 64),
                ("start_ori", C.c_float), ("end_ori", C.c_float), ("ms_gpu", C.c_float)]


class Pinhole(C.Structure):
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("k1", C.c_double), ("k2", C.c_double), ("p1", C.c_double), ("p2", C.c_double),
                ("width", C.c_int32), ("height", C.c_int32), ("kernel_type", C.c_int32),
                ("kernel_size", C.c_int32), ("blur_type", C.c_int32)]


def build(force: bool = False) -> str:
    """Compile every CUDA source for sm_100a into csrc/liblmono_b200.so (nvcc cross-compiles
    without a GPU)."""
    srcs = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cu", ".cuh", "Makefile"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "lmono.h"))
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.run(["make", "-s", "-j8", "-C", _CSRC], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise ImportError(f"{_SO} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        L = C.CDLL(_SO)
        L.lmono_strerror.restype = C.c_char_p
        L.lmono_launch_count.restype = C.c_int64
        L.lmono_launch_count.argtypes = [C.c_void_p]
        L.lmono_create.argtypes = [C.c_int, C.POINTER(Params), C.c_void_p, C.POINTER(C.c_void_p)]
        _lib = L
    return _lib


def _xyzi(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] != 4:
        raise ValueError("expected float32 [n,4] x,y,z,intensity")
    return a


def view_of(a: np.ndarray) -> CloudView:
    """CloudView over a float32 [n,4] (KITTI layout) or [n,8] (pcl::PointXYZI 32-byte) array."""
    if a.dtype != np.float32 or a.ndim != 2 or not a.flags["C_CONTIGUOUS"]:
        raise ValueError("expected a C-contiguous float32 2-D array")
    if a.shape[1] == 4:
        return CloudView(a.ctypes.data, a.shape[0], 16, 12)
    if a.shape[1] == 8:
        return CloudView(a.ctypes.data, a.shape[0], 32, 16)
    if a.shape[1] == 3:
        return CloudView(a.ctypes.data, a.shape[0], 12, -1)
    raise ValueError("unsupported point width")


def _out(cap):
    buf = np.zeros((max(cap, 1), 4), np.float32)
    return buf, CloudOut(buf.ctypes.data, cap, 16, 12, 0)


class Context:
    """One lmono_ctx: owns the device-resident map and per-stage state of one sequence."""

    def __init__(self, device: int = 0, stream=None, **params):
        L = lib()
        p = Params()
        L.lmono_default_params(C.byref(p))
        for k, v in params.items():
            if not hasattr(p, k):
                raise TypeError(f"unknown parameter {k}")
            setattr(p, k, v)
        self.params = p
        self._h = C.c_void_p()
        rc = L.lmono_create(device, C.byref(p), C.c_void_p(stream) if stream else None, C.byref(self._h))
        if rc:
            raise LmonoError(rc, "lmono_create")
        self.L = L

    def close(self):
        if getattr(self, "_h", None):
            self.L.lmono_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _chk(self, rc, what):
        if rc:
            raise LmonoError(rc, what)

    # -- plumbing
    def sync(self):
        self._chk(self.L.lmono_sync(self._h), "sync")

    def launch_count(self) -> int:
        return int(self.L.lmono_launch_count(self._h))

    def stage_times(self):
        """(ms uploads + kernels, ms kernels only) of the last scan_register / odom_step / project_color call"""
        a, b = C.c_float(0), C.c_float(0)
        self._chk(self.L.lmono_stage_times(self._h, C.byref(a), C.byref(b)), "stage_times")
        return a.value, b.value

    def last_fault(self) -> int:
        bits = C.c_uint32(0)
        self._chk(self.L.lmono_last_fault(self._h, C.byref(bits)), "last_fault")
        return bits.value

    PROFILE_PHASES = ("window", "index", "voxel", "assoc", "solve", "insert", "refilter", "misc")

    def profile_enable(self, on=True):
        self._chk(self.L.lmono_profile_enable(self._h, 1 if on else 0), "profile_enable")

    def profile_read(self):
        ms = (C.c_float * 8)()
        cnt = (C.c_int32 * 8)()
        self._chk(self.L.lmono_profile_read(self._h, ms, cnt), "profile_read")
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.PROFILE_PHASES)}

    def set_concurrency_hint(self, n):
        """n sequences run side by side on this GPU outside the batch calls: n >= 4 picks the throughput kernel forms."""
        self._chk(self.L.lmono_set_concurrency_hint(self._h, C.c_int32(int(n))), "set_concurrency_hint")

    def kernel_marks_enable(self, on=True):
        """Per-launch CUDA events on every kernel of map_step* (the step then runs as plain launches, not a graph replay)."""
        self._chk(self.L.lmono_kmarks_enable(self._h, 1 if on else 0), "kmarks_enable")

    def kernel_marks(self):
        """{kernel name: (launches, total_ms)} since the marks were enabled / last read."""
        import re
        buf = C.create_string_buffer(1 << 16)
        self._chk(self.L.lmono_kmarks_dump(self._h, buf, len(buf)), "kmarks_dump")
        out = {}
        for line in buf.value.decode().splitlines():
            site, n, ms = line.split()
            f, l = site.split(":")
            name = site
            try:
                src = open(os.path.join(_CSRC, f)).read().splitlines()
                for k in range(int(l) - 1, max(int(l) - 8, -1), -1):
                    m = re.search(r"(k_\w+)(?:<\w+>)?\s*(?:<<<|,)", src[k]) if k < len(src) else None
                    if m:
                        name = m.group(1)
                        break
            except OSError:
                pass
            a = out.setdefault(name, [0, 0.0])
            a[0] += int(n)
            a[1] += float(ms)
        return {k: (v[0], v[1]) for k, v in out.items()}

    # -- laserMapping
    def map_step(self, corner_last, surf_last, q_odom, t_odom, full_res=None):
        cl = corner_last if corner_last.dtype == np.float32 and corner_last.flags["C_CONTIGUOUS"] else _xyzi(corner_last)
        sl = surf_last if surf_last.dtype == np.float32 and surf_last.flags["C_CONTIGUOUS"] else _xyzi(surf_last)
        odom = Pose.make(q_odom, t_odom)
        w = Pose()
        wm = Pose()
        rep = MapReport()
        reg = None
        if full_res is not None:
            fr = _xyzi(full_res)
            buf, reg = _out(len(fr))
            fv = view_of(fr)
        else:
            fv = CloudView(None, 0, 16, 12)
        rc = self.L.lmono_map_step(self._h, view_of(cl), view_of(sl), C.byref(odom), C.byref(w), C.byref(wm),
                                   C.byref(rep), fv, C.byref(reg) if reg is not None else None)
        self._chk(rc, "map_step")
        q, t = w.as_np()
        return q, t, rep, (buf[: reg.n_out] if reg is not None else None)

    def map_step_device(self, d_corner_ptr, n_corner, d_surf_ptr, n_surf, q_odom, t_odom):
        odom = Pose.make(q_odom, t_odom)
        self._chk(self.L.lmono_map_step_device(self._h, C.c_void_p(d_corner_ptr), n_corner, C.c_void_p(d_surf_ptr), n_surf,
                                               C.byref(odom)), "map_step_device")

    def map_collect(self):
        w = Pose()
        wm = Pose()
        rep = MapReport()
        self._chk(self.L.lmono_map_collect(self._h, C.byref(w), C.byref(wm), C.byref(rep)), "map_collect")
        q, t = w.as_np()
        return q, t, rep

    def map_get_state(self):
        p = Pose()
        cen = (C.c_int32 * 3)()
        self._chk(self.L.lmono_map_get_state(self._h, C.byref(p), cen), "map_get_state")
        q, t = p.as_np()
        return q, t, list(cen)

    def map_set_state(self, q, t):
        p = Pose.make(q, t)
        self._chk(self.L.lmono_map_set_state(self._h, C.byref(p)), "map_set_state")

    def map_import(self, which, pts):
        pts = _xyzi(pts)
        step = (1 << 21) - 1
        if len(pts) > step:
            raise ValueError("import at most 2^21-1 points per call")
        self._chk(self.L.lmono_map_import(self._h, which, view_of(pts)), "map_import")

    def map_export(self, which, scope=1):
        cap = 1 << 16
        while True:
            buf, out = _out(cap)
            rc = self.L.lmono_map_export(self._h, which, scope, C.byref(out))
            if rc == -2:
                cap = out.n_out
                continue
            self._chk(rc, "map_export")
            return buf[: out.n_out].copy()

    def map_evict(self, keep_cubes):
        n = C.c_int32(0)
        self._chk(self.L.lmono_map_evict(self._h, int(keep_cubes), C.byref(n)), "map_evict")
        return n.value

    def map_clear(self):
        self._chk(self.L.lmono_map_clear(self._h), "map_clear")

    def map_prepare_window(self, t_w_curr):
        t = (C.c_double * 3)(*[float(v) for v in t_w_curr])
        self._chk(self.L.lmono_map_prepare_window(self._h, t), "map_prepare_window")

    def knn5(self, which, queries_world):
        q = _xyzi(queries_world)
        idx = np.zeros((len(q), 5), np.int32)
        d2 = np.zeros((len(q), 5), np.float32)
        self._chk(self.L.lmono_knn5(self._h, which, view_of(q), idx.ctypes.data_as(C.c_void_p), d2.ctypes.data_as(C.c_void_p)), "knn5")
        return idx, d2

    def knn5_device(self, which, d_q_ptr, n, d_idx_ptr, d_d2_ptr):
        self._chk(self.L.lmono_knn5_device(self._h, which, C.c_void_p(d_q_ptr), n, C.c_void_p(d_idx_ptr), C.c_void_p(d_d2_ptr)), "knn5_device")

    def map_normal_eq(self, corner_stack, surf_stack, q, t):
        cs = _xyzi(corner_stack)
        ss = _xyzi(surf_stack)
        H = np.zeros(36)
        g = np.zeros(6)
        cost = C.c_double(0)
        nc = C.c_int32(0)
        ns = C.c_int32(0)
        pose = Pose.make(q, t)
        self._chk(self.L.lmono_map_normal_eq(self._h, view_of(cs), view_of(ss), C.byref(pose), H.ctypes.data_as(C.c_void_p),
                                             g.ctypes.data_as(C.c_void_p), C.byref(cost), C.byref(nc), C.byref(ns)), "map_normal_eq")
        return H.reshape(6, 6), g, cost.value, nc.value, ns.value

    # -- cube-sharded map, peer-memory mode (include/lmono.h: lmono_shard_xchg_*)
    def shard_xchg_create(self):
        """(64-byte cudaIpcMemHandle_t as bytes, device pointer) of this rank's exchange block"""
        h = C.create_string_buffer(64)
        p = C.c_void_p()
        self._chk(self.L.lmono_shard_xchg_create(self._h, h, C.byref(p)), "shard_xchg_create")
        return h.raw, p.value

    def shard_xchg_open(self, rank, nranks, handles=None, same_process_ptrs=None):
        hb = None
        if handles is not None:
            assert len(handles) == nranks and all(len(x) == 64 for x in handles)
            hb = C.create_string_buffer(b"".join(handles), 64 * nranks)
        pp = None
        if same_process_ptrs is not None:
            pp = (C.c_void_p * nranks)(*[C.c_void_p(x) if x else None for x in same_process_ptrs])
        self._chk(self.L.lmono_shard_xchg_open(self._h, rank, nranks, hb, pp), "shard_xchg_open")

    def shard_xchg_stats(self, reset=False):
        out = (C.c_uint64 * 3)()
        self._chk(self.L.lmono_shard_xchg_stats(self._h, out, 1 if reset else 0), "shard_xchg_stats")
        return {"epoch": int(out[0]), "wait_ns": int(out[1]), "exchanges": int(out[2])}

    # -- scanRegistration
    def scan_register(self, raw, want_debug=False):
        """raw: float32 [n,4] (KITTI .bin layout) or [n,3].  Returns dict of clouds, labels, report."""
        raw = np.ascontiguousarray(raw, np.float32)
        n = len(raw)
        cap = max(n, 1)
        bufs, outs = {}, {}
        for k, c in (("full", cap), ("sharp", 64 * 12), ("less_sharp", 64 * 120), ("flat", 64 * 24), ("less_flat", cap)):
            bufs[k], outs[k] = _out(c)
        labels = np.zeros(cap, np.int32)
        rep = ScanReport()
        rc = self.L.lmono_scan_register(self._h, view_of(raw), C.byref(outs["full"]), C.byref(outs["sharp"]),
                                        C.byref(outs["less_sharp"]), C.byref(outs["flat"]), C.byref(outs["less_flat"]),
                                        labels.ctypes.data_as(C.c_void_p), C.byref(rep))
        self._chk(rc, "scan_register")
        res = {k: bufs[k][: outs[k].n_out] for k in bufs}
        res["labels"] = labels[: rep.n_kept]
        res["report"] = rep
        if want_debug:
            curv = np.zeros(max(rep.n_kept, 1), np.float32)
            src = np.zeros(max(rep.n_kept, 1), np.int32)
            self._chk(self.L.lmono_scan_debug(self._h, curv.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p), rep.n_kept), "scan_debug")
            res["curvature"] = curv[: rep.n_kept]
            res["src_index"] = src[: rep.n_kept]
        return res

    # -- laserOdometry
    def sweep_step(self, raw):
        """Fused scanRegistration -> laserOdometry -> laserMapping of one raw sweep (lmono_sweep_step).
        Returns ((q, t) last_curr, (q, t) odometry w_curr, (q, t) mapped w_curr, scan report, odom report, map report)."""
        raw = np.ascontiguousarray(raw, np.float32)
        lc, ow, mw, wm = Pose(), Pose(), Pose(), Pose()
        srep, orep, mrep = ScanReport(), OdomReport(), MapReport()
        self._chk(self.L.lmono_sweep_step(self._h, view_of(raw), C.byref(lc), C.byref(ow), C.byref(mw), C.byref(wm),
                                          C.byref(srep), C.byref(orep), C.byref(mrep)), "sweep_step")
        return lc.as_np(), ow.as_np(), mw.as_np(), srep, orep, mrep

    def sweep_submit(self, raw, own_stream=True):
        """enqueue-only fused sweep (lmono_sweep_submit): `raw` (float32 [n, 3|4], C-contiguous; page-locked memory makes the
        upload asynchronous) must stay untouched until sweep_wait()"""
        if raw.dtype != np.float32 or not raw.flags["C_CONTIGUOUS"]:
            raise ValueError("sweep_submit needs a C-contiguous float32 array (it is read after the call returns)")
        self._sweep_keep = raw
        self._chk(self.L.lmono_sweep_submit(self._h, view_of(raw), 1 if own_stream else 0), "sweep_submit")

    def sweep_wait(self):
        """results of the outstanding sweep_submit(): the same tuple sweep_step returns"""
        lc, ow, mw, wm = Pose(), Pose(), Pose(), Pose()
        srep, orep, mrep = ScanReport(), OdomReport(), MapReport()
        self._chk(self.L.lmono_sweep_wait(self._h, C.byref(lc), C.byref(ow), C.byref(mw), C.byref(wm),
                                          C.byref(srep), C.byref(orep), C.byref(mrep)), "sweep_wait")
        self._sweep_keep = None
        return lc.as_np(), ow.as_np(), mw.as_np(), srep, orep, mrep

    def odom_step(self, sharp, less_sharp, flat, less_flat):
        a, b, c, d = (_xyzi(x) for x in (sharp, less_sharp, flat, less_flat))
        lc, wc, rep = Pose(), Pose(), OdomReport()
        self._chk(self.L.lmono_odom_step(self._h, view_of(a), view_of(b), view_of(c), view_of(d),
                                         C.byref(lc), C.byref(wc), C.byref(rep)), "odom_step")
        return lc.as_np(), wc.as_np(), rep

    def odom_reset(self):
        self._chk(self.L.lmono_odom_reset(self._h), "odom_reset")

    def odom_debug(self, which_pass, n_sharp, n_flat):
        ci = np.full((max(n_sharp, 1), 2), -9, np.int32)
        pi = np.full((max(n_flat, 1), 3), -9, np.int32)
        self._chk(self.L.lmono_odom_debug(self._h, which_pass, ci.ctypes.data_as(C.c_void_p), n_sharp,
                                          pi.ctypes.data_as(C.c_void_p), n_flat), "odom_debug")
        return ci[:n_sharp], pi[:n_flat]

    # -- colour projection (mono_lidar_mapping map builder)
    def project_color(self, pts, bgr, cam: "Pinhole", q, t, T_cam_lidar=None, want_cam=True, out: "ColorBuffers" = None):
        """pts: float32 [n,3|4] (camera frame, or LiDAR frame with T_cam_lidar = 3x4); bgr: uint8 [H,W,3].
        out: caller-owned output buffers to reuse (ColorBuffers; page-locked ones make the read-back a DMA)."""
        p = pts if (pts.dtype == np.float32 and pts.flags["C_CONTIGUOUS"]) else np.ascontiguousarray(pts, np.float32)
        img = bgr if (bgr.dtype == np.uint8 and bgr.flags["C_CONTIGUOUS"]) else np.ascontiguousarray(bgr, np.uint8)
        H, W = img.shape[:2]
        assert (W, H) == (cam.width, cam.height)
        npix = W * H
        b = out or ColorBuffers(W, H)
        assert (b.W, b.H) == (W, H)
        n = C.c_int32(0)
        pose = Pose.make(q, t)
        T = None
        if T_cam_lidar is not None:
            T = np.ascontiguousarray(T_cam_lidar, np.float64).reshape(12)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = self.L.lmono_project_color(self._h, view_of(p), vp(T) if T is not None else None, vp(img), C.c_int32(img.strides[0]),
                                        C.byref(cam), C.byref(pose), vp(b.raw), vp(b.fill), vp(b.cc) if want_cam else None,
                                        vp(b.cw), vp(b.rgb), npix, C.byref(n))
        self._chk(rc, "project_color")
        return {"depth_raw": b.raw, "depth": b.fill, "cloud_cam": b.cc[: n.value], "cloud_world": b.cw[: n.value], "rgb": b.rgb[: n.value]}

    def color_projection(self, n):
        """(u, v, z) of the first n points of the last project_color call (u = NaN: not rasterised)"""
        uvz = np.zeros((max(n, 1), 3), np.float32)
        self._chk(self.L.lmono_color_projection(self._h, uvz.ctypes.data_as(C.c_void_p), n), "color_projection")
        return uvz[:n]

    def voxel_grid(self, pts, leaf):
        p = _xyzi(pts)
        buf, out = _out(len(p))
        self._chk(self.L.lmono_voxel_grid(self._h, view_of(p), C.c_float(leaf), C.byref(out)), "voxel_grid")
        return buf[: out.n_out].copy()


class SweepBatch:
    """n independent sequences, one fused sweep each per call (lmono_sweep_step_batch, BASELINE config C-4): every ctx runs
    its sweep on its own stream, all are enqueued before the first is waited for."""

    def __init__(self, contexts):
        self.ctxs = list(contexts)
        n = self.n = len(self.ctxs)
        self.L = self.ctxs[0].L
        self._h = (C.c_void_p * n)(*[c._h for c in self.ctxs])
        self._views = (CloudView * n)()
        self._ow, self._mw = (Pose * n)(), (Pose * n)()
        self._srep, self._orep, self._mrep = (ScanReport * n)(), (OdomReport * n)(), (MapReport * n)()

    def step(self, raws):
        """raws: n float32 [n_i, 3|4] arrays.  Returns [((q, t) odometry, (q, t) mapped, scan, odom, map report)] * n;
        the reports are views of buffers the next call overwrites."""
        assert len(raws) == self.n
        for i, r in enumerate(raws):
            if r.dtype != np.float32 or not r.flags["C_CONTIGUOUS"]:
                raise ValueError("C-contiguous float32 arrays expected")
            self._views[i] = view_of(r)
        rc = self.L.lmono_sweep_step_batch(self._h, self.n, self._views, self._ow, self._mw, self._srep, self._orep, self._mrep)
        if rc:
            raise LmonoError(rc, "sweep_step_batch")
        return [(self._ow[i].as_np(), self._mw[i].as_np(), self._srep[i], self._orep[i], self._mrep[i]) for i in range(self.n)]


class ColorBuffers:
    """caller-owned outputs of lmono_project_color for a W x H frame (pinned=True: page-locked, via torch)"""

    def __init__(self, W, H, pinned=False):
        self.W, self.H = W, H
        npix = W * H
        if pinned:
            import torch
            self._keep = [torch.zeros(s_, dtype=d_).pin_memory() for s_, d_ in (((H, W), torch.uint8), ((H, W), torch.uint8), ((npix, 3), torch.float32),
                                                                                ((npix, 3), torch.float32), ((npix, 3), torch.uint8))]
            self.raw, self.fill, self.cc, self.cw, self.rgb = (t_.numpy() for t_ in self._keep)
        else:
            self.raw = np.zeros((H, W), np.uint8)
            self.fill = np.zeros((H, W), np.uint8)
            self.cc = np.zeros((npix, 3), np.float32)
            self.cw = np.zeros((npix, 3), np.float32)
            self.rgb = np.zeros((npix, 3), np.uint8)


class BatchArgs:
    """The per-step argument arrays of a SequenceBatch (ctypes, n entries each).  A caller that cycles through a
    fixed set of inputs builds one BatchArgs per input set once and passes it to step / step_device."""

    def __init__(self, n):
        self.n = n
        self.cv = (CloudView * n)()
        self.sv = (CloudView * n)()
        self.dc = (C.c_void_p * n)()
        self.ds = (C.c_void_p * n)()
        self.nc = (C.c_int32 * n)()
        self.ns = (C.c_int32 * n)()
        self.odom = (Pose * n)()
        self.wmap_in = (Pose * n)()
        self.use_wmap_in = False
        self._keep = None

    @staticmethod
    def _poses(arr, poses):
        for i, (q, t) in enumerate(poses):
            arr[i].q[:] = [float(v) for v in q]
            arr[i].t[:] = [float(v) for v in t]

    def set_odom(self, poses):
        """wodom_curr of the step of every sequence: [(q, t)] * n"""
        self._poses(self.odom, poses)
        return self

    def set_wmap_in(self, poses):
        """q/t_wmap_wodom to install before the step (None: keep what the previous step left)"""
        self.use_wmap_in = poses is not None
        if poses is not None:
            self._poses(self.wmap_in, poses)
        return self

    def set_device_inputs(self, corner_ptrs, corner_ns, surf_ptrs, surf_ns):
        for i in range(self.n):
            self.dc[i] = corner_ptrs[i]
            self.ds[i] = surf_ptrs[i]
            self.nc[i] = corner_ns[i]
            self.ns[i] = surf_ns[i]
        return self

    def set_host_inputs(self, corners, surfs):
        """float32 [n_i, 4] arrays (page-locked memory makes the uploads asynchronous); kept alive by this object"""
        self._keep = (list(corners), list(surfs))
        for i in range(self.n):
            self.cv[i] = view_of(corners[i])
            self.sv[i] = view_of(surfs[i])
        return self


class SequenceBatch:
    """n independent sequences on one GPU (BASELINE config C-4): one Context per sequence (sharing one stream, or
    each on its own), driven together through lmono_map_step_batch / lmono_map_step_device_batch: the n
    registrations of a step are parallel branches of one CUDA graph."""

    def __init__(self, contexts):
        self.ctxs = list(contexts)
        n = self.n = len(self.ctxs)
        assert n > 0
        self.L = self.ctxs[0].L
        self._h = (C.c_void_p * n)(*[c._h for c in self.ctxs])
        self.args = BatchArgs(n)
        self._w = (Pose * n)()
        self._wm = (Pose * n)()
        self._rep = (MapReport * n)()

    # convenience setters on the default argument set
    def set_odom(self, poses):
        self.args.set_odom(poses)

    def set_wmap_in(self, poses):
        self.args.set_wmap_in(poses)

    def set_device_inputs(self, *a):
        self.args.set_device_inputs(*a)

    def set_host_inputs(self, *a):
        self.args.set_host_inputs(*a)

    def step_device(self, join_stream=0, args=None):
        """enqueue one registration per sequence on device-resident inputs (no host synchronisation)"""
        a = args or self.args
        rc = self.L.lmono_map_step_device_batch(self._h, self.n, a.dc, a.nc, a.ds, a.ns, a.odom,
                                                a.wmap_in if a.use_wmap_in else None,
                                                C.c_void_p(join_stream) if join_stream else None)
        if rc:
            raise LmonoError(rc, "map_step_device_batch")

    def step(self, args=None):
        """one registration per sequence through the host API (upload, step, read-back): [(q, t, report)] * n.
        The reports are views of a buffer the next call overwrites."""
        a = args or self.args
        rc = self.L.lmono_map_step_batch(self._h, self.n, a.cv, a.sv, a.odom,
                                         a.wmap_in if a.use_wmap_in else None, self._w, self._wm, self._rep)
        if rc:
            raise LmonoError(rc, "map_step_batch")
        return [(np.array(self._w[i].q[:]), np.array(self._w[i].t[:]), self._rep[i]) for i in range(self.n)]

    def submit(self, args=None):
        """enqueue one registration per sequence on host inputs without waiting (lmono_map_submit_batch).  At most two
        submissions may be outstanding; page-locked inputs are fetched by the step itself and must stay untouched until
        the matching wait()."""
        a = args or self.args
        rc = self.L.lmono_map_submit_batch(self._h, self.n, a.cv, a.sv, a.odom, a.wmap_in if a.use_wmap_in else None)
        if rc:
            raise LmonoError(rc, "map_submit_batch")

    def wait(self):
        """results of the oldest outstanding submit(): [(q, t, report)] * n (report views are overwritten by the next wait)"""
        rc = self.L.lmono_map_wait_batch(self._h, self.n, self._w, self._wm, self._rep)
        if rc:
            raise LmonoError(rc, "map_wait_batch")
        return [(np.array(self._w[i].q[:]), np.array(self._w[i].t[:]), self._rep[i]) for i in range(self.n)]

    def collect(self):
        return [c.map_collect() for c in self.ctxs]

    def close(self):
        for c in self.ctxs:
            c.close()
