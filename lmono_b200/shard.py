"""Host side of the cube-sharded global map (BASELINE config C-5).

This is an extension beyond the reference: A-LOAM keeps one process-local 21x21x11 array of 50 m
cube clouds (Aloam/src/laserMapping.cpp:74-104).  Here the cubes are distributed over the ranks of
a ``torch.distributed`` group (one rank = one GPU = one ``lmono_ctx``) by a hash of the absolute
cube coordinate, every cube stored with a voxel-complete 1 m halo.  Queries and the pose are
replicated; a query is associated by the rank that owns the cube it falls in (accepted neighbours
have d2 < 1.0, laserMapping.cpp:584,652, so its 5-NN is local and exact); the only data-path
exchange is a SUM all-reduce of 35 doubles per LM evaluation
``[J^T J upper 21 | J^T r 6 | cost | n_corner | n_surf | owned map counts 2 | pad 3]``
(NCCL over NVLink on the GPU box; gloo in the CPU tests), after which every rank advances the
same deterministic trust-region controller.  Map insertions are routed by the same ownership /
halo rule on every rank, so no point exchange is needed: each rank sees all new points (they are
the replicated queries) and keeps its share.

`ShardedMapper` drives an *engine* (the eight enqueue-only ``lmono_shard_*`` C-ABI calls of
include/lmono.h wrapped by `CtxEngine`).  The sequencing, ownership arithmetic and collective
placement live here so they can be exercised on CPU with world_size 2 over gloo.
"""
from __future__ import annotations

import numpy as np

WS_DOUBLES = 35
HALO = 1.25          # LM_SHARD_HALO in csrc/common.cuh
MAX_IMPORT = (1 << 21) - 1


# ----------------------------------------------------------------------------- ownership arithmetic
def cube_coord(v):
    """(int)((v + 25.0) / 50.0), minus one if v + 25.0 < 0  (laserMapping.cpp:312-321,741-750), cen = 0."""
    v = np.asarray(v, np.float64) + 25.0
    c = np.trunc(v / 50.0).astype(np.int64)
    return np.where(v < 0, c - 1, c)


def cube_owner(gi, gj, gk, nranks):
    """lm_cube_owner of csrc/common.cuh: cyclic, (gi + 3 gj + 5 gk) mod n -- a 5x5x3 window spreads evenly over the ranks."""
    gi, gj, gk = (np.asarray(a, np.int64) for a in (gi, gj, gk))
    if nranks <= 1:
        return np.zeros(np.broadcast(gi, gj, gk).shape, np.int64)
    return np.mod(gi + 3 * gj + 5 * gk, nranks).astype(np.int64)


def keep_mask(pts, leaf, rank, nranks):
    """d_shard_keep of csrc/common.cuh in numpy: True where `rank` stores the point (owner or halo)."""
    pts = np.asarray(pts, np.float32)
    n = len(pts)
    if nranks <= 1:
        return np.ones(n, bool)
    leaf32 = np.float32(leaf)
    inv = np.float32(1.0) / leaf32
    g = np.stack([cube_coord(pts[:, a]) for a in range(3)], 1)
    v = np.floor(pts[:, :3] * inv).astype(np.float64)
    vlo, vhi = v * float(leaf32), (v + 1.0) * float(leaf32)
    lo = vlo < (50.0 * g - 25.0) + HALO
    hi = vhi > (50.0 * g + 25.0) - HALO
    keep = cube_owner(g[:, 0], g[:, 1], g[:, 2], nranks) == rank
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if dx == dy == dz == 0:
                    continue
                ok = np.ones(n, bool)
                for a, d in enumerate((dx, dy, dz)):
                    if d < 0:
                        ok &= lo[:, a]
                    elif d > 0:
                        ok &= hi[:, a]
                keep |= ok & (cube_owner(g[:, 0] + dx, g[:, 1] + dy, g[:, 2] + dz, nranks) == rank)
    return keep


def owner_of_points(pts, nranks):
    pts = np.asarray(pts, np.float32)
    return cube_owner(cube_coord(pts[:, 0]), cube_coord(pts[:, 1]), cube_coord(pts[:, 2]), nranks)


def cube_chunks(pts, max_points=MAX_IMPORT):
    """Split a point set into chunks of whole cubes (lmono_map_import fills EMPTY cubes only)."""
    pts = np.asarray(pts, np.float32)
    if len(pts) == 0:
        return
    g = np.stack([cube_coord(pts[:, a]) for a in range(3)], 1)
    key = (g[:, 0] + 4096) * (1 << 26) + (g[:, 1] + 4096) * (1 << 13) + (g[:, 2] + 4096)
    order = np.argsort(key, kind="stable")          # stable: arrival order inside a cube is preserved
    ks = key[order]
    bounds = np.flatnonzero(np.diff(ks)) + 1
    starts = np.concatenate([[0], bounds, [len(ks)]])
    lo = 0
    for i in range(1, len(starts)):
        if starts[i] - starts[lo] > max_points:
            if i - 1 == lo:
                raise ValueError("a single cube exceeds the import chunk size")
            yield pts[order[starts[lo]:starts[i - 1]]]
            lo = i - 1
    yield pts[order[starts[lo]:]]


# ----------------------------------------------------------------------------- engines
class CtxEngine:
    """The product engine: enqueue-only lmono_shard_* calls on one lmono_ctx (no CPU fallback)."""

    def __init__(self, ctx, rank, nranks, device):
        import ctypes as C
        import torch
        self.C = C
        self.ctx = ctx
        self.ws = torch.zeros(64, dtype=torch.float64, device=device)
        ctx._chk(ctx.L.lmono_shard_configure(ctx._h, rank, nranks, C.c_void_p(self.ws.data_ptr())), "shard_configure")

    def _call(self, name, *args):
        self.ctx._chk(getattr(self.ctx.L, name)(self.ctx._h, *args), name)

    def begin(self, d_corner, nc, d_surf, ns, q_odom, t_odom):
        from .api import Pose
        C = self.C
        odom = Pose.make(q_odom, t_odom)
        self._call("lmono_shard_begin", C.c_void_p(d_corner), nc, C.c_void_p(d_surf), ns, C.byref(odom))

    def gate(self):
        self._call("lmono_shard_gate")

    def associate(self):
        self._call("lmono_shard_associate")

    def lm_begin(self, k):
        self._call("lmono_shard_lm_begin", k)

    def lm_eval(self, k):
        self._call("lmono_shard_lm_eval", k)

    def lm_control(self, k):
        self._call("lmono_shard_lm_control", k)

    def end(self):
        self._call("lmono_shard_end")

    def collect(self):
        return self.ctx.map_collect()

    def import_points(self, which, pts):
        self.ctx.map_import(which, pts)


class PeerMemoryMapper:
    """One rank of the cube-sharded registration in PEER-MEMORY mode (csrc/shard.cu): the 35-double exchange is done by
    the kernels over NVLink-mapped exchange blocks, a registration is one enqueue-only call with no collective issued
    by the host.  Same ownership / halo rules and the same import path as ShardedMapper."""

    def __init__(self, ctx, rank, nranks, leaves=(0.4, 0.8)):
        self.ctx, self.rank, self.nranks, self.leaves = ctx, rank, nranks, leaves

    @classmethod
    def connect(cls, ctx, group=None, leaves=(0.4, 0.8)):
        """one process per GPU: exchange the cudaIpc handles over torch.distributed (setup only)"""
        import torch.distributed as dist
        rank, n = dist.get_rank(group), dist.get_world_size(group)
        handle, _ = ctx.shard_xchg_create()
        handles = [None] * n
        dist.all_gather_object(handles, handle, group=group)
        ctx.shard_xchg_open(rank, n, handles=handles)
        dist.barrier(group)                  # nobody starts exchanging before every rank has mapped every block
        return cls(ctx, rank, n, leaves)

    @classmethod
    def connect_local(cls, ctxs, leaves=(0.4, 0.8)):
        """ranks = contexts of THIS process on one device (tests): plain device pointers, no IPC"""
        ptrs = [c.shard_xchg_create()[1] for c in ctxs]
        out = []
        for r, c in enumerate(ctxs):
            c.shard_xchg_open(r, len(ctxs), same_process_ptrs=[None if k == r else p for k, p in enumerate(ptrs)])
            out.append(cls(c, r, len(ctxs), leaves))
        return out

    def import_global(self, which, pts, prefilter=True):
        pts = np.ascontiguousarray(pts, np.float32)
        mine = pts[keep_mask(pts, self.leaves[which], self.rank, self.nranks)] if prefilter else pts
        for chunk in cube_chunks(mine):
            self.ctx.map_import(which, np.ascontiguousarray(chunk))
        return len(mine)

    def step(self, d_corner, nc, d_surf, ns, q_odom, t_odom):
        self.ctx.map_step_device(d_corner, nc, d_surf, ns, q_odom, t_odom)

    def collect(self):
        return self.ctx.map_collect()


class ShardedMapper:
    """One rank of the cube-sharded scan-to-map registration (laserMapping.cpp:307-801 semantics)."""

    LM_EVALS = 5          # max_num_iterations = 4 -> IterationZero + 4 candidate evaluations

    def __init__(self, engine, rank, nranks, group=None, leaves=(0.4, 0.8)):
        self.e = engine
        self.rank, self.nranks, self.group = rank, nranks, group
        self.leaves = leaves
        self.n_allreduce = 0

    @classmethod
    def on_gpu(cls, ctx, device, group=None, leaves=(0.4, 0.8), rank=None, nranks=None):
        import torch.distributed as dist
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
            nranks = dist.get_world_size(group) if dist.is_initialized() else 1
        return cls(CtxEngine(ctx, rank, nranks, device), rank, nranks, group, leaves)

    def _allreduce(self):
        if self.nranks > 1:
            import torch.distributed as dist
            dist.all_reduce(self.e.ws[:WS_DOUBLES], op=dist.ReduceOp.SUM, group=self.group)
            self.n_allreduce += 1

    def import_global(self, which, pts):
        """Every rank is handed (its view of) the global point set and keeps owner + halo points.
        The host prefilter only saves H2D traffic: the device applies the same rule again."""
        pts = np.ascontiguousarray(pts, np.float32)
        mine = pts[keep_mask(pts, self.leaves[which], self.rank, self.nranks)]
        for chunk in cube_chunks(mine):
            self.e.import_points(which, np.ascontiguousarray(chunk))
        return len(mine)

    def step_phases(self, d_corner, nc, d_surf, ns, q_odom, t_odom):
        """Generator over one registration; it yields wherever the ranks must SUM-all-reduce the
        workspace before continuing (so a test can drive several ranks of one process in lockstep)."""
        e = self.e
        e.begin(d_corner, nc, d_surf, ns, q_odom, t_odom)
        yield "gate"                         # owned map points in the window -> global :554 gate
        e.gate()
        for k in range(2):                   # laserMapping.cpp:562
            e.associate()
            e.lm_begin(k)
            for _ in range(self.LM_EVALS):
                e.lm_eval(k)
                yield "lm"                   # 35 doubles: J^T J, J^T r, cost, factor counts
                e.lm_control(k)
        e.end()

    def step(self, d_corner, nc, d_surf, ns, q_odom, t_odom):
        for _ in self.step_phases(d_corner, nc, d_surf, ns, q_odom, t_odom):
            self._allreduce()

    def collect(self):
        return self.e.collect()
