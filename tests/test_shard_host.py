"""CPU tests of the multi-GPU host logic (lmono_b200/shard.py) with world_size 2 over gloo.

The product engine (CtxEngine) needs CUDA; here ShardedMapper drives an engine built from the CPU
oracle, so what is under test is the host side: the ownership hash and halo rule, the chunked
whole-cube import, the placement of the 35-double all-reduces, and that a 2-rank sharded
registration reproduces the 1-rank result (poses, factor counts, owned cubes bit for bit)."""
import os
import socket

import numpy as np
import pytest

from lmono_b200 import shard, synth


# ----------------------------------------------------------------------------- ownership arithmetic
def test_owner_is_a_partition_and_balanced():
    g = np.stack(np.meshgrid(np.arange(-10, 11), np.arange(-10, 11), np.arange(-5, 6), indexing="ij"), -1).reshape(-1, 3)
    for n in (2, 4, 8):
        o = shard.cube_owner(g[:, 0], g[:, 1], g[:, 2], n)
        assert o.min() == 0 and o.max() == n - 1
        cnt = np.bincount(o, minlength=n)
        assert cnt.min() > 0.8 * len(g) / n and cnt.max() < 1.2 * len(g) / n
        # a 5x5x3 search window is spread over the ranks, not parked on one
        w = shard.cube_owner(*np.stack(np.meshgrid(np.arange(3, 8), np.arange(-2, 3), np.arange(-1, 2), indexing="ij"), -1).reshape(-1, 3).T, n)
        assert len(np.unique(w)) == n


def test_owner_matches_the_cuda_library_symbol():
    import ctypes as C
    from lmono_b200 import api
    L = api.lib()                      # host-side symbol of the CUDA library: no device needed
    L.lmono_shard_owner_of_cube.restype = C.c_int32
    rng = np.random.default_rng(0)
    g = rng.integers(-40, 40, (500, 3))
    for n in (1, 2, 3, 8):
        ref = [L.lmono_shard_owner_of_cube(int(a), int(b), int(c), n) for a, b, c in g]
        assert list(shard.cube_owner(g[:, 0], g[:, 1], g[:, 2], n)) == ref


def test_halo_covers_every_possible_neighbour():
    """Every point within 1 m (per axis) of a cube owned by rank r is kept by r, for both leaf sizes."""
    rng = np.random.default_rng(1)
    pts = np.concatenate([rng.uniform(-130, 130, (60000, 2)), rng.uniform(-30, 30, (60000, 1)), np.zeros((60000, 1))], 1).astype(np.float32)
    # concentrate samples near cube borders (borders at 50 k - 25)
    snap = (np.round((pts[:30000, :3] + 25.0) / 50.0) * 50.0 - 25.0 + rng.uniform(-1.6, 1.6, (30000, 3))).astype(np.float32)
    pts[:30000, :3] = snap
    for n in (2, 8):
        for leaf in (0.4, 0.8):
            total = np.zeros(len(pts), int)
            for r in range(n):
                keep = shard.keep_mask(pts, leaf, r, n)
                total += keep
                # brute force: would a query inside an owned cube within |d| < 1 per axis exist?
                need = np.zeros(len(pts), bool)
                for dx in (-1.0, 0.0, 1.0):
                    for dy in (-1.0, 0.0, 1.0):
                        for dz in (-1.0, 0.0, 1.0):
                            q = pts[:, :3].astype(np.float64) + np.array([dx, dy, dz]) * 0.999
                            need |= shard.cube_owner(shard.cube_coord(q[:, 0]), shard.cube_coord(q[:, 1]), shard.cube_coord(q[:, 2]), n) == r
                assert not np.any(need & ~keep)
            assert total.min() >= 1                       # every point has an owner
            own = shard.owner_of_points(pts, n)
            assert np.array_equal(np.bincount(own, minlength=n).sum(), len(pts))


def test_cube_chunks_keep_cubes_whole_and_order_stable():
    rng = np.random.default_rng(2)
    pts = np.concatenate([rng.uniform(-120, 120, (5000, 3)), np.arange(5000)[:, None]], 1).astype(np.float32)
    chunks = list(shard.cube_chunks(pts, max_points=700))
    assert sum(len(c) for c in chunks) == len(pts) and max(len(c) for c in chunks) <= 700
    seen = set()
    for c in chunks:
        g = {tuple(x) for x in np.stack([shard.cube_coord(c[:, a]) for a in range(3)], 1)}
        assert not (g & seen)
        seen |= g
        for cube in g:                                     # arrival order inside a cube preserved
            m = np.all(np.stack([shard.cube_coord(c[:, a]) for a in range(3)], 1) == cube, axis=1)
            assert np.all(np.diff(c[m, 3]) > 0)


# ----------------------------------------------------------------------------- 2-rank registration over gloo
class OracleEngine:
    """Test double for CtxEngine: same eight calls, CPU oracle arithmetic, replicated Gauss-Newton/LM controller."""

    def __init__(self, rank, nranks):
        import torch
        import oracle_lib as O
        self.O, self.rank, self.n = O, rank, nranks
        self.m = O.Mapper(0.4, 0.8, 0, 0)
        self.ws = torch.zeros(64, dtype=torch.float64)
        self.q = np.array([0, 0, 0, 1.0]); self.t = np.zeros(3)
        self.counts = []

    def import_points(self, which, pts):
        self.m.import_points(which, pts)

    def begin(self, corner, nc, surf, ns, q_odom, t_odom):
        O = self.O
        self.q, self.t = np.array(q_odom, float), np.array(t_odom, float)      # wmap_wodom = identity in this test
        self.cs, self.ss = O.voxel_grid(corner, 0.4), O.voxel_grid(surf, 0.8)
        self.m.prepare_window(self.t)
        self.ws.zero_()
        for which in (0, 1):
            w = self.m.export(which, 0)
            self.ws[30 + which] = float((shard.owner_of_points(w, self.n) == self.rank).sum()) if len(w) else 0.0

    def gate(self):
        self.optimize = self.ws[30] > 10 and self.ws[31] > 50

    def _world(self, pts):
        R = synth.quat_to_rot(self.q)
        out = pts.copy()
        out[:, :3] = (pts[:, :3].astype(np.float64) @ R.T + self.t).astype(np.float32)
        return out

    def associate(self):
        mine_c = shard.owner_of_points(self._world(self.cs), self.n) == self.rank
        mine_s = shard.owner_of_points(self._world(self.ss), self.n) == self.rank
        self.fac, nc, ns = self.m.associate(self.cs[mine_c], self.ss[mine_s], self.q, self.t)
        self.local_counts = (nc, ns)

    def lm_begin(self, k):
        self.lam, self.phase, self.done = 1e-4, 0, not self.optimize

    def lm_eval(self, k):
        if self.done:
            return
        x = self.cand if self.phase else (self.q, self.t)
        H, g, cost = self.O.normal_eq(self.fac, x[0], x[1]) if len(self.fac) else (np.zeros((6, 6)), np.zeros(6), 0.0)
        self.ws[:21] = __import__("torch").from_numpy(H[np.triu_indices(6)].copy())
        self.ws[21:27] = __import__("torch").from_numpy(g.copy())
        self.ws[27] = cost
        self.ws[28], self.ws[29] = self.local_counts

    def _step(self):
        H = np.zeros((6, 6)); H[np.triu_indices(6)] = self.H21; H = H + H.T - np.diag(np.diag(H))
        d = -np.linalg.solve(H + self.lam * np.diag(np.diag(H)), self.g)
        dq = np.concatenate([np.sin(np.linalg.norm(d[:3])) * d[:3] / max(np.linalg.norm(d[:3]), 1e-300), [np.cos(np.linalg.norm(d[:3]))]])
        self.cand = (synth.quat_mul(dq, self.q), self.t + d[3:])

    def lm_control(self, k):
        if self.done:
            return
        w = self.ws.numpy()
        if self.phase == 0:
            self.counts.append((int(w[28]), int(w[29])))
            self.H21, self.g, self.cost, self.phase = w[:21].copy(), w[21:27].copy(), float(w[27]), 1
        elif w[27] < self.cost:
            self.q, self.t = self.cand
            self.H21, self.g, self.cost, self.lam = w[:21].copy(), w[21:27].copy(), float(w[27]), self.lam / 3
        else:
            self.lam *= 4
        self._step()

    def end(self):
        for which, st in ((0, self.cs), (1, self.ss)):
            pw = self._world(st)
            keep = shard.keep_mask(pw, (0.4, 0.8)[which], self.rank, self.n)
            self.m.import_points(which, pw[keep])

    def collect(self):
        return self.q, self.t, self.counts


def _scenario():
    w = synth.make_world()
    _, t0 = synth.loop_pose(w, 0.0)
    cm, sm = synth.sample_map(w, t0, half_xy=60.0, n_surf=40_000, n_corner=10_000)
    rng = np.random.default_rng(4)
    sweeps = []
    for k in range(2):
        q, t = synth.loop_pose(w, 30.0 * k)          # the second pose sits near a cube border region
        c, s = synth.sample_sweep_features(w, q, t, rng, 500, 2500)
        qp, tp = synth.perturb_pose(q, t, rng, 0.1, 0.5)
        sweeps.append((c.astype(np.float32), s.astype(np.float32), qp, tp))
    return cm.astype(np.float32), sm.astype(np.float32), sweeps


def _run_rank(rank, world, port, out_q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    cm, sm, sweeps = _scenario()
    eng = OracleEngine(rank, world)
    sm_ = shard.ShardedMapper(eng, rank, world)
    kept = [sm_.import_global(0, cm), sm_.import_global(1, sm)]
    poses = []
    for c, s, qp, tp in sweeps:
        sm_.step(c, len(c), s, len(s), qp, tp)
        q, t, _ = sm_.collect()
        poses.append(np.concatenate([q, t]))
    maps = [eng.m.export(0, 1), eng.m.export(1, 1)]
    out_q.put((rank, kept, np.array(poses), eng.counts, maps, sm_.n_allreduce))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def test_two_rank_sharded_registration_matches_one_rank(oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    # 1 rank (reference result of the same engine)
    p = ctx.Process(target=_run_rank, args=(0, 1, 0, q)); p.start(); single = q.get(timeout=300); p.join()
    port = _free_port()
    procs = [ctx.Process(target=_run_rank, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda r: r[0])
    for p in procs:
        p.join()
        assert p.exitcode == 0
    _, kept1, poses1, counts1, maps1, nar1 = single
    assert nar1 == 0
    for rank, kept, poses, counts, maps, nar in res:
        assert nar == 2 * (1 + 2 * 5)                                   # per registration: 1 gate + 2 x 5 LM all-reduces
        assert counts == counts1                                        # global factor counts = single-rank counts
        assert np.abs(poses - poses1).max() < 1e-9                      # same poses (sum order differs only)
        assert kept[0] < kept1[0] and kept[1] < kept1[1]                # it really holds a subset
        for which in (0, 1):                                            # owned cubes are bit-identical to the 1-rank map
            own = shard.owner_of_points(maps[which], 2) == rank
            own1 = shard.owner_of_points(maps1[which], 2) == rank
            assert np.array_equal(maps[which][own].view(np.uint32), maps1[which][own1].view(np.uint32))
    assert np.array_equal(res[0][2], res[1][2])                          # ranks agree bit for bit
    tot = sum(r[1][1] for r in res)
    assert tot < 1.6 * kept1[1]                                          # halo overhead stays moderate
