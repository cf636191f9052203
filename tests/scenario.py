"""Seeded scenarios shared by the parity tests (small enough for the oracle to finish in seconds)."""
from __future__ import annotations

import functools

import numpy as np

from lmono_b200 import synth


@functools.lru_cache(maxsize=None)
def world():
    return synth.make_world()


@functools.lru_cache(maxsize=None)
def small_map(half_xy=100.0, n_surf=600_000, n_corner=150_000, s0=0.0):
    """Raw (unfiltered) world-frame samples around the pose at arc length s0."""
    w = world()
    _, t0 = synth.loop_pose(w, s0)
    cm, sm = synth.sample_map(w, t0, half_xy=half_xy, n_surf=n_surf, n_corner=n_corner)
    return cm, sm


def sweeps(n, s0=0.0, ds=1.0, seed=3, n_corner=3000, n_surf=20000, dt=0.2, drot=1.0):
    """n consecutive sweeps along the loop: (corner, surf, q_gt, t_gt, q_odom, t_odom)."""
    w = world()
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        q, t = synth.loop_pose(w, s0 + ds * k)
        c, s = synth.sample_sweep_features(w, q, t, rng, n_corner, n_surf)
        qp, tp = synth.perturb_pose(q, t, rng, dt, drot)
        out.append((c, s, q, t, qp, tp))
    return out
