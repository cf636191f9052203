"""The refilter (laserMapping.cpp:788-801) updates the centroids of a cube IN PLACE when the sweep's new points open no new
voxel there (mapstore.cu: k_rf_tailscan) instead of rewriting the cube through merge / scan / scatter.  Both paths must
leave the same map and the same search index: the same drive is run in two processes, LMONO_RF_INPLACE=1 and =0 (the
switch is read once per process), and every pose, report, map export and 5-NN answer is compared bit for bit; the
in-place process must actually have taken the in-place path.  The drive revisits its sweeps, so later passes see cubes
without a new voxel (in place) next to cubes that still grow (merge) and centroids that change their search cell (index of
that cube rebuilt at the start of the next step)."""
import hashlib
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import ctypes as C, hashlib, json, sys
import numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
from lmono_b200 import api
import scenario
cm, sm = scenario.small_map(half_xy=60.0, n_surf=150_000, n_corner=40_000)
ctx = api.Context(device=0)
ctx.map_import(0, cm); ctx.map_import(1, sm)
sw = scenario.sweeps(6, seed=11, n_corner=1500, n_surf=8000)
h = hashlib.sha256()
n_inplace = n_merge = n_lazy = 0
for rep in range(3):
    for (c, s, q, t, qp, tp) in sw:
        gq, gt, grep, _ = ctx.map_step(c, s, qp, tp)
        h.update(np.asarray(gq, np.float64).tobytes()); h.update(np.asarray(gt, np.float64).tobytes())
        h.update(bytes(grep.corner_num)); h.update(bytes(grep.surf_num))
        mo = (C.c_int32 * 1200)()
        ctx.L.lmono_debug_rf_meta(ctx._h, mo, 1200)
        m = np.array(mo[:]).reshape(150, 8)
        n_inplace += int((m[:, 0] >= 2).sum()); n_merge += int((m[:, 0] == 1).sum()); n_lazy += int((m[:, 0] == 3).sum())
for which in (0, 1):
    for scope in (0, 1):
        h.update(np.ascontiguousarray(ctx.map_export(which, scope)).tobytes())
ctx.map_prepare_window(gt)            # rebuilds the search index of the cubes whose centroids changed cell
qs = np.ascontiguousarray(sw[0][1][:4000], np.float32)
for which in (0, 1):
    idx, d2 = ctx.knn5(which, qs)
    h.update(np.ascontiguousarray(idx).tobytes()); h.update(np.ascontiguousarray(d2).tobytes())
ctx.close()
print(json.dumps({"sha": h.hexdigest(), "inplace": n_inplace, "merge": n_merge, "lazy_index": n_lazy}))
"""


def _run(flag):
    env = dict(os.environ, LMONO_RF_INPLACE=flag)
    out = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.gpu
def test_inplace_refilter_equals_merge_refilter():
    a, b = _run("1"), _run("0")
    assert b["inplace"] == 0 and b["merge"] > 0
    assert a["inplace"] > 0, a            # the path under test ran
    assert a["inplace"] + a["merge"] == b["merge"]
    assert a["sha"] == b["sha"], (a, b)
