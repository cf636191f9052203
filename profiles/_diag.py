import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib
from lmono_b200 import api
from test_oracle_vs_ref import _edge_sweeps
raw, n_scans, mr = _edge_sweeps()["elevations outside the ring rule (:169-200)"]
ctx = api.Context(device=0, scan_line=n_scans, minimum_range=mr, max_cubes_corner=8, max_cubes_surf=8, cube_capacity_corner=1024, cube_capacity_surf=1024)
got = ctx.scan_register(raw, want_debug=True)
ref = oracle_lib.ref_scan_register(raw, n_scans, mr)
ora = oracle_lib.scan_register(raw, n_scans, mr)
d = np.abs(got["full"][:, 3] - ref["full"][:, 3])
print("full intensity: max", d.max(), "n>1e-4", int((d > 1e-4).sum()), "first idx", np.where(d > 1e-4)[0][:10])
print("report start/end ori gpu", got["report"].start_ori, got["report"].end_ori, "oracle", ora["report"].start_ori, ora["report"].end_ori)
bad = np.where(d > 1e-4)[0]
for i in bad[:6]:
    print(" pt", i, "gpu", got["full"][i], "ref", ref["full"][i], "src", got["src_index"][i])
dl = np.abs(got["less_flat"] - ref["less_flat"])
print("less_flat: per-column max", dl.max(axis=0), "rows>1e-4", np.where(dl.max(axis=1) > 1e-4)[0][:10])
for i in np.where(dl.max(axis=1) > 1e-4)[0][:4]:
    print(" lf", i, "gpu", got["less_flat"][i], "ref", ref["less_flat"][i], "oracle", ora["less_flat"][i])
do = np.abs(ora["less_flat"] - ref["less_flat"])
print("oracle(canonical) vs ref less_flat per-column max", do.max(axis=0))
