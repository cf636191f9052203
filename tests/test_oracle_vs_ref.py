"""CPU tests (-m "not gpu"): the oracle against the REFERENCE'S OWN SOURCES compiled in this container.

`make -C oracle ref` compiles translation units of /root/reference where they lie (a driver #includes them; nothing is
copied) against the functional stand-ins of oracle/refstubs/ for ROS / PCL / Ceres / Eigen, into oracle/_ref/*.so.
Everything between the library calls is reference code built by this toolchain: these tests are what pins the oracle.
They run wherever oracle/_ref exists (built here; the files travel to the GPU box with the snapshot) and skip otherwise."""
import numpy as np
import pytest

import oracle_lib
from lmono_b200 import synth


def _need(name):
    if oracle_lib.ref_lib(name) is None:
        pytest.skip(f"oracle/_ref/libref_{name}.so not built (no reference tree here)")


def _bits(a):
    return a.view(np.uint32) if a.dtype == np.float32 else a


# scanRegistration.cpp:113-459.  The node's std::sort and libm calls are this toolchain's, so the oracle is asked for the
# same ones (sort_mode 1 = libstdc++ std::sort on the curvature comparator, voxel_order_mode 1 = std::sort on the voxel
# index alone inside VoxelGrid); `atan` / `sqrt` at :166 resolve to the double overloads here, the oracle's default.
@pytest.mark.parametrize("n_scans,min_range,azimuths,seed", [(16, 0.3, 900, 3), (32, 0.3, 1875, 4), (64, 5.0, 1875, 5), (64, 5.0, 1875, 6)])
def test_scan_registration_oracle_equals_reference_source(oracle, n_scans, min_range, azimuths, seed):
    _need("scanreg")
    w = synth.make_world()
    rng = np.random.default_rng(seed)
    q, t = synth.loop_pose(w, 2.0 * seed)
    raw = synth.raycast_sweep(w, q, t, n_scans, azimuths, rng)
    raw[7, 0] = np.nan                                                     # removeNaNFromPointCloud :136
    raw[11, :3] = 0.01                                                     # removeClosedPointCloud :137
    ref = oracle_lib.ref_scan_register(raw, n_scans, min_range)
    ora = oracle.scan_register(raw, n_scans, min_range, voxel_order_mode=1, sort_mode=1)
    assert len(ref["full"]) > 10000
    for k in ("full", "curvature", "labels", "sharp", "less_sharp", "flat", "less_flat"):
        assert ref[k].shape == ora[k].shape, k
        assert np.array_equal(_bits(ref[k]), _bits(ora[k])), k             # every coordinate, ring id, relTime, curvature bit and label
    assert len(ref["sharp"]) > 0 and len(ref["flat"]) > 0 and len(ref["less_flat"]) > 1000
    # the canonical orders the CUDA path implements (stable sort; voxel members summed in input order) differ from the
    # toolchain's introsort only where the reference is not well defined: same labels, same picks, centroids within fp32 summation order
    can = oracle.scan_register(raw, n_scans, min_range)
    for k in ("full", "curvature", "labels", "sharp", "less_sharp", "flat"):
        assert np.array_equal(_bits(ref[k]), _bits(can[k])), k
    assert ref["less_flat"].shape == can["less_flat"].shape
    assert np.abs(ref["less_flat"] - can["less_flat"]).max() <= 3e-5              # a few fp32 ulps at 50-100 m (summation order of the voxel members)
