"""Sequence batches (BASELINE config C-4): n independent sequences driven through lmono_map_step_batch /
lmono_map_step_device_batch must give, sequence by sequence, exactly what the same sequence gives alone
through lmono_map_step (the reference runs one laserMapping process per sequence,
Aloam/src/laserMapping.cpp:931-934) -- overlapping the registrations on the device may not change a bit."""
import numpy as np
import pytest

import scenario

pytestmark = pytest.mark.gpu

NSEQ = 3
NSWEEP = 6


def _seq_sweeps(s):
    return scenario.sweeps(NSWEEP, s0=2.0 * s, seed=40 + s, dt=0.1, drot=0.5)


def _alone(gpu_ctx_factory, s):
    cm, sm = scenario.small_map()
    ctx = gpu_ctx_factory()
    ctx.map_import(0, cm)
    ctx.map_import(1, sm)
    out = []
    for (c, su, q, t, qp, tp) in _seq_sweeps(s):
        gq, gt, rep, _ = ctx.map_step(c, su, qp, tp)
        out.append((gq, gt, list(rep.corner_num), list(rep.surf_num)))
    maps = [ctx.map_export(w, 1) for w in (0, 1)]
    ctx.close()
    return out, maps


@pytest.mark.parametrize("shared_stream", [False, True])
def test_batch_equals_sequences_alone(gpu_ctx_factory, shared_stream):
    import torch
    from lmono_b200 import api
    cm, sm = scenario.small_map()
    ctxs = []
    common = torch.cuda.Stream() if shared_stream else None
    for s in range(NSEQ):
        # stream=NULL: every ctx creates its own non-blocking stream; shared: all sequences on one stream (the batch
        # graph forks / joins internally)
        c = gpu_ctx_factory(stream=common.cuda_stream) if shared_stream else gpu_ctx_factory()
        c.map_import(0, cm)
        c.map_import(1, sm)
        ctxs.append(c)
    batch = api.SequenceBatch(ctxs)
    sweeps = [_seq_sweeps(s) for s in range(NSEQ)]
    got = [[] for _ in range(NSEQ)]
    for k in range(NSWEEP):
        # host API on even sweeps (page-locked inputs), device-resident API on odd sweeps
        batch.set_odom([(sweeps[s][k][4], sweeps[s][k][5]) for s in range(NSEQ)])
        if k % 2 == 0:
            hc = [torch.from_numpy(sweeps[s][k][0]).pin_memory() for s in range(NSEQ)]
            hs = [torch.from_numpy(sweeps[s][k][1]).pin_memory() for s in range(NSEQ)]
            batch.set_host_inputs([a.numpy() for a in hc], [a.numpy() for a in hs])
            res = batch.step()
            for s in range(NSEQ):
                q, t, rep = res[s]
                got[s].append((q, t, list(rep.corner_num), list(rep.surf_num)))
        else:
            dc = [torch.from_numpy(sweeps[s][k][0]).cuda() for s in range(NSEQ)]
            ds = [torch.from_numpy(sweeps[s][k][1]).cuda() for s in range(NSEQ)]
            torch.cuda.synchronize()
            batch.set_device_inputs([a.data_ptr() for a in dc], [a.shape[0] for a in dc],
                                    [a.data_ptr() for a in ds], [a.shape[0] for a in ds])
            js = common if shared_stream else torch.cuda.Stream()
            batch.step_device(join_stream=js.cuda_stream)
            js.synchronize()                               # the join makes the caller's stream wait for the whole batch
            for s, (q, t, rep) in enumerate(batch.collect()):
                got[s].append((q, t, list(rep.corner_num), list(rep.surf_num)))
    for s in range(NSEQ):
        ref, ref_maps = _alone(gpu_ctx_factory, s)
        for k in range(NSWEEP):
            assert np.array_equal(got[s][k][0], ref[k][0]) and np.array_equal(got[s][k][1], ref[k][1]), (s, k)
            assert got[s][k][2:] == ref[k][2:], (s, k)
            assert np.linalg.norm(got[s][k][1] - sweeps[s][k][3]) < 0.05          # and the registration is good
        for w in (0, 1):
            m = ctxs[s].map_export(w, 1)
            assert m.shape == ref_maps[w].shape and np.array_equal(m.view(np.uint32), ref_maps[w].view(np.uint32)), (s, w)


def test_batch_wmap_in(gpu_ctx_factory):
    """wmap_wodom_in replaces q/t_wmap_wodom before the step, like lmono_map_set_state"""
    from lmono_b200 import api
    cm, sm = scenario.small_map()
    a = gpu_ctx_factory(); b = gpu_ctx_factory()
    for c in (a, b):
        c.map_import(0, cm)
        c.map_import(1, sm)
    sw = scenario.sweeps(2, seed=51, dt=0.1, drot=0.5)
    for (c, su, q, t, qp, tp) in sw:
        a.map_set_state([0, 0, 0, 1], [0, 0, 0])
        ra = a.map_step(c, su, qp, tp)
        batch = api.SequenceBatch([b])
        batch.set_odom([(qp, tp)])
        batch.set_wmap_in([([0, 0, 0, 1], [0, 0, 0])])
        batch.set_host_inputs([c], [su])
        rb = batch.step()[0]
        assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1])
