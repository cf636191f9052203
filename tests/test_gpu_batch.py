"""Sequence batches (BASELINE config C-4): n independent sequences driven through lmono_map_step_batch /
lmono_map_step_device_batch must give, sequence by sequence, exactly what the same sequence gives alone
through lmono_map_step (the reference runs one laserMapping process per sequence,
Aloam/src/laserMapping.cpp:931-934) -- overlapping the registrations on the device may not change a bit."""
import numpy as np
import pytest

import scenario

pytestmark = pytest.mark.gpu

NSEQ = 4          # >= 4 sequences: the batch runs the throughput forms of the kernels (thread-per-query kNN, 8-CTA LM clusters),
                  # the sequences alone the latency forms -- the comparison below is bit for bit across both
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


def _pinned(a):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()


def test_pipelined_submit_wait_equals_sequences_alone(gpu_ctx_factory):
    """lmono_map_submit_batch / lmono_map_wait_batch with two submissions in flight per sequence (sweep k+1 is
    submitted before sweep k is waited for; page-locked inputs fetched by the step itself): same bits as lmono_map_step."""
    import torch
    from lmono_b200 import api
    cm, sm = scenario.small_map()
    common = torch.cuda.Stream()
    ctxs = []
    for s in range(NSEQ):
        c = gpu_ctx_factory(stream=common.cuda_stream)
        c.map_import(0, cm)
        c.map_import(1, sm)
        ctxs.append(c)
    batch = api.SequenceBatch(ctxs)
    sweeps = [_seq_sweeps(s) for s in range(NSEQ)]
    keep, args = [], []
    for k in range(NSWEEP):
        a = api.BatchArgs(NSEQ)
        a.set_odom([(sweeps[s][k][4], sweeps[s][k][5]) for s in range(NSEQ)])
        hc = [_pinned(sweeps[s][k][0]) for s in range(NSEQ)]
        hs = [_pinned(sweeps[s][k][1]) for s in range(NSEQ)]
        keep.append((hc, hs))
        a.set_host_inputs([x[1] for x in hc], [x[1] for x in hs])
        args.append(a)
    got = [[] for _ in range(NSEQ)]

    def take(res):
        for s in range(NSEQ):
            q, t, rep = res[s]
            got[s].append((q, t, list(rep.corner_num), list(rep.surf_num)))

    batch.submit(args[0])
    for k in range(1, NSWEEP):
        batch.submit(args[k])            # two in flight
        with pytest.raises(api.LmonoError):
            batch.submit(args[k])        # a third is refused: both result mirrors are unread
        take(batch.wait())
    take(batch.wait())
    with pytest.raises(api.LmonoError):
        batch.wait()                     # nothing outstanding
    for s in range(NSEQ):
        ref, ref_maps = _alone(gpu_ctx_factory, s)
        for k in range(NSWEEP):
            assert np.array_equal(got[s][k][0], ref[k][0]) and np.array_equal(got[s][k][1], ref[k][1]), (s, k)
            assert got[s][k][2:] == ref[k][2:], (s, k)
        for w in (0, 1):
            m = ctxs[s].map_export(w, 1)
            assert np.array_equal(m.view(np.uint32), ref_maps[w].view(np.uint32)), (s, w)


@pytest.mark.parametrize("layout", ["xyzi32", "xyz12", "pageable", "staged_env"])
def test_fused_upload_layouts(gpu_ctx_factory, layout, monkeypatch):
    """page-locked pcl::PointXYZI records (32 B, intensity at byte 16 -- Aloam/include/aloam_velodyne/common.h:43),
    bare XYZ records and pageable memory all register exactly like the packed float4 upload"""
    if layout == "staged_env":
        monkeypatch.setenv("LMONO_NO_ZEROCOPY", "1")
    cm, sm = scenario.small_map()
    a = gpu_ctx_factory(); b = gpu_ctx_factory()
    for c in (a, b):
        c.map_import(0, cm)
        c.map_import(1, sm)
    for (c, su, q, t, qp, tp) in scenario.sweeps(2, seed=77, dt=0.1, drot=0.5):
        def conv(p):
            if layout == "xyzi32":
                w = np.zeros((len(p), 8), np.float32); w[:, :3] = p[:, :3]; w[:, 4] = p[:, 3]; w[:, 3] = 1.0; w[:, 5:] = 7.0
                return _pinned(w)
            if layout == "xyz12":
                return _pinned(p[:, :3])
            if layout == "pageable":
                return None, p.copy()
            return _pinned(p)
        pc, ps = conv(c), conv(su)
        if layout == "xyz12":     # the reference copy sees zero intensities as well
            c = c.copy(); su = su.copy(); c[:, 3] = 0; su[:, 3] = 0
        ra = a.map_step(c, su, qp, tp)
        rb = b.map_step(pc[1], ps[1], qp, tp)
        assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1])
        assert list(ra[2].corner_num) == list(rb[2].corner_num) and list(ra[2].surf_num) == list(rb[2].surf_num)
    for w in (0, 1):
        assert np.array_equal(a.map_export(w, 1).view(np.uint32), b.map_export(w, 1).view(np.uint32))
