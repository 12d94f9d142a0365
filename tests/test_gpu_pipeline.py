"""GPU tests of the host-side streaming executor (mvsformer_b200/pipeline.py): the host-to-host results
must be bit-identical to calling the cascade directly on resident tensors, for plain samples, packed
samples and scan samples served from the per-view feature cache (including evictions)."""
import pytest
import torch

from mvsformer_b200 import synthetic as S
from mvsformer_b200.mvsformer_model import CascadeMVS
from mvsformer_b200.pipeline import PackedSample, ScanSample, StreamedCascade
from tests.helpers import CASCADE_ARGS

pytestmark = pytest.mark.gpu
DEV = "cuda"
H, W, V = 128, 192, 3


def _net():
    net = CascadeMVS(dict(CASCADE_ARGS)).eval()
    full = {}
    for s in range(4):
        sd = S.fill_state_dict(net.fusions[s].state_dict(), seed=s)
        full.update({"fusions.%d.%s" % (s, k): v for k, v in sd.items()})
    net.load_state_dict(full, strict=True)
    return net.to(DEV)


def _direct(net, feats, cams, dv):
    with torch.no_grad():
        out = net({k: v.to(DEV) for k, v in feats.items()}, {k: v.to(DEV) for k, v in cams.items()}, dv.to(DEV), tmp=list(S.EVAL_TMP))
    return out["refined_depth"].cpu(), out["photometric_confidence"].cpu()


@pytest.mark.parametrize("lanes", [1, 2])
def test_streamed_cascade_matches_direct_calls(lanes):
    """lanes = 2: consecutive reference views alternate between two compute streams (CascadeLanes) — same bits."""
    net = _net()
    cams = {k: v.pin_memory() for k, v in S.make_cameras(1, V, H, W).items()}
    dv = S.make_depth_range(1).pin_memory()
    samples = [{k: v.pin_memory() for k, v in S.make_features(1, V, H, W, seed=100 + i).items()} for i in range(4)]
    want = [_direct(net, f, cams, dv) for f in samples]
    streamer = StreamedCascade(net, DEV, list(S.EVAL_TMP), lanes=lanes)
    got = [(d.clone(), c.clone()) for d, c in streamer.run((f, cams, dv) for f in samples)]
    assert len(got) == len(want)
    for (gd, gc), (wd, wc) in zip(got, want):
        assert torch.equal(gd, wd) and torch.equal(gc, wc)
    assert streamer.h2d_bytes == sum(4 * v.numel() for v in samples[0].values()) + sum(4 * v.numel() for v in cams.values()) + 4 * dv.numel()
    packed = [PackedSample(f, cams, dv) for f in samples]
    got = [(d.clone(), c.clone()) for d, c in streamer.run(iter(packed))]
    for (gd, gc), (wd, wc) in zip(got, want):
        assert torch.equal(gd, wd) and torch.equal(gc, wc)


@pytest.mark.parametrize("lanes", [1, 2])
def test_run_scan_serves_shared_views_from_the_cache(lanes):
    """A scan of 7 views, reference view i with source views i+1, i+2 (wrapping): every view crosses
    PCIe once with a large cache; with 4 slots (V + 1) views are evicted and re-uploaded; results are
    bit-identical to dense calls either way."""
    net = _net()
    nviews = 7
    cams = {k: v.pin_memory() for k, v in S.make_cameras(1, V, H, W).items()}
    dv = S.make_depth_range(1).pin_memory()
    per_view = []
    for i in range(nviews):
        f = S.make_features(1, 1, H, W, seed=500 + i)
        per_view.append({k: v[0, 0].contiguous().pin_memory() for k, v in f.items()})
    loads = []

    def load(vid):
        loads.append(vid)
        return per_view[vid]

    ids = [[i, (i + 1) % nviews, (i + 2) % nviews] for i in range(nviews)]
    want = []
    for view_ids in ids:
        dense = {k: torch.stack([per_view[v][k] for v in view_ids]).unsqueeze(0) for k in per_view[0]}
        want.append(_direct(net, dense, cams, dv))
    for capacity, expect_loads in ((16, nviews), (4, None)):
        loads.clear()
        streamer = StreamedCascade(net, DEV, list(S.EVAL_TMP), lanes=lanes)
        samples = [ScanSample(v, load, cams, dv) for v in ids]
        got = [(d.clone(), c.clone()) for d, c in streamer.run_scan(iter(samples), capacity=capacity)]
        assert len(got) == nviews
        for (gd, gc), (wd, wc) in zip(got, want):
            assert torch.equal(gd, wd) and torch.equal(gc, wc)
        if expect_loads is not None:
            assert sorted(loads) == list(range(nviews))             # each view uploaded exactly once
            assert streamer.cache.hits == 3 * nviews - nviews
        else:
            assert len(loads) > nviews                              # evictions forced re-uploads


def test_two_scans_with_the_same_view_ids_on_one_streamer():
    """View ids are per scan (the reference numbers every scan's views 0..48): a second scan on the same streamer
    must not be served the first scan's features; ``keep_cache=True`` is the explicit opt-in to continue a scan."""
    net = _net()
    cams = {k: v.pin_memory() for k, v in S.make_cameras(1, V, H, W).items()}
    dv = S.make_depth_range(1).pin_memory()
    streamer = StreamedCascade(net, DEV, list(S.EVAL_TMP))
    results = []
    for scan in range(2):
        per_view = []
        for i in range(V):
            f = S.make_features(1, 1, H, W, seed=900 + 10 * scan + i)
            per_view.append({k: v[0, 0].contiguous().pin_memory() for k, v in f.items()})
        dense = {k: torch.stack([per_view[v][k] for v in range(V)]).unsqueeze(0) for k in per_view[0]}
        want = _direct(net, dense, cams, dv)
        loads = []

        def load(vid, per_view=per_view, loads=loads):
            loads.append(vid)
            return per_view[vid]

        got = [(d.clone(), c.clone()) for d, c in streamer.run_scan(iter([ScanSample(list(range(V)), load, cams, dv)]), capacity=8)]
        assert sorted(loads) == list(range(V))                      # nothing served from the previous scan
        assert torch.equal(got[0][0], want[0]) and torch.equal(got[0][1], want[1])
        results.append(got[0][0])
    assert not torch.equal(results[0], results[1])
    # same scan continued: no uploads at all
    loads.clear()
    again = [(d.clone(), c.clone()) for d, c in streamer.run_scan(iter([ScanSample(list(range(V)), load, cams, dv)]), capacity=8,
                                                                    keep_cache=True)]
    assert loads == [] and torch.equal(again[0][0], results[1])


def test_yielded_results_survive_one_more_step():
    """ring + 1 pinned buffers: a yielded (depth, confidence) pair is still intact after the next result was requested."""
    net = _net()
    cams = {k: v.pin_memory() for k, v in S.make_cameras(1, V, H, W).items()}
    dv = S.make_depth_range(1).pin_memory()
    samples = [{k: v.pin_memory() for k, v in S.make_features(1, V, H, W, seed=700 + i).items()} for i in range(4)]
    want = [_direct(net, f, cams, dv) for f in samples]
    streamer = StreamedCascade(net, DEV, list(S.EVAL_TMP))
    held = []
    for i, (d, c) in enumerate(streamer.run((f, cams, dv) for f in samples)):
        held.append((d, c))
        if i >= 1:                                                  # the previous pair, one next() later
            assert torch.equal(held[i - 1][0], want[i - 1][0]) and torch.equal(held[i - 1][1], want[i - 1][1])


def test_cascade_lanes_match_sequential_calls():
    """CascadeLanes.submit: six reference views round-robin over two streams, each result bit-identical to a plain call."""
    from mvsformer_b200.pipeline import CascadeLanes
    net = _net()
    cams = {k: v.to(DEV) for k, v in S.make_cameras(1, V, H, W).items()}
    dv = S.make_depth_range(1).to(DEV)
    feats = [{k: v.to(DEV) for k, v in S.make_features(1, V, H, W, seed=300 + i).items()} for i in range(6)]
    with torch.no_grad():
        want = [net(f, cams, dv, tmp=list(S.EVAL_TMP))["refined_depth"].clone() for f in feats]
        torch.cuda.synchronize()
        lanes = CascadeLanes(net, DEV, 2)
        outs = [lanes.submit(f, cams, dv, tmp=list(S.EVAL_TMP))[0] for f in feats]
        lanes.join()
        torch.cuda.synchronize()
    for o, w in zip(outs, want):
        assert torch.equal(o["refined_depth"], w)
