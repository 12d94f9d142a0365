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


def test_streamed_cascade_matches_direct_calls():
    net = _net()
    cams = {k: v.pin_memory() for k, v in S.make_cameras(1, V, H, W).items()}
    dv = S.make_depth_range(1).pin_memory()
    samples = [{k: v.pin_memory() for k, v in S.make_features(1, V, H, W, seed=100 + i).items()} for i in range(4)]
    want = [_direct(net, f, cams, dv) for f in samples]
    streamer = StreamedCascade(net, DEV, list(S.EVAL_TMP))
    got = [(d.clone(), c.clone()) for d, c in streamer.run((f, cams, dv) for f in samples)]
    assert len(got) == len(want)
    for (gd, gc), (wd, wc) in zip(got, want):
        assert torch.equal(gd, wd) and torch.equal(gc, wc)
    assert streamer.h2d_bytes == sum(4 * v.numel() for v in samples[0].values()) + sum(4 * v.numel() for v in cams.values()) + 4 * dv.numel()
    packed = [PackedSample(f, cams, dv) for f in samples]
    got = [(d.clone(), c.clone()) for d, c in streamer.run(iter(packed))]
    for (gd, gc), (wd, wc) in zip(got, want):
        assert torch.equal(gd, wd) and torch.equal(gc, wc)


def test_run_scan_serves_shared_views_from_the_cache():
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
        streamer = StreamedCascade(net, DEV, list(S.EVAL_TMP))
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
