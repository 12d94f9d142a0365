"""CPU checks of the host-side operand packing (pure index logic, no GPU): every packed weight lands
where the kernels' descriptor arithmetic expects it, TF32 rounding matches cvt.rna semantics, and the
packed host sample round-trips."""
import numpy as np
import torch

from mvsformer_b200 import engine
from mvsformer_b200.pipeline import PackedSample


def _w(kd, cin, cout, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(kd, 3, 3, cin, cout, generator=g)


def test_round_tf32_matches_rna():
    x = torch.tensor([1.0, 1.0 + 2 ** -11, 1.0 + 2 ** -11 + 2 ** -20, -3.14159265, 1e-30, 65504.0])
    r = engine.round_tf32(x)
    bits = r.view(torch.int32)
    assert (bits & 0x1FFF).eq(0).all()                       # low 13 mantissa bits cleared
    assert r[0] == 1.0 and r[1] == 1.0 + 2 ** -10            # tie rounds away from zero
    assert torch.all((r - x).abs() <= x.abs() * 2 ** -11 + 1e-38)


def test_pack_tc_weights_layout():
    kd, cin, cout = 3, 64, 32
    w = _w(kd, cin, cout)
    hi, lo, nt = engine.pack_tc_weights(w, x3=True)
    cs = engine.tc_channel_slice(cin)
    assert nt == 32 and hi.shape == (1, kd, 3, cin // cs, 3, cs // 4, nt, 4)
    wr = engine.round_tf32(w)
    for (kz, kh, kw, ci, co) in [(0, 0, 0, 0, 0), (2, 1, 2, 37, 31), (1, 2, 0, 63, 5)]:
        ch, q, e = ci // cs, (ci % cs) // 4, ci % 4
        assert hi[0, kz, kh, ch, kw, q, co, e] == wr[kz, kh, kw, ci, co]
        both = hi[0, kz, kh, ch, kw, q, co, e] + lo[0, kz, kh, ch, kw, q, co, e]
        assert abs(float(both - w[kz, kh, kw, ci, co])) <= abs(float(w[kz, kh, kw, ci, co])) * 2 ** -20   # hi + lo ~ fp32


def test_pack_tcz_and_tcr_layouts():
    kd, cin, cout = 3, 16, 8
    w = _w(kd, cin, cout, 1)
    wz, nt = engine.pack_tcz_weights(w, stride2=False)
    assert nt == 16 and wz.shape == (1, 3, 1, kd, 3, 4, 16, 4)
    wr = engine.round_tf32(w)
    for (kz, kh, kw, ci, co) in [(0, 0, 0, 0, 0), (2, 1, 2, 13, 7), (1, 2, 0, 5, 3)]:
        assert wz[0, kh, 0, kz, kw, ci // 4, co, ci % 4] == wr[kz, kh, kw, ci, co]
    assert wz[0, :, :, :, :, :, 8:, :].abs().sum() == 0           # Cout padded to the N tile with zeros
    wt, nt2 = engine.pack_tcr_weights(w)
    assert nt2 == 16 and wt.shape == (1, kd, 3, 3, 4, 16, 4)
    for (kz, kh, kw, ci, co) in [(2, 1, 2, 13, 7), (1, 2, 0, 5, 3)]:
        assert wt[0, kz, kh, kw, ci // 4, co, ci % 4] == wr[kz, kh, kw, ci, co]


def test_pack_deconv_layouts():
    kd, cin, cout = 3, 32, 16
    w = _w(kd, cin, cout, 2)
    wr = engine.round_tf32(w)
    taps = {0: [(1, 0), (1, 1), (1, 2), (2, 0), (2, 1), (2, 2)], 1: [(0, 0), (0, 1), (0, 2)]}
    hi, lo, nt = engine.pack_tc_deconv_weights(w, x3=False)
    wz, ntz = engine.pack_tcz_deconv_weights(w)
    assert lo is None and nt == ntz == 16
    for dy, lst in taps.items():
        for t, (kh, kw) in enumerate(lst):
            for (kz, ci, co) in [(0, 0, 0), (2, 29, 15)]:
                assert hi[0, kz, dy, 0, t, ci // 4, co, ci % 4] == wr[kz, kh, kw, ci, co]
                assert wz[0, dy, 0, kz, t, ci // 4, co, ci % 4] == wr[kz, kh, kw, ci, co]
    assert hi[0, :, 1, :, 3:].abs().sum() == 0                      # the three unused tap slots of dy = 1


def test_tcz_shape_rules():
    assert engine.tcz_supported(16, 16, 4, 3) and engine.tcz_supported(64, 64, 8, 3)
    assert not engine.tcz_supported(64, 64, 2, 3)                   # weight ring would not fit shared memory
    assert engine.tcz_supported(64, 32, 8, 3, transposed=True) and not engine.tcz_supported(8, 8, 4, 3, transposed=True)
    assert engine.tcr_supported(16, 16, 768) and engine.tcr_supported(16, 8, 1536)
    assert not engine.tcr_supported(16, 16, 192) and not engine.tcr_supported(64, 64, 768)


def test_packed_sample_roundtrip():
    feats = {"stage1": torch.randn(1, 3, 4, 2, 3), "stage2": torch.randn(1, 3, 2, 4, 6)}
    cams = {"stage1": torch.randn(1, 3, 2, 4, 4), "stage2": torch.randn(1, 3, 2, 4, 4)}
    dv = torch.arange(192, dtype=torch.float32).view(1, 192)
    try:
        ps = PackedSample(feats, cams, dv)
    except RuntimeError:                                            # pin_memory needs a CUDA runtime
        import pytest
        pytest.skip("pinned memory unavailable")
    f, c, d = ps.unpack(ps.flat)
    assert all(torch.equal(f[k], feats[k]) for k in feats) and all(torch.equal(c[k], cams[k]) for k in cams)
    assert torch.equal(d, dv)
    assert all(off % 64 == 0 for _, _, _, off, _ in ps.layout)      # 256-byte aligned views (TMA needs 16)


def test_feature_cache_lru_and_pinning():
    from mvsformer_b200.pipeline import FeatureCache
    c = FeatureCache(3)
    assert [c.reserve(v)[0] for v in "abc"] == [0, 1, 2]
    assert c.lookup("a") == 0 and c.lookup("zz") is None
    slot, evicted = c.reserve("d")                        # LRU is now b
    assert (slot, evicted) == (1, "b")
    slot, evicted = c.reserve("e", pinned=("c", "e"))     # c is LRU but pinned -> a goes
    assert (slot, evicted) == (0, "a")
    assert c.lookup("c") == 2 and c.hits == 2 and c.misses == 5
    import pytest
    with pytest.raises(RuntimeError):
        c.reserve("f", pinned=("c", "d", "e"))
    with pytest.raises(ValueError):
        FeatureCache(1)
