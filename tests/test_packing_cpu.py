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


def test_pack_tma_weights_layout():
    """[kd,3,3,Cin,Cout] -> [Cout tiles][kh][kw][Cin/4][kd][n_tile][4]: per (tap, channel quad) the B rows are [kz][n]
    (operand of the kz-fused MMA of csrc/conv3d_tma.cu)."""
    kd, cin, cout = 3, 16, 8
    w = _w(kd, cin, cout, 1)
    wr = engine.round_tf32(w)
    for mode in (engine.TMA_S1, engine.TMA_S2, engine.TMA_DECONV):
        wt, nt = engine.pack_tma_weights(w, mode)
        assert nt == engine.tma_n_tile(cin, cout, mode) and wt.shape == (1, 3, 3, cin // 4, kd, nt, 4)
        for (kz, kh, kw, ci, co) in [(0, 0, 0, 0, 0), (2, 1, 2, 13, 7), (1, 2, 0, 5, 3)]:
            assert wt[0, kh, kw, ci // 4, kz, co, ci % 4] == wr[kz, kh, kw, ci, co]
        assert wt[..., 8:, :].abs().sum() == 0                      # Cout padded to the N tile with zeros
    wt, nt = engine.pack_tma_weights(_w(3, 32, 64, 3))             # two Cout tiles of 32
    assert nt == 32 and wt.shape == (2, 3, 3, 8, 3, 32, 4)
    assert wt[1, 1, 2, 5, 0, 7, 3] == engine.round_tf32(_w(3, 32, 64, 3))[0, 1, 2, 23, 39]


def test_pack_vis_fused_weights_layout():
    """[Cout,16,3,3] -> [kw][4 quads][rows][4], rows = [kh][cout] zero-padded (operand of the kh-fused MMA of csrc/vis_fused.cu)."""
    g = torch.Generator().manual_seed(5)
    for cout, rows in ((16, 48), (8, 32)):
        w = torch.randn(cout, 16, 3, 3, generator=g)
        p = engine.pack_vis_fused_weights(w, rows)
        wr = engine.round_tf32(w)
        assert p.shape == (3, 4, rows, 4)
        for (co, ci, kh, kw) in [(0, 0, 0, 0), (cout - 1, 15, 2, 1), (3, 6, 1, 2)]:
            assert p[kw, ci // 4, kh * cout + co, ci % 4] == wr[co, ci, kh, kw]
        assert p[:, :, 3 * cout:].abs().sum() == 0


def test_pack_deconv_layouts():
    kd, cin, cout = 3, 32, 16
    w = _w(kd, cin, cout, 2)
    wr = engine.round_tf32(w)
    taps = {0: [(1, 0), (1, 1), (1, 2), (2, 0), (2, 1), (2, 2)], 1: [(0, 0), (0, 1), (0, 2)]}
    hi, lo, nt = engine.pack_tc_deconv_weights(w, x3=False)
    assert lo is None and nt == 16
    for dy, lst in taps.items():
        for t, (kh, kw) in enumerate(lst):
            for (kz, ci, co) in [(0, 0, 0), (2, 29, 15)]:
                assert hi[0, kz, dy, 0, t, ci // 4, co, ci % 4] == wr[kz, kh, kw, ci, co]
    assert hi[0, :, 1, :, 3:].abs().sum() == 0                      # the three unused tap slots of dy = 1


def test_tma_shape_rules():
    """Every CostRegNet3D layer with depth stride 1 of the cfg-2 cascade is covered; TMEM bounds the rest."""
    for d in (4, 8):
        for cin, cout in ((16, 16), (32, 32), (64, 64)):
            assert engine.tma_supported(cin, cout, d, 3)
        for cin, cout in ((8, 16), (16, 32), (32, 64)):
            assert engine.tma_supported(cin, cout, d, 3, stride2=True)
        for cin, cout in ((64, 32), (32, 16), (16, 8)):
            assert engine.tma_supported(cin, cout, d, 3, transposed=True)
    assert not engine.tma_supported(64, 32, 16, 3, transposed=True)        # 4 parity classes x 16 slices x 16 columns > 512
    assert not engine.tma_supported(16, 16, 64, 3)                         # 64 slices x 16 columns > 512 TMEM columns
    assert not engine.tma_supported(24, 16, 4, 3) and not engine.tma_supported(16, 16, 4, 2)


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
