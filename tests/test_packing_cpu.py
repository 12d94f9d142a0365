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


def _emulate_tcz_kzf(x, w_packed_flat, nt, cin, cout, kd, depth, zc):
    """Host emulation of the MMA issue loop of conv3d_tcz_kernel<.., KZF=true> (conv3d_tcz_kzf.cu), stride 1:
    same group / slab order, same window, column, accumulate-flag and B-descriptor arithmetic, with the tensor-core
    product written as a matrix product.  x [D,H,W,Cin] (numpy), returns y [D,H,W,Cout]."""
    cs = engine.tc_channel_slice(cin)
    ch_n, nch, pd = cs // 4, cin // cs, kd // 2
    d_, h_, w_ = x.shape[:3]
    ngroups = 3 * nch
    plane = kd * nt * 16                      # bytes between the K chunks of a tap
    btap = ch_n * plane                       # bytes per kw
    bgroup = 3 * btap
    xp = np.pad(x, ((0, 0), (1, 1), (1, 1), (0, 0)))
    y = np.zeros((d_, h_, w_, cout), np.float64)
    ntiles = (cout + nt - 1) // nt
    for ct in range(ntiles):
        for z0 in range(0, d_, zc):
            nz = min(zc, d_ - z0)
            tmem = np.full((h_ * w_, zc * nt), np.nan)          # garbage until an accumulate=0 MMA writes it
            started = 0
            iz_lo, iz_hi = max(z0 - pd, 0), min(z0 + nz - 1 + pd, d_ - 1)
            for g in range(ngroups):
                kh, ch = g // nch, g % nch
                group = w_packed_flat[(ct * ngroups + g) * (bgroup // 4):(ct * ngroups + g + 1) * (bgroup // 4)]
                for iz in range(iz_lo, iz_hi + 1):
                    kz_lo, kz_hi = max(0, iz + pd - (z0 + nz - 1)), min(kd - 1, iz + pd - z0)
                    if kz_lo > kz_hi:
                        continue
                    zi_top = iz + pd - kz_lo - z0
                    dwin = (nz - 1 - zi_top) * nt
                    for kw in range(3):
                        for kk in range(cs // 8):
                            a = xp[iz, kh:kh + h_, kw:kw + w_, ch * cs + kk * 8: ch * cs + kk * 8 + 8].reshape(-1, 8)

                            def b_rows(first_row, nrows):
                                out = np.empty((nrows, 8))
                                for r in range(nrows):
                                    for k in range(8):
                                        byte = kw * btap + (2 * kk + k // 4) * plane + (first_row + r) * 16 + (k % 4) * 4
                                        out[r, k] = group[byte // 4]
                                return out
                            if g == 0 and kw == 0 and kk == 0:
                                for kz in range(kz_lo, kz_hi + 1):
                                    zi = iz + pd - kz - z0
                                    col = (nz - 1 - zi) * nt
                                    prod = a @ b_rows(kz * nt, nt).T
                                    tmem[:, col:col + nt] = prod + (tmem[:, col:col + nt] if (started >> zi) & 1 else 0.0)
                                    started |= 1 << zi
                            else:
                                n = (kz_hi - kz_lo + 1) * nt
                                tmem[:, dwin:dwin + n] += a @ b_rows(kz_lo * nt, n).T
            for zi in range(nz):
                col = (nz - 1 - zi) * nt
                ncout = min(nt, cout - ct * nt)
                y[z0 + zi, :, :, ct * nt: ct * nt + ncout] = tmem[:, col:col + ncout].reshape(h_, w_, ncout)
    return y


def test_tcz_kzf_issue_loop_and_packing():
    """The opt-in kz-fused convolution: packed-weight layout + window / column / accumulate logic of the kernel's
    MMA issue loop, emulated on the host, reproduce conv3d for every depth chunking (no accumulator is read before
    it is initialised: uninitialised TMEM is NaN here)."""
    import torch.nn.functional as F
    for cin, cout, depth in ((16, 16, 4), (64, 32, 4), (8, 8, 8)):
        kd = 3
        g = torch.Generator().manual_seed(cin + cout)
        w = engine.round_tf32(torch.randn(kd, 3, 3, cin, cout, generator=g))
        x = engine.round_tf32(torch.randn(depth, 5, 6, cin, generator=g))
        wk, nt = engine.pack_tcz_kzf_weights(w, stride2=False)
        cs = engine.tc_channel_slice(cin)
        assert wk.shape == ((cout + nt - 1) // nt, 3, cin // cs, 3, cs // 4, kd, nt, 4)
        want = F.conv3d(x.permute(3, 0, 1, 2).unsqueeze(0).double(), w.permute(4, 3, 0, 1, 2).double(), padding=1)[0]
        want = want.permute(1, 2, 3, 0).numpy()
        for zc in (1, 2, 4):
            got = _emulate_tcz_kzf(x.double().numpy(), wk.reshape(-1).double().numpy(), nt, cin, cout, kd, depth, zc)
            assert np.isfinite(got).all(), (cin, cout, zc)
            assert np.abs(got - want).max() < 1e-9, (cin, cout, zc)


def _emulate_deconv_tcz_kzf(x, flat, nt, cin, cout, kd, zc):
    """Host emulation of the MMA issue loop of deconv3d_tcz_kzf_kernel (conv3d_tcz_kzf.cu): groups (dy, channel
    slice), slabs iz, taps -> parity class, accumulators [class][slice], per-slice initialisation in group 0,
    fused N = 3 * NT windows elsewhere.  x [D,H,W,Cin] -> y [D,2H,2W,Cout] (stride (1,2,2), pad 1, output_padding (0,1,1))."""
    cs = engine.tc_channel_slice(cin)
    ch_n, nch, pd = cs // 4, cin // cs, kd // 2
    d_, h_, w_ = x.shape[:3]
    ngroups = 2 * nch
    plane = kd * nt * 16
    btap = ch_n * plane
    bgroup = 6 * btap
    xp = np.pad(x, ((0, 0), (0, 1), (0, 1), (0, 0)))          # zero row / column past the image (dy = 1, sh = 1)
    y = np.zeros((d_, 2 * h_, 2 * w_, cout))
    for ct in range((cout + nt - 1) // nt):
        for z0 in range(0, d_, zc):
            nz = min(zc, d_ - z0)
            tmem = np.full((h_ * w_, nz * 4 * nt), np.nan)
            started = 0
            iz_lo, iz_hi = max(z0 - pd, 0), min(z0 + nz - 1 + pd, d_ - 1)
            for g in range(ngroups):
                dy, ch = g // nch, g % nch
                group = flat[(ct * ngroups + g) * (bgroup // 4):(ct * ngroups + g + 1) * (bgroup // 4)]
                for iz in range(iz_lo, iz_hi + 1):
                    kz_lo, kz_hi = max(0, z0 - iz + pd), min(kd - 1, z0 + nz - 1 - iz + pd)
                    if kz_lo > kz_hi:
                        continue
                    zi_lo = iz - pd + kz_lo - z0
                    for t in range(6 if dy == 0 else 3):
                        kh = 1 + t // 3 if dy == 0 else 0
                        kw = t % 3
                        cls = (0 if kh == 1 else 2) + (0 if kw == 1 else 1)
                        sh = 1 if kw == 0 else 0
                        dwin = (cls * nz + zi_lo) * nt
                        starter = g == 0 and t in (0, 1, 3, 4)
                        for kk in range(cs // 8):
                            a = xp[iz, dy:dy + h_, sh:sh + w_, ch * cs + kk * 8: ch * cs + kk * 8 + 8].reshape(-1, 8)

                            def b_rows(first_row, nrows):
                                out = np.empty((nrows, 8))
                                for r in range(nrows):
                                    for k in range(8):
                                        byte = t * btap + (2 * kk + k // 4) * plane + (first_row + r) * 16 + (k % 4) * 4
                                        out[r, k] = group[byte // 4]
                                return out
                            if starter and kk == 0:
                                for kz in range(kz_lo, kz_hi + 1):
                                    slot = cls * nz + (iz - pd + kz - z0)
                                    prod = a @ b_rows(kz * nt, nt).T
                                    col = slot * nt
                                    tmem[:, col:col + nt] = prod + (tmem[:, col:col + nt] if (started >> slot) & 1 else 0.0)
                                    started |= 1 << slot
                            else:
                                n = (kz_hi - kz_lo + 1) * nt
                                tmem[:, dwin:dwin + n] += a @ b_rows(kz_lo * nt, n).T
            ncout = min(nt, cout - ct * nt)
            for zi in range(nz):
                for cls in range(4):
                    col = (cls * nz + zi) * nt
                    y[z0 + zi, (cls >> 1)::2, (cls & 1)::2, ct * nt: ct * nt + ncout] = \
                        tmem[:, col:col + ncout].reshape(h_, w_, ncout)
    return y


def test_deconv_tcz_kzf_issue_loop_and_packing():
    import torch.nn.functional as F
    for cin, cout, depth in ((16, 8, 4), (32, 16, 4), (64, 32, 8)):
        kd = 3
        g = torch.Generator().manual_seed(cin * 7 + cout)
        w = engine.round_tf32(torch.randn(kd, 3, 3, cin, cout, generator=g))            # packed [kd,3,3,Cin,Cout]
        x = engine.round_tf32(torch.randn(depth, 4, 5, cin, generator=g))
        wk, nt = engine.pack_tcz_kzf_deconv_weights(w)
        cs = engine.tc_channel_slice(cin)
        assert wk.shape == ((cout + nt - 1) // nt, 2, cin // cs, 6, cs // 4, kd, nt, 4)
        want = F.conv_transpose3d(x.permute(3, 0, 1, 2).unsqueeze(0).double(), w.permute(3, 4, 0, 1, 2).double(),
                                  stride=(1, 2, 2), padding=1, output_padding=(0, 1, 1))[0].permute(1, 2, 3, 0).numpy()
        for zc in (1, 2, 4):
            got = _emulate_deconv_tcz_kzf(x.double().numpy(), wk.reshape(-1).double().numpy(), nt, cin, cout, kd, zc)
            assert np.isfinite(got).all(), (cin, cout, zc)
            assert np.abs(got - want).max() < 1e-9, (cin, cout, zc)


def _emulate_tcr_khf(x, flat, nt, cin, cout, kd, rows, zc):
    """Host emulation of the MMA issue loop of conv3d_tcr_khf_kernel (conv3d_tcz_kzf.cu) for one 128-column block:
    CTAs of `rows` output rows x `zc` slices, iterations over input rows (iz, iy), accumulators [slice][row descending],
    an MMA fused over kh whenever its whole window is initialised.  x [D,H,W,Cin] with W <= 128."""
    ch_n, pd = cin // 4, kd // 2
    d_, h_, w_ = x.shape[:3]
    plane = 3 * nt * 16
    btap = ch_n * plane
    b_bytes = kd * 3 * btap
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (0, 0)))
    y = np.zeros((d_, h_, w_, cout))
    fused = unfused = 0
    for ct in range((cout + nt - 1) // nt):
        wts = flat[ct * (b_bytes // 4):(ct + 1) * (b_bytes // 4)]
        for z0 in range(0, d_, zc):
            nz = min(zc, d_ - z0)
            for y0 in range(0, h_, rows):
                nr = min(rows, h_ - y0)
                tmem = np.full((w_, rows * zc * nt), np.nan)
                started = 0
                iz_lo, iz_hi = max(z0 - pd, 0), min(z0 + nz - 1 + pd, d_ - 1)
                iy_lo, iy_hi = max(y0 - 1, 0), min(y0 + nr, h_ - 1)
                for iz in range(iz_lo, iz_hi + 1):
                    for iy in range(iy_lo, iy_hi + 1):
                        kh_lo, kh_hi = max(0, iy + 1 - (y0 + nr - 1)), min(2, iy + 1 - y0)
                        for kz in range(kd):
                            oz = iz + pd - kz
                            if oz < z0 or oz >= z0 + nz or kh_lo > kh_hi:
                                continue
                            nk = kh_hi - kh_lo + 1
                            slot0 = (oz - z0) * rows + (nr - 1 - (iy + 1 - kh_lo - y0))
                            wmask = ((1 << nk) - 1) << slot0
                            for kw in range(3):
                                for kk in range(cin // 8):
                                    a = xp[iz, iy, kw:kw + w_, kk * 8:kk * 8 + 8]

                                    def b_rows(first_row, nrows):
                                        out = np.empty((nrows, 8))
                                        for r in range(nrows):
                                            for k in range(8):
                                                byte = (kz * 3 + kw) * btap + (2 * kk + k // 4) * plane + (first_row + r) * 16 + (k % 4) * 4
                                                out[r, k] = wts[byte // 4]
                                        return out
                                    if (started & wmask) == wmask:
                                        tmem[:, slot0 * nt:(slot0 + nk) * nt] += a @ b_rows(kh_lo * nt, nk * nt).T
                                        fused += 1
                                    else:
                                        for kh in range(kh_lo, kh_hi + 1):
                                            slot = slot0 + kh - kh_lo
                                            prod = a @ b_rows(kh * nt, nt).T
                                            tmem[:, slot * nt:(slot + 1) * nt] = prod + (tmem[:, slot * nt:(slot + 1) * nt]
                                                                                         if (started >> slot) & 1 else 0.0)
                                            started |= 1 << slot
                                            unfused += 1
                ncout = min(nt, cout - ct * nt)
                for zi in range(nz):
                    for ri in range(nr):
                        col = (zi * rows + (nr - 1 - ri)) * nt
                        y[z0 + zi, y0 + ri, :, ct * nt: ct * nt + ncout] = tmem[:, col:col + ncout]
    return y, fused, unfused


def test_tcr_khf_issue_loop_and_packing():
    import torch.nn.functional as F
    for cin, cout, kd, depth, height, rows, zc in ((16, 16, 1, 3, 19, 8, 1), (16, 8, 1, 2, 8, 8, 1), (32, 32, 3, 4, 7, 2, 4),
                                                   (8, 16, 3, 2, 5, 2, 2)):
        g = torch.Generator().manual_seed(cin * 11 + cout + kd)
        w = engine.round_tf32(torch.randn(kd, 3, 3, cin, cout, generator=g))
        x = engine.round_tf32(torch.randn(depth, height, 9, cin, generator=g))
        wk, nt = engine.pack_tcr_khf_weights(w)
        assert wk.shape == ((cout + nt - 1) // nt, kd, 3, cin // 4, 3, nt, 4)
        want = F.conv3d(x.permute(3, 0, 1, 2).unsqueeze(0).double(), w.permute(4, 3, 0, 1, 2).double(),
                        padding=(kd // 2, 1, 1))[0].permute(1, 2, 3, 0).numpy()
        got, fused, unfused = _emulate_tcr_khf(x.double().numpy(), wk.reshape(-1).double().numpy(), nt, cin, cout, kd, rows, zc)
        assert np.isfinite(got).all() and np.abs(got - want).max() < 1e-9, (cin, cout, kd)
        assert fused > unfused / 2                      # most MMAs take the fused form
