"""tcgen05 tensor-core path: operand-layout probe and implicit-GEMM conv3d parity (GPU)."""
import pytest
import torch
import torch.nn.functional as F

from mvsformer_b200 import engine, synthetic as S
from tests.helpers import rel_l1

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _image(mat, lbo_rows):
    """mat [R, K] (K multiple of 4) -> no-swizzle K-major operand image [K/4][lbo_rows][4]."""
    r, k = mat.shape
    img = torch.zeros(k // 4, lbo_rows, 4)
    img[:, :r] = mat.reshape(r, k // 4, 4).permute(1, 0, 2)
    return img.contiguous()


@pytest.mark.parametrize("n", [16, 32, 64])
def test_tc_probe_layout(n):
    """D = A (128 x K) * B (n x K)^T with the plane layout [K/4][rows][4 floats]:
    rows 16 B apart, SBO = 128 B between 8-row groups, LBO = plane pitch between 16-byte K chunks."""
    g = S._gen(5 + n)
    nk = 6
    k = 8 * nk
    a = engine.round_tf32(torch.randn(128, k, generator=g))
    b = engine.round_tf32(torch.randn(n, k, generator=g))
    rows_a, rows_b = 132, n
    a_img, b_img = _image(a, rows_a), _image(b, rows_b)
    want = a.double() @ b.double().t()
    got = engine.tc_probe(a_img.to(DEV), b_img.to(DEV), rows_a * 16, 128, rows_b * 16, 128, n, nk,
                          2 * rows_a * 16, 2 * rows_b * 16).cpu()
    err = rel_l1(got, want)
    if err > 1e-5:
        # diagnose the convention before failing: try LBO/SBO swapped
        alt = engine.tc_probe(a_img.to(DEV), b_img.to(DEV), 128, rows_a * 16, 128, rows_b * 16, n, nk,
                              2 * rows_a * 16, 2 * rows_b * 16).cpu()
        print("probe mismatch: rel_l1 %.3e; swapped LBO/SBO rel_l1 %.3e" % (err, rel_l1(alt, want)))
        print("got[:4,:4]", got[:4, :4], "\nwant[:4,:4]", want[:4, :4].float())
    assert err < 1e-5
    # a row-shifted window of the same image (what the conv uses for the kw taps)
    got2 = engine.tc_probe(a_img.to(DEV).view(-1)[8:].contiguous(), b_img.to(DEV), rows_a * 16, 128, rows_b * 16, 128, n, nk,
                           2 * rows_a * 16, 2 * rows_b * 16).cpu()
    a_shift = torch.cat([a[2:], torch.zeros(2, k)], dim=0)
    assert rel_l1(got2[:126], (a_shift.double() @ b.double().t())[:126]) < 1e-5


TC_CASES = [
    # cin, cout, kd, stride, D, H, W
    (8, 16, 3, (1, 2, 2), 4, 16, 24),
    (8, 16, 3, (2, 2, 2), 8, 16, 24),
    (16, 16, 3, (1, 1, 1), 4, 10, 14),
    (16, 16, 1, (1, 1, 1), 1, 33, 47),
    (16, 8, 1, (1, 1, 1), 1, 20, 20),
    (16, 32, 3, (2, 2, 2), 8, 8, 12),
    (16, 32, 3, (1, 2, 2), 4, 9, 13),
    (32, 32, 3, (1, 1, 1), 2, 6, 10),
    (32, 64, 3, (1, 2, 2), 3, 6, 10),
    (64, 64, 3, (1, 1, 1), 2, 5, 7),
    (16, 16, 3, (1, 1, 1), 2, 3, 300),
]


@pytest.mark.parametrize("x3", [False, True])
@pytest.mark.parametrize("cin,cout,kd,stride,D,H,W", TC_CASES)
def test_conv3d_tc(cin, cout, kd, stride, D, H, W, x3):
    g = S._gen(cin * 100 + cout + kd + W)
    w = torch.randn(cout, cin, kd, 3, 3, generator=g) * (2.0 / (cin * kd * 9)) ** 0.5
    shift = 0.1 * torch.randn(cout, generator=g)
    x = torch.randn(2, cin, D, H, W, generator=g)
    want = torch.relu(F.conv3d(x.double(), w.double(), stride=stride, padding=(kd // 2, 1, 1)) + shift.double().view(1, -1, 1, 1, 1))
    skip = torch.randn(want.shape, generator=g)
    w_packed = w.permute(2, 3, 4, 1, 0).contiguous().to(DEV)
    hi, lo, nt = engine.pack_tc_weights(w_packed, x3)
    x_cl = x.permute(0, 2, 3, 4, 1).contiguous().to(DEV)
    skip_cl = skip.permute(0, 2, 3, 4, 1).contiguous().to(DEV)
    got = engine.conv3d_tc(x_cl, hi, lo, nt, cout, kd, shift.to(DEV), skip_cl, stride, relu=True)
    got = got.permute(0, 4, 1, 2, 3).cpu()
    assert got.shape == want.shape
    tol = 1e-5 if x3 else 2e-3
    assert rel_l1(got, want + skip.double()) < tol


TCD_CASES = [
    # cin, cout, kd, sd, D, H, W
    (64, 32, 3, 2, 2, 3, 5), (32, 16, 3, 2, 3, 6, 10), (16, 8, 3, 2, 4, 8, 12),
    (64, 32, 3, 1, 2, 3, 5), (32, 16, 3, 1, 4, 6, 10), (16, 8, 3, 1, 4, 8, 12), (16, 8, 1, 1, 3, 8, 12),
    (32, 16, 3, 1, 2, 5, 200),
]


@pytest.mark.parametrize("x3", [False, True])
@pytest.mark.parametrize("cin,cout,kd,sd,D,H,W", TCD_CASES)
def test_deconv3d_tc(cin, cout, kd, sd, D, H, W, x3):
    g = S._gen(cin * 7 + cout + kd + sd + W)
    w = torch.randn(cin, cout, kd, 3, 3, generator=g) * (2.0 / (cin * kd * 9 / 4)) ** 0.5
    shift = 0.1 * torch.randn(cout, generator=g)
    x = torch.randn(2, cin, D, H, W, generator=g)
    want = torch.relu(F.conv_transpose3d(x.double(), w.double(), stride=(sd, 2, 2), padding=(kd // 2, 1, 1),
                                         output_padding=(sd - 1, 1, 1)) + shift.double().view(1, -1, 1, 1, 1))
    skip = torch.randn(want.shape, generator=g)
    w_packed = w.permute(2, 3, 4, 0, 1).contiguous().to(DEV)
    hi, lo, nt = engine.pack_tc_deconv_weights(w_packed, x3)
    x_cl = x.permute(0, 2, 3, 4, 1).contiguous().to(DEV)
    skip_cl = skip.permute(0, 2, 3, 4, 1).contiguous().to(DEV)
    got = engine.deconv3d_tc(x_cl, hi, lo, nt, cout, kd, shift.to(DEV), skip_cl, sd, relu=True)
    got = got.permute(0, 4, 1, 2, 3).cpu()
    assert got.shape == want.shape
    assert rel_l1(got, want + skip.double()) < (1e-5 if x3 else 2e-3)


@pytest.mark.parametrize("mode,tol", [("tf32x3", 1e-5), ("tf32", 3e-3), ("fp32", 5e-6)])
@pytest.mark.parametrize("kind", ["CostRegNet", "CostRegNet3D"])
def test_cost_reg_precision_modes(kind, mode, tol):
    from mvsformer_b200 import config, module as M
    from oracle import mvs_oracle as O
    g = S._gen(123)
    net = getattr(M, kind)(8, 8).eval()
    sd = S.fill_state_dict(net.state_dict(), seed=6)
    net.load_state_dict(sd)
    x = torch.randn(1, 8, 8, 32, 48, generator=g)
    fn = {"CostRegNet": O.cost_reg_net, "CostRegNet3D": O.cost_reg_net_3d}[kind]
    want = fn(x.double(), {"cost_reg." + k: v.double() for k, v in sd.items()})
    old = config.conv_precision()
    try:
        config.set_conv_precision(mode)
        got = net.to(DEV)(x.to(DEV)).cpu()
    finally:
        config.set_conv_precision(old)
    assert rel_l1(got, want) < tol


def test_cascade_tf32_meets_north_star_tolerance():
    """TF32 tensor-core convolutions (the bench default): refined depth within 1e-3 relative L1 of the
    reference's golden output (BASELINE.json north_star tolerance)."""
    from mvsformer_b200 import config
    from mvsformer_b200.mvsformer_model import CascadeMVS
    from tests.helpers import STAGE_ARGS, load_golden
    g = load_golden("cascade.npz")
    height, width, batch, views = int(g["height"]), int(g["width"]), int(g["batch"]), int(g["views"])
    feats = S.make_features(batch, views, height, width, seed=int(g["feat_seed"]))
    cams = S.make_cameras(batch, views, height, width)
    dv = S.make_depth_range(batch)
    args = dict(STAGE_ARGS, ndepths=list(S.NDEPTHS), depth_interals_ratio=list(S.DEPTH_INTERVAL_RATIO), inverse_depth=True)
    net = CascadeMVS(args).eval()
    full = {}
    for s in range(4):
        sd = S.fill_state_dict(net.fusions[s].state_dict(), seed=int(g["weight_seed0"]) + s)
        full.update({"fusions.%d.%s" % (s, k): v for k, v in sd.items()})
    net.load_state_dict(full, strict=True)
    net = net.to(DEV)
    old = config.conv_precision()
    try:
        config.set_conv_precision("tf32")
        out = net({k: v.to(DEV) for k, v in feats.items()}, {k: v.to(DEV) for k, v in cams.items()}, dv.to(DEV), tmp=list(S.EVAL_TMP))
    finally:
        config.set_conv_precision(old)
    err = rel_l1(out["refined_depth"].cpu(), g["refined_depth"])
    print("tf32 refined_depth rel-L1 = %.3e" % err)
    assert err < 1e-3


def test_vis_net_tensor_core_route():
    """StageNet.vis through the fused tcgen05 kernel (csrc/vis_fused.cu: layer 1 on CUDA cores, the 16->16 and 16->8 layers
    as TF32 MMAs, 1x1 conv + sigmoid in registers) and through the FP32 kernel vs the oracle's fp32 visibility net."""
    from mvsformer_b200 import config
    from mvsformer_b200.mvsformer_model import StageNet
    from oracle import mvs_oracle as O
    from tests.helpers import STAGE_ARGS
    g = S._gen(77)
    net = StageNet(dict(STAGE_ARGS), 8, 2).eval()
    sd = S.fill_state_dict(net.state_dict(), seed=5)
    net.load_state_dict(sd)
    net = net.to(DEV)
    for (b, n, h, w) in [(1, 4, 40, 56), (2, 3, 17, 130), (1, 10, 24, 32)]:
        ent = 2.0 * torch.rand(b, n, h, w, generator=g)
        want = O.vis_weight(ent.view(b * n, 1, h, w), sd, "vis").view(b, n, h, w)
        old = config.conv_precision()
        try:
            config.set_conv_precision("tf32")
            got = net._vis_weight(ent.to(DEV)).cpu()
            config.set_conv_precision("fp32")
            exact = net._vis_weight(ent.to(DEV)).cpu()
        finally:
            config.set_conv_precision(old)
        assert rel_l1(exact, want) < 1e-5
        assert rel_l1(got, want) < 2e-3


@pytest.mark.parametrize("n,shift", [(16, 0), (16, 2), (32, 1), (64, 0)])
def test_tc_probe_a_from_tmem(n, shift):
    """A operand copied smem -> TMEM (tcgen05.cp.128x256b, row-shifted descriptor) and read from TMEM by
    tcgen05.mma: D = A[shift:shift+128] * B^T."""
    g = S._gen(50 + n + shift)
    nk = 6
    k = 8 * nk
    a = engine.round_tf32(torch.randn(132, k, generator=g))
    b = engine.round_tf32(torch.randn(n, k, generator=g))
    a_img, b_img = _image(a, 132), _image(b, n)
    want = a[shift:shift + 128].double() @ b.double().t()
    got = engine.tc_probe_ts(a_img.to(DEV), b_img.to(DEV), 132 * 16, 128, n * 16, 128, n, nk, 2 * 132 * 16, 2 * n * 16, shift * 16).cpu()
    err = rel_l1(got, want)
    if err > 1e-5:
        print("TS probe mismatch", err, got[:3, :4], want[:3, :4].float())
    assert err < 1e-5


# ---- persistent TMA-fed convolutions (csrc/conv3d_tma.cu) --------------------------------------------------------
TMA_CASES = [
    # cin, cout, kd, D, H, W, batch
    (16, 16, 3, 4, 16, 8, 1), (16, 16, 3, 4, 10, 14, 2), (16, 16, 1, 1, 33, 47, 3), (16, 8, 1, 1, 20, 20, 2),
    (32, 32, 3, 8, 6, 10, 1), (64, 64, 3, 4, 5, 7, 2), (64, 64, 3, 8, 18, 24, 1), (16, 16, 3, 2, 3, 300, 1),
    (16, 16, 3, 5, 70, 90, 1), (32, 32, 3, 4, 144, 192, 1), (16, 16, 3, 8, 100, 60, 2),
]


@pytest.mark.parametrize("cin,cout,kd,D,H,W,batch", TMA_CASES)
def test_conv3d_tma(cin, cout, kd, D, H, W, batch):
    """Persistent warp-specialised kernel vs fp64 F.conv3d: many work items per CTA (ring + TMEM double buffering wrap
    around), partial tiles, Cout tiles, kd = 1 (visibility-net layers)."""
    assert engine.tma_supported(cin, cout, D, kd)
    g = S._gen(cin * 100 + cout + kd + W + D)
    w = engine.round_tf32(torch.randn(cout, cin, kd, 3, 3, generator=g) * (2.0 / (cin * kd * 9)) ** 0.5)
    shift = 0.1 * torch.randn(cout, generator=g)
    x = engine.round_tf32(torch.randn(batch, cin, D, H, W, generator=g))
    want = torch.relu(F.conv3d(x.double(), w.double(), padding=(kd // 2, 1, 1)) + shift.double().view(1, -1, 1, 1, 1))
    skip = torch.randn(want.shape, generator=g)
    wt, nt = engine.pack_tma_weights(w.permute(2, 3, 4, 1, 0).contiguous().to(DEV))
    x_cl = x.permute(0, 2, 3, 4, 1).contiguous().to(DEV)
    skip_cl = skip.permute(0, 2, 3, 4, 1).contiguous().to(DEV)
    got = engine.conv3d_tma(x_cl, wt, nt, cout, kd, shift.to(DEV), skip_cl, relu=True).permute(0, 4, 1, 2, 3).cpu()
    assert got.shape == want.shape
    assert torch.equal(got, engine.round_tf32(got))
    assert rel_l1(got, want + skip.double()) < 5e-4
    got2 = engine.conv3d_tma(x_cl, wt, nt, cout, kd, None, None, relu=False).permute(0, 4, 1, 2, 3).cpu()
    want2 = F.conv3d(x.double(), w.double(), padding=(kd // 2, 1, 1))
    assert rel_l1(got2, want2) < 5e-4


TMA_S2_CASES = [(8, 16, 3, 4, 16, 24, 1), (8, 16, 3, 4, 35, 50, 2), (16, 32, 3, 4, 9, 13, 2), (16, 32, 3, 8, 64, 96, 1), (32, 64, 3, 3, 6, 10, 1),
                (32, 64, 3, 4, 72, 96, 1), (8, 16, 3, 8, 144, 192, 1)]


@pytest.mark.parametrize("cin,cout,kd,D,H,W,batch", TMA_S2_CASES)
def test_conv3d_tma_stride2(cin, cout, kd, D, H, W, batch):
    assert engine.tma_supported(cin, cout, D, kd, stride2=True)
    g = S._gen(cin * 100 + cout + kd + W + D)
    w = engine.round_tf32(torch.randn(cout, cin, kd, 3, 3, generator=g) * (2.0 / (cin * kd * 9)) ** 0.5)
    shift = 0.1 * torch.randn(cout, generator=g)
    x = engine.round_tf32(torch.randn(batch, cin, D, H, W, generator=g))
    want = torch.relu(F.conv3d(x.double(), w.double(), stride=(1, 2, 2), padding=(kd // 2, 1, 1)) + shift.double().view(1, -1, 1, 1, 1))
    wt, nt = engine.pack_tma_weights(w.permute(2, 3, 4, 1, 0).contiguous().to(DEV), engine.TMA_S2)
    x_cl = x.permute(0, 2, 3, 4, 1).contiguous().to(DEV)
    got = engine.conv3d_tma(x_cl, wt, nt, cout, kd, shift.to(DEV), None, True, engine.TMA_S2).permute(0, 4, 1, 2, 3).cpu()
    assert got.shape == want.shape
    assert rel_l1(got, want) < 5e-4


TMA_DECONV_CASES = [(64, 32, 3, 4, 3, 5, 2), (64, 32, 3, 8, 18, 24, 1), (32, 16, 3, 4, 6, 10, 2), (32, 16, 3, 8, 36, 48, 1), (16, 8, 3, 4, 8, 12, 2),
                    (16, 8, 3, 8, 33, 50, 1), (16, 8, 3, 1, 9, 9, 1), (16, 8, 3, 4, 144, 192, 1)]


@pytest.mark.parametrize("cin,cout,kd,D,H,W,batch", TMA_DECONV_CASES)
def test_conv3d_tma_transposed(cin, cout, kd, D, H, W, batch):
    assert engine.tma_supported(cin, cout, D, kd, transposed=True)
    g = S._gen(cin * 7 + cout + kd + W + D)
    w = engine.round_tf32(torch.randn(cin, cout, kd, 3, 3, generator=g) * (2.0 / (cin * kd * 9 / 4)) ** 0.5)
    shift = 0.1 * torch.randn(cout, generator=g)
    x = engine.round_tf32(torch.randn(batch, cin, D, H, W, generator=g))
    want = torch.relu(F.conv_transpose3d(x.double(), w.double(), stride=(1, 2, 2), padding=(kd // 2, 1, 1),
                                         output_padding=(0, 1, 1)) + shift.double().view(1, -1, 1, 1, 1))
    skip = torch.randn(want.shape, generator=g)
    wt, nt = engine.pack_tma_weights(w.permute(2, 3, 4, 0, 1).contiguous().to(DEV), engine.TMA_DECONV)
    x_cl = x.permute(0, 2, 3, 4, 1).contiguous().to(DEV)
    skip_cl = skip.permute(0, 2, 3, 4, 1).contiguous().to(DEV)
    got = engine.conv3d_tma(x_cl, wt, nt, cout, kd, shift.to(DEV), skip_cl, True, engine.TMA_DECONV).permute(0, 4, 1, 2, 3).cpu()
    assert got.shape == want.shape
    assert rel_l1(got, want + skip.double()) < 5e-4


@pytest.mark.parametrize("kd,D,H,W,batch,with_skip", [(3, 4, 8, 12, 2, True), (3, 8, 33, 50, 1, True), (3, 1, 9, 9, 1, False),
                                                       (3, 4, 144, 192, 1, True), (1, 4, 20, 28, 1, True)])
def test_conv3d_tma_prob_matches_two_kernel_route(kd, D, H, W, batch, with_skip):
    """mvs_conv3d_tma_prob (transposed 16 -> 8 with the 1x1x1 `prob` conv in its epilogue, models/module.py:575,582) is
    BIT-identical to mvs_conv3d_tma + mvs_prob_conv_cl, partial tiles included, and stays within TF32 rounding of fp64."""
    import numpy as np
    g = S._gen(900 + kd + D + H + W)
    w = engine.round_tf32(torch.randn(16, 8, kd, 3, 3, generator=g) * (2.0 / (16 * kd * 9 / 4)) ** 0.5)
    shift = 0.1 * torch.randn(8, generator=g)
    x = engine.round_tf32(torch.randn(batch, 16, D, H, W, generator=g))
    skip = torch.randn(batch, 8, D, 2 * H, 2 * W, generator=g) if with_skip else None
    pw = torch.randn(8, generator=g)
    pb = float(torch.randn(1, generator=g))
    wt, nt = engine.pack_tma_weights(w.permute(2, 3, 4, 0, 1).contiguous().to(DEV), engine.TMA_DECONV)
    x_cl = x.permute(0, 2, 3, 4, 1).contiguous().to(DEV)
    skip_cl = skip.permute(0, 2, 3, 4, 1).contiguous().to(DEV) if with_skip else None
    pw_host = np.ascontiguousarray(pw.numpy().astype(np.float32))
    y = engine.conv3d_tma(x_cl, wt, nt, 8, kd, shift.to(DEV), skip_cl, True, engine.TMA_DECONV)
    two = engine.prob_conv_cl(y, pw_host.reshape(1, 8), np.array([pb], dtype=np.float32), 1)
    one = engine.conv3d_tma_prob(x_cl, wt, kd, shift.to(DEV), skip_cl, pw_host, pb, True)
    assert one.shape == two.shape == (batch, D, 2 * H, 2 * W)
    assert torch.equal(one, two)
    ref = torch.relu(F.conv_transpose3d(x.double(), w.double(), stride=(1, 2, 2), padding=(kd // 2, 1, 1), output_padding=(0, 1, 1))
                     + shift.double().view(1, -1, 1, 1, 1))
    if with_skip:
        ref = ref + skip.double()
    want = (ref * pw.double().view(1, 8, 1, 1, 1)).sum(1) + pb
    assert float((one.cpu().double() - want).abs().mean() / want.abs().mean()) < 2e-3


def test_cost_reg_net_3d_prob_fused_is_bit_identical():
    """CostRegNet3D eval in TF32 mode: config.prob_fused on / off give the same prob_volume_pre bit for bit."""
    from mvsformer_b200 import config, module as M
    net = M.CostRegNet3D(8, 8).eval()
    net.load_state_dict(S.fill_state_dict(net.state_dict(), seed=16))
    net = net.to(DEV)
    x = torch.randn(1, 8, 4, 64, 96, generator=S._gen(17)).to(DEV)
    old = config.conv_precision()
    try:
        config.set_conv_precision("tf32")
        config.set_prob_fused(True)
        before = engine._lib.load().mvs_launch_count()
        fused = net(x)
        n_fused = engine._lib.load().mvs_launch_count() - before
        config.set_prob_fused(False)
        before = engine._lib.load().mvs_launch_count()
        plain = net(x)
        n_plain = engine._lib.load().mvs_launch_count() - before
    finally:
        config.set_conv_precision(old)
        config.set_prob_fused(True)
    assert torch.equal(fused, plain)
    assert n_fused == n_plain - 1                           # the separate `prob` kernel is gone
