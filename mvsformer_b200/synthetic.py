"""Deterministic synthetic inputs for the plane-sweep cascade.

No dataset or checkpoint exists offline, so tests, the golden-vector generator and
``bench.py`` all draw their inputs from here.  Layouts follow the reference's dataset
contract (``datasets/general_eval.py:210-255``): per stage ``proj[b, v, 0]`` is the 4x4
extrinsic and ``proj[b, v, 1, :3, :3]`` the stage-scaled intrinsic; ``depth_values`` is the
``[B, 192]`` DTU-like depth range (``datasets/general_eval.py:94-104,220``).

Weights are NOT drawn from module initialisers (their RNG consumption order would tie the
fixtures to a class layout); every tensor of a ``state_dict`` is filled from a generator
seeded by its *name*, so the reference modules, the oracle and the CUDA modules can all be
given bit-identical parameters.
"""
import math
import zlib

import torch

# Cascade constants of the shipped config (configs/config_mvsformer.json:15-18).
NDEPTHS = (32, 16, 8, 4)
FEAT_CHS = (64, 32, 16, 8)            # channels of stage 1..4 features (1/8 .. 1/1 resolution)
STAGE_SCALES = (0.125, 0.25, 0.5, 1.0)
DEPTH_INTERVAL_RATIO = (4.0, 2.67, 1.5, 1.0)
GROUPS = 8
EVAL_TMP = (5.0, 5.0, 5.0, 1.0)        # README.md:147 / test.py:224-227


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed) & 0x7FFFFFFF)
    return g


def make_cameras(batch, views, height, width, dtype=torch.float32):
    """Arc of cameras looking at a point 680 mm in front of the reference view.

    Returns ``{"stage1": [B,V,2,4,4], ...}``.  K_full is DTU-like (fx=2892.33 at 1600 px
    width), E_i = [R_y(theta_i) | (I - R_y(theta_i)) (0,0,680)^T].
    """
    thetas = [0.0]
    k = 1
    while len(thetas) < views:
        thetas.append(0.10 * k)
        if len(thetas) < views:
            thetas.append(-0.10 * k)
        k += 1
    fx = 2892.33 * width / 1600.0
    fy = 2883.18 * height / 1200.0
    k_full = torch.tensor([[fx, 0.0, width / 2.0], [0.0, fy, height / 2.0], [0.0, 0.0, 1.0]], dtype=torch.float64)
    out = {}
    for s, scale in enumerate(STAGE_SCALES):
        proj = torch.zeros(batch, views, 2, 4, 4, dtype=torch.float64)
        ks = k_full.clone()
        ks[:2, :] *= scale
        for b in range(batch):
            for v, th in enumerate(thetas):
                th_b = th * (1.0 + 0.05 * b)        # slightly different rigs per batch item
                c, sn = math.cos(th_b), math.sin(th_b)
                rot = torch.tensor([[c, 0.0, sn], [0.0, 1.0, 0.0], [-sn, 0.0, c]], dtype=torch.float64)
                pivot = torch.tensor([0.0, 0.0, 680.0], dtype=torch.float64)
                ext = torch.eye(4, dtype=torch.float64)
                ext[:3, :3] = rot
                ext[:3, 3] = pivot - rot @ pivot
                # a small tilt + lateral offset so epipolar lines are not exactly horizontal
                tilt = 0.02 * v
                ct, st = math.cos(tilt), math.sin(tilt)
                rx = torch.tensor([[1.0, 0.0, 0.0], [0.0, ct, -st], [0.0, st, ct]], dtype=torch.float64)
                ext[:3, :3] = rx @ ext[:3, :3]
                ext[:3, 3] = rx @ ext[:3, 3] + torch.tensor([0.0, 3.0 * v, 0.0], dtype=torch.float64)
                proj[b, v, 0] = ext
                proj[b, v, 1, :3, :3] = ks
        out["stage%d" % (s + 1)] = proj.to(dtype)
    return out


def make_depth_range(batch, numdepth=192, dtype=torch.float32):
    """``depth_values [B, numdepth] = 425 + 2.65 k`` (DTU: 2.5 mm x 1.06)."""
    d = 425.0 + 2.65 * torch.arange(numdepth, dtype=torch.float64)
    return d.unsqueeze(0).repeat(batch, 1).to(dtype)


def make_features(batch, views, height, width, seed=1234, smooth=True, dtype=torch.float32,
                  stages=(0, 1, 2, 3), feat_chs=FEAT_CHS):
    """Per-stage NCHW feature maps ``{"stage1": [B,V,C,h,w], ...}``.

    ``smooth=True`` low-passes white noise (5x5 box) so that the correlation volume has
    non-degenerate structure; ``smooth=False`` is white noise (worst case for interpolation
    error)."""
    out = {}
    for s in stages:
        c = feat_chs[s]
        h, w = int(height * STAGE_SCALES[s]), int(width * STAGE_SCALES[s])
        x = torch.randn(batch * views, c, h, w, generator=_gen(seed + 17 * s), dtype=torch.float32)
        if smooth:
            x = torch.nn.functional.avg_pool2d(x, 5, stride=1, padding=2, count_include_pad=False) * 2.5
        out["stage%d" % (s + 1)] = x.view(batch, views, c, h, w).to(dtype).contiguous()
    return out


def fill_state_dict(state_dict, seed=0):
    """Return a new state_dict with every entry drawn from a generator keyed by its name.

    conv / deconv weights ~ N(0, 2/fan_in) (keeps activations O(1) through the U-Net);
    BN weight in [0.5, 1.5], BN bias ~ 0.1 N(0,1), running_mean ~ 0.1 N(0,1),
    running_var in [0.5, 1.5]; conv biases ~ 0.1 N(0,1); counters untouched.
    """
    out = {}
    for name, t in state_dict.items():
        g = _gen(zlib.crc32(name.encode()) + 7919 * seed)
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            out[name] = t.clone()
        elif t.dim() >= 3:                                   # conv kernels
            fan_in = t[0].numel() if t.dim() > 1 else t.numel()
            # ConvTranspose kernels are [Cin, Cout, k..]: fan-in per output is Cin*taps/stride^d; keep simple
            out[name] = (torch.randn(t.shape, generator=g) * math.sqrt(2.0 / fan_in)).to(t.dtype)
        elif leaf == "running_var":
            out[name] = (0.5 + torch.rand(t.shape, generator=g)).to(t.dtype)
        elif leaf == "weight":                               # BN affine scale
            out[name] = (0.5 + torch.rand(t.shape, generator=g)).to(t.dtype)
        else:                                                # BN bias / running_mean / conv bias
            out[name] = (0.1 * torch.randn(t.shape, generator=g)).to(t.dtype)
    return out


def stage_hw(height, width, stage_idx):
    return int(height * STAGE_SCALES[stage_idx]), int(width * STAGE_SCALES[stage_idx])


def cost_volume_algorithmic_bytes(views, height, width, batch=1, stages=(0, 1, 2, 3)):
    """4 * [(N+1) C h w + D h w + G D h w] per stage, per batch item."""
    total = 0
    for s in stages:
        h, w = stage_hw(height, width, s)
        total += 4 * (views * FEAT_CHS[s] * h * w + NDEPTHS[s] * h * w + GROUPS * NDEPTHS[s] * h * w)
    return total * batch


def voxels_per_ref_view(height, width, stages=(0, 1, 2, 3)):
    return sum(NDEPTHS[s] * stage_hw(height, width, s)[0] * stage_hw(height, width, s)[1] for s in stages)


def narrow_hypotheses(stage, height, width, batch):
    """Hypotheses a stage would see inside the cascade: stage 1 = full range, later stages a
    window around a smooth synthetic depth map (per-pixel)."""
    nd = NDEPTHS[stage]
    h, w = stage_hw(height, width, stage)
    dv = make_depth_range(batch)
    if stage == 0:
        inv_near, inv_far = 1.0 / dv[:, 0], 1.0 / dv[:, -1]
        k = torch.arange(nd, dtype=torch.float32).view(1, -1, 1, 1) / (nd - 1)
        inv = inv_far.view(-1, 1, 1, 1) + (inv_near - inv_far).view(-1, 1, 1, 1) * k
        return (1.0 / inv).repeat(1, 1, h, w)
    ys = torch.linspace(0, 1, h).view(1, h, 1)
    xs = torch.linspace(0, 1, w).view(1, 1, w)
    depth = 600.0 + 120.0 * torch.sin(3.0 * xs + 0.5) * torch.cos(2.0 * ys) + 40.0 * ys
    depth = depth.repeat(batch, 1, 1)
    itv1 = (1.0 / 425.0 - 1.0 / 931.15) / 31
    spacing = itv1 * [1.0, 0.356, 0.1526, 0.1017][stage]
    k = (torch.arange(nd, dtype=torch.float32) - (nd - 1) / 2).view(1, -1, 1, 1)
    return 1.0 / (1.0 / depth.unsqueeze(1) + k * spacing)


def make_fusion_case(views, height, width, seed=0, noise=0.3, outlier_frac=0.05):
    """Consistent depth maps of a tilted plane seen by ``views`` cameras of make_cameras (+ noise, + outlier blocks), in the
    tensor layout of the reference's fusion code (misc/fusion.py, test.py:413-431):
    ref_depth [1,1,h,w], src_depths [1,v-1,1,h,w], ref_cam [1,2,4,4], src_cams [1,v-1,2,4,4], ref_conf [1,3,1,h,w]."""
    g = _gen(seed)
    cams = make_cameras(1, views, height, width, dtype=torch.float64)["stage4"][0]            # [V,2,4,4]
    normal = torch.tensor([0.15, -0.1, 1.0], dtype=torch.float64)
    normal = normal / normal.norm()
    offset = 660.0                                                                             # plane n . X = offset
    ys, xs = torch.meshgrid(torch.arange(height, dtype=torch.float64) + 0.5, torch.arange(width, dtype=torch.float64) + 0.5,
                            indexing="ij")
    pix = torch.stack([xs, ys, torch.ones_like(xs)], dim=-1)                                   # [h,w,3]
    depths = []
    for v in range(views):
        ext, kmat = cams[v, 0], cams[v, 1, :3, :3]
        rot, trans = ext[:3, :3], ext[:3, 3]
        rays = pix @ torch.linalg.inv(kmat).T                                                  # z = 1
        nr = rot @ normal                                                                      # n^T R^T r = (R n) . r
        depth = (offset + nr @ trans) / (rays @ nr)
        depth = depth + noise * torch.randn(height, width, generator=g, dtype=torch.float64)
        bad = torch.rand(height // 4, width // 4, generator=g) < outlier_frac                  # 4x4 blocks of gross errors
        bad = bad.repeat_interleave(4, 0).repeat_interleave(4, 1)
        depth = torch.where(bad, depth * 1.2, depth)
        depths.append(depth.float())
    conf = torch.rand(1, 3, 1, height, width, generator=g)
    return {"ref_depth": depths[0].view(1, 1, height, width), "src_depths": torch.stack(depths[1:]).view(1, views - 1, 1, height, width),
            "ref_cam": cams[0:1].float().contiguous(), "src_cams": cams[1:].float().unsqueeze(0).contiguous(), "ref_conf": conf}
