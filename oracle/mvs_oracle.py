"""CPU oracle for the plane-sweep hot path of ewrfcas/MVSFormer.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU, the arithmetic of the reference's per-reference-view cascade
(homography warp -> group-wise correlation -> visibility-weighted aggregation -> 3D-CNN
regularisation -> temperature-softmax regression -> hypothesis re-scheduling).  It exists so the
CUDA path can be checked against it.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product package
``mvsformer_b200`` never does and raises if its CUDA library is missing.

Parity pinning: the reference ships no tests or golden vectors, so the oracle is
pinned against the reference code itself, imported unmodified from ``/root/reference`` in the
build container (``oracle/ref_import.py``), by ``oracle/make_golden.py`` -> ``tests/golden/*.npz``
and by ``tests/test_oracle_vs_reference.py`` (live, skipped where the reference is absent).

Arithmetic that the reference delegates to torch (conv3d / conv_transpose3d / batch-norm) is
delegated to the same torch CPU primitives here; everything the reference composes itself
(projection algebra, sampling grid, bilinear gather, correlation, entropy, schedules, head) is
written out explicitly and differently from the reference (explicit 4-tap gather instead of
``F.grid_sample``, per-plane bilinear instead of ``F.interpolate(trilinear)``) so that it is an
independent check of the semantics.  All functions take a ``dtype`` implicitly from their
inputs: float32 reproduces the reference, float64 gives a high-precision referee.

Citations are ``file:line`` in ``/root/reference``.
"""
import math

import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# A1. cameras                                                   models/mvsformer_model.py:69-72
# ----------------------------------------------------------------------------------------------


def compose_projection(proj_pair):
    """``proj_pair [B,2,4,4]`` (extrinsic, intrinsic) -> ``P [B,4,4]`` with P[:3,:4] = K @ E[:3,:4].

    models/mvsformer_model.py:69-72 (src) and :71-72 (ref)."""
    ext = proj_pair[:, 0]
    kmat = proj_pair[:, 1, :3, :3]
    out = ext.clone()
    out[:, :3, :4] = torch.matmul(kmat, ext[:, :3, :4])
    return out


def relative_projection(src_proj, ref_proj):
    """``M = P_src @ inverse(P_ref)``; returns (R [B,3,3], t [B,3]).  models/warping.py:80-82."""
    m = torch.matmul(src_proj, torch.inverse(ref_proj))
    return m[:, :3, :3], m[:, :3, 3]


# ----------------------------------------------------------------------------------------------
# A2. warp                                                               models/warping.py:69-109
# ----------------------------------------------------------------------------------------------


def _expand_depth(depth_values, batch, height, width):
    if depth_values.dim() == 2:                                   # [B, D] -> broadcast
        return depth_values.view(batch, -1, 1, 1).expand(batch, depth_values.shape[1], height, width)
    return depth_values


def projected_pixels(rot, trans, depth_values, height, width):
    """Source-image pixel coordinates of every (d, y, x).  models/warping.py:84-93.

    q = R (x, y, 1)^T d + t ;  px = q.x / (q.z + 1e-6) ; py likewise.  Returns px, py, z each
    [B, D, h, w]."""
    batch = rot.shape[0]
    dt = rot.dtype
    depth = _expand_depth(depth_values, batch, height, width).to(dt)
    ys = torch.arange(height, dtype=dt).view(1, 1, height, 1)
    xs = torch.arange(width, dtype=dt).view(1, 1, 1, width)
    r = rot.view(batch, 3, 3, 1, 1, 1)
    # R (x, y, 1)^T, accumulated in the same order as the reference's matmul over k = 0, 1, 2
    rx = r[:, 0, 0] * xs + r[:, 0, 1] * ys + r[:, 0, 2]
    ry = r[:, 1, 0] * xs + r[:, 1, 1] * ys + r[:, 1, 2]
    rz = r[:, 2, 0] * xs + r[:, 2, 1] * ys + r[:, 2, 2]
    t = trans.view(batch, 3, 1, 1, 1)
    qx = rx * depth + t[:, 0]
    qy = ry * depth + t[:, 1]
    qz = rz * depth + t[:, 2]
    denom = qz + 1e-6
    return qx / denom, qy / denom, qz


def sample_positions(px, py, height, width):
    """The normalise / un-normalise round trip of the reference, kept in the working dtype.

    models/warping.py:94-95 (``x / ((W-1)/2) - 1``) followed by ``F.grid_sample`` with
    ``align_corners=True`` (``((g + 1) / 2) * (W - 1)``, models/warping.py:105-106)."""
    gx = px / ((width - 1) / 2) - 1
    gy = py / ((height - 1) / 2) - 1
    ix = ((gx + 1) / 2) * (width - 1)
    iy = ((gy + 1) / 2) * (height - 1)
    return ix, iy, gx, gy


def bilinear_gather(src_fea, ix, iy):
    """Zero-padded bilinear sampling, one output per (d, y, x).  models/warping.py:105-107.

    ``src_fea [B,C,h,w]``, ``ix, iy [B,D,h,w]`` -> ``[B,C,D,h,w]``.  Each of the four taps
    contributes only when it lies inside the image (padding_mode='zeros')."""
    batch, chans, height, width = src_fea.shape
    depth = ix.shape[1]
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    wx1 = ix - x0
    wy1 = iy - y0
    wx0 = (x0 + 1) - ix
    wy0 = (y0 + 1) - iy
    flat = src_fea.reshape(batch, chans, height * width)
    out = torch.zeros(batch, chans, depth, height, width, dtype=src_fea.dtype)
    for dy, wy in ((0, wy0), (1, wy1)):
        for dx, wx in ((0, wx0), (1, wx1)):
            xi = x0 + dx
            yi = y0 + dy
            inside = (xi >= 0) & (xi <= width - 1) & (yi >= 0) & (yi <= height - 1)
            # clamp before the integer cast: wild coordinates (|x| > 2^31) must not wrap
            xl = xi.clamp(0, width - 1).to(torch.int64)
            yl = yi.clamp(0, height - 1).to(torch.int64)
            idx = (yl * width + xl).view(batch, 1, -1).expand(batch, chans, -1)
            tap = torch.gather(flat, 2, idx).view(batch, chans, depth, height, width)
            wgt = (wy * wx * inside.to(src_fea.dtype)).unsqueeze(1)
            out = out + tap * wgt
    return out


def homo_warping_3D_with_mask(src_fea, src_proj, ref_proj, depth_values):
    """models/warping.py:69-109 -> (warped [B,C,D,h,w], proj_mask [B,D,h,w] bool)."""
    batch, _, height, width = src_fea.shape
    rot, trans = relative_projection(src_proj, ref_proj)
    px, py, qz = projected_pixels(rot, trans, depth_values, height, width)
    ix, iy, gx, gy = sample_positions(px, py, height, width)
    mask = (gx > 1) | (gx < -1) | (gy > 1) | (gy < -1) | (qz <= 0)        # models/warping.py:99-103
    return bilinear_gather(src_fea, ix, iy), mask


def homo_warping_3D(src_fea, src_proj, ref_proj, depth_values):
    """models/warping.py:155-189."""
    return homo_warping_3D_with_mask(src_fea, src_proj, ref_proj, depth_values)[0]


# ----------------------------------------------------------------------------------------------
# A3-A6. cost volume                                          models/mvsformer_model.py:61-105
# ----------------------------------------------------------------------------------------------


def group_correlation(ref_fea, warped, groups):
    """corr[b,g,d,y,x] = mean_{c'} ref[b,g*cpg+c',y,x] * warped[b,g*cpg+c',d,y,x].  :75-79."""
    batch, chans, depth, height, width = warped.shape
    cpg = chans // groups
    ref = ref_fea.view(batch, groups, cpg, 1, height, width)
    return (ref * warped.view(batch, groups, cpg, depth, height, width)).mean(dim=2)


def cosine_similarity_volume(ref_fea, warped, groups):
    """Eval-only similarity (:81-85): L2-normalise ref and warped ACROSS THE GROUP AXIS for each
    (c', d, y, x) with eps 1e-12 (F.normalize), multiply, mean over c', sum over g -> [B,D,h,w]."""
    batch, chans, depth, height, width = warped.shape
    cpg = chans // groups
    ref = ref_fea.view(batch, groups, cpg, 1, height, width)
    wv = warped.view(batch, groups, cpg, depth, height, width)
    ref_n = ref / ref.pow(2).sum(dim=1, keepdim=True).sqrt().clamp_min(1e-12)
    wv_n = wv / wv.pow(2).sum(dim=1, keepdim=True).sqrt().clamp_min(1e-12)
    return (ref_n * wv_n).mean(dim=2).sum(dim=1)


def view_entropy(corr):
    """:88-90.  s = sum_g corr ; p = softmax_d(s) ; e = -sum_d p log(p + 1e-7) -> [B,1,h,w]."""
    s = corr.sum(dim=1)
    p = torch.softmax(s, dim=1)
    return -(p * torch.log(p + 1e-7)).sum(dim=1, keepdim=True)


def _bn_eval(x, sd, prefix, eps=1e-5):
    shape = [1, -1] + [1] * (x.dim() - 2)
    mean = sd[prefix + ".running_mean"].to(x.dtype).view(shape)
    var = sd[prefix + ".running_var"].to(x.dtype).view(shape)
    gamma = sd[prefix + ".weight"].to(x.dtype).view(shape)
    beta = sd[prefix + ".bias"].to(x.dtype).view(shape)
    return (x - mean) / torch.sqrt(var + eps) * gamma + beta


def _bn_train(x, sd, prefix, eps=1e-5):
    shape = [1, -1] + [1] * (x.dim() - 2)
    dims = [0] + list(range(2, x.dim()))
    mean = x.mean(dim=dims, keepdim=True)
    var = x.var(dim=dims, unbiased=False, keepdim=True)
    gamma = sd[prefix + ".weight"].to(x.dtype).view(shape)
    beta = sd[prefix + ".bias"].to(x.dtype).view(shape)
    return (x - mean) / torch.sqrt(var + eps) * gamma + beta


def _bn(x, sd, prefix, training):
    return _bn_train(x, sd, prefix) if training else _bn_eval(x, sd, prefix)


def vis_weight(entropy, sd, prefix="vis", training=False):
    """Visibility net (:37, :91): 3x [conv3x3(pad1,no bias)+BN2d+ReLU] 1->16->16->8, conv1x1 8->1
    (+bias), sigmoid.  ConvBnReLU = models/module.py:168-197."""
    x = entropy
    for i in range(3):
        x = F.conv2d(x, sd["%s.%d.conv.weight" % (prefix, i)].to(x.dtype), padding=1)
        x = torch.relu(_bn(x, sd, "%s.%d.bn" % (prefix, i), training))
    x = F.conv2d(x, sd[prefix + ".3.weight"].to(x.dtype), sd[prefix + ".3.bias"].to(x.dtype))
    return torch.sigmoid(x)


def build_cost_volume(features, proj_matrices, depth_values, sd, groups=8, training=False, want_parts=False):
    """models/mvsformer_model.py:52-105 (fusion_type='cnn').

    ``features [B,V,C,h,w]``, ``proj_matrices [B,V,2,4,4]``, ``depth_values [B,D,h,w]``.
    Returns (volume_mean [B,G,D,h,w], sim_sum [B,D,h,w] or None, parts)."""
    dt = features.dtype
    ref = features[:, 0]
    ref_p = compose_projection(proj_matrices[:, 0].to(dt))
    views = features.shape[1]
    vol_sum = 0.0
    w_sum = 0.0
    sim_sum = None if training else 0.0
    parts = {"entropy": [], "weight": [], "corr": []}
    for v in range(1, views):
        src_p = compose_projection(proj_matrices[:, v].to(dt))
        warped, _ = homo_warping_3D_with_mask(features[:, v], src_p, ref_p, depth_values)
        corr = group_correlation(ref, warped, groups)
        if not training:
            sim_sum = sim_sum + cosine_similarity_volume(ref, warped, groups)
        ent = view_entropy(corr.detach())                                # :88 detaches the softmax input
        wgt = vis_weight(ent, sd, "vis", training)
        vol_sum = vol_sum + corr * wgt.unsqueeze(1)                      # :101
        w_sum = w_sum + wgt                                              # :102
        if want_parts:
            parts["entropy"].append(ent)
            parts["weight"].append(wgt)
            parts["corr"].append(corr)
    volume_mean = vol_sum / (w_sum.unsqueeze(1) + 1e-6)                  # :105
    return volume_mean, sim_sum, parts


# ----------------------------------------------------------------------------------------------
# A7. regularisers                                  models/module.py:469-505 / :550-594 / :508-547
# ----------------------------------------------------------------------------------------------


def _conv_block(x, sd, prefix, stride, training, kernel=(3, 3, 3)):
    """Conv3d block of models/module.py:83-117: conv(no bias) -> BN3d -> ReLU."""
    pad = tuple(k // 2 for k in kernel)
    x = F.conv3d(x, sd[prefix + ".conv.weight"].to(x.dtype), stride=stride, padding=pad)
    return torch.relu(_bn(x, sd, prefix + ".bn", training))


def _deconv_block(x, sd, wname, bnname, stride, out_pad, training, kernel=(3, 3, 3)):
    """Transposed block: ConvTranspose3d(k, stride, pad k//2, output_padding, no bias) -> BN3d -> ReLU.
    models/module.py:126-159 (Deconv3d) and :562-575 (nn.Sequential form)."""
    pad = tuple(k // 2 for k in kernel)
    x = F.conv_transpose3d(x, sd[wname].to(x.dtype), stride=stride, padding=pad, output_padding=out_pad)
    return torch.relu(_bn(x, sd, bnname, training))


def cost_reg_net(x, sd, prefix="cost_reg", training=False):
    """CostRegNet (stages with ndepth > 8).  models/module.py:469-505: stride 2 in D,H,W; skip adds
    AFTER the ReLU of the transposed block; prob = conv 3x3x3 8->1 without bias."""
    p = prefix + "."
    c0 = x
    c2 = _conv_block(_conv_block(c0, sd, p + "conv1", 2, training), sd, p + "conv2", 1, training)
    c4 = _conv_block(_conv_block(c2, sd, p + "conv3", 2, training), sd, p + "conv4", 1, training)
    y = _conv_block(_conv_block(c4, sd, p + "conv5", 2, training), sd, p + "conv6", 1, training)
    y = c4 + _deconv_block(y, sd, p + "conv7.conv.weight", p + "conv7.bn", 2, 1, training)
    y = c2 + _deconv_block(y, sd, p + "conv9.conv.weight", p + "conv9.bn", 2, 1, training)
    y = c0 + _deconv_block(y, sd, p + "conv11.conv.weight", p + "conv11.bn", 2, 1, training)
    return F.conv3d(y, sd[p + "prob.weight"].to(x.dtype), padding=1)


def cost_reg_net_3d(x, sd, prefix="cost_reg", training=False):
    """CostRegNet3D (ndepth <= 8).  models/module.py:550-594: stride (1,2,2); transposed blocks are
    nn.Sequential (keys conv7.0.weight / conv7.1.*); prob = conv 1x1x1 8->1 WITH bias."""
    p = prefix + "."
    st, op = (1, 2, 2), (0, 1, 1)
    c0 = x
    c2 = _conv_block(_conv_block(c0, sd, p + "conv1", st, training), sd, p + "conv2", 1, training)
    c4 = _conv_block(_conv_block(c2, sd, p + "conv3", st, training), sd, p + "conv4", 1, training)
    y = _conv_block(_conv_block(c4, sd, p + "conv5", st, training), sd, p + "conv6", 1, training)
    y = c4 + _deconv_block(y, sd, p + "conv7.0.weight", p + "conv7.1", st, op, training)
    y = c2 + _deconv_block(y, sd, p + "conv9.0.weight", p + "conv9.1", st, op, training)
    y = c0 + _deconv_block(y, sd, p + "conv11.0.weight", p + "conv11.1", st, op, training)
    return F.conv3d(y, sd[p + "prob.weight"].to(x.dtype), sd[p + "prob.bias"].to(x.dtype))


def cost_reg_net_2d(x, sd, prefix="cost_reg", training=False):
    """CostRegNet2D (models/module.py:508-547; not used by shipped configs): the strided and
    transposed layers use (1,3,3) kernels, the stride-1 layers 3x3x3."""
    p = prefix + "."
    st, op, k2 = (1, 2, 2), (0, 1, 1), (1, 3, 3)
    c0 = x
    c2 = _conv_block(_conv_block(c0, sd, p + "conv1", st, training, k2), sd, p + "conv2", 1, training)
    c4 = _conv_block(_conv_block(c2, sd, p + "conv3", st, training, k2), sd, p + "conv4", 1, training)
    y = _conv_block(_conv_block(c4, sd, p + "conv5", st, training, k2), sd, p + "conv6", 1, training)
    y = c4 + _deconv_block(y, sd, p + "conv7.0.weight", p + "conv7.1", st, op, training, k2)
    y = c2 + _deconv_block(y, sd, p + "conv9.0.weight", p + "conv9.1", st, op, training, k2)
    y = c0 + _deconv_block(y, sd, p + "conv11.0.weight", p + "conv11.1", st, op, training, k2)
    return F.conv3d(y, sd[p + "prob.weight"].to(x.dtype), sd[p + "prob.bias"].to(x.dtype))


# ----------------------------------------------------------------------------------------------
# A8. head                                   models/mvsformer_model.py:110-125, module.py:597-619
# ----------------------------------------------------------------------------------------------


def depth_regression(p, depth_values):
    """models/module.py:597-603: sum_d p[b,d,y,x] * depth_values[b,d(,y,x)]."""
    if depth_values.dim() <= 2:
        depth_values = depth_values.view(*depth_values.shape, 1, 1)
    return (p * depth_values).sum(dim=1)


def conf_regression(p, n=4):
    """models/module.py:606-619: windowed probability sum (window n, left pad n//2-1 for even n,
    n//2 for odd) gathered at floor(sum_d p*d)."""
    batch, nd, height, width = p.shape
    left = n // 2 if n % 2 == 1 else n // 2 - 1
    padded = F.pad(p, (0, 0, 0, 0, left, n // 2))
    win = sum(padded[:, k:k + nd] for k in range(n))
    idx = depth_regression(p, torch.arange(nd, dtype=torch.float32)).long().clamp(0, nd - 1)
    return torch.gather(win, 1, idx.unsqueeze(1)).squeeze(1)


def regression_head(prob_volume_pre, depth_values, tmp, training=False):
    """models/mvsformer_model.py:110-125 for depth_type in ('ce', 'was')."""
    prob = torch.softmax(prob_volume_pre, dim=1)
    if training:
        idx = prob.argmax(dim=1, keepdim=True)
        depth = torch.gather(depth_values, 1, idx).squeeze(1)
    else:
        depth = depth_regression(torch.softmax(prob_volume_pre * tmp, dim=1), depth_values)
    conf = prob.max(dim=1)[0]
    return prob, depth, conf


# ----------------------------------------------------------------------------------------------
# A9. hypothesis schedules                                         models/module.py:622-699
# ----------------------------------------------------------------------------------------------


def init_range(cur_depth, ndepths, height, width):
    """models/module.py:622-630."""
    dmin, dmax = cur_depth[:, 0], cur_depth[:, -1]
    itv = (dmax - dmin) / (ndepths - 1)
    k = torch.arange(ndepths, dtype=cur_depth.dtype).view(1, -1)
    d = dmin.unsqueeze(1) + k * itv.unsqueeze(1)
    return d.view(d.shape[0], ndepths, 1, 1).repeat(1, 1, height, width)


def init_inverse_range(cur_depth, ndepths, height, width):
    """models/module.py:633-639: linear in 1/d from 1/d[-1] (k=0, far) to 1/d[0] (k=D-1, near)."""
    inv_near = 1.0 / cur_depth[:, 0]
    inv_far = 1.0 / cur_depth[:, -1]
    frac = (torch.arange(ndepths, dtype=cur_depth.dtype) / (ndepths - 1)).view(1, -1, 1, 1)
    frac = frac.repeat(1, 1, height, width)
    inv = inv_far.view(-1, 1, 1, 1) + (inv_near - inv_far).view(-1, 1, 1, 1) * frac
    return 1.0 / inv


def _upsample_planes_align_corners(x, height, width):
    """Per-plane bilinear resize with align_corners=True; equals F.interpolate(trilinear,
    align_corners=True) when the depth count is unchanged (models/module.py:652)."""
    batch, nd, hin, win = x.shape
    dt = x.dtype

    def axis(n_out, n_in):
        if n_out == 1:
            src = torch.zeros(1, dtype=dt)
        else:
            src = torch.arange(n_out, dtype=dt) * ((n_in - 1) / (n_out - 1))
        i0 = src.floor().clamp(0, n_in - 1).long()
        i1 = (i0 + 1).clamp(max=n_in - 1)
        lam = src - i0.to(dt)
        return i0, i1, lam

    y0, y1, ly = axis(height, hin)
    x0, x1, lx = axis(width, win)
    ly = ly.view(1, 1, -1, 1)
    lx = lx.view(1, 1, 1, -1)
    top = x[:, :, y0][:, :, :, x0] * (1 - lx) + x[:, :, y0][:, :, :, x1] * lx
    bot = x[:, :, y1][:, :, :, x0] * (1 - lx) + x[:, :, y1][:, :, :, x1] * lx
    return top * (1 - ly) + bot * ly


def schedule_inverse_range(depth, depth_hypo, ndepths, split_itv, height, width):
    """models/module.py:642-653.  ``depth [B,h/2,w/2]``, ``depth_hypo [B,D',h/2,w/2]`` ->
    ``[B,ndepths,h,w]``: k=0 is 1/depth - split*itv (far), k=D-1 is 1/depth + split*itv (near)."""
    itv = 1.0 / depth_hypo[:, 2] - 1.0 / depth_hypo[:, 1]
    inv_hi = 1.0 / depth + split_itv * itv
    inv_lo = 1.0 / depth - split_itv * itv
    frac = (torch.arange(ndepths, dtype=depth.dtype) / (ndepths - 1)).view(1, -1, 1, 1)
    inv = inv_lo.unsqueeze(1) + (inv_hi - inv_lo).unsqueeze(1) * frac
    inv = _upsample_planes_align_corners(inv, height, width)
    return 1.0 / inv


def schedule_range(cur_depth, ndepth, depth_interval_pixel, height, width):
    """models/module.py:687-699 (linear-depth variant)."""
    half = ndepth / 2 * depth_interval_pixel.view(-1, 1, 1)
    dmin = (cur_depth - half).clamp_min(0.01)
    dmax = cur_depth + half
    itv = (dmax - dmin) / (ndepth - 1)
    k = torch.arange(ndepth, dtype=cur_depth.dtype).view(1, -1, 1, 1)
    d = dmin.unsqueeze(1) + k * itv.unsqueeze(1)
    return _upsample_planes_align_corners(d, height, width)


# ----------------------------------------------------------------------------------------------
# StageNet / cascade                                models/mvsformer_model.py:51-158, :410-449
# ----------------------------------------------------------------------------------------------


def stage_forward(features, proj_matrices, depth_values, sd, ndepth, tmp, groups=8, model_th=8,
                  training=False, want_parts=False):
    """One StageNet.forward (fusion_type='cnn', depth_type='ce').  ``sd`` holds this stage's
    state_dict entries with keys ``vis.*`` / ``cost_reg.*``."""
    volume, sim_sum, parts = build_cost_volume(features, proj_matrices, depth_values, sd, groups,
                                               training, want_parts)
    if ndepth <= model_th:
        pre = cost_reg_net_3d(volume, sd, "cost_reg", training).squeeze(1)
    else:
        pre = cost_reg_net(volume, sd, "cost_reg", training).squeeze(1)
    prob, depth, conf = regression_head(pre, depth_values, tmp, training)
    out = {"depth": depth, "prob_volume": prob, "photometric_confidence": conf,
           "depth_values": depth_values, "prob_volume_pre": pre}
    if not training:
        idx = sim_sum.argmax(dim=1, keepdim=True)                                  # :151-156
        out["sim_depth"] = torch.gather(depth_values, 1, idx).squeeze(1)
        out["sim_volume"] = sim_sum                                                # oracle extra
    if want_parts:
        out["volume_mean"] = volume
        out["parts"] = parts
    return out


def split_stage_state(full_sd, stage_idx):
    """``fusions.{i}.xxx`` -> ``xxx`` for one stage."""
    pre = "fusions.%d." % stage_idx
    return {k[len(pre):]: v for k, v in full_sd.items() if k.startswith(pre)}


def cascade_forward(features, proj_matrices, depth_values, stage_sds, ndepths=(32, 16, 8, 4),
                    ratios=(4.0, 2.67, 1.5, 1.0), tmp=(5.0, 5.0, 5.0, 1.0), inverse_depth=True,
                    groups=8, training=False, full_hw=None):
    """The cascade loop of models/mvsformer_model.py:410-449 over pre-extracted features.

    ``features``: {"stageK": [B,V,C,h,w]}, ``proj_matrices``: {"stageK": [B,V,2,4,4]},
    ``depth_values [B,192]``.  ``full_hw`` = image (H, W) used for the confidence accumulation
    (defaults to the last stage's resolution)."""
    outputs = {}
    nst = len(ndepths)
    last = None
    batch = depth_values.shape[0]
    if full_hw is None:
        full_hw = tuple(features["stage%d" % nst].shape[-2:])
    prob_maps = torch.zeros(batch, full_hw[0], full_hw[1], dtype=depth_values.dtype)
    depth_interval = depth_values[:, 1] - depth_values[:, 0]
    for s in range(nst):
        feats = features["stage%d" % (s + 1)]
        h, w = feats.shape[-2:]
        if s == 0:
            hyp = init_inverse_range(depth_values, ndepths[s], h, w) if inverse_depth \
                else init_range(depth_values, ndepths[s], h, w)
        elif inverse_depth:
            hyp = schedule_inverse_range(last["depth"], last["depth_values"], ndepths[s], ratios[s], h, w)
        else:
            hyp = schedule_range(last["depth"], ndepths[s], ratios[s] * depth_interval, h, w)
        t = tmp[s] if isinstance(tmp, (list, tuple)) else tmp
        last = stage_forward(feats, proj_matrices["stage%d" % (s + 1)], hyp, stage_sds[s], ndepths[s], t,
                             groups, training=training)
        conf = last["photometric_confidence"]
        if conf.shape[-2:] != prob_maps.shape[-2:]:                       # nearest upsample (:438-441)
            conf = F.interpolate(conf.unsqueeze(1), list(prob_maps.shape[-2:]), mode="nearest").squeeze(1)
            last["photometric_confidence"] = conf
        prob_maps = prob_maps + conf
        outputs["stage%d" % (s + 1)] = last
        outputs.update(last)
    outputs["refined_depth"] = last["depth"]
    outputs["photometric_confidence"] = prob_maps / nst
    return outputs
