"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference on CPU.  TEST INFRASTRUCTURE.

Run in the build container (where ``/root/reference`` exists):

    python -m oracle.make_golden

Inputs and weights are NOT stored: they are regenerated from ``mvsformer_b200.synthetic`` with the
seeds recorded in each file (plus a checksum so a drifting RNG is detected, not silently
compared).  Outputs are what the reference functions returned:

* ``warp_*.npz``        models/warping.py:69-109,155-189
* ``schedules.npz``     models/module.py:597-699
* ``stage{1..4}.npz``   StageNet.forward (eval)  models/mvsformer_model.py:51-158
* ``stage2_train.npz``  StageNet.forward (train mode: batch-stat BN, argmax depth)
* ``stage{2,4}_train_grads.npz``  gradients of a cross-entropy loss on ``prob_volume_pre`` (models/losses.py:340-341)
                        through StageNet.forward in train mode, from torch autograd over the reference
* ``epipole.npz``        StageNet.forward with fusion_type 'epipole' / 'epipoleV2' (train gradients, eval outputs)
* ``heads.npz``          StageNet.forward (eval) with depth_type 'mixup_ce' / 're' (models/mvsformer_model.py:126-146)
* ``fusion.npz``         misc/fusion.py:69-118 + test.py:433-435 (prob_filter, get_reproj, vis_filter, ave_fusion, points)
* ``state_dict_keys.json``  names/shapes of the 302 ``fusions.*`` checkpoint entries
* ``cascade.npz``       the cascade loop models/mvsformer_model.py:410-449 driven over synthetic features
                        (the loop is re-stated here in 15 lines because the reference only has it
                        inline in TwinMVSNet.forward behind the ViT feature extractor)
"""
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from mvsformer_b200 import synthetic as S  # noqa: E402
from oracle import ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
STAGE_ARGS = {"base_ch": 8, "fusion_type": "cnn", "depth_type": "ce"}


def checksum(t):
    return float(t.double().abs().sum())


def np_(t):
    return t.detach().cpu().numpy()


def gen_warp(ns):
    feats = S.make_features(2, 3, 24, 40, seed=5, stages=(3,), feat_chs=(0, 0, 0, 8))["stage4"]     # [2,3,8,24,40]
    cams = S.make_cameras(2, 3, 24, 40)["stage4"]
    # exaggerate the baseline so a good share of samples leaves the image
    cams = cams.clone()
    cams[:, 2, 0, 0, 3] += 150.0
    O_compose = lambda p: _compose(p)
    ref_p = O_compose(cams[:, 0])
    dv_map = S.make_depth_range(2)[:, ::32][:, :6]                                                   # [2,6]
    dv_px = dv_map.view(2, 6, 1, 1) * (1.0 + 0.1 * torch.rand(2, 6, 24, 40, generator=S._gen(3)))
    out = {"seed": 5, "feat_checksum": checksum(feats)}
    for v in (1, 2):
        src_p = O_compose(cams[:, v])
        w1, m1 = ns.warping.homo_warping_3D_with_mask(feats[:, v], src_p, ref_p, dv_map)
        w2, m2 = ns.warping.homo_warping_3D_with_mask(feats[:, v], src_p, ref_p, dv_px)
        w3 = ns.warping.homo_warping_3D(feats[:, v], src_p, ref_p, dv_px)
        assert torch.equal(w2, w3)
        out["warped_bd_v%d" % v] = np_(w1)
        out["mask_bd_v%d" % v] = np_(m1)
        out["warped_px_v%d" % v] = np_(w2)
        out["mask_px_v%d" % v] = np_(m2)
    np.savez_compressed(os.path.join(OUT, "warp.npz"), **out)


def _compose(pair):
    p = pair[:, 0].clone()
    p[:, :3, :4] = torch.matmul(pair[:, 1, :3, :3], pair[:, 0, :3, :4])
    return p


def gen_schedules(ns):
    M = ns.module
    dv = S.make_depth_range(2)
    g = S._gen(11)
    out = {}
    out["init_inverse_range"] = np_(M.init_inverse_range(dv, 32, "cpu", torch.float32, 8, 12))
    out["init_range"] = np_(M.init_range(dv, 32, "cpu", torch.float32, 8, 12))
    hyp = M.init_inverse_range(dv, 32, "cpu", torch.float32, 8, 12)
    depth = 500.0 + 300.0 * torch.rand(2, 8, 12, generator=g)
    out["sched_depth_in"] = np_(depth)
    out["schedule_inverse_range"] = np_(M.schedule_inverse_range(depth, hyp, 16, 2.67, 16, 24))
    out["schedule_range"] = np_(M.schedule_range(depth, 16, 2.67 * (dv[:, 1] - dv[:, 0]), 16, 24))
    p = torch.softmax(3.0 * torch.randn(2, 32, 8, 12, generator=g), dim=1)
    out["prob_in"] = np_(p)
    out["depth_regression_map"] = np_(M.depth_regression(p, hyp))
    out["depth_regression_vec"] = np_(M.depth_regression(p, dv[:, :32]))
    for n in (2, 3, 4):
        out["conf_regression_n%d" % n] = np_(M.conf_regression(p, n=n))
    np.savez_compressed(os.path.join(OUT, "schedules.npz"), **out)


def stage_inputs(stage, height, width, batch=1, views=3, seed=1234):
    feats = S.make_features(batch, views, height, width, seed=seed, stages=(stage,))["stage%d" % (stage + 1)]
    cams = S.make_cameras(batch, views, height, width)["stage%d" % (stage + 1)]
    return feats, cams


def gen_stages(ns):
    R = ns.mvsformer_model
    height, width = 128, 192
    for s in range(4):
        net = R.StageNet(dict(STAGE_ARGS), S.NDEPTHS[s], s).eval()
        sd = S.fill_state_dict(net.state_dict(), seed=s)
        net.load_state_dict(sd)
        feats, cams = stage_inputs(s, height, width, batch=2 if s < 2 else 1)
        hyp = S.narrow_hypotheses(s, height, width, feats.shape[0])
        with torch.no_grad():
            out = net(feats, cams, hyp, tmp=list(S.EVAL_TMP))
        np.savez_compressed(
            os.path.join(OUT, "stage%d.npz" % (s + 1)),
            height=height, width=width, batch=feats.shape[0], views=3, weight_seed=s,
            feat_checksum=checksum(feats), hyp_checksum=checksum(hyp),
            **{k: np_(v) for k, v in out.items() if k != "depth_values"})
    # training mode (batch-stat BN, argmax depth) for one stage
    s = 1
    net = R.StageNet(dict(STAGE_ARGS), S.NDEPTHS[s], s).train()
    sd = S.fill_state_dict(net.state_dict(), seed=s)
    net.load_state_dict(sd)
    feats, cams = stage_inputs(s, height, width, batch=2)
    hyp = S.narrow_hypotheses(s, height, width, 2)
    with torch.no_grad():
        out = net(feats, cams, hyp, tmp=list(S.EVAL_TMP))
    np.savez_compressed(os.path.join(OUT, "stage2_train.npz"), height=height, width=width, batch=2, views=3,
                        weight_seed=s, feat_checksum=checksum(feats),
                        **{k: np_(v) for k, v in out.items() if k != "depth_values"})


TRAIN_GRAD_FULL = ("vis.0.conv.weight", "vis.0.bn.weight", "vis.2.bn.bias", "vis.3.weight", "vis.3.bias",
                   "cost_reg.conv1.conv.weight", "cost_reg.conv1.bn.weight", "cost_reg.conv2.bn.bias",
                   "cost_reg.prob.weight")


def train_target(stage, batch, h, w):
    """Synthetic ground-truth hypothesis index per pixel (the role of gt_index_volume, models/losses.py:338)."""
    return torch.randint(0, S.NDEPTHS[stage], (batch, h, w), generator=S._gen(700 + stage))


def gen_train_grads(ns):
    """Backward of the training branch: reference StageNet.train(), loss = CE(prob_volume_pre, target)."""
    R = ns.mvsformer_model
    height, width = 64, 96
    for s in (1, 3):
        net = R.StageNet(dict(STAGE_ARGS), S.NDEPTHS[s], s).train()
        sd = S.fill_state_dict(net.state_dict(), seed=50 + s)
        net.load_state_dict(sd)
        feats, cams = stage_inputs(s, height, width, batch=2, seed=77 + s)
        hyp = S.narrow_hypotheses(s, height, width, 2)
        feats = feats.clone().requires_grad_(True)
        out = net(feats, cams, hyp, tmp=list(S.EVAL_TMP))
        target = train_target(s, 2, feats.shape[-2], feats.shape[-1])
        loss = F.cross_entropy(out["prob_volume_pre"], target)
        loss.backward()
        grads = {"grad_features": np_(feats.grad), "loss": float(loss), "prob_volume_pre": np_(out["prob_volume_pre"])}
        names = []
        for name, p in net.named_parameters():
            names.append(name)
            grads["abs_sum/" + name] = float(p.grad.double().abs().sum())
            grads["sum/" + name] = float(p.grad.double().sum())
            if name in TRAIN_GRAD_FULL or name.endswith("conv11.0.weight") or name.endswith("conv11.conv.weight"):
                grads["grad/" + name] = np_(p.grad)
        for name, buf in net.named_buffers():
            if "running" in name and (name.startswith("vis.0") or name.startswith("cost_reg.conv1.")):
                grads["buf/" + name] = np_(buf)
        np.savez_compressed(os.path.join(OUT, "stage%d_train_grads.npz" % (s + 1)), height=height, width=width, batch=2,
                            views=3, weight_seed=50 + s, feat_seed=77 + s, feat_checksum=checksum(feats),
                            param_names=np.array(names), **grads)


def gen_heads(ns):
    """The non-default heads of StageNet.forward (models/mvsformer_model.py:126-146): depth_type 'mixup_ce' and 're',
    eval mode, for ndepth 16 (windowed confidence n = 3) and ndepth 4 (max-probability confidence)."""
    R = ns.mvsformer_model
    height, width = 64, 96
    out = {}
    for kind in ("mixup_ce", "re"):
        for s in (1, 3):
            net = R.StageNet(dict(STAGE_ARGS, depth_type=kind), S.NDEPTHS[s], s).eval()
            net.load_state_dict(S.fill_state_dict(net.state_dict(), seed=70 + s))
            feats, cams = stage_inputs(s, height, width, batch=1, seed=90 + s)
            hyp = S.narrow_hypotheses(s, height, width, 1)
            with torch.no_grad():
                res = net(feats, cams, hyp, tmp=list(S.EVAL_TMP))
            for key in ("depth", "photometric_confidence", "prob_volume"):
                out["%s_s%d_%s" % (kind, s + 1, key)] = np_(res[key])
    np.savez_compressed(os.path.join(OUT, "heads.npz"), height=height, width=width, **out)


def gen_epipole(ns):
    """fusion_type 'epipole' / 'epipoleV2' (models/mvsformer_model.py:92-104; no shipped config uses them): the reference
    StageNet in training (CE loss, gradients) and in eval."""
    R = ns.mvsformer_model
    height, width = 64, 96
    out = {}
    for kind in ("epipole", "epipoleV2"):
        s = 2                                                   # stage 3: C = 16, D = 8, half resolution
        args = dict(STAGE_ARGS, fusion_type=kind, attn_temp=2.0)
        feats, cams = stage_inputs(s, height, width, batch=1, seed=60)
        cams = cams.clone()
        cams[:, 2, 0, 0, 3] += 90.0                             # part of a view leaves the image: V2's mask matters
        hyp = S.narrow_hypotheses(s, height, width, 1)
        target = train_target(s, 1, feats.shape[-2], feats.shape[-1])
        net = R.StageNet(args, S.NDEPTHS[s], s).train()
        net.load_state_dict(S.fill_state_dict(net.state_dict(), seed=80))
        if kind == "epipoleV2":
            with torch.no_grad():
                net.attn_temp.fill_(1.7)
        f = feats.clone().requires_grad_(True)
        res = net(f, cams, hyp, tmp=list(S.EVAL_TMP))
        F.cross_entropy(res["prob_volume_pre"], target).backward()
        out[kind + "_train_pre"] = np_(res["prob_volume_pre"])
        out[kind + "_train_gfeat"] = np_(f.grad)
        out[kind + "_train_gconv1"] = np_(net.cost_reg.conv1.conv.weight.grad)
        if kind == "epipoleV2":
            out[kind + "_train_gtemp"] = np_(net.attn_temp.grad)
        net2 = R.StageNet(args, S.NDEPTHS[s], s).eval()
        net2.load_state_dict(S.fill_state_dict(net2.state_dict(), seed=80))
        with torch.no_grad():
            ev = net2(feats, cams, hyp, tmp=list(S.EVAL_TMP))
        for key in ("prob_volume_pre", "depth", "sim_depth"):
            out["%s_eval_%s" % (kind, key)] = np_(ev[key])
    np.savez_compressed(os.path.join(OUT, "epipole.npz"), height=height, width=width, **out)


def gen_fusion(ns):
    """Depth-map fusion: the reference's misc/fusion.py functions on a synthetic consistent scene.  Its
    get_pixel_grids calls ``.cuda()``; that call is patched to a no-op for the duration (nothing else is touched)."""
    import importlib
    sys.path.insert(0, ref_import.REFERENCE_ROOT)
    fusion = importlib.import_module("misc.fusion")
    case = S.make_fusion_case(4, 24, 32, seed=21)
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        prob_mask = fusion.prob_filter(case["ref_conf"], [0.1, 0.2, 0.3])
        reproj_xyd, in_range = fusion.get_reproj(case["ref_depth"], case["src_depths"], case["ref_cam"], case["src_cams"])
        masks, mask = fusion.vis_filter(case["ref_depth"], reproj_xyd, in_range, 1.0, 0.01, 2)
        ave = fusion.ave_fusion(case["ref_depth"], reproj_xyd, masks)
        idx_img = fusion.get_pixel_grids(24, 32).unsqueeze(0)
        points = fusion.idx_cam2world(fusion.idx_img2cam(idx_img, ave, case["ref_cam"]), case["ref_cam"])[..., :3, 0].permute(0, 3, 1, 2)
        # dynamic consistency checking: misc/fusion.py:116-168 + the vote / averaging of test.py:502-511 (restated verbatim)
        dyn_xyd = fusion.get_reproj_dynamic(case["ref_depth"], case["src_depths"], case["ref_cam"], case["src_cams"])
        dyn_masks, dyn_mask = fusion.vis_filter_dynamic(case["ref_depth"], dyn_xyd, dist_base=4, rel_diff_base=1300)
        dy_range = case["src_depths"].shape[1] + 1
        reproj_depth = dyn_xyd[:, :, -1].clone()
        reproj_depth[~dyn_mask.squeeze(2)] = 0
        geo_mask_sums = dyn_masks.sum(dim=1)
        geo_mask_sum = dyn_mask.sum(dim=1)
        dyn_ave = (torch.sum(reproj_depth, dim=1, keepdim=True) + case["ref_depth"]) / (geo_mask_sum + 1)
        geo_mask = geo_mask_sum >= dy_range
        for i in range(2, dy_range):
            geo_mask = torch.logical_or(geo_mask, geo_mask_sums[:, i - 2] >= i)
    finally:
        torch.Tensor.cuda = orig_cuda
    np.savez_compressed(os.path.join(OUT, "fusion.npz"), views=4, height=24, width=32, seed=21,
                        depth_checksum=checksum(case["src_depths"]), prob_mask=np_(prob_mask), reproj_xyd=np_(reproj_xyd),
                        in_range=np_(in_range), masks=np_(masks), mask=np_(mask), ave=np_(ave), points=np_(points),
                        dyn_xyd=np_(dyn_xyd), dyn_level_counts=np_(geo_mask_sums.float()), dyn_mask=np_(dyn_mask),
                        dyn_ave=np_(dyn_ave), dyn_geo_mask=np_(geo_mask))


def reference_cascade(ns, features, cams, depth_values, nets, tmp, ratios):
    """The loop of models/mvsformer_model.py:410-449 (TwinMVSNet.forward after feature extraction),
    calling the reference's own schedulers and StageNets."""
    M = ns.module
    outputs, last = {}, None
    full_h, full_w = features["stage4"].shape[-2:]
    prob_maps = torch.zeros(depth_values.shape[0], full_h, full_w)
    for s in range(4):
        f = features["stage%d" % (s + 1)]
        h, w = f.shape[-2:]
        if s == 0:
            hyp = M.init_inverse_range(depth_values, S.NDEPTHS[s], "cpu", torch.float32, h, w)
        else:
            hyp = M.schedule_inverse_range(last["depth"].detach(), last["depth_values"], S.NDEPTHS[s], ratios[s], h, w)
        last = nets[s](f, cams["stage%d" % (s + 1)], hyp, tmp=tmp)
        conf = last["photometric_confidence"]
        if conf.shape[-2:] != prob_maps.shape[-2:]:
            conf = F.interpolate(conf.unsqueeze(1), [full_h, full_w], mode="nearest").squeeze(1)
        prob_maps = prob_maps + conf
        outputs["stage%d" % (s + 1)] = last
    outputs["refined_depth"] = last["depth"]
    outputs["photometric_confidence"] = prob_maps / 4
    return outputs


def gen_cascade(ns):
    R = ns.mvsformer_model
    height, width, batch, views = 128, 192, 1, 4
    feats = S.make_features(batch, views, height, width, seed=77)
    cams = S.make_cameras(batch, views, height, width)
    dv = S.make_depth_range(batch)
    nets = []
    for s in range(4):
        net = R.StageNet(dict(STAGE_ARGS), S.NDEPTHS[s], s).eval()
        net.load_state_dict(S.fill_state_dict(net.state_dict(), seed=10 + s))
        nets.append(net)
    with torch.no_grad():
        out = reference_cascade(ns, feats, cams, dv, nets, list(S.EVAL_TMP), S.DEPTH_INTERVAL_RATIO)
    blob = {"height": height, "width": width, "batch": batch, "views": views, "feat_seed": 77, "weight_seed0": 10,
            "feat_checksum": checksum(feats["stage4"]),
            "refined_depth": np_(out["refined_depth"]), "photometric_confidence": np_(out["photometric_confidence"])}
    for s in range(4):
        st = out["stage%d" % (s + 1)]
        blob["stage%d_depth" % (s + 1)] = np_(st["depth"])
        blob["stage%d_prob_volume_pre" % (s + 1)] = np_(st["prob_volume_pre"])
        blob["stage%d_depth_values" % (s + 1)] = np_(st["depth_values"])
        blob["stage%d_sim_depth" % (s + 1)] = np_(st["sim_depth"])
    np.savez_compressed(os.path.join(OUT, "cascade.npz"), **blob)


def gen_state_dict_keys(ns):
    """Key names and shapes of ``fusions.*`` exactly as the reference's TwinMVSNet registers them
    (models/mvsformer_model.py:347): the checkpoint contract of the drop-in modules."""
    import json
    R = ns.mvsformer_model
    ml = torch.nn.ModuleList([R.StageNet(dict(STAGE_ARGS), S.NDEPTHS[i], i) for i in range(4)])
    keys = {"fusions." + k: list(v.shape) for k, v in ml.state_dict().items()}
    with open(os.path.join(OUT, "state_dict_keys.json"), "w") as f:
        json.dump(keys, f, indent=0, sort_keys=True)


def main():
    warnings.simplefilter("ignore")
    torch.set_num_threads(os.cpu_count())
    os.makedirs(OUT, exist_ok=True)
    ns = ref_import.load_reference()
    gen_warp(ns)
    gen_schedules(ns)
    gen_stages(ns)
    gen_cascade(ns)
    gen_train_grads(ns)
    gen_fusion(ns)
    gen_heads(ns)
    gen_epipole(ns)
    gen_state_dict_keys(ns)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
