"""Drive the UNMODIFIED reference's hot path (its own StageNet modules and schedulers, loaded by
``oracle/ref_import.py`` from /root/reference or the staged ``oracle/_ref``) over pre-extracted features, on the
CPU or on a CUDA device.  TEST / BENCH INFRASTRUCTURE — never imported by the product package.

The only restated code is the ~15-line cascade loop of ``TwinMVSNet.forward`` (models/mvsformer_model.py:417-447),
because in the reference it sits behind the feature extraction; every operator it calls is the reference's own.
"""
import time
import warnings

import torch
import torch.nn.functional as F

from mvsformer_b200 import synthetic as S
from oracle import ref_import

STAGE_ARGS = {"base_ch": 8, "fusion_type": "cnn", "depth_type": "ce"}


def available():
    return ref_import.reference_available()


def build_stage_nets(state_dicts, device="cpu"):
    """The reference's StageNet x4 (eval) with the given per-stage state dicts (strict load)."""
    ns = ref_import.load_reference()
    nets = []
    for s in range(4):
        net = ns.mvsformer_model.StageNet(dict(STAGE_ARGS), S.NDEPTHS[s], s).eval()
        net.load_state_dict(state_dicts[s], strict=True)
        nets.append(net.to(device))
    return nets


def cascade(nets, features, cams, depth_values, tmp=None, ratios=None):
    """models/mvsformer_model.py:417-447 over given features (dict stageK -> [B,V,C,h,w]) on their device."""
    M = ref_import.load_reference().module
    tmp = list(S.EVAL_TMP) if tmp is None else tmp
    ratios = S.DEPTH_INTERVAL_RATIO if ratios is None else ratios
    dev = depth_values.device
    outputs, last = {}, None
    full_h, full_w = features["stage4"].shape[-2:]
    prob_maps = torch.zeros(depth_values.shape[0], full_h, full_w, dtype=torch.float32, device=dev)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for s in range(4):
            f = features["stage%d" % (s + 1)]
            h, w = f.shape[-2:]
            if s == 0:
                hyp = M.init_inverse_range(depth_values, S.NDEPTHS[s], dev, torch.float32, h, w)
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


def time_cpu(nets, features, cams, depth_values, steps, warmup):
    """seconds per cascade on the host (time.perf_counter, like test.py:233-249 without the cuda syncs)."""
    with torch.no_grad():
        for _ in range(warmup):
            cascade(nets, features, cams, depth_values)
        t0 = time.perf_counter()
        for _ in range(steps):
            cascade(nets, features, cams, depth_values)
        return (time.perf_counter() - t0) / steps


def time_cuda(nets, features, cams, depth_values, steps, warmup):
    """CUDA-event ms per cascade and per part (reference eager + cuDNN on this GPU): whole cascade, the four
    ``cost_reg`` regularisers alone (fed volumes of the right shape), and the rest (cost-volume build + head)."""
    def timed(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    with torch.no_grad():
        for _ in range(warmup):
            cascade(nets, features, cams, depth_values)
        whole = timed(lambda: cascade(nets, features, cams, depth_values), steps)
        vols = []
        for s in range(4):
            f = features["stage%d" % (s + 1)]
            vols.append(torch.randn(f.shape[0], 8, S.NDEPTHS[s], f.shape[-2], f.shape[-1], device=f.device))

        def cnn():
            for s in range(4):
                nets[s].cost_reg(vols[s])

        for _ in range(warmup):
            cnn()
        t_cnn = timed(cnn, steps)
    return {"ms_per_step": whole, "cnn_ms": t_cnn, "cost_volume_and_head_ms": whole - t_cnn}
