"""Drop-in mirror of the hot-path half of the reference's ``models/mvsformer_model.py``.

* ``StageNet``     models/mvsformer_model.py:26-160 — same constructor, ``state_dict`` keys, forward
                   signature and output dict; the whole forward is ten-odd kernel launches in
                   libmvs_b200.so (fused warp + correlation, fused vis net, channels-last 3D CNN,
                   fused head) and no PyTorch arithmetic.
* ``CascadeMVS``   the cascade loop of ``TwinMVSNet.forward`` / ``DINOMVSNet.forward``
                   (models/mvsformer_model.py:410-449 / :273-308) over pre-extracted per-stage
                   features; ``fusions`` has the reference's name so ``fusions.*`` checkpoint
                   entries load unchanged.
* ``install_into(reference_module)``  swaps the reference's StageNet and schedulers for these,
                   which is how ``models/mvsformer_model.py`` uses the engine as a drop-in
                   (see INTEGRATION.md).
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import autograd, config, engine
from .module import (ConvBnReLU, CostRegNet, CostRegNet2D, CostRegNet3D, _FoldCache, _bn_scale_shift,
                     init_inverse_range, init_range, schedule_inverse_range, schedule_range)


class StageNet(nn.Module):
    def __init__(self, args, ndepth, stage_idx):
        super().__init__()
        self.args = args
        self.fusion_type = args.get("fusion_type", "cnn")
        self.ndepth = ndepth
        self.stage_idx = stage_idx
        in_channels = args["base_ch"]
        if self.fusion_type == "cnn":
            model_th = args.get("model_th", 8)
            self.vis = nn.Sequential(ConvBnReLU(1, 16), ConvBnReLU(16, 16), ConvBnReLU(16, 8), nn.Conv2d(8, 1, 1),
                                     nn.Sigmoid())
            if ndepth <= model_th:
                self.cost_reg = CostRegNet3D(in_channels, args["base_ch"])
            else:
                self.cost_reg = CostRegNet(in_channels, args["base_ch"])
        elif self.fusion_type == "epipole":
            # models/mvsformer_model.py:42-44 (in the reference's code, used by no shipped config)
            self.attn_temp = args.get("attn_temp", 2.0)
            self.cost_reg = CostRegNet2D(in_channels, args["base_ch"])
        elif self.fusion_type == "epipoleV2":
            self.attn_temp = nn.Parameter(torch.tensor(1.0, dtype=torch.float32), requires_grad=True)    # :45-47
            self.cost_reg = CostRegNet3D(in_channels, args["base_ch"])
        else:
            raise NotImplementedError
        self._vis_cache = _FoldCache()

    # -- visibility-net parameters, BN folded, packed for mvs_vis_weight ---------------------------
    def _vis_params_host(self):
        def build():
            chunks = []
            for i in range(3):
                blk = self.vis[i]
                scale, shift = _bn_scale_shift(blk.bn)
                w = blk.conv.weight.detach().float() * scale.view(-1, 1, 1, 1)      # [Co,Ci,3,3]
                chunks += [w.reshape(-1), shift.reshape(-1)]
            last = self.vis[3]
            chunks += [last.weight.detach().float().reshape(-1), last.bias.detach().float().reshape(-1)]
            return np.ascontiguousarray(torch.cat(chunks).cpu().numpy().astype(np.float32))

        tensors = []
        for i in range(3):
            blk = self.vis[i]
            tensors += [blk.conv.weight, blk.bn.weight, blk.bn.bias, blk.bn.running_mean, blk.bn.running_var]
        tensors += [self.vis[3].weight, self.vis[3].bias]
        return self._vis_cache.get(tensors, build)

    def _vis_params_fused(self):
        """Operands of the fused kernel: host params (layer 1, the two BN shifts, the 1x1 conv) and the packed TF32 weights
        of the 16->16 and 16->8 layers."""
        def build(_):
            host = self._vis_params_host()
            first = host[:16 * 9 + 16]
            last = host[-9:]
            packed, shifts = [], []
            for i in (1, 2):
                blk = self.vis[i]
                scale, shift = _bn_scale_shift(blk.bn)
                w = blk.conv.weight.detach().float() * scale.view(-1, 1, 1, 1)          # [Co,Ci,3,3]
                packed.append(engine.pack_vis_fused_weights(w, 48 if i == 1 else 32))
                shifts.append(shift.detach().float().cpu().numpy().astype(np.float32))
            params = np.ascontiguousarray(np.concatenate([first, shifts[0], shifts[1], last]).astype(np.float32))
            return params, packed[0], packed[1]
        self._vis_params_host()                                   # refreshes the cache key
        return self._vis_cache.get_derived("fused", build)

    def _vis_weight(self, entropy):
        """entropy [B,N,H,W] -> visibility weight [B,N,H,W] (models/mvsformer_model.py:91)."""
        b, n, h, w = entropy.shape
        maps = entropy.view(b * n, h, w)
        if config.conv_precision() == "tf32" and config.vis_fused():
            params, w2p, w3p = self._vis_params_fused()
            return engine.vis_fused(maps, params, w2p, w3p).view(b, n, h, w)
        return engine.vis_weight(maps, self._vis_params_host()).view(b, n, h, w)

    def build_cost_volume(self, features, proj_matrices, depth_values, features_cl=None, view_slots=None):
        """models/mvsformer_model.py:52-105 -> (volume channels-last [B,D,H,W,G], sim, entropy [B,N,H,W], vis_weight
        [B,N,H,W]); sim = sim_depth [B,H,W] from the channels-last kernels, the summed similarity volume [B,D,H,W] from the
        NCHW kernels, None in training.  ``features_cl`` [B,V,H,W,C]: the same features already re-laid
        out channels-last (CascadeMVS converts all stages in one launch); made here when absent.  With ``view_slots``
        (scan mode, B = 1) ``features`` may be None and ``features_cl`` is a per-scan pool [slots,H,W,C] whose maps
        ``view_slots`` = [reference, sources...] are this sample's views — nothing is gathered."""
        groups = self.args["base_ch"]
        round_tf32 = config.conv_precision() == "tf32"
        if view_slots is not None:
            if features_cl is None or features_cl.dim() != 4:
                raise RuntimeError("view_slots need a channels-last feature pool [slots,H,W,C]")
            assert len(view_slots) == proj_matrices.shape[1], "Different number of images and projection matrices"
            relproj = engine.relative_projections(proj_matrices)
            built = engine.cost_volume_cl_entropy(features_cl, relproj, depth_values, groups, not self.training, view_slots)
            if built is None:
                raise RuntimeError("no channels-last cost-volume kernel for %d channels x %d hypotheses"
                                   % (features_cl.shape[-1], depth_values.shape[1]))
            entropy, sim, corr = built
            weight = self._vis_weight(entropy)
            if corr is not None:
                return engine.corr_aggregate(corr, weight, round_tf32), sim, entropy, weight
            volume = engine.cost_volume_cl_aggregate(features_cl, relproj, depth_values, weight, groups, round_tf32, view_slots)
            return volume, sim, entropy, weight
        if features.dim() != 5:
            raise RuntimeError("features must be [B,V,C,H,W]")
        b, v = features.shape[:2]
        assert v == proj_matrices.shape[1], "Different number of images and projection matrices"
        if features.shape[2] % groups:
            raise RuntimeError("shape '[%d, %d, -1, ...]' is invalid for %d feature channels"
                               % (b, groups, features.shape[2]))
        relproj = engine.relative_projections(proj_matrices)
        if config.cv_layout() == "cl" and engine.cl_supported(features.shape[2], depth_values.shape[1], groups):
            if features_cl is None:
                features_cl = engine.features_to_cl([features])[0]
            built = engine.cost_volume_cl_entropy(features_cl, relproj, depth_values, groups, want_sim=not self.training)
            if built is not None:
                entropy, sim, corr = built
                weight = self._vis_weight(entropy)
                if corr is not None:                       # one sampling pass: the stored correlation is streamed back
                    return engine.corr_aggregate(corr, weight, round_tf32), sim, entropy, weight
                volume = engine.cost_volume_cl_aggregate(features_cl, relproj, depth_values, weight, groups, round_tf32)
                return volume, sim, entropy, weight
        entropy, sim = engine.cost_volume_entropy(features, relproj, depth_values, groups, want_sim=not self.training)
        weight = self._vis_weight(entropy)
        volume = engine.cost_volume_aggregate(features, relproj, depth_values, weight, groups, round_tf32=round_tf32)
        return volume, sim, entropy, weight

    def _vis_weight_train(self, entropy):
        """Training form of ``self.vis`` (batch-statistics BatchNorm, autograd): the reference calls the
        net once per source view (mvsformer_model.py:91), so statistics are per view.  entropy [B,H,W]."""
        x = entropy.contiguous().unsqueeze(1).unsqueeze(-1)                  # depth-1 volume [B,1,H,W,1]
        for i in range(3):
            x = self.vis[i].forward_cl(x)
        return autograd.thin_conv_module(x, self.vis[3], act=2).squeeze(-1).squeeze(1)      # sigmoid -> [B,H,W]

    def _forward_train(self, features, proj_matrices, depth_values, tmp):
        """models/mvsformer_model.py:51-158 with ``self.training``: differentiable w.r.t. the features, the
        visibility net and the regulariser; argmax depth; no similarity branch."""
        if features.dim() != 5:
            raise RuntimeError("features must be [B,V,C,H,W]")
        assert features.shape[1] == proj_matrices.shape[1], "Different number of images and projection matrices"
        groups = self.args["base_ch"]
        if features.shape[2] % groups:
            raise RuntimeError("shape '[%d, %d, -1, ...]' is invalid for %d feature channels"
                               % (features.shape[0], groups, features.shape[2]))
        relproj = engine.relative_projections(proj_matrices)
        corr = autograd.group_correlation(features, relproj, depth_values, groups)          # [B,N,D,H,W,G]
        entropy = autograd.corr_entropy(corr.detach())                                      # [B,N,H,W], :88 detach
        weight = torch.stack([self._vis_weight_train(entropy[:, v]) for v in range(entropy.shape[1])], dim=1)
        volume = autograd.aggregate(corr, weight)
        prob_volume_pre = self.cost_reg.forward_cl(volume)
        prob_volume, depth, conf = self._head(prob_volume_pre, depth_values, tmp)
        return {"depth": depth, "prob_volume": prob_volume, "photometric_confidence": conf,
                "depth_values": depth_values, "prob_volume_pre": prob_volume_pre}

    def _forward_epipole(self, features, proj_matrices, depth_values, tmp):
        """fusion_type 'epipole' / 'epipoleV2' (models/mvsformer_model.py:92-104): per-hypothesis softmax view weights.
        Used by no shipped config, so it runs through the materialised correlation of the training path in eval too."""
        if features.dim() != 5:
            raise RuntimeError("features must be [B,V,C,H,W]")
        assert features.shape[1] == proj_matrices.shape[1], "Different number of images and projection matrices"
        groups, chans = self.args["base_ch"], features.shape[2]
        if chans % groups:
            raise RuntimeError("shape '[%d, %d, -1, ...]' is invalid for %d feature channels" % (features.shape[0], groups, chans))
        relproj = engine.relative_projections(proj_matrices)
        corr = autograd.group_correlation(features, relproj, depth_values, groups)
        if self.fusion_type == "epipoleV2":
            volume = autograd.epipole_aggregate(corr, self.attn_temp, autograd.proj_mask(relproj, depth_values),
                                                math.sqrt(groups), clamp=(0.1, 10.0))
        else:
            volume = autograd.epipole_aggregate(corr, self.attn_temp, None, math.sqrt(chans))
        if not self.training and config.conv_precision() == "tf32":
            volume = engine.round_tf32(volume)                     # operand contract of the TF32 tensor-core convolutions
        prob_volume_pre = self.cost_reg.forward_cl(volume)
        prob_volume, depth, conf = self._head(prob_volume_pre, depth_values, tmp)
        outputs = {"depth": depth, "prob_volume": prob_volume, "photometric_confidence": conf,
                   "depth_values": depth_values, "prob_volume_pre": prob_volume_pre}
        if not self.training:
            _, sim = engine.cost_volume_entropy(features, relproj, depth_values, groups, want_sim=True)
            outputs["sim_depth"] = engine.argmax_gather(sim, depth_values)
        return outputs

    def _head(self, prob_volume_pre, depth_values, tmp):
        """models/mvsformer_model.py:110-146 -> (prob_volume, depth, photometric_confidence) for every depth_type:
        'ce' / 'was' (argmax depth in training, temperature regression in eval, max-probability confidence),
        'mixup_ce' (best adjacent pair, :126-136), anything else = regression ('re': expectation depth, windowed
        confidence by ndepth, :137-146)."""
        kind = self.args["depth_type"]
        if kind in ("ce", "was"):
            if self.training:
                return autograd.train_head(prob_volume_pre, depth_values, tmp)
            return engine.regression_head(prob_volume_pre, depth_values, tmp, False)
        if self.training:                                               # softmax stays differentiable
            prob_volume, _, max_prob = autograd.train_head(prob_volume_pre, depth_values, 1.0)
        else:
            prob_volume, _, max_prob = engine.regression_head(prob_volume_pre, depth_values, 1.0, False)
        if kind == "mixup_ce":
            depth, conf = autograd.mixup_head(prob_volume, depth_values)
            return prob_volume, depth, conf
        depth = autograd.depth_regression(prob_volume, depth_values) if self.training \
            else engine.depth_regression(prob_volume, depth_values)
        window = 4 if self.ndepth >= 32 else (3 if self.ndepth == 16 else (2 if self.ndepth == 8 else 0))
        conf = engine.conf_regression(prob_volume.detach(), window) if window else max_prob
        return prob_volume, depth, conf

    def forward(self, features, proj_matrices, depth_values, tmp=2.0, features_cl=None, view_slots=None):
        """features [B,V,C,H,W], proj_matrices [B,V,2,4,4], depth_values [B,D,H,W] (the reference's signature);
        ``features_cl`` / ``view_slots``: optional channels-last copy of ``features`` / feature pool (see build_cost_volume)."""
        depth_values = depth_values.float().contiguous()
        if self.fusion_type != "cnn":
            if type(tmp) == list or type(tmp) == tuple:
                tmp = tmp[self.stage_idx]
            return self._forward_epipole(features, proj_matrices, depth_values, tmp)
        if self.training:
            if type(tmp) == list or type(tmp) == tuple:
                tmp = tmp[self.stage_idx]
            return self._forward_train(features, proj_matrices, depth_values, tmp)
        volume, sim, _, _ = self.build_cost_volume(features, proj_matrices, depth_values, features_cl, view_slots)
        prob_volume_pre = self.cost_reg.forward_cl(volume)
        if type(tmp) == list or type(tmp) == tuple:
            tmp = tmp[self.stage_idx]
        prob_volume, depth, conf = self._head(prob_volume_pre, depth_values, tmp)
        outputs = {"depth": depth, "prob_volume": prob_volume, "photometric_confidence": conf,
                   "depth_values": depth_values, "prob_volume_pre": prob_volume_pre}
        if not self.training:
            # the channels-last kernels take the argmax themselves ([B,H,W]); the NCHW kernels return the volume
            outputs["sim_depth"] = sim if sim.dim() == 3 else engine.argmax_gather(sim, depth_values)
        return outputs


class CascadeMVS(nn.Module):
    """Coarse-to-fine cascade over pre-extracted features (the body of ``TwinMVSNet.forward`` after
    feature extraction, models/mvsformer_model.py:410-449)."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.ndepths = args["ndepths"]
        self.depth_interals_ratio = args["depth_interals_ratio"]
        self.inverse_depth = args.get("inverse_depth", False)
        self.fusions = nn.ModuleList([StageNet(args, self.ndepths[i], i) for i in range(len(self.ndepths))])
        self._side = {}                     # compute stream -> its side stream (calls on different streams do not share one)

    def _features_cl(self, features):
        """-> (channels-last copies of all stages' features, event the compute stream must wait for before stage 2) — eval,
        'cnn' fusion, shapes the channels-last kernels cover; (None, None) otherwise: every StageNet then decides for
        itself."""
        nst = len(self.ndepths)
        if self.training or nst > 4 or config.cv_layout() != "cl":
            return None, None
        feats = [features["stage%d" % (s + 1)] for s in range(nst)]
        groups = self.args["base_ch"]
        for s, f in enumerate(feats):
            if self.fusions[s].fusion_type != "cnn" or f.dim() != 5 or not f.is_cuda \
                    or not engine.cl_supported(f.shape[2], self.ndepths[s], groups):
                return None, None
        # stage 1 on the compute stream; stages 2-4 (HBM-bound copies) on a side stream, under the latency-bound kernels
        # of stage 1 — each stage waits for its own features only
        main = torch.cuda.current_stream(feats[0].device)
        side = self._side.get(main.cuda_stream)
        if side is None:
            side = self._side[main.cuda_stream] = torch.cuda.Stream(device=feats[0].device)
        out = engine.features_to_cl(feats[:1])
        done = None
        if nst > 1:
            # destinations are allocated on the compute stream (its allocator pool; no cross-stream frees), only the copies
            # run on the side stream; the compute stream waits for them before stage 2, i.e. before anything is released
            rest = [torch.empty(f.shape[:2] + (f.shape[3], f.shape[4], f.shape[2]), device=f.device, dtype=torch.float32)
                    for f in feats[1:]]
            side.wait_stream(main)
            with torch.cuda.stream(side):
                engine.features_to_cl(feats[1:], outs=rest)
                done = torch.cuda.Event()
                done.record(side)
            out += rest
        return out, done

    def forward(self, features, proj_matrices, depth_values, tmp=2.0, full_hw=None, pools_cl=None, view_slots=None):
        """features {"stageK": [B,V,C,h,w]}, proj_matrices {"stageK": [B,V,2,4,4]}, depth_values [B,ND].
        Scan mode (eval, B = 1): ``features=None``, ``pools_cl`` {"stageK": channels-last pool [slots,h,w,C]} and
        ``view_slots`` = the pool maps of [reference view, source views...] (StreamedCascade.run_scan)."""
        nst = len(self.ndepths)
        if pools_cl is not None:
            if self.training or view_slots is None:
                raise RuntimeError("feature pools are an eval-mode input and need view_slots")
            last_pool = pools_cl["stage%d" % nst]
            b = 1
            if full_hw is None:
                full_hw = tuple(last_pool.shape[1:3])
            device = last_pool.device
        else:
            last_feat = features["stage%d" % nst]
            b = last_feat.shape[0]
            if full_hw is None:
                full_hw = tuple(last_feat.shape[-2:])
            device = last_feat.device
        outputs = {}
        outputs_stage = None
        use_conf = self.args["depth_type"] in ("ce", "mixup_ce")
        prob_maps = torch.zeros(b, full_hw[0], full_hw[1], dtype=torch.float32, device=device) if use_conf else None
        depth_interval = depth_values[:, 1] - depth_values[:, 0]
        feats_cl, cl_ready = (None, None) if pools_cl is not None else self._features_cl(features)
        for s in range(nst):
            feats = None if pools_cl is not None else features["stage%d" % (s + 1)]
            h, w = pools_cl["stage%d" % (s + 1)].shape[1:3] if pools_cl is not None else feats.shape[-2:]
            if s == 0:
                rng = init_inverse_range if self.inverse_depth else init_range
                depth_samples = rng(depth_values, self.ndepths[s], device, torch.float32, h, w)
            elif self.inverse_depth:
                depth_samples = schedule_inverse_range(outputs_stage["depth"].detach(), outputs_stage["depth_values"],
                                                       self.ndepths[s], self.depth_interals_ratio[s], h, w)
            else:
                depth_samples = schedule_range(outputs_stage["depth"].detach(), self.ndepths[s],
                                               self.depth_interals_ratio[s] * depth_interval, h, w)
            if pools_cl is not None:
                outputs_stage = self.fusions[s](None, proj_matrices["stage%d" % (s + 1)], depth_samples, tmp=tmp,
                                                features_cl=pools_cl["stage%d" % (s + 1)], view_slots=view_slots)
            elif feats_cl is not None:
                if s == 1 and cl_ready is not None:
                    torch.cuda.current_stream(device).wait_event(cl_ready)            # stages 2-4 were re-laid out on the side stream
                outputs_stage = self.fusions[s](feats, proj_matrices["stage%d" % (s + 1)], depth_samples, tmp=tmp,
                                                features_cl=feats_cl[s])
            else:
                outputs_stage = self.fusions[s](feats, proj_matrices["stage%d" % (s + 1)], depth_samples, tmp=tmp)
            outputs["stage%d" % (s + 1)] = outputs_stage
            if use_conf:
                conf = outputs_stage["photometric_confidence"]
                if tuple(conf.shape[-2:]) != tuple(full_hw):
                    # the reference replaces the stage's confidence by its nearest upsampling (:439-441)
                    outputs_stage["photometric_confidence"] = engine.confidence_upsample_accumulate(conf, prob_maps, 1.0)
                else:
                    engine.confidence_accumulate(conf, prob_maps, 1.0)
            outputs.update(outputs_stage)
        outputs["refined_depth"] = outputs_stage["depth"]
        if use_conf:
            outputs["photometric_confidence"] = prob_maps / nst
        return outputs


def install_into(reference_model_module, reference_module_module=None):
    """Point the reference's ``models.mvsformer_model`` (already imported) at this engine: its
    ``TwinMVSNet`` / ``DINOMVSNet`` then build our StageNet and call our schedulers, and their
    checkpoints load unchanged.  See INTEGRATION.md."""
    m = reference_model_module
    m.StageNet = StageNet
    m.init_inverse_range = init_inverse_range
    m.init_range = init_range
    m.schedule_inverse_range = schedule_inverse_range
    m.schedule_range = schedule_range
    m.CostRegNet, m.CostRegNet3D, m.CostRegNet2D = CostRegNet, CostRegNet3D, CostRegNet2D
    if reference_module_module is not None:
        for name in ("CostRegNet", "CostRegNet3D", "CostRegNet2D", "init_inverse_range", "init_range",
                     "schedule_inverse_range", "schedule_range", "depth_regression", "conf_regression"):
            setattr(reference_module_module, name, globals().get(name) or getattr(__import__(
                "mvsformer_b200.module", fromlist=[name]), name))
    return m
