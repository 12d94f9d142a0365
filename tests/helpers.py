"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STAGE_ARGS = {"base_ch": 8, "fusion_type": "cnn", "depth_type": "ce"}
CASCADE_ARGS = dict(STAGE_ARGS, ndepths=[32, 16, 8, 4], depth_interals_ratio=[4.0, 2.67, 1.5, 1.0], inverse_depth=True)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def rel_l1(a, b):
    a = torch.as_tensor(a).detach().double()
    b = torch.as_tensor(b).detach().double()
    return float((a - b).abs().mean() / b.abs().mean().clamp_min(1e-30))


def max_abs(a, b):
    return float((torch.as_tensor(a).double() - torch.as_tensor(b).double()).abs().max())


def checksum(t):
    return float(t.double().abs().sum())


def reference_state_dict_template(stage_idx, ndepth):
    """Names/shapes of one StageNet's state_dict, built WITHOUT the reference: from our own
    drop-in module (same keys by contract; test_state_dict_contract pins that against a
    recorded key list)."""
    from mvsformer_b200.mvsformer_model import StageNet

    return StageNet(dict(STAGE_ARGS), ndepth, stage_idx).state_dict()
