"""Drop-in mirror of the reference's ``models/warping.py`` warping functions.

``homo_warping_3D_with_mask`` (models/warping.py:69-109), ``homo_warping_3D`` (:155-189) and
``diff_homo_warping_3D_with_mask`` (:112-152) keep their signatures and results, and are
differentiable w.r.t. ``src_fea`` (mvs_homo_warp_bwd); the ``diff_`` variant also w.r.t. depth and cameras.  StageNet does
NOT call them (its cost-volume kernels sample on the fly); they exist for API completeness.
"""
from . import autograd, engine


def homo_warping_3D_with_mask(src_fea, src_proj, ref_proj, depth_values):
    """src_fea [B,C,H,W], src_proj/ref_proj [B,4,4], depth_values [B,D] or [B,D,H,W]
    -> (warped [B,C,D,H,W], proj_mask [B,D,H,W] bool)."""
    relproj = engine.relative_projection_pair(src_proj, ref_proj)
    return autograd.homo_warp(src_fea, relproj, depth_values, True)


def homo_warping_3D(src_fea, src_proj, ref_proj, depth_values):
    relproj = engine.relative_projection_pair(src_proj, ref_proj)
    return autograd.homo_warp(src_fea, relproj, depth_values, False)[0]


def diff_homo_warping_3D_with_mask(src_fea, src_proj, ref_proj, depth_values):
    """The reference variant whose sampling grid stays in the autograd graph (:112-152): same forward values, and
    src_fea, depth_values, src_proj and ref_proj all receive gradients (mvs_homo_warp_bwd, mvs_homo_warp_bwd_grid; the
    4x4 projection algebra and its gradient are torch's).  No model of the reference calls it."""
    return autograd.diff_homo_warp(src_fea, src_proj, ref_proj, depth_values)
