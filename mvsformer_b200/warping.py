"""Drop-in mirror of the reference's ``models/warping.py`` warping functions.

``homo_warping_3D_with_mask`` (models/warping.py:69-109), ``homo_warping_3D`` (:155-189) and
``diff_homo_warping_3D_with_mask`` (:112-152) keep their signatures and results, and are
differentiable w.r.t. ``src_fea`` (mvs_homo_warp_bwd).  StageNet does
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
    """The reference variant that lets gradients flow into the sampling grid (:112-152).  Forward
    values are identical and src_fea is differentiable; the backward through cameras / depth (used by no
    model in the reference) is not built, so such inputs that require grad are rejected, not silently detached."""
    for t in (src_proj, ref_proj, depth_values):
        if t.requires_grad:
            raise NotImplementedError("diff_homo_warping_3D_with_mask: gradients w.r.t. cameras / depth are not built")
    return homo_warping_3D_with_mask(src_fea, src_proj, ref_proj, depth_values)
