"""Drop-in mirror of the depth-map fusion functions of the reference's ``misc/fusion.py``,
as ``test.py:404-435`` (``filter_depth``) drives them:

* ``prob_filter(ref_prob, prob_thresh)``                                     misc/fusion.py:69-76
* ``get_reproj(ref_depth, srcs_depth, ref_cam, srcs_cam)``                   misc/fusion.py:79-98
* ``vis_filter(ref_depth, reproj_xyd, in_range, img_dist_thresh, depth_thresh, vthresh)``   :101-109
* ``ave_fusion(ref_depth, reproj_xyd, masks)``                               :112-114
* ``world_points(depth, cam)`` = ``idx_cam2world(idx_img2cam(get_pixel_grids(h, w), depth, cam), cam)``  test.py:433-435
* ``get_reproj_dynamic`` / ``vis_filter_dynamic``                           misc/fusion.py:116-168 (``dynamic_filter_depth``, test.py:475-514)
* ``filter_view(...)`` / ``dynamic_filter_view(...)`` — the whole per-reference-view chain in three launches

Same tensor shapes as the reference (``n1hw`` depths, ``nv1hw`` source depths, ``n244`` cameras).  The per-pixel
work runs in libmvs_b200.so (``csrc/fusion.cu``): the intermediate per-source-pixel image of ``get_reproj`` is
evaluated on the fly at the four bilinear taps and never stored.  Camera matrices are inverted once per view pair
in fp64 on the host side of this module (16-element algebra; plumbing).  The PLY writer stays the reference's.
"""
import ctypes

import numpy as np
import torch

from . import _lib


def bin_op_reduce(lst, func):
    """misc/fusion.py:16-20."""
    result = lst[0]
    for i in range(1, len(lst)):
        result = func(result, lst[i])
    return result


def _call(name, *args):
    _lib.check(getattr(_lib.load(), name)(*args, _lib.stream()), name)


def _cam_block(cam):
    """cam [..., 2, 4, 4] (extrinsic, intrinsic) -> [..., 50] = Kinv (9) Einv (16) E (16) K (9), inverted in fp64."""
    cam = cam.double()
    ext, kmat = cam[..., 0, :, :], cam[..., 1, :3, :3]
    lead = cam.shape[:-3]
    parts = [torch.linalg.inv(kmat).reshape(*lead, 9), torch.linalg.inv(ext).reshape(*lead, 16), ext.reshape(*lead, 16),
             kmat.reshape(*lead, 9)]
    return torch.cat(parts, dim=-1)


def _pair_mats(ref_cam, srcs_cam):
    """ref_cam [n,2,4,4], srcs_cam [n,v,2,4,4] -> [n,v,100] float32 on the cameras' device."""
    n, v = srcs_cam.shape[:2]
    ref = _cam_block(ref_cam).unsqueeze(1).expand(n, v, 50)
    return torch.cat([ref, _cam_block(srcs_cam)], dim=-1).float().contiguous()


def prob_filter(ref_prob, prob_thresh, greater=True):
    """ref_prob [n,c,1,h,w] (-> bool [n,1,1,h,w], as the reference's indexing gives) or [n,c,h,w] (-> [n,1,h,w]);
    mask = AND_i ref_prob[:, i] > prob_thresh[i]."""
    p = ref_prob.float()
    five = p.dim() == 5
    if five:
        p = p.squeeze(2)
    p = p.contiguous()
    _lib.require_cuda(p)
    n, c, h, w = p.shape
    th = np.ascontiguousarray(np.asarray(list(prob_thresh), dtype=np.float32))
    mask = torch.empty(n, h, w, device=p.device, dtype=torch.float32)
    _call("mvs_fusion_prob_filter", _lib.ptr(p), th.ctypes.data_as(ctypes.c_void_p), int(th.size), _lib.ptr(mask), n, c, h, w)
    mask = mask.unsqueeze(1) > 0.5
    return mask.unsqueeze(2) if five else mask


def get_reproj(ref_depth, srcs_depth, ref_cam, srcs_cam):
    """ref_depth [n,1,h,w], srcs_depth [n,v,1,h,w], ref_cam [n,2,4,4], srcs_cam [n,v,2,4,4]
    -> (reproj_xyd [n,v,3,h,w], in_range [n,v,1,h,w])."""
    n, v, _, h, w = srcs_depth.shape
    rd = ref_depth.float().reshape(n, h, w).contiguous()
    sd = srcs_depth.float().reshape(n, v, h, w).contiguous()
    mats = _pair_mats(ref_cam, srcs_cam).to(rd.device)
    _lib.require_cuda(rd, sd, mats)
    xyd = torch.empty(n, v, 3, h, w, device=rd.device, dtype=torch.float32)
    inr = torch.empty(n, v, h, w, device=rd.device, dtype=torch.float32)
    _call("mvs_fusion_reproject", _lib.ptr(rd), _lib.ptr(sd), _lib.ptr(mats), _lib.ptr(xyd), _lib.ptr(inr), n, v, h, w)
    return xyd, inr.unsqueeze(2)


def _filter(ref_depth, reproj_xyd, in_range, img_dist_thresh, depth_thresh, vthresh):
    n, v, _, h, w = reproj_xyd.shape
    rd = ref_depth.float().reshape(n, h, w).contiguous()
    xyd = reproj_xyd.float().contiguous()
    inr = in_range.float().reshape(n, v, h, w).contiguous()
    _lib.require_cuda(rd, xyd, inr)
    masks = torch.empty(n, v, h, w, device=rd.device, dtype=torch.float32)
    mask = torch.empty(n, h, w, device=rd.device, dtype=torch.float32)
    ave = torch.empty(n, h, w, device=rd.device, dtype=torch.float32)
    _call("mvs_fusion_filter", _lib.ptr(rd), _lib.ptr(xyd), _lib.ptr(inr), float(img_dist_thresh), float(depth_thresh),
          float(vthresh), _lib.ptr(masks), _lib.ptr(mask), _lib.ptr(ave), n, v, h, w)
    return masks.unsqueeze(2), mask.unsqueeze(1) > 0.5, ave.unsqueeze(1)


def vis_filter(ref_depth, reproj_xyd, in_range, img_dist_thresh, depth_thresh, vthresh):
    """-> (masks [n,v,1,h,w] float 1/0, mask [n,1,h,w] bool)."""
    masks, mask, _ = _filter(ref_depth, reproj_xyd, in_range, img_dist_thresh, depth_thresh, vthresh)
    return masks, mask


def ave_fusion(ref_depth, reproj_xyd, masks):
    """(sum_v reproj_depth * masks + ref_depth) / (sum_v masks + 1)  ->  [n,1,h,w].  ``masks`` are vis_filter's; the
    kernel recomputes nothing: it is called with thresholds that reproduce the given masks exactly (masks act as the
    in-range input, the distance / depth tests are disabled)."""
    n, v, _, h, w = reproj_xyd.shape
    _, _, ave = _filter(ref_depth, reproj_xyd, masks.reshape(n, v, 1, h, w), float("inf"), float("inf"), 0.0)
    return ave


def get_reproj_dynamic(ref_depth, srcs_depth, ref_cam, srcs_cam):
    """misc/fusion.py:116-152 -> reproj_xyd [n,v,3,h,w]."""
    n, v, _, h, w = srcs_depth.shape
    rd = ref_depth.float().reshape(n, h, w).contiguous()
    sd = srcs_depth.float().reshape(n, v, h, w).contiguous()
    mats = _pair_mats(ref_cam, srcs_cam).to(rd.device)
    _lib.require_cuda(rd, sd, mats)
    xyd = torch.empty(n, v, 3, h, w, device=rd.device, dtype=torch.float32)
    _call("mvs_fusion_reproject_dynamic", _lib.ptr(rd), _lib.ptr(sd), _lib.ptr(mats), _lib.ptr(xyd), n, v, h, w)
    return xyd


def _filter_dynamic(ref_depth, reproj_xyd, dist_base, rel_diff_base, want_levels):
    n, v, _, h, w = reproj_xyd.shape
    rd = ref_depth.float().reshape(n, h, w).contiguous()
    xyd = reproj_xyd.float().contiguous()
    _lib.require_cuda(rd, xyd)
    vis = torch.empty(n, v, h, w, device=rd.device, dtype=torch.float32)
    geo = torch.empty(n, h, w, device=rd.device, dtype=torch.float32)
    ave = torch.empty(n, h, w, device=rd.device, dtype=torch.float32)
    levels = torch.empty(n, v - 1, h, w, device=rd.device, dtype=torch.float32) if want_levels and v > 1 else None
    _call("mvs_fusion_filter_dynamic", _lib.ptr(rd), _lib.ptr(xyd), float(dist_base), float(rel_diff_base), _lib.ptr(vis),
          _lib.ptr(geo), _lib.ptr(ave), _lib.ptr(levels), n, v, h, w)
    return vis, geo, ave, levels


def vis_filter_dynamic(ref_depth, reproj_xyd, dist_base=4, rel_diff_base=1300):
    """misc/fusion.py:155-168.  The reference returns (masks [n,v,v-1,h,w], mask [n,v,1,h,w]) and test.py:505-511 only
    uses ``masks.sum(dim=1)`` and ``mask``; this mirror returns (level_counts [n,v-1,h,w] = masks.sum(dim=1) as float,
    mask [n,v,1,h,w] bool) and never materialises the v x (v-1) mask stack."""
    vis, _, _, levels = _filter_dynamic(ref_depth, reproj_xyd, dist_base, rel_diff_base, True)
    return levels, vis.unsqueeze(2) > 0.5


def dynamic_filter_view(ref_depth, srcs_depth, ref_cam, srcs_cam, dist_base=4, rel_diff_base=1300, ref_conf=None,
                        prob_thresh=None):
    """The per-reference-view chain of test.py:487-514 (dynamic consistency checking): -> dict(mask, depth_ave, points,
    geo_mask, vis_mask, reproj_xyd)."""
    reproj_xyd = get_reproj_dynamic(ref_depth, srcs_depth, ref_cam, srcs_cam)
    vis, geo, ave, _ = _filter_dynamic(ref_depth, reproj_xyd, dist_base, rel_diff_base, False)
    geo_mask = geo.unsqueeze(1) > 0.5
    mask = geo_mask
    if ref_conf is not None and prob_thresh is not None:
        mask = bin_op_reduce([prob_filter(ref_conf, prob_thresh), geo_mask], torch.min)
    ave = ave.unsqueeze(1)
    return {"mask": mask, "depth_ave": ave, "points": world_points(ave, ref_cam), "geo_mask": geo_mask,
            "vis_mask": vis.unsqueeze(2) > 0.5, "reproj_xyd": reproj_xyd}


def world_points(depth, cam):
    """depth [n,1,h,w], cam [n,2,4,4] -> points [n,3,h,w] in world coordinates (test.py:433-435)."""
    n, _, h, w = depth.shape
    d = depth.float().reshape(n, h, w).contiguous()
    mats = _cam_block(cam)[..., :25].float().contiguous().to(d.device)
    _lib.require_cuda(d, mats)
    pts = torch.empty(n, 3, h, w, device=d.device, dtype=torch.float32)
    _call("mvs_fusion_points", _lib.ptr(d), _lib.ptr(mats), _lib.ptr(pts), n, h, w)
    return pts


def filter_view(ref_depth, srcs_depth, ref_cam, srcs_cam, img_dist_thresh, depth_thresh, vthresh, ref_conf=None,
                prob_thresh=None):
    """The per-reference-view chain of test.py:413-435: -> dict(mask [n,1,h,w] bool, depth_ave [n,1,h,w],
    points [n,3,h,w], vis_masks, vis_mask, reproj_xyd, in_range)."""
    reproj_xyd, in_range = get_reproj(ref_depth, srcs_depth, ref_cam, srcs_cam)
    vis_masks, vis_mask, ave = _filter(ref_depth, reproj_xyd, in_range, img_dist_thresh, depth_thresh, vthresh)
    mask = vis_mask
    if ref_conf is not None and prob_thresh is not None:
        mask = bin_op_reduce([prob_filter(ref_conf, prob_thresh), vis_mask], torch.min)
    return {"mask": mask, "depth_ave": ave, "points": world_points(ave, ref_cam), "vis_masks": vis_masks, "vis_mask": vis_mask,
            "reproj_xyd": reproj_xyd, "in_range": in_range}
