"""CPU oracle of the depth-map fusion step (misc/fusion.py:69-118, test.py:413-435).  TEST INFRASTRUCTURE ONLY.

An independent restatement in torch (CPU, fp32 or fp64): camera algebra written per component with the reference's
"+ 1e-9" homogeneous divisions and pixel-centre convention, explicit four-tap bilinear gather instead of
``F.grid_sample``.  Pinned against the reference's own functions by ``oracle/make_golden.py::gen_fusion`` ->
``tests/golden/fusion.npz`` (the reference calls ``.cuda()`` inside ``get_pixel_grids``; the generator patches that
call to a no-op, nothing else).  Only tests import this file.
"""
import torch


def _grid(height, width, dtype):
    ys, xs = torch.meshgrid(torch.arange(height, dtype=dtype) + 0.5, torch.arange(width, dtype=dtype) + 0.5, indexing="ij")
    return xs, ys                                                              # misc/fusion.py:8-13


def _img2cam(kmat, u, v, depth):                                               # :24-29
    kinv = torch.linalg.inv(kmat.double()).to(u.dtype)
    px = kinv[0, 0] * u + kinv[0, 1] * v + kinv[0, 2]
    py = kinv[1, 0] * u + kinv[1, 1] * v + kinv[1, 2]
    pz = kinv[2, 0] * u + kinv[2, 1] * v + kinv[2, 2]
    s = pz + 1e-9
    return torch.stack([px / s * depth, py / s * depth, pz / s * depth, torch.ones_like(px)])


def _transform(mat, p):                                                        # :32-41 (E^-1 or E), / (w + 1e-9)
    q = torch.einsum("ij,j...->i...", mat.to(p.dtype), p)
    return q / (q[3:4] + 1e-9)


def _cam2img(kmat, c):                                                         # :44-48
    c3 = c[:3] / (c[3:4] + 1e-9)
    i = torch.einsum("ij,j...->i...", kmat.to(c.dtype), c3)
    return i / (i[2:3] + 1e-9)


def _chain(depth, u, v, cam_from, cam_to):
    """pixels (u, v) with `depth` in camera `cam_from` -> (image coords in `cam_to` [3,...], camera coords in `cam_to` [4,...])."""
    ext_from, k_from, ext_to, k_to = cam_from[0], cam_from[1, :3, :3], cam_to[0], cam_to[1, :3, :3]
    world = _transform(torch.linalg.inv(ext_from.double()), _img2cam(k_from, u, v, depth))
    cam = _transform(ext_to, world)
    return _cam2img(k_to, cam), cam


def get_reproj(ref_depth, srcs_depth, ref_cam, srcs_cam):
    """misc/fusion.py:79-98 (+ project_img :51-66).  Shapes as the reference: n1hw, nv1hw, n244, nv244."""
    n, v, _, h, w = srcs_depth.shape
    dt = ref_depth.dtype
    xs, ys = _grid(h, w, dt)
    out = torch.zeros(n, v, 3, h, w, dtype=dt)
    in_range = torch.zeros(n, v, 1, h, w, dtype=dt)
    for b in range(n):
        for s in range(v):
            # per source pixel: its surface point seen from the reference view  (:87-91)
            img_r, cam_r = _chain(srcs_depth[b, s, 0], xs, ys, srcs_cam[b, s], ref_cam[b])
            xyd = torch.stack([img_r[0], img_r[1], cam_r[2]])                                  # [3,h,w]
            # per reference pixel: where it lands in the source image  (:51-65)
            img_s, _ = _chain(ref_depth[b, 0], xs, ys, ref_cam[b], srcs_cam[b, s])
            gx = (img_s[0] / w * 2 - 1).clamp(-1.1, 1.1)
            gy = (img_s[1] / h * 2 - 1).clamp(-1.1, 1.1)
            in_range[b, s, 0] = ((gx >= -1) & (gx <= 1) & (gy >= -1) & (gy <= 1)).to(dt)
            ix, iy = (gx + 1) / 2 * (w - 1), (gy + 1) / 2 * (h - 1)                            # align_corners=True
            x0, y0 = torch.floor(ix), torch.floor(iy)
            acc = torch.zeros(3, h, w, dtype=dt)
            for dy in (0, 1):
                for dx in (0, 1):
                    fx, fy = x0 + dx, y0 + dy
                    wgt = (ix - x0 if dx else 1 - (ix - x0)) * (iy - y0 if dy else 1 - (iy - y0))
                    ok = (fx >= 0) & (fx <= w - 1) & (fy >= 0) & (fy <= h - 1)
                    qx, qy = fx.clamp(0, w - 1).long(), fy.clamp(0, h - 1).long()
                    acc = acc + torch.where(ok, wgt, torch.zeros_like(wgt)) * xyd[:, qy, qx]
            out[b, s] = acc
    return out, in_range


def vis_filter(ref_depth, reproj_xyd, in_range, img_dist_thresh, depth_thresh, vthresh):
    """misc/fusion.py:101-109."""
    n, v, _, h, w = reproj_xyd.shape
    xs, ys = _grid(h, w, ref_depth.dtype)
    dist = torch.sqrt((reproj_xyd[:, :, 0] - xs) ** 2 + (reproj_xyd[:, :, 1] - ys) ** 2)
    rd = reproj_xyd[:, :, 2]
    dref = ref_depth[:, 0].unsqueeze(1)
    ok = (dist < img_dist_thresh) & ((dref - rd).abs() < torch.maximum(dref.expand_as(rd), rd) * depth_thresh) & (in_range[:, :, 0] > 0.5)
    masks = ok.to(ref_depth.dtype).unsqueeze(2)
    return masks, masks.sum(dim=1) >= (vthresh - 1.1)


def ave_fusion(ref_depth, reproj_xyd, masks):
    """misc/fusion.py:112-114."""
    return ((reproj_xyd[:, :, 2:3] * masks).sum(dim=1) + ref_depth) / (masks.sum(dim=1) + 1)


def world_points(depth, cam):
    """test.py:433-435."""
    n, _, h, w = depth.shape
    xs, ys = _grid(h, w, depth.dtype)
    pts = []
    for b in range(n):
        world = _transform(torch.linalg.inv(cam[b, 0].double()), _img2cam(cam[b, 1, :3, :3], xs, ys, depth[b, 0]))
        pts.append(world[:3])
    return torch.stack(pts)


def prob_filter(ref_prob, prob_thresh):
    """misc/fusion.py:69-76."""
    mask = None
    for i, p in enumerate(prob_thresh):
        m = ref_prob[:, [i]] > p
        mask = m if mask is None else mask & m
    return mask
