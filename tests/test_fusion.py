"""Depth-map fusion: the CPU oracle is pinned to what the reference's own misc/fusion.py
functions returned (tests/golden/fusion.npz), and the kernels of csrc/fusion_kernels.cuh are run on the CPU thread
emulation (tests/emu) through the package's drop-in mirror mvsformer_b200/fusion.py and compared with both.
The GPU run of the same kernels is in tests/test_gpu_variants.py."""
import pytest
import torch

from mvsformer_b200 import fusion as Fu
from mvsformer_b200 import synthetic as S
from oracle import fusion_oracle as FO
from tests.emu import harness
from tests.helpers import checksum, load_golden, rel_l1


@pytest.fixture()
def emu(monkeypatch):
    return harness.install(monkeypatch.setattr)


def _case():
    g = load_golden("fusion.npz")
    case = S.make_fusion_case(int(g["views"]), int(g["height"]), int(g["width"]), seed=int(g["seed"]))
    assert checksum(case["src_depths"]) == pytest.approx(float(g["depth_checksum"]), rel=1e-12)
    return g, case


def _agree(a, b):
    return float((torch.as_tensor(a).bool() == torch.as_tensor(b).bool()).float().mean())


def test_fusion_oracle_matches_reference_golden():
    g, c = _case()
    xyd, inr = FO.get_reproj(c["ref_depth"], c["src_depths"], c["ref_cam"], c["src_cams"])
    assert _agree(inr, g["in_range"]) > 0.995
    both = torch.from_numpy(g["in_range"]).bool() & (inr > 0.5)                  # compare where both are inside the image
    sel = both.expand(-1, -1, 3, -1, -1)
    assert rel_l1(xyd[sel], torch.from_numpy(g["reproj_xyd"])[sel]) < 1e-4
    masks, mask = FO.vis_filter(c["ref_depth"], torch.from_numpy(g["reproj_xyd"]), torch.from_numpy(g["in_range"]), 1.0, 0.01, 2)
    assert _agree(masks, g["masks"]) == 1.0 and _agree(mask, g["mask"]) == 1.0
    ave = FO.ave_fusion(c["ref_depth"], torch.from_numpy(g["reproj_xyd"]), torch.from_numpy(g["masks"]))
    assert rel_l1(ave, g["ave"]) < 1e-6
    assert rel_l1(FO.world_points(torch.from_numpy(g["ave"]), c["ref_cam"]), g["points"]) < 1e-5
    assert _agree(FO.prob_filter(c["ref_conf"], [0.1, 0.2, 0.3]), g["prob_mask"]) == 1.0


def test_fusion_kernels_emulated_vs_reference_and_oracle(emu):
    g, c = _case()
    xyd, inr = Fu.get_reproj(c["ref_depth"], c["src_depths"], c["ref_cam"], c["src_cams"])
    assert xyd.shape == g["reproj_xyd"].shape and inr.shape == g["in_range"].shape
    assert _agree(inr, g["in_range"]) > 0.995
    both = (torch.from_numpy(g["in_range"]) > 0.5) & (inr > 0.5)
    sel = both.expand(-1, -1, 3, -1, -1)
    assert rel_l1(xyd[sel], torch.from_numpy(g["reproj_xyd"])[sel]) < 1e-4
    # fp64 oracle as the referee for the reprojection itself
    c64 = {k: v.double() for k, v in c.items()}
    xyd64, _ = FO.get_reproj(c64["ref_depth"], c64["src_depths"], c64["ref_cam"], c64["src_cams"])
    assert rel_l1(xyd[sel], xyd64[sel]) < 1e-4
    gx, gi = torch.from_numpy(g["reproj_xyd"]), torch.from_numpy(g["in_range"])
    masks, mask = Fu.vis_filter(c["ref_depth"], gx, gi, 1.0, 0.01, 2)
    assert masks.shape == g["masks"].shape and mask.shape == g["mask"].shape and mask.dtype == torch.bool
    assert _agree(masks, g["masks"]) == 1.0 and _agree(mask, g["mask"]) == 1.0
    ave = Fu.ave_fusion(c["ref_depth"], gx, torch.from_numpy(g["masks"]))
    assert ave.shape == g["ave"].shape and rel_l1(ave, g["ave"]) < 1e-6
    pts = Fu.world_points(torch.from_numpy(g["ave"]), c["ref_cam"])
    assert pts.shape == g["points"].shape and rel_l1(pts, g["points"]) < 1e-5
    pm = Fu.prob_filter(c["ref_conf"], [0.1, 0.2, 0.3])
    assert pm.shape == g["prob_mask"].shape and _agree(pm, g["prob_mask"]) == 1.0


def test_filter_view_chain_emulated(emu):
    """The whole per-reference-view chain (test.py:413-435) vs the reference's result: masks agree except for pixels
    whose reprojection error sits on a threshold, averaged depth and points agree where both masks hold."""
    g, c = _case()
    out = Fu.filter_view(c["ref_depth"], c["src_depths"], c["ref_cam"], c["src_cams"], 1.0, 0.01, 2,
                         ref_conf=c["ref_conf"], prob_thresh=[0.1, 0.2, 0.3])
    assert _agree(out["vis_masks"], g["masks"]) > 0.99 and _agree(out["vis_mask"], g["mask"]) > 0.99
    want_mask = torch.from_numpy(g["prob_mask"]).bool() & torch.from_numpy(g["mask"]).bool()
    assert _agree(out["mask"], want_mask) > 0.99
    same = (out["vis_masks"] == torch.from_numpy(g["masks"])).all(dim=1)         # [n,1,h,w]
    assert rel_l1(out["depth_ave"][same], torch.from_numpy(g["ave"])[same]) < 1e-5
    assert rel_l1(out["points"][same.expand(-1, 3, -1, -1)], torch.from_numpy(g["points"])[same.expand(-1, 3, -1, -1)]) < 1e-4


def test_fusion_edge_cases_emulated(emu):
    """Zero (filtered-out) source depths, points behind the camera and projections outside the image follow the
    reference's conventions: no consistent view, depth stays the reference's, nothing reads out of bounds."""
    c = S.make_fusion_case(3, 16, 24, seed=5, noise=0.0, outlier_frac=0.0)
    src = c["src_depths"].clone()
    src[:, 0] = 0.0                                                              # a view removed by the photometric filter
    cams = c["src_cams"].clone()
    cams[:, 1, 0, 0, 3] += 5000.0                                                # a view that sees nothing of the scene
    out = Fu.filter_view(c["ref_depth"], src, c["ref_cam"], cams, 1.0, 0.01, 2)
    assert float(out["vis_masks"].sum()) == 0.0 and not bool(out["vis_mask"].any())
    assert torch.equal(out["depth_ave"], c["ref_depth"])
    want = FO.vis_filter(c["ref_depth"], *FO.get_reproj(c["ref_depth"], src, c["ref_cam"], cams), 1.0, 0.01, 2)[0]
    assert float(want.sum()) == 0.0


def test_dynamic_fusion_emulated_vs_reference(emu):
    """get_reproj_dynamic / vis_filter_dynamic (misc/fusion.py:116-168) and the vote + averaging of test.py:502-511."""
    g, c = _case()
    xyd = Fu.get_reproj_dynamic(c["ref_depth"], c["src_depths"], c["ref_cam"], c["src_cams"])
    gx = torch.from_numpy(g["dyn_xyd"])
    assert xyd.shape == gx.shape
    finite = torch.isfinite(gx) & torch.isfinite(xyd)
    assert float(finite.float().mean()) > 0.99 and rel_l1(xyd[finite], gx[finite]) < 1e-4
    # filter on the reference's reprojection: every threshold decision must match
    levels, mask = Fu.vis_filter_dynamic(c["ref_depth"], gx, dist_base=4, rel_diff_base=1300)
    assert levels.shape == g["dyn_level_counts"].shape and torch.equal(levels, torch.from_numpy(g["dyn_level_counts"]))
    assert mask.shape == g["dyn_mask"].shape and _agree(mask, g["dyn_mask"]) == 1.0
    # whole chain from the depth maps
    out = Fu.dynamic_filter_view(c["ref_depth"], c["src_depths"], c["ref_cam"], c["src_cams"], 4, 1300)
    assert _agree(out["vis_mask"], g["dyn_mask"]) > 0.99 and _agree(out["geo_mask"], g["dyn_geo_mask"]) > 0.99
    same = (out["vis_mask"] == torch.from_numpy(g["dyn_mask"])).all(dim=1)
    assert rel_l1(out["depth_ave"][same], torch.from_numpy(g["dyn_ave"])[same]) < 1e-5
