"""GPU parity against the committed golden fixtures (outputs of the real reference, tests/golden/):
the CUDA path, called through the C ABI, on the very inputs the reference produced them from."""
import os

import numpy as np
import pytest
import torch

from dmhomo_b200 import ops
from dmhomo_b200.compat import dgm, flow_and_mapping_operations as fmo, hem_net, hem_utils, losses, pixel_wise_mapping as pwm

pytestmark = pytest.mark.gpu
DEV = "cuda"
ATOL = 1e-4
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: (torch.from_numpy(z[k]) if z[k].dtype.kind in "fiub" and z[k].ndim > 0 else z[k]) for k in z.files}


def rel_fro(a, b):
    return ((a - b).norm() / b.norm()).item()


def test_dlt_golden():
    d = load("dlt")
    src, off = d["src"].to(DEV), d["off"].to(DEV)
    B = src.shape[0]
    h, w = [int(v) for v in d["hw"]]
    assert rel_fro(hem_utils.DLT(B)(src, src + off).cpu(), d["H"]) < 1e-5
    assert rel_fro(hem_net.DLT_solve(src.reshape(B, 8), off.reshape(B, 8)).cpu(), d["H_net"]) < 1e-5
    assert rel_fro(hem_utils.WarpMat(off.reshape(B, 8).clone(), (w, h), (w, h)).cpu(), d["H_warpmat"]) < 1e-5
    assert rel_fro(hem_utils.DLT_solve(d["mesh"].to(DEV), d["mesh_off"].to(DEV)).cpu(), d["H_mesh"]) < 1e-5


def test_get_flow_golden():
    d = load("get_flow")
    h, w = [int(v) for v in d["hw"]]
    H, Hm = d["H"].to(DEV), d["H_mesh"].to(DEV)
    assert torch.equal(ops.homography_to_flow(H, h, w).cpu(), d["flow"])                 # bit-exact
    assert torch.equal(ops.homography_to_flow(H, h, w, start=3).cpu(), d["flow_start3"])
    assert torch.equal(ops.homography_to_flow(Hm, h, w, divide=2).cpu(), d["flow_mesh"])


def test_s1_warp_golden():
    d = load("s1_warp")
    img = d["img"].to(DEV).requires_grad_(True)
    flow = d["flow"].to(DEV).requires_grad_(True)
    out = hem_utils.get_warp_flow(img, flow, start=int(d["start"]))
    assert torch.equal(out.detach().cpu(), d["out"])                                     # bit-exact pixels
    (out * d["grad_out"].to(DEV)).sum().backward()
    assert (img.grad.cpu() - d["grad_img"]).abs().max().item() < ATOL
    assert (flow.grad.cpu() - d["grad_flow"]).abs().max().item() < ATOL
    assert torch.equal(hem_utils.transformer(d["img"].to(DEV), d["vgrid"].to(DEV)).cpu(), d["out_transformer"])


def test_pipeline_golden():
    d = load("pipeline_h4pt")
    B, _, h, w = d["img1"].shape
    i1, i2 = d["img1"].to(DEV), d["img2"].to(DEV)
    Hf, Hb = d["Hf"].to(DEV), d["Hb"].to(DEV)   # stage isolation: the reference's own H
    w2, mf, ff = ops.warp(i2, Hf, kind=ops.PARAM_HOMOGRAPHY, return_mask=True, return_flow=True)
    w1, mb, fb = ops.warp(i1, Hb, kind=ops.PARAM_HOMOGRAPHY, return_mask=True, return_flow=True)
    assert torch.equal(ff.cpu(), d["flow_f"]) and torch.equal(fb.cpu(), d["flow_b"])
    assert torch.equal(mf.cpu(), d["mask_f"][:, 0].bool()) and torch.equal(mb.cpu(), d["mask_b"][:, 0].bool())
    assert torch.equal(w2.cpu(), d["w2"]) and torch.equal(w1.cpu(), d["w1"])
    # the lean forward kernel (out + mask only) must agree with the general one bit for bit
    w2b, mfb = ops.warp(i2, Hf, kind=ops.PARAM_HOMOGRAPHY, return_mask=True)
    assert torch.equal(w2b, w2) and torch.equal(mfb, mf)
    loss = ops.warp_loss([ops.WarpTerm(i2, i1, Hf), ops.WarpTerm(i1, i2, Hb)], kind=ops.PARAM_HOMOGRAPHY)
    assert abs(loss.item() - float(d["loss"])) < 1e-5
    # chained: our DLT feeding our warp
    src = torch.tensor([[0, 0], [w - 1, 0], [0, h - 1], [w - 1, h - 1]], dtype=torch.float32, device=DEV).repeat(B, 1, 1)
    assert rel_fro(ops.dlt4(src, src + d["off_f"].to(DEV)).cpu(), d["Hf"]) < 1e-5


def test_pipeline_basis_golden_through_the_reference_names():
    """cfg 2 as the reference's own statements issue it (tests/golden/make_golden.py: gen_basis -> basis product ->
    corner offsets -> DLT -> get_flow -> get_warp_flow x2 -> create_border_mask x2 -> LossL1 x2 -> backward), run here
    through the names compat.patch_reference() installs, against what the reference produced; then the fused op on the
    same inputs."""
    d = load("pipeline_basis")
    h, w = [int(v) for v in d["hw"]]
    B = d["img1"].shape[0]
    basis = hem_utils.gen_basis(h, w).to(DEV)
    leaves = [d[k].clone().to(DEV).requires_grad_(True) for k in ("img1", "img2", "w_f", "w_b")]
    c1, c2, wf, wb = leaves
    src = torch.tensor([[0, 0], [w - 1, 0], [0, h - 1], [w - 1, h - 1]], dtype=torch.float32, device=DEV).repeat(B, 1, 1)

    def corner_offsets(wt):
        fl = (basis.reshape(1, 8, -1) * wt).sum(1).reshape(B, 2, h, w)          # net.py:808-809, inline torch in the reference
        return torch.stack([fl[:, :, 0, 0], fl[:, :, 0, w - 1], fl[:, :, h - 1, 0], fl[:, :, h - 1, w - 1]], 1)

    Hf, Hb = hem_utils.DLT(B)(src, src + corner_offsets(wf)), hem_utils.DLT(B)(src, src + corner_offsets(wb))
    grid = hem_utils.get_grid(B, h, w, 0)
    ff, _ = hem_utils.get_flow(Hf.view(B, 1, 3, 3), grid, h, w, 1)
    fb, _ = hem_utils.get_flow(Hb.view(B, 1, 3, 3), grid, h, w, 1)
    w2, w1 = hem_utils.get_warp_flow(c2, ff), hem_utils.get_warp_flow(c1, fb)
    mf, mb = fmo.create_border_mask(ff).unsqueeze(1), fmo.create_border_mask(fb).unsqueeze(1)
    l1 = losses.LossL1(reduction="mean")
    loss = l1(mf * c1, mf * w2) + l1(mb * c2, mb * w1)
    loss.backward()
    assert rel_fro(Hf.detach().cpu(), d["Hf"]) < 1e-5 and rel_fro(Hb.detach().cpu(), d["Hb"]) < 1e-5
    assert (ff.detach().cpu() - d["flow_f"]).abs().max().item() < ATOL and (fb.detach().cpu() - d["flow_b"]).abs().max().item() < ATOL
    assert (w2.detach().cpu() - d["w2"]).abs().max().item() < ATOL and (w1.detach().cpu() - d["w1"]).abs().max().item() < ATOL
    assert abs(loss.item() - d["loss"].item()) < 1e-5
    assert (c1.grad.cpu() - d["g_img1"]).abs().max().item() < ATOL and (c2.grad.cpu() - d["g_img2"]).abs().max().item() < ATOL
    assert rel_fro(wf.grad.cpu(), d["g_wf"]) < 1e-3 and rel_fro(wb.grad.cpu(), d["g_wb"]) < 1e-3

    # the same step as ONE op (basis -> H -> both warps + masks + L1 + every gradient in one launch chain)
    f1, f2, fwf, fwb = [d[k].clone().to(DEV).requires_grad_(True) for k in ("img1", "img2", "w_f", "w_b")]
    floss, fHf, fHb = ops.basis_warp_loss(basis, f1, f2, fwf, fwb, return_homographies=True)
    floss.backward()
    assert rel_fro(fHf.cpu(), d["Hf"]) < 1e-5 and rel_fro(fHb.cpu(), d["Hb"]) < 1e-5
    assert abs(floss.item() - d["loss"].item()) < 1e-5
    assert (f1.grad.cpu() - d["g_img1"]).abs().max().item() < ATOL and (f2.grad.cpu() - d["g_img2"]).abs().max().item() < ATOL
    assert rel_fro(fwf.grad.cpu().reshape(B, 8, 1), d["g_wf"]) < 1e-3 and rel_fro(fwb.grad.cpu().reshape(B, 8, 1), d["g_wb"]) < 1e-3


def test_warp_images_golden():
    d = load("warp_images")
    out, flow = hem_utils.WarpImages(d["img"].to(DEV), d["H"].to(DEV), d["start"].to(DEV), tuple(int(v) for v in d["patch_wh"]))
    assert (flow.cpu() - d["flow"]).abs().max().item() < 1e-5
    assert (out.cpu() - d["out"]).abs().max().item() < ATOL


def test_grid_sample_warps_golden():
    d = load("grid_sample_warps")
    img, flow = d["img"].to(DEV), d["flow"].to(DEV)
    assert (pwm.warp(img, flow).cpu() - d["warp_zeros"]).abs().max().item() < ATOL
    assert (pwm.warp_with_mapping(img, flow + 3).cpu() - d["warp_mapping"]).abs().max().item() < ATOL
    assert (dgm.flow_warp(img, flow).cpu() - d["flow_warp"]).abs().max().item() < ATOL


def test_masks_golden():
    d = load("masks")
    flow = d["flow"].to(DEV)
    assert torch.equal(fmo.get_gt_correspondence_mask(flow).cpu(), d["corr"].bool())
    assert torch.equal(fmo.create_border_mask(flow).cpu(), d["border"])
    assert torch.equal(fmo.define_mask_zero_borders(d["image"].to(DEV)).cpu(), d["zero_border"].bool())
    assert torch.equal(fmo.convert_flow_to_mapping(flow).cpu(), d["mapping"])


def test_basis_golden():
    d = load("basis")
    h, w = [int(v) for v in d["hw"]]
    flow = ops.basis_combine(d["basis"].to(DEV), d["weight"].to(DEV), h, w)
    assert torch.equal(flow.cpu(), d["flow"])                                            # bit-exact


def test_dgm_golden():
    d = load("dgm_flow")
    h, w = [int(v) for v in d["hw"]]
    flow = ops.homography_to_flow_f64(d["H"].to(DEV), h, w)
    assert torch.equal(flow.cpu(), d["flow"])                                            # bit-exact (fp64 -> fp32)
    rgb = ops.flow_to_rgb(flow, in_channels_last=True, out_channels_last=True)
    assert (rgb.cpu() - d["rgb"]).abs().max().item() < 1e-5
    p = load("warp_perspective")
    out = ops.warp_perspective(p["img"].to(DEV), p["H"].to(DEV), (80, 64), channels_last=True)
    assert (out.cpu() - p["out"]).abs().max().item() < 1e-5                              # vs cv2 itself


def test_homo_gen_and_eval_points_golden():
    d = load("homo_gen")
    H = dgm.homo_gen(d["flow"].to(DEV)).cpu()
    for b in range(H.shape[0]):
        assert rel_fro(H[b], d["H"][b].double()) < 1e-8
    e = load("eval_points")
    err = torch.stack(losses.compute_eval_results({"pt_set": e["pts"].to(DEV)},
                                                  {"flow_f": e["flow_f"].to(DEV), "flow_b": e["flow_b"].to(DEV)}))
    assert (err.cpu() - e["err"]).abs().max().item() < 1e-5
