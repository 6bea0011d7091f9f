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
