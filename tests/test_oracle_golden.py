"""Pins oracle/port.py against tests/golden/*.npz - outputs of the REAL reference, generated in the
build container by tests/golden/make_golden.py (the reference ships no tests or vectors of its own,
SURVEY.md section 4).  Runs anywhere, CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import port

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: (torch.from_numpy(z[k]) if z[k].dtype.kind in "fiub" and z[k].ndim > 0 else z[k]) for k in z.files}


def rel_fro(a, b):
    return ((a - b).norm() / b.norm()).item()


def test_dlt_golden():
    d = load("dlt")
    src, off = d["src"], d["off"]
    B = src.shape[0]
    h, w = [int(v) for v in d["hw"]]
    # explicit inverse + matmul: same primitive as the reference -> same bits on the same host
    H = port.dlt4(src, src + off)
    assert rel_fro(H, d["H"]) < 1e-6
    assert rel_fro(port.dlt_solve_h4pt(src.reshape(B, 8), off.reshape(B, 8)), d["H_net"]) < 1e-6
    assert rel_fro(port.warp_mat(off.reshape(B, 8), (w, h), (w, h)), d["H_warpmat"]) < 1e-6
    assert torch.equal(port.mesh_source_points(B, h, w, 2), d["mesh"])
    assert rel_fro(port.dlt_solve_mesh(d["mesh"], d["mesh_off"]), d["H_mesh"]) < 1e-6


def test_get_flow_golden():
    d = load("get_flow")
    h, w = [int(v) for v in d["hw"]]
    assert torch.equal(port.homography_to_flow(d["H"], h, w)[0], d["flow"])
    assert torch.equal(port.homography_to_flow(d["H"], h, w, start=3)[0], d["flow_start3"])
    assert torch.equal(port.homography_to_flow(d["H_mesh"], h, w, divide=2)[0], d["flow_mesh"])


def test_s1_warp_golden():
    d = load("s1_warp")
    img = d["img"].clone().requires_grad_(True)
    flow = d["flow"].clone().requires_grad_(True)
    out = port.get_warp_flow(img, flow, start=int(d["start"]))
    assert torch.equal(out, d["out"])
    (out * d["grad_out"]).sum().backward()
    assert (img.grad - d["grad_img"]).abs().max().item() < 1e-5
    assert (flow.grad - d["grad_flow"]).abs().max().item() < 1e-4
    assert torch.equal(port.s1_sample(d["img"], d["vgrid"][:, 0], d["vgrid"][:, 1]), d["out_transformer"])


def test_pipeline_golden():
    d = load("pipeline_h4pt")
    r = port.pipeline_h4pt(d["img1"], d["img2"], d["off_f"], d["off_b"])
    assert rel_fro(r["Hf"], d["Hf"]) < 1e-6 and rel_fro(r["Hb"], d["Hb"]) < 1e-6
    # stage-isolated: the reference's own H through the port's flow / warp / mask / loss
    ff = port.homography_to_flow(d["Hf"], 72, 128)[0]
    fb = port.homography_to_flow(d["Hb"], 72, 128)[0]
    assert torch.equal(ff, d["flow_f"]) and torch.equal(fb, d["flow_b"])
    assert torch.equal(port.get_warp_flow(d["img2"], ff), d["w2"])
    assert torch.equal(port.get_warp_flow(d["img1"], fb), d["w1"])
    mf, mb = port.border_mask(ff).unsqueeze(1), port.border_mask(fb).unsqueeze(1)
    assert torch.equal(mf, d["mask_f"]) and torch.equal(mb, d["mask_b"])
    loss = port.masked_l1(mf, d["img1"], d["w2"]) + port.masked_l1(mb, d["img2"], d["w1"])
    assert abs(loss.item() - float(d["loss"])) < 1e-6


def test_warp_images_golden():
    d = load("warp_images")
    out, flow = port.warp_images_s1b(d["img"], d["H"], d["start"], tuple(int(v) for v in d["patch_wh"]))
    assert (flow - d["flow"]).abs().max().item() < 1e-5      # bmm rounding belongs to the host BLAS
    assert (out - d["out"]).abs().max().item() < 1e-4


def test_grid_sample_warps_golden():
    d = load("grid_sample_warps")
    assert (port.warp_zeros(d["img"], d["flow"]) - d["warp_zeros"]).abs().max().item() < 1e-6
    assert (port.warp_with_mapping(d["img"], d["flow"] + 3) - d["warp_mapping"]).abs().max().item() < 1e-6
    assert (port.flow_warp(d["img"], d["flow"]) - d["flow_warp"]).abs().max().item() < 1e-6


def test_masks_golden():
    d = load("masks")
    assert torch.equal(port.correspondence_mask(d["flow"]), d["corr"].bool())
    assert torch.equal(port.border_mask(d["flow"]), d["border"])
    assert torch.equal(port.zero_border_mask(d["image"]), d["zero_border"].bool())
    assert torch.equal(d["flow"] + port.pixel_grid(2, 20, 30, homogeneous=False), d["mapping"])


def test_basis_golden():
    d = load("basis")
    h, w = [int(v) for v in d["hw"]]
    basis = port.gen_basis(h, w)
    # fp32 QR is LAPACK-backend dependent: same host -> same bits; elsewhere close
    assert (basis.reshape(8, -1) - d["basis"]).abs().max().item() < 1e-4
    flow = port.basis_combine(d["basis"].reshape(1, 8, -1), d["weight"], h, w)
    assert torch.equal(flow, d["flow"])


def test_dgm_flow_golden():
    d = load("dgm_flow")
    h, w = [int(v) for v in d["hw"]]
    Hs = d["H"].numpy()
    for i in range(Hs.shape[0]):
        f = port.homo_to_flow_np(Hs[i], h, w)
        assert np.array_equal(f, d["flow"][i].numpy())
        assert np.abs(port.flow_to_image(f) - d["rgb"][i].numpy()).max() < 1e-6
    mx, my = port.homography_to_mapping_np((h, w), Hs[0])
    assert np.abs(mx - d["map_x"].numpy()).max() < 1e-5 and np.abs(my - d["map_y"].numpy()).max() < 1e-5


def test_warp_perspective_golden():
    d = load("warp_perspective")
    imgs, Hs, out = d["img"].numpy(), d["H"].numpy(), d["out"].numpy()
    for i in range(imgs.shape[0]):
        # the restatement of cv2's 1/32-pixel fixed point path (what the CUDA kernel implements)
        assert np.abs(port.warp_perspective_emul(imgs[i], Hs[i], (80, 64)) - out[i]).max() < 1e-5


def test_homo_gen_golden():
    d = load("homo_gen")
    H = port.homo_gen(d["flow"])
    for b in range(H.shape[0]):
        assert rel_fro(H[b].double(), d["H"][b].double()) < 1e-8


def test_eval_points_golden():
    d = load("eval_points")
    err = torch.stack(port.eval_point_errors(d["pts"], d["flow_f"], d["flow_b"]))
    assert (err - d["err"]).abs().max().item() < 1e-6


# ---- the data formats either side of the path ----------------------------------
def test_pairs_u8_golden():
    """The uint8 pair format -> grey full / patch / RGB tensors, against DGMTrainData.data_aug itself."""
    d = load("pairs_u8")
    crop = tuple(int(v) for v in d["crop"])
    for b in range(d["img12"].shape[0]):
        full, patch, rgb = port.pairs_u8_to_gray(d["img12"][b].numpy(), d["start"][b].tolist(), crop)
        assert torch.equal(full, d["gray_full"][b])
        assert torch.equal(patch, d["gray_patch"][b])
        assert torch.equal(rgb, d["rgb_full"][b])


def test_gt_flow_golden():
    d = load("gt_flow")
    h, w = [int(v) for v in d["hw"]]
    for b in range(d["H"].shape[0]):
        Hm = d["H"][b].numpy()
        assert torch.equal(port.homo_convert_to_flow(Hm, (h, w))[0], d["flow"][b])
        assert np.array_equal(port.homo_scale(360, 640, Hm, h, w), d["H_scaled"][b].numpy())


def test_upsample_golden():
    d = load("upsample")
    fl = d["flow"]
    for name, (rate,) in {"x4_rate": (True,), "odd_rate": (True,), "down": (False,), "x2_norate": (False,)}.items():
        tgt = torch.zeros(1, 1, *d[name].shape[-2:])
        assert torch.equal(port.upsample2d_flow_as(fl, tgt, if_rate=rate), d[name]), name


def test_pipeline_basis_golden():
    """cfg 2 (the headline configuration) chained through the reference's own functions, gradients included."""
    d = load("pipeline_basis")
    h, w = [int(v) for v in d["hw"]]
    leaves = [d[k].clone().requires_grad_(True) for k in ("img1", "img2", "w_f", "w_b")]
    out = port.pipeline_basis(leaves[0], leaves[1], port.gen_basis(h, w).reshape(1, 8, -1), leaves[2], leaves[3],
                              variant="dlt", backward=True)
    assert rel_fro(out["Hf"].detach(), d["Hf"]) < 1e-6 and rel_fro(out["Hb"].detach(), d["Hb"]) < 1e-6
    assert (out["flow_f"].detach() - d["flow_f"]).abs().max().item() < 1e-4
    assert (out["w2"].detach() - d["w2"]).abs().max().item() < 1e-4 and (out["w1"].detach() - d["w1"]).abs().max().item() < 1e-4
    assert abs(out["loss"].item() - d["loss"].item()) < 1e-6
    assert (leaves[0].grad - d["g_img1"]).abs().max().item() < 1e-6
    assert (leaves[1].grad - d["g_img2"]).abs().max().item() < 1e-6
    assert rel_fro(leaves[2].grad, d["g_wf"]) < 1e-3 and rel_fro(leaves[3].grad, d["g_wb"]) < 1e-3


def test_warp_perspective_u8_emulation_matches_cv2():
    """The uint8 path of cv2.warpPerspective (the reference warps the uint8 halves of its generated samples,
    DGM/generate_nyps_to_single_case.py:15): the oracle's restatement of OpenCV's fixed-point remap is bit-identical to
    cv2 itself - OpenCV is unpinned by the reference, cv2 (4.13 here) is the oracle of record."""
    import cv2
    import numpy as np

    from dmhomo_b200 import synth

    rs = np.random.default_rng(7)
    Hs = synth.homographies_360x640(3, synth.generator(), 32.0)
    for i, (h, w) in enumerate(((256, 256), (90, 160), (100, 132))):
        img = rs.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        Hm = np.diag([w / 640, h / 360, 1.0]) @ Hs[i] @ np.diag([640 / w, 360 / h, 1.0])
        assert np.array_equal(cv2.warpPerspective(img, Hm, (w, h)), port.warp_perspective_emul_u8(img, Hm, (w, h)))
