"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): validity masks and integer sample indices bit-exact;
homographies within 1e-5 relative (Frobenius); warped pixels, losses, gradients within 1e-4
absolute in fp32.  Where the kernels reproduce the reference's rounding order the tests ask
for more (bit-exact flows / coordinates / warped pixels).
"""
import os

import numpy as np
import pytest
import torch

from dmhomo_b200 import ops, synth
from dmhomo_b200.compat import dgm, flow_and_mapping_operations as fmo, hem_net, hem_utils, losses, pixel_wise_mapping as pwm
from oracle import port

pytestmark = pytest.mark.gpu
DEV = "cuda"
ATOL = 1e-4  # north_star tolerance for pixels / losses / gradients


def g(seed):
    return torch.Generator().manual_seed(seed)


def rel_fro(a, b):
    return ((a - b).norm() / b.norm()).item()


def close_to_fp64(name, cuda, ref32, ref64, atol=ATOL):
    """Sums of many signed terms (dL/dH, dL/dweights): the CUDA result has to be as close to an fp64 evaluation of the
    same pipeline as the reference's own fp32 arithmetic is, or within north_star's 1e-4 - whichever is larger."""
    e_cuda = (cuda.detach().double().cpu() - ref64).abs().max().item()
    e_ref = (ref32.detach().double() - ref64).abs().max().item()
    if os.environ.get("DMH_TEST_REPORT"):
        print(f"[margin] {name}: cuda {e_cuda:.3e} oracle {e_ref:.3e} bound {max(atol, 1.05 * e_ref):.3e}")
    assert e_cuda <= max(atol, 1.05 * e_ref), f"{name}: |cuda - fp64| = {e_cuda:.3e} vs |oracle fp32 - fp64| = {e_ref:.3e}"


# ---------------------------------------------------------------------------------- DLT
@pytest.mark.parametrize("B,h,w,rho", [(16, 360, 640, 32.0), (64, 320, 576, 32.0), (7, 1080, 1920, 64.0)])
def test_dlt4_matches_oracle(B, h, w, rho):
    src = port.corner_points(B, h, w)
    dst = src + synth.corner_offsets(B, rho, g(1))
    H_ref = port.dlt4(src, dst)
    H = ops.dlt4(src.to(DEV), dst.to(DEV)).cpu()
    assert H.shape == (B, 3, 3)
    assert torch.all(H[:, 2, 2] == 1)
    for b in range(B):
        assert rel_fro(H[b], H_ref[b]) < 1e-5
    # zero offsets -> identity (SURVEY section 4 property 1)
    I = ops.dlt4(src.to(DEV), src.to(DEV)).cpu()
    assert torch.allclose(I, torch.eye(3).expand(B, 3, 3), atol=1e-6)


def test_dlt4_backward_matches_autograd():
    B, h, w = 8, 90, 160
    src = port.corner_points(B, h, w)
    dst = (src + synth.corner_offsets(B, 8.0, g(2)))
    gH = torch.randn(B, 3, 3, generator=g(3))
    d64 = dst.double().requires_grad_(True)
    s64 = src.double().requires_grad_(True)
    A, b = port._dlt_system(s64, d64)
    h8 = torch.linalg.solve(A, b).view(B, 8)
    H64 = torch.cat([h8, h8.new_ones(B, 1)], 1).view(B, 3, 3)
    (H64 * gH.double()).sum().backward()
    dg = dst.to(DEV).requires_grad_(True)
    sg = src.to(DEV).requires_grad_(True)
    (ops.dlt4(sg, dg) * gH.to(DEV)).sum().backward()
    scale = d64.grad.abs().max().item()
    assert (dg.grad.cpu().double() - d64.grad).abs().max().item() < 1e-4 * max(scale, 1.0)
    assert (sg.grad.cpu().double() - s64.grad).abs().max().item() < 1e-4 * max(s64.grad.abs().max().item(), 1.0)


def test_compat_dlt_variants():
    B, h, w = 5, 64, 96
    src = port.corner_points(B, h, w)
    off = synth.corner_offsets(B, 6.0, g(4))
    a = hem_utils.DLT(B)(src.to(DEV), (src + off).to(DEV)).cpu()
    assert rel_fro(a, port.dlt4(src, src + off)) < 1e-5
    a = hem_net.DLT_solve(src.reshape(B, 8).to(DEV), off.reshape(B, 8).to(DEV)).cpu()
    assert rel_fro(a, port.dlt_solve_h4pt(src.reshape(B, 8), off.reshape(B, 8))) < 1e-5
    off2 = off.reshape(B, 8).clone().to(DEV)
    keep = off2.clone()
    a = hem_utils.WarpMat(off2, (w, h), (w, h)).cpu()
    assert rel_fro(a, port.warp_mat(off.reshape(B, 8), (w, h), (w, h))) < 1e-5
    assert torch.allclose(off2, keep, atol=1e-5)  # scaled and scaled back
    d = 2
    mesh = port.mesh_source_points(B, h, w, d)
    moff = torch.randn(B, 2, d + 1, d + 1, generator=g(5))
    a = hem_utils.DLT_solve(mesh.to(DEV), moff.to(DEV)).cpu()
    assert rel_fro(a, port.dlt_solve_mesh(mesh, moff)) < 1e-5
    assert torch.equal(hem_utils.get_src_p(B, h, w, d).cpu(), mesh)
    with pytest.raises(AssertionError):
        hem_utils.DLT(B)(src.to(DEV), src.to(DEV), method="bogus")


# ---------------------------------------------------------------------------------- H -> flow
@pytest.mark.parametrize("divide,start", [(1, 0), (2, 0), (1, 3)])
def test_homography_to_flow_bit_exact(divide, start):
    B, h, w = 4, 96, 160
    if divide == 1:
        src = port.corner_points(B, h, w)
        H = port.dlt4(src, src + synth.corner_offsets(B, 8.0, g(6)))
    else:
        mesh = port.mesh_source_points(B, h, w, divide)
        H = port.dlt_solve_mesh(mesh, torch.randn(B, 2, divide + 1, divide + 1, generator=g(7)))
    f_ref, _ = port.homography_to_flow(H, h, w, start=start, divide=divide)
    f = ops.homography_to_flow(H.to(DEV), h, w, divide=divide, start=start).cpu()
    assert torch.equal(f, f_ref)
    # compat get_flow with the reference's calling convention
    grid = port.pixel_grid(B, h, w, start)
    f2, vg = hem_utils.get_flow(H.reshape(B, divide * divide, 3, 3).to(DEV), grid.to(DEV), h, w, divide)
    assert torch.equal(f2.cpu(), f_ref) and torch.equal(vg.cpu(), grid[:, :2])


def test_homography_to_flow_epsilon_rule():
    # T == 0 on a line of pixels: the reference adds 1e-6 there
    H = torch.tensor([[[1.0, 0, 0], [0, 1, 0], [1.0, 0, -5.0]]])
    f_ref, _ = port.homography_to_flow(H, 4, 12)
    f = ops.homography_to_flow(H.to(DEV), 4, 12).cpu()
    assert torch.equal(f, f_ref)


def test_homography_to_flow_backward():
    B, h, w = 3, 40, 56
    src = port.corner_points(B, h, w)
    H = port.dlt4(src, src + synth.corner_offsets(B, 5.0, g(8)))
    gf = torch.randn(B, 2, h, w, generator=g(9))
    Hc = H.clone().requires_grad_(True)
    (port.homography_to_flow(Hc, h, w)[0] * gf).sum().backward()
    Hg = H.to(DEV).requires_grad_(True)
    (ops.homography_to_flow(Hg, h, w) * gf.to(DEV)).sum().backward()
    ref = Hc.grad
    assert ((Hg.grad.cpu() - ref).abs() / (ref.abs() + 1.0)).max().item() < 1e-4


def test_homography_to_flow_f64():
    rng = np.random.default_rng(10)
    Hs = np.stack([np.eye(3) + rng.normal(size=(3, 3)) * np.array([[1e-2, 1e-2, 3], [1e-2, 1e-2, 3], [1e-5, 1e-5, 0]])
                   for _ in range(3)])
    out = ops.homography_to_flow_f64(torch.as_tensor(Hs, device=DEV), 40, 56).cpu().numpy()
    for i in range(3):
        ref = port.homo_to_flow_np(Hs[i], 40, 56)
        assert np.array_equal(out[i], ref)
        assert np.array_equal(dgm.homo_to_flow(Hs[i].reshape(1, 1, 3, 3), 40, 56), ref)
    mx, my = fmo.from_homography_to_pixel_wise_mapping((40, 56), Hs[0])
    rx, ry = port.homography_to_mapping_np((40, 56), Hs[0])
    assert np.abs(mx - rx).max() < 1e-5 and np.abs(my - ry).max() < 1e-5


# ---------------------------------------------------------------------------------- S1 warp
@pytest.mark.parametrize("C,h,w,Hs,Ws,start", [(1, 45, 70, 45, 70, 0), (3, 33, 65, 40, 80, 2), (5, 16, 16, 16, 16, 0)])
def test_get_warp_flow_bit_exact(C, h, w, Hs, Ws, start):
    B = 3
    img = torch.rand(B, C, Hs, Ws, generator=g(11))
    flow = torch.randn(B, 2, h, w, generator=g(12)) * 7
    ref, idx_ref = port.get_warp_flow(img, flow, start=start, return_indices=True)
    out, mask, idx = ops.warp(img.to(DEV), flow.to(DEV), start=start, return_mask=True, return_indices=True)
    assert torch.equal(idx.cpu(), idx_ref)                       # integer sample indices: bit-exact
    assert torch.equal(mask.cpu(), port.correspondence_mask(flow))  # validity mask: bit-exact
    assert torch.equal(out.cpu(), ref)                           # pixels: bit-exact (same rounding order)
    assert torch.equal(hem_utils.get_warp_flow(img.to(DEV), flow.to(DEV), start).cpu(), ref)


def test_identity_flow_zeroes_last_row_and_col():
    img = torch.rand(2, 1, 20, 30, generator=g(13))
    out = hem_utils.get_warp_flow(img.to(DEV), torch.zeros(2, 2, 20, 30, device=DEV)).cpu()
    assert torch.equal(out[:, :, :-1, :-1], img[:, :, :-1, :-1])
    assert torch.all(out[:, :, -1, :] == 0) and torch.all(out[:, :, :, -1] == 0)


def test_transformer_coords():
    B, C, h, w = 2, 2, 24, 40
    img = torch.rand(B, C, h, w, generator=g(14))
    vgrid = port.pixel_grid(B, h, w)[:, :2] + torch.randn(B, 2, h, w, generator=g(15)) * 4
    ref = port.s1_sample(img, vgrid[:, 0], vgrid[:, 1])
    assert torch.equal(hem_utils.transformer(img.to(DEV), vgrid.to(DEV)).cpu(), ref)
    nhwc = hem_utils.transformer(img.to(DEV), vgrid.to(DEV), train=False).cpu()
    assert torch.equal(nhwc, ref.permute(0, 2, 3, 1))


def test_warp_backward_flow_param():
    B, C, h, w = 2, 3, 24, 36
    img = torch.rand(B, C, h + 4, w + 2, generator=g(16))
    flow = torch.randn(B, 2, h, w, generator=g(17)) * 5
    go = torch.randn(B, C, h, w, generator=g(18))
    ic, fc = img.clone().requires_grad_(True), flow.clone().requires_grad_(True)
    (port.get_warp_flow(ic, fc) * go).sum().backward()
    ig, fg = img.to(DEV).requires_grad_(True), flow.to(DEV).requires_grad_(True)
    (ops.warp(ig, fg) * go.to(DEV)).sum().backward()
    assert (ig.grad.cpu() - ic.grad).abs().max().item() < ATOL
    assert (fg.grad.cpu() - fc.grad).abs().max().item() < ATOL


def test_warp_homography_param_matches_flow_path():
    B, C, h, w = 4, 1, 90, 160
    img = synth.noise_images(B, C, h, w, g(19))
    src = port.corner_points(B, h, w)
    H = port.dlt4(src, src + synth.corner_offsets(B, 8.0, g(20)))
    flow_ref, _ = port.homography_to_flow(H, h, w)
    ref, idx_ref = port.get_warp_flow(img, flow_ref, return_indices=True)
    out, mask, flow, idx = ops.warp(img.to(DEV), H.to(DEV), kind=ops.PARAM_HOMOGRAPHY, return_mask=True,
                                    return_flow=True, return_indices=True)
    assert torch.equal(flow.cpu(), flow_ref)
    assert torch.equal(idx.cpu(), idx_ref)
    assert torch.equal(mask.cpu(), port.correspondence_mask(flow_ref))
    assert torch.equal(out.cpu(), ref)
    # backward to the image and to H
    go = torch.randn(B, C, h, w, generator=g(21))
    ic, Hc = img.clone().requires_grad_(True), H.clone().requires_grad_(True)
    (port.get_warp_flow(ic, port.homography_to_flow(Hc, h, w)[0]) * go).sum().backward()
    ig, Hg = img.to(DEV).requires_grad_(True), H.to(DEV).requires_grad_(True)
    (ops.warp(ig, Hg, kind=ops.PARAM_HOMOGRAPHY) * go.to(DEV)).sum().backward()
    assert (ig.grad.cpu() - ic.grad).abs().max().item() < ATOL
    assert ((Hg.grad.cpu() - Hc.grad).abs() / (Hc.grad.abs() + 1.0)).max().item() < 1e-3


def test_warp_mesh_homography():
    B, C, h, w, d = 2, 1, 64, 96, 2
    img = torch.rand(B, C, h, w, generator=g(22))
    mesh = port.mesh_source_points(B, h, w, d)
    H = port.dlt_solve_mesh(mesh, torch.randn(B, 2, d + 1, d + 1, generator=g(23)))
    flow_ref, _ = port.homography_to_flow(H, h, w, divide=d)
    out = ops.warp(img.to(DEV), H.to(DEV), kind=ops.PARAM_HOMOGRAPHY, divide=d).cpu()
    assert torch.equal(out, port.get_warp_flow(img, flow_ref))
    go = torch.randn(B, C, h, w, generator=g(24))
    Hc = H.clone().requires_grad_(True)
    (port.get_warp_flow(img, port.homography_to_flow(Hc, h, w, divide=d)[0]) * go).sum().backward()
    Hg = H.to(DEV).requires_grad_(True)
    (ops.warp(img.to(DEV), Hg, kind=ops.PARAM_HOMOGRAPHY, divide=d) * go.to(DEV)).sum().backward()
    assert ((Hg.grad.cpu() - Hc.grad).abs() / (Hc.grad.abs() + 1.0)).max().item() < 1e-3


def test_warp_basis_param():
    B, C, h, w = 3, 1, 32, 48
    img = torch.rand(B, C, h, w, generator=g(25))
    basis = hem_utils.gen_basis(h, w)
    assert torch.equal(basis, port.gen_basis(h, w))
    wt = synth.basis_weights(B, g(26))
    flow_ref = port.basis_combine(basis.reshape(1, 8, -1), wt, h, w)
    assert torch.equal(ops.basis_combine(basis.to(DEV), wt.to(DEV), h, w).cpu(), flow_ref)
    assert torch.equal(hem_net.basis_flow(basis.reshape(1, 8, -1).to(DEV), wt.to(DEV), h, w).cpu(), flow_ref)
    out, flow = ops.warp(img.to(DEV), wt.to(DEV), kind=ops.PARAM_BASIS8, basis=basis.to(DEV), return_flow=True)
    assert torch.equal(flow.cpu(), flow_ref)
    assert torch.equal(out.cpu(), port.get_warp_flow(img, flow_ref))
    off = ops.basis_corner_offsets(basis.to(DEV), wt.to(DEV), h, w).cpu()
    assert torch.equal(off, port.basis_corner_offsets(basis.reshape(1, 8, -1), wt, h, w))
    # gradients to the 8 weights through both routes
    go = torch.randn(B, C, h, w, generator=g(27))
    wc = wt.clone().requires_grad_(True)
    (port.get_warp_flow(img, port.basis_combine(basis.reshape(1, 8, -1), wc, h, w)) * go).sum().backward()
    wg = wt.to(DEV).requires_grad_(True)
    (ops.warp(img.to(DEV), wg, kind=ops.PARAM_BASIS8, basis=basis.to(DEV)) * go.to(DEV)).sum().backward()
    assert (wg.grad.cpu() - wc.grad).abs().max().item() < 1e-3 * max(1.0, wc.grad.abs().max().item())
    wg2 = wt.to(DEV).requires_grad_(True)
    (ops.warp(img.to(DEV), ops.basis_combine(basis.to(DEV), wg2, h, w)) * go.to(DEV)).sum().backward()
    assert (wg2.grad.cpu() - wc.grad).abs().max().item() < 1e-3 * max(1.0, wc.grad.abs().max().item())
    goff = torch.randn(B, 4, 2, generator=g(28))
    wc2 = wt.clone().requires_grad_(True)
    (port.basis_corner_offsets(basis.reshape(1, 8, -1), wc2, h, w) * goff).sum().backward()
    wg3 = wt.to(DEV).requires_grad_(True)
    (ops.basis_corner_offsets(basis.to(DEV), wg3, h, w) * goff.to(DEV)).sum().backward()
    assert (wg3.grad.cpu() - wc2.grad).abs().max().item() < 1e-5


def test_warp_images_s1b():
    B, C, h, w = 2, 1, 48, 64
    img = torch.rand(B, C, h, w, generator=g(29))
    src = port.corner_points(B, 32, 40)
    H = port.dlt4(src, src + synth.corner_offsets(B, 4.0, g(30)))
    start = torch.tensor([[3.0, 5.0], [20.0, 12.0]]).view(B, 2, 1, 1)
    ref, flow_ref = port.warp_images_s1b(img, H, start, (40, 32))
    out, flow = hem_utils.WarpImages(img.to(DEV), H.to(DEV), start.to(DEV), (40, 32))
    # the reference forms H @ grid with torch.bmm, whose summation order (FMA chain) belongs to the host
    # BLAS: the kernel follows MKL's order (bit-exact on that host), anything else is <= 1 ulp away
    assert (flow.cpu() - flow_ref).abs().max().item() < 1e-5
    assert (out.cpu() - ref).abs().max().item() < ATOL
    out2, _ = hem_utils.Transform(H.to(DEV), img.to(DEV), start.to(DEV), (40, 32), start_zero=True)
    ref2, _ = port.warp_images_s1b(img, H, torch.zeros_like(start), (40, 32))
    assert (out2.cpu() - ref2).abs().max().item() < ATOL


# ---------------------------------------------------------------------------------- S2 / S3
def test_grid_sample_warps():
    B, C, h, w = 2, 3, 32, 48
    img = torch.rand(B, C, h, w, generator=g(31))
    flow = torch.randn(B, 2, h, w, generator=g(32)) * 6
    assert (pwm.warp(img.to(DEV), flow.to(DEV)).cpu() - port.warp_zeros(img, flow)).abs().max().item() < ATOL
    assert (pwm.warp_with_mapping(img.to(DEV), (flow + 3).to(DEV)).cpu() -
            port.warp_with_mapping(img, flow + 3)).abs().max().item() < ATOL
    assert (dgm.flow_warp(img.to(DEV), flow.to(DEV)).cpu() - port.flow_warp(img, flow)).abs().max().item() < ATOL
    # gradients of the DGM warp (to the image; and to the flow)
    go = torch.randn(B, C, h, w, generator=g(33))
    for fn_ref, fn in ((port.flow_warp, dgm.flow_warp), (port.warp_zeros, pwm.warp)):
        ic, fc = img.clone().requires_grad_(True), flow.clone().requires_grad_(True)
        (fn_ref(ic, fc) * go).sum().backward()
        ig, fg = img.to(DEV).requires_grad_(True), flow.to(DEV).requires_grad_(True)
        (fn(ig, fg) * go.to(DEV)).sum().backward()
        assert (ig.grad.cpu() - ic.grad).abs().max().item() < ATOL
        assert (fg.grad.cpu() - fc.grad).abs().max().item() < 2e-4


# ---------------------------------------------------------------------------------- masks, L1
def test_masks_bit_exact():
    flow = torch.randn(2, 2, 20, 30, generator=g(34)) * 12
    assert torch.equal(fmo.get_gt_correspondence_mask(flow.to(DEV)).cpu(), port.correspondence_mask(flow))
    assert torch.equal(fmo.create_border_mask(flow.to(DEV)).cpu(), port.border_mask(flow))
    assert torch.equal(fmo.get_gt_correspondence_mask(flow[0].to(DEV)).cpu(), port.correspondence_mask(flow[0]))
    cl = flow.permute(0, 2, 3, 1).contiguous()
    assert torch.equal(fmo.get_gt_correspondence_mask(cl.to(DEV)).cpu(), port.correspondence_mask(flow))
    img = torch.rand(2, 3, 8, 9, generator=g(35))
    img[:, :, :3] = 0
    assert torch.equal(fmo.define_mask_zero_borders(img.to(DEV)).cpu(), port.zero_border_mask(img))
    m = fmo.convert_flow_to_mapping(flow.to(DEV)).cpu()
    assert torch.equal(m, flow + port.pixel_grid(2, 20, 30, homogeneous=False))
    assert torch.equal(fmo.convert_mapping_to_flow(m.to(DEV)).cpu(), m - port.pixel_grid(2, 20, 30, homogeneous=False))


def test_l1_loss():
    a = torch.rand(2, 1, 33, 47, generator=g(36))
    b = torch.rand(2, 1, 33, 47, generator=g(37))
    m = (torch.rand(2, 1, 33, 47, generator=g(38)) > 0.3).float()
    ref = port.masked_l1(m, a, b)
    ag = a.to(DEV).requires_grad_(True)
    out = losses.LossL1()(m.to(DEV) * ag, m.to(DEV) * b.to(DEV))
    assert abs(out.item() - ref.item()) < 1e-6
    out.backward()
    ac = a.clone().requires_grad_(True)
    port.masked_l1(m, ac, b).backward()
    assert (ag.grad.cpu() - ac.grad).abs().max().item() < 1e-7


# ---------------------------------------------------------------------------------- fused loss
def _pipeline_inputs(B, C, h, w, rho, seed, smooth):
    gen = g(seed)
    mk = synth.smooth_images if smooth else synth.noise_images
    img1, img2 = mk(B, C, h, w, gen), mk(B, C, h, w, gen)
    return img1, img2, synth.corner_offsets(B, rho, gen), synth.corner_offsets(B, rho, gen)


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("C", [1, 3, 4, 6])
def test_fused_bidirectional_loss_stage_isolated(fused, C):
    """Oracle H fed to both sides (stage isolation, noise images): loss + gradients to both images and H.  C = 4 / 6:
    feature maps, walked as channel groups of 1 / 3 whose loss and dL/dH sums add up per sample."""
    B, h, w = 4, 72, 128
    img1, img2, off_f, off_b = _pipeline_inputs(B, C, h, w, 8.0, 40, smooth=False)
    src = port.corner_points(B, h, w)
    Hf, Hb = port.dlt4(src, src + off_f), port.dlt4(src, src + off_b)
    i1c, i2c = img1.clone().requires_grad_(True), img2.clone().requires_grad_(True)
    Hfc, Hbc = Hf.clone().requires_grad_(True), Hb.clone().requires_grad_(True)
    ff, fb = port.homography_to_flow(Hfc, h, w)[0], port.homography_to_flow(Hbc, h, w)[0]
    mf, mb = port.border_mask(ff).unsqueeze(1), port.border_mask(fb).unsqueeze(1)
    ref = port.masked_l1(mf, i1c, port.get_warp_flow(i2c, ff)) + port.masked_l1(mb, i2c, port.get_warp_flow(i1c, fb))
    ref.backward()
    i1g, i2g = img1.to(DEV).requires_grad_(True), img2.to(DEV).requires_grad_(True)
    Hfg, Hbg = Hf.to(DEV).requires_grad_(True), Hb.to(DEV).requires_grad_(True)
    loss = ops.warp_loss([ops.WarpTerm(i2g, i1g, Hfg), ops.WarpTerm(i1g, i2g, Hbg)], kind=ops.PARAM_HOMOGRAPHY,
                         fused=fused)
    assert abs(loss.item() - ref.item()) < 1e-5
    loss.backward()
    assert (i1g.grad.cpu() - i1c.grad).abs().max().item() < ATOL
    assert (i2g.grad.cpu() - i2c.grad).abs().max().item() < ATOL
    # gradients to H are sums over all pixels of signed terms: yardstick = the same pipeline in fp64
    i1d, i2d = img1.double().requires_grad_(True), img2.double().requires_grad_(True)
    Hfd, Hbd = Hf.double().requires_grad_(True), Hb.double().requires_grad_(True)
    ffd, fbd = port.homography_to_flow(Hfd, h, w)[0], port.homography_to_flow(Hbd, h, w)[0]
    mfd, mbd = port.border_mask(ffd).unsqueeze(1).double(), port.border_mask(fbd).unsqueeze(1).double()
    (port.masked_l1(mfd, i1d, port.get_warp_flow(i2d, ffd)) + port.masked_l1(mbd, i2d, port.get_warp_flow(i1d, fbd))).backward()
    close_to_fp64("dL/dHf", Hfg.grad, Hfc.grad, Hfd.grad)
    close_to_fp64("dL/dHb", Hbg.grad, Hbc.grad, Hbd.grad)


def test_fused_loss_upstream_scaling_and_weight():
    B, C, h, w = 2, 1, 40, 64
    img1, img2, off_f, _ = _pipeline_inputs(B, C, h, w, 5.0, 41, smooth=True)
    src = port.corner_points(B, h, w)
    H = port.dlt4(src, src + off_f)
    outs = []
    for fused in (True, False):
        i2g = img2.to(DEV).requires_grad_(True)
        loss = ops.warp_loss([ops.WarpTerm(i2g, img1.to(DEV), H.to(DEV))], kind=ops.PARAM_HOMOGRAPHY, weight=3.0,
                             fused=fused)
        (loss * 0.25).backward()
        outs.append((loss.item(), i2g.grad.cpu()))
    i2c = img2.clone().requires_grad_(True)
    flow = port.homography_to_flow(H, h, w)[0]
    ref = 3.0 * port.masked_l1(port.border_mask(flow).unsqueeze(1), img1, port.get_warp_flow(i2c, flow))
    (ref * 0.25).backward()
    for val, grad in outs:
        assert abs(val - ref.item()) < 1e-5
        assert (grad - i2c.grad).abs().max().item() < 1e-6


@pytest.mark.parametrize("C", [1, 6])
def test_fused_loss_soft_mask_and_flow_param(C):
    """OSNet's 'unsup' term: soft masks from the mask net, warp by an explicit flow, grads to flow and masks
    (C = 6: feature maps, net.py:817-818 warps img2_fea)."""
    B, h, w = 2, 32, 48
    gen = g(42)
    f1, f2 = synth.smooth_images(B, C, h, w, gen), synth.smooth_images(B, C, h, w, gen)
    flow_f = torch.randn(B, 2, h, w, generator=gen) * 3
    flow_b = torch.randn(B, 2, h, w, generator=gen) * 3
    mf, mb = torch.rand(B, 1, h, w, generator=gen), torch.rand(B, 1, h, w, generator=gen)
    leaves_c = [t.clone().requires_grad_(True) for t in (f1, f2, flow_f, flow_b, mf, mb)]
    a, b, ffc, fbc, mfc, mbc = leaves_c
    ref = 2.0 * (port.masked_l1(mfc, a, port.get_warp_flow(b, ffc)) + port.masked_l1(mbc, b, port.get_warp_flow(a, fbc)))
    ref.backward()
    for fused in (True, False):
        leaves_g = [t.to(DEV).requires_grad_(True) for t in (f1, f2, flow_f, flow_b, mf, mb)]
        ag, bg, ffg, fbg, mfg, mbg = leaves_g
        loss = losses.unsup_loss(ag, bg, ffg, fbg, mask_f=mfg, mask_b=mbg, weight=2.0, fused=fused)
        assert abs(loss.item() - ref.item()) < 1e-5
        loss.backward()
        for tg, tc in zip(leaves_g, leaves_c):
            assert (tg.grad.cpu() - tc.grad).abs().max().item() < ATOL


def test_fused_loss_one_mask_tensor_shared_by_both_terms():
    """params.normalize_mask: mask_b = mask_f = mask_fusion (HEM/loss/losses.py:129) - ONE soft-mask tensor (and here
    also one flow tensor) feeds both terms; its gradient is the sum of both terms' contributions, fused or not."""
    B, C, h, w = 2, 1, 32, 48
    gen = g(44)
    f1, f2 = synth.smooth_images(B, C, h, w, gen), synth.smooth_images(B, C, h, w, gen)
    flow = torch.randn(B, 2, h, w, generator=gen) * 3
    m = torch.rand(B, 1, h, w, generator=gen)
    a, b, fc, mc = [t.clone().requires_grad_(True) for t in (f1, f2, flow, m)]
    ref = port.masked_l1(mc, a, port.get_warp_flow(b, fc)) + port.masked_l1(mc, b, port.get_warp_flow(a, fc))
    ref.backward()
    for fused in (True, False):
        ag, bg, fg, mg = [t.to(DEV).requires_grad_(True) for t in (f1, f2, flow, m)]
        loss = losses.unsup_loss(ag, bg, fg, fg, mask_f=mg, mask_b=mg, fused=fused)
        assert abs(loss.item() - ref.item()) < 1e-5
        loss.backward()
        for tg, tc, name in ((ag, a, "img1"), (bg, b, "img2"), (fg, fc, "flow"), (mg, mc, "mask")):
            assert (tg.grad.cpu() - tc.grad).abs().max().item() < ATOL, (name, fused)


def test_fused_loss_aliasing_leaves_get_their_own_gradients():
    """Two distinct leaves that view one storage are different inputs: each receives its own gradient."""
    B, C, h, w = 2, 1, 32, 48
    gen = g(45)
    base = synth.smooth_images(B, C, h, w, gen).to(DEV)
    other = synth.smooth_images(B, C, h, w, gen).to(DEV)
    src = port.corner_points(B, h, w)
    H = port.dlt4(src, src + synth.corner_offsets(B, 4.0, gen)).to(DEV)
    x1 = base.detach().requires_grad_(True)
    x2 = base.detach().requires_grad_(True)        # same storage, different leaf
    assert x1.data_ptr() == x2.data_ptr()
    loss = ops.warp_loss([ops.WarpTerm(x1, other, H), ops.WarpTerm(other, x2, H)], kind=ops.PARAM_HOMOGRAPHY)
    loss.backward()
    assert x1.grad is not None and x2.grad is not None
    y = base.clone().requires_grad_(True)
    ops.warp_loss([ops.WarpTerm(y, other, H)], kind=ops.PARAM_HOMOGRAPHY).backward()
    assert (x1.grad - y.grad).abs().max().item() < 1e-7      # source-side gradient only
    z = base.clone().requires_grad_(True)
    ops.warp_loss([ops.WarpTerm(other, z, H)], kind=ops.PARAM_HOMOGRAPHY).backward()
    assert (x2.grad - z.grad).abs().max().item() < 1e-7      # target-side gradient only


def test_fused_loss_second_backward_with_retain_graph():
    """retain_graph=True with a non-unit loss scale: the second backward must not rescale (or hand out again) the
    buffers of the first; basis_warp_loss refuses a second backward loudly."""
    B, C, h, w = 2, 1, 40, 64
    img1, img2, off_f, _ = _pipeline_inputs(B, C, h, w, 5.0, 46, smooth=True)
    src = port.corner_points(B, h, w)
    H = port.dlt4(src, src + off_f).to(DEV)
    i2 = img2.to(DEV).requires_grad_(True)
    loss = ops.warp_loss([ops.WarpTerm(i2, img1.to(DEV), H)], kind=ops.PARAM_HOMOGRAPHY, fused=True)
    (loss * 3.0).backward(retain_graph=True)
    g1 = i2.grad.clone()
    i2.grad = None
    (loss * 3.0).backward()
    assert (i2.grad - g1).abs().max().item() < 1e-7
    i2c = img2.clone().requires_grad_(True)
    flow = port.homography_to_flow(H.cpu(), h, w)[0]
    (3.0 * port.masked_l1(port.border_mask(flow).unsqueeze(1), img1, port.get_warp_flow(i2c, flow))).backward()
    assert (g1.cpu() - i2c.grad).abs().max().item() < 1e-6
    # accumulation into .grad across the two backwards (autograd may have adopted the first buffer as .grad)
    i2b = img2.to(DEV).requires_grad_(True)
    loss = ops.warp_loss([ops.WarpTerm(i2b, img1.to(DEV), H)], kind=ops.PARAM_HOMOGRAPHY, fused=True)
    (loss * 3.0).backward(retain_graph=True)
    (loss * 3.0).backward()
    assert (i2b.grad - 2 * g1).abs().max().item() < 1e-6

    basis = hem_utils.gen_basis(h, w).to(DEV)
    wf = synth.basis_weights(B, g(47), 2.0).to(DEV).requires_grad_(True)
    wb = synth.basis_weights(B, g(48), 2.0).to(DEV).requires_grad_(True)
    l2 = ops.basis_warp_loss(basis, img1.to(DEV), img2.to(DEV), wf, wb)
    l2.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="second backward"):
        l2.backward()


def test_dgm_photo_loss():
    B, C, h, w = 3, 3, 32, 32
    gen = g(43)
    im1, im2 = synth.smooth_images(B, C, h, w, gen), synth.smooth_images(B, C, h, w, gen)
    flow = torch.randn(B, 2, h, w, generator=gen) * 3
    mask = (torch.rand(B, 1, h, w, generator=gen) > 0.2).float()
    abar = torch.rand(B, generator=gen)
    a, b = im1.clone().requires_grad_(True), im2.clone().requires_grad_(True)
    ref = port.dgm_photo_loss(a, b, flow, mask, abar)
    ref.backward()
    for fused in (True, False):
        ag, bg = im1.to(DEV).requires_grad_(True), im2.to(DEV).requires_grad_(True)
        loss = dgm.photo_loss(ag, bg, flow.to(DEV), mask.to(DEV), abar.to(DEV), fused=fused)
        assert abs(loss.item() - ref.item()) < 1e-5
        loss.backward()
        assert (ag.grad.cpu() - a.grad).abs().max().item() < ATOL
        assert (bg.grad.cpu() - b.grad).abs().max().item() < ATOL


def test_cfg1_pipeline_chained_smooth():
    """cfg 1 end to end on smooth images (chained: our own DLT feeds our own warp)."""
    c = synth.CONFIGS["cfg1"]
    B, C, h, w = 4, c["C"], c["h"], c["w"]
    img1, img2, off_f, off_b = _pipeline_inputs(B, C, h, w, c["rho"], 230, smooth=True)
    ref = port.pipeline_h4pt(img1, img2, off_f, off_b)
    src = port.corner_points(B, h, w).to(DEV)
    H = ops.dlt4(torch.cat([src, src]), torch.cat([src + off_f.to(DEV), src + off_b.to(DEV)]))
    assert rel_fro(H[:B].cpu(), ref["Hf"]) < 1e-5 and rel_fro(H[B:].cpu(), ref["Hb"]) < 1e-5
    w2, mf = ops.warp(img2.to(DEV), H[:B], kind=ops.PARAM_HOMOGRAPHY, return_mask=True)
    w1, mb = ops.warp(img1.to(DEV), H[B:], kind=ops.PARAM_HOMOGRAPHY, return_mask=True)
    # masks may differ only where the reference's own DLT noise moves a pixel across the border
    assert (mf.cpu() != ref["mf"][:, 0].bool()).float().mean().item() < 1e-4
    assert (w2.cpu() - ref["w2"]).abs().max().item() < ATOL
    assert (w1.cpu() - ref["w1"]).abs().max().item() < ATOL
    loss = ops.warp_loss([ops.WarpTerm(img2.to(DEV), img1.to(DEV), H[:B]),
                          ops.WarpTerm(img1.to(DEV), img2.to(DEV), H[B:])], kind=ops.PARAM_HOMOGRAPHY)
    assert abs(loss.item() - ref["loss"].item()) < ATOL


@pytest.mark.parametrize("variant", ["dlt", "direct"])
def test_cfg2_pipeline(variant):
    """cfg 2: 8 basis weights -> (corner offsets -> DLT -> H | basis flow) -> bidirectional S1 warp ->
    masked L1, forward + backward to the images and the weights."""
    B, C, h, w = 4, 1, 64, 96
    gen = g(231)
    img1, img2 = synth.smooth_images(B, C, h, w, gen), synth.smooth_images(B, C, h, w, gen)
    basis = hem_utils.gen_basis(h, w)
    wf, wb = synth.basis_weights(B, gen, 2.0), synth.basis_weights(B, gen, 2.0)
    leaves_c = [t.clone().requires_grad_(True) for t in (img1, img2, wf, wb)]
    ref = port.pipeline_basis(leaves_c[0], leaves_c[1], basis.reshape(1, 8, -1), leaves_c[2], leaves_c[3],
                              variant=variant, backward=True)
    leaves_g = [t.to(DEV).requires_grad_(True) for t in (img1, img2, wf, wb)]
    i1, i2, wfg, wbg = leaves_g
    bd = basis.to(DEV)
    if variant == "dlt":
        src = port.corner_points(B, h, w).to(DEV)
        Hf = ops.dlt4(src, src + ops.basis_corner_offsets(bd, wfg, h, w))
        Hb = ops.dlt4(src, src + ops.basis_corner_offsets(bd, wbg, h, w))
        loss = ops.warp_loss([ops.WarpTerm(i2, i1, Hf), ops.WarpTerm(i1, i2, Hb)], kind=ops.PARAM_HOMOGRAPHY)
    else:
        loss = ops.warp_loss([ops.WarpTerm(i2, i1, wfg), ops.WarpTerm(i1, i2, wbg)], kind=ops.PARAM_BASIS8, basis=bd)
    assert abs(loss.item() - ref["loss"].item()) < ATOL
    loss.backward()
    for tg, tc in zip(leaves_g[:2], leaves_c[:2]):
        assert (tg.grad.cpu() - tc.grad).abs().max().item() < ATOL
    leaves_d = [t.double().requires_grad_(True) for t in (img1, img2, wf, wb)]
    port.pipeline_basis(leaves_d[0], leaves_d[1], basis.double().reshape(1, 8, -1), leaves_d[2], leaves_d[3], variant=variant,
                        backward=True)
    for name, tg, tc, td in zip(("dL/dw_f", "dL/dw_b"), leaves_g[2:], leaves_c[2:], leaves_d[2:]):
        close_to_fp64(name, tg.grad, tc.grad, td.grad)


def test_basis_homography_fused_matches_unfused():
    """The one-launch cfg-2 prologue (weights -> corner offsets -> DLT, both directions) against the two-step ops
    and the oracle, forward and backward."""
    B, h, w = 5, 64, 96
    basis = hem_utils.gen_basis(h, w)
    gen = g(233)
    wf, wb = synth.basis_weights(B, gen, 2.0), synth.basis_weights(B, gen, 2.0)
    bd = basis.to(DEV)
    src = port.corner_points(B, h, w)
    gH = torch.randn(2, B, 3, 3, generator=gen)
    # oracle
    wfc, wbc = wf.clone().requires_grad_(True), wb.clone().requires_grad_(True)
    Hf_o = port.dlt4(src, src + port.basis_corner_offsets(basis.reshape(1, 8, -1), wfc, h, w))
    Hb_o = port.dlt4(src, src + port.basis_corner_offsets(basis.reshape(1, 8, -1), wbc, h, w))
    ((Hf_o * gH[0]).sum() + (Hb_o * gH[1]).sum()).backward()
    # fused
    wfg, wbg = wf.to(DEV).requires_grad_(True), wb.to(DEV).requires_grad_(True)
    Hf, Hb = ops.basis_homography(bd, h, w, wfg, wbg)
    assert rel_fro(Hf.detach().cpu(), Hf_o.detach()) < 1e-5 and rel_fro(Hb.detach().cpu(), Hb_o.detach()) < 1e-5
    ((Hf * gH[0].to(DEV)).sum() + (Hb * gH[1].to(DEV)).sum()).backward()
    # unfused CUDA path: bit-identical forward (same arithmetic, one launch instead of three)
    sd = src.to(DEV)
    Hf2 = ops.dlt4(sd, sd + ops.basis_corner_offsets(bd, wf.to(DEV), h, w))
    assert torch.equal(Hf.detach(), Hf2)
    wfd, wbd = wf.double().requires_grad_(True), wb.double().requires_grad_(True)
    b64, s64 = basis.double().reshape(1, 8, -1), src.double()
    ((port.dlt4(s64, s64 + port.basis_corner_offsets(b64, wfd, h, w)) * gH[0].double()).sum() +
     (port.dlt4(s64, s64 + port.basis_corner_offsets(b64, wbd, h, w)) * gH[1].double()).sum()).backward()
    for name, tg, tc, td in (("dL/dw_f", wfg, wfc, wfd), ("dL/dw_b", wbg, wbc, wbd)):
        # relative yardstick: these gradients are O(1e2..1e4) (random dL/dH through an 8x8 solve of condition ~1e6)
        scale = max(1.0, td.grad.abs().max().item())
        close_to_fp64(name, tg.grad / scale, tc.grad / scale, td.grad / scale)


# ---------------------------------------------------------------------------------- DGM rendering
def test_flow_to_rgb():
    flow = (np.random.default_rng(50).normal(size=(2, 24, 32, 2)) * 20).astype(np.float32)
    flow[0, 0, 0] = 0
    for i in range(2):
        ref = port.flow_to_image(flow[i])
        out = dgm.flow_to_image(flow[i])
        assert np.abs(out - ref).max() < 1e-5
    t = torch.from_numpy(flow.transpose(0, 3, 1, 2).copy())
    assert (dgm.visulize_flow(t.to(DEV)).cpu() - port.visualize_flow(t)).abs().max().item() < 1e-5


def test_warp_perspective_matches_cv2():
    rng = np.random.default_rng(51)
    B = 3
    imgs = rng.random((B, 64, 80, 3), dtype=np.float32)
    Hs = np.stack([np.eye(3) + rng.normal(size=(3, 3)) * np.array([[2e-2, 2e-2, 4], [2e-2, 2e-2, 4], [1e-4, 1e-4, 0]])
                   for _ in range(B)])
    out = ops.warp_perspective(torch.from_numpy(imgs).to(DEV), torch.from_numpy(Hs).to(DEV), (80, 64),
                               channels_last=True).cpu().numpy()
    out_nchw = ops.warp_perspective(torch.from_numpy(imgs.transpose(0, 3, 1, 2).copy()).to(DEV),
                                    torch.from_numpy(Hs).to(DEV), (72, 48)).cpu().numpy()
    for i in range(B):
        assert np.abs(out[i] - port.warp_perspective_cv2(imgs[i], Hs[i], (80, 64))).max() < 1e-5
        assert np.abs(out_nchw[i].transpose(1, 2, 0) - port.warp_perspective_cv2(imgs[i], Hs[i], (72, 48))).max() < 1e-5


def test_cfg3_dgm_render():
    c = synth.CONFIGS["cfg3"]
    B, C, h, w = 5, c["C"], c["h"], c["w"]
    gen = g(232)
    im2 = synth.noise_images(B, C, h, w, gen)
    homos = synth.homographies_360x640(B, gen, c["rho"])
    ref = port.pipeline_dgm_render(im2, homos)
    Hs = np.stack([dgm.adapt_homography_to_preprocessing_v3(360, 640, homos[i], h, w) for i in range(B)])
    Ht = torch.as_tensor(Hs, device=DEV)
    persp = ops.warp_perspective(im2.to(DEV), Ht, (w, h)).cpu()
    flow = ops.homography_to_flow_f64(Ht, h, w, channels_last=False)
    rgb = ops.flow_to_rgb(flow).cpu()
    fw = dgm.flow_warp(im2.to(DEV), flow).cpu()
    assert (persp - ref["persp"]).abs().max().item() < 1e-5
    assert torch.equal(flow.cpu(), ref["flow"])
    assert (rgb - ref["rgb"]).abs().max().item() < 1e-5
    assert (fw - ref["flow_warp"]).abs().max().item() < ATOL


def test_render_conditions_one_call_equals_the_separate_calls():
    """ops.render_conditions (forked stream branches) against the four separate calls, eagerly and replayed from a CUDA
    graph: bit-identical outputs."""
    B, h, w = 5, 64, 96
    gen = g(310)
    im2 = torch.rand(B, 3, h, w, generator=gen).to(DEV)
    H360 = synth.homographies_360x640(B, gen, 24.0)
    homo = torch.stack([torch.from_numpy(dgm.adapt_homography_to_preprocessing_v3(360, 640, H360[b], h, w)) for b in range(B)]).to(DEV)
    ref_warp = ops.warp_perspective(im2, homo, (w, h))
    ref_flow = ops.homography_to_flow_f64(homo, h, w, eps=1e-6, channels_last=False)
    ref_rgb = ops.flow_to_rgb(ref_flow, 256.0)
    ref_fw = dgm.flow_warp(im2, ref_flow)
    out = dgm.render_conditions(im2, homo)
    torch.cuda.synchronize()
    for key, ref in (("warp", ref_warp), ("flow", ref_flow), ("flow_rgb", ref_rgb), ("flow_warp", ref_fw)):
        assert torch.equal(out[key], ref), key
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        dgm.render_conditions(im2, homo)
        stream.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=stream):
            captured = dgm.render_conditions(im2, homo)
        for t in captured.values():
            t.zero_()
        gr.replay()
        stream.synchronize()
    for key, ref in (("warp", ref_warp), ("flow", ref_flow), ("flow_rgb", ref_rgb), ("flow_warp", ref_fw)):
        assert torch.equal(captured[key], ref), key


def test_post_process():
    B, h = 2, 256
    gen = g(52)
    t = torch.rand(B, 6, h, h, generator=gen)
    mask = (torch.rand(B, 1, h, h, generator=gen) > 0.5).float()
    flows = torch.randn(B, 2, h, h, generator=gen) * 4
    b1, b2 = dgm.postProcess(t.to(DEV), mask.to(DEV), flows.to(DEV))
    assert b1.shape == (B, 3, h, 4 * h) and b2.shape == (B, 3, h, 4 * h)
    assert (b2[:, :, :, h:2 * h].cpu() - port.flow_warp(t[:, 3:6], flows)).abs().max().item() < ATOL
    imgs = (t.numpy() * 255).astype(np.uint8)
    homos = synth.homographies_360x640(B, gen, 8.0)
    homos = np.stack([port.homo_scale(360, 640, homos[i], h, h) for i in range(B)])
    c1, c2 = dgm.postProcess_cv2(imgs, homos, 0)
    for i in range(B):
        img1 = (imgs[i, :3].astype(np.float32) / 255.0).transpose(1, 2, 0)
        ref = port.warp_perspective_cv2(np.ascontiguousarray(img1), homos[i], (256, 256)).transpose(2, 0, 1)
        assert np.abs(c1[i, :, :, h:].cpu().numpy() - ref).max() < 1e-5


# ---------------------------------------------------------------------------------- eval metric, LS homography
def test_eval_point_error():
    gen = g(53)
    B, h, w = 3, 20, 30
    ff, fb = torch.randn(B, h, w, 2, generator=gen), torch.randn(B, h, w, 2, generator=gen)
    pts = torch.rand(B, 6, 2, 2, generator=gen) * torch.tensor([w - 1.0, h - 1.0])
    ref = torch.stack(port.eval_point_errors(pts, ff, fb))
    out = torch.stack(losses.compute_eval_results({"pt_set": pts.to(DEV)}, {"flow_f": ff.to(DEV), "flow_b": fb.to(DEV)}))
    assert (out.cpu() - ref).abs().max().item() < 1e-5
    e = losses.ComputeErrFlow(pts[0, 0, 0].to(DEV), pts[0, 0, 1].to(DEV), ff[0].to(DEV))
    assert abs(e.item() - port.err_flow(pts[0, 0, 0], pts[0, 0, 1], ff[0]).item()) < 1e-5


def test_flow_to_homography_ls():
    B, h, w = 3, 64, 64
    src = port.corner_points(B, h, w)
    H = port.dlt4(src, src + synth.corner_offsets(B, 6.0, g(54)))
    flow = port.homography_to_flow(H, h, w)[0] + torch.randn(B, 2, h, w, generator=g(55)) * 0.05
    ref = port.homo_gen(flow)
    out = dgm.homo_gen(flow.to(DEV)).cpu()
    assert out.shape == ref.shape and out.dtype == torch.float64
    for b in range(B):
        assert rel_fro(out[b], ref[b]) < 1e-8


# ---------------------------------------------------------------------------------- error behaviour
def test_no_cpu_fallback():
    with pytest.raises(RuntimeError):
        ops.warp(torch.zeros(1, 1, 4, 4), torch.zeros(1, 2, 4, 4))
    with pytest.raises(RuntimeError):
        ops.dlt4(torch.zeros(1, 4, 2), torch.zeros(1, 4, 2))


def test_basis_warp_loss_one_op_matches_composition_and_graph_replay():
    """ops.basis_warp_loss (one op, forked stream branches) == basis_homography + warp_loss (same kernels): loss,
    homographies and every gradient; also against the oracle, and replayed from a CUDA graph (the form bench.py runs)."""
    B, h, w = 3, 64, 96
    gen = synth.generator()
    img1, img2 = synth.smooth_images(B, 1, h, w, gen), synth.smooth_images(B, 1, h, w, gen)
    basis = hem_utils.gen_basis(h, w)
    wf, wb = synth.basis_weights(B, gen, 2.0), synth.basis_weights(B, gen, 2.0)
    bd = basis.to(DEV)

    def leaves():
        return [t.clone().to(DEV).requires_grad_(True) for t in (img1, img2, wf, wb)]

    a = leaves()
    Hf, Hb = ops.basis_homography(bd, h, w, a[2], a[3])
    la = ops.warp_loss([ops.WarpTerm(a[1], a[0], Hf), ops.WarpTerm(a[0], a[1], Hb)], kind=ops.PARAM_HOMOGRAPHY)
    la.backward()
    b = leaves()
    lb, Hf2, Hb2 = ops.basis_warp_loss(bd, b[0], b[1], b[2], b[3], return_homographies=True)
    lb.backward()
    assert torch.equal(Hf2, Hf.detach()) and torch.equal(Hb2, Hb.detach())
    assert abs(la.item() - lb.item()) < 1e-7
    for x, y in zip(a, b):
        assert (x.grad - y.grad).abs().max().item() <= 1e-6 * max(1.0, x.grad.abs().max().item())

    lc = [t.clone().requires_grad_(True) for t in (img1, img2, wf, wb)]
    ref = port.pipeline_basis(lc[0], lc[1], basis.reshape(1, 8, -1), lc[2], lc[3], variant="dlt", backward=True)
    assert abs(lb.item() - ref["loss"].item()) < 1e-4
    for x, y in zip(b[:2], lc[:2]):
        assert (x.grad.cpu() - y.grad).abs().max().item() < 1e-4

    # CUDA-graph capture of the whole step (forked branches become parallel nodes), replayed twice
    c = leaves()
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        for _ in range(2):
            for t in c:
                t.grad = None
            ops.basis_warp_loss(bd, c[0], c[1], c[2], c[3]).backward()
        stream.synchronize()
        for t in c:
            t.grad = None
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            lg = ops.basis_warp_loss(bd, c[0], c[1], c[2], c[3])
            lg.backward()
        for _ in range(2):
            g.replay()
        stream.synchronize()
    assert abs(lg.item() - lb.item()) < 1e-7
    for x, y in zip(c, b):
        assert (x.grad - y.grad).abs().max().item() <= 1e-6 * max(1.0, y.grad.abs().max().item())


@pytest.mark.parametrize("h,w", [(5, 7), (8, 12), (3, 4), (17, 33), (2, 64)])
def test_elementwise_kernels_scalar_and_vector_paths(h, w):
    """The one-pass kernels take four pixels per thread with 128-bit accesses when w % 4 == 0 and one otherwise;
    both paths, on tiny and ragged planes, bit-exact against the oracle (get_flow, the fp64 flows / mappings in both
    layouts, the M1 mask as bool and as float)."""
    B = 3
    src = port.corner_points(B, max(h, 4), max(w, 4))
    H = port.dlt4(src, src + synth.corner_offsets(B, 1.5, g(91)))
    for start in (0, 2):
        assert torch.equal(ops.homography_to_flow(H.to(DEV), h, w, start=start).cpu(), port.homography_to_flow(H, h, w, start=start)[0])
    H64 = H.double().numpy()
    f64 = ops.homography_to_flow_f64(torch.from_numpy(H64).to(DEV), h, w).cpu()                      # (B,h,w,2), eps 1e-6
    fcf = ops.homography_to_flow_f64(torch.from_numpy(H64).to(DEV), h, w, channels_last=False).cpu()
    gt = ops.homography_to_flow_f64(torch.from_numpy(H64).to(DEV), h, w, eps=1e-8, channels_last=False, as_mapping=2).cpu()
    mp = ops.homography_to_flow_f64(torch.from_numpy(H64).to(DEV), h, w, eps=1e-8, channels_last=False, as_mapping=True).cpu()
    for b in range(B):
        ref = torch.from_numpy(port.homo_to_flow_np(H64[b], h, w))
        assert torch.equal(f64[b], ref) and torch.equal(fcf[b], ref.permute(2, 0, 1))
        assert torch.equal(gt[b], port.homo_convert_to_flow(H64[b], (h, w))[0])
        mx, my = port.homography_to_mapping_np((h, w), H64[b])
        assert torch.equal(mp[b, 0], torch.from_numpy(mx)) and torch.equal(mp[b, 1], torch.from_numpy(my))
    flow = torch.randn(B, 2, h, w, generator=g(92)) * (w / 2)
    assert torch.equal(ops.border_mask(flow.to(DEV)).cpu(), port.correspondence_mask(flow))
    assert torch.equal(ops.border_mask(flow.to(DEV), as_float=True).cpu(), port.border_mask(flow).reshape(B, h, w))


@pytest.mark.parametrize("B,h,w", [(11, 64, 96), (64, 320, 576), (3, 50, 70), (5, 45, 71)])
def test_basis_combine_backward_groups(B, h, w):
    """flow = sum_k w_k basis_k (HEM/model/net.py:808-815) and dL/dweights: four pixels per thread when the plane allows
    (45 x 71 does not: scalar kernels), sample groups of 8 per CTA with a ragged last group - the forward bit-exact, the
    backward against an fp64 evaluation."""
    gen = g(260)
    basis = hem_utils.gen_basis(h, w)
    wt = synth.basis_weights(B, gen, 3.0)
    gflow = torch.randn(B, 2, h, w, generator=gen)
    w64 = wt.double().requires_grad_(True)
    (port.basis_combine(basis.double().reshape(1, 8, -1), w64, h, w) * gflow.double()).sum().backward()
    w32 = wt.clone().requires_grad_(True)
    (port.basis_combine(basis.reshape(1, 8, -1), w32, h, w) * gflow).sum().backward()
    wg = wt.to(DEV).requires_grad_(True)
    flow = ops.basis_combine(basis.to(DEV), wg, h, w)
    assert torch.equal(flow.detach().cpu(), port.basis_combine_sequential(basis.reshape(1, 8, -1), wt, h, w))
    if (2 * h * w) % 32 == 0:   # torch's CPU sum associates the tail elements of a row differently (oracle/port.py)
        assert torch.equal(flow.detach().cpu(), port.basis_combine(basis.reshape(1, 8, -1), wt, h, w))
    (flow * gflow.to(DEV)).sum().backward()
    scale = max(1.0, w64.grad.abs().max().item())
    close_to_fp64("dL/dw", wg.grad.reshape(B, 8) / scale, w32.grad.reshape(B, 8) / scale, w64.grad.reshape(B, 8) / scale)
