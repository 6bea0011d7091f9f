"""Full-size property tests (BASELINE.json sizes: cfg2 = 64 pairs 1x320x576, a cfg4-shaped 3x512x512 batch): an
integer translation is a homography whose every coordinate, weight and tap is exact, so the fused kernel's outputs
have a closed form that plain torch ops on the GPU can state at any size - no oracle run needed:

    warp(img, T)[y, x] = img[y + ty, x + tx]   if 0 <= x + tx < W - 1 and 0 <= y + ty < H - 1, else 0
                         (S1 clamps both taps to the border, where its weights cancel: utils.py:463-523)
    M1 mask            = 0 <= x + tx <= W and 0 <= y + ty <= H                      (flow_and_mapping_operations.py:66-69)
    loss               = mean |m * target - m * warp|, its gradients by autograd through the same closed form.
"""
import os

import pytest
import torch

from dmhomo_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def closed_form(img, tx, ty):
    B, C, H, W = img.shape
    ys = torch.arange(H, device=img.device).view(H, 1) + ty
    xs = torch.arange(W, device=img.device).view(1, W) + tx
    inside = (xs >= 0) & (xs < W - 1) & (ys >= 0) & (ys < H - 1)
    mask = (xs >= 0) & (xs <= W) & (ys >= 0) & (ys <= H)
    gathered = img[:, :, ys.clamp(0, H - 1).expand(H, W), xs.clamp(0, W - 1).expand(H, W)]
    return gathered * inside.to(img.dtype), mask.expand(B, H, W)


@pytest.mark.parametrize("B,C,h,w,tx,ty", [(64, 1, 320, 576, 5, -3), (64, 1, 320, 576, -17, 40), (8, 3, 512, 512, 9, 2)])
def test_integer_translation_full_size(B, C, h, w, tx, ty):
    gen = torch.Generator(device=DEV).manual_seed(230)
    img1 = torch.rand(B, C, h, w, generator=gen, device=DEV)
    img2 = torch.rand(B, C, h, w, generator=gen, device=DEV)
    H = torch.eye(3, device=DEV).repeat(B, 1, 1)
    H[:, 0, 2], H[:, 1, 2] = float(tx), float(ty)
    out, mask = ops.warp(img2, H, kind=ops.PARAM_HOMOGRAPHY, return_mask=True)
    ref, mref = closed_form(img2, tx, ty)
    assert torch.equal(mask, mref)
    assert torch.equal(out, ref)

    # fused loss + gradients, both directions (the second with the inverse shift)
    Hb = torch.eye(3, device=DEV).repeat(B, 1, 1)
    Hb[:, 0, 2], Hb[:, 1, 2] = float(-tx), float(-ty)
    i1, i2 = img1.clone().requires_grad_(True), img2.clone().requires_grad_(True)
    loss = ops.warp_loss([ops.WarpTerm(i2, i1, H), ops.WarpTerm(i1, i2, Hb)], kind=ops.PARAM_HOMOGRAPHY)
    loss.backward()
    r1, r2 = img1.clone().double().requires_grad_(True), img2.clone().double().requires_grad_(True)
    w2, mf = closed_form(r2, tx, ty)
    w1, mb = closed_form(r1, -tx, -ty)
    mf, mb = mf.unsqueeze(1).double(), mb.unsqueeze(1).double()
    lref = (mf * r1 - mf * w2).abs().mean() + (mb * r2 - mb * w1).abs().mean()
    lref.backward()
    assert abs(loss.item() - lref.item()) < 1e-6
    assert (i1.grad.double() - r1.grad).abs().max().item() < 1e-9 + 1e-6 * r1.grad.abs().max().item()
    assert (i2.grad.double() - r2.grad).abs().max().item() < 1e-9 + 1e-6 * r2.grad.abs().max().item()


# ------------------------------------------------------------------------------------------------------------------
# Full-size parity against the CPU oracle with random projective homographies (cfg2 whole,
# a 64-pair subset of cfg4, an 8-frame subset of cfg5).  Stage-isolated: the oracle and the kernels get the same fp32 H,
# so flows, coordinates, integer sample indices, masks and S1 pixels must be bit-identical; loss and gradients are
# sums in a different order and are held to north_star's 1e-4.  The gradients to H (and to the basis weights) are sums
# of millions of signed terms: both the CUDA result and the oracle's fp32 autograd result are compared with an fp64
# evaluation of the same pipeline, and the kernel has to be at least as close to it as the reference's own arithmetic
# is (or within 1e-4).
# ------------------------------------------------------------------------------------------------------------------
from dmhomo_b200 import _lib, synth  # noqa: E402
from dmhomo_b200.compat import hem_utils  # noqa: E402
from oracle import port  # noqa: E402

ATOL = 1e-4


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def _random_h(B, h, w, rho, seed):
    src = port.corner_points(B, h, w)
    return port.dlt4(src, src + synth.corner_offsets(B, rho, _gen(seed)))


def _oracle_terms(img1, img2, Hf, Hb, dtype=torch.float32):
    """Loss and gradients of the bidirectional term through the oracle, in `dtype` (fp64: the yardstick)."""
    i1, i2 = img1.detach().clone().to(dtype).requires_grad_(True), img2.detach().clone().to(dtype).requires_grad_(True)
    hf, hb = Hf.detach().clone().to(dtype).requires_grad_(True), Hb.detach().clone().to(dtype).requires_grad_(True)
    h, w = img1.shape[-2:]
    ff, fb = port.homography_to_flow(hf, h, w)[0], port.homography_to_flow(hb, h, w)[0]
    mf, mb = port.border_mask(ff).unsqueeze(1).to(dtype), port.border_mask(fb).unsqueeze(1).to(dtype)
    loss = port.masked_l1(mf, i1, port.get_warp_flow(i2, ff)) + port.masked_l1(mb, i2, port.get_warp_flow(i1, fb))
    loss.backward()
    return loss.detach(), i1.grad, i2.grad, hf.grad, hb.grad


def _check_against_fp64(name, cuda, ref32, ref64):
    """|cuda - fp64| <= max(1e-4, |oracle_fp32 - fp64|), elementwise maxima."""
    e_cuda = (cuda.double().cpu() - ref64).abs().max().item()
    e_ref = (ref32.double() - ref64).abs().max().item()
    if os.environ.get("DMH_TEST_REPORT"):
        print(f"[margin] {name}: cuda {e_cuda:.3e} oracle {e_ref:.3e} bound {max(ATOL, 1.05 * e_ref):.3e}")
    assert e_cuda <= max(ATOL, 1.05 * e_ref), f"{name}: |cuda - fp64| = {e_cuda:.3e}, |oracle fp32 - fp64| = {e_ref:.3e}"


def _forward_parity(img, H, tile):
    """warp + mask + integer corner indices against the oracle, bit for bit (given the same fp32 H)."""
    B, C, h, w = img.shape
    _lib.set_tuning(tile=tile)
    try:
        out, mask = ops.warp(img.to(DEV), H.to(DEV), kind=ops.PARAM_HOMOGRAPHY, return_mask=True)
        kernel = ops.last_warp_kernel
    finally:
        _lib.set_tuning(tile=3)
    flow, _ = port.homography_to_flow(H, h, w)
    ref, idx_ref = port.get_warp_flow(img, flow, return_indices=True)
    assert torch.equal(mask.cpu(), port.correspondence_mask(flow)), "validity mask differs"
    assert torch.equal(out.cpu(), ref), "warped pixels differ"
    return kernel, flow, idx_ref


def test_cfg2_full_batch_random_homographies_against_oracle():
    """cfg2 whole: B = 64 pairs 1x320x576, rho = 32."""
    c = synth.CONFIGS["cfg2"]
    B, C, h, w = c["B"], c["C"], c["h"], c["w"]
    gen = _gen(2301)
    img1, img2 = synth.noise_images(B, C, h, w, gen), synth.noise_images(B, C, h, w, gen)
    Hf, Hb = _random_h(B, h, w, c["rho"], 2302), _random_h(B, h, w, c["rho"], 2303)
    kernel, flow, idx_ref = _forward_parity(img2, Hf, tile=3)
    assert "tile" in kernel
    _, _, idx = ops.warp(img2[:8].to(DEV), Hf[:8].to(DEV), kind=ops.PARAM_HOMOGRAPHY, return_mask=True, return_indices=True)
    assert torch.equal(idx.cpu(), idx_ref[:, :8]), "integer sample indices differ"

    l32, g1_32, g2_32, ghf_32, ghb_32 = _oracle_terms(img1, img2, Hf, Hb)
    l64, g1_64, g2_64, ghf_64, ghb_64 = _oracle_terms(img1, img2, Hf, Hb, torch.float64)
    for fused in (True, False):
        i1, i2 = img1.to(DEV).requires_grad_(True), img2.to(DEV).requires_grad_(True)
        hf, hb = Hf.to(DEV).requires_grad_(True), Hb.to(DEV).requires_grad_(True)
        loss = ops.warp_loss([ops.WarpTerm(i2, i1, hf), ops.WarpTerm(i1, i2, hb)], kind=ops.PARAM_HOMOGRAPHY, fused=fused)
        loss.backward()
        assert abs(loss.item() - l32.item()) < 1e-5
        assert (i1.grad.cpu() - g1_32).abs().max().item() < ATOL
        assert (i2.grad.cpu() - g2_32).abs().max().item() < ATOL
        _check_against_fp64(f"dL/dHf fused={fused}", hf.grad, ghf_32, ghf_64)
        _check_against_fp64(f"dL/dHb fused={fused}", hb.grad, ghb_32, ghb_64)


@pytest.mark.parametrize("tile", [1, 2, 3])
def test_cfg4_subset_random_homographies_against_oracle(tile):
    """A cfg4 subset (3x512x512 pairs, rho = 32) on the scalar kernels (tile = 1), on the tile kernel for the
    gradient-free launches only (2) and on the tile kernel throughout (3, the default)."""
    c = synth.CONFIGS["cfg4"]
    B, C, h, w = (64 if tile == 3 else 16), c["C"], c["h"], c["w"]   # 64 pairs on the default dispatch
    gen = _gen(2304)
    img1, img2 = synth.noise_images(B, C, h, w, gen), synth.noise_images(B, C, h, w, gen)
    Hf, Hb = _random_h(B, h, w, c["rho"], 2305), _random_h(B, h, w, c["rho"], 2306)
    kernel, _, _ = _forward_parity(img2, Hf, tile=tile)
    assert ("tile" in kernel) == (tile >= 2)
    l32, g1_32, g2_32, ghf_32, ghb_32 = _oracle_terms(img1, img2, Hf, Hb)
    l64, _, _, ghf_64, ghb_64 = _oracle_terms(img1, img2, Hf, Hb, torch.float64)
    _lib.set_tuning(tile=tile)
    try:
        i1, i2 = img1.to(DEV).requires_grad_(True), img2.to(DEV).requires_grad_(True)
        hf, hb = Hf.to(DEV).requires_grad_(True), Hb.to(DEV).requires_grad_(True)
        loss = ops.warp_loss([ops.WarpTerm(i2, i1, hf), ops.WarpTerm(i1, i2, hb)], kind=ops.PARAM_HOMOGRAPHY)
        assert ("tile" in ops.last_warp_kernel) == (tile >= 3)
        loss.backward()
    finally:
        _lib.set_tuning(tile=3)
    assert abs(loss.item() - l32.item()) < 1e-5
    assert (i1.grad.cpu() - g1_32).abs().max().item() < ATOL
    assert (i2.grad.cpu() - g2_32).abs().max().item() < ATOL
    _check_against_fp64("dL/dHf", hf.grad, ghf_32, ghf_64)
    _check_against_fp64("dL/dHb", hb.grad, ghb_32, ghb_64)


@pytest.mark.parametrize("tile", [1, 2])
def test_cfg5_subset_frames_against_oracle(tile):
    """An 8-frame subset of cfg5 (3x1080x1920, rho = 64): forward warp + mask + integer indices, bit-exact."""
    c = synth.CONFIGS["cfg5"]
    B, C, h, w = 8, c["C"], c["h"], c["w"]
    frames = synth.noise_images(B, C, h, w, _gen(2307))
    H = _random_h(B, h, w, c["rho"], 2308)
    kernel, _, idx_ref = _forward_parity(frames, H, tile=tile)
    assert ("tile" in kernel) == (tile >= 2)
    _, _, idx = ops.warp(frames[:2].to(DEV), H[:2].to(DEV), kind=ops.PARAM_HOMOGRAPHY, return_mask=True, return_indices=True)
    assert torch.equal(idx.cpu(), idx_ref[:, :2]), "integer sample indices differ"


def test_cfg2_basis_weight_gradients_against_fp64():
    """cfg2 chained (8 basis weights -> corner offsets -> DLT -> warp -> loss): dL/dweights against an fp64 evaluation
    of the oracle pipeline - the CUDA result is as close to it as the reference's own fp32 autograd (or within 1e-4)."""
    c = synth.CONFIGS["cfg2"]
    B, C, h, w = 16, c["C"], c["h"], c["w"]
    gen = _gen(2309)
    img1, img2 = synth.smooth_images(B, C, h, w, gen), synth.smooth_images(B, C, h, w, gen)
    wf, wb = synth.basis_weights(B, gen), synth.basis_weights(B, gen)
    basis = hem_utils.gen_basis(h, w)

    def oracle(dtype):
        t = [x.detach().clone().to(dtype).requires_grad_(True) for x in (img1, img2, wf, wb)]
        r = port.pipeline_basis(t[0], t[1], basis.to(dtype).reshape(1, 8, -1), t[2], t[3], variant="dlt", backward=True)
        return r["loss"].detach(), [x.grad for x in t]

    l32, g32 = oracle(torch.float32)
    l64, g64 = oracle(torch.float64)
    lg = [x.to(DEV).requires_grad_(True) for x in (img1, img2, wf, wb)]
    loss = ops.basis_warp_loss(basis.to(DEV), *lg)
    loss.backward()
    assert abs(loss.item() - l64.item()) < ATOL
    for name, tg, a32, a64 in zip(("dL/dimg1", "dL/dimg2", "dL/dw_f", "dL/dw_b"), lg, g32, g64):
        _check_against_fp64(name, tg.grad, a32, a64)


@pytest.mark.parametrize("dyn", [0, 100])
def test_tile_schedule_share_does_not_change_results(dyn):
    """The dynamic tail of the tile schedule only changes who processes a tile: forward outputs are bit-identical, loss
    and gradients agree to summation order, whatever the share (0 = static split only, 100 = every tile claimed; the default depends on the launch kind)."""
    B, C, h, w = 24, 1, 320, 576
    gen = _gen(2310)
    img1, img2 = synth.noise_images(B, C, h, w, gen).to(DEV), synth.noise_images(B, C, h, w, gen).to(DEV)
    Hf, Hb = _random_h(B, h, w, 32.0, 2311).to(DEV), _random_h(B, h, w, 32.0, 2312).to(DEV)

    def run():
        out, mask = ops.warp(img2, Hf, kind=ops.PARAM_HOMOGRAPHY, return_mask=True)
        i1, i2 = img1.clone().requires_grad_(True), img2.clone().requires_grad_(True)
        hf = Hf.clone().requires_grad_(True)
        loss = ops.warp_loss([ops.WarpTerm(i2, i1, hf), ops.WarpTerm(i1, i2, Hb)], kind=ops.PARAM_HOMOGRAPHY)
        loss.backward()
        return out, mask, loss.detach(), i1.grad, i2.grad, hf.grad

    base = run()
    _lib.set_tuning(tile_dyn=dyn)
    try:
        other = run()
    finally:
        _lib.set_tuning(tile_dyn=-1)
    assert torch.equal(base[0], other[0]) and torch.equal(base[1], other[1])
    assert abs(base[2].item() - other[2].item()) < 1e-6
    assert (base[3] - other[3]).abs().max().item() < 1e-7 and (base[4] - other[4]).abs().max().item() < 1e-7
    assert ((base[5] - other[5]).norm() / base[5].norm()).item() < 1e-4


def test_pair_major_order_does_not_change_results():
    """Two-term launches walk the tile list (sample, term, tile) at C = 3 - both directions of a pair back to back, so the
    second use of every image / gradient line hits the L2 - or (term, sample, tile): same tiles, same results."""
    B, C, h, w = 6, 3, 128, 192
    gen = _gen(2313)
    img1, img2 = synth.noise_images(B, C, h, w, gen).to(DEV), synth.noise_images(B, C, h, w, gen).to(DEV)
    Hf, Hb = _random_h(B, h, w, 12.0, 2314).to(DEV), _random_h(B, h, w, 12.0, 2315).to(DEV)

    def run():
        loss_e, outs, masks = ops.warp_eval([ops.WarpTerm(img2, img1, Hf), ops.WarpTerm(img1, img2, Hb)], kind=ops.PARAM_HOMOGRAPHY)
        i1, i2 = img1.clone().requires_grad_(True), img2.clone().requires_grad_(True)
        hf, hb = Hf.clone().requires_grad_(True), Hb.clone().requires_grad_(True)
        loss = ops.warp_loss([ops.WarpTerm(i2, i1, hf), ops.WarpTerm(i1, i2, hb)], kind=ops.PARAM_HOMOGRAPHY)
        assert "tile" in ops.last_warp_kernel
        loss.backward()
        return loss_e, outs, masks, loss.detach(), i1.grad, i2.grad, hf.grad, hb.grad

    res = {}
    for pm in (1, 0):
        _lib.set_tuning(tile_pair_major=pm)
        try:
            res[pm] = run()
        finally:
            _lib.set_tuning(tile_pair_major=-1)
    a, b = res[1], res[0]
    assert torch.equal(a[1][0], b[1][0]) and torch.equal(a[1][1], b[1][1]) and torch.equal(a[2][0], b[2][0])
    assert abs(a[0].item() - b[0].item()) < 1e-6 and abs(a[3].item() - b[3].item()) < 1e-6
    assert (a[4] - b[4]).abs().max().item() < 1e-6 and (a[5] - b[5]).abs().max().item() < 1e-6
    assert ((a[6] - b[6]).norm() / b[6].norm()).item() < 1e-4 and ((a[7] - b[7]).norm() / b[7].norm()).item() < 1e-4
