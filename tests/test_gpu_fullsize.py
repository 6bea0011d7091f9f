"""Full-size property tests (BASELINE.json sizes: cfg2 = 64 pairs 1x320x576, a cfg4-shaped 3x512x512 batch): an
integer translation is a homography whose every coordinate, weight and tap is exact, so the fused kernel's outputs
have a closed form that plain torch ops on the GPU can state at any size - no oracle run needed:

    warp(img, T)[y, x] = img[y + ty, x + tx]   if 0 <= x + tx < W - 1 and 0 <= y + ty < H - 1, else 0
                         (S1 clamps both taps to the border, where its weights cancel: utils.py:463-523)
    M1 mask            = 0 <= x + tx <= W and 0 <= y + ty <= H                      (flow_and_mapping_operations.py:66-69)
    loss               = mean |m * target - m * warp|, its gradients by autograd through the same closed form.
"""
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
