"""GPU parity of the rows either side of the path: the uint8 pair format,
the loaders' ground-truth flow and the flow upsample + pyramid-level feature warp - against the committed
fixtures of the REAL reference (tests/golden) and against the CPU oracle on other shapes.
Bars: integer / byte / fp64-derived values bit-exact; interpolated fp32 values within 1e-5 absolute."""
import os

import numpy as np
import pytest
import torch

from dmhomo_b200 import ops
from dmhomo_b200.compat import data_loader as cdl, hem_utils
from oracle import port

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def g(seed):
    return torch.Generator().manual_seed(seed)


def test_pairs_u8_golden_bit_exact():
    d = load("pairs_u8")
    crop = tuple(int(v) for v in d["crop"])
    out = cdl.pairs_to_batch(d["img12"], d["start"], crop)
    assert torch.equal(out["imgs_gray_full"].cpu(), d["gray_full"])
    assert torch.equal(out["imgs_gray_patch"].cpu(), d["gray_patch"])
    assert torch.equal(out["imgs_rgb_full"].cpu(), d["rgb_full"])
    assert out["start"].shape == (3, 2, 1, 1)


@pytest.mark.parametrize("B,H,W,ph,pw", [(2, 360, 640, 320, 576), (5, 64, 100, 17, 33), (1, 8, 4, 8, 4)])
def test_pairs_u8_matches_oracle(B, H, W, ph, pw):
    rs = np.random.default_rng(7)
    img12 = rs.integers(0, 256, size=(B, 6, H, W), dtype=np.uint8)
    start = np.stack([rs.integers(0, W - pw + 1, size=B), rs.integers(0, H - ph + 1, size=B)], 1).astype(np.int32)
    full, patch, rgb = ops.pairs_u8_to_gray(torch.from_numpy(img12).to(DEV), start=torch.from_numpy(start), patch_size=(ph, pw))
    for b in range(B):
        f, p, r = port.pairs_u8_to_gray(img12[b], start[b].tolist(), (ph, pw))
        assert torch.equal(full[b].cpu(), f) and torch.equal(patch[b].cpu(), p) and torch.equal(rgb[b].cpu(), r)


def test_pairs_u8_errors():
    x = torch.zeros(1, 6, 8, 6, dtype=torch.uint8, device=DEV)          # W % 4 != 0
    with pytest.raises(Exception):
        ops.pairs_u8_to_gray(x)
    x = torch.zeros(1, 6, 8, 8, dtype=torch.uint8, device=DEV)
    with pytest.raises(ValueError):
        ops.pairs_u8_to_gray(x, start=torch.tensor([[4, 0]]), patch_size=(8, 8))   # window leaves the image
    with pytest.raises(ValueError):
        ops.pairs_u8_to_gray(x.float())


def test_gt_flow_golden_bit_exact():
    d = load("gt_flow")
    h, w = [int(v) for v in d["hw"]]
    flow = cdl.homo_convert_to_flow(d["H"].numpy(), (h, w))
    assert torch.equal(flow.cpu(), d["flow"])
    assert np.array_equal(cdl.homo_scale(360, 640, d["H"][0].numpy(), h, w), d["H_scaled"][0].numpy())


def test_gt_flow_matches_oracle_360x640():
    rs = np.random.default_rng(9)
    Hm = np.stack([np.eye(3) + rs.normal(size=(3, 3)) * np.array([[1e-2, 1e-2, 8.0], [1e-2, 1e-2, 8.0], [2e-5, 2e-5, 0.0]])
                   for _ in range(2)])
    flow = cdl.homo_convert_to_flow(Hm, (360, 640)).cpu()
    for b in range(2):
        assert torch.equal(flow[b], port.homo_convert_to_flow(Hm[b], (360, 640))[0])


def test_upsample_golden():
    d = load("upsample")
    fl = d["flow"]
    for name, rate in {"x4_rate": True, "odd_rate": True, "down": False, "x2_norate": False}.items():
        ref = d[name]
        out = ops.flow_upsample(fl.to(DEV), ref.shape[-2:], if_rate=rate).cpu()
        assert (out - ref).abs().max().item() < 1e-5, name


def test_upsample_compat_side_effect_and_backward():
    """compat.upsample2d_flow_as scales its input in place like the reference (HEM/model/utils.py:562-565);
    the adjoint kernel against autograd through F.interpolate."""
    fl = torch.randn(3, 2, 12, 20, generator=g(71)) * 2
    tgt = torch.zeros(3, 5, 48, 80)
    a = fl.clone().to(DEV)
    out = hem_utils.upsample2d_flow_as(a, tgt.to(DEV), if_rate=True)
    b = fl.clone()
    ref = port.upsample2d_flow_as(b, tgt, if_rate=True)
    assert (out.cpu() - ref).abs().max().item() < 1e-5
    assert torch.equal(a.cpu()[:, 0], fl[:, 0] * (80 / 20)) and torch.equal(a.cpu()[:, 1], fl[:, 1] * (48 / 12))
    for (ho, wo, rate) in [(48, 80, True), (30, 33, False), (6, 10, False), (12, 20, True)]:
        x = fl.clone().requires_grad_(True)
        go = torch.randn(3, 2, ho, wo, generator=g(72))
        port.upsample2d_flow_as(x, torch.zeros(1, 1, ho, wo), if_rate=rate).backward(go)
        xg = fl.clone().to(DEV).requires_grad_(True)
        ops.flow_upsample(xg, (ho, wo), if_rate=rate).backward(go.to(DEV))
        assert (xg.grad.cpu() - x.grad).abs().max().item() < 1e-4, (ho, wo, rate)


@pytest.mark.parametrize("C", [12, 24])
def test_pyramid_level_feature_warp(C):
    """swin_multi.py:161-166: basis flow at patch resolution -> upsample2d_flow_as(if_rate) to the pyramid level ->
    get_warp_flow of C = 12 / 24 feature maps; forward and the gradients to the features and to the basis weights."""
    B, hp, wp, hl, wl = 2, 32, 48, 16, 24
    basis = port.gen_basis(hp, wp)
    wgt = ((torch.rand(B, 8, 1, generator=g(81)) * 2 - 1) * 2.0)
    feat = torch.randn(B, C, hl, wl, generator=g(82))
    go = torch.randn(B, C, hl, wl, generator=g(83))

    wc, fc = wgt.clone().requires_grad_(True), feat.clone().requires_grad_(True)
    flow_c = (basis.reshape(1, 8, -1) * wc).sum(1).reshape(B, 2, hp, wp)
    up_c = port.upsample2d_flow_as(flow_c, feat, if_rate=True)
    out_c = port.get_warp_flow(fc, up_c)
    out_c.backward(go)

    wg, fg = wgt.clone().to(DEV).requires_grad_(True), feat.clone().to(DEV).requires_grad_(True)
    flow_g = ops.basis_combine(basis.to(DEV), wg, hp, wp)
    up_g = ops.flow_upsample(flow_g, (hl, wl), if_rate=True)
    out_g = hem_utils.get_warp_flow(fg, up_g)
    out_g.backward(go.to(DEV))
    assert (up_g.detach().cpu() - up_c.detach()).abs().max().item() < 1e-5
    assert (out_g.detach().cpu() - out_c.detach()).abs().max().item() < 1e-4
    assert (fg.grad.cpu() - fc.grad).abs().max().item() < 1e-4
    assert ((wg.grad.cpu() - wc.grad).norm() / wc.grad.norm()).item() < 1e-3
