"""GPU parity of the rows either side of the path: the uint8 pair format,
the loaders' ground-truth flow and the flow upsample + pyramid-level feature warp - against the committed
fixtures of the REAL reference (tests/golden) and against the CPU oracle on other shapes.
Bars: integer / byte / fp64-derived values bit-exact; interpolated fp32 values within 1e-5 absolute."""
import os

import numpy as np
import pytest
import torch

from dmhomo_b200 import ops
from dmhomo_b200.compat import data_loader as cdl, dgm, hem_utils
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


@pytest.mark.parametrize("C", [2, 6, 7, 12])
def test_multichannel_flow_warp_channel_groups(C):
    """Feature maps with C other than 1 / 3 run on the lean kernel as channel groups of 3 (C % 3 == 0) or 1: pixels
    bit-exact against get_warp_flow (HEM/model/utils.py:443-553), dL/dfeat and dL/dflow (summed over the groups of a
    sample) within 1e-4; the S3 sampler (DGM flow_warp) through the same grouping."""
    B, h, w = 3, 40, 72
    feat = torch.randn(B, C, h, w, generator=g(91))
    flow = torch.randn(B, 2, h, w, generator=g(92)) * 5
    go = torch.randn(B, C, h, w, generator=g(93))
    fc, lc = feat.clone().requires_grad_(True), flow.clone().requires_grad_(True)
    out_c = port.get_warp_flow(fc, lc)
    out_c.backward(go)
    fg, lg = feat.clone().to(DEV).requires_grad_(True), flow.clone().to(DEV).requires_grad_(True)
    out_g = hem_utils.get_warp_flow(fg, lg)
    out_g.backward(go.to(DEV))
    assert torch.equal(out_g.detach().cpu(), out_c.detach())
    assert (fg.grad.cpu() - fc.grad).abs().max().item() < 1e-4
    scale = max(1.0, lc.grad.abs().max().item())
    assert ((lg.grad.cpu() - lc.grad).abs().max().item() / scale) < 1e-4
    # mask + S3
    out_m, mask = ops.warp(feat.to(DEV), flow.to(DEV), kind=ops.PARAM_FLOW, return_mask=True)
    assert torch.equal(out_m.cpu(), out_c.detach()) and torch.equal(mask.cpu(), port.correspondence_mask(flow))
    f3, l3 = feat.clone().requires_grad_(True), flow.clone().requires_grad_(True)
    o3 = port.flow_warp(f3, l3)
    o3.backward(go)
    g3, m3 = feat.clone().to(DEV).requires_grad_(True), flow.clone().to(DEV).requires_grad_(True)
    og = dgm.flow_warp(g3, m3)
    og.backward(go.to(DEV))
    assert (og.detach().cpu() - o3.detach()).abs().max().item() < 1e-4
    assert (g3.grad.cpu() - f3.grad).abs().max().item() < 1e-4
    assert ((m3.grad.cpu() - l3.grad).abs().max().item() / max(1.0, l3.grad.abs().max().item())) < 1e-4


# ---------------------------------------------------------------------------------- remaining helpers of the module files
def test_normalize_family_bit_exact():
    from dmhomo_b200.compat import flow_and_mapping_operations as fmo

    gen = torch.Generator().manual_seed(61)
    pix = torch.rand(3, 2, 37, 52, generator=gen) * 60 - 4
    nrm = torch.rand(3, 2, 37, 52, generator=gen) * 2.4 - 1.2
    assert torch.equal(fmo.normalize(pix.to(DEV)).cpu(), port.grid_normalize(pix, 0))
    assert torch.equal(fmo.unnormalize(nrm.to(DEV)).cpu(), port.grid_normalize(nrm, 1))
    assert torch.equal(fmo.unormalise_flow_or_mapping(nrm.to(DEV)).cpu(), port.grid_normalize(nrm, 1))
    assert torch.equal(fmo.unormalise_and_convert_mapping_to_flow(nrm.to(DEV)).cpu(), port.grid_normalize(nrm, 2))
    # channel-last in / out and 3-D inputs, as the reference accepts them
    cl = fmo.unnormalize(nrm.permute(0, 2, 3, 1).to(DEV), output_channel_first=False).cpu()
    assert torch.equal(cl, port.grid_normalize(nrm, 1).permute(0, 2, 3, 1))
    assert torch.equal(fmo.normalize(pix[0].to(DEV)).cpu(), port.grid_normalize(pix[:1], 0)[0])


def test_crop_patch_from_full():
    gen = torch.Generator().manual_seed(62)
    img = torch.rand(3, 2, 40, 56, generator=gen)
    start_i = torch.tensor([[[3, 2]], [[0, 0]], [[20, 12]]])
    a = hem_utils.CropPatchFromFull((24, 20), img.to(DEV), start_i.to(DEV), rescale=False)
    assert torch.equal(a.cpu(), port.crop_patch_from_full((24, 20), img, start_i, rescale=False))
    start_f = torch.tensor([[[3.25, 2.5]], [[-1.5, 0.75]], [[40.5, 27.25]]])
    b = hem_utils.CropPatchFromFull((24, 20), img.to(DEV), start_f.to(DEV), rescale=True)
    assert (b.cpu() - port.crop_patch_from_full((24, 20), img, start_f, rescale=True)).abs().max().item() < 1e-6


def test_resize_flow_matches_cv2():
    from dmhomo_b200.compat import dgm

    rs = np.random.default_rng(63)
    for (h, w, size) in ((36, 64, 24), (45, 80, 128), (360, 640, 256)):
        fl = rs.standard_normal((h, w, 2)).astype(np.float32) * 5
        ours = dgm.resize_flow(fl, size)
        ref = port.resize_flow(fl.copy(), size)
        assert ours.shape == ref.shape
        assert np.abs(ours - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())


# ---------------------------------------------------------------------------------- formats / resident pipelines of bench.py
def test_u8_to_f32_and_planar_patch():
    rs = np.random.default_rng(64)
    u8 = torch.from_numpy(rs.integers(0, 256, size=(2, 3, 1, 40, 52), dtype=np.uint8))
    out = ops.u8_to_f32(u8.to(DEV), 1.0 / 255.0, 0.0)
    assert torch.equal(out.cpu(), u8.float() * (1.0 / 255.0))
    odd = u8.flatten()[:1003]                                   # scalar tail + unaligned length
    assert torch.equal(ops.u8_to_f32(odd.to(DEV), 0.5, -3.0).cpu(), odd.float() * 0.5 + (-3.0))
    with pytest.raises(ValueError):
        ops.u8_to_f32(u8.float().to(DEV))
    # planar patch layout (2,B,ph,pw) == the (B,2,ph,pw) patch transposed, written into a caller-owned buffer
    img12 = torch.from_numpy(rs.integers(0, 256, size=(3, 6, 48, 64), dtype=np.uint8)).to(DEV)
    start = torch.tensor([[7, 11], [0, 0], [24, 24]])
    _, patch, _ = ops.pairs_u8_to_gray(img12, start=start, patch_size=(24, 40), want_rgb=False)
    buf = torch.empty(2, 3, 1, 24, 40, device=DEV)
    full, planar, rgb = ops.pairs_u8_to_gray(img12, start=start.to(DEV), patch_size=(24, 40), want_rgb=False, want_full=False,
                                            patch_planar=True, patch_out=buf)
    assert full is None and rgb is None and planar.data_ptr() == buf.data_ptr()
    assert torch.equal(planar, patch.transpose(0, 1))


@pytest.mark.parametrize("C,h,w", [(1, 360, 640), (3, 128, 192), (1, 72, 100)])
def test_warp_eval_and_warp_into(C, h, w):
    """The evaluation pass (warped + mask + loss, one launch) and the buffer-reusing forward warp against warp() /
    warp_loss(); (1, 72, 100) has a width the tile kernel does not take (general kernel)."""
    B = 4
    gen = g(65)
    img1, img2 = torch.rand(B, C, h, w, generator=gen).to(DEV), torch.rand(B, C, h, w, generator=gen).to(DEV)
    src = port.corner_points(2 * B, h, w)
    H = port.dlt4(src, src + (torch.rand(2 * B, 4, 2, generator=gen) * 2 - 1) * 12.0).to(DEV)
    loss, outs, masks = ops.warp_eval([ops.WarpTerm(img2, img1, H[:B]), ops.WarpTerm(img1, img2, H[B:])], kind=ops.PARAM_HOMOGRAPHY)
    w2, m2 = ops.warp(img2, H[:B], kind=ops.PARAM_HOMOGRAPHY, return_mask=True)
    w1, m1 = ops.warp(img1, H[B:], kind=ops.PARAM_HOMOGRAPHY, return_mask=True)
    assert torch.equal(outs[0], w2) and torch.equal(outs[1], w1)
    assert torch.equal(masks[0], m2) and torch.equal(masks[1], m1)
    ref = ops.warp_loss([ops.WarpTerm(img2, img1, H[:B]), ops.WarpTerm(img1, img2, H[B:])], kind=ops.PARAM_HOMOGRAPHY, fused=False)
    assert abs(loss.item() - ref.item()) < 1e-6
    out = torch.full_like(img2, -7.0)
    valid = torch.full((B, h, w), 9, dtype=torch.uint8, device=DEV)
    ops.warp_into(img2, H[:B], out, valid, kind=ops.PARAM_HOMOGRAPHY)
    assert torch.equal(out, w2) and torch.equal(valid.bool(), m2)


def test_warp_perspective_u8_matches_cv2_bit_for_bit():
    """uint8 S4: the {"imgs","homos"} sample batches and their cv2.warpPerspective check."""
    import cv2

    from dmhomo_b200 import synth
    from dmhomo_b200.compat import dgm

    rs = np.random.default_rng(66)
    N, h, w = 5, 256, 256
    imgs = rs.integers(0, 256, size=(N, 6, h, w), dtype=np.uint8)
    H360 = synth.homographies_360x640(N, synth.generator(), 32.0)
    homos = np.stack([dgm.adapt_homography_to_preprocessing_v3(360, 640, H360[i], h, w) for i in range(N)])
    ours = dgm.warp_pairs_u8(imgs, homos)
    assert ours.dtype == np.uint8 and ours.shape == (N, 3, h, w)
    for i in range(N):
        ref = cv2.warpPerspective(np.ascontiguousarray(imgs[i, :3].transpose(1, 2, 0)), homos[i], (w, h)).transpose(2, 0, 1)
        assert np.array_equal(ours[i], ref), f"sample {i}: {np.abs(ours[i].astype(int) - ref.astype(int)).max()} LSB off"
    # channels-last, non-square, partial blocks
    img = torch.from_numpy(rs.integers(0, 256, size=(2, 100, 132, 3), dtype=np.uint8))
    Hm = np.stack([np.diag([132 / 640, 100 / 360, 1.0]) @ H360[i] @ np.diag([640 / 132, 360 / 100, 1.0]) for i in range(2)])
    out = ops.warp_perspective(img.to(DEV), torch.from_numpy(Hm).to(DEV), (132, 100), channels_last=True).cpu().numpy()
    for i in range(2):
        assert np.array_equal(out[i], cv2.warpPerspective(img[i].numpy(), Hm[i], (132, 100)))
    samples = dgm.split_sample_batches([{"imgs": imgs[:2], "homos": homos[:2]}, {"imgs": imgs[2:], "homos": homos[2:]}])
    assert len(samples) == N and samples[3]["img12"].shape == (6, h, w) and np.array_equal(samples[3]["homo12"], homos[3])


def test_remap_matches_cv2():
    """remap_using_flow_fields / remap_using_correspondence_map (HEM/utils_operations/pixel_wise_mapping.py:7-52) against
    cv2.remap itself, fp32 and uint8 images, 1 / 3 channels, coordinates inside, on the border, far outside, non-finite:
    bit-identical."""
    from dmhomo_b200.compat import pixel_wise_mapping as pwm

    rs = np.random.default_rng(5)
    for (h, w, c) in ((48, 80, 3), (37, 53, 1), (64, 64, 3)):
        img_f = rs.random((h, w, c), dtype=np.float32) if c > 1 else rs.random((h, w), dtype=np.float32)
        img_u = rs.integers(0, 256, size=img_f.shape, dtype=np.uint8)
        dx = (rs.standard_normal((h, w)) * 6).astype(np.float32)
        dy = (rs.standard_normal((h, w)) * 6).astype(np.float32)
        dx[::9, ::7] += 500.0
        dy[3::11, 2::5] -= 70000.0          # beyond OpenCV's int16 coordinate range
        dx[5, 5], dy[6, 6] = np.inf, np.nan
        dx[7, 7] = -3.0e9
        for img in (img_f, img_u):
            ref = port.remap_using_flow_fields(img, dx, dy)
            out = pwm.remap_using_flow_fields(img, dx, dy)
            assert out.dtype == ref.dtype and out.shape == ref.shape
            assert np.array_equal(out, ref), (h, w, c, img.dtype)
            X, Y = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
            mx, my = (X * 0.9 + 2.25 + dx * 0.1).astype(np.float32), (Y * 1.05 - 1.5).astype(np.float32)
            assert np.array_equal(pwm.remap_using_correspondence_map(img, mx, my), port.remap_using_correspondence_map(img, mx, my))
    # batched tensor form, planar layout
    B, C, h, w = 3, 3, 40, 56
    img = torch.rand(B, C, h, w, generator=g(401))
    coords = torch.rand(B, 2, h, w, generator=g(402)) * torch.tensor([w + 8.0, h + 8.0]).view(1, 2, 1, 1) - 4.0
    out = ops.remap(img.to(DEV), coords.to(DEV)).cpu()
    for b in range(B):
        ref = port.remap_using_correspondence_map(img[b].permute(1, 2, 0).contiguous().numpy(), coords[b, 0].numpy(), coords[b, 1].numpy())
        assert np.array_equal(out[b].permute(1, 2, 0).numpy(), ref)

