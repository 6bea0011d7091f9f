"""Pins oracle/port.py against the real reference executed in place (build container
only: skipped wherever /root/reference is not mounted, e.g. on the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import port, ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference not mounted")


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()


def _rand_offsets(B, rho, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(B, 4, 2, generator=g) * 2 - 1) * rho


def _images(B, C, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, C, h, w, generator=g)


def test_grid(ref):
    assert torch.equal(ref.utils.get_grid(3, 7, 9), port.pixel_grid(3, 7, 9))
    assert torch.equal(ref.utils.get_grid(2, 5, 4, 3), port.pixel_grid(2, 5, 4, 3))


def test_dlt4_bit_exact(ref):
    B, h, w = 8, 90, 160
    src = port.corner_points(B, h, w)
    dst = src + _rand_offsets(B, 8, 1)
    assert torch.equal(ref.utils.DLT(B)(src, dst), port.dlt4(src, dst))
    off = _rand_offsets(B, 8, 2).reshape(B, 8)
    ref_H = ref.utils.WarpMat(off.clone(), (w, h), (w, h))
    assert torch.equal(ref_H, port.warp_mat(off, (w, h), (w, h)))


def test_dlt_solve_variants(ref):
    B, h, w = 5, 64, 96
    src = port.corner_points(B, h, w).reshape(B, 8)
    off = _rand_offsets(B, 6, 3).reshape(B, 8)
    assert torch.equal(ref.net.DLT_solve(src, off), port.dlt_solve_h4pt(src, off))
    # 2x2 mesh (9 points) through the flat-index variant
    d = 2
    mesh = port.mesh_source_points(B, h, w, d)
    assert torch.equal(mesh, ref.utils.get_src_p(B, h, w, d))
    moff = torch.randn(B, 2, d + 1, d + 1, generator=torch.Generator().manual_seed(4))
    assert torch.equal(ref.utils.DLT_solve(mesh, moff), port.dlt_solve_mesh(mesh, moff))
    flat = mesh.permute(0, 2, 3, 1).reshape(B, -1)
    foff = moff.permute(0, 2, 3, 1).reshape(B, -1)
    assert torch.equal(ref.net.DLT_solve(flat, foff), port.dlt_solve_h4pt(flat, foff))


def test_get_flow_bit_exact(ref):
    B, h, w = 4, 90, 160
    src = port.corner_points(B, h, w)
    H = port.dlt4(src, src + _rand_offsets(B, 8, 5))
    f_ref, vg_ref = ref.utils.get_flow(H.reshape(B, 1, 3, 3), ref.utils.get_grid(B, h, w), h, w, 1)
    f, vg = port.homography_to_flow(H, h, w)
    assert torch.equal(f_ref, f) and torch.equal(vg_ref, vg)
    # mesh of 2x2 homographies
    d = 2
    mesh = port.mesh_source_points(B, h, w, d)
    Hm = port.dlt_solve_mesh(mesh, torch.randn(B, 2, d + 1, d + 1, generator=torch.Generator().manual_seed(6)))
    f_ref, _ = ref.utils.get_flow(Hm, ref.utils.get_grid(B, h, w), h, w, d)
    f, _ = port.homography_to_flow(Hm, h, w, divide=d)
    assert torch.equal(f_ref, f)


def test_get_warp_flow_bit_exact_and_grads(ref):
    B, C, h, w = 3, 2, 40, 56
    img = _images(B, C, h + 6, w + 4, 7)  # source larger than output
    flow = torch.randn(B, 2, h, w, generator=torch.Generator().manual_seed(8)) * 6
    a = ref.utils.get_warp_flow(img, flow, start=2)
    b = port.get_warp_flow(img, flow, start=2)
    assert torch.equal(a, b)
    i1 = img.clone().requires_grad_(True)
    f1 = flow.clone().requires_grad_(True)
    i2 = img.clone().requires_grad_(True)
    f2 = flow.clone().requires_grad_(True)
    go = torch.randn(B, C, h, w, generator=torch.Generator().manual_seed(9))
    (ref.utils.get_warp_flow(i1, f1) * go).sum().backward()
    (port.get_warp_flow(i2, f2) * go).sum().backward()
    assert torch.allclose(i1.grad, i2.grad, atol=1e-6)
    assert torch.allclose(f1.grad, f2.grad, atol=1e-6)


def test_warp_images_s1b(ref):
    B, C, h, w = 2, 1, 48, 64
    img = _images(B, C, h, w, 10)
    src = port.corner_points(B, 32, 40)
    H = port.dlt4(src, src + _rand_offsets(B, 4, 11))
    start = torch.tensor([[3.0, 5.0], [10.0, 7.0]]).view(B, 2, 1, 1)
    a, fa = ref.utils.WarpImages(img, H, start, (40, 32))
    b, fb = port.warp_images_s1b(img, H, start, (40, 32))
    assert torch.equal(a, b) and torch.equal(fa, fb)


def test_grid_sample_warps(ref):
    B, C, h, w = 2, 3, 32, 48
    img = _images(B, C, h, w, 12)
    flow = torch.randn(B, 2, h, w, generator=torch.Generator().manual_seed(13)) * 5
    assert torch.equal(ref.pwm.warp(img, flow), port.warp_zeros(img, flow))
    assert torch.equal(ref.pwm.warp_with_mapping(img, flow + 3), port.warp_with_mapping(img, flow + 3))
    assert torch.equal(ref.ddpm.flow_warp(img, flow), port.flow_warp(img, flow))
    assert torch.equal(ref.data_loader.flow_warp(img, flow), port.flow_warp(img, flow))


def test_masks(ref):
    flow = torch.randn(2, 2, 20, 30, generator=torch.Generator().manual_seed(14)) * 12
    assert torch.equal(ref.fmo.get_gt_correspondence_mask(flow), port.correspondence_mask(flow))
    assert torch.equal(ref.fmo.create_border_mask(flow), port.border_mask(flow))
    assert torch.equal(ref.fmo.get_gt_correspondence_mask(flow[0]), port.correspondence_mask(flow[0]))
    img = _images(2, 3, 8, 9, 15)
    img[:, :, :3] = 0
    assert torch.equal(ref.fmo.define_mask_zero_borders(img), port.zero_border_mask(img))


def test_basis(ref):
    h, w = 32, 48
    a = ref.utils.gen_basis(h, w)
    b = port.gen_basis(h, w)
    assert torch.equal(a, b)
    basis = b.reshape(1, 8, -1)
    wt = torch.randn(3, 8, 1, generator=torch.Generator().manual_seed(16))
    flow = port.basis_combine(basis, wt, h, w)
    # sequential, separately rounded accumulation == the reference's (basis*w).sum(1)
    acc = basis[:, 0] * wt[:, 0]
    for k in range(1, 8):
        acc = acc + basis[:, k] * wt[:, k]
    assert torch.equal(flow, acc.reshape(3, 2, h, w))
    off = port.basis_corner_offsets(basis, wt, h, w)
    fl = flow
    exp = torch.stack([fl[:, :, 0, 0], fl[:, :, 0, w - 1], fl[:, :, h - 1, 0], fl[:, :, h - 1, w - 1]], 1)
    assert torch.equal(off, exp)


def test_losses(ref):
    a = _images(2, 1, 16, 16, 17)
    b = _images(2, 1, 16, 16, 18)
    m = (_images(2, 1, 16, 16, 19) > 0.3).float()
    assert torch.equal(ref.losses.LossL1()(m * a, m * b), port.masked_l1(m, a, b))


def test_numpy_flow_helpers(ref):
    rng = np.random.default_rng(20)
    Hm = np.eye(3) + rng.normal(size=(3, 3)) * np.array([[1e-2, 1e-2, 3], [1e-2, 1e-2, 3], [1e-5, 1e-5, 0]])
    a = ref.ddpm.homo_to_flow(Hm.reshape(1, 1, 3, 3), 40, 56)
    assert np.array_equal(a, port.homo_to_flow_np(Hm, 40, 56))
    mx, my = ref.fmo.from_homography_to_pixel_wise_mapping((40, 56), Hm)
    px, py = port.homography_to_mapping_np((40, 56), Hm)
    assert np.array_equal(mx, px) and np.array_equal(my, py)
    assert torch.equal(ref.data_loader.homo_convert_to_flow(Hm, (40, 56)), port.homo_convert_to_flow(Hm, (40, 56)))
    assert np.array_equal(ref.data_loader.homo_scale(360, 640, Hm, 256, 256), port.homo_scale(360, 640, Hm, 256, 256))
    assert np.array_equal(ref.ddpm.adapt_homography_to_preprocessing_v3(360, 640, Hm, 256, 256),
                          port.homo_scale(360, 640, Hm, 256, 256))


def test_flow_to_image(ref):
    flow = (np.random.default_rng(21).normal(size=(24, 32, 2)) * 20).astype(np.float32)
    flow[0, 0] = 0
    assert np.array_equal(ref.ddpm.flow_to_image(flow), port.flow_to_image(flow))
    t = torch.from_numpy(flow.transpose(2, 0, 1)[None])
    assert torch.equal(ref.ddpm.visulize_flow(t), port.visualize_flow(t))


def test_warp_perspective_emulation_matches_cv2():
    rng = np.random.default_rng(22)
    img = rng.random((64, 80, 3), dtype=np.float32)
    Hm = np.eye(3) + rng.normal(size=(3, 3)) * np.array([[2e-2, 2e-2, 4], [2e-2, 2e-2, 4], [1e-4, 1e-4, 0]])
    a = port.warp_perspective_cv2(img, Hm, (80, 64))
    b = port.warp_perspective_emul(img, Hm, (80, 64))
    assert np.abs(a - b).max() < 5e-7


def test_homo_gen(ref):
    B, h, w = 2, 24, 32
    src = port.corner_points(B, h, w)
    H = port.dlt4(src, src + _rand_offsets(B, 3, 23))
    flow, _ = port.homography_to_flow(H, h, w)
    a = ref.ddpm.homo_gen(flow)
    b = port.homo_gen(flow)
    assert torch.allclose(a, b, rtol=0, atol=1e-12)
    assert torch.allclose(b.reshape(B, 3, 3).float(), H, atol=2e-4)


def test_eval_point_error(ref):
    g = torch.Generator().manual_seed(24)
    flow_f = torch.randn(2, 20, 30, 2, generator=g)
    flow_b = torch.randn(2, 20, 30, 2, generator=g)
    pts = torch.rand(2, 6, 2, 2, generator=g) * torch.tensor([29.0, 19.0])
    a = ref.losses.compute_eval_results({"imgs_gray_full": torch.zeros(2, 2, 20, 30), "pt_set": pts},
                                        {"flow_f": flow_f, "flow_b": flow_b})
    b = port.eval_point_errors(pts, flow_f, flow_b)
    assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_upsample_flow(ref):
    fl = torch.randn(2, 2, 10, 12, generator=torch.Generator().manual_seed(25))
    tgt = torch.zeros(2, 1, 15, 20)
    a = ref.utils.upsample2d_flow_as(fl.clone(), tgt, if_rate=True)
    assert torch.equal(a, port.upsample2d_flow_as(fl, tgt, if_rate=True))


def test_pairs_u8_against_data_aug(ref):
    """oracle restatement of __getitem__ + data_aug vs the reference's own method."""
    import types

    DL = ref.data_loader
    rs = np.random.default_rng(3)
    img12 = rs.integers(0, 256, size=(6, 48, 64), dtype=np.uint8)
    me = types.SimpleNamespace(mean_I=np.array([118.93, 113.97, 102.60]).reshape(1, 1, 3),
                               std_I=np.array([69.85, 68.81, 72.45]).reshape(1, 1, 3), crop_size=(24, 40), rho=0)
    hwc = img12.transpose(1, 2, 0)
    o = DL.DGMTrainData.data_aug(me, hwc[..., :3], hwc[..., 3:], np.eye(3), np.eye(3), start=[7, 11])
    full, patch, rgb = port.pairs_u8_to_gray(img12, [7, 11], (24, 40))
    assert torch.equal(full, torch.cat((o[0], o[1]), dim=2).permute(2, 0, 1).float())
    assert torch.equal(patch, torch.cat((o[2], o[3]), dim=2).permute(2, 0, 1).float())
    assert o[8] == [7, 11]


def test_normalize_family(ref):
    g = torch.Generator().manual_seed(31)
    pix = torch.rand(2, 2, 9, 14, generator=g) * 12
    nrm = torch.rand(2, 2, 9, 14, generator=g) * 2 - 1
    assert torch.equal(ref.fmo.normalize(pix.clone()), port.grid_normalize(pix, 0))
    assert torch.equal(ref.fmo.unnormalize(nrm.clone()), port.grid_normalize(nrm, 1))
    assert torch.equal(ref.fmo.unormalise_flow_or_mapping(nrm.clone()), port.grid_normalize(nrm, 1))
    assert torch.equal(ref.fmo.unormalise_and_convert_mapping_to_flow(nrm.clone()), port.grid_normalize(nrm, 2))


def test_crop_patch_from_full(ref):
    g = torch.Generator().manual_seed(32)
    img = torch.rand(3, 2, 20, 28, generator=g)
    start_i = torch.tensor([[[3, 2]], [[0, 0]], [[9, 6]]])
    a = ref.utils.CropPatchFromFull((12, 10), img, start_i, rescale=False)
    assert torch.equal(a, port.crop_patch_from_full((12, 10), img, start_i, rescale=False))
    start_f = torch.tensor([[[3.25, 2.5]], [[-1.5, 0.75]], [[17.5, 11.25]]])     # windows that leave the image on both sides
    b = ref.utils.CropPatchFromFull((12, 10), img, start_f, rescale=True)
    assert torch.equal(b, port.crop_patch_from_full((12, 10), img, start_f, rescale=True))


def test_resize_flow(ref):
    if ref.ddpm is None:
        pytest.skip(f"ddpm module not importable here: {getattr(ref, 'ddpm_error', '')}")
    rs = np.random.default_rng(33)
    fl = rs.standard_normal((36, 64, 2)).astype(np.float32) * 5
    assert np.array_equal(ref.ddpm.resize_flow(fl.copy(), 24), port.resize_flow(fl.copy(), 24))
