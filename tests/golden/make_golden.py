#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REAL reference (imported in place from /root/reference)
on small seeded inputs.  Only runnable in the build container; the fixtures it writes are committed so
that the oracle and the CUDA path can be checked against the reference's own outputs anywhere
(the GPU box has no /root/reference).

    CUDA_VISIBLE_DEVICES="" python tests/golden/make_golden.py [fixture names ...]

Every array is stored with the inputs that produced it, so a test never has to re-derive inputs.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def g(seed):
    return torch.Generator().manual_seed(seed)


def corners(B, h, w):
    c = torch.tensor([[0, 0], [w - 1, 0], [0, h - 1], [w - 1, h - 1]], dtype=torch.float32)
    return c.view(1, 4, 2).repeat(B, 1, 1)


ONLY = set(sys.argv[1:])   # optional: fixture names to (re)write; the others are left untouched


def save(name, **arrs):
    if ONLY and name not in ONLY:
        return
    out = {}
    for k, v in arrs.items():
        out[k] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}.npz: " + ", ".join(f"{k}{tuple(v.shape)}" for k, v in out.items()))


def main():
    torch.manual_seed(230)
    torch.set_num_threads(1)
    r = ref_loader.load()
    U, F, P, L = r.utils, r.fmo, r.pwm, r.losses

    # ---- A1/A2/A3: DLT ------------------------------------------------------------------------
    B, h, w = 6, 90, 160
    src = corners(B, h, w)
    off = (torch.rand(B, 4, 2, generator=g(1)) * 2 - 1) * 8
    H = U.DLT(B)(src, src + off)
    H_net = r.net.DLT_solve(src.reshape(B, 8), off.reshape(B, 8))
    H_wm = U.WarpMat(off.reshape(B, 8).clone(), (w, h), (w, h))
    d = 2
    mesh = U.get_src_p(B, h, w, d)
    moff = torch.randn(B, 2, d + 1, d + 1, generator=g(2))
    H_mesh = U.DLT_solve(mesh, moff)
    save("dlt", src=src, off=off, H=H, H_net=H_net, H_warpmat=H_wm, mesh=mesh, mesh_off=moff, H_mesh=H_mesh,
         hw=np.array([h, w]))

    # ---- A4/A5: grid + get_flow ------------------------------------------------------------------
    grid = U.get_grid(B, h, w, 0)
    flow, vgrid = U.get_flow(H.view(B, 1, 3, 3), grid, h, w, 1)
    grid3 = U.get_grid(B, h, w, 3)
    flow3, _ = U.get_flow(H.view(B, 1, 3, 3), grid3, h, w, 1)
    flow_mesh, _ = U.get_flow(H_mesh, grid, h, w, d)
    save("get_flow", H=H, H_mesh=H_mesh, flow=flow, flow_start3=flow3, flow_mesh=flow_mesh, hw=np.array([h, w]))

    # ---- A6: get_warp_flow / transformer (S1), + autograd gradients ----------------------------------
    B2, C, hs, ws, ho, wo = 3, 3, 40, 72, 33, 65
    img = torch.rand(B2, C, hs, ws, generator=g(3)).requires_grad_(True)
    fl = (torch.randn(B2, 2, ho, wo, generator=g(4)) * 6).requires_grad_(True)
    out = U.get_warp_flow(img, fl, start=2)
    go = torch.randn(B2, C, ho, wo, generator=g(5))
    (out * go).sum().backward()
    vg = (U.get_grid(B2, ho, wo, 0)[:, :2] + fl.detach())
    out_tr = U.transformer(img.detach(), vg)
    save("s1_warp", img=img, flow=fl, start=np.array(2), out=out, grad_out=go, grad_img=img.grad, grad_flow=fl.grad,
         vgrid=vg, out_transformer=out_tr)

    # ---- A6 through a homography (cfg1 pipeline at small size): warped pair + masks + loss ---------
    B3, h3, w3 = 4, 72, 128
    i1 = torch.rand(B3, 1, h3, w3, generator=g(6))
    i2 = torch.rand(B3, 1, h3, w3, generator=g(7))
    s3 = corners(B3, h3, w3)
    of = (torch.rand(B3, 4, 2, generator=g(8)) * 2 - 1) * 8
    ob = (torch.rand(B3, 4, 2, generator=g(9)) * 2 - 1) * 8
    Hf, Hb = U.DLT(B3)(s3, s3 + of), U.DLT(B3)(s3, s3 + ob)
    g3 = U.get_grid(B3, h3, w3, 0)
    ff, _ = U.get_flow(Hf.view(B3, 1, 3, 3), g3, h3, w3, 1)
    fb, _ = U.get_flow(Hb.view(B3, 1, 3, 3), g3, h3, w3, 1)
    w2, w1 = U.get_warp_flow(i2, ff), U.get_warp_flow(i1, fb)
    mf, mb = F.create_border_mask(ff).unsqueeze(1), F.create_border_mask(fb).unsqueeze(1)
    l1 = L.LossL1(reduction="mean")
    loss = l1(mf * i1, mf * w2) + l1(mb * i2, mb * w1)
    save("pipeline_h4pt", img1=i1, img2=i2, off_f=of, off_b=ob, Hf=Hf, Hb=Hb, flow_f=ff, flow_b=fb, w2=w2, w1=w1,
         mask_f=mf, mask_b=mb, loss=loss)

    # ---- A7: WarpImages (S1b) -----------------------------------------------------------------------
    img7 = torch.rand(2, 1, 48, 64, generator=g(10))
    s7 = corners(2, 32, 40)
    H7 = U.DLT(2)(s7, s7 + (torch.rand(2, 4, 2, generator=g(11)) * 2 - 1) * 4)
    st7 = torch.tensor([[3.0, 5.0], [20.0, 12.0]]).view(2, 2, 1, 1)
    o7, f7 = U.WarpImages(img7, H7, st7, (40, 32))
    save("warp_images", img=img7, H=H7, start=st7, out=o7, flow=f7, patch_wh=np.array([40, 32]))

    # ---- A8/A9: grid_sample warps (S2, S3) ------------------------------------------------------------
    x8 = torch.rand(2, 3, 32, 48, generator=g(12))
    f8 = torch.randn(2, 2, 32, 48, generator=g(13)) * 6
    save("grid_sample_warps", img=x8, flow=f8, warp_zeros=P.warp(x8, f8), warp_mapping=P.warp_with_mapping(x8, f8 + 3),
         flow_warp=r.data_loader.flow_warp(x8, f8))

    # ---- A10/A11: masks --------------------------------------------------------------------------------
    f10 = torch.randn(2, 2, 20, 30, generator=g(14)) * 12
    im10 = torch.rand(2, 3, 8, 9, generator=g(15))
    im10[:, :, :3] = 0
    save("masks", flow=f10, corr=F.get_gt_correspondence_mask(f10), border=F.create_border_mask(f10), image=im10,
         zero_border=F.define_mask_zero_borders(im10), mapping=F.convert_flow_to_mapping(f10))

    # ---- A12: basis -----------------------------------------------------------------------------------
    hb, wb = 32, 48
    basis = U.gen_basis(hb, wb)                       # (8, 2*h*w) per the reference's reshape in callers
    basis = basis.reshape(8, -1) if basis.dim() != 2 else basis
    wt = (torch.rand(3, 8, 1, generator=g(16)) * 2 - 1) * 4
    bflow = (basis.unsqueeze(0) * wt).sum(1).reshape(3, 2, hb, wb)
    save("basis", basis=basis, weight=wt, flow=bflow, hw=np.array([hb, wb]))

    # ---- A15/A16/A17: DGM rendering ----------------------------------------------------------------------
    D = r.ddpm
    rng = np.random.default_rng(17)
    Hs = np.stack([np.eye(3) + rng.normal(size=(3, 3)) * np.array([[1e-2, 1e-2, 3], [1e-2, 1e-2, 3], [1e-5, 1e-5, 0]])
                   for _ in range(3)])
    fl_np = []
    for i in range(3):
        fl_np.append(np.asarray(D.homo_to_flow(Hs[i].reshape(1, 1, 3, 3), 40, 56)))   # (h, w, 2) float32
    fl_np = np.stack(fl_np)
    rgb = np.stack([D.flow_to_image(fl_np[i]) for i in range(3)])
    mx, my = F.from_homography_to_pixel_wise_mapping((40, 56), Hs[0])
    save("dgm_flow", H=Hs, flow=fl_np, rgb=rgb, map_x=mx, map_y=my, hw=np.array([40, 56]))

    import cv2
    imgs = rng.random((3, 64, 80, 3), dtype=np.float32)
    Hp = np.stack([np.eye(3) + rng.normal(size=(3, 3)) * np.array([[2e-2, 2e-2, 4], [2e-2, 2e-2, 4], [1e-4, 1e-4, 0]])
                   for _ in range(3)])
    persp = np.stack([cv2.warpPerspective(imgs[i], Hp[i], (80, 64)) for i in range(3)])
    save("warp_perspective", img=imgs, H=Hp, out=persp, cv2_version=np.array(cv2.__version__))

    # ---- A14: DGM photometric term (flow_warp + masked L1 weighting, cfg.py:784-806 restated from its lines) ----
    # (p_losses itself needs the U-Net; the golden pins flow_warp, the term is pinned in test_oracle_vs_reference)

    # ---- next row 1: homo_gen (least-squares DLT over all pixels) ---------------------------------------------
    Hg = U.DLT(2)(corners(2, 32, 32), corners(2, 32, 32) + (torch.rand(2, 4, 2, generator=g(18)) * 2 - 1) * 3)
    fg, _ = U.get_flow(Hg.view(2, 1, 3, 3), U.get_grid(2, 32, 32, 0), 32, 32, 1)
    fg = fg + torch.randn(2, 2, 32, 32, generator=g(19)) * 0.05
    save("homo_gen", flow=fg, H=D.homo_gen(fg))

    # ---- A18: eval point error ------------------------------------------------------------------------------------
    ffe, fbe = torch.randn(3, 20, 30, 2, generator=g(20)), torch.randn(3, 20, 30, 2, generator=g(21))
    pts = torch.rand(3, 6, 2, 2, generator=g(22)) * torch.tensor([29.0, 19.0])
    errs = L.compute_eval_results({"imgs_gray_full": torch.zeros(3, 2, 20, 30), "pt_set": pts}, {"flow_f": ffe, "flow_b": fbe})
    save("eval_points", pts=pts, flow_f=ffe, flow_b=fbe, err=torch.stack([torch.as_tensor(e) for e in errs]))



    # ---- cfg 2 chained through the reference's own functions, with gradients -----------
    # gen_basis -> (basis * w).sum(1) -> the flow at the four corners as 4-pt offsets -> DLT -> get_flow ->
    # get_warp_flow (both directions) -> create_border_mask -> LossL1, backward to both images and both weight sets
    Bc, hc, wc = 3, 64, 96
    basis_c = U.gen_basis(hc, wc).reshape(1, 8, -1)
    c1 = torch.rand(Bc, 1, hc, wc, generator=g(71)).requires_grad_(True)
    c2 = torch.rand(Bc, 1, hc, wc, generator=g(72)).requires_grad_(True)
    wfc = ((torch.rand(Bc, 8, 1, generator=g(73)) * 2 - 1) * 2.0).requires_grad_(True)
    wbc = ((torch.rand(Bc, 8, 1, generator=g(74)) * 2 - 1) * 2.0).requires_grad_(True)
    sc = corners(Bc, hc, wc)

    def corner_offsets(wt):
        fl = (basis_c * wt).sum(1).reshape(Bc, 2, hc, wc)           # net.py:808-809
        pts = [fl[:, :, 0, 0], fl[:, :, 0, wc - 1], fl[:, :, hc - 1, 0], fl[:, :, hc - 1, wc - 1]]
        return torch.stack(pts, 1)                                  # (B,4,2) in the corner order of `corners`

    Hfc, Hbc = U.DLT(Bc)(sc, sc + corner_offsets(wfc)), U.DLT(Bc)(sc, sc + corner_offsets(wbc))
    gc = U.get_grid(Bc, hc, wc, 0)
    ffc, _ = U.get_flow(Hfc.view(Bc, 1, 3, 3), gc, hc, wc, 1)
    fbc, _ = U.get_flow(Hbc.view(Bc, 1, 3, 3), gc, hc, wc, 1)
    w2c, w1c = U.get_warp_flow(c2, ffc), U.get_warp_flow(c1, fbc)
    mfc, mbc = F.create_border_mask(ffc).unsqueeze(1), F.create_border_mask(fbc).unsqueeze(1)
    l1c = L.LossL1(reduction="mean")
    lossc = l1c(mfc * c1, mfc * w2c) + l1c(mbc * c2, mbc * w1c)
    lossc.backward()
    save("pipeline_basis", img1=c1, img2=c2, w_f=wfc, w_b=wbc, Hf=Hfc, Hb=Hbc, flow_f=ffc, flow_b=fbc, w2=w2c, w1=w1c,
         loss=lossc, g_img1=c1.grad, g_img2=c2.grad, g_wf=wfc.grad, g_wb=wbc.grad, hw=np.array([hc, wc]))

    # ---- section 8f rows 2-4: uint8 pair format, loader GT flow, flow upsample ------------------------
    import types
    DL = r.data_loader
    rs = np.random.default_rng(61)
    Bp, Hp_, Wp_ = 3, 40, 56
    img12 = rs.integers(0, 256, size=(Bp, 6, Hp_, Wp_), dtype=np.uint8)
    starts = np.array([[4, 3], [0, 0], [24, 8]], dtype=np.int32)       # [x, y]
    crop = (32, 32)
    me = types.SimpleNamespace(mean_I=np.array([118.93, 113.97, 102.60]).reshape(1, 1, 3),
                               std_I=np.array([69.85, 68.81, 72.45]).reshape(1, 1, 3), crop_size=crop, rho=0)
    fulls, patches, rgbs = [], [], []
    for b in range(Bp):
        hwc = img12[b].transpose(1, 2, 0)
        i1, i2 = hwc[..., :3], hwc[..., 3:]
        rgbs.append(torch.cat((torch.Tensor(i1), torch.Tensor(i2)), dim=-1).permute(2, 0, 1).float() / 255.)  # data_loader.py:144-145
        o = DL.DGMTrainData.data_aug(me, i1, i2, np.eye(3), np.eye(3), start=[int(starts[b, 0]), int(starts[b, 1])])
        fulls.append(torch.cat((o[0], o[1]), dim=2).permute(2, 0, 1).float())
        patches.append(torch.cat((o[2], o[3]), dim=2).permute(2, 0, 1).float())
    save("pairs_u8", img12=img12, start=starts, crop=np.array(crop), gray_full=torch.stack(fulls), gray_patch=torch.stack(patches),
         rgb_full=torch.stack(rgbs))

    Hg = np.stack([np.eye(3) + rs.normal(size=(3, 3)) * np.array([[2e-2, 2e-2, 3.0], [2e-2, 2e-2, 3.0], [1e-4, 1e-4, 0.0]])
                   for _ in range(3)])
    gtf = torch.cat([DL.homo_convert_to_flow(Hg[b], (40, 56)) for b in range(3)], 0)
    save("gt_flow", H=Hg, flow=gtf, H_scaled=np.stack([DL.homo_scale(360, 640, Hg[b], 40, 56) for b in range(3)]),
         hw=np.array([40, 56]))

    flo = torch.randn(2, 2, 10, 18, generator=g(62)) * 3
    ups = {}
    for name, (ho, wo, rate, al) in {"x4_rate": (40, 72, True, True), "odd_rate": (23, 31, True, True),
                                     "down": (5, 9, False, True), "x2_norate": (20, 36, False, True)}.items():
        ups[name] = U.upsample2d_flow_as(flo.clone(), torch.zeros(1, 1, ho, wo), if_rate=rate, align_corners=al)
    save("upsample", flow=flo, **ups)


if __name__ == "__main__":
    main()
