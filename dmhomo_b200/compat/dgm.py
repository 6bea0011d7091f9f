"""Drop-in for the warp / flow helpers inside
DGM/denoising_diffusion_models/denoising_diffusion_pytorch.py (and the photometric term of
classifier_free_guidance.py)."""
import numpy as np
import torch

from .. import ops

__all__ = ["mesh_grid", "norm_grid", "flow_warp", "homo_to_flow", "adapt_homography_to_preprocessing_v3",
           "flow_to_image", "visulize_flow", "postProcess", "postProcess_cv2", "homo_gen", "photo_loss", "resize_flow", "warp_pairs_u8", "split_sample_batches"]


def mesh_grid(B, H, W):
    """ddpm.py:1283-1289 (host-side helper; kernels do not need it)."""
    xs = torch.arange(0, W).repeat(B, H, 1)
    ys = torch.arange(0, H).repeat(B, W, 1).transpose(1, 2)
    return torch.stack([xs, ys], 1)


def norm_grid(v_grid):
    """ddpm.py:1292-1299."""
    _, _, H, W = v_grid.size()
    out = torch.zeros_like(v_grid)
    out[:, 0] = 2.0 * v_grid[:, 0] / (W - 1) - 1.0
    out[:, 1] = 2.0 * v_grid[:, 1] / (H - 1) - 1.0
    return out.permute(0, 2, 3, 1)


def flow_warp(x, flow12, pad="border", mode="bilinear"):
    """ddpm.py:1262-1280 == HEM/dataset/data_loader.py:84-94 (S3 sampler)."""
    if mode != "bilinear":
        raise NotImplementedError("flow_warp: only mode='bilinear' occurs in the reference")
    if pad not in ("border", "zeros"):
        raise NotImplementedError(f"flow_warp: pad={pad!r} not supported")
    return ops.warp(x, flow12, kind=ops.PARAM_FLOW, sampler=ops.S3_BORDER if pad == "border" else ops.S2_ZEROS)


def homo_to_flow(homo, H=600, W=800):
    """ddpm.py:972-975: one homography (numpy, any shape holding 9 values) -> (H,W,2) fp32 numpy."""
    Ht = torch.as_tensor(np.asarray(homo, dtype=np.float64).reshape(1, 3, 3), device="cuda")
    return ops.homography_to_flow_f64(Ht, H, W, eps=1e-6, channels_last=True)[0].cpu().numpy()


def adapt_homography_to_preprocessing_v3(h0, w0, H, h1, w1):
    """ddpm.py:978-988 == HEM/dataset/data_loader.py:29-39 (3x3 fp64 host arithmetic)."""
    M0 = np.array([[w0 / 2.0, 0.0, w0 / 2.0], [0.0, h0 / 2.0, h0 / 2.0], [0.0, 0.0, 1.0]])
    M1 = np.array([[w1 / 2.0, 0.0, w1 / 2.0], [0.0, h1 / 2.0, h1 / 2.0], [0.0, 0.0, 1.0]])
    Hn = np.matmul(np.matmul(np.linalg.inv(M0), H), M0)
    return np.matmul(np.matmul(M1, Hn), np.linalg.inv(M1))


def flow_to_image(flow, max_flow=256):
    """ddpm.py:1471-1486: (H,W,2) numpy -> (H,W,3) numpy in [0,1]."""
    if max_flow is None:
        max_flow = float(np.max(flow))
    f = torch.as_tensor(np.ascontiguousarray(flow, dtype=np.float32), device="cuda").unsqueeze(0)
    return ops.flow_to_rgb(f, max_flow, in_channels_last=True, out_channels_last=True)[0].cpu().numpy()


def visulize_flow(all_images):
    """ddpm.py:1489-1502: (B,2,H,W) -> (B,3,H,W) RGB flow.  Stays on the device (the reference
    round-trips through numpy per sample and returns a CPU tensor)."""
    return ops.flow_to_rgb(all_images.detach(), 256.0)


def postProcess(torch_tensor, mask, flows):
    """ddpm.py:1505-1517."""
    img1s, img2s = torch_tensor[:, :3], torch_tensor[:, 3:6]
    warp_img2s = flow_warp(img2s, flows)
    flows_vis = visulize_flow(flows)
    mask_rgb = mask.repeat(1, 3, 1, 1)
    buf1 = torch.concat([img1s, img1s, mask_rgb, flows_vis], -1)
    buf2 = torch.concat([img2s, warp_img2s, mask_rgb, flows_vis], -1)
    return buf1, buf2


def postProcess_cv2(imgs, homos, rank):
    """ddpm.py:1520-1540: uint8 numpy (B,6,H,W) + numpy homographies (B,3,3) -> device buffers;
    img1 is warped by cv2.warpPerspective semantics to (256,256) as in the reference."""
    dev = torch.device("cuda", rank) if isinstance(rank, int) else torch.device(rank)
    t = torch.as_tensor(imgs, device=dev).float() / 255.0
    img1s, img2s = t[:, :3], t[:, 3:6]
    Hs = torch.as_tensor(np.asarray(homos, dtype=np.float64).reshape(-1, 3, 3), device=dev)
    warp_img2s = ops.warp_perspective(img1s, Hs, (256, 256))
    return torch.concat([img1s, warp_img2s], -1), torch.concat([img2s, img2s], -1)


def homo_gen(flow):
    """ddpm.py:1647-1661: least-squares homography of a dense flow, (B,1,3,3) float64."""
    return ops.flow_to_homography_ls(flow)


def photo_loss(im1, im2, flow, mask, alpha_bar_t, fused=True):
    """classifier_free_guidance.py:784, 799-806:
    mean_b( alpha_bar_t[b] * mean_chw( mask * |flow_warp(im2, flow) - im1| ) ), one fused launch."""
    term = ops.WarpTerm(im2, im1, flow, soft_mask=mask, sample_weight=alpha_bar_t.reshape(-1))
    return ops.warp_loss([term], kind=ops.PARAM_FLOW, sampler=ops.S3_BORDER, loss_form=ops.LOSS_DIFF_MASKED,
                         border_mask=False, fused=fused)


def resize_flow(flow, size):
    """ddpm.py:1249-1259: cv2.resize (bilinear, half-pixel centres) of an (h,w,2) numpy flow to (size,size,2) with the
    two channels scaled by size / w and size / h."""
    f = torch.as_tensor(np.ascontiguousarray(flow, dtype=np.float32), device="cuda").permute(2, 0, 1).unsqueeze(0)
    out = ops.flow_upsample(f.contiguous(), (int(size), int(size)), if_rate=True, align_corners=False)
    return out[0].permute(1, 2, 0).contiguous().cpu().numpy()


def warp_pairs_u8(imgs, homos):
    """The check DGM/generate_nyps_to_single_case.py:10-21 runs on the generated {"imgs": (N,6,H,W) uint8, "homos": (N,3,3)}
    sample batches (DGM/dgm_sample.py:62-76): cv2.warpPerspective(img1, homo12, (W,H)) on the uint8 image, batched on the
    GPU, bit-identical to OpenCV's fixed-point path.  numpy in, numpy (N,3,H,W) uint8 out."""
    x = torch.as_tensor(np.ascontiguousarray(imgs[:, :3]), device="cuda")
    H = torch.as_tensor(np.asarray(homos, dtype=np.float64).reshape(-1, 3, 3), device="cuda")
    N, _, h, w = x.shape
    return ops.warp_perspective(x, H, (w, h)).cpu().numpy()


def split_sample_batches(buf):
    """generate_nyps_to_single_case.py:29-47: the list of {"imgs", "homos"} batches a sampling run saved ->
    one {"img12": (6,H,W) uint8, "homo12": (3,3)} dict per sample, the on-disk pair format the HEM loader reads
    (HEM/dataset/data_loader.py:121-146).  Host-side bookkeeping, no arithmetic."""
    out = []
    for item in buf:
        imgs, homos = item["imgs"], item["homos"]
        for i in range(len(imgs)):
            out.append({"img12": imgs[i], "homo12": homos[i]})
    return out


def render_conditions(img2s, homos, max_flow=256):
    """The condition images a DGM sampling step derives from (img2s, condition homographies): postProcess_cv2's
    warpPerspective (ddpm.py:1520-1529), homo_to_flow + visulize_flow (ddpm.py:972-975, 1489-1502) and postProcess's
    flow_warp (ddpm.py:1505-1518), as ONE call whose independent launches overlap (ops.render_conditions).  Tensors in,
    dict of tensors out; every entry is bit-identical to the separate call of the same name."""
    return ops.render_conditions(img2s, homos, max_flow=max_flow)

