"""Drop-in replacements for the warp / flow helpers of HEM/model/utils.py (same names,
argument meaning and error behaviour), backed by libdmhomo.  Every tensor op that was a
chain of ATen launches in the reference is one hand-written kernel here; what remains in
torch is host-side plumbing (grid tensors callers ask for, index reshuffles, init-time basis
generation, F.interpolate)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops

__all__ = ["DLT", "WarpMat", "RescaleH", "Transform", "WarpImages", "CropPatchFromFull", "get_grid", "get_src_p", "get_point_pairs",
           "DLT_solve", "get_flow", "transformer", "get_warp_flow", "upsample2d_flow_as", "gen_basis"]


def RescaleH(H, rescale_size, patch_size):
    """HEM/model/utils.py:10-16 (3x3 conjugation; host-side, tiny)."""
    B = H.size()[0]
    M = torch.tensor([[patch_size[0] / rescale_size[0], 0, 0], [0, patch_size[1] / rescale_size[1], 0], [0, 0, 1]],
                     dtype=torch.float32, device=H.device).view(1, 3, 3).repeat(B, 1, 1)
    return torch.matmul(torch.matmul(torch.inverse(M), H), M)


class DLT(nn.Module):
    """HEM/model/utils.py:55-101.  forward(src_pt, dst_pt, method='Axb') -> (B,3,3)."""

    def __init__(self, batch_size, nums_pt=4):
        super().__init__()
        self.batch_size = batch_size
        self.nums_pt = nums_pt

    def forward(self, src_pt, dst_pt, method="Axb"):
        assert method in ["Ax0", "Axb"]
        self.batch_size, self.nums_pt = src_pt.shape[0], src_pt.shape[1]
        if method == "Ax0":
            raise NotImplementedError("DLT(method='Ax0'): the reference's SVD branch is numerically unusable on "
                                      "raw pixel coordinates (SURVEY.md App. D.5) and has no caller; use 'Axb'")
        if self.nums_pt != 4:
            raise ValueError("DLT: the 'Axb' branch inverts an 8x8 system, i.e. exactly 4 points")
        return ops.dlt4(src_pt, dst_pt).view(self.batch_size, 3, 3)


def WarpMat(offset, ori_size, patch_size):
    """HEM/model/utils.py:19-43.  Like the reference, `offset` is scaled in place to the original
    resolution and scaled back before returning (callers may hold a view of it)."""
    B = offset.size()[0]
    offset = offset.contiguous().view(B, -1, 2)
    offset[:, :, 0] = offset[:, :, 0] * (ori_size[0] / patch_size[0])
    offset[:, :, 1] = offset[:, :, 1] * (ori_size[1] / patch_size[1])
    src_pt = torch.tensor([[0, 0], [ori_size[0] - 1, 0], [0, ori_size[1] - 1], [ori_size[0] - 1, ori_size[1] - 1]],
                          dtype=torch.float32, device=offset.device).view(1, 4, 2).repeat(B, 1, 1)
    H = DLT(B, nums_pt=4)(src_pt=src_pt, dst_pt=src_pt + offset)
    offset[:, :, 0] = offset[:, :, 0] * (patch_size[0] / ori_size[0])
    offset[:, :, 1] = offset[:, :, 1] * (patch_size[1] / ori_size[1])
    return H


def WarpImages(input_map, H, start, patch_size):
    """HEM/model/utils.py:104-197 (S1b sampler).  patch_size = (w, h); start (B,2[,1,1]).
    Returns (warped (B,C,ph,pw), flow (B,ph,pw,2))."""
    pw, ph = patch_size
    out, flow = ops.warp(input_map, H, kind=ops.PARAM_HOMOGRAPHY, sampler=ops.S1B, out_hw=(ph, pw), start=start,
                         return_flow=True)
    return out, flow.permute(0, 2, 3, 1).contiguous()


def Transform(H, input_map, start, patch_size, start_zero=False):
    """HEM/model/utils.py:46-52."""
    if start_zero:
        start = torch.zeros_like(start)
    return WarpImages(input_map, H, start, patch_size)


def CropPatchFromFull(patch_size, full_img, start, rescale=False):
    """HEM/model/utils.py:200-291.  patch_size = (w, h); start (B,1,2) = per-sample (x, y) origin.
    rescale=True: bilinear interpolation at grid + start with the coordinate clamped to the image first - exactly the
    S1B sampler under an identity homography (one kernel).  rescale=False: an integer gather of the window (pure data
    movement, one torch indexing op)."""
    pw, ph = patch_size
    B, C, H, W = full_img.shape
    st = start.reshape(B, 2)
    if rescale:
        eye = torch.eye(3, device=full_img.device, dtype=torch.float32).repeat(B, 1, 1)
        return ops.warp(full_img, eye, kind=ops.PARAM_HOMOGRAPHY, sampler=ops.S1B, out_hw=(ph, pw), start=st.float())
    xs = st[:, 0].long().view(B, 1, 1) + torch.arange(pw, device=full_img.device).view(1, 1, pw)
    ys = st[:, 1].long().view(B, 1, 1) + torch.arange(ph, device=full_img.device).view(1, ph, 1)
    idx = (ys * W + xs).view(B, 1, ph * pw).expand(B, C, ph * pw)
    return torch.gather(full_img.reshape(B, C, H * W), 2, idx).view(B, C, ph, pw).contiguous()


def get_grid(batch_size, H, W, start=0):
    """HEM/model/utils.py:586-602: (B,3,H,W) grid of (x,y,1) + start, on start's device (CPU for an
    int start, as in the reference).  The kernels never need this tensor - they derive pixel
    coordinates from thread indices - it exists for callers that consume the grid itself."""
    dev = start.device if torch.is_tensor(start) else None
    xs = torch.arange(W, dtype=torch.float32, device=dev).view(1, 1, 1, W).expand(batch_size, 1, H, W)
    ys = torch.arange(H, dtype=torch.float32, device=dev).view(1, 1, H, 1).expand(batch_size, 1, H, W)
    grid = torch.cat([xs, ys, torch.ones(batch_size, 1, H, W, device=dev)], 1)
    grid[:, :2] = grid[:, :2] + start
    return grid


def get_src_p(batch_size, patch_size_h, patch_size_w, divides, axis_t=False):
    """HEM/model/utils.py:314-338: (B,2|3,d+1,d+1) mesh, last row/col at size-1; on CUDA when
    available (as the reference)."""
    gh, gw = patch_size_h // divides, patch_size_w // divides
    m = divides + 1
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    xx = (torch.arange(m, device=dev).view(1, m).repeat(m, 1) * gw)
    yy = (torch.arange(m, device=dev).view(m, 1).repeat(1, m) * gh)
    xx[:, -1] -= 1
    yy[-1, :] -= 1
    chans = [xx, yy] + ([torch.ones_like(xx)] if axis_t else [])
    return torch.stack(chans, 0).unsqueeze(0).repeat(batch_size, 1, 1, 1).float()


def get_point_pairs(src_p, divide):
    """HEM/model/utils.py:350-357: (B,2,d+1,d+1) -> (B,d*d,4,2), corners TL,TR,BL,BR per cell."""
    tl = src_p[:, :, :-1, :-1]
    tr = src_p[:, :, :-1, 1:]
    bl = src_p[:, :, 1:, :-1]
    br = src_p[:, :, 1:, 1:]
    B = src_p.shape[0]
    cells = torch.stack([tl, tr, bl, br], -1)  # (B,2,d,d,4)
    return cells.permute(0, 2, 3, 4, 1).reshape(B, divide * divide, 4, 2).contiguous()


def DLT_solve(src_p, off_set):
    """HEM/model/utils.py:360-397 (mesh variant): (B,2,d+1,d+1) x2 -> (B,d*d,3,3)."""
    B, _, m = src_p.shape[:3]
    d = m - 1
    src = get_point_pairs(src_p, d)
    dst = src + get_point_pairs(off_set, d)
    return ops.dlt4(src.reshape(-1, 4, 2), dst.reshape(-1, 4, 2)).view(B, d * d, 3, 3)


def get_flow(H_mat_mul, patch_indices, patch_size_h, patch_size_w, divide, point_use=False):
    """HEM/model/utils.py:400-440 -> (flow (B,2,h,w), vgrid (B,2,h,w)).  The pixel grid is
    regenerated in-kernel; only its origin (patch_indices[:, :2, 0, 0] = `start`) is read."""
    if point_use:
        raise NotImplementedError("get_flow(point_use=True) has no caller in the reference")
    vgrid = patch_indices[:, :2, ...]
    start = patch_indices[:, :2, 0, 0].to(H_mat_mul.device)
    flow = ops.homography_to_flow(H_mat_mul, patch_size_h, patch_size_w, divide=divide, start=start)
    return flow, vgrid


def transformer(I, vgrid, train=True):
    """HEM/model/utils.py:443-545: S1 bilinear gather at absolute coordinates vgrid (B,2,h,w)."""
    out = ops.warp(I, vgrid, kind=ops.PARAM_COORDS, sampler=ops.S1)
    if not train:
        out = out.permute(0, 2, 3, 1)
    return out


def get_warp_flow(img, flow, start=0):
    """HEM/model/utils.py:548-553."""
    return ops.warp(img, flow, kind=ops.PARAM_FLOW, sampler=ops.S1, start=start)


def upsample2d_flow_as(inputs, target_as, mode="bilinear", if_rate=False, align_corners=True):
    """HEM/model/utils.py:556-572 (scales `inputs` in place when if_rate, as the reference: callers see it)."""
    _, _, h, w = target_as.size()
    if if_rate:
        _, _, h_, w_ = inputs.size()
        inputs[:, 0, :, :] *= (w / w_)
        inputs[:, 1, :, :] *= (h / h_)
    if mode != "bilinear" or inputs.shape[1] != 2:
        # nearest / other channel counts: not on the hot path (no caller in the reference's active configuration)
        if mode == "nearest":
            return F.interpolate(inputs, [h, w], mode=mode)
        return F.interpolate(inputs, [h, w], mode=mode, align_corners=align_corners)
    return ops.flow_upsample(inputs, (h, w), if_rate=False, align_corners=align_corners)


def gen_basis(h, w, is_qr=True, is_scale=True):
    """HEM/model/utils.py:605-640: the 8 flow bases (8,2,h,w).  Init-time, on the CPU, through
    torch.qr exactly like the reference: the fp32 QR output is *data* the kernels consume
    (it deviates from the ideal polynomial span by ~1e-2), not a formula."""
    xs = torch.arange(w, dtype=torch.float32).view(1, w).expand(h, w)
    ys = torch.arange(h, dtype=torch.float32).view(h, 1).expand(h, w)
    z, o = torch.zeros(h, w), torch.ones(h, w)
    pairs = [(xs, z), (ys, z), (o, z), (z, xs), (z, ys), (z, o), (xs * xs, xs * ys), (xs * ys, ys * ys)]
    flows = torch.stack([torch.stack(p, -1) for p in pairs], 0)
    if is_qr:
        q, _ = torch.qr(flows.reshape(8, -1).t().contiguous())
        flows = q.t().reshape(8, h, w, 2).contiguous()
    if is_scale:
        flows = flows / flows.abs().reshape(8, -1).max(1)[0].reshape(8, 1, 1, 1)
    return flows.permute(0, 3, 1, 2).contiguous()
