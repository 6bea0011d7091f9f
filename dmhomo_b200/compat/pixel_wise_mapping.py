"""Drop-in for HEM/utils_operations/pixel_wise_mapping.py."""
import numpy as np
import torch

from .. import ops

__all__ = ["warp", "warp_with_mapping", "remap_using_flow_fields", "remap_using_correspondence_map"]

_CV_INTER_LINEAR, _CV_BORDER_CONSTANT = 1, 0       # cv2.INTER_LINEAR, cv2.BORDER_CONSTANT (cv2 itself is not needed here)


def _remap_numpy(image, a, b, displacement, interpolation, border_mode):
    if interpolation != _CV_INTER_LINEAR or border_mode != _CV_BORDER_CONSTANT:
        raise NotImplementedError("remap: only INTER_LINEAR with BORDER_CONSTANT (the reference's defaults) is implemented")
    img = np.ascontiguousarray(image)
    if img.dtype not in (np.float32, np.uint8):
        raise TypeError("remap: image must be float32 or uint8")
    squeeze = img.ndim == 2
    t = torch.from_numpy(img.reshape(img.shape[0], img.shape[1], -1)).unsqueeze(0).cuda()
    # (displacements: the reference adds an fp64 grid and rounds the sum to fp32 - the kernel's fp32 add is that value)
    m = torch.from_numpy(np.stack([np.asarray(a, dtype=np.float32), np.asarray(b, dtype=np.float32)], 0)).unsqueeze(0).cuda()
    out = ops.remap(t, m, displacement=displacement, channels_last=True)[0].cpu().numpy()
    return out[..., 0] if squeeze else out


def remap_using_flow_fields(image, disp_x, disp_y, interpolation=_CV_INTER_LINEAR, border_mode=_CV_BORDER_CONSTANT):
    """pixel_wise_mapping.py:7-32: cv2.remap at grid + (disp_x, disp_y); numpy HxWxC in, numpy out (as the reference)."""
    return _remap_numpy(image, disp_x, disp_y, True, interpolation, border_mode)


def remap_using_correspondence_map(image, map_x, map_y, interpolation=_CV_INTER_LINEAR, border_mode=_CV_BORDER_CONSTANT):
    """pixel_wise_mapping.py:35-52: cv2.remap at absolute coordinates (map_x, map_y); numpy in, numpy out."""
    return _remap_numpy(image, map_x, map_y, False, interpolation, border_mode)


def _sampler(padding_mode):
    if padding_mode == "zeros":
        return ops.S2_ZEROS
    if padding_mode == "border":
        return ops.S3_BORDER
    raise NotImplementedError(f"padding_mode={padding_mode!r}: only 'zeros' and 'border' occur in the reference")


def warp(x, flo, padding_mode="zeros"):
    """pixel_wise_mapping.py:55-88: grid_sample(align_corners=True) at grid + flo (torch>=1.3 branch)."""
    return ops.warp(x, flo, kind=ops.PARAM_FLOW, sampler=_sampler(padding_mode))


def warp_with_mapping(x, vgrid):
    """pixel_wise_mapping.py:91-113: as `warp` with absolute pixel coordinates."""
    return ops.warp(x, vgrid, kind=ops.PARAM_COORDS, sampler=ops.S2_ZEROS)
