"""Drop-in for the mask / mapping helpers of HEM/utils_operations/flow_and_mapping_operations.py
that sit on the warp path.  torch inputs run on the GPU; numpy inputs are a host convenience:
they are moved to the current CUDA device, processed by the same kernels, and returned as numpy."""
import numpy as np
import torch

from .. import ops

__all__ = ["get_gt_correspondence_mask", "create_border_mask", "define_mask_zero_borders", "convert_flow_to_mapping",
           "convert_mapping_to_flow", "from_homography_to_pixel_wise_mapping", "normalize", "unnormalize",
           "unormalise_flow_or_mapping", "unormalise_and_convert_mapping_to_flow"]


def _bchw(flow):
    """-> (tensor (B,2,H,W), was_3d).  Accepts channel-first or channel-last, 3-D or 4-D."""
    squeeze = flow.dim() == 3
    f = flow.unsqueeze(0) if squeeze else flow
    if f.shape[1] != 2:
        f = f.permute(0, 3, 1, 2)
    return f, squeeze


def get_gt_correspondence_mask(flow):
    """...operations.py:45-71 (torch branch): bool mask 0 <= x' <= W and 0 <= y' <= H."""
    if isinstance(flow, np.ndarray):
        raise NotImplementedError("numpy branch of get_gt_correspondence_mask: the reference crashes on "
                                  "NumPy >= 1.24 (np.bool, SURVEY.md App. D.2); pass a torch tensor")
    f, squeeze = _bchw(flow)
    m = ops.border_mask(f)
    return m[0] if squeeze else m


def create_border_mask(flow):
    """...operations.py:40-42."""
    f, squeeze = _bchw(flow)
    m = ops.border_mask(f, as_float=True)
    return m[0] if squeeze else m


def define_mask_zero_borders(image, epsilon=1e-6):
    """...operations.py:6-37 (torch branch)."""
    if isinstance(image, np.ndarray):
        raise NotImplementedError("numpy branch of define_mask_zero_borders crashes in the reference on "
                                  "NumPy >= 1.24 (np.bool); pass a torch tensor")
    squeeze = image.dim() == 3
    im = image.unsqueeze(0) if squeeze else image
    if im.shape[1] != 3:
        im = im.permute(0, 3, 1, 2)
    m = ops.zero_border_mask(im, epsilon)
    return m[0] if squeeze else m


def _grid_like(f):
    B, _, H, W = f.shape
    xs = torch.arange(W, dtype=torch.float32, device=f.device).view(1, 1, 1, W).expand(B, 1, H, W)
    ys = torch.arange(H, dtype=torch.float32, device=f.device).view(1, 1, H, 1).expand(B, 1, H, W)
    return torch.cat([xs, ys], 1)


def _convert(t, sign, output_channel_first):
    is_np = isinstance(t, np.ndarray)
    x = torch.from_numpy(np.ascontiguousarray(t)) if is_np else t
    squeeze = x.dim() == 3
    f = x.unsqueeze(0) if squeeze else x
    last = 2 if is_np else 1  # numpy inputs are detected channel-last by shape[3] != 2 in the reference
    if is_np:
        if f.shape[3] != 2:
            f = f.permute(0, 2, 3, 1)
        f = f.permute(0, 3, 1, 2)
    elif f.shape[1] != 2:
        f = f.permute(0, 3, 1, 2)
    out = (f.float() + sign * _grid_like(f)).float()
    if not output_channel_first:
        out = out.permute(0, 2, 3, 1)
    out = out[0] if squeeze else out
    return out.numpy().astype(np.float32) if is_np else out


def convert_flow_to_mapping(flow, output_channel_first=True):
    """...operations.py:84-152: mapping = flow + pixel grid (one elementwise add; stays in torch)."""
    return _convert(flow, 1.0, output_channel_first)


def convert_mapping_to_flow(mapping, output_channel_first=True):
    """...operations.py:155-224."""
    return _convert(mapping, -1.0, output_channel_first)


def from_homography_to_pixel_wise_mapping(shape, H):
    """...operations.py:454-484: numpy in, numpy out ((h,w) fp32 map_x, map_y); fp64 math on the GPU."""
    h, w = shape[:2]
    Ht = torch.as_tensor(np.asarray(H, dtype=np.float64).reshape(1, 3, 3), device="cuda")
    m = ops.homography_to_flow_f64(Ht, h, w, eps=1e-8, channels_last=False, as_mapping=True)[0].cpu().numpy()
    return m[0], m[1]


def _grid_op(t, mode, output_channel_first):
    """Layout handling of the reference's torch branches (...operations.py:227-451): 3-D or 4-D, channel-first or
    channel-last in, channel-first or channel-last out; numpy inputs are moved to the GPU and come back as numpy."""
    is_np = isinstance(t, np.ndarray)
    x = torch.from_numpy(np.ascontiguousarray(t)).cuda() if is_np else t
    squeeze = x.dim() == 3
    f = x.unsqueeze(0) if squeeze else x
    if f.shape[1] != 2:
        f = f.permute(0, 3, 1, 2)
    out = ops.grid_normalize(f, mode)
    if not output_channel_first:
        out = out.permute(0, 2, 3, 1)
    out = out[0] if squeeze else out
    return out.cpu().numpy().astype(np.float32) if is_np else out


def normalize(tensor, output_channel_first=True):
    """...operations.py:419-451: pixel coordinates -> [-1, 1] (2*t/(S-1) - 1)."""
    return _grid_op(tensor, 0, output_channel_first)


def unnormalize(tensor, output_channel_first=True):
    """...operations.py:384-416: [-1, 1] -> pixel coordinates ((t+1)*(S-1)/2)."""
    return _grid_op(tensor, 1, output_channel_first)


def unormalise_flow_or_mapping(map, output_channel_first=True):
    """...operations.py:318-381."""
    return _grid_op(map, 1, output_channel_first)


def unormalise_and_convert_mapping_to_flow(map, output_channel_first=True):
    """...operations.py:227-315: un-normalise a [-1, 1] mapping and subtract the pixel grid."""
    return _grid_op(map, 2, output_channel_first)
