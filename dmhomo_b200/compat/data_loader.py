"""Drop-in names of HEM/dataset/data_loader.py on the warp path: the
ground-truth flow generation and the uint8 pair format, batched on the GPU."""
import numpy as np
import torch

from .. import ops
from .dgm import flow_warp, mesh_grid, norm_grid  # noqa: F401  (data_loader.py:65-94 == ddpm.py:1262-1280)

__all__ = ["homo_scale", "homo_convert_to_flow", "flow_warp", "mesh_grid", "norm_grid", "pairs_to_batch", "MEAN_I", "STD_I"]

MEAN_I, STD_I = ops.MEAN_I, ops.STD_I


def homo_scale(h0, w0, H, h1, w1):
    """data_loader.py:29-39: conjugate H from an (h0, w0) image to an (h1, w1) image through the normalised frame.
    Host fp64 3x3 algebra (a per-sample constant, not per-pixel work); accepts (3,3) or (B,3,3)."""
    H = np.asarray(H, dtype=np.float64)
    M0 = np.array([[w0 / 2.0, 0.0, w0 / 2.0], [0.0, h0 / 2.0, h0 / 2.0], [0.0, 0.0, 1.0]])
    M1 = np.array([[w1 / 2.0, 0.0, w1 / 2.0], [0.0, h1 / 2.0, h1 / 2.0], [0.0, 0.0, 1.0]])
    Hn = np.matmul(np.matmul(np.linalg.inv(M0), H), M0)
    return np.matmul(np.matmul(M1, Hn), np.linalg.inv(M1))


def homo_convert_to_flow(H, size=(360, 640), device="cuda"):
    """data_loader.py:42-52: ground-truth flow (B,2,h,w) of homographies H ((3,3) or (B,3,3), fp64): the fp64
    mapping (+1e-8), rounded to fp32, minus the fp32 grid.  The reference returns a CPU tensor of batch 1 per call;
    here a whole batch stays on the GPU."""
    Ht = torch.as_tensor(np.asarray(H, dtype=np.float64) if not torch.is_tensor(H) else H, dtype=torch.float64)
    Ht = Ht.reshape(-1, 3, 3).to(device)
    return ops.homography_to_flow_f64(Ht, int(size[0]), int(size[1]), eps=1e-8, channels_last=False, as_mapping=2)


def pairs_to_batch(img12, start, crop_size, device="cuda"):
    """DGMTrainData.__getitem__ + data_aug (data_loader.py:121-146, 217-255) for a batch of on-disk pairs
    `img12` uint8 (B,6,H,W) (numpy or tensor): returns the dict entries the network consumes, computed on the GPU."""
    x = torch.as_tensor(img12).to(device)
    full, patch, rgb = ops.pairs_u8_to_gray(x, start=start, patch_size=crop_size)
    st = torch.as_tensor(start, dtype=torch.float32, device=device).reshape(-1, 2, 1, 1)
    return {"imgs_gray_full": full, "imgs_gray_patch": patch, "imgs_rgb_full": rgb, "start": st}
