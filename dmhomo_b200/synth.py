"""Seeded synthetic inputs of BASELINE.json's configs.

Small configs are generated on the CPU with torch.Generator(230 + rank) so the CPU oracle
and the CUDA path see identical bits; callers move them to the device.
"""
import math

import torch

SEED = 230  # the reference's seed (hem_evaluate.py:204, HEM/train.py:42)

CONFIGS = {
    # name: B, C, h, w, rho, what
    "cfg1": dict(B=16, C=1, h=360, w=640, rho=32.0),
    "cfg2": dict(B=64, C=1, h=320, w=576, rho=32.0),
    "cfg3": dict(B=25, C=3, h=256, w=256, rho=16.0),
    "cfg4": dict(B=4096, C=3, h=512, w=512, rho=32.0),
    "cfg5": dict(B=8192, C=3, h=1080, w=1920, rho=64.0),
}


def generator(rank=0, device="cpu"):
    return torch.Generator(device=device).manual_seed(SEED + rank)


def noise_images(B, C, h, w, gen):
    """U[0,1) images: worst-case gradients (stage-isolated parity)."""
    return torch.rand(B, C, h, w, generator=gen, device=gen.device)


def smooth_images(B, C, h, w, gen):
    """Sum of 4 random sinusoids with periods >= 16 px, scaled to [0,1] (chained parity)."""
    dev = gen.device
    ys = torch.arange(h, dtype=torch.float32, device=dev).view(1, 1, h, 1)
    xs = torch.arange(w, dtype=torch.float32, device=dev).view(1, 1, 1, w)
    img = torch.zeros(B, C, h, w, device=dev)
    for _ in range(4):
        period = 16.0 + 48.0 * torch.rand(B, C, 1, 1, generator=gen, device=dev)
        theta = 2 * math.pi * torch.rand(B, C, 1, 1, generator=gen, device=dev)
        phase = 2 * math.pi * torch.rand(B, C, 1, 1, generator=gen, device=dev)
        k = 2 * math.pi / period
        img = img + torch.sin(k * (xs * torch.cos(theta) + ys * torch.sin(theta)) + phase)
    return (img / 8.0 + 0.5).contiguous()


def corner_points(B, h, w, device="cpu"):
    c = torch.tensor([[0, 0], [w - 1, 0], [0, h - 1], [w - 1, h - 1]], dtype=torch.float32, device=device)
    return c.view(1, 4, 2).repeat(B, 1, 1)


def corner_offsets(B, rho, gen):
    return (torch.rand(B, 4, 2, generator=gen, device=gen.device) * 2 - 1) * rho


def basis_weights(B, gen, scale=4.0):
    return (torch.rand(B, 8, 1, generator=gen, device=gen.device) * 2 - 1) * scale


def homographies_360x640(B, gen, rho=32.0):
    """Random 4-pt homographies at 360x640 as float64 numpy (cfg 3 input), solved in fp64."""
    src = corner_points(B, 360, 640).double()
    dst = src + corner_offsets(B, rho, gen).double().cpu()
    x, y, u, v = src[..., 0], src[..., 1], dst[..., 0], dst[..., 1]
    one, zero = torch.ones_like(x), torch.zeros_like(x)
    top = torch.stack([x, y, one, zero, zero, zero, -u * x, -u * y], -1)
    bot = torch.stack([zero, zero, zero, x, y, one, -v * x, -v * y], -1)
    A = torch.stack([top, bot], 2).reshape(B, 8, 8)
    h8 = torch.linalg.solve(A, dst.reshape(B, 8, 1)).reshape(B, 8)
    return torch.cat([h8, torch.ones(B, 1, dtype=torch.float64)], 1).reshape(B, 3, 3).numpy()
