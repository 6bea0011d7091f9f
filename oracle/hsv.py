"""Restatement of ``matplotlib.colors.hsv_to_rgb`` (TEST INFRASTRUCTURE ONLY).

The reference calls it at DGM/denoising_diffusion_models/denoising_diffusion_pytorch.py:22,1485.
matplotlib is a third-party dependency that is neither vendored in /root/reference
nor pinned by it (no requirements file) and is not installed in this image, so this
follows the published algorithm (the classic sector formula):

    i = int(h*6); f = h*6 - i; p = v(1-s); q = v(1-s f); t = v(1-s(1-f))
    (r,g,b) = (v,t,p),(q,v,p),(p,v,t),(p,q,v),(t,p,v),(v,p,q)  for i%6 = 0..5
    s == 0 -> (v,v,v)

**Parity unpinned**: no reference test or fixture pins this function.
"""
import numpy as np


def hsv_to_rgb(hsv):
    hsv = np.asarray(hsv)
    if hsv.shape[-1] != 3:
        raise ValueError("last dimension of input array must be 3")
    if hsv.dtype.kind != "f":
        hsv = hsv.astype(np.float32)
    ft = hsv.dtype.type
    h, s, v = hsv[..., 0], hsv[..., 1], hsv[..., 2]
    one = ft(1.0)
    h6 = h * ft(6.0)
    i = h6.astype(np.int32)
    f = h6 - i.astype(hsv.dtype)
    p = v * (one - s)
    q = v * (one - s * f)
    t = v * (one - s * (one - f))
    i = i % 6
    r = np.choose(i, [v, q, p, p, t, v])
    g = np.choose(i, [t, v, v, q, p, p])
    b = np.choose(i, [p, p, t, v, v, q])
    grey = s == 0
    r = np.where(grey, v, r)
    g = np.where(grey, v, g)
    b = np.where(grey, v, b)
    return np.stack([r, g, b], axis=-1)
