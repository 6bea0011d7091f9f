"""CPU model of the tile kernel's per-tile proofs (dmhomo_b200/csrc/dmh_warp_tile.cu, producer warp).

The producer classifies every 64x64 output tile from its four corners only: `full` (every tap of the tile lies in
the staged 96 x 88 source window), `interior` (no clamp, no mask, no epsilon rule can apply to any pixel) and the
preconditions of the packed division.  The consumers then run bodies that ASSUME those facts.  This test restates the
classification in float32 numpy and checks, over thousands of tiles of random and adversarial homographies, that
each flag implies what the fast bodies rely on, using the reference-exact per-pixel coordinates of the oracle
(separately rounded fp32 chain, HEM/model/utils.py:400-440).  No GPU needed: it pins the *argument*, the GPU parity
tests pin the code."""
import numpy as np
import pytest
import torch

from dmhomo_b200 import synth
from oracle import port

TW = TH = 64
BW, BH = 96, (TH * 5) // 4 + 8
f32 = np.float32


def entry_sane(v):
    z = abs(float(v))
    return z == 0.0 or (9.094947017729282e-13 <= z <= 1048576.0)


def classify(hm, tx0, ty0, h, w, Hs, Ws):
    """Producer logic, start = 0 (float32 arithmetic; the kernel contracts some of it into FMAs and uses
    rcp.approx - differences of a few ulp, far inside the margins being tested)."""
    hm = hm.astype(f32)
    Wm1, Hm1 = Ws - 1, Hs - 1
    tx1, ty1 = min(tx0 + TW, w) - 1, min(ty0 + TH, h) - 1
    ok = robust = True
    us = []
    for i in range(4):
        px, py = f32(tx1 if i & 1 else tx0), f32(ty1 if i & 2 else ty0)
        T = f32(hm[6] * px + hm[7] * py + hm[8])
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            rT = f32(1.0) / T
            ux, uy = f32((hm[0] * px + hm[1] * py + hm[2]) * rT), f32((hm[3] * px + hm[4] * py + hm[5]) * rT)
        ok = ok and bool(T > 1e-4) and bool(abs(ux) < 1e7) and bool(abs(uy) < 1e7)
        Tm = abs(hm[6] * px) + abs(hm[7] * py) + abs(hm[8])
        Nm = abs(hm[0] * px) + abs(hm[1] * py) + abs(hm[2]) + abs(hm[3] * px) + abs(hm[4] * py) + abs(hm[5])
        robust = robust and bool(T > Tm * f32(0.015625)) and bool(Nm < T * f32(1048576.0))
        us.append((ux, uy))
    if not ok:
        return dict(ok=False, full=False, interior=False, mixed=False, wx0=0, wy0=0)
    mnx, mxx = min(u[0] for u in us), max(u[0] for u in us)
    mny, mxy = min(u[1] for u in us), max(u[1] for u in us)
    wx0 = max(int(np.floor(mnx)) - 1, 0) & ~3
    wy0 = max(int(np.floor(mny)) - 1, 0)
    have = wx0 <= Wm1 and wy0 <= Hm1
    full = have and min(int(np.floor(mxx)) + 2, Wm1) <= wx0 + BW - 1 and min(int(np.floor(mxy)) + 2, Hm1) <= wy0 + BH - 1
    sane = all(entry_sane(v) for v in hm)
    mixed = full and sane and robust and tx0 + TW <= w
    interior = (mixed and ty0 + TH <= h and mnx >= 2.0 and mny >= 2.0 and mxx <= float(min(Wm1, w) - 2)
                and mxy <= float(min(Hm1, h) - 2))
    return dict(ok=True, full=full, interior=interior, mixed=mixed, wx0=wx0, wy0=wy0)


def exact_coords(H, h, w):
    """Per-pixel sampling coordinates and T exactly as the kernels (and the reference) round them."""
    flow, _ = port.homography_to_flow(H, h, w)
    grid = port.pixel_grid(H.shape[0], h, w, homogeneous=False)
    c = (grid + flow).numpy()                                    # fl(g + fl(q/T - g))
    xs, ys = grid[:, 0].numpy(), grid[:, 1].numpy()
    Hn = H.numpy().reshape(-1, 9).astype(f32)
    T = (Hn[:, 6, None, None] * xs + Hn[:, 7, None, None] * ys).astype(f32) + Hn[:, 8, None, None]
    return c[:, 0], c[:, 1], T.astype(f32)


def homographies(kind, B, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    src = port.corner_points(B, h, w)
    if kind == "mild":
        return port.dlt4(src, src + synth.corner_offsets(B, 32.0, g))
    if kind == "strong":
        return port.dlt4(src, src + synth.corner_offsets(B, 0.35 * min(h, w), g))
    if kind == "affine":   # scale / shear / shift, no perspective (tiny or zero h6, h7)
        H = torch.eye(3).repeat(B, 1, 1)
        H[:, 0, 0] = 1 + (torch.rand(B, generator=g) - 0.5) * 0.6
        H[:, 1, 1] = 1 + (torch.rand(B, generator=g) - 0.5) * 0.6
        H[:, 0, 1] = (torch.rand(B, generator=g) - 0.5) * 0.3
        H[:, 1, 0] = (torch.rand(B, generator=g) - 0.5) * 0.3
        H[:, 0, 2] = (torch.rand(B, generator=g) - 0.5) * 80
        H[:, 1, 2] = (torch.rand(B, generator=g) - 0.5) * 80
        H[:, 2, 0] = (torch.rand(B, generator=g) - 0.5) * 2e-9
        return H
    if kind == "horizon":  # T crosses zero inside or near the image
        H = port.dlt4(src, src + synth.corner_offsets(B, 8.0, g))
        H[:, 2, 0] = -1.0 / (torch.rand(B, generator=g) * 2.0 * w + 8.0)
        return H
    raise ValueError(kind)


@pytest.mark.parametrize("kind,h,w", [("mild", 320, 576), ("mild", 360, 640), ("strong", 256, 256), ("affine", 192, 320),
                                      ("horizon", 128, 256)])
def test_tile_flags_imply_what_the_fast_bodies_assume(kind, h, w):
    B = 6
    H = homographies(kind, B, h, w, seed=sum(map(ord, kind)) + h + w)
    cx, cy, T = exact_coords(H, h, w)
    Ws, Hs = w, h
    n_tiles = n_full = n_int = 0
    for b in range(B):
        hm = H[b].reshape(9).numpy()
        for ty0 in range(0, h, TH):
            for tx0 in range(0, w, TW):
                f = classify(hm, tx0, ty0, h, w, Hs, Ws)
                n_tiles += 1
                sl = (b, slice(ty0, min(ty0 + TH, h)), slice(tx0, min(tx0 + TW, w)))
                x, y, t = cx[sl], cy[sl], T[sl]
                if f["full"]:
                    n_full += 1
                    # S1 taps: floor, +1, both clamped to the source (utils.py:463-490)
                    x0 = np.clip(np.floor(x), 0, Ws - 1); x1 = np.clip(np.floor(x) + 1, 0, Ws - 1)
                    y0 = np.clip(np.floor(y), 0, Hs - 1); y1 = np.clip(np.floor(y) + 1, 0, Hs - 1)
                    assert x0.min() >= f["wx0"] and x1.max() <= f["wx0"] + BW - 1, (kind, b, tx0, ty0)
                    assert y0.min() >= f["wy0"] and y1.max() <= f["wy0"] + BH - 1, (kind, b, tx0, ty0)
                if f["interior"]:
                    n_int += 1
                    assert x.min() >= 0 and x.max() < min(Ws - 1, w) and y.min() >= 0 and y.max() < min(Hs - 1, h)
                    assert np.abs(t).min() >= 1e-4 * 0.5       # nowhere near the |T| < 1e-7 epsilon rule
                if f["mixed"] or f["interior"]:
                    assert t.min() > 0
    assert n_tiles == B * ((h + TH - 1) // TH) * ((w + TW - 1) // TW)
    if kind == "mild":
        assert n_int > 0.3 * n_tiles and n_full > 0.9 * n_tiles      # the fast bodies carry the typical workload
    if kind == "horizon":
        assert n_int < n_tiles                                        # and are refused where they must be


def test_interior_weight_identities_are_exact():
    """The interior body replaces fl(x1 - cx) by fl(1 - fl(cx - x0)) (x1 = x0 + 1, no clamp) and takes floor(cx) from the
    mantissa of fl_down(cx + 2^23).  Both are claimed bit-exact for 0 <= cx < 2^22: ax0 = cx - floor(cx) is exact
    (Sterbenz), so 1 - ax0 and x1 - cx are the same real number before the single rounding."""
    rs = np.random.default_rng(5)
    cx = np.concatenate([
        rs.uniform(0, 4096, 2_000_000), rs.uniform(0, 2, 500_000), rs.uniform(0, 1e-3, 200_000),
        np.arange(0, 2048, dtype=np.float64), np.nextafter(np.arange(1, 2048, dtype=np.float32), np.float32(0)).astype(np.float64),
        np.nextafter(np.arange(0, 2048, dtype=np.float32), np.float32(1e9)).astype(np.float64),
        rs.uniform(4096, 2 ** 22 - 1, 500_000)]).astype(f32)
    x0 = np.floor(cx)                                   # exact in fp32
    ax0 = (cx - x0).astype(f32)
    assert np.array_equal(ax0.astype(np.float64), cx.astype(np.float64) - x0.astype(np.float64)), "cx - floor(cx) must be exact"
    ref = ((x0 + f32(1)).astype(f32) - cx).astype(f32)  # the reference: x1f - cx with x1 = x0 + 1
    new = (f32(1) - ax0).astype(f32)
    assert np.array_equal(ref, new)
    # round-down add of 2^23: emulated in fp64 (exact sum) + floor to the fp32 grid of [2^23, 2^24), which is the integers
    magic = np.floor(cx.astype(np.float64) + 8388608.0)
    assert np.array_equal(magic - 8388608.0, x0.astype(np.float64))
    assert np.array_equal((magic.astype(np.int64) & 0x7FFFFF), x0.astype(np.int64))   # mantissa bits = the integer
