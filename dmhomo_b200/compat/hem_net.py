"""Drop-in for the warp-path pieces of HEM/model/net.py."""
import math

import torch

from .. import ops
from .hem_utils import gen_basis  # noqa: F401  (net.py:118-154 duplicates utils.gen_basis)

__all__ = ["DLT_solve", "gen_basis", "basis_flow", "basis_homography"]


def DLT_solve(src_p, off_set):
    """HEM/model/net.py:24-92: (B, 2*(d+1)^2) mesh points + offsets -> (B, d*d, 3, 3); for the usual
    (B,8) 4-point input the corners are taken in the order p0,p1,p3,p2 and the result is (B,1,3,3)."""
    B, L = src_p.shape[:2]
    d = int(math.sqrt(L / 2) - 1)
    s = src_p.reshape(B, d + 1, d + 1, 2)
    o = off_set.reshape(B, d + 1, d + 1, 2)

    def cells(m):
        return torch.stack([m[:, :-1, :-1], m[:, :-1, 1:], m[:, 1:, 1:], m[:, 1:, :-1]], 3).reshape(B, d * d, 4, 2)

    src = cells(s)
    dst = src + cells(o)
    return ops.dlt4(src.reshape(-1, 4, 2), dst.reshape(-1, 4, 2)).view(B, d * d, 3, 3)


def basis_flow(basis, weight, h, w):
    """(basis * weight).sum(1).reshape(bs, 2, h, w)  (HEM/model/net.py:808-809, 814-815)."""
    return ops.basis_combine(basis, weight, h, w)


def basis_homography(basis, h, w, *weights):
    """8 basis weights -> basis flow at the 4 image corners -> 4-point DLT -> H (B,3,3), for up to four
    weight sets in one launch: the "8-coefficient basis-flow prediction -> 8x8 DLT solve" entry of the path
    (net.py:808-815 sampled at the corners, then utils.py:55-101)."""
    return ops.basis_homography(basis, h, w, *weights)
