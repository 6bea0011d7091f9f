"""Drop-in for HEM/utils_operations/pixel_wise_mapping.py (torch entry points)."""
from .. import ops

__all__ = ["warp", "warp_with_mapping"]


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
