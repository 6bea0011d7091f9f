"""dmhomo_b200 - B200-native (sm_100a) kernels for DMHomo's batched homography-warp hot path.

    from dmhomo_b200 import ops            # torch operators over the C ABI (include/dmhomo.h)
    from dmhomo_b200.compat import ...     # the reference's own function names, drop-in

Importing the package never touches CUDA (safe in forked DataLoader workers); the shared
library loads on the first op.  There is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (no CUDA work at import)


def library_path():
    return _lib.LIB_PATH
