"""Import the real DMHomo reference *in place* (TEST INFRASTRUCTURE ONLY).

Only usable where /root/reference (or $DMHOMO_REF) is mounted, i.e. in the build
container - never on the GPU box.  Arithmetic-free third-party modules that are
missing from the image are replaced by inert stubs; the one
stub that carries arithmetic is ``matplotlib.colors.hsv_to_rgb`` -> oracle/hsv.py.
No reference source is copied: modules are imported from where they lie.
"""
import importlib
import os
import sys
import types

_CANDIDATES = [os.environ.get("DMHOMO_REF"), "/root/reference"]


def reference_root():
    for c in _CANDIDATES:
        if c and os.path.isfile(os.path.join(c, "HEM", "model", "utils.py")):
            return c
    return None


def available():
    return reference_root() is not None


def _stub(name, **attrs):
    try:
        return importlib.import_module(name)
    except Exception:
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__dict__["__dmhomo_stub__"] = True
        sys.modules[name] = m
        return m


class _Ref:
    pass


_cache = None


def load():
    """Returns a namespace with the reference's hot-path modules."""
    global _cache
    if _cache is not None:
        return _cache
    root = reference_root()
    if root is None:
        raise RuntimeError("DMHomo reference not mounted (looked in $DMHOMO_REF, /root/reference)")
    for p in (os.path.join(root, "DGM"), root):
        if p not in sys.path:
            sys.path.insert(0, p)
    from . import hsv as _hsv

    _stub("imageio")
    kl = _stub("kornia.geometry.linalg", transform_points=None)
    kg = _stub("kornia.geometry", linalg=kl)
    _stub("kornia", geometry=kg)
    _stub("timm")
    _stub("timm.models")
    _stub("timm.models.layers", trunc_normal_=None, DropPath=None, to_2tuple=None)
    mc = _stub("matplotlib.colors", hsv_to_rgb=_hsv.hsv_to_rgb)
    _stub("matplotlib", colors=mc)
    _stub("ema_pytorch", EMA=None)
    _stub("accelerate", Accelerator=None)
    _stub("denoising_diffusion_pytorch")
    _stub("denoising_diffusion_pytorch.version", __version__="stub")
    _stub("pytorch_grad_cam")
    _stub("termcolor", colored=lambda s, *a, **k: s)
    _stub("coloredlogs")
    _stub("prettytable", PrettyTable=None)

    r = _Ref()
    r.root = root
    r.utils = importlib.import_module("HEM.model.utils")
    r.pwm = importlib.import_module("HEM.utils_operations.pixel_wise_mapping")
    r.fmo = importlib.import_module("HEM.utils_operations.flow_and_mapping_operations")
    r.losses = importlib.import_module("HEM.loss.losses")
    try:
        r.net = importlib.import_module("HEM.model.net")
    except Exception as e:  # needs torchvision/timm pieces; only DLT_solve is used
        r.net = None
        r.net_error = e
    try:
        r.data_loader = importlib.import_module("HEM.dataset.data_loader")
    except Exception as e:
        r.data_loader = None
        r.data_loader_error = e
    try:
        r.ddpm = importlib.import_module("denoising_diffusion_models.denoising_diffusion_pytorch")
    except Exception as e:
        r.ddpm = None
        r.ddpm_error = e
    _cache = r
    return r
