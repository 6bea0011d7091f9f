"""The reference's function names, backed by libdmhomo.

    from dmhomo_b200.compat import hem_utils            # HEM/model/utils.py
    from dmhomo_b200.compat import hem_net              # HEM/model/net.py (DLT_solve, basis flow)
    from dmhomo_b200.compat import pixel_wise_mapping, flow_and_mapping_operations
    from dmhomo_b200.compat import losses, dgm, data_loader   # HEM/dataset/data_loader.py helpers

`patch_reference()` rebinds those names inside an already imported reference tree (both
import roots the reference uses: `HEM.model.utils` and `model.utils`), so HEM training /
evaluation and DGM sampling pick the kernels up unchanged.
"""
import sys

from . import data_loader, dgm, flow_and_mapping_operations, hem_net, hem_utils, losses, pixel_wise_mapping  # noqa: F401

# defining module of the reference (path below either import root) -> (our module, names it defines that we replace)
_TARGETS = {
    "model.utils": (hem_utils, hem_utils.__all__),
    "model.net": (hem_net, ["DLT_solve"]),
    # (remap_using_* are host-side numpy helpers of the loaders' augmentation, flow_and_mapping_operations.py:74-81: left
    #  alone like every loader-side helper; their GPU forms are importable from compat.pixel_wise_mapping)
    "utils_operations.pixel_wise_mapping": (pixel_wise_mapping, ["warp", "warp_with_mapping"]),
    "utils_operations.flow_and_mapping_operations": (flow_and_mapping_operations,
                                                     ["get_gt_correspondence_mask", "create_border_mask",
                                                      "from_homography_to_pixel_wise_mapping"]),
    "loss.losses": (losses, ["LossL1", "ComputeErrFlow", "compute_eval_results"]),
    "denoising_diffusion_models.denoising_diffusion_pytorch": (dgm, ["flow_warp", "visulize_flow", "postProcess",
                                                                     "postProcess_cv2", "homo_gen"]),
}
# Left alone on purpose: HEM/dataset/data_loader.py's own `flow_warp`, `homo_convert_to_flow`, ... and ddpm.py's
# `homo_to_flow` / `flow_to_image` run on CPU data inside forked DataLoader workers (DGMTrainData.__getitem__,
# ddpm.py:1100-1246), where a CUDA-only op must not be called; batched GPU forms of those are compat.data_loader /
# compat.dgm, to be called from the training loop.
_PREFIXES = ("HEM.", "", "DGM.")


def patch_reference(verbose=False):
    """Rebind the hot-path names in every imported reference module.

    Step 1 replaces the definitions in their defining modules (under either import root).  Step 2 walks
    sys.modules and rebinds every module attribute that still IS one of the original objects (identity match):
    that catches each `from model.utils import get_warp_flow`-style by-value import (HEM/model/net.py:14,
    HEM/model/swin_multi.py:7, HEM/evaluate.py, hem_evaluate.py, classifier_free_guidance.py:18, ...) without a
    hand-kept importer list.  Modules imported after the call see the replacements through step 1.
    Returns the list of (module name, attribute) pairs that were rebound."""
    done, originals = [], {}
    for suffix, (ours, names) in _TARGETS.items():
        for prefix in _PREFIXES:
            mod = sys.modules.get(prefix + suffix)
            if mod is None:
                continue
            for n in names:
                if not (hasattr(mod, n) and hasattr(ours, n)):
                    continue
                old, new = getattr(mod, n), getattr(ours, n)
                if old is new:
                    continue
                originals[id(old)] = (old, new)
                setattr(mod, n, new)
                done.append((mod.__name__, n))
    if originals:
        for mname, mod in list(sys.modules.items()):
            if mod is None or mname.startswith("dmhomo_b200") or mname.startswith("oracle"):
                continue
            if mname.endswith("dataset.data_loader"):
                continue   # loader-side CPU helpers stay the reference's own (see above)
            try:
                items = list(vars(mod).items())
            except TypeError:
                continue
            for attr, val in items:
                hit = originals.get(id(val))
                if hit is not None and hit[0] is val:
                    setattr(mod, attr, hit[1])
                    done.append((mname, attr))
    if verbose:
        for m, n in done:
            print(f"dmhomo_b200: {m}.{n} -> CUDA")
    return done
