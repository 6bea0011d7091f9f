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

# reference module (either import root) -> (our module, names to rebind)
_TARGETS = {
    "model.utils": (hem_utils, hem_utils.__all__),
    "model.net": (hem_net, ["DLT_solve"]),
    "utils_operations.pixel_wise_mapping": (pixel_wise_mapping, pixel_wise_mapping.__all__),
    "utils_operations.flow_and_mapping_operations": (flow_and_mapping_operations,
                                                     ["get_gt_correspondence_mask", "create_border_mask",
                                                      "from_homography_to_pixel_wise_mapping"]),
    "loss.losses": (losses, ["LossL1", "ComputeErrFlow", "compute_eval_results"]),
    "dataset.data_loader": (dgm, ["flow_warp"]),
    "denoising_diffusion_models.denoising_diffusion_pytorch": (dgm, ["flow_warp", "visulize_flow", "postProcess",
                                                                     "postProcess_cv2", "homo_gen"]),
    "denoising_diffusion_models.classifier_free_guidance": (dgm, ["flow_warp"]),
}
# modules that `from model.utils import ...` the names above
_IMPORTERS = {
    "model.net": (hem_utils, ["get_warp_flow", "get_grid", "get_flow", "transformer", "upsample2d_flow_as"]),
    "model.swin_multi": (hem_utils, ["get_warp_flow", "upsample2d_flow_as"]),
}


def patch_reference(verbose=False):
    """Rebind the hot-path names in every imported reference module.  Returns the list of
    (module, name) pairs that were replaced."""
    done = []
    for table in (_TARGETS, _IMPORTERS):
        for suffix, (ours, names) in table.items():
            for prefix in ("HEM.", "", "DGM."):
                mod = sys.modules.get(prefix + suffix)
                if mod is None:
                    continue
                for n in names:
                    if hasattr(mod, n) and hasattr(ours, n):
                        setattr(mod, n, getattr(ours, n))
                        done.append((mod.__name__, n))
    if verbose:
        for m, n in done:
            print(f"dmhomo_b200: {m}.{n} -> CUDA")
    return done
