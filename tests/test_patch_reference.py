"""compat.patch_reference() executed against the real reference tree (build container only: needs /root/reference).

Runs in a subprocess: the patch is process-global and the oracle tests of this suite use the unpatched reference."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")

_SCRIPT = r"""
import importlib, os, sys
sys.path.insert(0, %r)
from oracle import ref_loader
r = ref_loader.load()                                   # import root 1: HEM.model.utils, HEM.model.net, ...
sys.path.insert(0, os.path.join(r.root, "HEM"))         # import root 2: the way train.py / hem_evaluate.py run
mods = {}
for name in ("model.utils", "model.net", "model.swin_multi", "loss.losses", "utils_operations.pixel_wise_mapping",
             "utils_operations.flow_and_mapping_operations", "dataset.data_loader"):
    mods[name] = importlib.import_module(name)
cfg = importlib.import_module("denoising_diffusion_models.classifier_free_guidance")
ddpm = importlib.import_module("denoising_diffusion_models.denoising_diffusion_pytorch")
hem = {n: importlib.import_module("HEM." + n) for n in ("model.utils", "model.net", "model.swin_multi", "loss.losses")}
loader_flow_warp = mods["dataset.data_loader"].flow_warp

from dmhomo_b200 import compat
from dmhomo_b200.compat import hem_utils, hem_net, losses, dgm, pixel_wise_mapping, flow_and_mapping_operations as fmo
done = compat.patch_reference()
assert done, "nothing was patched"

def is_ours(obj):
    return getattr(obj, "__module__", "").startswith("dmhomo_b200.")

# every name of the tables, in the defining modules, under both roots
for root in (mods, hem):
    u = root["model.utils"]
    for n in hem_utils.__all__:
        assert getattr(u, n) is getattr(hem_utils, n), ("model.utils", n)
    assert root["model.net"].DLT_solve is hem_net.DLT_solve
    for n in ("LossL1", "ComputeErrFlow", "compute_eval_results"):
        assert getattr(root["loss.losses"], n) is getattr(losses, n), n
    # by-value importers: HEM/model/net.py:14, HEM/model/swin_multi.py:7
    for n in ("get_warp_flow", "get_grid", "upsample2d_flow_as"):
        assert getattr(root["model.net"], n) is getattr(hem_utils, n), ("model.net", n)
    assert root["model.swin_multi"].get_warp_flow is hem_utils.get_warp_flow
for n in ("warp", "warp_with_mapping"):
    assert getattr(mods["utils_operations.pixel_wise_mapping"], n) is getattr(pixel_wise_mapping, n)
# the cv2 remap helpers serve the loaders' host-side augmentation (flow_and_mapping_operations.py:74-81): left alone
for n in ("remap_using_flow_fields", "remap_using_correspondence_map"):
    assert not is_ours(getattr(mods["utils_operations.pixel_wise_mapping"], n)), n
for n in ("get_gt_correspondence_mask", "create_border_mask", "from_homography_to_pixel_wise_mapping"):
    assert getattr(mods["utils_operations.flow_and_mapping_operations"], n) is getattr(fmo, n)
# HEM/loss/losses.py imports the mask helpers it calls in compute_losses by value
for mod in (mods["loss.losses"], hem["loss.losses"]):
    for n, v in vars(mod).items():
        if n in ("create_border_mask", "get_gt_correspondence_mask", "get_warp_flow"):
            assert is_ours(v), (mod.__name__, n)
# DGM: definitions in ddpm.py, by-value import in classifier_free_guidance.py:18
for n in ("flow_warp", "visulize_flow", "postProcess", "postProcess_cv2", "homo_gen"):
    assert getattr(ddpm, n) is getattr(dgm, n), n
assert cfg.flow_warp is dgm.flow_warp
# loader-side CPU helper left alone (runs in forked DataLoader workers)
assert mods["dataset.data_loader"].flow_warp is loader_flow_warp
# nothing of the original hot path is left bound anywhere but the loader
import types
left = []
for mname, mod in list(sys.modules.items()):
    if not isinstance(mod, types.ModuleType) or mname.endswith("dataset.data_loader"):
        continue
    f = getattr(mod, "__file__", None) or ""
    if not f.startswith(r.root):
        continue
    for n in ("get_warp_flow", "transformer", "get_flow", "warp", "warp_with_mapping", "create_border_mask", "LossL1"):
        v = vars(mod).get(n)
        if v is not None and not is_ours(v):
            left.append((mname, n))
assert not left, left
print("PATCH OK", len(done))
"""


def test_patch_reference_rebinds_every_hot_path_name():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-c", _SCRIPT % ROOT], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PATCH OK" in r.stdout, r.stdout[-3000:] + r.stderr[-5000:]
