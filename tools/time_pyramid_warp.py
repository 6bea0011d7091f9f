"""Pyramid-level feature warp (C = 12 / 24) forward and backward, eager calls for an ncu launch list (development tool)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dmhomo_b200 import ops, _lib
from dmhomo_b200.compat import hem_utils
if len(sys.argv) > 1:
    _lib.set_tuning(channels=int(sys.argv[1]))
smooth = len(sys.argv) > 2 and sys.argv[2] == "smooth"
for (B, C, h, w) in ((64, 12, 80, 144), (64, 24, 40, 72)):
    feat = torch.rand(B, C, h, w, device="cuda").requires_grad_(True)
    if smooth:
        flow = ops.basis_combine(hem_utils.gen_basis(h, w).cuda(), (torch.rand(B, 8, device="cuda") * 2 - 1) * 4.0, h, w).detach().requires_grad_(True)
    else:
        flow = (torch.randn(B, 2, h, w, device="cuda") * 4).requires_grad_(True)
    go = torch.randn(B, C, h, w, device="cuda")
    for _ in range(3):
        out = hem_utils.get_warp_flow(feat, flow)
        print(ops.last_warp_kernel)
        out.backward(go)
        print(ops.last_warp_kernel)
    torch.cuda.synchronize()
