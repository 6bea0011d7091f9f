"""Times ops.basis_combine forward / backward (development A/B; DMH_LIB selects the library)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dmhomo_b200 import ops
from dmhomo_b200.compat import hem_utils
h, w = 320, 576
basis = hem_utils.gen_basis(h, w).cuda()
for B in (64, 256):
    wt = [((torch.rand(B, 8, device="cuda") * 2 - 1) * 4).requires_grad_(True) for _ in range(4)]
    go = torch.randn(B, 2, h, w, device="cuda")
    for name, fn in (("fwd", lambda k: ops.basis_combine(basis, wt[k].detach(), h, w)),
                     ("fwd+bwd", lambda k: ops.basis_combine(basis, wt[k], h, w).backward(go))):
        for k in range(4): fn(k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                for k in range(8): fn(k % 4)
            g.replay(); s.synchronize()
            e0.record(s); g.replay(); e1.record(s); s.synchronize()
        print(f"B={B} {name}: {e0.elapsed_time(e1) / 8 * 1e3:.1f} us")
