#!/usr/bin/env python
"""Times the REAL reference (imported in place from /root/reference through oracle/ref_loader.py; build container only)
on the cfg2 step next to the oracle port, same inputs, same host: 16 pairs 1x320x576, 8 basis weights -> basis flow at
the corners -> DLT -> get_flow -> get_warp_flow x2 -> create_border_mask x2 -> LossL1 x2, forward + backward.
Usage: python tools/time_reference.py [pairs]"""
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dmhomo_b200 import synth  # noqa: E402
from oracle import port, ref_loader  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
C, h, w = 1, 320, 576
torch.set_num_threads(os.cpu_count() or 1)
R = ref_loader.load()
gen = synth.generator()
img1, img2 = synth.noise_images(B, C, h, w, gen), synth.noise_images(B, C, h, w, gen)
wf, wb = synth.basis_weights(B, gen), synth.basis_weights(B, gen)
basis = R.utils.gen_basis(h, w).reshape(1, 8, -1)          # the reference's own basis
l1 = R.losses.LossL1(reduction="mean")
src = port.corner_points(B, h, w)


def reference_step():
    t = [x.clone().requires_grad_(True) for x in (img1, img2, wf, wb)]
    i1, i2, a, b = t
    losses = []
    for (w8, s, tg) in ((a, i2, i1), (b, i1, i2)):
        off = port.basis_corner_offsets(basis, w8, h, w)                  # 4 corner samples of (basis * w).sum(1): 32 numbers
        H = R.utils.DLT(B, 4)(src, src + off)
        grid = R.utils.get_grid(B, h, w)
        flow, _ = R.utils.get_flow(H.reshape(B, 1, 3, 3), grid, h, w, 1)
        warped = R.utils.get_warp_flow(s, flow)
        m = R.fmo.create_border_mask(flow).unsqueeze(1)
        losses.append(l1(m * tg, m * warped))
    loss = losses[0] + losses[1]
    loss.backward()
    return loss.item()


def port_step():
    t = [x.clone().requires_grad_(True) for x in (img1, img2, wf, wb)]
    return port.pipeline_basis(t[0], t[1], basis, t[2], t[3], variant="dlt", backward=True)["loss"].item()


def med(fn, n=5):
    fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        v = fn()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts), v


px = 2 * B * h * w
tr, lr = med(reference_step)
tp, lp = med(port_step)
print(f"host: {os.cpu_count()} logical CPUs, torch {torch.__version__}, {torch.get_num_threads()} threads; {B} pairs {C}x{h}x{w}, fwd+bwd")
print(f"reference (in place): {tr * 1e3:8.1f} ms/step  {px / tr / 1e6:8.2f} Mpix/s  loss {lr:.6f}")
print(f"oracle port         : {tp * 1e3:8.1f} ms/step  {px / tp / 1e6:8.2f} Mpix/s  loss {lp:.6f}")
print(f"port / reference speed ratio: {tr / tp:.2f}x")
