#!/usr/bin/env python
"""Per-kernel roofline table of the path's other launches (one B200): CUDA-event time over rotating buffer sets
larger than L2, algorithmic bytes per output pixel as in DESIGN.md section 4.  Not the headline (bench.py is).

    python tools/bench_kernels.py > gpurun_out/kernels.txt
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dmhomo_b200 import ops, synth  # noqa: E402
from dmhomo_b200.compat import dgm, hem_utils  # noqa: E402

DEV = torch.device("cuda", 0)
PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.isfile(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, sets, iters=24):
    """Device time per call: every (fn, input set) is captured into its own CUDA graph and the graphs are replayed
    in rotation between two events, so Python / ctypes launch overhead (30-50 us per op) is not what is measured."""
    stream = torch.cuda.Stream(DEV)
    graphs = []
    with torch.cuda.stream(stream):
        for st in sets:
            for _ in range(2):
                fn(*st)
        stream.synchronize()
        pool = None
        keep = []
        for st in sets:
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, pool=pool, stream=stream):
                keep.append(fn(*st))
            pool = gr.pool()
            graphs.append(gr)
        for gr in graphs:
            gr.replay()
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(iters):
            graphs[i % len(graphs)].replay()
        e1.record(stream)
        stream.synchronize()
    return e0.elapsed_time(e1) / iters


def row(name, ms, px, bpp):
    gbs = px * bpp / (ms * 1e-3) / 1e9
    print(f"{name:58s} {ms * 1e3:9.1f} us  {px / ms * 1e-6:8.1f} Gpix/s  {bpp:5.1f} B/px  {gbs:7.0f} GB/s  frac {gbs / PEAK:.3f}", flush=True)


def Hs(B, h, w, rho, seed):
    gen = torch.Generator().manual_seed(seed)
    src = synth.corner_points(B, h, w)
    return ops.dlt4(src.to(DEV), (src + synth.corner_offsets(B, rho, gen)).to(DEV)).detach()


def main():
    print(f"# peak {PEAK:.0f} GB/s (MEASURED_PEAKS.json hbm_gbs); CUDA-graph replays, input sets rotate so that inputs exceed the 126 MB L2")
    gen = torch.Generator(device=DEV).manual_seed(230)
    with torch.no_grad():
        # forward warp + validity mask from a homography: 8C + 1 B/px
        for (name, B, C, h, w, rho) in [("cfg1 S1 forward warp+mask, H param   B=16x4 C=1 360x640", 64, 1, 360, 640, 32.0),
                                        ("cfg2 shape S1 forward warp+mask        B=64 C=1 320x576", 64, 1, 320, 576, 32.0),
                                        ("cfg5 frames S1 forward warp+mask       B=16 C=3 1080x1920", 16, 3, 1080, 1920, 64.0),
                                        ("cfg4 shape S1 forward warp+mask        B=128 C=3 512x512", 128, 3, 512, 512, 32.0)]:
            sets = [(torch.rand(B, C, h, w, generator=gen, device=DEV), Hs(B, h, w, rho, 5 + k)) for k in range(4)]
            ms = timeit(lambda img, H: ops.warp(img, H, kind=ops.PARAM_HOMOGRAPHY, return_mask=True), sets)
            row(name, ms, B * h * w, 8 * C + 1)
        # explicit flow parameterisation (drop-in get_warp_flow): 8C + 8 B/px
        # flows: what HEM predicts (a smooth 8-basis flow, weights U(-4, 4)) and, as the worst case for a gather, per-pixel noise
        def basis_flow(B, h, w):
            bs = hem_utils.gen_basis(h, w).to(DEV)
            return ops.basis_combine(bs, (torch.rand(B, 8, generator=gen, device=DEV) * 2 - 1) * 4.0, h, w)

        B, C, h, w = 64, 1, 320, 576
        sets = [(torch.rand(B, C, h, w, generator=gen, device=DEV), basis_flow(B, h, w)) for _ in range(4)]
        row("get_warp_flow(img, basis flow) S1              B=64 C=1 320x576", timeit(lambda i, f: hem_utils.get_warp_flow(i, f), sets),
            B * h * w, 8 * C + 8)
        sets = [(torch.rand(B, C, h, w, generator=gen, device=DEV), torch.randn(B, 2, h, w, generator=gen, device=DEV) * 8) for _ in range(4)]
        row("get_warp_flow(img, noise flow sigma 8) S1      B=64 C=1 320x576", timeit(lambda i, f: hem_utils.get_warp_flow(i, f), sets),
            B * h * w, 8 * C + 8)
        B, C, h, w = 64, 12, 80, 144
        sets = [(torch.rand(B, C, h, w, generator=gen, device=DEV), basis_flow(B, h, w)) for _ in range(6)]
        row("get_warp_flow(feat, basis flow) pyramid level  B=64 C=12 80x144", timeit(lambda i, f: hem_utils.get_warp_flow(i, f), sets),
            B * h * w, 8 * C + 8)
        sets = [(torch.rand(B, C, h, w, generator=gen, device=DEV), torch.randn(B, 2, h, w, generator=gen, device=DEV) * 4) for _ in range(6)]
        row("get_warp_flow(feat, noise flow sigma 4)        B=64 C=12 80x144", timeit(lambda i, f: hem_utils.get_warp_flow(i, f), sets),
            B * h * w, 8 * C + 8)
        # homography -> flow (fp32, bit-exact chain): 8 B/px written
        B, h, w = 256, 320, 576
        sets = [(Hs(B, h, w, 32.0, 40 + k),) for k in range(2)]
        row("get_flow(H) -> flow fp32                        B=256 320x576", timeit(lambda H: ops.homography_to_flow(H, h, w), sets), B * h * w, 8)
        # basis combine: 8 B/px written (+ basis once)
        basis = hem_utils.gen_basis(h, w).to(DEV)
        sets = [(((torch.rand(B, 8, generator=gen, device=DEV) * 2 - 1) * 4),) for _ in range(2)]
        row("basis flow = sum_k w_k basis_k                  B=256 320x576", timeit(lambda wt: ops.basis_combine(basis, wt, h, w), sets), B * h * w, 8)
        # M1 mask from a flow: 8 read + 1 written
        sets = [(torch.randn(B, 2, h, w, generator=gen, device=DEV) * 20,) for _ in range(2)]
        row("create_border_mask(flow)                        B=256 320x576", timeit(lambda f: ops.border_mask(f), sets), B * h * w, 9)
        # cfg3 DGM rendering
        B, h, w = 25 * 16, 256, 256
        H64 = [torch.eye(3, dtype=torch.float64, device=DEV).repeat(B, 1, 1) + torch.randn(B, 3, 3, generator=gen, device=DEV, dtype=torch.float64)
               * torch.tensor([[1e-2, 1e-2, 4.0], [1e-2, 1e-2, 4.0], [1e-5, 1e-5, 0.0]], device=DEV, dtype=torch.float64) for _ in range(2)]
        row("cfg3 homo_to_flow fp64 -> fp32 flow             B=400 256x256", timeit(lambda H: ops.homography_to_flow_f64(H, h, w), [(x,) for x in H64]), B * h * w, 8)
        fl = [(torch.randn(B, h, w, 2, generator=gen, device=DEV) * 20,) for _ in range(2)]
        row("cfg3 flow_to_image (HSV wheel)                  B=400 256x256", timeit(lambda f: ops.flow_to_rgb(f, in_channels_last=True, out_channels_last=True), fl), B * h * w, 20)
        im = [(torch.rand(B, 3, h, w, generator=gen, device=DEV), H64[k]) for k in range(2)]
        row("cfg3 warpPerspective (cv2-exact S4)             B=400 C=3 256x256", timeit(lambda i, H: ops.warp_perspective(i, H, (w, h)), im), B * h * w, 24)
        sets = [(torch.rand(B, 3, h, w, generator=gen, device=DEV), ops.homography_to_flow_f64(H64[k], h, w, eps=1e-6, channels_last=False)) for k in range(2)]
        row("cfg3 flow_warp (S3 border, homography flow)     B=400 C=3 256x256", timeit(lambda i, f: dgm.flow_warp(i, f), sets), B * h * w, 8 * 3 + 8)
        sets = [(torch.rand(B, 3, h, w, generator=gen, device=DEV), torch.randn(B, 2, h, w, generator=gen, device=DEV) * 6) for _ in range(2)]
        row("cfg3 flow_warp (S3 border, noise flow sigma 6)  B=400 C=3 256x256", timeit(lambda i, f: dgm.flow_warp(i, f), sets), B * h * w, 8 * 3 + 8)
        # section 8f rows
        B, H_, W_ = 64, 360, 640
        u8 = [(torch.randint(0, 256, (B, 6, H_, W_), generator=gen, device=DEV, dtype=torch.uint8),
               torch.tensor([[32, 20]] * B, dtype=torch.int32, device=DEV)) for _ in range(3)]
        row("uint8 pair format -> grey full+patch+RGB        B=64 6x360x640", timeit(lambda x, s: ops.pairs_u8_to_gray(x, s, (320, 576)), u8),
            B * H_ * W_, 6 + 8 + 24 + 8 * (320 * 576) / (H_ * W_))
        Hd = [(H64[k][:128],) for k in range(2)]
        row("loaders' GT flow homo_convert_to_flow           B=128 360x640", timeit(lambda H: ops.homography_to_flow_f64(H, H_, W_, eps=1e-8, channels_last=False, as_mapping=2), Hd),
            128 * H_ * W_, 8)
        lo = [(torch.randn(256, 2, 80, 144, generator=gen, device=DEV),) for _ in range(2)]
        row("upsample2d_flow_as x4 (if_rate)                 B=256 80x144->320x576", timeit(lambda f: ops.flow_upsample(f, (320, 576), if_rate=True), lo), 256 * 320 * 576, 8 + 8 / 16)


if __name__ == "__main__":
    main()
