"""A/B probe for the dense S1 homography kernels: runs forward (out + mask) and the fused loss + gradients
on seeded inputs and saves every result, or compares two such dumps.

    DMH_TILE=0 python tools/tile_probe.py run /tmp/a.pt [B C h w rho]
    DMH_TILE=1 python tools/tile_probe.py run /tmp/b.pt [B C h w rho]
    python tools/tile_probe.py cmp /tmp/a.pt /tmp/b.pt
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(path, B=8, C=1, h=320, w=576, rho=32.0, start=0):
    from dmhomo_b200 import ops, synth

    dev = torch.device("cuda", 0)
    gen = synth.generator()
    img1 = synth.noise_images(B, C, h, w, gen).to(dev)
    img2 = synth.noise_images(B, C, h, w, gen).to(dev)
    src = synth.corner_points(B, h, w).to(dev)
    Hf = ops.dlt4(src, src + synth.corner_offsets(B, rho, gen).to(dev))
    Hb = ops.dlt4(src, src + synth.corner_offsets(B, rho, gen).to(dev))
    Hf = Hf.detach().requires_grad_(True)
    Hb = Hb.detach().requires_grad_(True)
    out = {}
    w2, m = ops.warp(img2, Hf.detach(), kind=ops.PARAM_HOMOGRAPHY, return_mask=True, start=start)
    out["fwd_out"], out["fwd_mask"] = w2, m
    i1 = img1.clone().requires_grad_(True)
    i2 = img2.clone().requires_grad_(True)
    loss = ops.warp_loss([ops.WarpTerm(i2, i1, Hf), ops.WarpTerm(i1, i2, Hb)], kind=ops.PARAM_HOMOGRAPHY)
    loss.backward()
    torch.cuda.synchronize()
    out.update(loss=loss.detach(), g1=i1.grad, g2=i2.grad, gHf=Hf.grad, gHb=Hb.grad)
    torch.save({k: v.cpu() for k, v in out.items()}, path)
    print("saved", path, "loss", float(loss))


def cmp(pa, pb):
    a, b = torch.load(pa), torch.load(pb)
    bad = 0
    for k in a:
        x, y = a[k], b[k]
        if x.dtype in (torch.bool, torch.uint8):
            nd = int((x != y).sum())
            print(f"{k:9s} differing elements: {nd}")
            bad += nd
        else:
            d = (x.double() - y.double()).abs()
            eq = torch.equal(x, y)
            print(f"{k:9s} bit-equal {eq}  max abs diff {d.max().item():.3e}  (max |ref| {x.abs().max().item():.3e})")
            if k in ("fwd_out",) and not eq:
                bad += 1
    print("RESULT", "OK" if bad == 0 else "MISMATCH")


if __name__ == "__main__":
    if sys.argv[1] == "run":
        nums = [float(v) for v in sys.argv[3:]]
        kw = {}
        for name, v in zip(("B", "C", "h", "w"), nums[:4]):
            kw[name] = int(v)
        if len(nums) > 4:
            kw["rho"] = nums[4]
        if len(nums) > 5:
            kw["start"] = nums[5]
        run(sys.argv[2], **kw)
    else:
        cmp(sys.argv[2], sys.argv[3])
