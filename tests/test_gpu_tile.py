"""GPU parity of the persistent tiled kernel (dmhomo_b200/csrc/dmh_warp_tile.cu) against the CPU oracle.

The dense S1 homography launches (forward: warped image + validity mask; fused: masked L1 + all gradients)
of C = 1 images go through the tiled kernel by default.  These cases aim at its own machinery: partial tile
rows / columns, fewer tiles than CTAs, the dynamic tail of the tile schedule, windows that do not cover the
tile (global fallback taps), homographies outside the packed division's proven domain (scalar __fdiv_rn
path), start offsets, and - in a subprocess with DMH_TUNING=tile=3 - the C = 3 instantiation.
Bars: warped pixels and masks bit-exact, loss / gradients within 1e-4 absolute (north_star).
"""
import os
import subprocess
import sys

import pytest
import torch

from dmhomo_b200 import ops, synth
from dmhomo_b200.compat import hem_utils
from oracle import port

pytestmark = pytest.mark.gpu
DEV = "cuda"
ATOL = 1e-4
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def g(seed):
    return torch.Generator().manual_seed(seed)


def _homographies(B, h, w, rho, seed):
    src = port.corner_points(B, h, w)
    return port.dlt4(src, src + synth.corner_offsets(B, rho, g(seed)))


def _check_forward(img, H, start=0):
    B, C, h, w = img.shape
    flow, _ = port.homography_to_flow(H, h, w, start=start)
    ref = port.get_warp_flow(img, flow, start=start)
    out, mask = ops.warp(img.to(DEV), H.to(DEV), kind=ops.PARAM_HOMOGRAPHY, return_mask=True, start=start)
    assert torch.equal(mask.cpu(), port.correspondence_mask(flow)), "validity mask differs"
    assert torch.equal(out.cpu(), ref), f"warped pixels differ (max {(out.cpu() - ref).abs().max().item():.3e})"


@pytest.mark.parametrize(
    "B,h,w,rho,start",
    [
        (3, 360, 640, 32.0, 0),     # cfg1 shape: last tile row is partial (360 = 5 * 64 + 40)
        (2, 320, 576, 32.0, 0),     # cfg2 shape
        (2, 128, 100, 8.0, 0),      # partial tile column (w % 64 != 0), w % 4 == 0
        (1, 64, 64, 4.0, 0),        # one tile: far fewer tiles than CTAs, everything in the static chunk of CTA 0
        (5, 192, 320, 24.0, 3),     # start offset (get_grid start = 3)
        (2, 256, 256, 16.0, 0.5),   # fractional start
        (150, 64, 128, 6.0, 0),     # 300 tiles on 148 CTAs: static chunks of 1-2 tiles + dynamic tail
    ],
)
def test_tile_forward_bit_exact(B, h, w, rho, start):
    img = synth.noise_images(B, 1, h, w, g(101))
    _check_forward(img, _homographies(B, h, w, rho, 102), start)


def test_tile_forward_strong_warp_uses_global_taps():
    """Corner offsets of 40 % of the image: the pre-image of a 64 x 64 tile does not fit the staged window,
    so part of the taps come from global memory - the result must not depend on the window."""
    B, h, w = 3, 256, 256
    img = synth.noise_images(B, 1, h, w, g(111))
    _check_forward(img, _homographies(B, h, w, 100.0, 112))


def test_tile_forward_scaled_homography_takes_scalar_division():
    """H * 2^30 is the same projective map but lies outside the magnitude range for which the packed Newton
    division is proven exact: the tile must fall back to scalar IEEE division and stay bit-exact."""
    B, h, w = 2, 128, 192
    img = synth.noise_images(B, 1, h, w, g(121))
    H = _homographies(B, h, w, 12.0, 122) * float(2 ** 30)
    _check_forward(img, H)
    _check_forward(img, _homographies(B, h, w, 12.0, 122) * float(2.0 ** -30))


def test_tile_forward_horizon_inside_image():
    """A homography whose T changes sign inside the image (and hits the |T| < 1e-7 epsilon rule region):
    no window is staged for such tiles; coordinates, masks and pixels still follow the reference bit for bit."""
    B, h, w = 2, 128, 128
    img = synth.noise_images(B, 1, h, w, g(131))
    H = _homographies(B, h, w, 4.0, 132)
    H[:, 2, 0] = -1.0 / 70.0   # T = 1 - x / 70 + ...: zero near x = 70
    _check_forward(img, H)


def _affine(B, sx, sy, tx, ty, shear=0.0):
    H = torch.eye(3).repeat(B, 1, 1)
    H[:, 0, 0], H[:, 1, 1], H[:, 0, 1], H[:, 0, 2], H[:, 1, 2] = sx, sy, shear, tx, ty
    return H


@pytest.mark.parametrize("tx,ty", [(2.0, 2.0), (1.99, 2.01), (2.5, 1.0), (0.0, 0.0), (-3.0, 7.25), (61.0, 3.0), (62.0, 62.0),
                                   (63.5, 0.5), (130.0, -70.0)])
def test_tile_interior_flag_boundaries_translation(tx, ty):
    """Pure translations put the pre-image of every tile at a known distance from the source border: offsets
    around the 2-pixel margin of the interior-tile proof (clamp-free, mask-free body) flip tiles between the two
    bodies; both must reproduce the reference bit for bit, including the zeroed / clamped border taps."""
    B, h, w = 2, 192, 256
    img = synth.noise_images(B, 1, h, w, g(181))
    _check_forward(img, _affine(B, 1.0, 1.0, tx, ty))


@pytest.mark.parametrize("sx,sy,shear", [(1.07, 0.93, 0.0), (0.9, 1.1, 0.05), (1.0, 1.25, -0.1), (0.5, 0.5, 0.0)])
def test_tile_interior_scaled_rows_skip_and_repeat(sx, sy, shear):
    """Vertical scale != 1: tap rows are skipped (sy > 1) or repeated (sy < 1) between consecutive output rows, so
    the vertical merging of the scatter takes its rare-seam branch in interior tiles; forward bit-exact, loss and
    gradients within 1e-4."""
    B, h, w = 2, 256, 320
    img1, img2 = synth.noise_images(B, 1, h, w, g(191)), synth.noise_images(B, 1, h, w, g(192))
    H = _affine(B, sx, sy, 20.0 * (1 - sx) + 3.3, 30.0 * (1 - sy) + 4.7, shear)
    _check_forward(img2, H)
    i1c, i2c, Hc = img1.clone().requires_grad_(True), img2.clone().requires_grad_(True), H.clone().requires_grad_(True)
    ff = port.homography_to_flow(Hc, h, w)[0]
    ref = port.masked_l1(port.border_mask(ff).unsqueeze(1), i1c, port.get_warp_flow(i2c, ff))
    ref.backward()
    i1g, i2g = img1.to(DEV).requires_grad_(True), img2.to(DEV).requires_grad_(True)
    Hg = H.to(DEV).requires_grad_(True)
    loss = ops.warp_loss([ops.WarpTerm(i2g, i1g, Hg)], kind=ops.PARAM_HOMOGRAPHY)
    loss.backward()
    assert abs(loss.item() - ref.item()) < 1e-5
    assert (i1g.grad.cpu() - i1c.grad).abs().max().item() < ATOL
    assert (i2g.grad.cpu() - i2c.grad).abs().max().item() < ATOL
    assert ((Hg.grad.cpu() - Hc.grad).norm() / Hc.grad.norm()).item() < 1e-3


_AB_SCRIPT = r"""
import sys, torch
sys.path.insert(0, %r)
from dmhomo_b200 import ops, synth
from dmhomo_b200.compat import hem_utils
from oracle import port
g = lambda s: torch.Generator().manual_seed(s)
B, h, w = 6, 320, 576
img1, img2 = synth.noise_images(B, 1, h, w, g(201)).cuda(), synth.noise_images(B, 1, h, w, g(202)).cuda()
src = port.corner_points(B, h, w)
Hf = port.dlt4(src, src + synth.corner_offsets(B, 32.0, g(203))).cuda()
out, mask = ops.warp(img2, Hf, kind=ops.PARAM_HOMOGRAPHY, return_mask=True)
i1, i2, Hg = img1.clone().requires_grad_(True), img2.clone().requires_grad_(True), Hf.clone().requires_grad_(True)
loss = ops.warp_loss([ops.WarpTerm(i2, i1, Hg)], kind=ops.PARAM_HOMOGRAPHY)
loss.backward()
torch.save(dict(out=out.cpu(), mask=mask.cpu(), loss=loss.detach().cpu(), g1=i1.grad.cpu(), g2=i2.grad.cpu(), gH=Hg.grad.cpu()), sys.argv[1])
"""


def test_tile_interior_body_matches_general_body(tmp_path):
    """A/B in subprocesses: DMH_TUNING=tile_interior=0 forces every tile through the general (clamping, masking) body,
    1 allows whole interior tiles only, 3 (default) adds the per-row-pair vote inside border tiles.
    Forward output, mask and dL/dtarget (no atomics involved) must be bit-identical; the scattered dL/dsrc and the
    reductions only differ by fp32 summation order."""
    res = []
    for flag in ("3", "1", "0"):   # interior + mixed (default) / interior only / general body only
        path = str(tmp_path / f"ab{flag}.pt")
        env = dict(os.environ, DMH_TUNING="tile_interior=" + flag)
        r = subprocess.run([sys.executable, "-c", _AB_SCRIPT % ROOT, path], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
        res.append(torch.load(path))
    b = res[-1]
    for a in res[:-1]:
        assert torch.equal(a["out"], b["out"]) and torch.equal(a["mask"], b["mask"])
        assert torch.equal(a["g1"], b["g1"]), "dL/dtarget differs between the fast and the general body"
        assert (a["g2"] - b["g2"]).abs().max().item() < 1e-7
        assert abs(a["loss"].item() - b["loss"].item()) < 1e-6
        assert ((a["gH"] - b["gH"]).norm() / b["gH"].norm()).item() < 1e-4


def test_tile_identity_zeroes_last_row_and_col():
    B, h, w = 2, 64, 128
    img = synth.noise_images(B, 1, h, w, g(141))
    H = torch.eye(3).repeat(B, 1, 1)
    out, mask = ops.warp(img.to(DEV), H.to(DEV), kind=ops.PARAM_HOMOGRAPHY, return_mask=True)
    ref = img.clone()
    ref[:, :, -1, :] = 0
    ref[:, :, :, -1] = 0
    assert torch.equal(out.cpu(), ref)
    assert bool(mask.all())


@pytest.mark.parametrize("B,h,w,rho", [(3, 320, 576, 32.0), (2, 360, 640, 32.0), (160, 64, 64, 5.0), (2, 256, 256, 90.0)])
def test_tile_fused_loss_and_gradients(B, h, w, rho):
    img1, img2 = synth.noise_images(B, 1, h, w, g(151)), synth.noise_images(B, 1, h, w, g(152))
    Hf, Hb = _homographies(B, h, w, rho, 153), _homographies(B, h, w, rho, 154)
    i1c, i2c = img1.clone().requires_grad_(True), img2.clone().requires_grad_(True)
    Hfc, Hbc = Hf.clone().requires_grad_(True), Hb.clone().requires_grad_(True)
    ff, fb = port.homography_to_flow(Hfc, h, w)[0], port.homography_to_flow(Hbc, h, w)[0]
    mf, mb = port.border_mask(ff).unsqueeze(1), port.border_mask(fb).unsqueeze(1)
    ref = port.masked_l1(mf, i1c, port.get_warp_flow(i2c, ff)) + port.masked_l1(mb, i2c, port.get_warp_flow(i1c, fb))
    ref.backward()
    i1g, i2g = img1.to(DEV).requires_grad_(True), img2.to(DEV).requires_grad_(True)
    Hfg, Hbg = Hf.to(DEV).requires_grad_(True), Hb.to(DEV).requires_grad_(True)
    loss = ops.warp_loss([ops.WarpTerm(i2g, i1g, Hfg), ops.WarpTerm(i1g, i2g, Hbg)], kind=ops.PARAM_HOMOGRAPHY)
    assert abs(loss.item() - ref.item()) < 1e-5
    loss.backward()
    assert (i1g.grad.cpu() - i1c.grad).abs().max().item() < ATOL
    assert (i2g.grad.cpu() - i2c.grad).abs().max().item() < ATOL
    for a, b in ((Hfg.grad.cpu(), Hfc.grad), (Hbg.grad.cpu(), Hbc.grad)):
        assert ((a - b).norm() / b.norm()).item() < 1e-3


def test_tile_fused_repeatable_and_counter_slots():
    """More launches than counter slots (64): the self-resetting tile counters must leave every launch with the
    full tile list (identical loss every time)."""
    B, h, w = 4, 128, 192
    img1, img2 = synth.noise_images(B, 1, h, w, g(161)).to(DEV), synth.noise_images(B, 1, h, w, g(162)).to(DEV)
    Hf, Hb = _homographies(B, h, w, 10.0, 163).to(DEV), _homographies(B, h, w, 10.0, 164).to(DEV)
    first = None
    for _ in range(150):
        loss = ops.warp_loss([ops.WarpTerm(img2, img1, Hf), ops.WarpTerm(img1, img2, Hb)], kind=ops.PARAM_HOMOGRAPHY)
        v = loss.item()
        first = v if first is None else first
        assert abs(v - first) < 1e-6


_C3_SCRIPT = r"""
import sys, torch
sys.path.insert(0, %r)
from dmhomo_b200 import ops, synth
from dmhomo_b200.compat import hem_utils
from oracle import port
g = lambda s: torch.Generator().manual_seed(s)
B, C, h, w = 2, 3, 128, 192
img1, img2 = synth.noise_images(B, C, h, w, g(171)), synth.noise_images(B, C, h, w, g(172))
src = port.corner_points(B, h, w)
Hf = port.dlt4(src, src + synth.corner_offsets(B, 12.0, g(173)))
Hb = port.dlt4(src, src + synth.corner_offsets(B, 12.0, g(174)))
flow, _ = port.homography_to_flow(Hf, h, w)
out, mask = ops.warp(img2.cuda(), Hf.cuda(), kind=ops.PARAM_HOMOGRAPHY, return_mask=True)
assert torch.equal(out.cpu(), port.get_warp_flow(img2, flow)), "C=3 tiled forward differs"
assert torch.equal(mask.cpu(), port.correspondence_mask(flow))
i1c, i2c = img1.clone().requires_grad_(True), img2.clone().requires_grad_(True)
ff, fb = port.homography_to_flow(Hf, h, w)[0], port.homography_to_flow(Hb, h, w)[0]
mf, mb = port.border_mask(ff).unsqueeze(1), port.border_mask(fb).unsqueeze(1)
ref = port.masked_l1(mf, i1c, port.get_warp_flow(i2c, ff)) + port.masked_l1(mb, i2c, port.get_warp_flow(i1c, fb))
ref.backward()
i1g, i2g = img1.cuda().requires_grad_(True), img2.cuda().requires_grad_(True)
loss = ops.warp_loss([ops.WarpTerm(i2g, i1g, Hf.cuda().requires_grad_(True)), ops.WarpTerm(i1g, i2g, Hb.cuda().requires_grad_(True))],
                     kind=ops.PARAM_HOMOGRAPHY)
loss.backward()
assert abs(loss.item() - ref.item()) < 1e-5
assert (i1g.grad.cpu() - i1c.grad).abs().max().item() < 1e-4
assert (i2g.grad.cpu() - i2c.grad).abs().max().item() < 1e-4
print("C3 OK")
"""


def test_tile_c3_instantiation_in_subprocess():
    env = dict(os.environ, DMH_TUNING="tile=3")
    r = subprocess.run([sys.executable, "-c", _C3_SCRIPT % ROOT], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "C3 OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


# ---------------------------------------------------------------------------------- explicit flow on the tile kernel
def _flows(B, h, w, seed):
    """A smooth basis-like flow (what HEM predicts), a noisy one (taps all over the window), and one that throws part of
    the tile far outside the staged window (global fallback path)."""
    gen = g(seed)
    basis = hem_utils.gen_basis(h, w)
    smooth = ops.basis_combine(basis.to(DEV), synth.basis_weights(B, gen, 4.0).to(DEV), h, w).cpu()
    noisy = smooth + torch.randn(B, 2, h, w, generator=gen) * 3.0
    wild = smooth.clone()
    wild[:, :, ::7, ::5] += 90.0
    wild[:, 0, 40:60, 100:130] -= 300.0
    return {"smooth": smooth, "noisy": noisy, "wild": wild}


def _scatter_mass(flow, gout, Hs, Ws):
    """Sum of |bilinear weight| * |upstream gradient| over every tap that lands on a source pixel (S1 taps and weights
    as in HEM/model/utils.py:463-523, fp64): the magnitude scale of the scattered sum dL/dimg, (B,1,Hs,Ws)."""
    B, _, h, w = flow.shape
    c = port.pixel_grid(B, h, w)[:, :2].double() + flow.double()
    x, y = c[:, 0].reshape(-1), c[:, 1].reshape(-1)
    x0, y0 = torch.floor(x), torch.floor(y)
    x1, y1 = (x0 + 1).clamp(0, Ws - 1), (y0 + 1).clamp(0, Hs - 1)
    x0, y0 = x0.clamp(0, Ws - 1), y0.clamp(0, Hs - 1)
    base = (torch.arange(B) * (Hs * Ws)).repeat_interleave(h * w)
    ga = gout.double().abs().sum(1).reshape(-1)
    m = torch.zeros(B * Hs * Ws, dtype=torch.float64)
    for xi, yi, wt in ((x0, y0, (x1 - x) * (y1 - y)), (x0, y1, (x1 - x) * (y - y0)), (x1, y0, (x - x0) * (y1 - y)), (x1, y1, (x - x0) * (y - y0))):
        m.index_add_(0, base + yi.long() * Ws + xi.long(), wt.abs() * ga)
    return m.reshape(B, 1, Hs, Ws)


@pytest.mark.parametrize("kind", ["smooth", "noisy", "wild"])
def test_tile_flow_forward_backward_against_oracle_and_scalar_kernel(kind):
    from dmhomo_b200 import _lib

    B, C, h, w = 3, 1, 128, 192
    img = synth.noise_images(B, C, h, w, g(151))
    flow = _flows(B, h, w, 152)[kind]
    gout = torch.randn(B, C, h, w, generator=g(153))
    ic, fc = img.clone().requires_grad_(True), flow.clone().requires_grad_(True)
    ref = port.get_warp_flow(ic, fc)
    (ref * gout).sum().backward()
    # Far out-of-bounds coordinates clamp both taps to the border pixel with weights +d and -d (d = distance, up to 300
    # here): the border pixels of dL/dimg are sums of thousands of cancelling terms, so the reference's own fp32 result
    # is noise there.  Yardstick: the same pipeline in fp64; the kernels must be as close to it as the fp32 oracle is.
    i64, f64 = img.double().requires_grad_(True), flow.double().requires_grad_(True)
    (port.get_warp_flow(i64, f64) * gout.double()).sum().backward()
    # ... within one fp32 ulp of the MASS of the sum (sum of |weight * upstream gradient| over the taps that land on a
    # pixel: 1.6e6 at the corner pixel of the wild case, where the fp32 oracle itself is off by 4e-3), or 1e-4.
    # The order of the atomic adds differs from run to run, so a multiple of the oracle's own error is no stable bound.
    img_tol = 1e-4 + _scatter_mass(flow, gout, h, w) * 2.0 ** -24
    assert ((ic.grad.double() - i64.grad).abs() <= img_tol).all()      # the yardstick holds for the reference's own fp32

    res = {}
    for tile_flow in (1, 0):
        _lib.set_tuning(tile_flow=tile_flow)
        try:
            ig, fg = img.to(DEV).requires_grad_(True), flow.to(DEV).requires_grad_(True)
            out = hem_utils.get_warp_flow(ig, fg)
            kern_f = ops.last_warp_kernel
            (out * gout.to(DEV)).sum().backward()
            kern_b = ops.last_warp_kernel
        finally:
            _lib.set_tuning(tile_flow=1)
        assert ("tile" in kern_f) == bool(tile_flow) and ("tile" in kern_b) == bool(tile_flow), (kern_f, kern_b)
        assert torch.equal(out.detach().cpu(), ref.detach()), "warped pixels differ from the oracle"
        assert (ig.grad.cpu()[..., 1:-1, 1:-1] - ic.grad[..., 1:-1, 1:-1]).abs().max().item() < 1e-4
        assert ((ig.grad.cpu().double() - i64.grad).abs() <= img_tol).all()
        assert (fg.grad.cpu() - fc.grad).abs().max().item() < 1e-4 * max(1.0, fc.grad.abs().max().item())
        res[tile_flow] = (out.detach(), ig.grad, fg.grad)
    assert torch.equal(res[1][2], res[0][2]), "dL/dflow differs between the tile and the scalar kernel"
    assert ((res[1][1] - res[0][1]).abs().cpu().double() <= 2 * img_tol).all()      # scattered sums: order differs

    # only one of the two gradients wanted
    ig = img.to(DEV).requires_grad_(True)
    (hem_utils.get_warp_flow(ig, flow.to(DEV)) * gout.to(DEV)).sum().backward()
    assert ((ig.grad - res[1][1]).abs().cpu().double() <= 2 * img_tol).all()
    fg = flow.to(DEV).requires_grad_(True)
    (hem_utils.get_warp_flow(img.to(DEV), fg) * gout.to(DEV)).sum().backward()
    assert torch.equal(fg.grad, res[1][2])


def test_tile_flow_fused_loss_direct_variant():
    """The reference's own training warp (HEM/model/net.py:808-818): warp by the basis flow, masked L1, gradients to both
    images and both flows in one launch."""
    from dmhomo_b200 import _lib

    B, C, h, w = 4, 1, 160, 256
    gen = g(154)
    img1, img2 = synth.smooth_images(B, C, h, w, gen), synth.smooth_images(B, C, h, w, gen)
    fl = _flows(B, h, w, 155)
    ff, fb = fl["smooth"], fl["noisy"]
    leaves_c = [t.clone().requires_grad_(True) for t in (img1, img2, ff, fb)]
    a, b, ffc, fbc = leaves_c
    mf, mb = port.border_mask(ffc).unsqueeze(1), port.border_mask(fbc).unsqueeze(1)
    ref = port.masked_l1(mf, a, port.get_warp_flow(b, ffc)) + port.masked_l1(mb, b, port.get_warp_flow(a, fbc))
    ref.backward()
    for tile_flow in (1, 0):
        _lib.set_tuning(tile_flow=tile_flow)
        try:
            lg = [t.to(DEV).requires_grad_(True) for t in (img1, img2, ff, fb)]
            loss = ops.warp_loss([ops.WarpTerm(lg[1], lg[0], lg[2]), ops.WarpTerm(lg[0], lg[1], lg[3])], kind=ops.PARAM_FLOW,
                                 border_mask=True)
            assert ("tile" in ops.last_warp_kernel) == bool(tile_flow)
            loss.backward()
        finally:
            _lib.set_tuning(tile_flow=1)
        assert abs(loss.item() - ref.item()) < 1e-5
        for tg, tc in zip(lg, leaves_c):
            assert (tg.grad.cpu() - tc.grad).abs().max().item() < 1e-4
