"""torch-facing operators over libdmhomo (C ABI in include/dmhomo.h).

PyTorch is only plumbing here: it owns device memory, streams and the autograd graph.
Every op enqueues hand-written sm_100a kernels on the caller's current stream through
ctypes; there is no CPU path and no PyTorch re-implementation to fall back on.
"""
import ctypes as C

import torch

from . import _lib as L
from ._lib import (LOSS_DIFF_MASKED, LOSS_MASKED_DIFF, LOSS_NONE, PARAM_BASIS8, PARAM_COORDS, PARAM_FLOW,
                   PARAM_HOMOGRAPHY, S1, S1B, S2_ZEROS, S3_BORDER)

__all__ = [
    "S1", "S1B", "S2_ZEROS", "S3_BORDER", "PARAM_FLOW", "PARAM_COORDS", "PARAM_HOMOGRAPHY", "PARAM_BASIS8", "render_conditions", "remap",
    "LOSS_NONE", "LOSS_MASKED_DIFF", "LOSS_DIFF_MASKED", "dlt4", "homography_to_flow", "homography_to_flow_f64",
    "basis_combine", "basis_corner_offsets", "basis_homography", "warp", "warp_into", "warp_loss", "warp_eval", "WarpTerm", "u8_to_f32", "pairs_u8_to_gray", "border_mask", "zero_border_mask",
    "l1_loss", "flow_to_rgb", "warp_perspective", "eval_point_error", "flow_to_homography_ls",
]


# ----------------------------------------------------------------------------------------------
# plumbing
# ----------------------------------------------------------------------------------------------
def _cuda(*tensors):
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise L.DmhError("dmhomo_b200 ops take CUDA tensors only (there is no CPU fallback); got a "
                             f"{t.device} tensor")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise L.DmhError(f"dmhomo_b200: tensors on different devices ({dev} vs {t.device})")
    return dev


def _f32(t):
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _p(t):
    return None if t is None else t.data_ptr()


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _start_args(start, B, dev):
    """`start` of get_grid(): a number (added to x and y), an (sx, sy) pair, or a tensor
    broadcastable to (B,2,1,1).  Returns (sx, sy, per_sample_tensor_or_None)."""
    if torch.is_tensor(start):
        s = _f32(start.to(dev)).reshape(-1)
        if s.numel() == 1:
            s = s.expand(2 * B)
        elif s.numel() == 2:
            s = s.repeat(B)
        elif s.numel() != 2 * B:
            raise ValueError(f"start tensor must broadcast to (B,2,1,1); got {tuple(start.shape)}")
        return 0.0, 0.0, s.contiguous().view(B, 2)
    if isinstance(start, (tuple, list)):
        return float(start[0]), float(start[1]), None
    return float(start), float(start), None


def _desc(sampler, kind, src, param, h, w, **kw):
    B, Cc, Hs, Ws = src.shape
    d = L.WarpDesc()
    d.struct_size = C.sizeof(L.WarpDesc)
    d.sampler, d.param_kind = sampler, kind
    d.loss_form = kw.get("loss_form", LOSS_NONE)
    d.B, d.C, d.Hs, d.Ws, d.h, d.w = B, Cc, Hs, Ws, h, w
    d.divide = kw.get("divide", 1)
    d.use_border_mask = int(bool(kw.get("use_border_mask", False)))
    d.compute_grads = int(bool(kw.get("compute_grads", False)))
    d.start_x, d.start_y = kw.get("start_x", 0.0), kw.get("start_y", 0.0)
    d.grad_loss_scale = kw.get("grad_loss_scale", 0.0)
    d.src, d.param = _p(src), _p(param)
    for name in ("basis", "start", "target", "soft_mask", "sample_weight", "grad_out", "grad_loss", "out", "valid",
                 "flow_out", "indices", "loss_acc", "grad_src", "grad_target", "grad_param", "grad_soft_mask"):
        setattr(d, name, _p(kw.get(name)))
    return d


warp_timing_events = None   # optional (begin, end) torch.cuda.Event pair recorded tightly around the next warp launches
last_warp_kernel = ""   # diagnostic: the kernel the last warp launch of this process went to (bench.py's roofline label)


def _run_warp(descs, dev, backward=False):
    arr = (L.WarpDesc * len(descs))(*descs)
    fn = L.lib().dmh_warp_backward if backward else L.lib().dmh_warp_forward
    ev = warp_timing_events
    with torch.cuda.device(dev):
        if ev is not None:
            ev[0].record()
        L.check(fn(arr, len(descs), _stream(dev)), "warp_backward" if backward else "warp_forward")
        if ev is not None:
            ev[1].record()
    global last_warp_kernel
    last_warp_kernel = L.last_kernel_name()


def _param_shape_check(kind, param, B, h, w, divide):
    if kind in (PARAM_FLOW, PARAM_COORDS):
        if tuple(param.shape) != (B, 2, h, w):
            raise ValueError(f"flow/coords must be (B,2,h,w)={(B, 2, h, w)}, got {tuple(param.shape)}")
    elif kind == PARAM_HOMOGRAPHY:
        if param.numel() != B * divide * divide * 9:
            raise ValueError(f"homography must hold B*divide^2*9 values, got {tuple(param.shape)}")
    elif kind == PARAM_BASIS8:
        if param.numel() != B * 8:
            raise ValueError(f"basis weights must hold B*8 values, got {tuple(param.shape)}")
    else:
        raise ValueError(f"bad param kind {kind}")


# ----------------------------------------------------------------------------------------------
# DLT (A1-A3)
# ----------------------------------------------------------------------------------------------
class _DLT4(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, dst):
        dev = _cuda(src, dst)
        src_c, dst_c = _f32(src), _f32(dst)
        N = src_c.numel() // 8
        H = torch.empty(N, 3, 3, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            L.check(L.lib().dmh_dlt4_forward(_p(src_c), _p(dst_c), _p(H), N, _stream(dev)), "dlt4_forward")
        ctx.save_for_backward(src_c, dst_c, H)
        ctx.shapes = (src.shape, dst.shape)
        return H

    @staticmethod
    def backward(ctx, gH):
        src, dst, H = ctx.saved_tensors
        dev = src.device
        N = H.shape[0]
        gH = _f32(gH)
        g_dst = torch.empty_like(dst)
        g_src = torch.empty_like(src) if ctx.needs_input_grad[0] else None
        with torch.cuda.device(dev):
            L.check(L.lib().dmh_dlt4_backward(_p(src), _p(dst), _p(H), _p(gH), _p(g_dst), _p(g_src), N, _stream(dev)),
                    "dlt4_backward")
        return (g_src.view(ctx.shapes[0]) if g_src is not None else None), g_dst.view(ctx.shapes[1])


def dlt4(src_pt, dst_pt):
    """N independent 4-point DLT solves: (..,4,2),(..,4,2) -> (N,3,3), H[2,2] = 1.
    Replaces DLT.forward('Axb') (HEM/model/utils.py:55-101)."""
    if src_pt.numel() % 8 or src_pt.numel() != dst_pt.numel():
        raise ValueError("dlt4: src/dst must hold N*4*2 values each")
    return _DLT4.apply(src_pt, dst_pt)


# ----------------------------------------------------------------------------------------------
# homography -> flow (A5, A15)
# ----------------------------------------------------------------------------------------------
class _H2Flow(torch.autograd.Function):
    @staticmethod
    def forward(ctx, H, h, w, divide, start):
        dev = _cuda(H)
        Hc = _f32(H)
        B = Hc.numel() // (9 * divide * divide)
        sx, sy, per = _start_args(start, B, dev)
        flow = torch.empty(B, 2, h, w, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            L.check(L.lib().dmh_homography_to_flow(_p(Hc), _p(flow), B, h, w, divide, sx, sy, _p(per), _stream(dev)),
                    "homography_to_flow")
        ctx.save_for_backward(Hc, per)
        ctx.cfg = (B, h, w, divide, sx, sy, H.shape)
        return flow

    @staticmethod
    def backward(ctx, gflow):
        Hc, per = ctx.saved_tensors
        B, h, w, divide, sx, sy, shape = ctx.cfg
        dev = Hc.device
        gflow = _f32(gflow)
        gH = torch.zeros_like(Hc)
        with torch.cuda.device(dev):
            L.check(L.lib().dmh_homography_to_flow_backward(_p(Hc), _p(gflow), _p(gH), B, h, w, divide, sx, sy,
                                                            _p(per), _stream(dev)), "homography_to_flow_backward")
        return gH.view(shape), None, None, None, None


def homography_to_flow(H, h, w, divide=1, start=0):
    """get_flow() (HEM/model/utils.py:400-440): H (B,[divide^2,]3,3) -> flow (B,2,h,w), fp32 with the
    reference's rounding order and epsilon rule.  `start` as in get_grid() (number, pair, or a
    tensor broadcastable to (B,2,1,1))."""
    return _H2Flow.apply(H, int(h), int(w), int(divide), start)


def homography_to_flow_f64(H, h, w, eps=1e-6, channels_last=True, as_mapping=False):
    """homo_to_flow()/get_flow_np() (ddpm.py:913-975; eps=1e-6) and
    from_homography_to_pixel_wise_mapping() (flow_and_mapping_operations.py:454-484; eps=1e-8,
    as_mapping=True): fp64 arithmetic, fp32 result.  as_mapping=2: the loaders' ground-truth flow
    homo_convert_to_flow (HEM/dataset/data_loader.py:42-52) = fl32(fl32(mapping) - grid).
    H: (B,3,3) float64 CUDA tensor."""
    dev = _cuda(H)
    Hc = H.to(torch.float64).contiguous()
    B = Hc.numel() // 9
    shape = (B, h, w, 2) if channels_last else (B, 2, h, w)
    out = torch.empty(shape, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        L.check(L.lib().dmh_homography_to_flow_f64(_p(Hc), _p(out), B, h, w, float(eps), int(channels_last),
                                                   int(as_mapping), _stream(dev)), "homography_to_flow_f64")
    return out


# ----------------------------------------------------------------------------------------------
# basis flows (A12)
# ----------------------------------------------------------------------------------------------
class _BasisCombine(torch.autograd.Function):
    @staticmethod
    def forward(ctx, basis, weight, h, w):
        dev = _cuda(basis, weight)
        bc, wc = _f32(basis), _f32(weight)
        B = wc.numel() // 8
        flow = torch.empty(B, 2, h, w, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            L.check(L.lib().dmh_basis_combine(_p(bc), _p(wc), _p(flow), B, h, w, _stream(dev)), "basis_combine")
        ctx.save_for_backward(bc)
        ctx.cfg = (B, h, w, weight.shape)
        return flow

    @staticmethod
    def backward(ctx, gflow):
        (bc,) = ctx.saved_tensors
        B, h, w, wshape = ctx.cfg
        dev = bc.device
        gflow = _f32(gflow)
        gw = torch.zeros(B, 8, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            L.check(L.lib().dmh_basis_combine_backward(_p(bc), _p(gflow), _p(gw), B, h, w, _stream(dev)),
                    "basis_combine_backward")
        return None, gw.view(wshape), None, None


def basis_combine(basis, weight, h, w):
    """flow = sum_k w_k basis_k (HEM/model/net.py:808-815).  basis holds 8*2*h*w values laid out
    (8,2,h,w) (== the reference's (1,8,2*h*w)); weight (B,8[,1]) -> (B,2,h,w)."""
    if basis.numel() != 16 * h * w:
        raise ValueError("basis must hold 8*2*h*w values")
    return _BasisCombine.apply(basis, weight, int(h), int(w))


class _BasisCorners(torch.autograd.Function):
    @staticmethod
    def forward(ctx, basis, weight, h, w):
        dev = _cuda(basis, weight)
        bc, wc = _f32(basis), _f32(weight)
        B = wc.numel() // 8
        off = torch.empty(B, 4, 2, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            L.check(L.lib().dmh_basis_corner_offsets(_p(bc), _p(wc), _p(off), B, h, w, _stream(dev)),
                    "basis_corner_offsets")
        ctx.save_for_backward(bc)
        ctx.cfg = (B, h, w, weight.shape)
        return off

    @staticmethod
    def backward(ctx, goff):
        (bc,) = ctx.saved_tensors
        B, h, w, wshape = ctx.cfg
        dev = bc.device
        goff = _f32(goff)
        gw = torch.zeros(B, 8, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            L.check(L.lib().dmh_basis_corner_offsets_backward(_p(bc), _p(goff), _p(gw), B, h, w, _stream(dev)),
                    "basis_corner_offsets_backward")
        return None, gw.view(wshape), None, None


def basis_corner_offsets(basis, weight, h, w):
    """The basis flow sampled at the 4 image corners (TL,TR,BL,BR) as 4-point offsets (B,4,2)."""
    if basis.numel() != 16 * h * w:
        raise ValueError("basis must hold 8*2*h*w values")
    return _BasisCorners.apply(basis, weight, int(h), int(w))


class _BasisHomography(torch.autograd.Function):
    """weights (B,8) x n_sets -> H (B,3,3) x n_sets through corner offsets and the 4-point DLT: one launch."""

    @staticmethod
    def forward(ctx, basis, h, w, *weights):
        dev = _cuda(basis, *weights)
        bc = _f32(basis)
        ws = [_f32(t).reshape(-1, 8) for t in weights]
        B = ws[0].shape[0]
        if any(t.shape[0] != B for t in ws) or not 1 <= len(ws) <= 4:
            raise ValueError("basis_homography: 1..4 weight sets of equal batch size")
        Hs = [torch.empty(B, 3, 3, device=dev, dtype=torch.float32) for _ in ws]
        n = len(ws)
        wp = (C.c_void_p * n)(*[t.data_ptr() for t in ws])
        hp = (C.c_void_p * n)(*[t.data_ptr() for t in Hs])
        with torch.cuda.device(dev):
            L.check(L.lib().dmh_basis_homography_forward(_p(bc), wp, hp, n, B, h, w, _stream(dev)),
                    "basis_homography_forward")
        ctx.save_for_backward(bc, *Hs)
        ctx.cfg = (B, h, w, [t.shape for t in weights])
        return tuple(Hs)

    @staticmethod
    def backward(ctx, *gHs):
        bc, *Hs = ctx.saved_tensors
        B, h, w, shapes = ctx.cfg
        dev = bc.device
        n = len(Hs)
        gH = [_f32(g) if g is not None else torch.zeros_like(Hs[i]) for i, g in enumerate(gHs)]
        gw = [torch.empty(B, 8, device=dev, dtype=torch.float32) for _ in range(n)]
        hp = (C.c_void_p * n)(*[t.data_ptr() for t in Hs])
        gp = (C.c_void_p * n)(*[t.data_ptr() for t in gH])
        wp = (C.c_void_p * n)(*[t.data_ptr() for t in gw])
        with torch.cuda.device(dev):
            L.check(L.lib().dmh_basis_homography_backward(_p(bc), hp, gp, wp, n, B, h, w, _stream(dev)),
                    "basis_homography_backward")
        return (None, None, None) + tuple(g.view(s) for g, s in zip(gw, shapes))


def basis_homography(basis, h, w, *weights):
    """8 basis weights -> basis flow at the 4 image corners -> 4-point DLT -> H, for up to four weight sets
    (e.g. forward and backward direction) in ONE launch.  Equivalent to
    dlt4(corners, corners + basis_corner_offsets(basis, w_i, h, w)) per set (cfg 2's "8-basis flow -> DLT")."""
    if basis.numel() != 16 * h * w:
        raise ValueError("basis must hold 8*2*h*w values")
    out = _BasisHomography.apply(basis, int(h), int(w), *weights)
    return out[0] if len(out) == 1 else out


# ----------------------------------------------------------------------------------------------
# warp (A6-A9)
# ----------------------------------------------------------------------------------------------
class _Warp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, param, basis, cfg):
        dev = _cuda(img, param, basis)
        img_c, par_c, bas_c = _f32(img), _f32(param), _f32(basis)
        B, Cc, Hs, Ws = img_c.shape
        h, w = cfg["h"], cfg["w"]
        _param_shape_check(cfg["kind"], par_c, B, h, w, cfg["divide"])
        sx, sy, per = _start_args(cfg["start"], B, dev)
        out = torch.empty(B, Cc, h, w, device=dev, dtype=torch.float32)
        valid = torch.empty(B, h, w, device=dev, dtype=torch.uint8) if cfg["mask"] else None
        flow = torch.empty(B, 2, h, w, device=dev, dtype=torch.float32) if cfg["flow"] else None
        idx = torch.empty(4, B, h, w, device=dev, dtype=torch.int32) if cfg["indices"] else None
        d = _desc(cfg["sampler"], cfg["kind"], img_c, par_c, h, w, basis=bas_c, start=per, start_x=sx, start_y=sy,
                  divide=cfg["divide"], out=out, valid=valid, flow_out=flow, indices=idx)
        _run_warp([d], dev)
        ctx.save_for_backward(img_c, par_c, bas_c, per)
        ctx.cfg = dict(cfg, sx=sx, sy=sy, pshape=param.shape)
        extras = tuple(t for t in (valid, flow, idx) if t is not None)
        ctx.mark_non_differentiable(*extras)
        return (out,) + extras

    @staticmethod
    def backward(ctx, gout, *_):
        img_c, par_c, bas_c, per = ctx.saved_tensors
        cfg = ctx.cfg
        dev = img_c.device
        h, w, kind = cfg["h"], cfg["w"], cfg["kind"]
        need_img, need_par = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g_src = torch.zeros_like(img_c) if need_img else None
        g_par = None
        if need_par:
            g_par = torch.empty_like(par_c) if kind in (PARAM_FLOW, PARAM_COORDS) else torch.zeros_like(par_c)
        d = _desc(cfg["sampler"], kind, img_c, par_c, h, w, basis=bas_c, start=per, start_x=cfg["sx"],
                  start_y=cfg["sy"], divide=cfg["divide"], grad_out=_f32(gout), grad_src=g_src, grad_param=g_par)
        _run_warp([d], dev, backward=True)
        return g_src, (g_par.view(cfg["pshape"]) if g_par is not None else None), None, None


def warp(img, param, kind=PARAM_FLOW, sampler=S1, out_hw=None, start=0, basis=None, divide=1, return_mask=False,
         return_flow=False, return_indices=False):
    """Bilinear warp of `img` (B,C,Hs,Ws) at coordinates given by `param` (see dmh_param_kind).

    Returns out (B,C,h,w); with return_mask also the M1 validity mask (B,h,w) bool; with
    return_flow the generated flow (B,2,h,w); with return_indices the clamped integer corners
    (4,B,h,w) int32 [x0,y0,x1,y1].  Differentiable w.r.t. img and param."""
    if img.dim() != 4:
        raise ValueError("img must be (B,C,H,W)")
    if kind in (PARAM_FLOW, PARAM_COORDS):
        h, w = param.shape[-2:]
    else:
        h, w = out_hw if out_hw is not None else img.shape[-2:]
    cfg = dict(sampler=sampler, kind=kind, h=int(h), w=int(w), start=start, divide=int(divide), mask=return_mask,
               flow=return_flow, indices=return_indices)
    res = _Warp.apply(img, param, basis, cfg)
    out, extras = res[0], list(res[1:])
    ret = [out]
    if return_mask:
        ret.append(extras.pop(0).view(torch.bool))      # the kernels write exactly 0 / 1: reinterpret, no copy
    if return_flow:
        ret.append(extras.pop(0))
    if return_indices:
        ret.append(extras.pop(0))
    return ret[0] if len(ret) == 1 else tuple(ret)


def warp_into(img, param, out, valid=None, kind=PARAM_HOMOGRAPHY, sampler=S1, start=0, basis=None, divide=1):
    """warp() into caller-owned buffers, no autograd: out (B,C,h,w) fp32 and, optionally, valid (B,h,w) uint8 (the M1
    mask) are overwritten.  For resident pipelines that reuse their output buffers (frame-sequence warps, cfg 5)."""
    dev = _cuda(img, param, out, valid, basis)
    img_c, par_c = _f32(img), _f32(param)
    B, Cc, _, _ = img_c.shape
    h, w = out.shape[-2:]
    if out.dtype != torch.float32 or tuple(out.shape) != (B, Cc, h, w) or not out.is_contiguous():
        raise ValueError("warp_into: out must be a contiguous fp32 (B,C,h,w) tensor")
    if valid is not None and (valid.dtype != torch.uint8 or valid.numel() != B * h * w or not valid.is_contiguous()):
        raise ValueError("warp_into: valid must be a contiguous uint8 (B,h,w) tensor")
    _param_shape_check(kind, par_c, B, h, w, divide)
    sx, sy, per = _start_args(start, B, dev)
    d = _desc(sampler, kind, img_c, par_c, h, w, basis=_f32(basis), start=per, start_x=sx, start_y=sy, divide=divide,
              out=out, valid=valid)
    _run_warp([d], dev)
    return out, valid


# ----------------------------------------------------------------------------------------------
# fused warp + mask + masked-L1 (+ gradients) (A13, A14)
# ----------------------------------------------------------------------------------------------
class WarpTerm:
    """One loss term: mean-style masked L1 between `target` and warp(`src`, `param`)."""

    def __init__(self, src, target, param, soft_mask=None, sample_weight=None):
        self.src, self.target, self.param = src, target, param
        self.soft_mask, self.sample_weight = soft_mask, sample_weight


_SLOTS = 5  # tensors per term: src, target, param, soft_mask, sample_weight


def _plan_grad_slots(flat, terms, needs, kind, base):
    """(term, slot) -> (offset, numel) into one flat gradient buffer starting at `base`.

    The kernels ACCUMULATE (atomics / TMA reduce-add) the gradients of src, target and of a HOMOGRAPHY / BASIS8
    parameter, so one tensor object used in several of those slots (img1 is the target of one term and the source
    of the other) shares one zeroed buffer and is handed back once.  Gradients of soft_mask and of a FLOW / COORDS
    parameter are plain stores: every (term, slot) gets its own buffer and autograd sums them (the reference passes
    one mask to both terms when params.normalize_mask is set, HEM/loss/losses.py:129).  Sharing is by tensor
    identity, never by storage address: two distinct leaves that alias one storage each get their own gradient.
    `flat` holds the tensors (or their ids) in apply() order.  Returns (slots, zero_end, total): only [0, zero_end)
    needs zeroing, the plain-store buffers behind it are fully overwritten by the kernels."""
    accumulated = (0, 1) + ((2,) if kind in (PARAM_HOMOGRAPHY, PARAM_BASIS8) else ())
    slots, owners, total = {}, {}, base
    zero_end = base
    for want_accumulated in (True, False):      # accumulated buffers first: only [0, zero_end) has to be zeroed
        for i, tl in enumerate(terms):
            for s in (0, 1, 2, 3):
                if tl[s] is None or not needs[i * _SLOTS + s] or (s in accumulated) != want_accumulated:
                    continue
                obj = flat[i * _SLOTS + s]
                key = (obj if isinstance(obj, int) else id(obj)) if want_accumulated else ("own", i, s)
                if key not in owners:
                    owners[key] = (total, tl[s].numel())
                    total += tl[s].numel()
                slots[(i, s)] = owners[key]
        if want_accumulated:
            zero_end = total
    return slots, zero_end, total


class _WarpLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cfg, basis, *flat):
        n = len(flat) // _SLOTS
        dev = _cuda(basis, *flat)
        bas_c = _f32(basis)
        terms = [[_f32(t) for t in flat[i * _SLOTS:(i + 1) * _SLOTS]] for i in range(n)]
        B, Cc, Hs, Ws = terms[0][0].shape
        h, w = terms[0][1].shape[-2:]
        for src, tgt, par, soft, sw in terms:
            if tuple(tgt.shape) != (B, Cc, h, w) or src.shape[:2] != (B, Cc):
                raise ValueError("warp_loss: src (B,C,Hs,Ws) / target (B,C,h,w) shapes disagree")
            _param_shape_check(cfg["kind"], par, B, h, w, cfg["divide"])
            if soft is not None and soft.numel() != B * h * w:
                raise ValueError("warp_loss: soft_mask must be (B,1,h,w)")
            if sw is not None and sw.numel() != B:
                raise ValueError("warp_loss: sample_weight must be (B,)")
        sx, sy, per = _start_args(cfg["start"], B, dev)
        scale = float(cfg["weight"]) / float(B * Cc * h * w)

        needs = ctx.needs_input_grad[2:]
        any_grad = any(needs)
        fused = bool(cfg["fused"]) and any_grad
        # one flat zeroed workspace: [loss accumulators (double) | gradients of every distinct input]
        acc_floats = 2 * n * B
        slots, zero_end, total = _plan_grad_slots(flat, terms, needs, cfg["kind"], acc_floats) if fused else ({}, acc_floats, acc_floats)
        ws = torch.empty(total, device=dev, dtype=torch.float32)
        ws[:zero_end].zero_()
        acc = ws[:acc_floats].view(torch.float64)

        def gbuf(i, s):
            if (i, s) not in slots:
                return None
            o, m = slots[(i, s)]
            return ws[o:o + m]

        descs = []
        for i, (src, tgt, par, soft, sw) in enumerate(terms):
            descs.append(_desc(cfg["sampler"], cfg["kind"], src, par, h, w, basis=bas_c, start=per, start_x=sx,
                               start_y=sy, divide=cfg["divide"], target=tgt, soft_mask=soft, sample_weight=sw,
                               use_border_mask=cfg["border_mask"], loss_form=cfg["loss_form"], grad_loss_scale=scale,
                               compute_grads=fused, loss_acc=acc[i * B:(i + 1) * B], grad_src=gbuf(i, 0),
                               grad_target=gbuf(i, 1), grad_param=gbuf(i, 2), grad_soft_mask=gbuf(i, 3)))
        _run_warp(descs, dev)
        loss = torch.empty((), device=dev, dtype=torch.float32)
        acc_ptrs = (C.c_void_p * n)(*[acc[i * B:(i + 1) * B].data_ptr() for i in range(n)])
        sw_ptrs = (C.c_void_p * n)(*[_p(tl[4]) for tl in terms])
        with torch.cuda.device(dev):
            L.check(L.lib().dmh_loss_finish(acc_ptrs, sw_ptrs, n, B, scale, _p(loss), _stream(dev)), "loss_finish")

        ctx.cfg = dict(cfg, sx=sx, sy=sy, scale=scale, n=n, fused=fused, B=B, h=h, w=w, acc_floats=acc_floats)
        ctx.slots = slots
        ctx.input_ids = [id(t) for t in flat]
        ctx.shapes = [None if t is None else t.shape for t in flat]
        # The fused pass has left the gradients (for an upstream gradient of 1) in ws: the first backward rescales them
        # in place and hands out views.  The inputs are saved as well, so that any further backward through the same
        # graph (retain_graph=True, checkpointing) recomputes with the separate backward kernel instead of rescaling
        # - and possibly handing out again - buffers autograd may already own.
        ctx.save_for_backward(bas_c, per, *[t for tl in terms for t in tl], *([ws] if fused else []))
        ctx.consumed = False
        return loss

    @staticmethod
    def backward(ctx, g):
        cfg = ctx.cfg
        n, B, h, w = cfg["n"], cfg["B"], cfg["h"], cfg["w"]
        needs = ctx.needs_input_grad[2:]
        g = _f32(g)
        dev = g.device
        if cfg["fused"] and not ctx.consumed:
            ctx.consumed = True
            ws = ctx.saved_tensors[-1]
            slots = ctx.slots
            nf = ws.numel() - cfg["acc_floats"]
            if nf > 0:
                with torch.cuda.device(dev):
                    L.check(L.lib().dmh_scale_inplace(ws[cfg["acc_floats"]:].data_ptr(), nf, _p(g), _stream(dev)),
                            "scale_inplace")
        else:
            saved = list(ctx.saved_tensors)
            bas_c, per = saved[0], saved[1]
            flat = saved[2:2 + n * _SLOTS]
            terms = [flat[i * _SLOTS:(i + 1) * _SLOTS] for i in range(n)]
            slots, zero_end, total = _plan_grad_slots(ctx.input_ids, terms, needs, cfg["kind"], 0)
            ws = torch.empty(max(total, 1), device=dev, dtype=torch.float32)
            ws[:zero_end].zero_()
            cfgacc = 0

            def gbuf(i, s):
                if (i, s) not in slots:
                    return None
                o, m = slots[(i, s)]
                return ws[o:o + m]

            descs = []
            for i, (src, tgt, par, soft, sw) in enumerate(terms):
                descs.append(_desc(cfg["sampler"], cfg["kind"], src, par, h, w, basis=bas_c, start=per,
                                   start_x=cfg["sx"], start_y=cfg["sy"], divide=cfg["divide"], target=tgt,
                                   soft_mask=soft, sample_weight=sw, use_border_mask=cfg["border_mask"],
                                   loss_form=cfg["loss_form"], grad_loss_scale=cfg["scale"], grad_loss=g,
                                   grad_src=gbuf(i, 0), grad_target=gbuf(i, 1), grad_param=gbuf(i, 2),
                                   grad_soft_mask=gbuf(i, 3)))
            _run_warp(descs, dev, backward=True)
            cfg = dict(cfg, acc_floats=cfgacc)
        # hand every distinct buffer back exactly once (autograd sums duplicates of the same input)
        grads, seen = [], set()
        for i in range(n):
            for s in range(_SLOTS):
                k = (i, s)
                if k in slots and slots[k] not in seen:
                    seen.add(slots[k])
                    o, m = slots[k]
                    grads.append(ws[o:o + m].view(ctx.shapes[i * _SLOTS + s]))
                else:
                    grads.append(None)
        return (None, None) + tuple(grads)


def warp_loss(terms, kind=PARAM_HOMOGRAPHY, sampler=S1, loss_form=LOSS_MASKED_DIFF, border_mask=True, weight=1.0,
              basis=None, divide=1, start=0, fused=True):
    """sum over `terms` of  weight * mean_{b,c,y,x}( sample_weight[b] * L(m, target, warp(src, param)) ).

    One kernel launch evaluates every term (e.g. both directions of a pair): flow generation,
    bilinear sampling, the M1 validity mask, the masked L1 and - when `fused` and an input
    requires grad - the gradients to src / target / param / soft_mask in the same pass (the
    backward then only rescales them by the upstream gradient, a no-op when it is 1).
    With fused=False gradients are produced by the separate backward kernel.
    L = |m*t - m*w| (LOSS_MASKED_DIFF, HEM/loss/losses.py:142-146) or m*|w - t| (LOSS_DIFF_MASKED,
    classifier_free_guidance.py:799-806); m = M1 (if border_mask) * soft_mask."""
    if isinstance(terms, WarpTerm):
        terms = [terms]
    flat = []
    for t in terms:
        flat += [t.src, t.target, t.param, t.soft_mask, t.sample_weight]
    cfg = dict(kind=kind, sampler=sampler, loss_form=loss_form, border_mask=bool(border_mask), weight=float(weight),
               divide=int(divide), start=start, fused=bool(fused))
    return _WarpLoss.apply(cfg, basis, *flat)


def warp_eval(terms, kind=PARAM_HOMOGRAPHY, sampler=S1, loss_form=LOSS_MASKED_DIFF, border_mask=True, weight=1.0, basis=None,
              divide=1, start=0, masks_as_bool=True):
    """Evaluation pass, one launch for all `terms` (e.g. both directions of a pair), no gradients: per term the warped
    image (get_warp_flow, HEM/model/utils.py:548-553) and the M1 validity mask (get_gt_correspondence_mask,
    flow_and_mapping_operations.py:45-71), plus the masked-L1 total of warp_loss() (HEM/loss/losses.py:142-146).
    Returns (loss scalar, [warped (B,C,h,w) per term], [mask (B,h,w) bool (uint8 with masks_as_bool=False) per term])."""
    if isinstance(terms, WarpTerm):
        terms = [terms]
    n = len(terms)
    dev = _cuda(basis, *[t for tm in terms for t in (tm.src, tm.target, tm.param, tm.soft_mask, tm.sample_weight)])
    bas_c = _f32(basis)
    with torch.no_grad():
        tl = [[_f32(t) for t in (tm.src, tm.target, tm.param, tm.soft_mask, tm.sample_weight)] for tm in terms]
        B, Cc, Hs, Ws = tl[0][0].shape
        h, w = tl[0][1].shape[-2:]
        for src, tgt, par, soft, sw in tl:
            if tuple(tgt.shape) != (B, Cc, h, w) or src.shape[:2] != (B, Cc):
                raise ValueError("warp_eval: src (B,C,Hs,Ws) / target (B,C,h,w) shapes disagree")
            _param_shape_check(kind, par, B, h, w, divide)
        sx, sy, per = _start_args(start, B, dev)
        scale = float(weight) / float(B * Cc * h * w)
        acc = torch.zeros(n * B, device=dev, dtype=torch.float64)
        outs = [torch.empty(B, Cc, h, w, device=dev, dtype=torch.float32) for _ in range(n)]
        valids = [torch.empty(B, h, w, device=dev, dtype=torch.uint8) for _ in range(n)]
        descs = [_desc(sampler, kind, src, par, h, w, basis=bas_c, start=per, start_x=sx, start_y=sy, divide=divide,
                       target=tgt, soft_mask=soft, sample_weight=sw, use_border_mask=border_mask, loss_form=loss_form,
                       grad_loss_scale=scale, loss_acc=acc[i * B:(i + 1) * B], out=outs[i], valid=valids[i])
                 for i, (src, tgt, par, soft, sw) in enumerate(tl)]
        _run_warp(descs, dev)
        loss = torch.empty((), device=dev, dtype=torch.float32)
        acc_ptrs = (C.c_void_p * n)(*[acc[i * B:(i + 1) * B].data_ptr() for i in range(n)])
        sw_ptrs = (C.c_void_p * n)(*[_p(t[4]) for t in tl])
        with torch.cuda.device(dev):
            L.check(L.lib().dmh_loss_finish(acc_ptrs, sw_ptrs, n, B, scale, _p(loss), _stream(dev)), "loss_finish")
    return loss, outs, ([v.view(torch.bool) for v in valids] if masks_as_bool else valids)


# ----------------------------------------------------------------------------------------------
# the whole unsupervised step of cfg 2 in one op: 8 basis weights -> H -> bidirectional warp loss + all gradients
# ----------------------------------------------------------------------------------------------
_side_streams = {}


def _side_stream(dev, idx=0):
    """Auxiliary streams per device for the short fork / join branches of basis_warp_loss and render_conditions."""
    key = (dev.type, dev.index, idx)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(dev)
    return _side_streams[key]


class _BasisWarpLoss(torch.autograd.Function):
    """basis_homography + warp_loss(two directions, fused gradients) + the adjoint DLT, with the independent
    pieces on forked branches (they become parallel nodes when the step is captured into a CUDA graph):

        zero the gradient workspace  ||  weights -> corner offsets -> DLT -> Hf, Hb
                            fused warp kernel (loss sums + dL/dimg + dL/dH)
        loss finish                  ||  dL/dH -> adjoint DLT -> dL/dweights

    The backward only rescales the stored gradients by the upstream gradient."""

    @staticmethod
    def forward(ctx, basis, img1, img2, w_f, w_b, weight):
        dev = _cuda(basis, img1, img2, w_f, w_b)
        bc, i1, i2 = _f32(basis), _f32(img1), _f32(img2)
        wf, wb = _f32(w_f).reshape(-1, 8), _f32(w_b).reshape(-1, 8)
        B, Cc, h, w = i1.shape
        if i2.shape != i1.shape or wf.shape[0] != B or wb.shape[0] != B or bc.numel() != 16 * h * w:
            raise ValueError("basis_warp_loss: img1/img2 (B,C,h,w), weights (B,8), basis (8,2,h,w) shapes disagree")
        needs = ctx.needs_input_grad
        any_grad = bool(needs[1] or needs[2] or needs[3] or needs[4])
        n_img, acc_floats = i1.numel(), 4 * B
        # workspace: [loss accumulators (fp64, 2 x B) | dL/dimg1 | dL/dimg2 | dL/dHf | dL/dHb | dL/dwf | dL/dwb]
        off_g1, off_g2 = acc_floats, acc_floats + n_img
        off_hf = off_g2 + n_img
        off_hb, off_wf, off_wb = off_hf + 9 * B, off_hf + 18 * B, off_hf + 26 * B
        total = off_hf + 34 * B if any_grad else acc_floats
        ws = torch.empty(total, device=dev, dtype=torch.float32)
        Hf = torch.empty(B, 3, 3, device=dev, dtype=torch.float32)
        Hb = torch.empty(B, 3, 3, device=dev, dtype=torch.float32)
        loss = torch.empty((), device=dev, dtype=torch.float32)
        cur, side = torch.cuda.current_stream(dev), _side_stream(dev)
        lib = L.lib()
        scale = float(weight) / float(B * Cc * h * w)
        with torch.cuda.device(dev):
            fork, join = torch.cuda.Event(), torch.cuda.Event()
            fork.record(cur)
            side.wait_event(fork)
            with torch.cuda.stream(side):
                ws.zero_()
                join.record(side)
            wp = (C.c_void_p * 2)(wf.data_ptr(), wb.data_ptr())
            hp = (C.c_void_p * 2)(Hf.data_ptr(), Hb.data_ptr())
            L.check(lib.dmh_basis_homography_forward(_p(bc), wp, hp, 2, B, h, w, _stream(dev)), "basis_homography_forward")
            cur.wait_event(join)
            acc = ws[:acc_floats].view(torch.float64)
            g1 = ws[off_g1:off_g1 + n_img] if any_grad else None
            g2 = ws[off_g2:off_g2 + n_img] if any_grad else None
            ghf = ws[off_hf:off_hf + 9 * B] if any_grad else None
            ghb = ws[off_hb:off_hb + 9 * B] if any_grad else None
            common = dict(use_border_mask=True, loss_form=LOSS_MASKED_DIFF, grad_loss_scale=scale, compute_grads=any_grad)
            descs = [_desc(S1, PARAM_HOMOGRAPHY, i2, Hf, h, w, target=i1, loss_acc=acc[:B], grad_src=g2, grad_target=g1,
                           grad_param=ghf, **common),
                     _desc(S1, PARAM_HOMOGRAPHY, i1, Hb, h, w, target=i2, loss_acc=acc[B:], grad_src=g1, grad_target=g2,
                           grad_param=ghb, **common)]
            _run_warp(descs, dev)
            fork2, join2 = torch.cuda.Event(), torch.cuda.Event()
            fork2.record(cur)
            side.wait_event(fork2)
            with torch.cuda.stream(side):
                acc_ptrs = (C.c_void_p * 2)(acc[:B].data_ptr(), acc[B:].data_ptr())
                sw_ptrs = (C.c_void_p * 2)(None, None)
                L.check(lib.dmh_loss_finish(acc_ptrs, sw_ptrs, 2, B, scale, _p(loss), C.c_void_p(side.cuda_stream)), "loss_finish")
                join2.record(side)
            if any_grad:
                gp = (C.c_void_p * 2)(ghf.data_ptr(), ghb.data_ptr())
                gwp = (C.c_void_p * 2)(ws[off_wf:off_wf + 8 * B].data_ptr(), ws[off_wb:off_wb + 8 * B].data_ptr())
                L.check(lib.dmh_basis_homography_backward(_p(bc), hp, gp, gwp, 2, B, h, w, _stream(dev)),
                        "basis_homography_backward")
            cur.wait_event(join2)
        ctx.cfg = (any_grad, acc_floats, n_img, off_g1, off_g2, off_wf, off_wb, B, i1.shape, w_f.shape, w_b.shape)
        ctx.save_for_backward(ws)
        ctx.consumed = False
        ctx.mark_non_differentiable(Hf, Hb)
        return loss, Hf, Hb

    @staticmethod
    def backward(ctx, g, *_):
        any_grad, acc_floats, n_img, off_g1, off_g2, off_wf, off_wb, B, ishape, wfs, wbs = ctx.cfg
        if not any_grad:
            return (None,) * 6
        if ctx.consumed:
            raise RuntimeError("basis_warp_loss: its gradients were produced by the forward pass and have been handed "
                               "out; a second backward through the same graph is not supported - use "
                               "basis_homography() + warp_loss() where retain_graph=True is needed")
        ctx.consumed = True
        (ws,) = ctx.saved_tensors
        g = _f32(g)
        dev = g.device
        with torch.cuda.device(dev):
            L.check(L.lib().dmh_scale_inplace(ws[acc_floats:].data_ptr(), ws.numel() - acc_floats, _p(g), _stream(dev)),
                    "scale_inplace")
        needs = ctx.needs_input_grad
        return (None,
                ws[off_g1:off_g1 + n_img].view(ishape) if needs[1] else None,
                ws[off_g2:off_g2 + n_img].view(ishape) if needs[2] else None,
                ws[off_wf:off_wf + 8 * B].view(wfs) if needs[3] else None,
                ws[off_wb:off_wb + 8 * B].view(wbs) if needs[4] else None,
                None)


def basis_warp_loss(basis, img1, img2, w_f, w_b, weight=1.0, return_homographies=False):
    """The unsupervised HEM term on the DLT variant of cfg 2 as ONE op:
        Hf, Hb = basis_homography(basis, h, w, w_f, w_b)
        loss   = weight * (L1(m_f*img1, m_f*warp(img2, Hf)) + L1(m_b*img2, m_b*warp(img1, Hb)))     (losses.py:142-146)
    with the gradients to both images and both weight sets computed in the same pass.  Numerically identical to
    warp_loss([WarpTerm(img2, img1, Hf), WarpTerm(img1, img2, Hb)]) after basis_homography - the same kernels - but the
    independent launches (workspace zeroing / weights -> H; loss finish / adjoint DLT) sit on forked stream branches."""
    loss, Hf, Hb = _BasisWarpLoss.apply(basis, img1, img2, w_f, w_b, float(weight))
    return (loss, Hf, Hb) if return_homographies else loss


# ----------------------------------------------------------------------------------------------
# masks, plain L1
# ----------------------------------------------------------------------------------------------
def border_mask(flow, as_float=False):
    """get_gt_correspondence_mask / create_border_mask (flow_and_mapping_operations.py:40-71)."""
    dev = _cuda(flow)
    f = _f32(flow)
    B, _, h, w = f.shape
    if as_float:
        out = torch.empty(B, h, w, device=dev, dtype=torch.float32)
        a, b = None, _p(out)
    else:
        out = torch.empty(B, h, w, device=dev, dtype=torch.uint8)
        a, b = _p(out), None
    with torch.cuda.device(dev):
        L.check(L.lib().dmh_border_mask(_p(f), a, b, B, h, w, _stream(dev)), "border_mask")
    return out if as_float else out.view(torch.bool)    # 0 / 1 bytes reinterpreted, no copy


def zero_border_mask(image, eps=1e-6):
    """define_mask_zero_borders (flow_and_mapping_operations.py:6-37): image (B,3,h,w) -> bool (B,h,w)."""
    dev = _cuda(image)
    im = _f32(image)
    B, c, h, w = im.shape
    if c != 3:
        raise ValueError("zero_border_mask expects (B,3,h,w)")
    out = torch.empty(B, h, w, device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        L.check(L.lib().dmh_zero_border_mask(_p(im), _p(out), B, h, w, float(eps), _stream(dev)), "zero_border_mask")
    return out.view(torch.bool)


class _L1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, reduction):
        dev = _cuda(a, b)
        ac, bc = _f32(a), _f32(b.expand_as(a) if b.shape != a.shape else b)
        n = ac.numel()
        acc = torch.zeros(1, device=dev, dtype=torch.float64)
        with torch.cuda.device(dev):
            L.check(L.lib().dmh_l1_sum(_p(ac), _p(bc), n, _p(acc), _stream(dev)), "l1_sum")
            scale = 1.0 / n if reduction == "mean" else 1.0
            loss = torch.empty((), device=dev, dtype=torch.float32)
            ptrs = (C.c_void_p * 1)(acc.data_ptr())
            L.check(L.lib().dmh_loss_finish(ptrs, None, 1, 1, scale, _p(loss), _stream(dev)), "loss_finish")
        ctx.save_for_backward(ac, bc)
        ctx.scale = scale
        ctx.shapes = (a.shape, b.shape)
        return loss

    @staticmethod
    def backward(ctx, g):
        ac, bc = ctx.saved_tensors
        dev = ac.device
        ga = torch.empty_like(ac) if ctx.needs_input_grad[0] else None
        gb = torch.empty_like(bc) if ctx.needs_input_grad[1] else None
        if ga is None and gb is None:
            return None, None, None
        with torch.cuda.device(dev):
            L.check(L.lib().dmh_l1_backward(_p(ac), _p(bc), ac.numel(), _p(_f32(g)), ctx.scale, _p(ga), _p(gb),
                                            _stream(dev)), "l1_backward")
        if gb is not None and ctx.shapes[1] != ctx.shapes[0]:
            gb = gb.sum_to_size(ctx.shapes[1])
        return ga, gb, None


def l1_loss(a, b, reduction="mean"):
    """nn.L1Loss(reduction)(a, b) as used by LossL1 (HEM/loss/losses.py:10-17)."""
    if reduction not in ("mean", "sum"):
        raise ValueError(f"l1_loss: unsupported reduction {reduction!r}")
    return _L1.apply(a, b, reduction)


# ----------------------------------------------------------------------------------------------
# DGM condition rendering (A16, A17), eval metric (A18), least-squares homography (next row 1)
# ----------------------------------------------------------------------------------------------
def flow_to_rgb(flow, max_flow=256.0, in_channels_last=False, out_channels_last=False):
    """flow_to_image()/visulize_flow() (ddpm.py:1471-1502).  flow (B,2,h,w) [or (B,h,w,2)] -> rgb in [0,1]."""
    dev = _cuda(flow)
    f = _f32(flow)
    if in_channels_last:
        B, h, w, _ = f.shape
    else:
        B, _, h, w = f.shape
    max_flow = max(float(max_flow), 1.0)
    out = torch.empty((B, h, w, 3) if out_channels_last else (B, 3, h, w), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        L.check(L.lib().dmh_flow_to_rgb(_p(f), _p(out), B, h, w, max_flow, int(in_channels_last),
                                        int(out_channels_last), _stream(dev)), "flow_to_rgb")
    return out


def warp_perspective(img, H, dsize, channels_last=False):
    """cv2.warpPerspective(img, H, dsize=(w,h)) with default flags, batched (ddpm.py:1520-1529).
    img (B,C,Hs,Ws) [or (B,Hs,Ws,C)] fp32 - or uint8, which takes OpenCV's fixed-point path and returns uint8
    (generate_nyps_to_single_case.py:15); H (B,3,3) (any float dtype; used as float64)."""
    dev = _cuda(img, H)
    u8 = img.dtype == torch.uint8
    im = img.contiguous() if u8 else _f32(img)
    Hc = H.to(torch.float64).contiguous()
    if channels_last:
        B, Hs, Ws, Cc = im.shape
    else:
        B, Cc, Hs, Ws = im.shape
    w, h = int(dsize[0]), int(dsize[1])
    if Hc.numel() != B * 9:
        raise ValueError("warp_perspective: H must be (B,3,3)")
    out = torch.empty((B, h, w, Cc) if channels_last else (B, Cc, h, w), device=dev, dtype=im.dtype)
    fn = L.lib().dmh_warp_perspective_u8 if u8 else L.lib().dmh_warp_perspective
    with torch.cuda.device(dev):
        L.check(fn(_p(im), _p(Hc), _p(out), B, Cc, Hs, Ws, h, w, int(channels_last), _stream(dev)), "warp_perspective")
    return out


def remap(img, coords, displacement=False, channels_last=False):
    """cv2.remap(img, map_x, map_y, INTER_LINEAR, BORDER_CONSTANT), batched (pixel_wise_mapping.py:7-52): img
    (B,C,Hs,Ws) [or (B,Hs,Ws,C)] fp32 or uint8, coords (B,2,h,w) fp32 absolute source coordinates (x, y) - or
    displacements to the pixel grid with displacement=True.  Bit-identical to OpenCV (1/32 px fixed-point coordinates)."""
    dev = _cuda(img, coords)
    u8 = img.dtype == torch.uint8
    im = img.contiguous() if u8 else _f32(img)
    mp = _f32(coords)
    if channels_last:
        B, Hs, Ws, Cc = im.shape
    else:
        B, Cc, Hs, Ws = im.shape
    if mp.dim() != 4 or mp.shape[0] != B or mp.shape[1] != 2:
        raise ValueError("remap: coords must be (B,2,h,w)")
    h, w = int(mp.shape[2]), int(mp.shape[3])
    out = torch.empty((B, h, w, Cc) if channels_last else (B, Cc, h, w), device=dev, dtype=im.dtype)
    fn = L.lib().dmh_remap_u8 if u8 else L.lib().dmh_remap
    with torch.cuda.device(dev):
        L.check(fn(_p(im), _p(mp), _p(out), B, Cc, Hs, Ws, h, w, int(channels_last), int(bool(displacement)), _stream(dev)), "remap")
    return out


def render_conditions(im2, homo, max_flow=256.0, timing_events=None):
    """Everything the DGM condition rendering derives from one batch (im2 (B,3,h,w) in [0,1], condition homographies
    (B,3,3) at that resolution) in one call: cv2.warpPerspective(im2, homo) (postProcess_cv2, ddpm.py:1520-1529), the
    fp64 homography flow (homo_to_flow, ddpm.py:913-975), its colour wheel image (flow_to_image, ddpm.py:1471-1502) and
    flow_warp(im2, flow) (postProcess, ddpm.py:1505-1518).  The four launches are the same kernels the separate calls
    use; here the independent ones sit on forked stream branches (parallel nodes once captured into a CUDA graph):

        warpPerspective                 ||  homography -> flow  ->  flow_warp (S3)
                                        ||                      ->  flow -> RGB

    Returns {"warp": (B,3,h,w), "flow": (B,2,h,w), "flow_rgb": (B,3,h,w), "flow_warp": (B,3,h,w)}.  No autograd.
    timing_events: optional (begin, end) events recorded around the warpPerspective launch on its own stream."""
    dev = _cuda(im2, homo)
    im = _f32(im2)
    Hc = homo.to(torch.float64).contiguous()
    B, Cc, h, w = im.shape
    if Hc.numel() != B * 9:
        raise ValueError("render_conditions: homo must be (B,3,3)")
    lib = L.lib()
    cur, s_a, s_b = torch.cuda.current_stream(dev), _side_stream(dev, 0), _side_stream(dev, 1)
    with torch.cuda.device(dev), torch.no_grad():
        warp_out = torch.empty(B, Cc, h, w, device=dev, dtype=torch.float32)      # every output belongs to the caller's stream
        flow = torch.empty(B, 2, h, w, device=dev, dtype=torch.float32)
        rgb = torch.empty(B, 3, h, w, device=dev, dtype=torch.float32)
        fwarp = torch.empty(B, Cc, h, w, device=dev, dtype=torch.float32)
        fork, have_flow, join_a, join_b = (torch.cuda.Event() for _ in range(4))
        fork.record(cur)
        s_a.wait_event(fork)
        with torch.cuda.stream(s_a):
            if timing_events is not None:
                timing_events[0].record(s_a)
            L.check(lib.dmh_warp_perspective(_p(im), _p(Hc), _p(warp_out), B, Cc, h, w, h, w, 0, C.c_void_p(s_a.cuda_stream)),
                    "warp_perspective")
            if timing_events is not None:
                timing_events[1].record(s_a)
            join_a.record(s_a)
        L.check(lib.dmh_homography_to_flow_f64(_p(Hc), _p(flow), B, h, w, 1e-6, 0, 0, _stream(dev)), "homography_to_flow_f64")
        have_flow.record(cur)
        s_b.wait_event(have_flow)
        with torch.cuda.stream(s_b):
            L.check(lib.dmh_flow_to_rgb(_p(flow), _p(rgb), B, h, w, max(float(max_flow), 1.0), 0, 0, C.c_void_p(s_b.cuda_stream)),
                    "flow_to_rgb")
            join_b.record(s_b)
        warp_into(im, flow, fwarp, kind=PARAM_FLOW, sampler=S3_BORDER)
        cur.wait_event(join_a)
        cur.wait_event(join_b)
    return {"warp": warp_out, "flow": flow, "flow_rgb": rgb, "flow_warp": fwarp}


def eval_point_error(pts, flow_f, flow_b):
    """compute_eval_results() (HEM/loss/losses.py:263-296): pts (B,P,2,2), flows (B,h,w,2) -> (B,).
    With flow_b=None only the forward error is returned (ComputeErrFlow, losses.py:208-211)."""
    dev = _cuda(pts, flow_f, flow_b)
    p, ff, fb = _f32(pts), _f32(flow_f), _f32(flow_b)  # flow_b None: forward direction only
    B, P = p.shape[:2]
    _, h, w, _ = ff.shape
    err = torch.empty(B, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        L.check(L.lib().dmh_eval_point_error(_p(p), _p(ff), _p(fb), _p(err), B, P, h, w, _stream(dev)),
                "eval_point_error")
    return err


def flow_to_homography_ls(flow):
    """homo_gen() (ddpm.py:1647-1661): least-squares DLT over all pixels.  (B,2,h,w) -> (B,1,3,3) float64."""
    dev = _cuda(flow)
    f = _f32(flow)
    B, _, h, w = f.shape
    H = torch.empty(B, 1, 3, 3, device=dev, dtype=torch.float64)
    ws = torch.empty(B * 45, device=dev, dtype=torch.float64)
    with torch.cuda.device(dev):
        L.check(L.lib().dmh_flow_to_homography_ls(_p(f), _p(H), _p(ws), B, h, w, _stream(dev)),
                "flow_to_homography_ls")
    return H


# ----------------------------------------------------------------------------------------------
# data formats either side of the path
# ----------------------------------------------------------------------------------------------
MEAN_I = (118.93, 113.97, 102.60)   # HEM/dataset/data_loader.py:103-104
STD_I = (69.85, 68.81, 72.45)


def pairs_u8_to_gray(img12, start=None, patch_size=None, want_rgb=True, mean=MEAN_I, std=STD_I, want_full=True,
                     patch_planar=False, patch_out=None):
    """The on-disk pair format batched as uint8 (B,6,H,W) (CUDA) -> (imgs_gray_full (B,2,H,W), imgs_gray_patch
    (B,2,ph,pw) or None, imgs_rgb_full (B,6,H,W) or None) exactly as DGMTrainData.__getitem__ / data_aug produce them
    on the host in numpy fp64 (HEM/dataset/data_loader.py:121-146, 217-255).  start: (B,2) int (x, y) crop origins.
    patch_planar: the patch comes back as (2,B,ph,pw) (image 1 / image 2 as two dense batches, what the warp ops
    take); patch_out: write the patch into this fp32 buffer of the right size."""
    dev = _cuda(img12)
    if img12.dtype != torch.uint8 or img12.dim() != 4 or img12.shape[1] != 6:
        raise ValueError("pairs_u8_to_gray: img12 must be uint8 (B,6,H,W)")
    x = img12.contiguous()
    B, _, H, W = x.shape
    full = torch.empty(B, 2, H, W, device=dev, dtype=torch.float32) if want_full else None
    rgb = torch.empty(B, 6, H, W, device=dev, dtype=torch.float32) if want_rgb else None
    patch, st, ph, pw = None, None, 0, 0
    if patch_size is not None:
        if start is None:
            raise ValueError("pairs_u8_to_gray: a patch needs start (B,2)")
        ph, pw = int(patch_size[0]), int(patch_size[1])
        if not (torch.is_tensor(start) and start.is_cuda):
            # host-side crop origins are validated (no device round trip); device-resident ones are trusted - the
            # kernel only ever writes patch pixels that exist, a window leaving the image leaves them unwritten
            sh = torch.as_tensor(start).reshape(B, 2)
            if bool(((sh[:, 0] < 0) | (sh[:, 0] + pw > W) | (sh[:, 1] < 0) | (sh[:, 1] + ph > H)).any()):
                raise ValueError("pairs_u8_to_gray: crop window outside the image")
        st = torch.as_tensor(start, device=dev).to(torch.int32).reshape(B, 2).contiguous()
        shape = (2, B, ph, pw) if patch_planar else (B, 2, ph, pw)
        if patch_out is not None:
            if patch_out.dtype != torch.float32 or patch_out.numel() != 2 * B * ph * pw or not patch_out.is_contiguous():
                raise ValueError("pairs_u8_to_gray: patch_out must be a contiguous fp32 buffer of 2*B*ph*pw values")
            patch = patch_out.view(shape)
        else:
            # device-resident origins are not validated on the host: pixels of a window that leaves the image stay 0
            patch = (torch.zeros if (torch.is_tensor(start) and start.is_cuda) else torch.empty)(shape, device=dev, dtype=torch.float32)
    if full is None and rgb is None and patch is None:
        raise ValueError("pairs_u8_to_gray: no output requested")
    m3, s3 = (C.c_double * 3)(*[float(v) for v in mean]), (C.c_double * 3)(*[float(v) for v in std])
    with torch.cuda.device(dev):
        L.check(L.lib().dmh_pairs_u8_to_gray(_p(x), _p(st), _p(full), _p(patch), _p(rgb), m3, s3, B, H, W, ph, pw,
                                             int(bool(patch_planar)), _stream(dev)), "pairs_u8_to_gray")
    return full, patch, rgb


def u8_to_f32(src, scale=1.0 / 255.0, bias=0.0, out=None):
    """float32(u8) * scale + bias (separately rounded): uint8 frames / grey patches shipped over PCIe at one byte per
    pixel and expanded in HBM (torch.Tensor(img).float() / 255 of the loaders, HEM/dataset/data_loader.py:139-146)."""
    dev = _cuda(src, out)
    if src.dtype != torch.uint8:
        raise ValueError("u8_to_f32: src must be uint8")
    x = src.contiguous()
    if out is None:
        out = torch.empty(x.shape, device=dev, dtype=torch.float32)
    elif out.dtype != torch.float32 or out.numel() != x.numel() or not out.is_contiguous():
        raise ValueError("u8_to_f32: out must be a contiguous fp32 tensor of the same size")
    with torch.cuda.device(dev):
        L.check(L.lib().dmh_u8_to_f32(_p(x), _p(out), x.numel(), float(scale), float(bias), _stream(dev)), "u8_to_f32")
    return out


def grid_normalize(t, mode):
    """normalize (mode 0) / unnormalize (1) / unnormalize-and-subtract-grid (2) of a channel-first (B,2,H,W) tensor
    (HEM/utils_operations/flow_and_mapping_operations.py:227-451)."""
    dev = _cuda(t)
    x = _f32(t)
    if x.dim() != 4 or x.shape[1] != 2:
        raise ValueError("grid_normalize: expected (B,2,H,W)")
    B, _, H, W = x.shape
    out = torch.empty_like(x)
    with torch.cuda.device(dev):
        L.check(L.lib().dmh_grid_normalize(_p(x), _p(out), B, H, W, int(mode), _stream(dev)), "grid_normalize")
    return out


class _FlowUpsample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flow, ho, wo, if_rate, align_corners):
        dev = _cuda(flow)
        f = _f32(flow)
        if f.dim() != 4 or f.shape[1] != 2:
            raise ValueError("flow_upsample: flow must be (B,2,h,w)")
        B, _, hi, wi = f.shape
        out = torch.empty(B, 2, ho, wo, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            L.check(L.lib().dmh_flow_upsample(_p(f), _p(out), B, hi, wi, ho, wo, int(if_rate), int(align_corners),
                                              _stream(dev)), "flow_upsample")
        ctx.cfg = (B, hi, wi, ho, wo, int(if_rate), int(align_corners))
        return out

    @staticmethod
    def backward(ctx, g):
        B, hi, wi, ho, wo, if_rate, align = ctx.cfg
        gc = _f32(g)
        gin = torch.empty(B, 2, hi, wi, device=gc.device, dtype=torch.float32)
        with torch.cuda.device(gc.device):
            L.check(L.lib().dmh_flow_upsample_backward(_p(gc), _p(gin), B, hi, wi, ho, wo, if_rate, align,
                                                       _stream(gc.device)), "flow_upsample_backward")
        return gin, None, None, None, None


def flow_upsample(flow, size, if_rate=False, align_corners=True):
    """upsample2d_flow_as without the in-place side effect: bilinear resize of a flow field (B,2,h,w) to `size`
    = (ho, wo); if_rate scales channel 0 by wo/w and channel 1 by ho/h (HEM/model/utils.py:556-572)."""
    return _FlowUpsample.apply(flow, int(size[0]), int(size[1]), bool(if_rate), bool(align_corners))
