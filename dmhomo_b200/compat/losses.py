"""Drop-in for the warped-loss terms of HEM/loss/losses.py."""
import torch
import torch.nn as nn

from .. import ops

__all__ = ["LossL1", "unsup_loss", "ComputeErrFlow", "compute_eval_results"]


class LossL1(nn.Module):
    """HEM/loss/losses.py:10-17: nn.L1Loss(reduction)(input, target)."""

    def __init__(self, reduction="mean"):
        super().__init__()
        if reduction not in ("mean", "sum"):
            raise ValueError(f"{reduction} is not a valid value for reduction")
        self.reduction = reduction

    def forward(self, input, target):
        return ops.l1_loss(input, target, self.reduction)


def unsup_loss(img1_fea, img2_fea, param_f, param_b, mask_f=None, mask_b=None, weight=1.0, kind=ops.PARAM_FLOW,
               basis=None, border_mask=False, fused=True):
    """compute_losses()['unsup'] (HEM/loss/losses.py:142-146) *including* the two get_warp_flow
    calls that feed it (HEM/model/net.py:817-818), as one fused launch:

        weight * ( L1(mask_f*img1_fea, mask_f*warp(img2_fea, f)) + L1(mask_b*img2_fea, mask_b*warp(img1_fea, b)) )

    param_f / param_b are flows (B,2,h,w), homographies (B,3,3) or basis weights (B,8[,1]) per `kind`."""
    terms = [ops.WarpTerm(img2_fea, img1_fea, param_f, soft_mask=mask_f),
             ops.WarpTerm(img1_fea, img2_fea, param_b, soft_mask=mask_b)]
    return ops.warp_loss(terms, kind=kind, sampler=ops.S1, loss_form=ops.LOSS_MASKED_DIFF, border_mask=border_mask,
                         weight=weight, basis=basis, fused=fused)


def ComputeErrFlow(src, dst, flow):
    """HEM/loss/losses.py:208-211 for one point: || dst - (src + flow[int(y), int(x)]) ||."""
    pts = torch.stack([src, dst], 0).view(1, 1, 2, 2)
    f = flow.unsqueeze(0)
    return ops.eval_point_error(pts, f, None)[0]


def compute_eval_results(data_batch, output_batch):
    """HEM/loss/losses.py:263-296: list of per-sample mean point errors (min over fwd / bwd flow)."""
    pts = data_batch["pt_set"]
    ff, fb = output_batch["flow_f"], output_batch["flow_b"]
    err = ops.eval_point_error(pts[:, :6].to(ff.device), ff, fb)
    return list(err.unbind(0))
