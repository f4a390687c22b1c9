"""Training step around the hot path (BASELINE config 5; reference: train_flow.py:94-203).

What is on the path: the ground-truth flow is rotated into the orthogonal view with the fused flo_rotate kernel
(train_flow.py:124 -> flo_A2B), the forward runs the lookup / volume / warp kernels with autograd attached, and the
backward runs their adjoint kernels (pf_lookup_dual_bwd, pf_pyramid_fold_bwd, pf_warp_groupcorr_bwd, pf_remap_bwd)
plus two cuBLAS GEMMs for dV -> dF.  Multi-GPU is DDP over NCCL (one gradient all-reduce per step) instead of the
reference's single-process nn.DataParallel.  The loss is the reference's latitude-weighted sequence L1
(uniform_loss, train_flow.py:55-79; weights from core/utils/spherical.py:11-17).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch

from . import geometry as geo
from . import ops
from .distributed import Context, ddp_loss_scale

MAX_FLOW = 400.0


def latitude_weights(H: int, W: int, device) -> torch.Tensor:
    """cos(latitude) per ERP row, normalised to sum 1 over the image (spherical_mask)."""
    n = torch.arange(H, device=device, dtype=torch.float32)
    phi = (0.5 - (n + 0.5) / H) * math.pi
    m = torch.cos(phi).view(H, 1).expand(H, W)
    return (m / m.sum())[None]


def sequence_loss(preds: Sequence[torch.Tensor], flow_gt: torch.Tensor, valid: torch.Tensor, weights: torch.Tensor,
                  gamma: float = 0.8) -> Tuple[torch.Tensor, Dict[str, float]]:
    mag = torch.sum(flow_gt ** 2, dim=1).sqrt()
    ok = ((valid >= 0.5) & (mag < MAX_FLOW)).float()
    loss = flow_gt.new_zeros(())
    n = len(preds)
    fused = flow_gt.is_cuda and all(p.dtype == torch.float32 for p in preds)
    lat = weights.reshape(weights.shape[-2], -1)[:, 0].contiguous() if fused else None    # [H]: the weights are constant along a row
    for i, p in enumerate(preds):
        if fused:      # one launch per term (and one in backward) instead of sub / abs / sum / mul / mul / sum
            loss = loss + ops.uniform_loss_term(p, flow_gt, ok, lat, gamma ** (n - i - 1))
        else:
            loss = loss + gamma ** (n - i - 1) * torch.sum(ok * weights * torch.sum((p - flow_gt).abs(), dim=1))
    epe = torch.sum((preds[-1].detach() - flow_gt) ** 2, dim=1).sqrt()[ok > 0]
    return loss, {"epe": float(epe.mean()) if epe.numel() else float("nan")}


def train_step(model: torch.nn.Module, optimizer: torch.optim.Optimizer, batch, ctx: Context, iters: int = 12,
               clip: float = 1.0, scaler: "torch.amp.GradScaler | None" = None) -> Dict[str, float]:
    """One optimisation step (train_flow.py:118-146).  `model` may be DDP-wrapped."""
    image1, image2, flow_gt, valid = batch
    with torch.no_grad():   # ground truth in the orthogonal view: full-resolution flo_rotate (train_flow.py:124-126)
        flow_gt_B = geo.flo_A2B(flow_gt)
        valid_B = ((flow_gt_B[:, 0].abs() < 1000) & (flow_gt_B[:, 1].abs() < 1000)).float()
    optimizer.zero_grad(set_to_none=True)
    preds_A, preds_B = model(image1, image2, iters=iters)
    H, W = image1.shape[-2:]
    wts = latitude_weights(H, W, image1.device)
    loss_A, met_A = sequence_loss(preds_A, flow_gt, valid, wts)
    loss_B, met_B = sequence_loss(preds_B, flow_gt_B, valid_B, wts)
    loss = (loss_A + loss_B) * ddp_loss_scale(ctx)   # DDP averages; the reference's DataParallel sums
    if scaler is not None:
        scaler.scale(loss).backward()
        scaler.unscale_(optimizer)
    else:
        loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), clip)
    if scaler is not None:
        scaler.step(optimizer)
        scaler.update()
    else:
        optimizer.step()
    return {"loss": float(loss.detach()) / ddp_loss_scale(ctx), "epe_A": met_A["epe"], "epe_B": met_B["epe"]}


def spherical_epe(flow_pred: torch.Tensor, flow_gt: torch.Tensor, radius: float = 1.0) -> torch.Tensor:
    """SEPE map of evaluate.py:354 — great-circle distance between predicted and true ERP endpoints (core/utils/spherical.py:20-53)."""
    return ops.great_circle_distance(flow_pred.float(), flow_gt.float(), radius)
