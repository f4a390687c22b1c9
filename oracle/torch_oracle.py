"""TEST INFRASTRUCTURE ONLY — the hot path restated with the same ATen ops the reference calls.

Device-agnostic (the reference hard-codes `.cuda()`), so it serves three purposes:
  * on CPU: the multi-threaded CPU baseline (`bench.py --impl reference`, `cpu_baseline`) — the
    reference is pure PyTorch, so "the reference's own CPU implementation" is exactly this op
    sequence (matmul, avg_pool2d, grid_sample, gather ...);
  * on the GPU box: the "plain PyTorch fp32 reference" the CUDA kernels are compared with at
    full size, where the numpy oracle is too slow — same ATen CUDA kernels the reference would run;
  * pinned, like np_oracle, against tests/golden (generated from the unmodified reference).
Only tests/, smoke() and bench.py's baseline legs import it; the product never does.
Each function cites the reference lines it follows (paths relative to PriOr-RAFT/).
"""
from __future__ import annotations

import math
from typing import List, Sequence

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- samplers
def _px_to_norm(p: torch.Tensor, size: int) -> torch.Tensor:
    return 2 * p / (size - 1) - 1


def sample_px(img: torch.Tensor, pts: torch.Tensor, cyclic: bool) -> torch.Tensor:
    """core/utils/utils.py:61-95 (both wrappers).  img [B,C,H,W]; pts [B,Ho,Wo,2] in pixels."""
    H, W = img.shape[-2:]
    x, y = pts[..., 0:1], pts[..., 1:2]
    if cyclic:
        x = x % W
    grid = torch.cat([_px_to_norm(x, W), _px_to_norm(y, H)], dim=-1)
    return F.grid_sample(img, grid, align_corners=True)


def cycle_bilinear_sampler(img, pts):
    return sample_px(img, pts, True)


def bilinear_sampler(img, pts):
    return sample_px(img, pts, False)


def coords_grid(batch: int, ht: int, wd: int, device) -> torch.Tensor:
    """core/utils/utils.py:98-101."""
    ys, xs = torch.meshgrid(torch.arange(ht, device=device), torch.arange(wd, device=device), indexing="ij")
    return torch.stack([xs, ys], dim=0).float()[None].repeat(batch, 1, 1, 1)


# ----------------------------------------------------------------------------- volume + pyramid
def corr_volume(fmap1: torch.Tensor, fmap2: torch.Tensor) -> torch.Tensor:
    """core/prior_raft.py:69-75."""
    B, C, h, w = fmap1.shape
    v = torch.matmul(fmap1.reshape(B, C, h * w).transpose(1, 2), fmap2.reshape(B, C, h * w))
    return v.view(B, h, w, h, w) / torch.sqrt(torch.tensor(C).float())


def build_pyramid(volume: torch.Tensor, num_levels: int = 4) -> List[torch.Tensor]:
    """core/corr.py:99-111."""
    B, h, w, h2, w2 = volume.shape
    lvl = volume.reshape(B * h * w, 1, h2, w2)
    pyr = [lvl]
    for _ in range(num_levels - 1):
        lvl = F.avg_pool2d(lvl, 2, stride=2)
        pyr.append(lvl)
    return pyr


# ----------------------------------------------------------------------------- lookups
def _window(coords: torch.Tensor, level: int, radius: int) -> torch.Tensor:
    """core/corr.py:120-126 — [B*h*w, k, k, 2]; entry [n,a,b] = (cx/2^l + a - r, cy/2^l + b - r)."""
    B, _, h, w = coords.shape
    d = torch.linspace(-radius, radius, 2 * radius + 1, device=coords.device)
    delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1)
    centre = coords.permute(0, 2, 3, 1).reshape(B * h * w, 1, 1, 2) / 2 ** level
    return centre + delta[None]


def dccl_lookup(coords, pyr_own: Sequence[torch.Tensor], pyr_other: Sequence[torch.Tensor], grid_w2c, grid_c2w,
                radius: int = 4):
    """core/corr.py:113-144 — DCCL.__call__ -> (out_own, out_other), each [B, L*k*k, h, w]."""
    B, _, h, w = coords.shape
    k = 2 * radius + 1
    own_out, other_out = [], []
    for lvl in range(len(pyr_own)):
        win = _window(coords, lvl, radius)
        own_out.append(sample_px(pyr_own[lvl], win, True).view(B, h, w, k * k))
        mapped = sample_px(grid_w2c, win.reshape(B, h * w, k * k, 2), True)          # [B,2,hw,kk]
        mapped = mapped.permute(0, 2, 3, 1).reshape(B * h * w, k, k, 2)
        raw = sample_px(pyr_other[lvl], mapped, True).view(B, h, w, k * k).permute(0, 3, 1, 2)
        other_out.append(img_rotate(raw, grid_c2w).permute(0, 2, 3, 1))
    cat = lambda xs: torch.cat(xs, dim=-1).permute(0, 3, 1, 2).contiguous().float()
    return cat(own_out), cat(other_out)


def corrblock_lookup(coords, pyramid: Sequence[torch.Tensor], radius: int = 4):
    """core/corr.py:30-51 — CorrBlock.__call__."""
    B, _, h, w = coords.shape
    outs = [sample_px(vol, _window(coords, lvl, radius), False).view(B, h, w, -1) for lvl, vol in enumerate(pyramid)]
    return torch.cat(outs, dim=-1).permute(0, 3, 1, 2).contiguous().float()


# ----------------------------------------------------------------------------- feature warp
def warp_groupcorr(fmap1, fmap2, coords, num_groups: int = 4):
    """core/prior_raft.py:173-174 + :77-83."""
    B, C, H, W = fmap1.shape
    warped = sample_px(fmap2, coords.permute(0, 2, 3, 1), True)
    return (fmap1 * warped).view(B, num_groups, C // num_groups, H, W).mean(dim=2)


# ----------------------------------------------------------------------------- ERP geometry
def rotation_matrix(theta_list=None, axis_list=None, device="cpu") -> torch.Tensor:
    """core/utils/projection_prim_ortho.py:23-48."""
    axis_list = ["z", "y", "x"] if axis_list is None else axis_list
    theta_list = [0.0, 0.0, 0.0] if theta_list is None else theta_list
    R = torch.eye(3)
    for axis, theta in zip(axis_list, theta_list):
        c = torch.cos(torch.tensor(theta)).float().item()
        s = torch.sin(torch.tensor(theta)).float().item()
        M = {"x": [[1, 0, 0], [0, c, -s], [0, s, c]],
             "y": [[c, 0, s], [0, 1, 0], [-s, 0, c]],
             "z": [[c, -s, 0], [s, c, 0], [0, 0, 1]]}[axis]
        R = R @ torch.tensor(M, dtype=torch.float32)
    return R.to(device)


def _nudge_zero(t: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    return t + torch.sign(t) * (t.abs() < eps) * eps


def generate_samplegrid(size, R: torch.Tensor) -> torch.Tensor:
    """core/utils/projection_prim_ortho.py:432-443 (+ helpers :10-20,:51-89,:247-261,:397-429)."""
    B, _, H, W = size
    dev = R.device
    m = torch.arange(0, W, device=dev).view(1, -1).repeat(H, 1).float()
    n = torch.arange(0, H, device=dev).view(-1, 1).repeat(1, W).float()
    theta = ((m + 0.5) / W - 0.5) * 2 * math.pi
    phi = (0.5 - (n + 0.5) / H) * math.pi
    xyz = torch.stack([torch.cos(phi) * torch.cos(theta), torch.cos(phi) * torch.sin(theta), torch.sin(phi)], dim=0)
    rot = torch.matmul(R.view(1, 1, 3, 3), xyz.permute(1, 2, 0).unsqueeze(-1)).squeeze(-1)  # [H,W,3]
    phi2 = torch.arcsin(rot[..., 2])
    theta2 = torch.atan2(_nudge_zero(rot[..., 1]), _nudge_zero(rot[..., 0]))
    m2 = (theta2 / (2 * math.pi) + 0.5) * W - 0.5
    n2 = (0.5 - phi2 / math.pi) * H - 0.5
    return torch.stack([m2, n2], dim=0)[None].repeat(B, 1, 1, 1)


def img_rotate(image: torch.Tensor, sample_grid: torch.Tensor) -> torch.Tensor:
    """core/utils/projection_prim_ortho.py:507-514."""
    return sample_px(image, sample_grid.permute(0, 2, 3, 1), True)


def cycle_grid_sample(src: torch.Tensor, grid: torch.Tensor, is_grid: bool = False) -> torch.Tensor:
    """core/utils/my_cycle_sample.py:6-97 (input grid is not modified)."""
    B, C, H, W = src.shape
    Hg, Wg = grid.shape[-2:]
    flat = src.reshape(B, C, H * W)
    g = grid.reshape(B, 2, -1)
    gx, gy = g[:, 0] % W, g[:, 1]
    fx, fy = gx.floor(), gy.floor()
    xw, yw = gx - fx, gy - fy
    wa, wb, wc, wd = (1 - xw) * (1 - yw), (1 - xw) * yw, xw * (1 - yw), xw * yw
    x0, y0 = fx.long(), fy.long()
    x1 = (x0 + 1) % W
    x0 = x0 % W
    y1 = torch.clamp(y0 + 1, 0, H - 1)
    y0 = torch.clamp(y0, 0, H - 1)
    take = lambda yy, xx: torch.gather(flat, 2, (yy * W + xx).unsqueeze(1).expand(B, C, -1))
    Ia, Ib, Ic, Id = take(y0, x0), take(y1, x0), take(y0, x1), take(y1, x1)
    if is_grid:
        def recentre(I):
            m = Ia[:, 0] + ((I[:, 0] - Ia[:, 0]) + W / 2) % W - W / 2
            return torch.cat([m.unsqueeze(1), I[:, 1:]], dim=1)
        Ib, Ic, Id = recentre(Ib), recentre(Ic), recentre(Id)
    out = wa.unsqueeze(1) * Ia + wb.unsqueeze(1) * Ib + wc.unsqueeze(1) * Ic + wd.unsqueeze(1) * Id
    return out.reshape(B, C, Hg, Wg).contiguous()


def flo_rotate(flow: torch.Tensor, grid_w2c: torch.Tensor, grid_c2w: torch.Tensor) -> torch.Tensor:
    """core/utils/projection_prim_ortho.py:531-546 (flow2endpoint :200-218, u_clip :234-244)."""
    B, _, H, W = flow.shape
    end = coords_grid(B, H, W, flow.device) + flow
    ex = (end[:, 0] + 0.5) % W - 0.5
    ey = torch.clamp(end[:, 1], min=-0.5, max=H - 0.5)
    end_c = cycle_grid_sample(grid_w2c, torch.stack([ex, ey], dim=1), is_grid=True)
    flow_c = end_c - grid_w2c
    fm = (flow_c[:, 0] + W / 2) % W - W / 2
    flow_c = torch.stack([fm, flow_c[:, 1]], dim=1)
    return cycle_grid_sample(flow_c, grid_c2w, is_grid=False)
