"""Host-side mirror of the geometry / sampler helpers `PriOr_RAFT.forward` resolves by name:

  core/utils/utils.py                  cycle_bilinear_sampler (:78-95), bilinear_sampler (:61-75), coords_grid (:98-101)
  core/utils/projection_prim_ortho.py  generate_rotation_metrix (:23-48), generate_samplegrid (:432-443),
                                       img_rotate (:507-514), flo_rotate (:531-546) and the thin wrappers
                                       img_A2B / img_B2A / flo_A2B / flo_B2A (:517-524, :563-570)

Same names, argument meaning and return conventions (including the reference's spelling `metrix`), backed by
the sm_100a kernels.  Sample grids depend only on (size, rotation), never on the input, yet the reference
rebuilds all eight of them on every forward (core/prior_raft.py:115-125); here they are cached per device.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops

A2B = (0.0, 0.0, -math.pi / 2)   # primitive ERP (A) -> orthogonal view (B): Rx(-pi/2)   (prior_raft.py:115)
B2A = (0.0, 0.0, math.pi / 2)    #                                           Rx(+pi/2)   (prior_raft.py:121)


# ---------------------------------------------------------------------------- rotation matrices (host)
def rotation_matrix_host(theta_list: Optional[Sequence[float]] = None,
                         axis_list: Optional[Sequence[str]] = None) -> torch.Tensor:
    """R = prod_k R_axis_k(theta_k) as a CPU fp32 tensor, built with the reference's torch ops (fp32 cos/sin of
    `torch.tensor(theta)`, fp32 3x3 products) so the entries are bit-identical, e.g. cos(-pi/2) = -4.3711e-08."""
    axis_list = ["z", "y", "x"] if axis_list is None else list(axis_list)
    theta_list = [0.0, 0.0, 0.0] if theta_list is None else list(theta_list)
    R = torch.eye(3)
    for axis, theta in zip(axis_list, theta_list):
        c = torch.cos(torch.tensor(theta)).float().item()
        s = torch.sin(torch.tensor(theta)).float().item()
        if axis == "x":
            M = [[1, 0, 0], [0, c, -s], [0, s, c]]
        elif axis == "y":
            M = [[c, 0, s], [0, 1, 0], [-s, 0, c]]
        elif axis == "z":
            M = [[c, -s, 0], [s, c, 0], [0, 0, 1]]
        else:
            continue
        R = R @ torch.tensor(M, dtype=torch.float32)
    return R


_rot_cache: Dict[Tuple, torch.Tensor] = {}      # (axes, thetas, device) -> device tensor (shared, read-only)
_rot_host: Dict[Tuple, torch.Tensor] = {}       # (data_ptr, strides, device) -> the same matrix on the host


def _remember_host(R_dev: torch.Tensor, R_host: torch.Tensor) -> None:
    dev = (R_dev.device.type, R_dev.device.index)
    _rot_host[(R_dev.data_ptr(), tuple(R_dev.stride()), dev)] = R_host.contiguous()
    _rot_host[(R_dev.data_ptr(), tuple(R_dev.T.stride()), dev)] = R_host.T.contiguous()   # `R.T` is a view of the same storage


def _host_matrix(R: torch.Tensor) -> torch.Tensor:
    """The 3x3 matrix on the host without a device synchronisation when it is one `generate_rotation_metrix` handed out
    (or its `.T` view) — which keeps a forward of the patched reference free of D2H copies, i.e. CUDA-graph capturable."""
    if not R.is_cuda:
        return R.detach().float().contiguous()
    hit = _rot_host.get((R.data_ptr(), tuple(R.stride()), (R.device.type, R.device.index)))
    if hit is not None:
        return hit
    return R.detach().float().cpu().contiguous()


def generate_rotation_metrix(axis_list=None, theta_list=None) -> torch.Tensor:
    """core/utils/projection_prim_ortho.py:23-48 — returns a CUDA tensor like the reference.  The tensor is cached per
    (axes, angles, device) and shared between calls: treat it as read-only (every caller in the reference does)."""
    axes = tuple(["z", "y", "x"] if axis_list is None else axis_list)
    thetas = tuple(float(t) for t in ([0.0, 0.0, 0.0] if theta_list is None else theta_list))
    key = (axes, thetas, torch.cuda.current_device())
    R = _rot_cache.get(key)
    if R is None:
        R_host = rotation_matrix_host(list(thetas), list(axes))
        R = R_host.cuda()
        _rot_cache[key] = R
        _remember_host(R, R_host)
    return R


# ---------------------------------------------------------------------------- sample grids (cached)
_grid_cache: Dict[Tuple, torch.Tensor] = {}


def clear_cache() -> None:
    _grid_cache.clear()
    _rot_cache.clear()
    _rot_host.clear()


def samplegrid_cached(H: int, W: int, R_host: torch.Tensor, device) -> torch.Tensor:
    """[1,2,H,W] grid for a CPU rotation matrix; one kernel launch per (size, R, device, div mode), ever.
    The returned tensor is shared: treat it as read-only."""
    device = torch.device(device)
    key = (H, W, R_host.numpy().tobytes(), device.type, device.index if device.index is not None
           else torch.cuda.current_device(), ops.get_div_mode())
    g = _grid_cache.get(key)
    if g is None:
        g = ops.samplegrid((1, 3, H, W), R_host, device=device)
        _grid_cache[key] = g
    return g


def generate_samplegrid(tensor_size, rotate_metrix: torch.Tensor) -> torch.Tensor:
    """core/utils/projection_prim_ortho.py:432-443 -> [B,2,H,W] fp32 (a fresh, writable tensor per call)."""
    B, _, H, W = (int(s) for s in tensor_size)
    dev = rotate_metrix.device if rotate_metrix.is_cuda else torch.device("cuda")
    g = samplegrid_cached(H, W, _host_matrix(rotate_metrix), dev)
    return g.expand(B, 2, H, W).contiguous() if B > 1 else g.clone()


# ---------------------------------------------------------------------------- samplers
def _mask_of(coords: torch.Tensor, H: int, W: int, cyclic: bool) -> torch.Tensor:
    x, y = coords.split([1, 1], dim=-1)
    if cyclic:
        x = x % W
    gx, gy = 2 * x / (W - 1) - 1, 2 * y / (H - 1) - 1
    return ((gx > -1) & (gy > -1) & (gx < 1) & (gy < 1)).float()


def cycle_bilinear_sampler(img: torch.Tensor, coords: torch.Tensor, mode: str = "bilinear", mask: bool = False):
    """core/utils/utils.py:78-95.  img [B,C,H,W], coords [B,Ho,Wo,2] pixel (x,y)."""
    out = ops.remap_autograd(img.float(), coords.float(), "BHW2", cyclic=True)
    if mask:
        return out, _mask_of(coords, img.shape[-2], img.shape[-1], True)
    return out


def bilinear_sampler(img: torch.Tensor, coords: torch.Tensor, mode: str = "bilinear", mask: bool = False):
    """core/utils/utils.py:61-75."""
    out = ops.remap_autograd(img.float(), coords.float(), "BHW2", cyclic=False)
    if mask:
        return out, _mask_of(coords, img.shape[-2], img.shape[-1], False)
    return out


def coords_grid(batch: int, ht: int, wd: int, device) -> torch.Tensor:
    """core/utils/utils.py:98-101 — [B,2,ht,wd], channel 0 = x, channel 1 = y."""
    ys, xs = torch.meshgrid(torch.arange(ht, device=device), torch.arange(wd, device=device), indexing="ij")
    return torch.stack([xs, ys], dim=0).float()[None].repeat(batch, 1, 1, 1)


# ---------------------------------------------------------------------------- image / flow rotation
def img_rotate(image: torch.Tensor, EulerAngles_zyx=None, sample_grid: Optional[torch.Tensor] = None) -> torch.Tensor:
    """core/utils/projection_prim_ortho.py:507-514."""
    if sample_grid is None:
        assert EulerAngles_zyx is not None
        H, W = image.shape[-2:]
        sample_grid = samplegrid_cached(H, W, rotation_matrix_host(EulerAngles_zyx), image.device)
    return ops.remap_autograd(image.float(), sample_grid, "B2HW", cyclic=True)


def flo_rotate(flow: torch.Tensor, EulerAngles_zyx=None, sample_grid_W2C: Optional[torch.Tensor] = None,
               sample_grid_C2W: Optional[torch.Tensor] = None) -> torch.Tensor:
    """core/utils/projection_prim_ortho.py:531-546.  Unlike the reference's cycle_grid_sample
    (core/utils/my_cycle_sample.py:30-31) the caller's grids are never written to."""
    H, W = flow.shape[-2:]
    if sample_grid_W2C is None or sample_grid_C2W is None:
        assert EulerAngles_zyx is not None
        R = rotation_matrix_host(EulerAngles_zyx)
        if sample_grid_W2C is None:
            sample_grid_W2C = samplegrid_cached(H, W, R.T.contiguous(), flow.device)
        if sample_grid_C2W is None:
            sample_grid_C2W = samplegrid_cached(H, W, R, flow.device)
    ops._no_grad_wrt(flow, "the flow passed to flo_rotate")
    return ops.flo_rotate(flow.detach().float(), sample_grid_W2C, sample_grid_C2W)


def img_A2B(image_A):
    return img_rotate(image_A, EulerAngles_zyx=list(A2B))


def img_B2A(image_B):
    return img_rotate(image_B, EulerAngles_zyx=list(B2A))


def flo_A2B(flow_A):
    return flo_rotate(flow_A, EulerAngles_zyx=list(A2B))


def flo_B2A(flow_B):
    return flo_rotate(flow_B, EulerAngles_zyx=list(B2A))
