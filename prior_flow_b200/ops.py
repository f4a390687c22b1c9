"""Tensor-level entry points: thin PyTorch wrappers over the C ABI (include/priorcorr.h).

PyTorch is plumbing here — device memory (caching allocator), the current stream, autograd
bookkeeping.  Every function validates its inputs, allocates the outputs, and makes exactly one
C-ABI call on `torch.cuda.current_stream()`.  There is no fallback path: a CPU tensor or a missing
library raises.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib

_state = {"div_mode": _lib.DIV_ATEN_CUDA, "volume_mode": "fp32", "launches": 0}


def _count(n: int) -> None:
    """Kernels of ours launched so far (bench.py reports them as `gpu_launches`)."""
    _state["launches"] += n


def reset_launch_count() -> None:
    _state["launches"] = 0


def launch_count() -> int:
    return _state["launches"]


def set_div_mode(mode: str) -> None:
    """"aten_cuda" (default): `tensor / scalar` multiplies by the fp32 reciprocal, as ATen's CUDA kernels
    do — bit-exact coordinates against the reference running on a GPU.  "ieee": true division —
    bit-exact against the reference running on CPU (what tests/golden holds)."""
    _state["div_mode"] = {"aten_cuda": _lib.DIV_ATEN_CUDA, "ieee": _lib.DIV_IEEE}[mode]


def get_div_mode() -> str:
    return "aten_cuda" if _state["div_mode"] == _lib.DIV_ATEN_CUDA else "ieee"


def set_volume_mode(mode: str) -> None:
    if mode not in _lib.VOLUME_MODES:
        raise ValueError(f"volume mode must be one of {sorted(_lib.VOLUME_MODES)}")
    _state["volume_mode"] = mode


def get_volume_mode() -> str:
    return _state["volume_mode"]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(t: torch.Tensor, name: str, ndim: Optional[int] = None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise _lib.PriorCorrError(f"{name} is on {t.device}: prior_flow_b200 runs on CUDA (sm_100a) only, "
                                  "there is no CPU fallback")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (got {t.dtype})")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name} must have {ndim} dims (got shape {tuple(t.shape)})")
    return t


def _grid_arg(g: torch.Tensor, name: str, B: int, H: int, W: int) -> Tuple[torch.Tensor, int]:
    """Sample grids are [B,2,H,W]; batch-invariant grids may be passed expanded (batch stride 0)."""
    _chk(g, name, 4)
    if g.shape[1:] != (2, H, W) or g.shape[0] not in (1, B):
        raise ValueError(f"{name} must be [{B} or 1, 2, {H}, {W}] (got {tuple(g.shape)})")
    if g.stride()[1:] != (H * W, W, 1):
        g = g.contiguous()
    bs = 0 if (g.shape[0] == 1 or g.stride(0) == 0) else g.stride(0)
    return g, bs


# ------------------------------------------------------------------------------------------ (a)
def tcgen05_shape_ok(channels: int, h: int, w: int, num_levels: int = 4) -> bool:
    """Shapes the tensor-core volume kernel tiles exactly (pf_volume_build): other shapes (e.g. h = 55 after
    InputPadder) run the CUDA-core variant."""
    return w % 32 == 0 and h % 8 == 0 and (h * w) % 128 == 0 and channels % 64 == 0 and num_levels <= 4


def volume_pyramid(fmap1: torch.Tensor, fmap2: torch.Tensor, num_levels: int = 4,
                   mode: Optional[str] = None) -> List[torch.Tensor]:
    """Fused corr volume + avg-pool pyramid (core/prior_raft.py:69-75 + core/corr.py:99-111).
    Returns [level_l of shape [B*h*w, 1, h>>l, w>>l]] — the reference's pyramid layout."""
    lib = _lib.load()
    _chk(fmap1, "fmap1", 4), _chk(fmap2, "fmap2", 4)
    if fmap1.shape != fmap2.shape:
        raise ValueError("fmap1 and fmap2 must have the same shape")
    fmap1, fmap2 = fmap1.contiguous(), fmap2.contiguous()
    B, Cn, h, w = fmap1.shape
    mode_id = _lib.VOLUME_MODES[mode or _state["volume_mode"]]
    if mode_id != _lib.VOL_FP32_SIMT and not tcgen05_shape_ok(Cn, h, w, num_levels):
        mode_id = _lib.VOL_FP32_SIMT    # still on the GPU, exact fp32: the tcgen05 tiling needs 8x32 patches of the map
    with torch.cuda.device(fmap1.device):
        levels = [torch.empty((B * h * w, 1, h >> l, w >> l), device=fmap1.device, dtype=torch.float32)
                  for l in range(num_levels)]
        ws_bytes = lib.pf_volume_workspace_bytes(B, Cn, h, w, mode_id)
        ws = torch.empty((max(ws_bytes, 1) + 1024,), device=fmap1.device, dtype=torch.uint8)
        ws_ptr = (ws.data_ptr() + 1023) // 1024 * 1024
        a = _lib.VolumeArgs(B, Cn, h, w, num_levels, mode_id, fmap1.data_ptr(), fmap2.data_ptr(),
                            _lib.level_ptrs(levels), ws_ptr, ws_bytes)
        _lib.check(lib.pf_volume_build(C.byref(a), _stream()), "pf_volume_build")
        _count(num_levels if mode_id == _lib.VOL_FP32_SIMT else 3)   # simt: gemm + pools; tcgen05: absmax, split, gemm
        ws.record_stream(torch.cuda.current_stream())
    return levels


def avg_pool2x2(x: torch.Tensor) -> torch.Tensor:
    """F.avg_pool2d(x, 2, stride=2) for [..., H, W] (core/corr.py:108)."""
    lib = _lib.load()
    _chk(x, "x")
    x = x.contiguous()
    H, W = x.shape[-2:]
    out = torch.empty(x.shape[:-2] + (H // 2, W // 2), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(lib.pf_avg_pool2x2(x.data_ptr(), out.data_ptr(), x.numel() // (H * W), H, W, _stream()),
                   "pf_avg_pool2x2")
        _count(1)
    return out


# ------------------------------------------------------------------------------------------ (b)
def _as_nchw_view(t_cl: torch.Tensor) -> torch.Tensor:
    """[B,h,w,C] storage -> logical [B,C,h,w] tensor in torch.channels_last memory format (no copy)."""
    return t_cl.permute(0, 3, 1, 2)


def lookup(coords: torch.Tensor, pyr_own: Sequence[torch.Tensor], pyr_other: Optional[Sequence[torch.Tensor]] = None,
           grid_w2c: Optional[torch.Tensor] = None, grid_c2w: Optional[torch.Tensor] = None, radius: int = 4,
           cyclic: bool = True, debug: bool = False, channels_last: bool = False, fuse_sum: bool = False):
    """Pyramid lookup.  With `pyr_other` + grids: DCCL.__call__ (core/corr.py:113-144) -> (own, other);
    without: the single-view lookup of CorrBlock.__call__ (core/corr.py:30-51) -> own.
    channels_last=True returns [B,324,h,w] tensors in torch.channels_last memory format (what cuDNN's tensor-core
    convolutions consume without a layout conversion); fuse_sum=True returns the single tensor own + other
    (`corr_A + corr_B_A`, core/prior_raft.py:187-188) with the add folded into the rotate kernel.
    debug=True additionally returns the unnormalised sample coordinates [B*h*w, L, k*k, 2] per branch."""
    lib = _lib.load()
    _chk(coords, "coords", 4)
    coords = coords.contiguous()
    B, two, h, w = coords.shape
    if two != 2:
        raise ValueError("coords must be [B,2,h,w]")
    L = len(pyr_own)
    own = [_chk(t, f"pyr_own[{l}]").contiguous() for l, t in enumerate(pyr_own)]
    h2, w2 = own[0].shape[-2:]
    for l, t in enumerate(own):
        if t.numel() != B * h * w * (h2 >> l) * (w2 >> l):
            raise ValueError(f"pyr_own[{l}] has {t.numel()} elements, expected [{B * h * w},1,{h2 >> l},{w2 >> l}]")
    dual = pyr_other is not None
    if fuse_sum and not dual:
        raise ValueError("fuse_sum needs the dual lookup")
    K2 = (2 * radius + 1) ** 2
    dev = coords.device
    shape = (B, h, w, L * K2) if channels_last else (B, L * K2, h, w)
    with torch.cuda.device(dev):
        out_own = torch.empty(shape, device=dev, dtype=torch.float32)
        a = _lib.LookupArgs()
        a.batch, a.h, a.w, a.h2, a.w2 = B, h, w, h2, w2
        a.radius, a.num_levels, a.cyclic, a.div_mode = radius, L, int(cyclic), _state["div_mode"]
        a.out_channels_last, a.fuse_sum = int(channels_last), int(fuse_sum)
        a.coords = coords.data_ptr()
        a.own = _lib.level_ptrs(own)
        a.out_own = out_own.data_ptr()
        out_other = None
        if dual:
            other = [_chk(t, f"pyr_other[{l}]").contiguous() for l, t in enumerate(pyr_other)]
            if len(other) != L or any(o.shape != t.shape for o, t in zip(other, own)):
                raise ValueError("pyr_other must match pyr_own level by level")
            gw, bs_w = _grid_arg(grid_w2c, "grid_w2c", B, h, w)
            gc, bs_c = _grid_arg(grid_c2w, "grid_c2w", B, h, w)
            if bs_w != bs_c:
                gw, gc = gw.expand(B, 2, h, w).contiguous(), gc.expand(B, 2, h, w).contiguous()
                bs_w = gw.stride(0)
            scratch = torch.empty_like(out_own)
            a.other = _lib.level_ptrs(other)
            a.grid_w2c, a.grid_c2w, a.grid_batch_stride = gw.data_ptr(), gc.data_ptr(), bs_w
            a.scratch = scratch.data_ptr()
            if fuse_sum and not channels_last:
                # the own view is staged channels-last and summed in the rotate kernel's single NCHW write pass
                scratch_own = torch.empty_like(out_own)
                a.scratch_own = scratch_own.data_ptr()
            if not fuse_sum:
                out_other = torch.empty_like(out_own)
                a.out_other = out_other.data_ptr()
        dbg = None
        if debug:
            dbg = [torch.full((B * h * w, L, K2, 2), float("nan"), device=dev) for _ in range(2 if dual else 1)]
            a.dbg_own_xy = dbg[0].data_ptr()
            if dual:
                a.dbg_other_xy = dbg[1].data_ptr()
        _lib.check(lib.pf_lookup_dual(C.byref(a), _stream()), "pf_lookup_dual")
        _count(2 if dual else 1)
    if channels_last:
        out_own = _as_nchw_view(out_own)
        out_other = _as_nchw_view(out_other) if out_other is not None else None
    res = (out_own, out_other) if (dual and not fuse_sum) else out_own
    if debug:
        return res, dbg
    return res


# ------------------------------------------------------------------------------------------ (f1)
def prepare_conv_weight(weight: torch.Tensor) -> torch.Tensor:
    """Conv2d(324, 256, 1) weight -> the K-major fp16 hi/lo planes pf_dccl_conv streams with TMA.  The planes are cached ON the
    tensor object, keyed by its storage address and version counter (an optimizer step or load_state_dict bumps the version);
    a cache keyed by address alone would hand a new layer the planes of a freed one whose memory it reuses."""
    lib = _lib.load()
    _chk(weight, "weight")
    if weight.dim() == 4:
        if weight.shape[2:] != (1, 1):
            raise ValueError("pf_dccl_conv fuses a 1x1 convolution")
    elif weight.dim() != 2:
        raise ValueError("weight must be [256, 324] or [256, 324, 1, 1]")
    key = (weight.data_ptr(), weight._version, str(weight.device))
    hit = getattr(weight, "_pf_conv_planes", None)
    if hit is not None and hit[0] == key:
        return hit[1]
    w2 = weight.detach().reshape(weight.shape[0], weight.shape[1]).contiguous()
    with torch.cuda.device(weight.device):
        nbytes = lib.pf_dccl_conv_weight_bytes()
        buf = torch.empty(nbytes + 1024, device=weight.device, dtype=torch.uint8)
        off = (-buf.data_ptr()) % 1024
        prepared = buf[off:off + nbytes]
        _lib.check(lib.pf_dccl_conv_prepare(w2.data_ptr(), w2.shape[0], w2.shape[1], prepared.data_ptr(), _stream()),
                   "pf_dccl_conv_prepare")
        _count(2)
    try:
        weight._pf_conv_planes = (key, prepared)
    except AttributeError:      # an object that takes no attributes: prepared again on the next call
        pass
    return prepared


def lookup_conv(coords: torch.Tensor, pyr_own: Sequence[torch.Tensor], pyr_other: Sequence[torch.Tensor], grid_w2c: torch.Tensor,
                grid_c2w: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, channels_last: bool = False,
                fp32: bool = True) -> torch.Tensor:
    """relu(conv1x1(own + other)) of a DCCL call in two launches (core/corr.py:113-144 + core/prior_raft.py:187-188 +
    core/update.py:168,184 / :85,92): the gather kernel leaves both views channels-last, pf_dccl_conv rotates, sums and
    convolves on tcgen05 — the [B,324,h,w] tensor never exists.  fp32=True: three-product fp16 split (1e-5 of max|ref|);
    False: single product, TF32-class (2e-3).  Returns [B,256,h,w] (torch.channels_last memory format if asked)."""
    lib = _lib.load()
    _chk(coords, "coords", 4), _chk(bias, "bias", 1)
    coords = coords.contiguous()
    B, two, h, w = coords.shape
    L = len(pyr_own)
    if L != 4 or two != 2:
        raise ValueError("lookup_conv is built for the model's 4-level, radius-4 lookup")
    own = [_chk(t, f"pyr_own[{l}]").contiguous() for l, t in enumerate(pyr_own)]
    other = [_chk(t, f"pyr_other[{l}]").contiguous() for l, t in enumerate(pyr_other)]
    h2, w2 = own[0].shape[-2:]
    prepared = prepare_conv_weight(weight)
    dev = coords.device
    with torch.cuda.device(dev):
        own_cl = torch.empty((B, h, w, 324), device=dev, dtype=torch.float32)
        raw = torch.empty_like(own_cl)
        a = _lib.LookupArgs()
        a.batch, a.h, a.w, a.h2, a.w2 = B, h, w, h2, w2
        a.radius, a.num_levels, a.cyclic, a.div_mode = 4, L, 1, _state["div_mode"]
        a.out_channels_last, a.fuse_sum, a.no_rotate = 1, 0, 1
        a.coords = coords.data_ptr()
        a.own, a.other = _lib.level_ptrs(own), _lib.level_ptrs(other)
        gw, bs_w = _grid_arg(grid_w2c, "grid_w2c", B, h, w)
        gc, bs_c = _grid_arg(grid_c2w, "grid_c2w", B, h, w)
        if bs_w != bs_c:
            gw, gc = gw.expand(B, 2, h, w).contiguous(), gc.expand(B, 2, h, w).contiguous()
            bs_w = gw.stride(0)
        a.grid_w2c, a.grid_c2w, a.grid_batch_stride = gw.data_ptr(), gc.data_ptr(), bs_w
        a.out_own, a.scratch = own_cl.data_ptr(), raw.data_ptr()
        _lib.check(lib.pf_lookup_dual(C.byref(a), _stream()), "pf_lookup_dual")
        out = torch.empty((B, h, w, 256) if channels_last else (B, 256, h, w), device=dev, dtype=torch.float32)
        c = _lib.DcclConvArgs()
        c.batch, c.h, c.w, c.in_channels, c.out_channels = B, h, w, 324, 256
        c.div_mode, c.split, c.out_channels_last, c.after_lookup = _state["div_mode"], int(fp32), int(channels_last), 1
        c.raw, c.own_cl, c.grid_c2w, c.grid_batch_stride = raw.data_ptr(), own_cl.data_ptr(), gc.data_ptr(), bs_w
        c.prepared_weight, c.bias, c.out = prepared.data_ptr(), bias.contiguous().data_ptr(), out.data_ptr()
        _lib.check(lib.pf_dccl_conv(C.byref(c), _stream()), "pf_dccl_conv")
        _count(2)
    return _as_nchw_view(out) if channels_last else out


# ------------------------------------------------------------------------------------------ (d)
def samplegrid(size, R: torch.Tensor, device=None) -> torch.Tensor:
    """generate_samplegrid (core/utils/projection_prim_ortho.py:432-443) -> [B,2,H,W]."""
    lib = _lib.load()
    B, _, H, W = (int(s) for s in size)
    Rh = R.detach().to("cpu", torch.float32).contiguous().reshape(-1)
    if Rh.numel() != 9:
        raise ValueError("R must be 3x3")
    dev = torch.device(device) if device is not None else (R.device if R.is_cuda else torch.device("cuda"))
    if dev.type != "cuda":
        raise _lib.PriorCorrError("samplegrid: CUDA device required (no CPU fallback)")
    Rc = (C.c_float * 9)(*Rh.tolist())
    with torch.cuda.device(dev):
        out = torch.empty((B, 2, H, W), device=dev, dtype=torch.float32)
        _lib.check(lib.pf_samplegrid(out.data_ptr(), B, H, W, Rc, _state["div_mode"], _stream()), "pf_samplegrid")
        _count(1)
    return out


def remap(src: torch.Tensor, coords: torch.Tensor, coords_layout: str, cyclic: bool = True) -> torch.Tensor:
    """Bilinear remap with the reference wrappers' semantics (zeros padding, align_corners=True, optional
    x % W).  coords_layout "BHW2": [B,Ho,Wo,2] (the samplers of core/utils/utils.py:61-95);
    "B2HW": [B,2,Ho,Wo] sample grid (img_rotate, core/utils/projection_prim_ortho.py:507-514)."""
    lib = _lib.load()
    _chk(src, "src", 4), _chk(coords, "coords", 4)
    src = src.contiguous()
    B, Cn, H, W = src.shape
    if coords_layout == "BHW2":
        coords = coords.contiguous()
        Bc, Ho, Wo, two = coords.shape
        cbs, cps, cxs = Ho * Wo * 2, 2, 1
    elif coords_layout == "B2HW":
        Bc, two, Ho, Wo = coords.shape
        coords, cbs = _grid_arg(coords, "coords", B, Ho, Wo)
        cps, cxs = 1, Ho * Wo
    else:
        raise ValueError(coords_layout)
    if two != 2 or Bc not in (1, B):
        raise ValueError(f"bad coords shape {tuple(coords.shape)} for layout {coords_layout}")
    if coords_layout == "BHW2" and Bc == 1 and B > 1:
        cbs = 0
    with torch.cuda.device(src.device):
        out = torch.empty((B, Cn, Ho, Wo), device=src.device, dtype=torch.float32)
        a = _lib.RemapArgs(B, Cn, H, W, Ho, Wo, int(cyclic), _state["div_mode"], src.data_ptr(), coords.data_ptr(),
                           cbs, cps, cxs, out.data_ptr())
        _lib.check(lib.pf_remap(C.byref(a), _stream()), "pf_remap")
        _count(1)
    return out


def flo_rotate(flow: torch.Tensor, grid_w2c: torch.Tensor, grid_c2w: torch.Tensor) -> torch.Tensor:
    """flo_rotate with both sample grids given (core/utils/projection_prim_ortho.py:531-546)."""
    lib = _lib.load()
    _chk(flow, "flow", 4)
    flow = flow.contiguous()
    B, two, H, W = flow.shape
    if two != 2:
        raise ValueError("flow must be [B,2,H,W]")
    gw, bs_w = _grid_arg(grid_w2c, "grid_w2c", B, H, W)
    gc, bs_c = _grid_arg(grid_c2w, "grid_c2w", B, H, W)
    if bs_w != bs_c:
        gw, gc = gw.expand(B, 2, H, W).contiguous(), gc.expand(B, 2, H, W).contiguous()
        bs_w = gw.stride(0)
    with torch.cuda.device(flow.device):
        out = torch.empty_like(flow)
        _lib.check(lib.pf_flo_rotate(flow.data_ptr(), gw.data_ptr(), gc.data_ptr(), bs_w, out.data_ptr(), B, H, W,
                                     _stream()), "pf_flo_rotate")
        _count(1)
    return out


def warp_groupcorr(fmap1: torch.Tensor, fmap2: torch.Tensor, coords: torch.Tensor, groups: int = 4) -> torch.Tensor:
    """groupwise_corr(fmap1, cycle_bilinear_sampler(fmap2, coords)) fused (core/prior_raft.py:173-174,77-83).
    coords [B,2,h,w] -> [B,groups,h,w]."""
    lib = _lib.load()
    _chk(fmap1, "fmap1", 4), _chk(fmap2, "fmap2", 4), _chk(coords, "coords", 4)
    fmap1, fmap2, coords = fmap1.contiguous(), fmap2.contiguous(), coords.contiguous()
    B, Cn, h, w = fmap1.shape
    if fmap2.shape != fmap1.shape or coords.shape != (B, 2, h, w):
        raise ValueError("shape mismatch in warp_groupcorr")
    with torch.cuda.device(fmap1.device):
        out = torch.empty((B, groups, h, w), device=fmap1.device, dtype=torch.float32)
        _lib.check(lib.pf_warp_groupcorr(fmap1.data_ptr(), fmap2.data_ptr(), coords.data_ptr(), out.data_ptr(), B, Cn,
                                         h, w, groups, _state["div_mode"], _stream()), "pf_warp_groupcorr")
        _count(1)
    return out


# ------------------------------------------------------------------------------------------ (c)
def channels_last_pyramid(fmap: torch.Tensor, num_levels: int) -> List[torch.Tensor]:
    """[B,C,h,w] -> [level l as [B, h>>l, w>>l, C]] (avg-pooled, channels-last): the operand layout of
    the on-the-fly lookup (what AlternateCorrBlock feeds alt_cuda_corr, core/corr.py:66-72,82-83)."""
    _chk(fmap, "fmap", 4)
    levels, cur = [], fmap.contiguous()
    for l in range(num_levels):
        if l:
            cur = avg_pool2x2(cur)
        levels.append(cur.permute(0, 2, 3, 1).contiguous())
    return levels


class OnTheFlyPlanes:
    """Tensor-core operands of one view for pf_lookup_onthefly_tc: fp16 hi/lo planes of the channels-last query features and
    of every level of the pooled target pyramid (one shared split scale per tensor family).  The target planes hold every
    image row twice side by side ([B, Hl, 2 Wl, C]) so that windows across the ERP seam are one box."""
    TILE_H, TILE_W = 8, 16
    POOL_SEGMENTS_PER_QUERY = 2          # 16 KiB segments of the local-plane pool per query and view (32 KiB): O(N) scratch

    def __init__(self, f1: torch.Tensor, f2: Sequence[torch.Tensor]):
        lib = _lib.load()
        f1, f2 = f1.detach(), [t.detach() for t in f2]
        dev = f1.device
        with torch.cuda.device(dev):
            self.amax = torch.zeros(2, device=dev, dtype=torch.int32)
            st = _stream()
            a0, a1 = self.amax.data_ptr(), self.amax.data_ptr() + 4
            _lib.check(lib.pf_onthefly_absmax(f1.data_ptr(), f1.numel(), a0, st), "pf_onthefly_absmax")
            _lib.check(lib.pf_onthefly_absmax(f2[0].data_ptr(), f2[0].numel(), a1, st), "pf_onthefly_absmax")
            self.f1_hi, self.f1_lo = torch.empty_like(f1, dtype=torch.float16), torch.empty_like(f1, dtype=torch.float16)
            _lib.check(lib.pf_onthefly_split(f1.data_ptr(), f1.numel(), a0, self.f1_hi.data_ptr(), self.f1_lo.data_ptr(), 0, st), "pf_onthefly_split")
            self.f2_hi, self.f2_lo = [], []
            for t in f2:
                Bt, Hl, Wl, Ct = t.shape
                hi, lo = (torch.empty((Bt, Hl, 2 * Wl, Ct), device=dev, dtype=torch.float16) for _ in range(2))
                _lib.check(lib.pf_onthefly_split(t.data_ptr(), t.numel(), a1, hi.data_ptr(), lo.data_ptr(), Wl * Ct, st), "pf_onthefly_split")
                self.f2_hi.append(hi), self.f2_lo.append(lo)
            _count(3 + len(f2))
        self.shape, self.levels = tuple(f1.shape), len(f2)

    @staticmethod
    def supported(f1: torch.Tensor, radius: int = 4) -> bool:
        B, h, w, Cn = f1.shape
        return radius == 4 and h % OnTheFlyPlanes.TILE_H == 0 and w % OnTheFlyPlanes.TILE_W == 0 and Cn % 128 == 0 and Cn <= 512


_OTF_SCRATCH = {}


def _otf_scratch(dev, views: int, L: int, B: int, h: int, w: int):
    """The local-plane pool and the work buffer of pf_lookup_onthefly_tc: per device and shape, reused by every call (calls are
    stream-ordered; allocate outside CUDA-graph capture — PriOrRAFT.graphed() warms up eagerly first).  Entries are never
    evicted: a captured CUDA graph holds the raw pointers.  `ops._OTF_SCRATCH.clear()` releases them when no graph uses them."""
    key = (dev.index, views, L, B, h, w, OnTheFlyPlanes.POOL_SEGMENTS_PER_QUERY)
    got = _OTF_SCRATCH.get(key)
    if got is None:
        tiles = (h // OnTheFlyPlanes.TILE_H) * (w // OnTheFlyPlanes.TILE_W)
        T = views * L * B * tiles
        segs = max(8, int(views * B * h * w * OnTheFlyPlanes.POOL_SEGMENTS_PER_QUERY))
        pool = torch.empty((segs, 128, 32), device=dev, dtype=torch.float32)
        work = torch.empty(16 + 10 * T + 2 + 2 * (segs // 8 + T), device=dev, dtype=torch.int32)      # PF_OTF_WORK_INTS
        tap_xy = torch.empty((L, B, h * w, 81, 2), device=dev, dtype=torch.float32) if views == 2 else None
        got = _OTF_SCRATCH[key] = (pool, work, T, tap_xy)
    return got


def lookup_onthefly(coords: torch.Tensor, f1_own: torch.Tensor, f2_own: Sequence[torch.Tensor],
                    f1_other: Optional[torch.Tensor] = None, f2_other: Optional[Sequence[torch.Tensor]] = None,
                    grid_w2c: Optional[torch.Tensor] = None, grid_c2w: Optional[torch.Tensor] = None,
                    radius: int = 4, cyclic: bool = True, planes_own: Optional["OnTheFlyPlanes"] = None,
                    planes_other: Optional["OnTheFlyPlanes"] = None, _no_rotate: bool = False):
    """Volume-free lookup.  f1_* are channels-last [B,h,w,C]; f2_* are channels_last_pyramid() lists.  With `planes_*`
    (OnTheFlyPlanes of the same operands) the dot products run on tensor cores (pf_lookup_onthefly_tc).
    `_no_rotate` (tensor-core dual lookup only): returns (own, raw) channels-last [B,h,w,L*81] before img_rotate — the inputs
    of pf_dccl_conv, see lookup_onthefly_conv."""
    lib = _lib.load()
    _chk(coords, "coords", 4)
    coords = coords.contiguous()
    B, _, h, w = coords.shape
    L = len(f2_own)
    Cn = f1_own.shape[-1]
    _chk(f1_own, "f1_own", 4)
    if tuple(f1_own.shape) != (B, h, w, Cn) or not f1_own.is_contiguous():
        raise ValueError("f1_own must be contiguous channels-last [B,h,w,C]")
    for l, t in enumerate(f2_own):
        if tuple(t.shape) != (B, h >> l, w >> l, Cn) or not t.is_contiguous():
            raise ValueError(f"f2_own[{l}] must be contiguous [B,{h >> l},{w >> l},{Cn}]")
    dual = f1_other is not None
    K2 = (2 * radius + 1) ** 2
    dev = coords.device
    tc = planes_own is not None and cyclic and radius == 4 and (not dual or planes_other is not None)
    if _no_rotate and not (tc and dual):
        raise ValueError("_no_rotate needs the tensor-core dual lookup")
    with torch.cuda.device(dev):
        out_own = torch.empty((B, h, w, L * K2) if _no_rotate else (B, L * K2, h, w), device=dev, dtype=torch.float32)
        a = _lib.OnTheFlyArgs()
        a.batch, a.channels, a.h, a.w = B, Cn, h, w
        a.radius, a.num_levels, a.cyclic, a.div_mode = radius, L, int(cyclic), _state["div_mode"]
        a.coords, a.fmap1_own, a.fmap2_own = coords.data_ptr(), f1_own.data_ptr(), _lib.level_ptrs(list(f2_own))
        a.out_own = out_own.data_ptr()
        out_other = None
        if dual:
            gw, bs_w = _grid_arg(grid_w2c, "grid_w2c", B, h, w)
            gc, bs_c = _grid_arg(grid_c2w, "grid_c2w", B, h, w)
            if bs_w != bs_c:
                gw, gc = gw.expand(B, 2, h, w).contiguous(), gc.expand(B, 2, h, w).contiguous()
                bs_w = gw.stride(0)
            scratch = torch.empty_like(out_own)
            out_other = None if _no_rotate else torch.empty_like(out_own)
            a.fmap1_other, a.fmap2_other = f1_other.data_ptr(), _lib.level_ptrs(list(f2_other))
            a.grid_w2c, a.grid_c2w, a.grid_batch_stride = gw.data_ptr(), gc.data_ptr(), bs_w
            a.out_other, a.scratch = (out_other.data_ptr() if out_other is not None else None), scratch.data_ptr()
        if tc:
            t = _lib.OnTheFlyTcArgs()
            t.base = a
            t.no_rotate = int(_no_rotate)
            t.f1_hi_own, t.f1_lo_own = planes_own.f1_hi.data_ptr(), planes_own.f1_lo.data_ptr()
            t.f2_hi_own, t.f2_lo_own = _lib.level_ptrs(planes_own.f2_hi), _lib.level_ptrs(planes_own.f2_lo)
            t.amax_own = planes_own.amax.data_ptr()
            if dual:
                t.f1_hi_other, t.f1_lo_other = planes_other.f1_hi.data_ptr(), planes_other.f1_lo.data_ptr()
                t.f2_hi_other, t.f2_lo_other = _lib.level_ptrs(planes_other.f2_hi), _lib.level_ptrs(planes_other.f2_lo)
                t.amax_other = planes_other.amax.data_ptr()
            pool, work, T, tap_xy = _otf_scratch(dev, 2 if dual else 1, L, B, h, w)
            t.pool, t.pool_segments, t.worklist = pool.data_ptr(), pool.shape[0], work.data_ptr()
            t.tap_xy = tap_xy.data_ptr() if tap_xy is not None else None
            _lib.check(lib.pf_lookup_onthefly_tc(C.byref(t), _stream()), "pf_lookup_onthefly_tc")
            _state["otf_work"] = (work, T, 2 if dual else 1, L, B)      # diagnostics: scripts/probe/otf_tiles.py
            _count((6 if dual else 5) - int(_no_rotate))
            if _no_rotate:
                return out_own, scratch, gc, bs_w
        else:
            _lib.check(lib.pf_lookup_onthefly(C.byref(a), _stream()), "pf_lookup_onthefly")
            _count(2 if dual else 1)
    return (out_own, out_other) if dual else out_own


def lookup_onthefly_conv(coords: torch.Tensor, f1_own: torch.Tensor, f2_own: Sequence[torch.Tensor], f1_other: torch.Tensor,
                         f2_other: Sequence[torch.Tensor], grid_w2c: torch.Tensor, grid_c2w: torch.Tensor, planes_own: "OnTheFlyPlanes",
                         planes_other: "OnTheFlyPlanes", weight: torch.Tensor, bias: torch.Tensor, channels_last: bool = False,
                         fp32: bool = True) -> torch.Tensor:
    """lookup_conv for the volume-free mode: relu(conv1x1(own + img_rotate(other))) with the tensor-core on-the-fly lookup
    leaving both views channels-last and pf_dccl_conv rotating, summing and convolving (SURVEY §8 f1 behind kernel (c))."""
    lib = _lib.load()
    _chk(bias, "bias", 1)
    prepared = prepare_conv_weight(weight)
    own_cl, raw, gc, bs = lookup_onthefly(coords, f1_own, f2_own, f1_other, f2_other, grid_w2c, grid_c2w, 4, True, planes_own, planes_other,
                                          _no_rotate=True)
    B, h, w, K = own_cl.shape
    if K != 324:
        raise ValueError("lookup_onthefly_conv is built for the model's 4-level, radius-4 lookup")
    dev = coords.device
    with torch.cuda.device(dev):
        out = torch.empty((B, h, w, 256) if channels_last else (B, 256, h, w), device=dev, dtype=torch.float32)
        c = _lib.DcclConvArgs()
        c.batch, c.h, c.w, c.in_channels, c.out_channels = B, h, w, 324, 256
        c.div_mode, c.split, c.out_channels_last, c.after_lookup = _state["div_mode"], int(fp32), int(channels_last), 0
        c.raw, c.own_cl, c.grid_c2w, c.grid_batch_stride = raw.data_ptr(), own_cl.data_ptr(), gc.data_ptr(), bs
        c.prepared_weight, c.bias, c.out = prepared.data_ptr(), bias.contiguous().data_ptr(), out.data_ptr()
        _lib.check(lib.pf_dccl_conv(C.byref(c), _stream()), "pf_dccl_conv")
        _count(1)
    return _as_nchw_view(out) if channels_last else out


# ------------------------------------------------------------------------------------------ (c) backward
class OnTheFlyTape:
    """Volume-free backward of the on-the-fly lookups of one forward (both views).

    A lookup's backward needs d(loss)/d(volume) only as an intermediate: dF1 = dV F2^T, dF2 = dV^T F1.  The materialised path
    holds dV for the whole volume; here every lookup's backward just RECORDS (coords, grids, incoming gradients), and when autograd
    reaches the feature pyramids the tape replays all recorded calls chunk by chunk over the query pixels: scatter the chunk's rows
    of dV for both views (pf_lookup_dual_bwd with a query range), fold the levels, and contract them away at once (pf_volume_bwd
    on the chunk).  Memory O(chunk x N) instead of O(N^2); the contraction work is the same single pass as the materialised
    backward's, however many lookup calls there were."""

    CHUNK_BYTES = 1 << 30      # gradient-volume rows held at a time, both views together

    def __init__(self, radius: int, num_levels: int):
        self.radius, self.num_levels = radius, num_levels
        self.views = []            # (fmap1 [B,C,h,w], fmap2 [B,C,h,w])
        self.calls = []
        self.result = None
        self.seeded = False

    def add_view(self, fmap1, fmap2) -> int:
        self.views.append((fmap1, fmap2))
        return len(self.views) - 1

    def record(self, own_id, other_id, coords, grid_w2c, grid_c2w, g_own, g_other):
        self.calls.append((own_id, other_id, coords, grid_w2c, grid_c2w, g_own, g_other))

    def _reset(self):
        self.calls, self.result, self.seeded = [], None, False

    def run(self):
        """-> [(dfmap1, dfmap2) per view]; computed once per backward pass, then handed to each view's pyramid node."""
        if self.result is not None:
            return self.result
        f1_0 = self.views[0][0]
        B, Cn, h, w = f1_0.shape
        N, L, dev = h * w, self.num_levels, f1_0.device
        nv = len(self.views)
        per_row = nv * B * N * 4 * sum(0.25 ** l for l in range(L))
        Q = int(max(128, min(N, (self.CHUNK_BYTES // per_row) // 128 * 128))) if N % 128 == 0 else N
        d1 = [torch.zeros_like(v[0]) for v in self.views]
        d2 = [torch.zeros_like(v[1]) for v in self.views]
        ws = [volume_backward_workspace(v[0]) if volume_backward_shape_ok(Cn, h, w) else None for v in self.views]
        K2 = (2 * self.radius + 1) ** 2
        scratch = torch.empty((B, L * K2, h, w), device=dev, dtype=torch.float32)
        try:
            torch.autograd.Variable._execution_engine.queue_callback(self._reset)     # the tape is per backward pass
        except RuntimeError:
            pass
        for ci, q0 in enumerate(range(0, N, Q)):
            Qc = min(Q, N - q0)
            shapes = [(B * Qc, 1, h >> l, w >> l) for l in range(L)]
            dV = [[torch.zeros(s_, device=dev, dtype=torch.float32) for s_ in shapes] for _ in range(nv)]
            for own_id, other_id, coords, gw, gc, g_own, g_other in self.calls:
                lookup_backward(coords, g_own, g_other, shapes, gw, gc, self.radius, into_own=dV[own_id], into_other=dV[other_id],
                                query_range=(q0, Qc), scratch=scratch)
            for v in range(nv):
                g0 = pyramid_fold_backward(dV[v]).view(B, Qc, N)
                volume_backward_chunk(self.views[v][0], self.views[v][1], g0, q0, d1[v], d2[v], ws[v], first=ci == 0)
            del dV
        self.result = list(zip(d1, d2))
        return self.result


class _FeaturePyramidFn(torch.autograd.Function):
    """fmaps -> the channels-last operands of the on-the-fly lookup, with the tape's volume-free backward attached."""

    @staticmethod
    def forward(ctx, fmap1, fmap2, num_levels, tape, view_id):
        ctx.tape, ctx.view_id, ctx.num_levels = tape, view_id, num_levels
        f1 = fmap1.permute(0, 2, 3, 1).contiguous()
        return (f1, *channels_last_pyramid(fmap2, num_levels))

    @staticmethod
    def backward(ctx, g_f1, *g_f2):
        d1, d2 = ctx.tape.run()[ctx.view_id]
        # gradients that reached the operands directly (the first lookup of a pass seeds f1 with zeros so that this node runs)
        if g_f1 is not None and ctx.tape.views and not _is_seed(g_f1):
            d1 = d1 + g_f1.permute(0, 3, 1, 2)
        for l, g in enumerate(g_f2):
            if g is not None:
                up = g.permute(0, 3, 1, 2)
                if l:
                    up = torch.nn.functional.interpolate(up, scale_factor=2 ** l, mode="nearest") / float(4 ** l)
                d2 = d2 + up
        return d1, d2, None, None, None


_SEEDS = set()


def _is_seed(t: torch.Tensor) -> bool:
    return t.data_ptr() in _SEEDS


class _OnTheFlyLookupFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coords, grid_w2c, grid_c2w, radius, tape, own_id, other_id, planes_own, planes_other, f1_own, f1_other, *f2):
        L = len(f2) // 2
        ctx.save_for_backward(coords, grid_w2c, grid_c2w)
        ctx.meta = (tape, own_id, other_id, tuple(f1_own.shape))
        return lookup_onthefly(coords, f1_own, list(f2[:L]), f1_other, list(f2[L:]), grid_w2c, grid_c2w, radius, cyclic=True,
                               planes_own=planes_own, planes_other=planes_other)

    @staticmethod
    def backward(ctx, g_own, g_other):
        coords, gw, gc = ctx.saved_tensors
        tape, own_id, other_id, f1_shape = ctx.meta
        if g_own is None:
            g_own = torch.zeros_like(g_other)
        if g_other is None:
            g_other = torch.zeros_like(g_own)
        tape.record(own_id, other_id, coords, gw, gc, g_own.contiguous(), g_other.contiguous())
        n_f2 = len(ctx.needs_input_grad) - 11
        seed = None
        if not tape.seeded:        # one real (zero) gradient per pass makes sure autograd visits the pyramid nodes of BOTH views
            tape.seeded = True
            seed = (torch.zeros(f1_shape, device=coords.device), torch.zeros(f1_shape, device=coords.device))
            _SEEDS.clear()
            _SEEDS.update(t.data_ptr() for t in seed)
        return (None, None, None, None, None, None, None, None, None, seed[0] if seed else None, seed[1] if seed else None, *([None] * n_f2))


def lookup_onthefly_autograd(coords, pyr_own, pyr_other, grid_w2c, grid_c2w, radius=4):
    """DCCL lookup straight from the features with gradients to the feature maps (pyr_*: corr.FeaturePyramid built under autograd)."""
    return _OnTheFlyLookupFn.apply(coords.detach(), grid_w2c.detach(), grid_c2w.detach(), radius, pyr_own.tape, pyr_own.view_id,
                                   pyr_other.view_id, pyr_own.planes, pyr_other.planes, pyr_own.f1, pyr_other.f1, *pyr_own.f2, *pyr_other.f2)


# ------------------------------------------------------------------------------------------ (f2) / (f4)
def _convex_upsample_args(flow, mask):
    _chk(flow, "flow", 4), _chk(mask, "mask", 4)
    B, two, h, w = flow.shape
    if two != 2 or tuple(mask.shape) != (B, 576, h, w):
        raise ValueError("convex_upsample: flow [B,2,h,w], mask [B,576,h,w]")
    flow = flow.contiguous()
    cl = mask.is_contiguous(memory_format=torch.channels_last) and not mask.is_contiguous()
    if not cl:
        mask = mask.contiguous()
    return flow, mask, cl, B, h, w


def _convex_upsample_fwd(flow, mask):
    lib = _lib.load()
    flow, mask, cl, B, h, w = _convex_upsample_args(flow, mask)
    with torch.cuda.device(flow.device):
        out = torch.empty((B, 2, 8 * h, 8 * w), device=flow.device, dtype=torch.float32)
        _lib.check(lib.pf_convex_upsample(flow.data_ptr(), mask.data_ptr(), out.data_ptr(), B, h, w, int(cl), _stream()),
                   "pf_convex_upsample")
        _count(1)
    return out


class _ConvexUpsampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flow, mask):
        ctx.save_for_backward(flow, mask)
        return _convex_upsample_fwd(flow, mask)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        flow, mask = ctx.saved_tensors
        flow, mask, cl, B, h, w = _convex_upsample_args(flow, mask)
        g = g.contiguous().float()
        with torch.cuda.device(flow.device):
            dflow = torch.empty_like(flow)
            dmask = torch.empty_like(mask)          # same memory format as the mask
            _lib.check(lib.pf_convex_upsample_bwd(flow.data_ptr(), mask.data_ptr(), g.data_ptr(), dflow.data_ptr(), dmask.data_ptr(),
                                                  B, h, w, int(cl), _stream()), "pf_convex_upsample_bwd")
            _count(1)
        return dflow, dmask


def convex_upsample(flow: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """PriOr_RAFT.upsample_flow (core/prior_raft.py:58-67) in one launch: flow [B,2,h,w], mask [B,576,h,w] (NCHW or
    torch.channels_last) -> [B,2,8h,8w].  Under autograd the adjoint is one launch too (pf_convex_upsample_bwd: softmax
    recomputed, nothing but the two inputs is kept for backward — the eager chain keeps five 576-channel tensors per call)."""
    if torch.is_grad_enabled() and (flow.requires_grad or mask.requires_grad):
        return _ConvexUpsampleFn.apply(flow, mask)
    return _convex_upsample_fwd(flow, mask)


class _UniformLossTermFn(torch.autograd.Function):
    """term_weight * sum(ok * lat[y] * |pred - gt|_1) — one prediction's share of uniform_loss (train_flow.py:66-69)."""

    @staticmethod
    def forward(ctx, pred, gt, ok, lat, term_weight):
        lib = _lib.load()
        pred, gt, ok, lat = pred.contiguous(), gt.contiguous(), ok.contiguous(), lat.contiguous()
        B, _, H, W = pred.shape
        with torch.cuda.device(pred.device):
            acc = torch.zeros(1, device=pred.device, dtype=torch.float32)
            _lib.check(lib.pf_uniform_loss_fwd(pred.data_ptr(), gt.data_ptr(), ok.data_ptr(), lat.data_ptr(), acc.data_ptr(),
                                               float(term_weight), B, H, W, _stream()), "pf_uniform_loss_fwd")
        ctx.save_for_backward(pred, gt, ok, lat)
        ctx.term_weight = float(term_weight)
        return acc.view(())

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        pred, gt, ok, lat = ctx.saved_tensors
        B, _, H, W = pred.shape
        g = g.contiguous().float().view(1)
        with torch.cuda.device(pred.device):
            dpred = torch.empty_like(pred)
            _lib.check(lib.pf_uniform_loss_bwd(pred.data_ptr(), gt.data_ptr(), ok.data_ptr(), lat.data_ptr(), g.data_ptr(),
                                               ctx.term_weight, dpred.data_ptr(), B, H, W, _stream()), "pf_uniform_loss_bwd")
        return dpred, None, None, None, None


def uniform_loss_term(pred: torch.Tensor, gt: torch.Tensor, ok: torch.Tensor, lat: torch.Tensor, term_weight: float) -> torch.Tensor:
    """pred, gt [B,2,H,W]; ok [B,H,W] 0/1 floats; lat [H] normalised cos-latitude weights -> scalar tensor (differentiable in pred)."""
    _chk(pred, "pred", 4), _chk(gt, "gt", 4), _chk(ok, "ok", 3), _chk(lat, "lat", 1)
    if pred.shape != gt.shape or pred.shape[1] != 2 or tuple(ok.shape) != (pred.shape[0],) + tuple(pred.shape[2:]) or lat.numel() != pred.shape[2]:
        raise ValueError("uniform_loss_term: shape mismatch")
    return _UniformLossTermFn.apply(pred, gt, ok, lat, term_weight)


def great_circle_distance(pred: torch.Tensor, gt: torch.Tensor, radius: float = 1.0) -> torch.Tensor:
    """calculate_great_circle_distance(pred, gt, 'Haversine', R) (core/utils/spherical.py:20-53) -> [B,H,W]."""
    lib = _lib.load()
    _chk(pred, "pred", 4), _chk(gt, "gt", 4)
    if pred.shape != gt.shape or pred.shape[1] != 2:
        raise ValueError("great_circle_distance: [B,2,H,W] flows of equal shape")
    pred, gt = pred.contiguous(), gt.contiguous()
    B, _, H, W = pred.shape
    with torch.cuda.device(pred.device):
        out = torch.empty((B, H, W), device=pred.device, dtype=torch.float32)
        _lib.check(lib.pf_great_circle(pred.data_ptr(), gt.data_ptr(), out.data_ptr(), B, H, W, float(radius), _stream()), "pf_great_circle")
        _count(1)
    return out


# ------------------------------------------------------------------------------------------ measurement aids
def probe_gather(vol: torch.Tensor, pos_xy: torch.Tensor) -> None:
    """Issues the loads of an own-view lookup of `vol` [planes,H,W] at integer footprint corners pos_xy [planes,2] (int32) and
    nothing else — the HBM floor of the lookup's access pattern (bench.py times it beside the kernel)."""
    lib = _lib.load()
    _chk(vol, "vol", 3)
    if pos_xy.dtype != torch.int32 or not pos_xy.is_cuda or tuple(pos_xy.shape) != (vol.shape[0], 2):
        raise ValueError("pos_xy must be a CUDA int32 tensor [planes, 2]")
    vol, pos_xy = vol.contiguous(), pos_xy.contiguous()
    with torch.cuda.device(vol.device):
        sink = torch.empty(1, device=vol.device)
        _lib.check(lib.pf_probe_gather(vol.data_ptr(), vol.shape[0], vol.shape[1], vol.shape[2], pos_xy.data_ptr(),
                                       sink.data_ptr(), _stream()), "pf_probe_gather")


def probe_stream_read(buf: torch.Tensor) -> None:
    """Reads a contiguous fp32 CUDA tensor once (read-only streaming ceiling)."""
    lib = _lib.load()
    _chk(buf, "buf")
    with torch.cuda.device(buf.device):
        sink = torch.empty(1, device=buf.device)
        _lib.check(lib.pf_probe_stream_read(buf.data_ptr(), buf.numel(), sink.data_ptr(), _stream()), "pf_probe_stream_read")


# ------------------------------------------------------------------------------------------ (e)
def _grad_layout(g: torch.Tensor, channels_last: bool) -> torch.Tensor:
    """Brings an incoming [B,C,h,w] gradient into the memory layout the backward kernels read."""
    return g.permute(0, 2, 3, 1).contiguous() if channels_last else g.contiguous()


def lookup_backward(coords, grad_own, grad_other, level_shapes, grid_w2c=None, grid_c2w=None, radius=4, cyclic=True,
                    into_own=None, into_other=None, channels_last=False, query_range=None, scratch=None):
    """Adjoint of `lookup` w.r.t. the pyramids.  Returns (d_own levels, d_other levels); with `into_*` given the
    gradients are accumulated (+=) into those tensors instead of fresh zero tensors.  query_range=(begin, count) scatters only
    those queries of every batch item into gradient pyramids that hold `count` planes per batch item (`level_shapes` must say so):
    the chunked, volume-free backward of the on-the-fly lookup."""
    lib = _lib.load()
    coords = _chk(coords, "coords", 4).contiguous()
    B, _, h, w = coords.shape
    L = len(level_shapes)
    h2, w2 = level_shapes[0][-2:]
    dual = grad_other is not None
    K2 = (2 * radius + 1) ** 2
    dev = coords.device
    with torch.cuda.device(dev):
        d_own = into_own if into_own is not None else [torch.zeros(s, device=dev, dtype=torch.float32) for s in level_shapes]
        d_other = None
        ba = _lib.LookupBwdArgs()
        a = ba.fwd
        a.batch, a.h, a.w, a.h2, a.w2 = B, h, w, h2, w2
        a.radius, a.num_levels, a.cyclic, a.div_mode = radius, L, int(cyclic), _state["div_mode"]
        a.out_channels_last = int(channels_last)
        a.coords = coords.data_ptr()
        grad_own = _grad_layout(_chk(grad_own, "grad_own", 4), channels_last)
        ba.grad_own = grad_own.data_ptr()
        ba.dgrad_own = _lib.level_ptrs(d_own)
        if dual:
            d_other = into_other if into_other is not None else [torch.zeros(s, device=dev, dtype=torch.float32)
                                                                 for s in level_shapes]
            grad_other = _grad_layout(_chk(grad_other, "grad_other", 4), channels_last)
            gw, bs_w = _grid_arg(grid_w2c, "grid_w2c", B, h, w)
            gc, bs_c = _grid_arg(grid_c2w, "grid_c2w", B, h, w)
            if bs_w != bs_c:
                gw, gc = gw.expand(B, 2, h, w).contiguous(), gc.expand(B, 2, h, w).contiguous()
                bs_w = gw.stride(0)
            if scratch is None:
                scratch = torch.empty((B, L * K2, h, w), device=dev, dtype=torch.float32)
            a.grid_w2c, a.grid_c2w, a.grid_batch_stride, a.scratch = gw.data_ptr(), gc.data_ptr(), bs_w, scratch.data_ptr()
            ba.grad_other = grad_other.data_ptr()
            ba.dgrad_other = _lib.level_ptrs(d_other)
        if query_range is not None:
            ba.query_begin, ba.query_count = int(query_range[0]), int(query_range[1])
        _lib.check(lib.pf_lookup_dual_bwd(C.byref(ba), _stream()), "pf_lookup_dual_bwd")
    return d_own, d_other


def remap_backward(dout: torch.Tensor, coords: torch.Tensor, coords_layout: str, src_shape, cyclic: bool = True):
    """Adjoint of `remap` w.r.t. src."""
    lib = _lib.load()
    dout = _chk(dout, "dout", 4).contiguous()
    B, Cn, H, W = src_shape
    Ho, Wo = dout.shape[-2:]
    if coords_layout == "BHW2":
        coords = _chk(coords, "coords", 4).contiguous()
        cbs, cps, cxs = (0 if coords.shape[0] == 1 and B > 1 else Ho * Wo * 2), 2, 1
    else:
        coords, cbs = _grid_arg(coords, "coords", B, Ho, Wo)
        cps, cxs = 1, Ho * Wo
    with torch.cuda.device(dout.device):
        dsrc = torch.zeros((B, Cn, H, W), device=dout.device, dtype=torch.float32)
        a = _lib.RemapArgs(B, Cn, H, W, Ho, Wo, int(cyclic), _state["div_mode"], None, coords.data_ptr(), cbs, cps, cxs, None)
        _lib.check(lib.pf_remap_bwd(C.byref(a), dout.data_ptr(), dsrc.data_ptr(), _stream()), "pf_remap_bwd")
    return dsrc


def pyramid_fold_backward(grads: Sequence[torch.Tensor]) -> torch.Tensor:
    """Folds level gradients [planes,1,H>>l,W>>l] into level 0 (in place on grads[0]) and returns it."""
    lib = _lib.load()
    g0 = grads[0]
    H, W = g0.shape[-2:]
    planes = g0.numel() // (H * W)
    ptrs = (C.c_void_p * len(grads))(*[g.data_ptr() for g in grads])
    with torch.cuda.device(g0.device):
        _lib.check(lib.pf_pyramid_fold_bwd(ptrs, len(grads), planes, H, W, _stream()), "pf_pyramid_fold_bwd")
    return g0


def warp_groupcorr_backward(fmap1, fmap2, coords, dout, groups: int = 4):
    lib = _lib.load()
    fmap1, fmap2, coords, dout = (t.contiguous() for t in (fmap1, fmap2, coords, dout))
    B, Cn, h, w = fmap1.shape
    with torch.cuda.device(fmap1.device):
        df1 = torch.empty_like(fmap1)
        df2 = torch.zeros_like(fmap2)
        _lib.check(lib.pf_warp_groupcorr_bwd(fmap1.data_ptr(), fmap2.data_ptr(), coords.data_ptr(), dout.data_ptr(),
                                             df1.data_ptr(), df2.data_ptr(), B, Cn, h, w, groups, _state["div_mode"],
                                             _stream()), "pf_warp_groupcorr_bwd")
    return df1, df2


# ------------------------------------------------------------------------------------------ autograd
class _VolumePyramidFn(torch.autograd.Function):
    """volume_pyramid with gradients to the feature maps: fold the level gradients into level 0
    (pf_pyramid_fold_bwd), then dF1 = dV F2^T / sqrt(C) and dF2 = dV^T F1 / sqrt(C) (volume_backward)."""

    @staticmethod
    def forward(ctx, fmap1, fmap2, num_levels, mode, sink):
        ctx.save_for_backward(fmap1, fmap2)
        ctx.num_levels = num_levels
        ctx.sink = sink
        return tuple(volume_pyramid(fmap1, fmap2, num_levels, mode))

    @staticmethod
    def backward(ctx, *grads):
        fmap1, fmap2 = ctx.saved_tensors
        B, Cn, h, w = fmap1.shape
        N = h * w
        sink = ctx.sink
        if sink is not None and sink.bufs is not None:
            # the sink's in-place scatters are only valid if autograd handed us the very storage they went into: a
            # pyramid level with a consumer besides the DCCL lookups makes the engine sum out of place and lose them
            for g, buf in zip(grads, sink.bufs):
                if g is not None and g.data_ptr() != buf.data_ptr():
                    sink.reset()
                    raise RuntimeError("prior_flow_b200.GradSink: a pyramid level was consumed by something other than DCCL "
                                       "lookups; build the DCCL with accumulate_grads=False for this graph")
            sink.reset()     # the pass is over for this pyramid: a second backward (retain_graph) starts from zero
        gs = []
        for l, g in enumerate(grads):
            if g is None:
                gs.append(torch.zeros((B * N, 1, h >> l, w >> l), device=fmap1.device))
            else:
                gs.append(g.contiguous().clone() if (l == 0 and sink is None) else g.contiguous())  # level 0 is folded into in place
        g0 = pyramid_fold_backward(gs).view(B, N, N)
        d1, d2 = volume_backward(fmap1, fmap2, g0, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return d1, d2, None, None, None


def volume_backward_shape_ok(channels: int, h: int, w: int) -> bool:
    return channels == 256 and (h * w) % 128 == 0


def volume_backward_chunk(fmap1, fmap2, g0_chunk, q_begin, d1, d2, workspace, first: bool):
    """One chunk of query rows of the volume adjoints (pf_volume_bwd with query_begin / query_count): g0_chunk [B, Qc, N] holds
    the level-0 volume gradient of queries [q_begin, q_begin + Qc); d1 [B,C,h,w] receives those queries' columns, d2 [B,C,h,w] is
    written by the first chunk and accumulated by the others.  `workspace`: the 1 KiB-aligned uint8 tensor of
    pf_volume_bwd_workspace_bytes(), shared by all chunks (it keeps the bf16 planes of the feature maps)."""
    lib = _lib.load()
    B, Cn, h, w = fmap1.shape
    N = h * w
    Qc = g0_chunk.shape[1]
    if not (volume_backward_shape_ok(Cn, h, w) and Qc % 128 == 0 and q_begin % 64 == 0):
        scale = 1.0 / (Cn ** 0.5)
        f1, f2 = fmap1.reshape(B, Cn, N), fmap2.reshape(B, Cn, N)
        d1.view(B, Cn, N)[:, :, q_begin:q_begin + Qc] = torch.matmul(f2, g0_chunk.transpose(1, 2)) * scale
        upd = torch.matmul(f1[:, :, q_begin:q_begin + Qc], g0_chunk) * scale
        if first:
            d2.view(B, Cn, N).copy_(upd)
        else:
            d2.view(B, Cn, N).add_(upd)
        return
    g0_chunk = g0_chunk.contiguous()
    with torch.cuda.device(fmap1.device):
        a = _lib.VolumeBwdArgs(B, Cn, h, w, fmap1.data_ptr(), fmap2.data_ptr(), g0_chunk.data_ptr(), d1.data_ptr(), d2.data_ptr(),
                               workspace.data_ptr(), workspace.numel(), int(q_begin), int(Qc), int(not first), int(not first))
        _lib.check(lib.pf_volume_bwd(C.byref(a), _stream()), "pf_volume_bwd")


def volume_backward_workspace(fmap: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    B, Cn, h, w = fmap.shape
    nbytes = lib.pf_volume_bwd_workspace_bytes(B, Cn, h, w)
    buf = torch.empty(nbytes + 1024, device=fmap.device, dtype=torch.uint8)
    off = (-buf.data_ptr()) % 1024
    return buf[off:off + nbytes]


def volume_backward(fmap1, fmap2, g0, need1=True, need2=True, use_library: bool = False):
    """Adjoints of the volume GEMM (autograd of core/prior_raft.py:73-75): dF1[c,n] = sum_m dV[n,m] F2[c,m] / sqrt(C),
    dF2[c,m] = sum_n dV[n,m] F1[c,n] / sqrt(C).  g0: [B, N, N].  One tcgen05 launch for both (pf_volume_bwd: bf16 hi/lo
    split, three products); shapes it does not tile (C != 256 or h*w % 128 != 0) and use_library=True run two library GEMMs."""
    B, Cn, h, w = fmap1.shape
    N = h * w
    if not (need1 or need2):
        return None, None
    if use_library or not volume_backward_shape_ok(Cn, h, w):
        scale = 1.0 / (Cn ** 0.5)
        f1 = fmap1.reshape(B, Cn, N)
        f2 = fmap2.reshape(B, Cn, N)
        d1 = torch.matmul(f2, g0.transpose(1, 2)).mul_(scale).view_as(fmap1) if need1 else None
        d2 = torch.matmul(f1, g0).mul_(scale).view_as(fmap2) if need2 else None
        return d1, d2
    lib = _lib.load()
    fmap1, fmap2, g0 = fmap1.contiguous(), fmap2.contiguous(), g0.contiguous()
    with torch.cuda.device(fmap1.device):
        d1 = torch.empty_like(fmap1) if need1 else None
        d2 = torch.empty_like(fmap2) if need2 else None
        ws_bytes = lib.pf_volume_bwd_workspace_bytes(B, Cn, h, w)
        ws = torch.empty((ws_bytes + 1024,), device=fmap1.device, dtype=torch.uint8)
        ws_ptr = (ws.data_ptr() + 1023) // 1024 * 1024
        a = _lib.VolumeBwdArgs(B, Cn, h, w, fmap1.data_ptr(), fmap2.data_ptr(), g0.data_ptr(),
                               d1.data_ptr() if need1 else None, d2.data_ptr() if need2 else None, ws_ptr, ws_bytes)
        _lib.check(lib.pf_volume_bwd(C.byref(a), _stream()), "pf_volume_bwd")
        ws.record_stream(torch.cuda.current_stream())
    return d1, d2


def volume_pyramid_autograd(fmap1, fmap2, num_levels=4, mode=None, sink=None):
    if torch.is_grad_enabled() and (fmap1.requires_grad or fmap2.requires_grad):
        return list(_VolumePyramidFn.apply(fmap1, fmap2, num_levels, mode, sink))
    return volume_pyramid(fmap1, fmap2, num_levels, mode)


class _DualLookupFn(torch.autograd.Function):
    """DCCL lookup with gradients to both pyramids (coords and grids carry none, prior_raft.py:171,176)."""

    @staticmethod
    def forward(ctx, coords, grid_w2c, grid_c2w, radius, num_levels, channels_last, fuse_sum, sink_own, sink_other, *levels):
        own, other = levels[:num_levels], levels[num_levels:]
        out = lookup(coords, own, other, grid_w2c, grid_c2w, radius, channels_last=channels_last, fuse_sum=fuse_sum)
        ctx.save_for_backward(coords, grid_w2c, grid_c2w)
        ctx.meta = (radius, num_levels, [tuple(t.shape) for t in own], channels_last, fuse_sum)
        ctx.sinks = (sink_own, sink_other)
        return out

    @staticmethod
    def backward(ctx, *grads):
        coords, grid_w2c, grid_c2w = ctx.saved_tensors
        radius, L, shapes, channels_last, fuse_sum = ctx.meta
        sink_own, sink_other = ctx.sinks
        g_own = grads[0]
        g_other = g_own if fuse_sum else grads[1]      # d(own + other) flows to both branches unchanged
        if g_own is None:
            g_own = torch.zeros((coords.shape[0], L * (2 * radius + 1) ** 2) + tuple(coords.shape[2:]), device=coords.device)
        if g_other is None:
            g_other = torch.zeros_like(g_own)
        # With a GradSink the FIRST backward call of a pass hands autograd the freshly zeroed gradient pyramid and
        # every later call scatters straight into that same storage and returns None — instead of 24 zero-filled
        # 357 MB pyramids per view that autograd then adds up pairwise (profiles: 14 ms of a 115 ms training step).
        into_own = sink_own.bufs if sink_own is not None else None
        into_other = sink_other.bufs if sink_other is not None else None
        if sink_own is not None and sink_own is sink_other and into_own is None:
            into_own = [torch.zeros(s_, device=coords.device, dtype=torch.float32) for s_ in shapes]
            sink_own.bufs, into_other, first_own = into_own, into_own, True
        else:
            first_own = sink_own is not None and into_own is None
        first_other = sink_other is not None and sink_other is not sink_own and into_other is None
        d_own, d_other = lookup_backward(coords, g_own, g_other, shapes, grid_w2c, grid_c2w, radius,
                                         into_own=into_own, into_other=into_other, channels_last=channels_last)
        if first_own:
            sink_own.bufs = d_own
            sink_own.arm()
        if first_other:
            sink_other.bufs = d_other
            sink_other.arm()
        r_own = d_own if (sink_own is None or first_own) else [None] * L
        r_other = d_other if (sink_other is None or first_other) else [None] * L
        return (None, None, None, None, None, None, None, None, None, *r_own, *r_other)


class GradSink:
    """Opt-in, per pyramid: the shared in-place accumulator of d(loss)/d(pyramid levels) used by
    `_DualLookupFn.backward` (see there).  Valid when the pyramid's levels are consumed by DCCL lookups only — what
    `PriOrRAFT.forward` builds; `_VolumePyramidFn.backward` verifies that and raises otherwise.  The buffers live for one
    backward pass: they are dropped when the pyramid's own backward has consumed them and, as a safety net, by an
    engine callback at the end of the pass, so a second backward over a retained graph starts from zero again."""

    def __init__(self):
        self.bufs = None
        self._armed = False

    def reset(self):
        self.bufs = None
        self._armed = False

    def arm(self):
        """Called by the first lookup backward of a pass: queue the end-of-pass reset (only legal inside backward)."""
        if not self._armed:
            self._armed = True
            try:
                torch.autograd.Variable._execution_engine.queue_callback(self.reset)
            except RuntimeError:
                pass


class _SingleLookupFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coords, radius, cyclic, *levels):
        out = lookup(coords, levels, None, None, None, radius, cyclic)
        ctx.save_for_backward(coords)
        ctx.meta = (radius, cyclic, [tuple(t.shape) for t in levels])
        return out

    @staticmethod
    def backward(ctx, g):
        (coords,) = ctx.saved_tensors
        radius, cyclic, shapes = ctx.meta
        d_own, _ = lookup_backward(coords, g, None, shapes, None, None, radius, cyclic)
        return (None, None, None, *d_own)


def lookup_autograd(coords, pyr_own, pyr_other=None, grid_w2c=None, grid_c2w=None, radius=4, cyclic=True,
                    channels_last=False, fuse_sum=False, sink_own=None, sink_other=None):
    needs = torch.is_grad_enabled() and any(t.requires_grad for t in list(pyr_own) + list(pyr_other or []))
    if not needs:
        return lookup(coords, pyr_own, pyr_other, grid_w2c, grid_c2w, radius, cyclic, channels_last=channels_last,
                      fuse_sum=fuse_sum)
    if pyr_other is None:
        return _SingleLookupFn.apply(coords.detach(), radius, cyclic, *pyr_own)
    return _DualLookupFn.apply(coords.detach(), grid_w2c.detach(), grid_c2w.detach(), radius, len(pyr_own), channels_last,
                               fuse_sum, sink_own, sink_other, *pyr_own, *pyr_other)


class _RemapFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, coords, layout, cyclic):
        ctx.save_for_backward(coords)
        ctx.meta = (layout, cyclic, tuple(src.shape))
        return remap(src, coords, layout, cyclic)

    @staticmethod
    def backward(ctx, g):
        (coords,) = ctx.saved_tensors
        layout, cyclic, shape = ctx.meta
        return remap_backward(g, coords, layout, shape, cyclic), None, None, None


def _no_grad_wrt(t: torch.Tensor, what: str) -> None:
    """The kernels implement d/d(source) only.  The reference's forward never asks for more (coordinates are detached at
    the top of every iteration, core/prior_raft.py:171,176; the sample grids are constants) — any other caller must hear
    about it instead of silently training with a missing gradient."""
    if torch.is_grad_enabled() and t.requires_grad:
        raise NotImplementedError(f"prior_flow_b200: no gradient with respect to {what} (detach it, as "
                                  "PriOr_RAFT.forward does, or use the eager sampler)")


def remap_autograd(src, coords, layout, cyclic=True):
    _no_grad_wrt(coords, "sample coordinates")
    if torch.is_grad_enabled() and src.requires_grad:
        return _RemapFn.apply(src, coords.detach(), layout, cyclic)
    return remap(src, coords, layout, cyclic)


class _WarpGroupCorrFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fmap1, fmap2, coords, groups):
        ctx.save_for_backward(fmap1, fmap2, coords)
        ctx.groups = groups
        return warp_groupcorr(fmap1, fmap2, coords, groups)

    @staticmethod
    def backward(ctx, g):
        fmap1, fmap2, coords = ctx.saved_tensors
        d1, d2 = warp_groupcorr_backward(fmap1, fmap2, coords, g, ctx.groups)
        return d1, d2, None, None


def warp_groupcorr_autograd(fmap1, fmap2, coords, groups=4):
    if torch.is_grad_enabled() and (fmap1.requires_grad or fmap2.requires_grad):
        return _WarpGroupCorrFn.apply(fmap1, fmap2, coords.detach(), groups)
    return warp_groupcorr(fmap1, fmap2, coords, groups)
