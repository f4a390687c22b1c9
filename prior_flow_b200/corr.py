"""Host-side mirror of PriOr-RAFT/core/corr.py: `DCCL`, `CorrBlock`, `AlternateCorrBlock` with the
reference's constructor and call signatures, backed by the sm_100a kernels.

`PriOr_RAFT.forward` never inspects what `self.corr(...)` or `build_pyramid(...)` return
(core/prior_raft.py:151-159), so both are opaque handles here:

  corr(fmap1, fmap2)          -> CostVolume   (lazy: just the two feature maps)
  DCCL.build_pyramid(volume)  -> Pyramid      (fused tcgen05 GEMM + pooling)  or
                                 FeaturePyramid (on-the-fly mode: pooled channels-last fmaps, O(N*C) memory)
  DCCL.__call__(coords, pyr_own, pyr_other, grid_W2C_8x, grid_C2W_8x) -> (out_own, out_other)

A materialised [B,h,w,h,w] tensor is still accepted by build_pyramid (it is pooled with the
avg-pool kernel), and `CostVolume.materialize()` gives that tensor to callers that want it.
"""
from __future__ import annotations

import os
import time
from typing import List, Optional, Sequence

import torch

from . import ops

# mode="auto": materialise both views' pyramids while they fit comfortably in device memory — on a 180 GB B200 that
# includes 1024x2048 (2 x 5.7 GB), where the materialised lookups are 2.9x faster end to end than the on-the-fly
# kernel (profiles/r02f_bench_hires_*.json) — and fall back to the volume-free on-the-fly lookup beyond that.
AUTO_MATERIALIZE_FRACTION = 0.4     # of the currently free device memory, for the two pyramids + split workspace


DRIVER_FREE_TTL_S = 1.0             # how long a cudaMemGetInfo answer is reused (see available_device_memory)
_driver_free = {}                   # str(device) -> (monotonic time, driver-free bytes, torch-reserved bytes at that moment)


def available_device_memory(device) -> int:
    """Bytes a new allocation can use: what the driver reports free PLUS what PyTorch's caching allocator holds but has not
    handed out (reserved - allocated).  `mem_get_info` alone shrinks as a long-running process caches freed blocks, which would
    silently push `mode="auto"` onto the slower volume-free path.
    The driver query (cudaMemGetInfo) is reused for DRIVER_FREE_TTL_S: the reference builds a new DCCL per forward, so `auto`
    asks per forward, and inside a process with CUDA graphs and gigabytes of cached blocks the call was measured at 2-17 ms
    (r03z: it was the whole 5.6 ms of the drop-in's build_pyramid stage).  Between queries the estimate follows the caching
    allocator's own counters, which are exact and cost microseconds."""
    key = str(device)
    try:
        reserved, allocated = torch.cuda.memory_reserved(device), torch.cuda.memory_allocated(device)
    except Exception:  # noqa: BLE001 - a mocked / uninitialised device
        reserved = allocated = 0
    now = time.monotonic()
    hit = _driver_free.get(key)
    if hit is None or now - hit[0] > DRIVER_FREE_TTL_S:
        hit = _driver_free[key] = (now, int(torch.cuda.mem_get_info(device)[0]), reserved)
    driver_free = hit[1] - max(0, reserved - hit[2])          # what the allocator took from the driver since the query
    return int(max(0, driver_free) + reserved - allocated)


class CostVolume:
    """Lazy all-pairs volume: what `PriOr_RAFT.corr` returns (core/prior_raft.py:69-75)."""

    def __init__(self, fmap1: torch.Tensor, fmap2: torch.Tensor, mode: Optional[str] = None):
        if fmap1.shape != fmap2.shape or fmap1.dim() != 4:
            raise ValueError("fmap1/fmap2 must be [B,C,h,w] with equal shapes")
        self.fmap1, self.fmap2, self.mode = fmap1.float(), fmap2.float(), mode

    @property
    def shape(self):
        B, _, h, w = self.fmap1.shape
        return torch.Size((B, h, w, h, w))

    def materialize(self) -> torch.Tensor:
        """The reference's tensor: [B, h, w, h, w] fp32."""
        B, _, h, w = self.fmap1.shape
        return ops.volume_pyramid_autograd(self.fmap1, self.fmap2, 1, self.mode)[0].view(B, h, w, h, w)


class Pyramid(list):
    """Materialised pyramid: a list of [B*h*w, 1, h>>l, w>>l] tensors (the reference's layout, core/corr.py:99-111).
    `grad_sink` (training, opt-in): the shared in-place accumulator of the lookups' gradients, see ops.GradSink."""
    grad_sink = None


class FeaturePyramid:
    """On-the-fly operands: channels-last query features and the pooled channels-last target pyramid.  Built under autograd
    (`tape` given) the operands carry the volume-free backward of ops.OnTheFlyTape."""

    def __init__(self, fmap1: torch.Tensor, fmap2: torch.Tensor, num_levels: int, tape=None):
        self.num_levels, self.tape, self.view_id = num_levels, tape, None
        if tape is not None:
            self.view_id = tape.add_view(fmap1, fmap2)
            out = ops._FeaturePyramidFn.apply(fmap1, fmap2, num_levels, tape, self.view_id)
            self.f1, self.f2 = out[0], list(out[1:])
        else:
            self.f1 = fmap1.permute(0, 2, 3, 1).contiguous()
            self.f2 = ops.channels_last_pyramid(fmap2, num_levels)
        # tensor-core operands (fp16 hi/lo planes): PF_ONTHEFLY_TC=0 keeps the CUDA-core kernel of r01
        self.planes = (ops.OnTheFlyPlanes(self.f1, self.f2)
                       if os.environ.get("PF_ONTHEFLY_TC", "1") != "0" and ops.OnTheFlyPlanes.supported(self.f1) else None)

    def __len__(self):
        return self.num_levels


def corr(fmap1: torch.Tensor, fmap2: torch.Tensor) -> CostVolume:
    """Drop-in for PriOr_RAFT.corr (core/prior_raft.py:69-75)."""
    return CostVolume(fmap1, fmap2)


class DCCL:
    """Dual-cost correlation lookup — core/corr.py:94-144."""

    def __init__(self, num_levels: int = 4, radius: int = 4, mode: str = "auto", volume_mode: Optional[str] = None,
                 accumulate_grads: bool = False):
        if mode not in ("auto", "materialized", "onthefly"):
            raise ValueError("mode must be auto | materialized | onthefly")
        self.num_levels, self.radius, self.mode, self.volume_mode = num_levels, radius, mode, volume_mode
        # training: let all lookups of one backward pass scatter into ONE gradient pyramid per view (ops.GradSink); only
        # valid when the pyramids are consumed by lookups alone and backpropagated once, hence opt-in
        self.accumulate_grads = accumulate_grads
        self._auto = {}
        self._tape = None          # on-the-fly mode under autograd: shared by the pyramids this instance builds

    def _use_onthefly(self, fmap: torch.Tensor) -> bool:
        if self.mode != "auto":
            return self.mode == "onthefly"
        key = (tuple(fmap.shape), str(fmap.device))
        if key not in self._auto:      # decided once per instance and shape: both views must take the same path
            B, C, h, w = fmap.shape
            n = h * w
            need = 2 * (B * n * n * 4 * sum(0.25 ** l for l in range(self.num_levels)) + 4 * B * n * C * 2)   # both views
            self._auto[key] = need > AUTO_MATERIALIZE_FRACTION * available_device_memory(fmap.device) if fmap.is_cuda else True
        return self._auto[key]

    def build_pyramid(self, cost_volume_8):
        if isinstance(cost_volume_8, CostVolume):
            f1, f2 = cost_volume_8.fmap1, cost_volume_8.fmap2
            needs_grad = torch.is_grad_enabled() and (f1.requires_grad or f2.requires_grad)
            if self._use_onthefly(f1):
                if needs_grad:      # volume-free backward: the lookups of this DCCL instance record on one tape (ops.OnTheFlyTape)
                    if self._tape is None:
                        self._tape = ops.OnTheFlyTape(self.radius, self.num_levels)
                    return FeaturePyramid(f1, f2, self.num_levels, tape=self._tape)
                return FeaturePyramid(f1, f2, self.num_levels)
            sink = ops.GradSink() if (self.accumulate_grads and torch.is_grad_enabled()
                                      and (f1.requires_grad or f2.requires_grad)) else None
            pyr = Pyramid(ops.volume_pyramid_autograd(f1, f2, self.num_levels, cost_volume_8.mode or self.volume_mode, sink))
            pyr.grad_sink = sink
            return pyr
        # a materialised [B,h,w,h,w] tensor, as the reference passes (core/corr.py:102-109)
        B, h1, w1, h2, w2 = cost_volume_8.shape
        lvl = cost_volume_8.reshape(B * h1 * w1, 1, h2, w2).float()
        pyr = Pyramid([lvl])
        for _ in range(self.num_levels - 1):
            lvl = ops.avg_pool2x2(lvl) if not lvl.requires_grad else torch.nn.functional.avg_pool2d(lvl, 2, stride=2)
            pyr.append(lvl)
        return pyr

    def __call__(self, coords, corr_pyramid_A, corr_pyramid_B, sample_grid_A2B_W2C_8x, sample_grid_B2A_8x):
        coords = coords.float()
        if isinstance(corr_pyramid_A, FeaturePyramid):
            if corr_pyramid_A.tape is not None and torch.is_grad_enabled():
                return ops.lookup_onthefly_autograd(coords, corr_pyramid_A, corr_pyramid_B, sample_grid_A2B_W2C_8x,
                                                    sample_grid_B2A_8x, self.radius)
            return ops.lookup_onthefly(coords, corr_pyramid_A.f1, corr_pyramid_A.f2, corr_pyramid_B.f1, corr_pyramid_B.f2,
                                       sample_grid_A2B_W2C_8x, sample_grid_B2A_8x, self.radius, cyclic=True,
                                       planes_own=corr_pyramid_A.planes, planes_other=corr_pyramid_B.planes)
        return ops.lookup_autograd(coords, corr_pyramid_A, corr_pyramid_B, sample_grid_A2B_W2C_8x, sample_grid_B2A_8x,
                                   self.radius, cyclic=True, sink_own=getattr(corr_pyramid_A, "grad_sink", None),
                                   sink_other=getattr(corr_pyramid_B, "grad_sink", None))

    def summed(self, coords, corr_pyramid_A, corr_pyramid_B, sample_grid_A2B_W2C_8x, sample_grid_B2A_8x,
               channels_last: bool = False):
        """`corr_own + corr_other` as `PriOr_RAFT.forward` forms it right after the call (core/prior_raft.py:185-188),
        with the add fused into the rotate kernel; optionally in torch.channels_last memory format."""
        coords = coords.float()
        if isinstance(corr_pyramid_A, FeaturePyramid):
            a, b = self(coords, corr_pyramid_A, corr_pyramid_B, sample_grid_A2B_W2C_8x, sample_grid_B2A_8x)
            return a + b
        return ops.lookup_autograd(coords, corr_pyramid_A, corr_pyramid_B, sample_grid_A2B_W2C_8x, sample_grid_B2A_8x,
                                   self.radius, cyclic=True, channels_last=channels_last, fuse_sum=True,
                                   sink_own=getattr(corr_pyramid_A, "grad_sink", None),
                                   sink_other=getattr(corr_pyramid_B, "grad_sink", None))


    def summed_conv(self, coords, corr_pyramid_A, corr_pyramid_B, sample_grid_A2B_W2C_8x, sample_grid_B2A_8x, conv,
                    channels_last: bool = False, fp32: bool = True):
        """`F.relu(conv(corr_own + corr_other))` for the motion encoder's first layer `conv` = Conv2d(324, 256, 1)
        (core/update.py:168,184 / :85,92) without ever forming the [B,324,h,w] sum: inference only; materialised pyramids or
        the tensor-core on-the-fly lookup."""
        if isinstance(corr_pyramid_A, FeaturePyramid):
            if corr_pyramid_A.planes is not None and corr_pyramid_B.planes is not None and not torch.is_grad_enabled():
                return ops.lookup_onthefly_conv(coords.float(), corr_pyramid_A.f1, corr_pyramid_A.f2, corr_pyramid_B.f1, corr_pyramid_B.f2,
                                                sample_grid_A2B_W2C_8x, sample_grid_B2A_8x, corr_pyramid_A.planes, corr_pyramid_B.planes,
                                                conv.weight, conv.bias, channels_last=channels_last, fp32=fp32)
            x = self.summed(coords, corr_pyramid_A, corr_pyramid_B, sample_grid_A2B_W2C_8x, sample_grid_B2A_8x, channels_last)
            return torch.relu(conv(x))
        if torch.is_grad_enabled() and (conv.weight.requires_grad or corr_pyramid_A[0].requires_grad):
            x = self.summed(coords, corr_pyramid_A, corr_pyramid_B, sample_grid_A2B_W2C_8x, sample_grid_B2A_8x, channels_last)
            return torch.relu(conv(x))
        return ops.lookup_conv(coords.float(), corr_pyramid_A, corr_pyramid_B, sample_grid_A2B_W2C_8x, sample_grid_B2A_8x,
                               conv.weight, conv.bias, channels_last=channels_last, fp32=fp32)


class CorrBlock:
    """Plain RAFT correlation block — core/corr.py:13-61 (dead code in PriOr_RAFT.forward, kept for the signature)."""

    def __init__(self, fmap1, fmap2, num_levels: int = 4, radius: int = 4):
        self.num_levels, self.radius = num_levels, radius
        self.corr_pyramid = Pyramid(ops.volume_pyramid_autograd(fmap1.float(), fmap2.float(), num_levels))

    def __call__(self, coords):
        return ops.lookup_autograd(coords.float(), self.corr_pyramid, radius=self.radius, cyclic=False)

    @staticmethod
    def corr(fmap1, fmap2):
        B, _, h, w = fmap1.shape
        return CostVolume(fmap1, fmap2).materialize().view(B, h, w, 1, h, w)


class AlternateCorrBlock:
    """Memory-efficient block — core/corr.py:64-91.  The reference calls the unshipped `alt_cuda_corr`; here the
    on-the-fly kernel computes the same quantity as CorrBlock (non-cyclic sampler) without the volume."""

    def __init__(self, fmap1, fmap2, num_levels: int = 4, radius: int = 4):
        self.num_levels, self.radius = num_levels, radius
        self.pyramid = FeaturePyramid(fmap1.float(), fmap2.float(), num_levels)

    def __call__(self, coords):
        return ops.lookup_onthefly(coords.float(), self.pyramid.f1, self.pyramid.f2, radius=self.radius, cyclic=False)
