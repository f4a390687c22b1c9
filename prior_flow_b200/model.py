"""PriOr-RAFT with the correlation hot path on the sm_100a kernels — the caller either side of the path.

The hot path has no parameters; what surrounds it (feature/context encoders, the two GRU update blocks) stays on
PyTorch/cuDNN, as BASELINE.json's north_star prescribes.  This module exists because the reference checkout is not
available at run time on the GPU box, so `bench.py`, `smoke()` and the end-to-end parity test need a network to
drive the path with.  Parameter names and shapes are those of the reference (`fnet.*`, `cnet.*`, `ODDC.*`,
`update_block.*`; PriOr-RAFT/core/prior_raft.py:37-41, core/extractor.py:98-158, core/update.py:117-201), so its
checkpoints load with `load_state_dict(strict=True)` — including the DataParallel `module.` prefix handling of
`load_reference_checkpoint`.  Forward semantics follow core/prior_raft.py:107-215 line for line; the differences are
all in *how* the hot-path lines are executed:

  * the eight sample grids come from a per-device cache instead of 8 x 51 eager launches per forward (:115-125);
  * volume + pyramid is one fused tcgen05 launch per view (:151-159), or nothing at all in on-the-fly mode;
  * each DCCL call is 2 launches instead of 168, flo_rotate 1 instead of 123, warp + groupwise_corr 1 instead of 11;
  * in test_mode the convex upsampling runs once, on the last iteration — the reference computes and discards
    the other eleven (:193-213).
"""
from __future__ import annotations

import os
from types import SimpleNamespace
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import geometry as geo
from . import ops
from .corr import DCCL, CostVolume


# --------------------------------------------------------------------------------------- encoders
def _norm(kind: str, ch: int) -> nn.Module:
    if kind == "batch":
        return nn.BatchNorm2d(ch)
    if kind == "instance":
        return nn.InstanceNorm2d(ch)
    if kind == "group":
        return nn.GroupNorm(num_groups=ch // 8, num_channels=ch)
    return nn.Sequential()


FOLD_BN_INFERENCE = True   # eval-mode BatchNorm of the context encoder folded into the preceding convolution (inference only)


def _conv_norm_act(conv: nn.Conv2d, norm: nn.Module, x: torch.Tensor, relu: bool) -> torch.Tensor:
    """relu?(norm(conv(x))).  For an eval-mode BatchNorm in fp32 inference the normalisation is a per-channel affine map
    of the convolution's output, so it is folded into the weights and bias and the whole thing is ONE cuDNN launch
    (conv + bias [+ ReLU]) instead of conv, bias add, a bandwidth-bound BN pass over a full-resolution tensor and a clamp."""
    if (FOLD_BN_INFERENCE and isinstance(norm, nn.BatchNorm2d) and not norm.training and norm.track_running_stats
            and x.is_cuda and x.dtype == torch.float32 and not torch.is_grad_enabled() and not torch.is_autocast_enabled()):
        scale = norm.weight * torch.rsqrt(norm.running_var + norm.eps)
        w = conv.weight * scale.view(-1, 1, 1, 1)
        b = (conv.bias - norm.running_mean) * scale + norm.bias
        if relu:
            return torch.cudnn_convolution_relu(x, w, b, conv.stride, conv.padding, conv.dilation, conv.groups)
        return F.conv2d(x, w, b, conv.stride, conv.padding, conv.dilation, conv.groups)
    y = norm(conv(x))
    return F.relu(y) if relu else y


class ResidualBlock(nn.Module):
    def __init__(self, cin: int, cout: int, norm_fn: str, stride: int = 1):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1, stride=stride)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)
        self.norm1, self.norm2 = _norm(norm_fn, cout), _norm(norm_fn, cout)
        self.downsample = None
        if stride != 1:
            self.norm3 = _norm(norm_fn, cout)   # registered under both names, like the reference's state_dict
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride), self.norm3)

    def forward(self, x):
        y = _conv_norm_act(self.conv1, self.norm1, x, relu=True)
        y = _conv_norm_act(self.conv2, self.norm2, y, relu=True)
        skip = x if self.downsample is None else _conv_norm_act(self.downsample[0], self.downsample[1], x, relu=False)
        return self.relu(skip + y)


class Encoder(nn.Module):
    """BasicEncoder: 7x7/2 stem, three 2-block stages (64, 96/2, 128/2), 1x1 head -> 1/8 resolution."""

    def __init__(self, output_dim: int, norm_fn: str, dropout: float = 0.0):
        super().__init__()
        self.norm1 = _norm(norm_fn, 64) if norm_fn != "group" else nn.GroupNorm(8, 64)
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3)
        self.relu1 = nn.ReLU(inplace=True)
        stages, cin = [], 64
        for cout, stride in ((64, 1), (96, 2), (128, 2)):
            stages.append(nn.Sequential(ResidualBlock(cin, cout, norm_fn, stride), ResidualBlock(cout, cout, norm_fn, 1)))
            cin = cout
        self.layer1, self.layer2, self.layer3 = stages
        self.conv2 = nn.Conv2d(128, output_dim, 1)
        self.dropout = nn.Dropout2d(dropout) if dropout > 0 else None
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, (nn.BatchNorm2d, nn.InstanceNorm2d, nn.GroupNorm)) and m.weight is not None:
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        many = isinstance(x, (tuple, list))
        if many:
            n = x[0].shape[0]
            x = torch.cat(x, dim=0)
        x = _conv_norm_act(self.conv1, self.norm1, x, relu=True)
        x = self.conv2(self.layer3(self.layer2(self.layer1(x))))
        if self.training and self.dropout is not None:
            x = self.dropout(x)
        return torch.split(x, n, dim=0) if many else x


# --------------------------------------------------------------------------------------- update blocks
FUSE_CONV_RELU = True   # inference only; set False to run every conv -> bias add -> ReLU as three ATen launches


def _conv_relu(conv: nn.Conv2d, x: torch.Tensor) -> torch.Tensor:
    """F.relu(conv(x)).  In fp32 inference on CUDA this is ONE cuDNN launch (conv + bias + ReLU in the epilogue) instead
    of cudnn_convolution, a separate bias add and a clamp — 168 of the ~420 element-wise launches of a 12-iteration
    forward (profiles/r02e_e2e_breakdown.txt)."""
    if (FUSE_CONV_RELU and x.is_cuda and x.dtype == torch.float32 and conv.bias is not None and not torch.is_grad_enabled()
            and not torch.is_autocast_enabled()):
        return torch.cudnn_convolution_relu(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation, conv.groups)
    return F.relu(conv(x))


class FlowHead(nn.Module):
    def __init__(self, cin: int = 128, hidden: int = 256):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, hidden, 3, padding=1)
        self.conv2 = nn.Conv2d(hidden, 2, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        return self.conv2(_conv_relu(self.conv1, x))


class SepConvGRU(nn.Module):
    def __init__(self, hidden: int = 128, cin: int = 320):
        super().__init__()
        for tag, k, p in (("1", (1, 5), (0, 2)), ("2", (5, 1), (2, 0))):
            for gate in "zrq":
                setattr(self, f"conv{gate}{tag}", nn.Conv2d(hidden + cin, hidden, k, padding=p))

    def _pass(self, h, x, tag):
        hx = torch.cat([h, x], dim=1)
        z = torch.sigmoid(getattr(self, "convz" + tag)(hx))
        r = torch.sigmoid(getattr(self, "convr" + tag)(hx))
        q = torch.tanh(getattr(self, "convq" + tag)(torch.cat([r * h, x], dim=1)))
        return (1 - z) * h + z * q

    def forward(self, h, x):
        return self._pass(self._pass(h, x, "1"), x, "2")


class MotionEncoder(nn.Module):
    """BasicMotionEncoder (core/update.py:81-99): consumer of the orthogonal view's [B,324,h,w] lookup."""

    def __init__(self, cor_planes: int):
        super().__init__()
        self.convc1 = nn.Conv2d(cor_planes, 256, 1)
        self.convc2 = nn.Conv2d(256, 192, 3, padding=1)
        self.convf1 = nn.Conv2d(2, 128, 7, padding=3)
        self.convf2 = nn.Conv2d(128, 64, 3, padding=1)
        self.conv = nn.Conv2d(64 + 192, 128 - 2, 3, padding=1)

    def forward(self, flow, corr, cor1=None):
        # cor1: relu(convc1(corr)) already computed by the fused rotate + sum + 1x1-conv kernel (ops.lookup_conv)
        cor = _conv_relu(self.convc2, cor1 if cor1 is not None else _conv_relu(self.convc1, corr))
        flo = _conv_relu(self.convf2, _conv_relu(self.convf1, flow))
        return torch.cat([_conv_relu(self.conv, torch.cat([cor, flo], dim=1)), flow], dim=1)


class DualMotionEncoder(nn.Module):
    """BasicMultiMotionEncoder (core/update.py:162-201): own lookup + own/rotated flows + the two 4-ch flaw maps."""

    def __init__(self, cor_planes: int):
        super().__init__()
        self.convc1_A = nn.Conv2d(cor_planes, 256, 1)
        self.convc2_A = nn.Conv2d(256, 128, 3, padding=1)
        self.convf1_A = nn.Conv2d(2, 128, 7, padding=3)
        self.convf2_A = nn.Conv2d(128, 64, 3, padding=1)
        self.convf1_B = nn.Conv2d(2, 128, 7, padding=3)
        self.convf2_B = nn.Conv2d(128, 64, 3, padding=1)
        self.conv_conf1 = nn.Conv2d(8, 32, 3, padding=1)
        self.conv_conf2 = nn.Conv2d(32, 16, 3, padding=1)
        self.conv_A = nn.Conv2d(128 + 64 + 64 + 16, 128 - 4, 3, padding=1)

    def forward(self, flow_A, corr_A, flaw_A, flow_B_A, flaw_B_A, cor1=None):
        cor = _conv_relu(self.convc2_A, cor1 if cor1 is not None else _conv_relu(self.convc1_A, corr_A))
        fa = _conv_relu(self.convf2_A, _conv_relu(self.convf1_A, flow_A))
        fb = _conv_relu(self.convf2_B, _conv_relu(self.convf1_B, flow_B_A))
        conf = _conv_relu(self.conv_conf2, _conv_relu(self.conv_conf1, torch.cat([flaw_A, flaw_B_A], dim=1)))
        out = _conv_relu(self.conv_A, torch.cat([cor, fa, fb, conf], dim=1))
        return torch.cat([out, flow_A, flow_B_A], dim=1)


class _UpdateBase(nn.Module):
    def __init__(self, encoder: nn.Module, hidden: int):
        super().__init__()
        self.encoder = encoder
        self.gru = SepConvGRU(hidden, 128 + hidden)
        self.flow_head = FlowHead(hidden, 256)
        self.mask = nn.Sequential(nn.Conv2d(hidden, 256, 3, padding=1), nn.ReLU(inplace=True), nn.Conv2d(256, 64 * 9, 1))

    def _step(self, net, inp, motion, want_mask=True):
        net = self.gru(net, torch.cat([inp, motion], dim=1))
        mask = 0.25 * self.mask[2](_conv_relu(self.mask[0], net)) if want_mask else None
        return net, mask, self.flow_head(net)


class UpdateBlock(_UpdateBase):          # BasicUpdateBlock, core/update.py:117-136
    def __init__(self, cor_planes: int, hidden: int = 128):
        super().__init__(MotionEncoder(cor_planes), hidden)

    def forward(self, net, inp, corr, flow, want_mask=True, cor1=None):
        return self._step(net, inp, self.encoder(flow, corr, cor1), want_mask)


class DualUpdateBlock(_UpdateBase):      # BasicMultiUpdateBlock ("ODDC"), core/update.py:139-159
    def __init__(self, cor_planes: int, hidden: int = 128):
        super().__init__(DualMotionEncoder(cor_planes), hidden)

    def forward(self, net, inp, flow_A, corr_A, flaw_A, flow_B_A, flaw_B_A, want_mask=True, cor1=None):
        return self._step(net, inp, self.encoder(flow_A, corr_A, flaw_A, flow_B_A, flaw_B_A, cor1), want_mask)


# --------------------------------------------------------------------------------------- the model
def convex_upsample(flow: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """core/prior_raft.py:58-67 — [B,2,h,w] -> [B,2,8h,8w] as a convex combination of the 3x3 neighbourhood."""
    B, _, h, w = flow.shape
    if flow.is_cuda and flow.dtype == torch.float32 and mask.dtype == torch.float32 and os.environ.get("PF_UPSAMPLE_KERNEL", "1") != "0":
        return ops.convex_upsample(flow, mask)       # one launch forward, one backward; eager chain below otherwise
    mask = torch.softmax(mask.view(B, 1, 9, 8, 8, h, w), dim=2)
    nb = F.unfold(8 * flow, [3, 3], padding=1).view(B, 2, 9, 1, 1, h, w)
    up = torch.sum(mask * nb, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(B, 2, 8 * h, 8 * w)


class PriOrRAFT(nn.Module):
    def __init__(self, mixed_precision: bool = False, dropout: float = 0.0, corr_mode: str = "auto",
                 volume_mode: Optional[str] = None):
        super().__init__()
        self.args = SimpleNamespace(mixed_precision=mixed_precision, dropout=dropout, corr_levels=4, corr_radius=4)
        self.hidden_dim = self.context_dim = 128
        self.corr_mode, self.volume_mode = corr_mode, volume_mode
        self.channels_last = False      # set by .to_channels_last(): lookups then emit torch.channels_last tensors
        # training: all lookups of a backward pass scatter into one gradient pyramid per view (ops.GradSink)
        self.accumulate_grads = os.environ.get("PF_GRAD_SINK", "1") != "0"
        # inference: img_rotate + `corr_A + corr_B_A` + the motion encoders' first layer (Conv2d(324, 256, 1) + ReLU) as one
        # tcgen05 kernel behind the gather; the [B,324,h,w] lookup tensor is then never formed
        self.fuse_conv1 = os.environ.get("PF_FUSE_CONV1", "1") != "0"
        self._auto_corr = {}
        cor_planes = 4 * 9 * 9
        self.fnet = Encoder(256, "instance", dropout)
        self.cnet = Encoder(256, "batch", dropout)
        self.ODDC = DualUpdateBlock(cor_planes, 128)
        self.update_block = UpdateBlock(cor_planes, 128)

    def to_channels_last(self):
        """NHWC weights for the context encoder and the two update blocks + NHWC lookup outputs: cuDNN's tensor-core
        kernels then run without an NCHW<->NHWC conversion around every convolution.  The feature encoder stays NCHW:
        its InstanceNorm layers are NCHW-only in ATen and would convert back and forth at full resolution."""
        self.channels_last = True
        for m in (self.cnet, self.ODDC, self.update_block):
            m.to(memory_format=torch.channels_last)
        return self

    def graphed(self, iters: int = 12) -> "GraphedForward":
        """CUDA-graph replay of `self(image1, image2, iters, test_mode=True)`."""
        return GraphedForward(self, iters)

    def freeze_bn(self):
        for m in self.modules():
            if isinstance(m, (nn.BatchNorm2d, nn.SyncBatchNorm)):
                m.eval()

    def load_reference_checkpoint(self, path: str, strict: bool = True):
        sd = torch.load(path, map_location="cpu")
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
        return self.load_state_dict(sd, strict=strict)

    # ---- hot-path geometry: input independent, cached per (size, device)
    def _grids(self, H: int, W: int, device):
        out = {}
        for tag, angles in (("A2B", geo.A2B), ("B2A", geo.B2A)):
            R = geo.rotation_matrix_host(list(angles))
            Rt = R.T.contiguous()
            out[tag] = geo.samplegrid_cached(H, W, R, device)
            out[tag + "_8x"] = geo.samplegrid_cached(H // 8, W // 8, R, device)
            out[tag + "_W2C_8x"] = geo.samplegrid_cached(H // 8, W // 8, Rt, device)
        return out

    def forward(self, image1, image2, iters: int = 12, init_flow=None, test_mode: bool = False):
        amp = lambda: torch.autocast("cuda", enabled=self.args.mixed_precision)
        image1 = (2 * (image1 / 255.0) - 1.0).contiguous()
        image2 = (2 * (image2 / 255.0) - 1.0).contiguous()
        B, _, H, W = image1.shape
        g = self._grids(H, W, image1.device)

        # orthogonal view of both frames: one 6-channel remap (prior_raft.py:127)
        both_B = geo.img_rotate(torch.cat([image1, image2], dim=1), sample_grid=g["A2B"])
        image1_B, image2_B = both_B[:, :3].contiguous(), both_B[:, 3:].contiguous()

        with amp():
            cA, cB = self.cnet([image1, image1_B])
            net_A, inp_A = torch.tanh(cA[:, :128]), torch.relu(cA[:, 128:])
            net_B, inp_B = torch.tanh(cB[:, :128]), torch.relu(cB[:, 128:])
            f1A, f2A, f1B, f2B = (f.float() for f in self.fnet([image1, image2, image1_B, image2_B]))

        lookup = DCCL(4, 4, mode=self.corr_mode, volume_mode=self.volume_mode, accumulate_grads=self.accumulate_grads)
        lookup._auto = self._auto_corr          # the materialise / on-the-fly decision is made once per shape for the model's lifetime
        pyr_A = lookup.build_pyramid(CostVolume(f1A, f2A))
        pyr_B = lookup.build_pyramid(CostVolume(f1B, f2B))

        h, w = H // 8, W // 8
        coords0 = geo.coords_grid(B, h, w, image1.device)
        coords1_A, coords1_B = coords0.clone(), coords0.clone()
        if init_flow is not None:
            coords1_A = coords1_A + init_flow
            coords1_B = coords1_B + geo.flo_rotate(init_flow, sample_grid_W2C=g["A2B_W2C_8x"], sample_grid_C2W=g["A2B_8x"])

        preds_A: List[torch.Tensor] = []
        preds_B: List[torch.Tensor] = []
        up_A = None
        for it in range(iters):
            last = it == iters - 1
            want_up = last or not test_mode
            coords1_A, coords1_B = coords1_A.detach(), coords1_B.detach()
            flow_A, flow_B = coords1_A - coords0, coords1_B - coords0
            flaw_A = ops.warp_groupcorr_autograd(f1A, f2A, coords1_A, 4)
            flow_B_A = geo.flo_rotate(flow_B, sample_grid_W2C=g["B2A_W2C_8x"], sample_grid_C2W=g["B2A_8x"])
            flaw_B_A = ops.warp_groupcorr_autograd(f1A, f2A, coords0 + flow_B_A, 4)
            with amp():
                # corr_A + corr_B_A and corr_B + corr_A_B (prior_raft.py:185-188), the adds fused into the rotate kernel
                fused1 = (self.fuse_conv1 and not torch.is_grad_enabled() and not self.args.mixed_precision
                          and (isinstance(pyr_A, list) or getattr(pyr_A, "planes", None) is not None))
                corr_A = corr_B = cor1_A = cor1_B = None
                if fused1:
                    fp32 = not torch.backends.cudnn.allow_tf32     # the precision cuDNN would run this layer in
                    cor1_A = lookup.summed_conv(coords1_A, pyr_A, pyr_B, g["A2B_W2C_8x"], g["B2A_8x"], self.ODDC.encoder.convc1_A,
                                                self.channels_last, fp32)
                    cor1_B = lookup.summed_conv(coords1_B, pyr_B, pyr_A, g["B2A_W2C_8x"], g["A2B_8x"], self.update_block.encoder.convc1,
                                                self.channels_last, fp32)
                else:
                    corr_A = lookup.summed(coords1_A, pyr_A, pyr_B, g["A2B_W2C_8x"], g["B2A_8x"], self.channels_last)
                    corr_B = lookup.summed(coords1_B, pyr_B, pyr_A, g["B2A_W2C_8x"], g["A2B_8x"], self.channels_last)
                if self.channels_last:
                    # torch.cat falls back to NCHW as soon as ONE input is NCHW (and every convolution behind it then
                    # pays a layout copy): hand the update blocks their 2- and 4-channel inputs in NHWC as well
                    cl = lambda t: t.contiguous(memory_format=torch.channels_last)
                    flow_A, flow_B, flow_B_A, flaw_A, flaw_B_A = cl(flow_A), cl(flow_B), cl(flow_B_A), cl(flaw_A), cl(flaw_B_A)
                net_A, mask_A, d_A = self.ODDC(net_A, inp_A, flow_A, corr_A, flaw_A, flow_B_A, flaw_B_A, want_mask=want_up, cor1=cor1_A)
                net_B, mask_B, d_B = self.update_block(net_B, inp_B, corr_B, flow_B, want_mask=want_up and not test_mode, cor1=cor1_B)
            coords1_A = coords1_A + d_A
            coords1_B = coords1_B + d_B
            if want_up:
                up_A = convex_upsample(coords1_A - coords0, mask_A)
                if not test_mode:
                    preds_A.append(up_A)
                    preds_B.append(convex_upsample(coords1_B - coords0, mask_B))
        if test_mode:
            return up_A
        return preds_A, preds_B


class GraphedForward:
    """Inference through a CUDA graph: the whole forward — sample-grid lookups, both volume builds, the
    12-iteration loop with its ~90 launches of ours and ~1100 of cuDNN/ATen — is captured once per input shape and replayed;
    the Python loop, the allocator and 1200 launch latencies leave the critical path.

        run = model.graphed(iters=12)
        flow = run(image1, image2)        # [B,2,H,W]; valid until the next call with the same shape (a static buffer)

    Inputs are copied into static buffers (device-to-device, or straight from pinned host memory), so callers may pass fresh
    tensors every time.  Weights are baked in by address: after `load_state_dict` / an optimizer step that REPLACES parameter
    tensors call `reset()`; in-place updates are picked up automatically."""

    def __init__(self, model: "PriOrRAFT", iters: int = 12, warmup: int = 2):
        self.model, self.iters, self.warmup = model, iters, warmup
        self._graphs = {}

    def reset(self) -> None:
        self._graphs.clear()

    def _capture(self, image1: torch.Tensor, image2: torch.Tensor):
        dev = next(self.model.parameters()).device
        s1 = torch.empty(image1.shape, device=dev, dtype=torch.float32)
        s2 = torch.empty(image2.shape, device=dev, dtype=torch.float32)
        s1.copy_(image1), s2.copy_(image2)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(self.warmup):           # cuDNN autotuning, sample-grid cache, prepared conv weights: all before capture
                self.model(s1, s2, iters=self.iters, test_mode=True)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph), torch.no_grad():
            out = self.model(s1, s2, iters=self.iters, test_mode=True)
        return graph, s1, s2, out

    def __call__(self, image1: torch.Tensor, image2: torch.Tensor) -> torch.Tensor:
        if self.model.training:
            raise RuntimeError("GraphedForward is an inference path: call model.eval() first")
        key = (tuple(image1.shape), tuple(image2.shape))
        entry = self._graphs.get(key)
        if entry is None:
            entry = self._graphs[key] = self._capture(image1, image2)
        graph, s1, s2, out = entry
        s1.copy_(image1, non_blocking=True)
        s2.copy_(image2, non_blocking=True)
        graph.replay()
        return out
