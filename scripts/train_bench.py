#!/usr/bin/env python
"""BASELINE config 5: one PriOr-RAFT training step (512x1024, 12 iters) with the backward kernels, DDP over NCCL.

    python scripts/train_bench.py [--batch 1] [--steps 5]                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/train_bench.py --batch 1
Prints one JSON line (rank 0): ms per step (max over ranks), loss trajectory, peak memory."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prior_flow_b200 import distributed as pfd  # noqa: E402
from prior_flow_b200.model import PriOrRAFT  # noqa: E402
from prior_flow_b200.train import train_step  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1, help="pairs per GPU")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=12)
    ap.add_argument("--memory-format", default="nchw", choices=["nchw", "channels_last"])
    ap.add_argument("--corr-mode", default="auto", choices=["auto", "materialized", "onthefly"],
                    help="onthefly: volume-free forward (tensor-core lookup) and backward (ops.OnTheFlyTape)")
    ap.add_argument("--ddp", default="ddp", choices=["ddp", "static", "nobroadcast", "none"],
                    help="ddp: DistributedDataParallel defaults; static: static_graph=True; nobroadcast: broadcast_buffers=False (BN is frozen); "
                         "none: independent replicas, no gradient all-reduce (isolates host contention from DDP's cost)")
    a = ap.parse_args()
    ctx = pfd.init_from_env("nccl")
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    model = PriOrRAFT(corr_mode=a.corr_mode).to(ctx.device)
    if a.memory_format == "channels_last":
        model = model.to_channels_last()
    model.train()
    model.freeze_bn()
    if a.ddp == "none" or ctx.world == 1:
        ddp = model
    else:
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[ctx.local_rank], gradient_as_bucket_view=True,
                                                        static_graph=a.ddp == "static", broadcast_buffers=a.ddp == "ddp")
    opt = torch.optim.AdamW(ddp.parameters(), lr=1e-4, weight_decay=1e-5, eps=1e-8)
    g = torch.Generator(device=ctx.device).manual_seed(100 + ctx.rank)
    B, H, W = a.batch, a.height, a.width
    im1 = torch.rand(B, 3, H, W, device=ctx.device, generator=g) * 255
    im2 = torch.rand(B, 3, H, W, device=ctx.device, generator=g) * 255
    low = torch.randn(B, 2, H // 64, W // 64, device=ctx.device, generator=g) * 8
    flow_gt = torch.nn.functional.interpolate(low, size=(H, W), mode="bicubic", align_corners=True)
    valid = torch.ones(B, H, W, device=ctx.device)
    losses = []
    for _ in range(a.warmup):
        losses.append(train_step(ddp, opt, (im1, im2, flow_gt, valid), ctx, iters=a.iters)["loss"])
    torch.cuda.synchronize()
    if ctx.world > 1:
        torch.distributed.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.steps):
        losses.append(train_step(ddp, opt, (im1, im2, flow_gt, valid), ctx, iters=a.iters)["loss"])
    e.record()
    torch.cuda.synchronize()
    ms = pfd.max_over_ranks(s.elapsed_time(e) / a.steps, ctx)
    if ctx.rank == 0:
        print(json.dumps({"what": "train step (config 5)", "n_gpus": ctx.world, "global_batch": B * ctx.world, "ms_per_step": round(ms, 2),
                          "pairs_per_s": round(B * ctx.world / ms * 1e3, 2), "losses": [round(x, 4) for x in losses],
                          "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2), "shape": [H, W], "iters": a.iters, "ddp": a.ddp if ctx.world > 1 else "n/a", "memory_format": a.memory_format,
                          "corr_mode": a.corr_mode}), flush=True)
    if ctx.world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
