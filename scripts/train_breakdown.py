"""Per-kernel GPU time of one PriOr-RAFT training step (512x1024, 12 iters, batch 1) via torch.profiler."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prior_flow_b200 import distributed as pfd  # noqa: E402
from prior_flow_b200.model import PriOrRAFT  # noqa: E402
from prior_flow_b200.train import train_step  # noqa: E402

ctx = pfd.init_from_env("nccl")
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
model = PriOrRAFT().to(ctx.device)
model.train()
model.freeze_bn()
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5, eps=1e-8)
g = torch.Generator(device=ctx.device).manual_seed(100)
B, H, W = 1, 512, 1024
batch = (torch.rand(B, 3, H, W, device=ctx.device, generator=g) * 255, torch.rand(B, 3, H, W, device=ctx.device, generator=g) * 255,
         torch.randn(B, 2, H, W, device=ctx.device, generator=g) * 5, torch.ones(B, H, W, device=ctx.device))
for _ in range(2):
    train_step(model, opt, batch, ctx)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    train_step(model, opt, batch, ctx)
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0 and not (e.key.startswith("aten::") or e.key.startswith("cudnn_") or e.key.startswith("cudaLaunch") or "Backward" in e.key or e.key.startswith("autograd::") or e.key.startswith("Optimizer"))]
ev.sort(key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in ev)
ours = sum(e.device_time_total for e in ev if "pf::" in e.key)
print(f"train step: total kernel time {tot / 1e3:.2f} ms ; ours (pf::) {ours / 1e3:.2f} ms = {100 * ours / tot:.1f}%")
for e in ev[:28]:
    print(f"{e.device_time_total / 1e3:9.3f} ms {100 * e.device_time_total / tot:5.1f}%  n={e.count:5d}  {e.key[:105]}")
