"""Per-kernel GPU time of one training step (512x1024, 12 iterations, batch 1): which kernels of ours and which eager ops weigh."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from prior_flow_b200 import distributed as pfd  # noqa: E402
from prior_flow_b200.model import PriOrRAFT  # noqa: E402
from prior_flow_b200.train import train_step  # noqa: E402

torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
ctx = pfd.init_from_env("nccl")
model = PriOrRAFT().cuda().train()
model.freeze_bn()
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5, eps=1e-8)
g = torch.Generator(device="cuda").manual_seed(1)
H, W = 512, 1024
im1, im2 = torch.rand(1, 3, H, W, device="cuda", generator=g) * 255, torch.rand(1, 3, H, W, device="cuda", generator=g) * 255
low = torch.randn(1, 2, H // 64, W // 64, device="cuda", generator=g) * 8
gt = torch.nn.functional.interpolate(low, size=(H, W), mode="bicubic", align_corners=True)
valid = torch.ones(1, H, W, device="cuda")
for _ in range(3):
    train_step(model, opt, (im1, im2, gt, valid), ctx, iters=12)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    train_step(model, opt, (im1, im2, gt, valid), ctx, iters=12)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
total = sum(e.device_time_total for e in rows)
ours = sum(e.device_time_total for e in rows if "pf::" in e.key)
print(f"total kernel time {total / 1e3:.2f} ms; ours (pf::) {ours / 1e3:.2f} ms = {100 * ours / total:.1f}%")
for e in rows[:40]:
    print(f"{e.device_time_total / 1e3:9.3f} ms {100 * e.device_time_total / total:5.1f}%  n={e.count:5d}  {e.key[:130]}")
