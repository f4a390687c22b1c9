"""ATen-level attribution of one PriOrRAFT forward (512x1024, 12 iters): device time per (op, input shapes).
Usage (GPU box): PF_CHANNELS_LAST=0|1 python scripts/e2e_ops.py"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prior_flow_b200.model import PriOrRAFT  # noqa: E402

CL = os.environ.get("PF_CHANNELS_LAST", "0") == "1"
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
model = PriOrRAFT().cuda().eval()
if CL:
    model = model.to_channels_last()
g = torch.Generator().manual_seed(1234)
im1 = (torch.rand(1, 3, 512, 1024, generator=g) * 255).cuda()
im2 = (torch.rand(1, 3, 512, 1024, generator=g) * 255).cuda()
if CL:
    im1, im2 = im1.contiguous(memory_format=torch.channels_last), im2.contiguous(memory_format=torch.channels_last)
with torch.no_grad():
    for _ in range(3):
        model(im1, im2, iters=12, test_mode=True)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
        model(im1, im2, iters=12, test_mode=True)
        torch.cuda.synchronize()
ev = [e for e in prof.key_averages(group_by_input_shape=True) if e.self_device_time_total > 0 and e.key.startswith("aten::")]
ev.sort(key=lambda e: -e.self_device_time_total)
tot = sum(e.self_device_time_total for e in ev)
print(f"channels_last={CL}: aten self device time {tot / 1e3:.2f} ms")
for e in ev[:45]:
    print(f"{e.self_device_time_total / 1e3:8.3f} ms n={e.count:4d} {e.key:32s} {str(e.input_shapes)[:110]}")
