"""Per-kernel GPU time of one PriOrRAFT forward (512x1024, 12 iters) via torch.profiler (CUPTI).
Shows how the step splits between our hot-path kernels and the unchanged cuDNN/ATen side."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prior_flow_b200.model import PriOrRAFT  # noqa: E402

CL = os.environ.get("PF_CHANNELS_LAST", "0") == "1"
torch.backends.cudnn.benchmark = os.environ.get("PF_CUDNN_BENCHMARK", "0") == "1"
TAG = os.environ.get("PF_TAG", "default")
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
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        model(im1, im2, iters=12, test_mode=True)
        torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
ev.sort(key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in ev if not e.key.startswith("aten::") and not e.key.startswith("cudnn"))
rows = []
ktot = 0.0
for e in ev:
    if e.key.startswith("aten::") or e.key.startswith("cudnn_") or e.key.startswith("cudaLaunch"):
        continue
    ktot += e.device_time_total
for e in ev:
    if e.key.startswith("aten::") or e.key.startswith("cudnn_") or e.key.startswith("cudaLaunch"):
        continue
    rows.append(f"{e.device_time_total / 1e3:9.3f} ms {100 * e.device_time_total / ktot:5.1f}%  n={e.count:5d}  {e.key[:110]}")
ours = sum(e.device_time_total for e in ev if "pf::" in e.key)
out = [f"[{TAG}] channels_last={CL} cudnn.benchmark={torch.backends.cudnn.benchmark}", f"total kernel time {ktot / 1e3:.3f} ms ; ours (pf::) {ours / 1e3:.3f} ms = {100 * ours / ktot:.1f}%"] + rows[:30]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", f"e2e_breakdown_{TAG}.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
