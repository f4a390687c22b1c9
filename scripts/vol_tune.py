"""Times volume_pyramid per view the way bench.py does (8 builds on 4 different fmap sets in one CUDA graph)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prior_flow_b200 import ops  # noqa: E402
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(7)
sets = [[torch.randn(1, 256, 64, 128, device=dev, generator=g) * 1.45 for _ in range(4)] for _ in range(4)]
calls = [lambda f=s, k=k: ops.volume_pyramid(f[k], f[k + 1], 4) for s in sets for k in (0, 2)]
for c in calls:
    c()
torch.cuda.synchronize()
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    calls[0]()
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    keep = [c() for c in calls]
ts = []
for _ in range(12):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); gr.replay(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
ts = sorted(ts[2:])
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("PF_")}, "us_per_view": round(ts[len(ts) // 2] / len(calls) * 1e3, 2)}))
