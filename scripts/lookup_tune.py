"""Times one DCCL lookup call (lookup + rotate kernels) the way bench.py's `roofline` does — CUDA-graph replay, L2
flushed before every replay, CUDA events — for the tuning variant selected by the PF_LOOKUP_* / PF_ROTATE_* environment
variables.  Usage (GPU box): PF_LOOKUP_Q=48 PF_LOOKUP_OCC=4 python scripts/lookup_tune.py [--smooth]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prior_flow_b200 import ops  # noqa: E402
from prior_flow_b200.model import PriOrRAFT  # noqa: E402
from prior_flow_b200 import geometry as geo  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--smooth", action="store_true", help="smooth ~5 px flow instead of per-pixel N(0, 5^2) noise")
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--fuse-sum", action="store_true")
    a = ap.parse_args()
    dev = "cuda"
    B, H, W = 1, 512, 1024
    h, w = H // 8, W // 8
    g = torch.Generator(device=dev).manual_seed(7)
    fm = [torch.randn(B, 256, h, w, device=dev, generator=g) * 1.45 for _ in range(4)]
    if a.smooth:
        low = torch.randn(B, 2, h // 8, w // 8, device=dev, generator=g) * 5
        coords = geo.coords_grid(B, h, w, dev) + torch.nn.functional.interpolate(low, size=(h, w), mode="bicubic", align_corners=True)
    else:
        coords = geo.coords_grid(B, h, w, dev) + torch.randn(B, 2, h, w, device=dev, generator=g) * 5.0
    grids = PriOrRAFT()._grids(H, W, dev)
    pa, pb = ops.volume_pyramid(fm[0], fm[1], 4), ops.volume_pyramid(fm[2], fm[3], 4)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    fn = lambda: ops.lookup(coords, pa, pb, grids["A2B_W2C_8x"], grids["B2A_8x"], 4, fuse_sum=a.fuse_sum)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        keep = fn()  # noqa: F841
    ts = []
    for _ in range(a.reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        gr.replay()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    # back-to-back replays without a flush (what the model's iteration loop looks like to the L2)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        gr.replay()
    e.record()
    torch.cuda.synchronize()
    # fine-grained flushed timing: the event timer of this box ticks in ~2 us steps, so time 10 x (flush, call) and
    # 10 x (flush) inside graphs and take the difference
    def graph_of(body):
        g_ = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_):
            keep_ = [body() for _ in range(10)]  # noqa: F841
        return g_, keep_
    def both():
        flush.zero_()
        return fn()
    g_both, k1 = graph_of(both)
    g_flush, k2 = graph_of(lambda: flush.zero_())
    def time_graph(g_):
        g_.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s_.record()
            g_.replay()
            e_.record()
            torch.cuda.synchronize()
            best = min(best, s_.elapsed_time(e_))
        return best
    fine_us = (time_graph(g_both) - time_graph(g_flush)) / 10 * 1e3
    env = {k: v for k, v in os.environ.items() if k.startswith("PF_")}
    look_bytes = B * (h * w * 2 * 4 * 100 * 4 + 2 * h * w * 324 * 4 + 3 * 2 * h * w * 4)
    med = ts[len(ts) // 2]
    print(json.dumps({"env": env, "smooth": a.smooth, "us_median_flushed": round(med * 1e3, 2), "us_best_flushed": round(ts[0] * 1e3, 2),
                      "us_back_to_back": round(s.elapsed_time(e) / 20 * 1e3, 2), "us_flushed_fine": round(fine_us, 2), "GBps": round(look_bytes / med / 1e6, 1)}), flush=True)


if __name__ == "__main__":
    main()
