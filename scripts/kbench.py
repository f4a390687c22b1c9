"""Per-kernel timings of the hot path at BASELINE config 2 (512x1024 -> 64x128, C=256), CUDA events, L2 flushed
between iterations.  Prints one JSON line per kernel with achieved GB/s (or TFLOP/s) against MEASURED_PEAKS.json.
Usage (GPU box): python scripts/kbench.py [--iters 20] [--h 64 --w 128]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prior_flow_b200 import ops  # noqa: E402
from oracle import torch_oracle as TO  # noqa: E402  (timed beside ours as the eager-PyTorch GPU bar)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return p["hbm_gbs"], p["bf16_tflops"], "measured"
    except Exception:
        return 6650.0, 1590.0, "fallback"


def timeit(fn, iters, flush):
    """Per-call GPU time, conservative: the call is captured in a CUDA graph (no host work in the timed region), the L2
    is flushed with a memset before every replay and CUDA events bracket ONE replay — so a number includes the graph
    launch latency (~4 us), the 2 us tick of the event timer and the write-back of the flush's dirty lines.  bench.py's
    `roofline` uses K calls on K different working sets in one graph instead; use that for the small kernels."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    run = fn
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep = fn()  # noqa: F841
        run = g.replay
    except Exception as ex:  # noqa: BLE001
        print(f"[kbench] graph capture failed ({type(ex).__name__}); eager timing", file=sys.stderr)
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        run()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--h", type=int, default=64)
    ap.add_argument("--w", type=int, default=128)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--skip-torch", action="store_true")
    ap.add_argument("--out", default="", help="write the rows as JSON here (never set this for a run under ncu: its times are not timings)")
    ap.add_argument("--only", default="", help="comma-separated substrings: run only the kernels whose name matches one")
    ap.add_argument("--cpu", action="store_true", help="also time the ATen restatement of each op on the host cores (reference's CPU path)")
    a = ap.parse_args()
    B, h, w, C = a.batch, a.h, a.w, 256
    N = h * w
    hbm, tflops, src = peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    fm = [torch.randn(B, C, h, w, device="cuda", generator=g) * 1.45 for _ in range(4)]
    # a smooth flow field (low-resolution noise, bicubically upsampled, ~5 px): neighbouring pixels move together
    low = torch.randn(B, 2, h // 8, w // 8, device="cuda", generator=g) * 5
    coords = TO.coords_grid(B, h, w, "cuda") + torch.nn.functional.interpolate(low, size=(h, w), mode="bicubic", align_corners=True)
    Ra, Rb = TO.rotation_matrix([0., 0., -np.pi / 2], device="cuda"), TO.rotation_matrix([0., 0., np.pi / 2], device="cuda")
    gw = ops.samplegrid((1, 3, h, w), Ra.T.contiguous())
    gc = ops.samplegrid((1, 3, h, w), Rb)
    img = torch.rand(B, 6, 8 * h, 8 * w, device="cuda")
    gfull = ops.samplegrid((1, 3, 8 * h, 8 * w), Ra)
    rows = []

    import time as _time
    cpu_threads = os.cpu_count() or 1
    torch.set_num_threads(cpu_threads)

    def cpu_time(make_fn):
        """best of 3 wall-clock runs of the same op restated with ATen ops on CPU tensors"""
        f = make_fn()
        f()
        ts = []
        for _ in range(3):
            t0 = _time.perf_counter()
            f()
            ts.append(_time.perf_counter() - t0)
        return min(ts) * 1e3

    only = [t for t in a.only.split(",") if t]

    def rec(name, fn, bytes_=None, flops=None, torch_fn=None, cpu_fn=None):
        if only and not any(t in name for t in only):
            return
        med, best = timeit(fn, a.iters, flush)
        r = {"kernel": name, "ms_median": round(med, 4), "ms_best": round(best, 4)}
        if cpu_fn is not None and a.cpu:
            r["cpu_reference_ms"] = round(cpu_time(cpu_fn), 2)
            r["cpu_threads"] = cpu_threads
        if bytes_:
            r["GBps"] = round(bytes_ / med / 1e6, 1)
            r["frac_hbm"] = round(bytes_ / med / 1e6 / hbm, 3)
        if flops:
            r["TFLOPs"] = round(flops / med / 1e9, 1)
            r["frac_bf16_peak"] = round(flops / med / 1e9 / tflops, 3)
        if torch_fn is not None and not a.skip_torch:
            r["torch_eager_ms"] = round(timeit(torch_fn, max(3, a.iters // 4), flush)[0], 4)
        r["peaks"] = src
        rows.append(r)
        print(json.dumps(r), flush=True)

    # what this GPU sustains for a pure-write and a copy stream (the volume build is write-only)
    big = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    big2 = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    rec("hbm_fill_1GiB (write-only stream)", lambda: big.fill_(1), float(1 << 30))
    rec("hbm_copy_1GiB (read+write stream)", lambda: big2.copy_(big), float(2 << 30))
    del big, big2
    pyr_bytes = sum(B * N * (h >> l) * (w >> l) * 4 for l in range(4))
    vol_bytes = pyr_bytes + 2 * B * C * N * 4
    vol_flops = 2.0 * B * N * N * C
    for mode in ("fp32", "f16", "fp32_simt"):
        rec(f"volume_pyramid[{mode}]", lambda m=mode: ops.volume_pyramid(fm[0], fm[1], 4, m), vol_bytes, vol_flops,
            (lambda: TO.build_pyramid(TO.corr_volume(fm[0], fm[1]))) if mode == "fp32" else None,
            (lambda: (lambda a_=fm[0].cpu(), b_=fm[1].cpu(): TO.build_pyramid(TO.corr_volume(a_, b_)))) if mode == "fp32" else None)
    pa, pb = ops.volume_pyramid(fm[0], fm[1], 4), ops.volume_pyramid(fm[2], fm[3], 4)
    K2 = 81
    look_bytes = B * N * 2 * 4 * 100 * 4 + 2 * B * N * 4 * K2 * 4 + 3 * B * 2 * N * 4
    def cpu_lookup():
        c_, pa_, pb_ = coords.cpu(), [t.cpu() for t in pa], [t.cpu() for t in pb]
        gw_, gc_ = gw.expand(B, -1, -1, -1).cpu(), gc.expand(B, -1, -1, -1).cpu()
        return lambda: TO.dccl_lookup(c_, pa_, pb_, gw_, gc_, 4)
    rec("lookup_dual", lambda: ops.lookup(coords, pa, pb, gw, gc, 4), look_bytes, None,
        lambda: TO.dccl_lookup(coords, pa, pb, gw.expand(B, -1, -1, -1), gc.expand(B, -1, -1, -1), 4), cpu_lookup)
    conv = torch.nn.Conv2d(324, 256, 1).cuda()
    def unfused(tf32):
        def f():
            torch.backends.cudnn.allow_tf32 = tf32
            with torch.no_grad():
                return torch.relu(conv(ops.lookup(coords, pa, pb, gw, gc, 4, fuse_sum=True, channels_last=True)))
        return f
    conv = conv.to(memory_format=torch.channels_last)
    rec("lookup_fused_sum(cl) + cuDNN conv1x1 + relu [fp32]", unfused(False), look_bytes)
    rec("lookup_fused_sum(cl) + cuDNN conv1x1 + relu [tf32]", unfused(True), look_bytes)
    rec("lookup_conv[fp32, cl] (pf_dccl_conv)", lambda: ops.lookup_conv(coords, pa, pb, gw, gc, conv.weight, conv.bias, channels_last=True, fp32=True), look_bytes)
    rec("lookup_conv[tf32-class, cl] (pf_dccl_conv)", lambda: ops.lookup_conv(coords, pa, pb, gw, gc, conv.weight, conv.bias, channels_last=True, fp32=False), look_bytes)
    rec("lookup_fused_sum(cl)", lambda: ops.lookup(coords, pa, pb, gw, gc, 4, fuse_sum=True, channels_last=True), look_bytes)
    rec("lookup_single", lambda: ops.lookup(coords, pa, radius=4, cyclic=True), look_bytes // 2)
    flow = coords - TO.coords_grid(B, h, w, "cuda")
    rec("flo_rotate", lambda: ops.flo_rotate(flow, gw, gc), 4 * B * 2 * N * 4, None,
        lambda: TO.flo_rotate(flow, gw.expand(B, -1, -1, -1), gc.expand(B, -1, -1, -1)),
        lambda: (lambda f_=flow.cpu(), a_=gw.expand(B, -1, -1, -1).cpu(), b_=gc.expand(B, -1, -1, -1).cpu(): TO.flo_rotate(f_, a_, b_)))
    rec("warp_groupcorr", lambda: ops.warp_groupcorr(fm[0], fm[1], coords, 4), 2 * B * C * N * 4 + B * 6 * N * 4, None,
        lambda: TO.warp_groupcorr(fm[0], fm[1], coords, 4),
        lambda: (lambda a_=fm[0].cpu(), b_=fm[1].cpu(), c_=coords.cpu(): TO.warp_groupcorr(a_, b_, c_, 4)))
    rec("img_rotate_fullres", lambda: ops.remap(img, gfull, "B2HW", True), (2 * 6 + 2) * B * 64 * N * 4, None,
        lambda: TO.img_rotate(img, gfull.expand(B, -1, -1, -1)),
        lambda: (lambda i_=img.cpu(), g_=gfull.expand(B, -1, -1, -1).cpu(): TO.img_rotate(i_, g_)))
    rec("samplegrid_fullres", lambda: ops.samplegrid((1, 3, 8 * h, 8 * w), Ra), 2 * 64 * N * 4, None,
        lambda: TO.generate_samplegrid((1, 3, 8 * h, 8 * w), Ra),
        lambda: (lambda r_=Ra.cpu(): TO.generate_samplegrid((1, 3, 8 * h, 8 * w), r_)))
    dV = torch.randn(B, N, N, device="cuda", generator=g)
    bwd_flops = 2 * 2.0 * B * N * N * C
    rec("volume_backward[tcgen05 bf16x2, both gradients]", lambda: ops.volume_backward(fm[0], fm[1], dV), 2 * B * N * N * 4 + 4 * B * C * N * 4, bwd_flops)
    torch.backends.cuda.matmul.allow_tf32 = False
    rec("volume_backward[cuBLAS fp32, two GEMMs]", lambda: ops.volume_backward(fm[0], fm[1], dV, use_library=True), 2 * B * N * N * 4 + 4 * B * C * N * 4, bwd_flops)
    del dV
    mask = torch.randn(B, 576, h, w, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
    from prior_flow_b200 import model as M
    rec("convex_upsample[kernel, channels_last mask]", lambda: ops.convex_upsample(flow, mask), B * (576 + 2 + 128) * N * 4, None,
        lambda: torch.sum(torch.softmax(mask.view(B, 1, 9, 8, 8, h, w), dim=2) * torch.nn.functional.unfold(8 * flow, [3, 3], padding=1).view(B, 2, 9, 1, 1, h, w), dim=2))
    cl = lambda t: t.permute(0, 2, 3, 1).contiguous()
    f1a, f2a, f1b, f2b = cl(fm[0]), ops.channels_last_pyramid(fm[1], 4), cl(fm[2]), ops.channels_last_pyramid(fm[3], 4)
    rec("lookup_onthefly", lambda: ops.lookup_onthefly(coords, f1a, f2a, f1b, f2b, gw, gc, 4))
    pla, plb = ops.OnTheFlyPlanes(f1a, f2a), ops.OnTheFlyPlanes(f1b, f2b)
    rec("lookup_onthefly[tcgen05 dots, smooth flow]", lambda: ops.lookup_onthefly(coords, f1a, f2a, f1b, f2b, gw, gc, 4, planes_own=pla, planes_other=plb))
    noisy = TO.coords_grid(B, h, w, "cuda") + torch.randn(B, 2, h, w, device="cuda", generator=g) * 5.0
    rec("lookup_onthefly[CUDA cores, iid sigma-5 flow]", lambda: ops.lookup_onthefly(noisy, f1a, f2a, f1b, f2b, gw, gc, 4))
    rec("lookup_onthefly[tcgen05 dots, iid sigma-5 flow]", lambda: ops.lookup_onthefly(noisy, f1a, f2a, f1b, f2b, gw, gc, 4, planes_own=pla, planes_other=plb))
    rec("onthefly_planes[absmax + split, one view]", lambda: ops.OnTheFlyPlanes(f1a, f2a))
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as fh:
            json.dump(rows, fh, indent=1)


if __name__ == "__main__":
    main()
