#!/usr/bin/env python
"""Headline benchmark: PriOr-RAFT inference pairs/s at 512x1024 ERP, 12 GRU iterations (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One JSON line on stdout (rank 0).  A "step" is one pass of the model (hot path on the sm_100a kernels, encoders and
GRU blocks on cuDNN) over one batch of synthetic image pairs; pairs shard one process per GPU with no data-path
collective ("scaling": "weak").  `value` is device-resident throughput (inputs in HBM, whole forward replayed from a
CUDA graph), `e2e` goes through the public API with pinned-host inputs and a device->host read of the flow inside the
timed region.  `roofline` is the dominant hot-path kernel pair (one DCCL lookup call = lookup_kernel + rotate_kernel),
timed live with CUDA events; `cpu_baseline` is the eager-ATen restatement of the reference forward on the host cores.
`--impl reference` times that CPU restatement alone (the reference is pure PyTorch and is not present on the box).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pairs/s @512x1024 ERP, 12 iters"
UNIT = "pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1, help="pairs per GPU per step")
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=12)
    ap.add_argument("--volume-mode", default="fp32", choices=["fp32", "f16", "fp32_simt"])
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA-graph replay")
    ap.add_argument("--memory-format", default="channels_last", choices=["nchw", "channels_last"],
                    help="memory format of the cuDNN side; channels_last also makes the lookups emit NHWC directly")
    ap.add_argument("--corr-mode", default="auto", choices=["auto", "materialized", "onthefly"],
                    help="DCCL mode: auto materialises the pyramids while they fit in device memory; BASELINE configs[3] names onthefly")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------- helpers
def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (pynvml, falling back to nvidia-smi)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": "nvmlClocksThrottleReasonHwSlowdown", "hw_thermal_slowdown": "nvmlClocksThrottleReasonHwThermalSlowdown",
                 "sw_thermal_slowdown": "nvmlClocksThrottleReasonSwThermalSlowdown", "sw_power_cap": "nvmlClocksThrottleReasonSwPowerCap"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for tag, attr in names.items():
                    if mask & getattr(nv, attr, 0):
                        self.reasons.add(tag)
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=10).stdout.split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [], "note": "single nvidia-smi sample after the run"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "no NVML / nvidia-smi"}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def synthetic_pair(batch, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(batch, 3, H, W, generator=g) * 255).pin_memory() if torch.cuda.is_available() else torch.rand(batch, 3, H, W, generator=g) * 255


# ----------------------------------------------------------------------------------------- CPU arm
def cpu_reference_run(a, steps, warmup, iters):
    """Eager-ATen restatement of the reference forward on the host cores (oracle/cpu_model.py)."""
    from oracle.cpu_model import EagerPriOrRAFT
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    model = EagerPriOrRAFT().eval()
    g = torch.Generator().manual_seed(1234)
    im1 = torch.rand(a.batch, 3, a.height, a.width, generator=g) * 255
    im2 = torch.rand(a.batch, 3, a.height, a.width, generator=g) * 255
    with torch.no_grad():
        for _ in range(warmup):
            model(im1, im2, iters=iters, test_mode=True)
        t0 = time.perf_counter()
        for _ in range(steps):
            model(im1, im2, iters=iters, test_mode=True)
        dt = time.perf_counter() - t0
    return a.batch * steps / dt, dt / steps, torch.get_num_threads()


def main_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, sec, threads = cpu_reference_run(a, a.steps, a.warmup, a.iters)
    sample = f"{a.steps} timed + {a.warmup} warm-up forwards of {a.batch} pair(s), {a.height}x{a.width}, {a.iters} iters, fp32, eager ATen on CPU"
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"PriOr-RAFT inference, synthetic {a.height}x{a.width} ERP pair, batch {a.batch}, {a.iters} iters",
                       "device": "cpu", "note": "reference is pure PyTorch and absent on the box: oracle port (oracle/cpu_model.py)"},
            "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- GPU arm
def main_ours(a):
    import torch.distributed as dist
    from prior_flow_b200 import ops
    from prior_flow_b200.model import PriOrRAFT
    from prior_flow_b200 import geometry as geo

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    ops.set_volume_mode(a.volume_mode)
    torch.backends.cudnn.benchmark = True     # let cuDNN pick its conv algorithms (the default picks a CUDA-core SGEMM for the 1x1)
    B, H, W = a.batch, a.height, a.width

    torch.manual_seed(0)
    model = PriOrRAFT(mixed_precision=False, corr_mode=a.corr_mode).to(dev).eval()
    if a.memory_format == "channels_last":
        model = model.to_channels_last()
    host1, host2 = synthetic_pair(B, H, W, 1234 + rank), synthetic_pair(B, H, W, 4321 + rank)
    d1, d2 = host1.to(dev), host2.to(dev)
    host_out = torch.empty(B, 2, H, W).pin_memory()

    def forward(x1, x2):
        with torch.no_grad():
            return model(x1, x2, iters=a.iters, test_mode=True)

    # ---- launch accounting: our kernels per forward (counted by the host layer while running eagerly)
    forward(d1, d2)
    torch.cuda.synchronize()
    ops.reset_launch_count()
    forward(d1, d2)
    torch.cuda.synchronize()
    launches_per_forward = ops.launch_count()

    # ---- CUDA graph of the whole forward (static input buffers)
    graph, static_out = None, None
    if not a.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    forward(d1, d2)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = forward(d1, d2)
            graph.replay()
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001 - report and fall back to eager timing
            print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); timing eager launches", file=sys.stderr)
            graph = None

    def step_resident():
        if graph is not None:
            graph.replay()
            return static_out
        return forward(d1, d2)

    def step_e2e():
        d1.copy_(host1, non_blocking=True)
        d2.copy_(host2, non_blocking=True)
        out = step_resident()
        host_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = torch.tensor([s.elapsed_time(e), wall * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms[0].item(), ms[1].item()

    for _ in range(max(a.warmup, 3)):
        step_resident()
    with ClockSampler(local) as clk:
        dev_ms, _ = timed(step_resident, a.steps)
    for _ in range(max(a.warmup, 3)):
        step_e2e()
    _, e2e_ms = timed(step_e2e, a.steps)
    value = world * B * a.steps / (dev_ms / 1e3)
    e2e_value = world * B * a.steps / (e2e_ms / 1e3)

    if rank == 0:
        hbm, peak_src = peaks()
        # ---- roofline of the dominant hot-path call, timed live on this stream (rank 0's GPU).
        # Method: K calls on K DIFFERENT working sets (4 pyramid pairs = 2.9 GB, 8 coordinate fields; every call reads
        # planes no earlier call touched, so nothing it needs is in the 126 MB L2 except the 64 KB grids, as in the model's
        # loop) captured in ONE CUDA graph; CUDA events bracket a replay; per-call time = replay / K.  No flush kernel in
        # the timed region (a memset flush leaves the L2 full of dirty lines whose write-back the next kernel pays for),
        # no subtraction, graph-launch latency (~4 us) amortised over K calls.
        h, w, N = H // 8, W // 8, (H // 8) * (W // 8)
        g = torch.Generator(device=dev).manual_seed(7)
        grids = model._grids(H, W, dev)
        sets = []
        n_sets = 4 if B * N * N * 4 * 1.33 * 2 <= (1 << 30) else 1   # two pyramids per set; one set is already >> L2 at high resolution
        for i in range(n_sets):
            fm = [torch.randn(B, 256, h, w, device=dev, generator=g) * 1.45 for _ in range(4)]
            sets.append((fm, ops.volume_pyramid(fm[0], fm[1], 4), ops.volume_pyramid(fm[2], fm[3], 4)))
        coords = [geo.coords_grid(B, h, w, dev) + torch.randn(B, 2, h, w, device=dev, generator=g) * 5.0 for _ in range(8)]

        def graph_ms(calls, reps=15):
            """median replay time of one graph holding `calls` (a list of thunks) / len(calls), in ms"""
            for c in calls:
                c()
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                calls[0]()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_):
                keep = [c() for c in calls]   # noqa: F841 - outputs live in the graph's pool
            ts = []
            for _ in range(reps + 2):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                g_.replay()
                e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e))
            ts = sorted(ts[2:])
            return ts[len(ts) // 2] / len(calls)

        def lookup_calls(fuse):
            out = []
            for j in range(8):
                _, pa_, pb_ = sets[j % n_sets]
                own, other = (pa_, pb_) if j < 4 else (pb_, pa_)
                gw, gc = (grids["A2B_W2C_8x"], grids["B2A_8x"]) if j < 4 else (grids["B2A_W2C_8x"], grids["A2B_8x"])
                out.append(lambda c=coords[j], o=own, t=other, gw=gw, gc=gc: ops.lookup(c, o, t, gw, gc, 4, fuse_sum=fuse))
            return out

        look_ms = graph_ms(lookup_calls(False))
        fused_ms = graph_ms(lookup_calls(True))
        vol_ms = graph_ms([lambda f=sets[i][0], k=k: ops.volume_pyramid(f[k], f[k + 1], 4) for i in range(n_sets) for k in (0, 2)])
        fill_buf = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
        fill_gbs = (1 << 30) / graph_ms([lambda: fill_buf.fill_(1)] * 4) / 1e6   # what a write-only stream sustains here
        del fill_buf
        look_bytes = B * (N * 2 * 4 * 100 * 4 + 2 * N * 324 * 4 + 3 * 2 * N * 4)          # 47.45 MB at B=1
        vol_bytes = B * (sum(N * (h >> l) * (w >> l) * 4 for l in range(4)) + 2 * 256 * N * 4)
        achieved = look_bytes / look_ms / 1e6
        roofline = {"kernel": "DCCL lookup call: lookup_rows_kernel + rotate_fwd_kernel (24 calls per pair)", "bound": "hbm",
                    "achieved": round(achieved, 1), "peak": hbm, "unit": "GB/s", "frac": round(achieved / hbm, 4),
                    # dram__bytes_read.sum + dram__bytes_write.sum of lookup_rows_kernel + rotate_fwd_kernel, one launch each, ncu
                    # --set full, cold cache (profiles/r02h_ncu_raw_*.csv; B = 1, 64x128): 56.31 + 2.91 + 9.62 + 0 MB.  Reads are
                    # 2.1x the algorithmic 26.2 MB because a 40-byte footprint row straddles 32-byte sectors; the outputs
                    # (21.2 MB) are still dirty in L2 when the kernels end
                    "traffic": 68.84e6 if (B, h, w) == (1, 64, 128) else None,
                    "peak_source": peak_src, "ms_per_launch": round(look_ms, 4),
                    "algorithmic_bytes_per_launch": look_bytes,
                    # the call as the model issues it (own + other summed, core/prior_raft.py:187): one output tensor
                    "ms_per_launch_fused_sum": round(fused_ms, 4),
                    "timing": "8 calls on 8 different working sets (2.9 GB of pyramids >> L2) in one CUDA graph, CUDA events around a replay, / 8",
                    "other_kernels": {"volume_pyramid(tcgen05, fp32 split) per view": {
                        "ms": round(vol_ms, 4), "GB/s": round(vol_bytes / vol_ms / 1e6, 1), "frac_hbm": round(vol_bytes / vol_ms / 1e6 / hbm, 4),
                        "frac_of_write_only_stream": round(vol_bytes / vol_ms / 1e6 / fill_gbs, 4), "write_only_stream_GB/s": round(fill_gbs, 1),
                        "TFLOP/s_algorithmic": round(2.0 * B * N * N * 256 / vol_ms / 1e9, 1)}}}
        del sets, coords
        torch.cuda.empty_cache()

        cpu_baseline = None
        if not a.skip_cpu_baseline:
            v, sec, threads = cpu_reference_run(a, steps=2, warmup=1, iters=a.iters)
            cpu_baseline = {"value": round(v, 4), "unit": UNIT, "cores": threads, "kind": "port",
                            "sample": f"2 timed + 1 warm-up forwards of {B} pair(s) at {H}x{W}, {a.iters} iters "
                                      f"(eager-ATen restatement of the reference forward, oracle/cpu_model.py), {sec:.2f} s/step"}
        cfg_tag = ("BASELINE configs[1]" if (B, H, W, a.iters) == (1, 512, 1024, 12) else
                   "BASELINE configs[2] per-GPU share" if (H, W, a.iters) == (512, 1024, 12) else
                   "BASELINE configs[3]" if (H, W, a.iters) == (1024, 2048, 32) else "custom")
        line = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
                "ms_per_step": round(dev_ms / a.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"PriOr-RAFT inference, synthetic {H}x{W} ERP pair, batch {B} per GPU, {a.iters} iters ({cfg_tag})",
                           "parallelism": f"pair-per-GPU x{world}, no collectives", "volume_mode": a.volume_mode, "corr_mode": a.corr_mode,
                           "cuda_graph": graph is not None, "weights": "random init (seed 0)", "memory_format": a.memory_format,
                           "l2": "no flush between steps: one step streams ~2.4 GB (2x340 MiB pyramids written, re-read by 24 lookups) >> 126 MB L2",
                           "cudnn_tf32": bool(torch.backends.cudnn.allow_tf32), "cudnn_benchmark": True, "hot_path": "fp32 (tcgen05 fp16x2 split, fp32 accumulate)"},
                "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": 2 * host1.numel() * 4,
                        "d2h_bytes_per_step": host_out.numel() * 4, "ms_per_step": round(e2e_ms / a.steps, 4)},
                "gpu_launches": launches_per_forward * a.steps,
                "gpu_launches_per_step": launches_per_forward,
                "clocks": clk.summary(), "roofline": roofline, "cpu_baseline": cpu_baseline}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        main_reference(a)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU baseline)")
        main_ours(a)


if __name__ == "__main__":
    main()
