#!/usr/bin/env python
"""Headline benchmark: PriOr-RAFT inference pairs/s at 512x1024 ERP, 12 GRU iterations (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One JSON line on stdout (rank 0).  A "step" is one pass of the model (hot path on the sm_100a kernels, encoders and
GRU blocks on cuDNN) over one batch of synthetic image pairs; pairs shard one process per GPU with no data-path
collective ("scaling": "weak").  `value` is device-resident throughput (inputs in HBM, whole forward replayed from a
CUDA graph), `e2e` goes through the public API with pinned-host inputs and a device->host read of the flow inside the
timed region.  `roofline` is the dominant hot-path kernel pair (one DCCL lookup call = lookup_rows_kernel + rotate_fwd_kernel),
timed live with CUDA events, with the DRAM traffic of the same two kernels measured by an ncu pass started from here and the
HBM floor of the lookup's access pattern (pf_probe_gather) timed beside it.  At N = 1 the line also carries
  `cpu_baseline`        the UNMODIFIED reference (baseline/_ref, sha256-checked) on the host cores, bounded sample;
  `gpu_eager_baseline`  the unmodified reference's eager ATen path on the same B200, per-stage CUDA-event split;
  `dropin`              the same unmodified model object with prior_flow_b200.install() (eager and CUDA-graph replay).
`--impl reference` times the unmodified reference on the host cores alone (oracle port only if the copy is missing).
`--global-batch G` shards G pairs over the ranks (BASELINE configs[2]: `"scaling": "strong"`).
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
    ap.add_argument("--cudnn-tf32", default="on", choices=["on", "off"],
                    help="TF32 convolutions on the cuDNN side (PyTorch's and hence the reference's default on a GPU); off = fp32 "
                         "convolutions, the setting the 1e-3 px flow gate is stated for (tests/test_gpu_dropin.py)")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-gpu-baselines", action="store_true", help="skip gpu_eager_baseline / dropin (reference on the same GPU)")
    ap.add_argument("--skip-traffic", action="store_true", help="skip the ncu pass that measures roofline.traffic")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="total pairs per step, sharded pair-per-GPU over the ranks (strong scaling; BASELINE configs[2] = 64)")
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
    """The reference forward on the host cores.  Preferred: the UNMODIFIED reference (baseline/_ref copy checked against
    baseline/ref_manifest.json, `.cuda()` patched to a no-op — SURVEY.md Appendix B); fallback when the copy is missing:
    the oracle's eager-ATen restatement (oracle/cpu_model.py).  Returns a dict (value, s_per_step, threads, kind, ...)."""
    try:
        from baseline import ref_runner
        if ref_runner.available():
            r = ref_runner.run_reference("cpu", a.height, a.width, a.batch, iters, steps, warmup, stages=True, warmup_iters=2)
            r["kind"] = "reference" if r["kind"] == "reference" else "port"
            return r
    except Exception as e:  # noqa: BLE001
        print(f"[bench] unmodified reference unavailable ({type(e).__name__}: {e}); timing the oracle port", file=sys.stderr)
    from oracle.cpu_model import EagerPriOrRAFT
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    model = EagerPriOrRAFT().eval()
    g = torch.Generator().manual_seed(1234)
    im1 = torch.rand(a.batch, 3, a.height, a.width, generator=g) * 255
    im2 = torch.rand(a.batch, 3, a.height, a.width, generator=g) * 255
    with torch.no_grad():
        for _ in range(warmup):
            model(im1, im2, iters=2, test_mode=True)
        t0 = time.perf_counter()
        for _ in range(steps):
            model(im1, im2, iters=iters, test_mode=True)
        dt = time.perf_counter() - t0
    return {"value": round(a.batch * steps / dt, 4), "s_per_step": round(dt / steps, 4), "threads": torch.get_num_threads(),
            "host_cores": os.cpu_count(), "kind": "port", "steps": steps, "warmup": warmup}


def cpu_baseline_object(a, r):
    what = ("unmodified reference (baseline/_ref, sha256-verified), Tensor.cuda patched to a no-op" if r["kind"] == "reference"
            else "oracle port of the reference forward (oracle/cpu_model.py)")
    out = {"value": r["value"], "unit": UNIT, "cores": r["threads"], "host_cores": r.get("host_cores"), "kind": r["kind"],
           "s_per_step": r["s_per_step"],
           "sample": f"{r['steps']} timed forwards (+ {r['warmup']} warm-up at 2 iters) of {a.batch} pair(s) at {a.height}x{a.width}, "
                     f"{a.iters} iters, fp32, {what}"}
    if "stages_ms" in r:
        out["stages_ms"] = r["stages_ms"]
    return out


def main_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(a, a.steps, max(a.warmup, 1), a.iters)
    cb = cpu_baseline_object(a, r)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": round(r["s_per_step"] * 1e3, 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"PriOr-RAFT inference, synthetic {a.height}x{a.width} ERP pair, batch {a.batch}, {a.iters} iters",
                       "device": "cpu", "note": "the reference's own forward on the host cores (it has no GPU-free switch: .cuda() is patched to a no-op)"},
            "cpu_baseline": cb,
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def measure_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one lookup_rows_kernel + one rotate_fwd_kernel launch (B = 1, 64x128),
    from an ncu pass over scripts/kbench.py started here.  None when ncu is unavailable."""
    import csv
    import io
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--print-units", "base",
           "-k", "regex:lookup_rows_kernel|rotate_fwd_kernel", "-s", "2", "-c", "2", "--csv",
           sys.executable, os.path.join(ROOT, "scripts", "kbench.py"), "--iters", "1", "--skip-torch", "--only", "lookup_dual"]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT).stdout
        lines = out.splitlines()
        start = next(i for i, ln in enumerate(lines) if ln.startswith('"ID"'))
        rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
        per = {}
        for r in rows:
            k = r["Kernel Name"].split("(")[0].split("<")[0].split("::")[-1]
            per.setdefault(k, {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        total = sum(sum(v.values()) for v in per.values())
        if len(per) < 2 or total <= 0:
            return None, f"ncu output not understood ({len(rows)} rows)"
        return total, {k: {m: round(x / 1e6, 2) for m, x in v.items()} for k, v in per.items()}
    except Exception as e:  # noqa: BLE001
        return None, f"ncu pass failed: {type(e).__name__}: {e}"


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
    torch.backends.cudnn.allow_tf32 = a.cudnn_tf32 == "on"
    torch.backends.cudnn.benchmark = True     # let cuDNN pick its conv algorithms (the default picks a CUDA-core SGEMM for the 1x1)
    strong = a.global_batch > 0
    if strong:
        if a.global_batch % world:
            raise SystemExit(f"--global-batch {a.global_batch} is not divisible by {world} ranks")
        a.batch = a.global_batch // world
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

    # ---- CUDA-graph inference through the product API (prior_flow_b200.model.GraphedForward)
    run = None
    if not a.no_graph:
        try:
            run = model.graphed(a.iters)
            run(d1, d2)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001 - report and fall back to eager timing
            print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); timing eager launches", file=sys.stderr)
            run = None

    def step_resident():                      # inputs already in HBM (the graph's static buffers are refreshed device-to-device)
        return run(d1, d2) if run is not None else forward(d1, d2)

    def step_e2e():                           # the call a user makes: pinned host images in, flow back on the host
        out = run(host1, host2) if run is not None else forward(host1.to(dev, non_blocking=True), host2.to(dev, non_blocking=True))
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
    # the same forward through the eager product API (`model(image1, image2, ...)`, no CUDA graph): launch-bound, reported beside it
    eager_steps = max(3, min(a.steps, 10))
    for _ in range(2):
        forward(d1, d2)
    eager_ms, _ = timed(lambda: forward(d1, d2), eager_steps)
    eager_value = world * B * eager_steps / (eager_ms / 1e3)

    # The timed, collective part is over: tear the process group down NOW, so that nothing below (rank 0's roofline section)
    # runs while other ranks spin in an NCCL barrier.  Ranks != 0 are done.
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    if run is not None:
        run.reset()
    torch.cuda.empty_cache()

    hbm, peak_src = peaks()
    # ---- roofline of the dominant hot-path call, timed live on this stream (rank 0's GPU).
    # Method: K calls on K DIFFERENT working sets (4 pyramid pairs = 2.9 GB, 8 coordinate fields; every call reads
    # planes no earlier call touched, so nothing it needs is in the 126 MB L2 except the 64 KB grids, as in the model's
    # loop) captured in ONE CUDA graph; CUDA events bracket a replay; per-call time = replay / K.  No flush kernel in
    # the timed region (a memset flush leaves the L2 full of dirty lines whose write-back the next kernel pays for),
    # no subtraction, graph-launch latency (~4 us) amortised over K calls.
    Br = 1                                   # roofline shapes are per pair (B = 1), whatever the step's batch
    h, w, N = H // 8, W // 8, (H // 8) * (W // 8)
    g = torch.Generator(device=dev).manual_seed(7)
    grids = model._grids(H, W, dev)
    onthefly = a.corr_mode == "onthefly"
    sets = []
    n_sets = 4 if Br * N * N * 4 * 1.33 * 2 <= (1 << 30) else 1   # two pyramids per set; one set is already >> L2 at high resolution
    for i in range(n_sets):
        fm = [torch.randn(Br, 256, h, w, device=dev, generator=g) * 1.45 for _ in range(4)]
        sets.append((fm, ops.volume_pyramid(fm[0], fm[1], 4), ops.volume_pyramid(fm[2], fm[3], 4)))
    coords = [geo.coords_grid(Br, h, w, dev) + torch.randn(Br, 2, h, w, device=dev, generator=g) * 5.0 for _ in range(8)]

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
    fbuf = fill_buf.view(torch.float32)
    read_gbs = (1 << 30) / graph_ms([lambda: ops.probe_stream_read(fbuf)] * 4) / 1e6   # ... and a read-only stream
    del fill_buf, fbuf
    # HBM floor of the lookup's access pattern: the level-0 footprint loads of BOTH views of a call (2 N planes, 10x10 footprint
    # each, random corners) and nothing else, on planes no earlier probe call touched
    pos = torch.stack([torch.randint(0, w - 10, (2 * Br * N,), device=dev, generator=g),
                       torch.randint(0, h - 10, (2 * Br * N,), device=dev, generator=g)], dim=1).to(torch.int32)
    probe_vols = [torch.cat([s_[1][0].view(Br * N, h, w), s_[2][0].view(Br * N, h, w)]) for s_ in sets]
    gather_ms = graph_ms([lambda v=v: ops.probe_gather(v, pos) for v in probe_vols] * 2)
    del probe_vols
    look_bytes = Br * (N * 2 * 4 * 100 * 4 + 2 * N * 324 * 4 + 3 * 2 * N * 4)          # 47.45 MB at B=1
    vol_bytes = Br * (sum(N * (h >> l) * (w >> l) * 4 for l in range(4)) + 2 * 256 * N * 4)
    achieved = look_bytes / look_ms / 1e6
    traffic, traffic_detail = (None, "skipped")
    if world == 1 and not a.skip_traffic and (h, w) == (64, 128):
        traffic, traffic_detail = measure_traffic()
    roofline = {"kernel": "DCCL lookup call: lookup_rows_kernel + rotate_fwd_kernel (24 calls per pair)", "bound": "hbm",
                "achieved": round(achieved, 1), "peak": hbm, "unit": "GB/s", "frac": round(achieved / hbm, 4),
                # dram__bytes_read.sum + dram__bytes_write.sum of one lookup_rows_kernel + one rotate_fwd_kernel launch, measured by
                # an ncu pass over scripts/kbench.py started from this run (B = 1, 64x128, cold caches)
                "traffic": traffic, "traffic_MB_per_kernel": traffic_detail,
                "peak_source": peak_src, "ms_per_launch": round(look_ms, 4),
                "algorithmic_bytes_per_launch": look_bytes,
                # the call as the model issues it (own + other summed, core/prior_raft.py:187): one output tensor
                "ms_per_launch_fused_sum": round(fused_ms, 4),
                "timing": "8 calls on 8 different working sets (2.9 GB of pyramids >> L2) in one CUDA graph, CUDA events around a replay, / 8",
                # what the HBM system gives this ACCESS PATTERN: the level-0 footprint loads of a call alone (no arithmetic, no
                # stores): 2N planes x ten 40-byte segments, each in its own DRAM page.  Level 0 is 2 of the 8 (view, level)
                # gathers of a call and 6.55 of its 26.2 MB of footprint bytes
                "access_pattern_floor": {"what": "pf_probe_gather: level-0 footprint loads of both views only (6.55 MB algorithmic)",
                                         "ms": round(gather_ms, 4), "algorithmic_GB/s": round(2 * Br * N * 400 / gather_ms / 1e6, 1),
                                         "read_only_stream_GB/s": round(read_gbs, 1)},
                "other_kernels": {"volume_pyramid(tcgen05, fp32 split) per view": {
                    "ms": round(vol_ms, 4), "GB/s": round(vol_bytes / vol_ms / 1e6, 1), "frac_hbm": round(vol_bytes / vol_ms / 1e6 / hbm, 4),
                    "frac_of_write_only_stream": round(vol_bytes / vol_ms / 1e6 / fill_gbs, 4), "write_only_stream_GB/s": round(fill_gbs, 1),
                    "TFLOP/s_algorithmic": round(2.0 * Br * N * N * 256 / vol_ms / 1e9, 1)}}}
    if onthefly:
        cl = lambda t: t.permute(0, 2, 3, 1).contiguous()
        fm = sets[0][0]
        f1a, f2a, f1b, f2b = cl(fm[0]), ops.channels_last_pyramid(fm[1], 4), cl(fm[2]), ops.channels_last_pyramid(fm[3], 4)
        use_tc = os.environ.get("PF_ONTHEFLY_TC", "1") != "0" and ops.OnTheFlyPlanes.supported(f1a)
        pla, plb = (ops.OnTheFlyPlanes(f1a, f2a), ops.OnTheFlyPlanes(f1b, f2b)) if use_tc else (None, None)
        otf_ms = graph_ms([lambda c=c: ops.lookup_onthefly(c, f1a, f2a, f1b, f2b, grids["A2B_W2C_8x"], grids["B2A_8x"], 4,
                                                           planes_own=pla, planes_other=plb) for c in coords[:4]])
        cc_ms = graph_ms([lambda c=c: ops.lookup_onthefly(c, f1a, f2a, f1b, f2b, grids["A2B_W2C_8x"], grids["B2A_8x"], 4) for c in coords[:4]])
        flops = 2.0 * 2 * 4 * Br * N * 100 * 256          # (2r+2)^2 = 100 lattice dot products of C = 256 per query, level and view
        tf_peak = 1652.6
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                tf_peak = float(json.load(fh)["bf16_tflops"])
        except Exception:
            pass
        # the step runs the on-the-fly lookup: THAT is the dominant kernel sequence of this configuration
        roofline = {"kernel": ("on-the-fly DCCL lookup call (pf_lookup_onthefly_tc: tile boxes, tcgen05 dots into local planes, blend, rotate)" if use_tc
                               else "on-the-fly DCCL lookup call (pf_lookup_onthefly + rotate, CUDA cores)"),
                    "bound": "tensor", "achieved": round(flops / otf_ms / 1e9, 2), "peak": tf_peak, "unit": "TFLOP/s",
                    "frac": round(flops / otf_ms / 1e9 / tf_peak, 5), "traffic": None, "ms_per_launch": round(otf_ms, 4),
                    "cuda_core_kernel_ms_per_launch": round(cc_ms, 4),
                    "algorithmic_flops_per_launch": flops,
                    "peak_source": "MEASURED_PEAKS.json bf16 burst; the algorithmic count is one product per lattice point, the kernel spends three fp16 "
                                   "products on the tile's whole bounding box to keep fp32 accuracy",
                    "materialized_lookup_for_comparison": roofline}
    del sets, coords
    torch.cuda.empty_cache()

    cpu_baseline = gpu_eager = dropin = None
    if world == 1 and not a.skip_gpu_baselines:
        try:
            from baseline import ref_runner
            if ref_runner.available():
                tf32 = bool(torch.backends.cudnn.allow_tf32)
                torch.backends.cudnn.benchmark = False       # the reference as it runs out of the box
                gpu_eager = ref_runner.run_reference(f"cuda:{local}", H, W, B, a.iters, steps=5, warmup=2)
                gpu_eager.update({"unit": UNIT, "what": "unmodified reference, eager ATen on this GPU (its default: cuDNN TF32 "
                                  f"{'on' if tf32 else 'off'}, fp32 matmul), CUDA events around 5 forwards"})
                d_eager = ref_runner.run_reference(f"cuda:{local}", H, W, B, a.iters, steps=5, warmup=2, install=True)
                d_graph = ref_runner.run_reference(f"cuda:{local}", H, W, B, a.iters, steps=10, warmup=2, install=True, stages=False, graph=True)
                dropin = {"unit": UNIT, "what": "the same unmodified reference model with prior_flow_b200.install(): hot path on the sm_100a "
                          "kernels, everything else (encoders, update blocks, 12 convex upsamplings, NCHW, Python loop) as the reference runs it",
                          "eager": {k: d_eager[k] for k in ("value", "s_per_step", "stages_ms", "hot_path_ms") if k in d_eager},
                          "cuda_graph": {k: d_graph[k] for k in ("value", "s_per_step")},
                          "kind": d_eager["kind"]}
                torch.backends.cudnn.benchmark = True
            else:
                gpu_eager = {"unavailable": "baseline/_ref is missing (run scripts/vendor_reference.py in the build container)"}
        except Exception as e:  # noqa: BLE001
            gpu_eager = gpu_eager or {"unavailable": f"{type(e).__name__}: {e}"}
            print(f"[bench] GPU reference legs failed: {type(e).__name__}: {e}", file=sys.stderr)
    if world == 1 and not a.skip_cpu_baseline:
        cpu_baseline = cpu_baseline_object(a, cpu_reference_run(a, steps=2, warmup=1, iters=a.iters))
    elif world > 1:
        cpu_baseline = {"skipped": "timed at N = 1 only (rank 0 of a multi-rank run shares the host with the other ranks' processes)"}
    cfg_tag = ("BASELINE configs[1]" if (B, H, W, a.iters) == (1, 512, 1024, 12) else
               f"BASELINE configs[2]: global batch {a.global_batch} sharded pair-per-GPU" if strong and (H, W, a.iters) == (512, 1024, 12) else
               "BASELINE configs[2] per-GPU share" if (H, W, a.iters) == (512, 1024, 12) else
               "BASELINE configs[3]" if (H, W, a.iters) == (1024, 2048, 32) else "custom")
    line = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": round(dev_ms / a.steps, 4), "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"PriOr-RAFT inference, synthetic {H}x{W} ERP pair, batch {B} per GPU, {a.iters} iters ({cfg_tag})",
                       "parallelism": f"pair-per-GPU x{world}, no collectives", "volume_mode": a.volume_mode, "corr_mode": a.corr_mode,
                       "global_batch": world * B,
                       "cuda_graph": run is not None, "api": "PriOrRAFT.graphed(iters)(image1, image2)" if run is not None else "PriOrRAFT.forward", "weights": "random init (seed 0)", "memory_format": a.memory_format,
                       "l2": "no flush between steps: one step streams ~2.4 GB (2x340 MiB pyramids written, re-read by 24 lookups) >> 126 MB L2",
                       "cudnn_tf32": bool(torch.backends.cudnn.allow_tf32), "cudnn_benchmark": True,
                       "flow_vs_reference": "mean EPE vs the unmodified reference at this shape (tests/test_gpu_dropin.py): 9.8e-6 px with fp32 "
                                            "convolutions; 3.4e-3 px with TF32 convolutions, where the reference's own TF32 run is 6.1e-3 px from its fp32 run", "hot_path": "fp32 (tcgen05 fp16x2 split, fp32 accumulate)"},
            "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": 2 * host1.numel() * 4,
                    "d2h_bytes_per_step": host_out.numel() * 4, "ms_per_step": round(e2e_ms / a.steps, 4)},
            "product_api_eager": {"value": round(eager_value, 3), "unit": "pairs/s", "ms_per_step": round(eager_ms / eager_steps, 4), "steps": eager_steps,
                                  "what": "PriOrRAFT.forward without a CUDA graph (inputs resident), device-timed like `value`"},
            "gpu_launches": launches_per_forward * a.steps,
            "gpu_launches_per_step": launches_per_forward,
            "clocks": clk.summary(), "roofline": roofline, "cpu_baseline": cpu_baseline,
            "gpu_eager_baseline": gpu_eager, "dropin": dropin}
    print(json.dumps(line), flush=True)


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
