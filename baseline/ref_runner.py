"""BASELINE INFRASTRUCTURE ONLY — runs the UNMODIFIED reference model (oracle/ref_shim.py finds it: the container's
read-only mount or the vendored copy under baseline/_ref) and times it.  Used by `bench.py` for

  * `--impl reference` and the `cpu_baseline` leg: the reference on the box's host cores (`.cuda()` patched to a no-op,
    SURVEY.md Appendix B — the reference has no CPU switch of its own);
  * `gpu_eager_baseline`: the reference's eager ATen path on the same B200 ("the practical bar"), with the
    per-stage split of SURVEY.md Appendix B (`time_ref`);
  * `dropin`: the same unmodified model object with `prior_flow_b200.install()` active — the hot path on the sm_100a
    kernels, everything else (encoders, update blocks, upsampling, the Python loop) exactly as the reference runs it.

Nothing under prior_flow_b200/ imports this file.
"""
from __future__ import annotations

import os
import sys
import time
from contextlib import contextmanager

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402


def available() -> bool:
    return ref_shim.available()


def synthetic_images(batch, H, W, device, seed=1234):
    g = torch.Generator().manual_seed(seed)
    im1 = torch.rand(batch, 3, H, W, generator=g) * 255
    im2 = torch.rand(batch, 3, H, W, generator=g) * 255
    return im1.to(device), im2.to(device)


class StageTimer:
    """Wraps the hot-path entry points and the four sub-networks of the reference with timers (outermost call wins, so
    `img_rotate` inside `DCCL.__call__` is charged to the lookup).  CUDA: event pairs on the current stream, resolved after a
    synchronize; CPU: perf_counter."""

    def __init__(self, ref, model, cuda: bool):
        self.ref, self.model, self.cuda = ref, model, cuda
        self.depth = 0
        self.pending, self.totals, self.saved = [], {}, []

    def _wrap(self, obj, name, label):
        fn = getattr(obj, name)
        timer = self

        def wrapped(*a, **k):
            if timer.depth:
                return fn(*a, **k)
            timer.depth += 1
            try:
                if timer.cuda:
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record()
                    out = fn(*a, **k)
                    e.record()
                    timer.pending.append((label, s, e))
                else:
                    t0 = time.perf_counter()
                    out = fn(*a, **k)
                    timer.totals[label] = timer.totals.get(label, 0.0) + (time.perf_counter() - t0) * 1e3
                return out
            finally:
                timer.depth -= 1

        self.saved.append((obj, name, fn))
        setattr(obj, name, wrapped)

    def __enter__(self):
        r, m = self.ref, self.model
        PR = r.prior_raft.PriOr_RAFT
        self._wrap(PR, "corr", "volume (PriOr_RAFT.corr)")
        dccl = r.prior_raft.DCCL                                  # the class prior_raft.py resolves (ours after install())
        self._wrap(dccl, "build_pyramid", "pyramid (DCCL.build_pyramid)")
        self._wrap(dccl, "__call__", "lookup (DCCL.__call__)")
        self._wrap(r.ppo, "flo_rotate", "flo_rotate")
        self._wrap(r.ppo, "generate_samplegrid", "generate_samplegrid")
        self._wrap(r.ppo, "img_rotate", "img_rotate (images)")
        self._wrap(r.prior_raft, "cycle_bilinear_sampler", "feature warp (cycle_bilinear_sampler)")
        self._wrap(PR, "groupwise_corr", "groupwise_corr")
        self._wrap(PR, "upsample_flow", "[cuDNN side] upsample_flow")
        for name in ("fnet", "cnet", "ODDC", "update_block"):
            self._wrap(getattr(m, name), "forward", f"[cuDNN side] {name}")
        return self

    def __exit__(self, *exc):
        while self.saved:
            obj, name, fn = self.saved.pop()
            if name == "forward" and isinstance(obj, torch.nn.Module):
                try:
                    delattr(obj, name)                            # instance attribute shadowing the class method
                except AttributeError:
                    setattr(obj, name, fn)
            else:
                setattr(obj, name, fn)

    def collect(self, per: int = 1):
        if self.cuda:
            torch.cuda.synchronize()
            for label, s, e in self.pending:
                self.totals[label] = self.totals.get(label, 0.0) + s.elapsed_time(e)
            self.pending = []
        return {k: round(v / per, 3) for k, v in sorted(self.totals.items())}


@contextmanager
def installed(active: bool):
    if not active:
        yield
        return
    import prior_flow_b200 as pfb
    pfb.install()
    try:
        yield
    finally:
        pfb.uninstall()


def run_reference(device: str, H: int, W: int, batch: int, iters: int, steps: int, warmup: int, install: bool = False,
                  stages: bool = True, graph: bool = False, tf32=None, warmup_iters=None):
    """Times `steps` forwards of the unmodified reference.  Returns a dict with pairs/s, s/step, threads, per-stage ms
    (one extra instrumented forward, so the timers' synchronisation never sits inside the timed region)."""
    cuda = device.startswith("cuda")
    ref = ref_shim.load(cpu=not cuda)
    try:
        if not cuda:
            torch.set_num_threads(os.cpu_count() or 1)
        if tf32 is not None:
            torch.backends.cudnn.allow_tf32 = bool(tf32)
        model = ref_shim.make_model(ref, seed=0).to(device).eval()
        im1, im2 = synthetic_images(batch, H, W, device)
        out = {"kind": "reference" if ref_shim.verified() else "reference (unverified copy)", "device": device,
               "threads": torch.get_num_threads(), "host_cores": os.cpu_count()}

        def fwd(n_iters=iters):
            with torch.no_grad():
                return model(im1, im2, iters=n_iters, test_mode=True)

        def sync():
            if cuda:
                torch.cuda.synchronize()

        with installed(install):
            for _ in range(warmup):
                fwd(warmup_iters or iters)
            sync()
            step = fwd
            if graph and cuda:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    fwd()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    static = fwd()  # noqa: F841
                g.replay()
                sync()
                step = g.replay
            if cuda:
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for _ in range(steps):
                    step()
                e.record()
                sync()
                sec = s.elapsed_time(e) / 1e3 / steps
            else:
                t0 = time.perf_counter()
                for _ in range(steps):
                    step()
                sec = (time.perf_counter() - t0) / steps
            out.update({"value": round(batch / sec, 4), "s_per_step": round(sec, 5), "steps": steps, "warmup": warmup})
            if stages:
                # three instrumented forwards, per-stage median: one forward alone can catch an allocator hiccup (a cudaMalloc
                # inside build_pyramid once read 3.0 ms instead of 0.26)
                runs, t0 = [], time.perf_counter()
                for _ in range(3):
                    with StageTimer(ref, model, cuda) as t:
                        fwd()
                        runs.append(t.collect())
                st = {k: round(sorted(r.get(k, 0.0) for r in runs)[1], 3) for k in runs[0]}
                out["stages_ms"] = st
                out["stages_note"] = ("median of three instrumented eager forwards; CUDA events per call, outermost call wins"
                                      if cuda else "median of three instrumented forwards; perf_counter per call, outermost call wins")
                hot = sum(v for k, v in st.items() if not k.startswith("[cuDNN side]"))
                out["hot_path_ms"] = round(hot, 3)
                out["instrumented_forward_ms"] = round((time.perf_counter() - t0) * 1e3 / 3, 2)
        return out
    finally:
        ref_shim.unpatch_cuda()
