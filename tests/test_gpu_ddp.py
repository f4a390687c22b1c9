"""GPU, >= 2 devices: a 2-GPU DDP training step equals a 1-GPU step with the same global batch — summed loss
identical, gradients after the NCCL all-reduce (x world, the DataParallel-compatible scale of distributed.ddp_loss_scale) equal
to the single-process gradients up to atomics / reduction order.  Skipped on a 1-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_gpu_ddp_step_equals_one_gpu_step_with_the_same_global_batch(tmp_path):
    worker = os.path.join(ROOT, "tests", "ddp_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PYTHONPATH=ROOT)
    two, one = str(tmp_path / "two.pt"), str(tmp_path / "one.pt")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", worker, two], env=env, capture_output=True, text=True, timeout=400)
    assert r.returncode == 0, r.stderr[-3000:]
    # the same 4 pairs on one GPU: the worker's world-1 path uses G = 2, so run it with the 2-rank batch size through an env override
    env1 = dict(env, PF_TEST_WORLD_BATCH="4")
    r = subprocess.run([sys.executable, worker, one], env=env1, capture_output=True, text=True, timeout=400)
    assert r.returncode == 0, r.stderr[-3000:]
    a, b = torch.load(two), torch.load(one)
    assert abs(a["loss"] - b["loss"]) <= 1e-4 * abs(b["loss"])
    # per parameter: max |difference| over max |reference|, with a floor of 1e-4 x the largest gradient of the whole model —
    # conv biases in front of an InstanceNorm have an analytically ZERO gradient, what is left of them is rounding noise
    gmax = max(float(g.abs().max()) for g in b["grads"].values())
    worst, worst_name = 0.0, ""
    for k, gb in b["grads"].items():
        ga = a["grads"][k]
        e = float((ga - gb).abs().max()) / max(float(gb.abs().max()), 1e-4 * gmax)
        if e > worst:
            worst, worst_name = e, k
    errs = sorted(((float((a["grads"][k] - gb).abs().max()), float(gb.abs().max()), k) for k, gb in b["grads"].items()), reverse=True)[:4]
    print("\n[ddp] largest absolute differences (|diff|, |ref|max, name):", errs, "gmax", gmax)
    print(f"\n[ddp] summed loss {a['loss']:.6f} vs {b['loss']:.6f}; worst relative gradient difference {worst:.2e} ({worst_name})")
    assert set(a["grads"]) == set(b["grads"])
    assert worst < 1e-2        # cuDNN picks other kernels for batch 2 and batch 4: 5.7e-3 between a batch of 4 and its two halves on ONE GPU
