"""Worker for test_two_rank_gloo_plumbing (launched by torchrun, backend gloo, CPU)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prior_flow_b200 import distributed as pfd  # noqa: E402


def main():
    ctx = pfd.init_from_env(backend="gloo")
    assert ctx.world == 2 and ctx.rank in (0, 1)
    mine = pfd.shard_pairs(7, ctx.rank, ctx.world)
    counts = torch.tensor([len(mine)], dtype=torch.int64)
    dist.all_reduce(counts)
    assert counts.item() == 7
    # step time is the max over ranks (bench.py contract)
    t = pfd.max_over_ranks(10.0 + 5.0 * ctx.rank, ctx)
    assert abs(t - 15.0) < 1e-9
    # DDP averages gradients; the reference's DataParallel sums a sum-reduced loss (train_flow.py:69): compensate
    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 1, bias=False)
    ddp = torch.nn.parallel.DistributedDataParallel(lin)
    x = torch.arange(8, dtype=torch.float32).view(2, 4) + ctx.rank
    loss = ddp(x).sum() * pfd.ddp_loss_scale(ctx)
    loss.backward()
    full = torch.cat([torch.arange(8, dtype=torch.float32).view(2, 4) + r for r in range(2)]).sum(0)
    assert torch.allclose(lin.weight.grad[0], full), (lin.weight.grad, full)
    dist.barrier()
    print(f"rank{ctx.rank} ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
