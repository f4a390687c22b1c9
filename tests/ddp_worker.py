"""Worker for tests/test_gpu_ddp.py (torchrun, NCCL, one rank per GPU): one training step of the model at a reduced size with
the pair-per-GPU sharding of BASELINE config 5; rank 0 saves the averaged gradients for comparison with a 1-GPU step over the
whole batch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from prior_flow_b200 import distributed as pfd  # noqa: E402
from prior_flow_b200.model import PriOrRAFT  # noqa: E402
from prior_flow_b200.train import latitude_weights, sequence_loss  # noqa: E402


def batch(global_batch, H, W, device):
    g = torch.Generator().manual_seed(77)
    im1 = torch.rand(global_batch, 3, H, W, generator=g) * 255
    im2 = torch.rand(global_batch, 3, H, W, generator=g) * 255
    gt = torch.randn(global_batch, 2, H, W, generator=g) * 3
    return im1.to(device), im2.to(device), gt.to(device)


def grads_of_step(model, im1, im2, gt, scale):
    model.zero_grad(set_to_none=True)
    pa, pb = model(im1, im2, iters=3)
    H, W = im1.shape[-2:]
    ok, lat = torch.ones(im1.shape[0], H, W, device=im1.device), latitude_weights(H, W, im1.device)
    # both views' predictions enter the loss, as in train_step: every parameter takes part (DDP without find_unused_parameters
    # never reduces a bucket that holds an unused parameter)
    loss = sequence_loss(pa, gt, ok, lat)[0] + sequence_loss(pb, gt, ok, lat)[0]
    (loss * scale).backward()
    return float(loss.detach())


def main():
    out_path = sys.argv[1]
    ctx = pfd.init_from_env("nccl")
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    model = PriOrRAFT().to(ctx.device)
    model.train()
    model.freeze_bn()
    manual = os.environ.get("PF_TEST_MANUAL_ALLREDUCE") == "1"      # debugging aid: no DDP wrapper, gradients summed by hand
    ddp = model if manual else pfd.wrap_ddp(model, ctx)
    G, H, W = int(os.environ.get("PF_TEST_WORLD_BATCH", 2 * ctx.world)), 128, 256
    im1, im2, gt = batch(G, H, W, ctx.device)
    mine = pfd.shard_pairs(G, ctx.rank, ctx.world)
    sl = slice(mine[0], mine[-1] + 1)
    loss = grads_of_step(ddp, im1[sl], im2[sl], gt[sl], 1.0 if manual else pfd.ddp_loss_scale(ctx))
    if manual and ctx.world > 1:
        for p in model.parameters():
            if p.grad is not None:
                torch.distributed.all_reduce(p.grad)
    total = torch.tensor([loss], device=ctx.device, dtype=torch.float64)
    if ctx.world > 1:
        torch.distributed.all_reduce(total)
    if ctx.rank == 0:
        torch.save({"loss": float(total), "grads": {k: p.grad.detach().cpu() for k, p in model.named_parameters() if p.grad is not None}}, out_path)
    if ctx.world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
