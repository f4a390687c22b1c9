"""Peak / live device memory of a training forward + backward at 1024x2048 per correlation mode, phase by phase."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from prior_flow_b200.model import PriOrRAFT  # noqa: E402
from prior_flow_b200.train import sequence_loss  # noqa: E402

GB = 2 ** 30
H, W, iters = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1024, 2048, 12)
for mode in ("materialized", "onthefly"):
    torch.manual_seed(0)
    model = PriOrRAFT(corr_mode=mode).cuda().train()
    model.freeze_bn()
    g = torch.Generator(device="cuda").manual_seed(1)
    im1, im2 = torch.rand(1, 3, H, W, device="cuda", generator=g) * 255, torch.rand(1, 3, H, W, device="cuda", generator=g) * 255
    for rep in range(2):
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        pa, pb = model(im1, im2, iters=iters)
        torch.cuda.synchronize()
        fwd_live, fwd_peak = torch.cuda.memory_allocated(), torch.cuda.max_memory_allocated()
        loss = sum(p.abs().mean() for p in pa) + sum(p.abs().mean() for p in pb)
        torch.cuda.reset_peak_memory_stats()
        loss.backward()
        torch.cuda.synchronize()
        bwd_peak = torch.cuda.max_memory_allocated()
        model.zero_grad(set_to_none=True)
        del pa, pb, loss
        if rep:
            print(f"[{mode}] {H}x{W}/{iters}: before {base / GB:.2f} GB, live after forward {fwd_live / GB:.2f}, peak in forward {fwd_peak / GB:.2f}, "
                  f"peak in backward {bwd_peak / GB:.2f}")
    del model
    torch.cuda.empty_cache()
