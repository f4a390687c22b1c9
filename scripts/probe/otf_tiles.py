"""Which query tiles of the tensor-core on-the-fly lookup take the tcgen05 path (tile box fits the level's local plane), per view
and level, for a smooth flow, for i.i.d. noise and for the flow a random-init model predicts.  Prints the fractions."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import torch_oracle as TO  # noqa: E402
from prior_flow_b200 import ops  # noqa: E402


def report(tag, coords, f1a, f2a, f1b, f2b, gw, gc, pla, plb):
    ops.lookup_onthefly(coords, f1a, f2a, f1b, f2b, gw, gc, 4, planes_own=pla, planes_other=plb)
    work, T, views, L, B = ops._state["otf_work"]
    work = work.cpu().long()
    ctr = work[:16]
    lo, hi = work[16:16 + 4 * T].view(views, L, B, -1, 4), work[16 + 4 * T:16 + 8 * T].view(views, L, B, -1, 4)
    alloc = work[16 + 8 * T:16 + 9 * T].view(views, L, B, -1)
    bw = torch.minimum(hi[..., 0] - lo[..., 0], hi[..., 1] - lo[..., 1]) + 1      # the narrower of the two column numberings
    bh = hi[..., 2] - lo[..., 2] + 1
    print(f"[{tag}] tiles on the CUDA-core path {int(ctr[0])}, work items of the dots kernel {int(ctr[2])}, pool segments asked for {int(ctr[3])}")
    for v in range(views):
        for l in range(L):
            w_, h_, a_ = bw[v, l].flatten(), bh[v, l].flatten(), alloc[v, l].flatten()
            valid = a_ != -2
            print(f"[{tag}] view {v} level {l}: tiles {w_.numel()}, with taps {int(valid.sum())}, of those on tensor cores "
                  f"{float((a_ >= 0).sum()) / max(int(valid.sum()), 1):.2f}; median box {int(w_[valid].median())} x {int(h_[valid].median())}, "
                  f"p90 {int(w_[valid].float().quantile(0.9))} x {int(h_[valid].float().quantile(0.9))}, max {int(w_[valid].max())} x {int(h_[valid].max())}")


def main():
    B, h, w, C = 1, 64, 128, 256
    g = torch.Generator(device="cuda").manual_seed(0)
    fm = [torch.randn(B, C, h, w, device="cuda", generator=g) * 1.45 for _ in range(4)]
    Ra, Rb = TO.rotation_matrix([0., 0., -np.pi / 2], device="cuda"), TO.rotation_matrix([0., 0., np.pi / 2], device="cuda")
    gw, gc = TO.generate_samplegrid((B, 3, h, w), Ra.T.contiguous()), TO.generate_samplegrid((B, 3, h, w), Rb)
    cl = lambda t: t.permute(0, 2, 3, 1).contiguous()
    f1a, f2a, f1b, f2b = cl(fm[0]), ops.channels_last_pyramid(fm[1], 4), cl(fm[2]), ops.channels_last_pyramid(fm[3], 4)
    pla, plb = ops.OnTheFlyPlanes(f1a, f2a), ops.OnTheFlyPlanes(f1b, f2b)
    grid = TO.coords_grid(B, h, w, "cuda")
    report("zero flow", grid, f1a, f2a, f1b, f2b, gw, gc, pla, plb)
    low = torch.randn(B, 2, h // 8, w // 8, device="cuda", generator=g) * 5
    report("smooth 5px/8", grid + torch.nn.functional.interpolate(low, size=(h, w), mode="bicubic", align_corners=True), f1a, f2a, f1b, f2b, gw, gc, pla, plb)
    low = torch.randn(B, 2, h // 16, w // 16, device="cuda", generator=g) * 3
    report("smooth 3px/16", grid + torch.nn.functional.interpolate(low, size=(h, w), mode="bicubic", align_corners=True), f1a, f2a, f1b, f2b, gw, gc, pla, plb)


if __name__ == "__main__":
    main()
