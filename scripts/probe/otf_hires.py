"""Times the CUDA-core and the tensor-core on-the-fly lookups at the 1024x2048 shape (128x256 features, BASELINE configs[4])."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import torch_oracle as TO  # noqa: E402
from prior_flow_b200 import ops  # noqa: E402


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 256)
    B, C = 1, 256
    g = torch.Generator(device="cuda").manual_seed(0)
    fm = [torch.randn(B, C, h, w, device="cuda", generator=g) * 1.45 for _ in range(4)]
    Ra, Rb = TO.rotation_matrix([0., 0., -np.pi / 2], device="cuda"), TO.rotation_matrix([0., 0., np.pi / 2], device="cuda")
    gw, gc = TO.generate_samplegrid((B, 3, h, w), Ra.T.contiguous()), TO.generate_samplegrid((B, 3, h, w), Rb)
    cl = lambda t: t.permute(0, 2, 3, 1).contiguous()
    f1a, f2a, f1b, f2b = cl(fm[0]), ops.channels_last_pyramid(fm[1], 4), cl(fm[2]), ops.channels_last_pyramid(fm[3], 4)
    pla, plb = ops.OnTheFlyPlanes(f1a, f2a), ops.OnTheFlyPlanes(f1b, f2b)
    low = torch.randn(B, 2, h // 8, w // 8, device="cuda", generator=g) * 5
    coords = TO.coords_grid(B, h, w, "cuda") + torch.nn.functional.interpolate(low, size=(h, w), mode="bicubic", align_corners=True)
    t_cc = timeit(lambda: ops.lookup_onthefly(coords, f1a, f2a, f1b, f2b, gw, gc, 4))
    t_tc = timeit(lambda: ops.lookup_onthefly(coords, f1a, f2a, f1b, f2b, gw, gc, 4, planes_own=pla, planes_other=plb))
    print(f"[{h}x{w} features, smooth flow] on-the-fly DCCL call: CUDA cores {t_cc:.3f} ms, tensor cores {t_tc:.3f} ms")


if __name__ == "__main__":
    main()
