"""GPU box: where does pf_samplegrid differ from the reference's generate_samplegrid on CUDA?  Stage-by-stage comparison of
the reference's intermediates with elementwise restatements, and candidate accumulation orders for its 3x3 matmul."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from prior_flow_b200 import ops  # noqa: E402

ref = ref_shim.load()
P = ref.ppo
torch.backends.cuda.matmul.allow_tf32 = False


def fma(a, b, c):          # fp32 fma emulated in fp64 (product exact, one extra rounding of the sum: rare double rounding)
    return (a.double() * b.double() + c.double()).float()


for (H, W) in ((64, 128), (512, 1024)):
    for ang in (-np.pi / 2, np.pi / 2):
        for transpose in (False, True):
            R = P.generate_rotation_metrix(theta_list=[0., 0., ang])
            R = R.T if transpose else R
            size = (1, 3, H, W)
            plane = P.generate_plane_grid(size)
            sph = P.ERP.plane2spherical(plane)
            cart = P.Spherical2Cartesian(sph)
            rot = P.rotate_cartesian(cart, R)
            sph2 = P.Cartesian2Spherical(rot)
            want = P.ERP.spherical2plane(sph2)
            x, y, z = cart[:, 0], cart[:, 1], cart[:, 2]
            Rc = R.contiguous()
            r = [[Rc[i, j] for j in range(3)] for i in range(3)]
            cands = {
                "fma k asc  fma(r2,z,fma(r1,y,r0*x))": lambda i: fma(r[i][2], z, fma(r[i][1], y, r[i][0] * x)),
                "fma k desc fma(r0,x,fma(r1,y,r2*z))": lambda i: fma(r[i][0], x, fma(r[i][1], y, r[i][2] * z)),
                "no fma (r0*x + r1*y) + r2*z": lambda i: (r[i][0] * x + r[i][1] * y) + r[i][2] * z,
                "fp64 dot rounded once": lambda i: (r[i][0].double() * x.double() + r[i][1].double() * y.double() + r[i][2].double() * z.double()).float(),
            }
            msg = []
            for name, f in cands.items():
                bad = sum(int((f(i) != rot[:, i]).sum()) for i in range(3))
                msg.append(f"{name}: {bad}")
            got = ops.samplegrid(size, Rc)
            dm = (got - want).abs()
            print(f"{H}x{W} ang {ang:+.2f} T={transpose}: kernel-vs-ref mismatches {int((got != want).sum())} / {got.numel()}, max |d| {float(dm.max()):.3e} | matmul candidates -> "
                  + " ; ".join(msg))
