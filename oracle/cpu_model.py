"""TEST / BASELINE INFRASTRUCTURE ONLY — PriOr-RAFT forward with the hot path on eager ATen ops.

`/root/reference` does not exist on the GPU box and the reference is pure PyTorch, so "the reference's own
CPU implementation of the path" is this: the network of `prior_flow_b200.model` (same parameters) with every
hot-path call routed to `oracle/torch_oracle.py` — op for op the ATen sequence of
PriOr-RAFT/core/prior_raft.py:107-215 (eight sample grids rebuilt per forward, separate matmul / div / 3x avg_pool,
168-launch DCCL, 123-launch flo_rotate, convex upsampling every iteration).  Used by `bench.py --impl reference`
and `bench.py`'s `cpu_baseline` leg (device "cpu"), and by the end-to-end parity test as the same-device eager
reference (device "cuda").  Never imported by the product.
"""
from __future__ import annotations

import math

import torch

from oracle import torch_oracle as TO
from prior_flow_b200.model import PriOrRAFT, convex_upsample


class EagerPriOrRAFT(PriOrRAFT):
    """Same modules and parameters as PriOrRAFT; forward restated on eager ops (no custom kernels)."""

    def forward(self, image1, image2, iters: int = 12, init_flow=None, test_mode: bool = False):
        dev = image1.device
        amp = lambda: torch.autocast(dev.type, enabled=self.args.mixed_precision and dev.type == "cuda")
        image1 = (2 * (image1 / 255.0) - 1.0).contiguous()
        image2 = (2 * (image2 / 255.0) - 1.0).contiguous()
        B, _, H, W = image1.shape
        size, size8 = (B, 3, H, W), (B, 3, H // 8, W // 8)
        g = {}
        for tag, ang in (("A2B", -math.pi / 2), ("B2A", math.pi / 2)):          # prior_raft.py:115-125
            R = TO.rotation_matrix([0.0, 0.0, ang], device=dev)
            g[tag] = TO.generate_samplegrid(size, R)
            g[tag + "_8x"] = TO.generate_samplegrid(size8, R)
            g[tag + "_W2C"] = TO.generate_samplegrid(size, R.T.contiguous())
            g[tag + "_W2C_8x"] = TO.generate_samplegrid(size8, R.T.contiguous())
        both_B = TO.img_rotate(torch.cat([image1, image2], dim=1), g["A2B"])
        image1_B, image2_B = both_B[:, :3].contiguous(), both_B[:, 3:].contiguous()
        with amp():
            cA, cB = self.cnet([image1, image1_B])
            net_A, inp_A = torch.tanh(cA[:, :128]), torch.relu(cA[:, 128:])
            net_B, inp_B = torch.tanh(cB[:, :128]), torch.relu(cB[:, 128:])
            f1A, f2A, f1B, f2B = (f.float() for f in self.fnet([image1, image2, image1_B, image2_B]))
        pyr_A = TO.build_pyramid(TO.corr_volume(f1A, f2A))
        pyr_B = TO.build_pyramid(TO.corr_volume(f1B, f2B))
        h, w = H // 8, W // 8
        coords0 = TO.coords_grid(B, h, w, dev)
        coords1_A, coords1_B = coords0.clone(), coords0.clone()
        if init_flow is not None:
            coords1_A = coords1_A + init_flow
            coords1_B = coords1_B + TO.flo_rotate(init_flow, g["A2B_W2C_8x"], g["A2B_8x"])
        preds_A, preds_B = [], []
        up_A = None
        for _ in range(iters):
            coords1_A, coords1_B = coords1_A.detach(), coords1_B.detach()
            flow_A, flow_B = coords1_A - coords0, coords1_B - coords0
            flaw_A = TO.warp_groupcorr(f1A, f2A, coords1_A)
            flow_B_A = TO.flo_rotate(flow_B, g["B2A_W2C_8x"], g["B2A_8x"])
            flaw_B_A = TO.warp_groupcorr(f1A, f2A, coords0 + flow_B_A)
            with amp():
                corr_A, corr_B_A = TO.dccl_lookup(coords1_A, pyr_A, pyr_B, g["A2B_W2C_8x"], g["B2A_8x"])
                corr_B, corr_A_B = TO.dccl_lookup(coords1_B, pyr_B, pyr_A, g["B2A_W2C_8x"], g["A2B_8x"])
                net_A, mask_A, d_A = self.ODDC(net_A, inp_A, flow_A, corr_A + corr_B_A, flaw_A, flow_B_A, flaw_B_A)
                net_B, mask_B, d_B = self.update_block(net_B, inp_B, corr_B + corr_A_B, flow_B)
            coords1_A = coords1_A + d_A
            coords1_B = coords1_B + d_B
            up_A = convex_upsample(coords1_A - coords0, mask_A)
            up_B = convex_upsample(coords1_B - coords0, mask_B)
            preds_A.append(up_A)
            preds_B.append(up_B)
        if test_mode:
            return up_A
        return preds_A, preds_B
