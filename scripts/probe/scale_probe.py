import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ddp_worker import batch, grads_of_step
from prior_flow_b200.model import PriOrRAFT
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(0)
m = PriOrRAFT().cuda(); m.train(); m.freeze_bn()
im1, im2, gt = batch(2, 128, 256, "cuda")
grads_of_step(m, im1, im2, gt, 1.0)
g1 = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
grads_of_step(m, im1, im2, gt, 2.0)
g2 = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
for k in ("ODDC.flow_head.conv2.bias", "fnet.conv1.weight", "update_block.flow_head.conv2.bias"):
    print(k, float(g1[k].abs().max()), float(g2[k].abs().max()), "ratio", float(g2[k].abs().max() / g1[k].abs().max()))
