"""GPU box: is d(loss)/d(theta) of a batch of 4 the sum of the gradients of its two halves?  (isolates a batch-dependent backward bug)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ddp_worker import batch, grads_of_step
from prior_flow_b200.model import PriOrRAFT
from prior_flow_b200 import ops

torch.backends.cudnn.allow_tf32 = False
def run(label, **kw):
    torch.manual_seed(0)
    m = PriOrRAFT().cuda(); m.train(); m.freeze_bn()
    for k, v in kw.items():
        setattr(m, k, v)
    im1, im2, gt = batch(4, 128, 256, "cuda")
    grads_of_step(m, im1, im2, gt, 1.0)
    full = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    acc = None
    for sl in (slice(0, 2), slice(2, 4)):
        grads_of_step(m, im1[sl], im2[sl], gt[sl], 1.0)
        g = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        acc = g if acc is None else {k: acc[k] + g[k] for k in g}
    grads_of_step(m, im1, im2, gt, 1.0)
    again = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    gmax = max(float(v.abs().max()) for v in full.values())
    def worst(a, b):
        w, n = 0.0, ""
        for k in b:
            e = float((a[k] - b[k]).abs().max()) / max(float(b[k].abs().max()), 1e-4 * gmax)
            if e > w: w, n = e, k
        return w, n
    print(label, "| full vs halves-summed:", worst(acc, full), "| full vs full again:", worst(again, full),
          "| |g| fnet.conv1", float(full["fnet.conv1.weight"].abs().max()), "gmax", gmax, flush=True)

run("default")
run("no grad sink", accumulate_grads=False)
_vb = ops.volume_backward
ops.volume_backward = lambda *a, **k: _vb(*a, **{**k, "use_library": True})
run("library volume backward")
