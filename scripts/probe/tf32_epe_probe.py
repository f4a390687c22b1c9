"""GPU box: which cuDNN-side option of model.py moves the flow away from the reference under TF32 convolutions?"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim
from prior_flow_b200 import model as M

ref = ref_shim.load()
def epe(a, b): return float(torch.sqrt(((a.double() - b.double()) ** 2).sum(1)).mean())
g = torch.Generator().manual_seed(1234)
im1, im2 = (torch.rand(1, 3, 512, 1024, generator=g) * 255).cuda(), (torch.rand(1, 3, 512, 1024, generator=g) * 255).cuda()
rm = ref_shim.make_model(ref, seed=0).cuda().eval()
def run(m):
    with torch.no_grad():
        return m(im1, im2, iters=12, test_mode=True)
torch.backends.cudnn.allow_tf32 = False
ref32 = run(rm)
torch.backends.cudnn.allow_tf32 = True
reftf = run(rm)
print(f"reference tf32 vs reference fp32: {epe(reftf, ref32):.3e} px")
for cl in (False, True):
    for fold in (False, True):
        for fuse in (False, True):
            for bench in (False, True):
                M.FOLD_BN_INFERENCE, M.FUSE_CONV_RELU = fold, fuse
                torch.backends.cudnn.benchmark = bench
                m = M.PriOrRAFT().cuda().eval()
                m.load_state_dict(rm.state_dict(), strict=True)
                if cl:
                    m = m.to_channels_last()
                out = run(m)
                print(f"channels_last={cl!s:5} fold_bn={fold!s:5} fuse_conv_relu={fuse!s:5} cudnn.benchmark={bench!s:5}: vs ref tf32 {epe(out, reftf):.3e}  vs ref fp32 {epe(out, ref32):.3e}")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cudnn.benchmark = False
M.FOLD_BN_INFERENCE, M.FUSE_CONV_RELU = True, True
m = M.PriOrRAFT().cuda().eval(); m.load_state_dict(rm.state_dict(), strict=True); m = m.to_channels_last()
print(f"fp32 convs, channels_last + fold + fuse: vs ref fp32 {epe(run(m), ref32):.3e}")
