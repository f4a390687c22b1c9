"""GPU: (f1) img_rotate + `own + other` + Conv2d(324, 256, 1) + bias + ReLU in one tcgen05 kernel (csrc/pf_conv.cu) against
`F.relu(conv(corr_own + corr_other))` (core/prior_raft.py:187-188 + core/update.py:168,184 / :85,92).  The operand of the
convolution is bit-identical to the un-fused path (same FMA chain, same fp32 add — checked through a unit weight matrix); the
contraction is toleranced: fp32 mode (fp16 hi/lo, three products) 1e-5 of max|ref| and no less accurate than cuDNN's fp32
kernel against an fp64 reference; TF32-class mode (one product) 2e-3.  Tolerance = max-abs error / max-abs reference."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import torch_oracle as TO

pytestmark = pytest.mark.gpu


def scene(B, h, w, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    fm = [torch.randn(B, 256, h, w, device="cuda", generator=g) * 1.45 for _ in range(4)]
    coords = TO.coords_grid(B, h, w, "cuda") + torch.randn(B, 2, h, w, device="cuda", generator=g) * 5.0
    Ra = TO.rotation_matrix([0., 0., -np.pi / 2], device="cuda")
    Rb = TO.rotation_matrix([0., 0., np.pi / 2], device="cuda")
    gw = TO.generate_samplegrid((B, 3, h, w), Ra.T.contiguous())
    gc = TO.generate_samplegrid((B, 3, h, w), Rb)
    conv = torch.nn.Conv2d(324, 256, 1).cuda()
    with torch.no_grad():
        conv.bias.mul_(10)          # make the bias matter
    return fm, coords, gw, gc, conv


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


@pytest.mark.parametrize("B,h,w", [(1, 64, 128), (2, 16, 32), (1, 24, 44)])
def test_fused_rotate_sum_conv_relu(B, h, w):
    from prior_flow_b200 import ops
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        fm, coords, gw, gc, conv = scene(B, h, w, 7 + h)
        mode = "fp32" if ops.tcgen05_shape_ok(256, h, w) else "fp32_simt"
        pa, pb = ops.volume_pyramid(fm[0], fm[1], 4, mode), ops.volume_pyramid(fm[2], fm[3], 4, mode)
        x = ops.lookup(coords, pa, pb, gw, gc, 4, fuse_sum=True)                      # the un-fused operand [B,324,h,w]
        with torch.no_grad():
            want32 = F.relu(conv(x))
            want64 = F.relu(F.conv2d(x.double(), conv.weight.double(), conv.bias.double()))
            for cl in (False, True):
                got = ops.lookup_conv(coords, pa, pb, gw, gc, conv.weight, conv.bias, channels_last=cl, fp32=True)
                assert got.shape == want32.shape
                if cl:
                    assert got.is_contiguous(memory_format=torch.channels_last)
                e, e_cudnn = rel(got, want64), rel(want32, want64)
                print(f"\n[f1 {B}x{h}x{w} cl={cl}] fused fp32-mode vs fp64 {e:.2e} (cuDNN fp32 vs fp64 {e_cudnn:.2e}); vs cuDNN fp32 {rel(got, want32):.2e}")
                assert rel(got, want32) < 1e-5
                assert e < 4 * e_cudnn + 1e-7
            fast = ops.lookup_conv(coords, pa, pb, gw, gc, conv.weight, conv.bias, fp32=False)
            assert rel(fast, want64) < 2e-3
            assert (got >= 0).all()
    finally:
        torch.backends.cudnn.allow_tf32 = old


def test_fused_operand_is_bit_identical_to_the_unfused_sum():
    """Identity-like weights: W = [I_256 | 0] scaled by a power of two, zero bias -> the kernel must return relu(X[:, :256])
    exactly (every product is exact in the three-product scheme when W has one power-of-two entry per row)."""
    from prior_flow_b200 import ops
    B, h, w = 1, 16, 32
    fm, coords, gw, gc, _ = scene(B, h, w, 3)
    pa, pb = ops.volume_pyramid(fm[0], fm[1], 4, "fp32_simt"), ops.volume_pyramid(fm[2], fm[3], 4, "fp32_simt")
    x = ops.lookup(coords, pa, pb, gw, gc, 4, fuse_sum=True)
    for shift in (0, 68):
        W = torch.zeros(256, 324, device="cuda")
        W[torch.arange(256), torch.arange(256) + shift] = 0.5
        got = ops.lookup_conv(coords, pa, pb, gw, gc, W, torch.zeros(256, device="cuda"), fp32=True)
        want = F.relu(0.5 * x[:, shift:shift + 256])
        # hi + lo carries 22 bits of X: the fused operand can differ from the fp32 X by at most 2^-22 relative
        assert float((got - want).abs().max()) <= 2.0 ** -21 * float(want.abs().max())


def test_model_with_fused_first_layer_matches_unfused():
    from prior_flow_b200.model import PriOrRAFT
    import cases
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(0)
        m = PriOrRAFT().cuda().eval()
        im1, im2 = (torch.from_numpy(x).cuda() for x in cases.e2e_images())
        with torch.no_grad():
            m.fuse_conv1 = False
            a = m(im1, im2, iters=6, test_mode=True)
            m.fuse_conv1 = True
            b = m(im1, im2, iters=6, test_mode=True)
            mc = PriOrRAFT().cuda().eval()
            mc.load_state_dict(m.state_dict())
            mc = mc.to_channels_last()
            mc.fuse_conv1 = True
            c = mc(im1, im2, iters=6, test_mode=True)
        epe = lambda p, q: float(torch.sqrt(((p.double() - q.double()) ** 2).sum(1)).mean())
        print(f"\n[f1 e2e] fused-vs-unfused first layer: NCHW {epe(a, b):.3e} px, channels_last {epe(a, c):.3e} px")
        assert epe(a, b) < 1e-3 and epe(a, c) < 1e-3
    finally:
        torch.backends.cudnn.allow_tf32 = old


def test_prepared_weight_planes_follow_the_tensor_not_its_address():
    """r03z regression: the prepared planes were cached by (address, version counter); a new layer allocated where a freed one
    had been (same address, version 0) got the old layer's planes.  They now live on the tensor object.  `.data` gives the
    deterministic stand-in for that: a fresh tensor object at the same address with its own version counter at 0."""
    from prior_flow_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    w1 = torch.randn(256, 324, 1, 1, device="cuda", generator=g)
    p1 = ops.prepare_conv_weight(w1)
    assert ops.prepare_conv_weight(w1) is p1                            # cached on the tensor
    old_planes = p1.clone()
    w1.data.copy_(torch.randn(256, 324, 1, 1, device="cuda", generator=g))      # new contents, w1's own counter untouched
    w2 = w1.data
    assert w2.data_ptr() == w1.data_ptr() and w2._version == w1._version == 0
    assert not torch.equal(ops.prepare_conv_weight(w2), old_planes)     # an (address, version) cache would return old_planes
    before = ops.prepare_conv_weight(w2).clone()
    w2.mul_(1.5)                                                        # in-place update: version bump -> prepared again
    assert not torch.equal(ops.prepare_conv_weight(w2), before)
