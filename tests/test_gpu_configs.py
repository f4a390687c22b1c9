"""GPU: op-level parity at the shapes of BASELINE configs 3, 4 and 5 — what the 1024x2048 and batched bench records rest on.

  config 3 / 5: batch 8 at 512x1024 (8 pairs per GPU)        -> B = 8, 64x128 features, N = 8192
  config 4:     1024x2048 ERP, fully materialised and on-the-fly -> B = 1, 128x256 features, N = 32768 (level 0 = 4 GiB / view)

Reference = the ATen ops the reference executes (oracle/torch_oracle.py on the device, ATen's native grid_sampler kernel:
see tests/test_gpu_torch_parity.py for why not cuDNN's).  Bars: lookups bit-exact; volume 1e-5 of max|ref| (max-abs error over
max-abs reference, conftest.rel_to_max); on-the-fly 1e-5 of max|ref| against the materialised lookup.
Also here: sample grids bit-exact against the UNMODIFIED reference on CUDA, and the GradSink double-backward regression.
"""
import numpy as np
import pytest
import torch

from oracle import ref_shim
from oracle import torch_oracle as TO
from test_gpu_torch_parity import native_aten

pytestmark = pytest.mark.gpu
C = 256


def rel(a, b):
    """max |a-b| / max |b| on the device (the tensors here are up to 4 GiB)."""
    return float((a - b).abs().max() / b.abs().max())


def make_scene(B, h, w, seed):
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(seed)
    fm = [torch.randn(B, C, h, w, device="cuda", generator=g) * 1.45 for _ in range(4)]
    coords = TO.coords_grid(B, h, w, "cuda") + torch.randn(B, 2, h, w, device="cuda", generator=g) * 5.0
    Ra = TO.rotation_matrix([0., 0., -np.pi / 2], device="cuda")
    Rb = TO.rotation_matrix([0., 0., np.pi / 2], device="cuda")
    gw = TO.generate_samplegrid((B, 3, h, w), Ra.T.contiguous())
    gc = TO.generate_samplegrid((B, 3, h, w), Rb)
    return fm, coords, gw, gc


def test_config4_shape_128x256_features():
    """1024x2048 ERP: N = 32768 query pixels, level 0 of one view = 4 GiB."""
    from prior_flow_b200 import ops
    B, h, w = 1, 128, 256
    fm, coords, gw, gc = make_scene(B, h, w, 41)
    pa = ops.volume_pyramid(fm[0], fm[1], 4, "fp32")
    ref_a = TO.build_pyramid(TO.corr_volume(fm[0], fm[1]))
    for l in range(4):
        assert pa[l].shape == ref_a[l].shape == (B * h * w, 1, h >> l, w >> l)
        e = rel(pa[l], ref_a[l])
        assert e < 1e-5, (l, e)
    del pa
    ref_b = TO.build_pyramid(TO.corr_volume(fm[2], fm[3]))
    own, other = ops.lookup(coords, ref_a, ref_b, gw, gc, 4)
    with native_aten():
        want_own, want_other = TO.dccl_lookup(coords, ref_a, ref_b, gw, gc, 4)
    assert torch.equal(own, want_own)
    assert torch.equal(other, want_other)
    fused = ops.lookup(coords, ref_a, ref_b, gw, gc, 4, fuse_sum=True)
    assert torch.equal(fused, want_own + want_other)
    del ref_a, ref_b
    torch.cuda.empty_cache()
    cl = lambda t: t.permute(0, 2, 3, 1).contiguous()
    o2, x2 = ops.lookup_onthefly(coords, cl(fm[0]), ops.channels_last_pyramid(fm[1], 4), cl(fm[2]),
                                 ops.channels_last_pyramid(fm[3], 4), gw, gc, 4)
    e_own, e_other = rel(o2, want_own), rel(x2, want_other)
    print(f"\n[config 4 shape] on-the-fly vs materialised: own {e_own:.2e}, other {e_other:.2e} of max|ref|")
    assert e_own < 1e-5 and e_other < 1e-5


def test_config3_shape_batch8_at_64x128_features():
    """8 pairs per GPU at 512x1024: every kernel with B = 8 against the per-sample results and against ATen."""
    from prior_flow_b200 import ops
    B, h, w = 8, 64, 128
    fm, coords, gw, gc = make_scene(B, h, w, 43)
    pa, pb = ops.volume_pyramid(fm[0], fm[1], 4, "fp32"), ops.volume_pyramid(fm[2], fm[3], 4, "fp32")
    N = h * w
    for b in (0, 3, 7):        # batched build == per-sample build, bit for bit
        one = ops.volume_pyramid(fm[0][b:b + 1], fm[1][b:b + 1], 4, "fp32")
        for l in range(4):
            assert torch.equal(pa[l][b * N:(b + 1) * N], one[l]), (b, l)
    ref_a = TO.build_pyramid(TO.corr_volume(fm[0], fm[1]))
    for l in range(4):
        assert rel(pa[l], ref_a[l]) < 1e-5, l
    del ref_a
    own, other = ops.lookup(coords, pa, pb, gw, gc, 4)
    with native_aten():
        want_own, want_other = TO.dccl_lookup(coords, pa, pb, gw, gc, 4)
    assert torch.equal(own, want_own)
    assert torch.equal(other, want_other)
    fused_cl = ops.lookup(coords, pa, pb, gw, gc, 4, channels_last=True, fuse_sum=True)
    assert torch.equal(fused_cl, want_own + want_other)
    flow = coords - TO.coords_grid(B, h, w, "cuda")
    assert torch.equal(ops.flo_rotate(flow, gw, gc), TO.flo_rotate(flow, gw, gc))
    got = ops.warp_groupcorr(fm[0], fm[1], coords, 4)
    assert rel(got, TO.warp_groupcorr(fm[0], fm[1], coords, 4)) < 1e-5
    cl = lambda t: t.permute(0, 2, 3, 1).contiguous()
    o2, x2 = ops.lookup_onthefly(coords, cl(fm[0]), ops.channels_last_pyramid(fm[1], 4), cl(fm[2]),
                                 ops.channels_last_pyramid(fm[3], 4), gw, gc, 4)
    assert rel(o2, want_own) < 1e-5 and rel(x2, want_other) < 1e-5


@pytest.mark.skipif(not ref_shim.available(), reason="vendored reference missing (scripts/vendor_reference.py)")
def test_samplegrids_bit_exact_against_the_unmodified_reference_on_cuda():
    """generate_samplegrid (projection_prim_ortho.py:432-443) run by the reference itself on this GPU vs pf_samplegrid:
    every one of the eight grids `PriOr_RAFT.forward` builds (prior_raft.py:115-125), `==`.  (For these rotations the
    3x3 product of rotate_cartesian has entries 0, +-1 and -4.37e-8, so cuBLAS's accumulation order cannot matter; a general
    rotation is only pinned to 2 ulp by tests/test_gpu_torch_parity.py::test_samplegrids_full_size.)"""
    from prior_flow_b200 import geometry as geo
    from prior_flow_b200 import ops
    ref = ref_shim.load()
    assert ref_shim.verified()
    for H, W in ((64, 128), (128, 256), (512, 1024)):
        for ang in (-np.pi / 2, np.pi / 2):
            R = ref.ppo.generate_rotation_metrix(theta_list=[0., 0., ang])
            assert torch.equal(R, geo.generate_rotation_metrix(theta_list=[0., 0., ang]))
            for Rm in (R, R.T):
                want = ref.ppo.generate_samplegrid((1, 3, H, W), Rm)
                got = ops.samplegrid((1, 3, H, W), Rm.contiguous())
                assert torch.equal(got, want), (H, W, ang)
                assert torch.equal(geo.generate_samplegrid((1, 3, H, W), Rm), want)


def test_grad_sink_survives_a_second_backward():
    """with the shared gradient sink a second backward over a retained graph used to scatter into stale buffers and
    hand the volume zero gradient.  Two backward passes must each equal the sink-free result."""
    from prior_flow_b200.corr import DCCL, CostVolume
    B, h, w = 1, 16, 32
    fm, coords, gw, gc = make_scene(B, h, w, 47)

    def run(accumulate):
        f = [t.clone().requires_grad_(True) for t in fm]
        look = DCCL(4, 4, mode="materialized", accumulate_grads=accumulate)
        pa, pb = look.build_pyramid(CostVolume(f[0], f[1])), look.build_pyramid(CostVolume(f[2], f[3]))
        loss = 0
        for k in range(3):      # three iterations, both call directions, like PriOr_RAFT.forward
            c = coords + 0.37 * k
            loss = loss + look.summed(c, pa, pb, gw, gc).square().mean() + look.summed(c, pb, pa, gc, gw).abs().mean()
        grads = []
        for _ in range(2):
            for t in f:
                t.grad = None
            loss.backward(retain_graph=True)
            grads.append([t.grad.clone() for t in f])
        return grads

    plain, sunk = run(False), run(True)
    for pass_ in range(2):
        for a, b in zip(plain[pass_], sunk[pass_]):
            assert float(b.abs().max()) > 0
            assert float((a - b).abs().max() / a.abs().max()) < 1e-4      # atomics order only
    for a, b in zip(sunk[0], sunk[1]):
        assert float((a - b).abs().max() / a.abs().max()) < 1e-4


def test_grad_sink_refuses_a_foreign_consumer():
    """A pyramid level that also feeds something other than a DCCL lookup invalidates the in-place sink: loud error, not a
    silently wrong gradient."""
    from prior_flow_b200.corr import DCCL, CostVolume
    fm, coords, gw, gc = make_scene(1, 16, 32, 53)
    f = [t.clone().requires_grad_(True) for t in fm]
    look = DCCL(4, 4, mode="materialized", accumulate_grads=True)
    pa, pb = look.build_pyramid(CostVolume(f[0], f[1])), look.build_pyramid(CostVolume(f[2], f[3]))
    loss = look.summed(coords, pa, pb, gw, gc).mean() + look.summed(coords + 1, pa, pb, gw, gc).mean() + pa[0].mean()
    with pytest.raises(RuntimeError, match="GradSink"):
        loss.backward()


@pytest.mark.parametrize("B,h,w", [(1, 64, 128), (2, 16, 32)])
def test_volume_backward_tcgen05_matches_fp32_gemms(B, h, w):
    """pf_volume_bwd (bf16 hi/lo, three products) against the two fp32 library GEMMs it replaces and an fp64 reference:
    stated tolerance 1e-4 of max|ref| (measured ~1e-5); gradients spanning 12 orders of magnitude need no scaling pass."""
    from prior_flow_b200 import ops
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(61 + h)
    f1 = torch.randn(B, C, h, w, device="cuda", generator=g) * 1.45
    f2 = torch.randn(B, C, h, w, device="cuda", generator=g) * 1.45
    N = h * w
    for mag in (1.0, 1e-9, 1e3):
        dV = torch.randn(B, N, N, device="cuda", generator=g) * mag
        dV[:, :, ::7] *= 1e-3                                  # mixed magnitudes inside one tile
        d1, d2 = ops.volume_backward(f1, f2, dV)
        l1, l2 = ops.volume_backward(f1, f2, dV, use_library=True)
        r1 = (torch.matmul(f2.double().view(B, C, N), dV.double().transpose(1, 2)) / 16.0).view_as(f1)
        r2 = (torch.matmul(f1.double().view(B, C, N), dV.double()) / 16.0).view_as(f2)
        e1, e2 = rel(d1.double(), r1), rel(d2.double(), r2)
        print(f"\n[volume bwd {B}x{h}x{w} |dV|~{mag:g}] tcgen05 vs fp64: dF1 {e1:.2e} dF2 {e2:.2e}; cuBLAS fp32 vs fp64: {rel(l1.double(), r1):.2e} {rel(l2.double(), r2):.2e}")
        assert e1 < 1e-4 and e2 < 1e-4
    only1, none2 = ops.volume_backward(f1, f2, dV, need2=False)
    assert none2 is None and torch.equal(only1, d1)


@pytest.mark.parametrize("B,h,w,chunk_bytes", [(1, 16, 32, 1 << 30), (2, 16, 32, 3 << 20), (1, 24, 40, 1 << 30)])
def test_onthefly_backward_matches_materialised_backward(B, h, w, chunk_bytes):
    """Volume-free backward of the on-the-fly lookup (ops.OnTheFlyTape: recorded calls replayed chunk by chunk over the queries,
    pf_lookup_dual_bwd with a query range + pf_volume_bwd on the chunk) against the materialised path's backward
    (gradient pyramids for the whole volume): same loss, gradients to 1e-4 of max|ref| (bf16x2 contraction, atomics order).
    The second case forces several chunks (3 MB of gradient rows at a time); the third is a shape pf_volume_bwd does not tile
    (h*w = 960), which takes the library-GEMM branch of volume_backward_chunk."""
    from prior_flow_b200 import ops
    from prior_flow_b200.corr import DCCL, CostVolume
    fm, coords, gw, gc = make_scene(B, h, w, 71)
    old = ops.OnTheFlyTape.CHUNK_BYTES
    ops.OnTheFlyTape.CHUNK_BYTES = chunk_bytes

    def run(mode):
        f = [t.clone().requires_grad_(True) for t in fm]
        look = DCCL(4, 4, mode=mode)
        pa, pb = look.build_pyramid(CostVolume(f[0], f[1])), look.build_pyramid(CostVolume(f[2], f[3]))
        loss = 0
        for k in range(3):
            c = coords + 0.41 * k
            o1, x1 = look(c, pa, pb, gw, gc)
            o2, x2 = look(c, pb, pa, gc, gw)
            loss = loss + (o1 * (k + 1)).square().mean() + x1.abs().mean() + (o2 + x2).square().mean()
        loss.backward()
        return float(loss), [t.grad.clone() for t in f]

    try:
        l_mat, g_mat = run("materialized")
        l_otf, g_otf = run("onthefly")
    finally:
        ops.OnTheFlyTape.CHUNK_BYTES = old
    assert abs(l_mat - l_otf) <= 1e-5 * abs(l_mat)
    worst = max(rel(a, b) for a, b in zip(g_otf, g_mat))
    print(f"\n[on-the-fly backward {B}x{h}x{w}] loss {l_otf:.6f} vs {l_mat:.6f}; worst gradient difference {worst:.2e} of max|ref|")
    assert all(float(g.abs().max()) > 0 for g in g_otf)
    assert worst < 1e-4
