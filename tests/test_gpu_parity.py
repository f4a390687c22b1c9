"""GPU: the CUDA path (through the C ABI) against the oracle and the golden vectors of the unmodified reference.

Bars (BASELINE.json north_star): sample coordinates and indices bit-exact; correlation and lookup values within
1e-5 of max|ref| on the fp32 path (stated separately for the f16 fast mode); everything through the same
`prior_flow_b200.ops` entry points the product uses.  The golden files hold the reference as executed on CPU
(true IEEE division), so these tests run the kernels in div_mode "ieee"; `test_gpu_torch_parity.py` checks the
"aten_cuda" flavour against ATen's own CUDA kernels at full size.
"""
import numpy as np
import pytest
import torch

import cases
from conftest import golden, rel_to_max
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _ieee_mode():
    from prior_flow_b200 import ops
    prev = ops.get_div_mode()
    ops.set_div_mode("ieee")
    yield
    ops.set_div_mode(prev)


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    torch.cuda.synchronize()
    return t.detach().cpu().numpy()


def test_library_is_the_native_one():
    from prior_flow_b200 import _lib
    lib = _lib.load()
    assert b"sm_100a" in lib.pf_build_info()
    assert torch.cuda.get_device_capability()[0] == 10


# ------------------------------------------------------------------ samplers / remaps (bit-exact)
def test_samplers_bit_exact():
    from prior_flow_b200 import ops
    img, pts = cases.small_sampler_case()
    g = golden("samplers")
    assert np.array_equal(host(ops.remap(cu(img), cu(pts), "BHW2", cyclic=True)), g["cyc"])
    assert np.array_equal(host(ops.remap(cu(img), cu(pts), "BHW2", cyclic=False)), g["lin"])


def test_img_rotate_bit_exact(geo):
    from prior_flow_b200 import ops
    out = ops.remap(cu(cases.image(seed=5)), cu(geo["a2b_64x128"]), "B2HW", cyclic=True)
    assert np.array_equal(host(out), golden("img_rotate")["out"])


def test_flo_rotate_bit_exact(geo):
    from prior_flow_b200 import ops
    g = golden("flo_rotate")
    gw, gc = np.repeat(geo["b2a_w2c_8x"], 2, 0), np.repeat(geo["b2a_8x"], 2, 0)
    assert np.array_equal(host(ops.flo_rotate(cu(cases.flow(seed=6)), cu(gw), cu(gc))), g["out"])
    assert np.array_equal(host(ops.flo_rotate(cu(cases.flow(seed=7, sigma=60.0)), cu(gw), cu(gc))), g["out_big"])
    # batch-invariant grids passed once (batch stride 0)
    assert np.array_equal(host(ops.flo_rotate(cu(cases.flow(seed=6)), cu(geo["b2a_w2c_8x"]), cu(geo["b2a_8x"]))), g["out"])


@pytest.mark.parametrize("tag,Rk,H,W,transpose", [
    ("a2b", "R_a2b", 128, 256, False), ("a2b_w2c_8x", "R_a2b", 16, 32, True),
    ("b2a_8x", "R_b2a", 16, 32, False), ("b2a_w2c", "R_b2a", 128, 256, True)])
def test_samplegrid(geo, tag, Rk, H, W, transpose):
    from prior_flow_b200 import ops
    R = geo[Rk].T if transpose else geo[Rk]
    got = host(ops.samplegrid((1, 3, H, W), torch.from_numpy(np.ascontiguousarray(R))))
    # libdevice vs the CPU libm differ by an ulp in cos/sin/asin/atan2: <= 1e-4 px at W = 256
    assert np.abs(got - geo[tag]).max() < 1e-4
    assert np.abs(got - O.generate_samplegrid((1, 3, H, W), R)).max() < 1e-4


def test_samplegrid_general_rotation_and_batch(geo):
    from prior_flow_b200 import ops
    got = host(ops.samplegrid((2, 3, 24, 40), torch.from_numpy(geo["R_gen"])))
    assert got.shape == (2, 2, 24, 40)
    assert np.abs(got - geo["gen_24x40"]).max() < 2e-5
    assert np.array_equal(got[0], got[1])


# ------------------------------------------------------------------ volume + pyramid
@pytest.mark.parametrize("mode,tol", [("fp32_simt", 1e-5), ("fp32", 1e-5), ("f16", 2e-3)])
def test_volume_pyramid_vs_golden(gold_pyramids, mode, tol):
    """fp32 path (CUDA-core FFMA and the tcgen05 3xFP16 split): 1e-5 of max|ref|.  f16 fast mode: 2e-3 (TF32-class)."""
    from prior_flow_b200 import ops
    f1a, f2a, f1b, f2b = cases.fmaps(seed=1)
    for (f1, f2), gold in zip(((f1a, f2a), (f1b, f2b)), gold_pyramids):
        pyr = ops.volume_pyramid(cu(f1), cu(f2), 4, mode)
        for l in range(4):
            got = host(pyr[l])
            assert got.shape == gold[l].shape
            assert rel_to_max(got, gold[l]) < tol, (mode, l)
        # pooling is exact given the finer level, in avg_pool2d's summation order
        for l in range(3):
            assert np.array_equal(O.avg_pool2x2(host(pyr[l])), host(pyr[l + 1])), (mode, l)


def test_volume_batched_matches_per_sample():
    from prior_flow_b200 import ops
    f1, f2 = cases.fmaps(seed=21, B=3, n=2)
    full = ops.volume_pyramid(cu(f1), cu(f2), 4, "fp32")
    for b in range(3):
        one = ops.volume_pyramid(cu(f1[b:b + 1]), cu(f2[b:b + 1]), 4, "fp32")
        for l in range(4):
            n = one[l].shape[0]
            # scales are per-tensor absmax powers of two, so per-sample and batched runs may round differently
            assert rel_to_max(host(full[l][b * n:(b + 1) * n]), host(one[l])) < 2e-6


def test_volume_scale_invariance():
    """The fp16 split rescales by a power of two chosen from absmax: results must not depend on input magnitude."""
    from prior_flow_b200 import ops
    f1, f2 = cases.fmaps(seed=22, n=2)
    base = host(ops.volume_pyramid(cu(f1), cu(f2), 1, "fp32")[0])
    for s1, s2 in ((1024.0, 1.0 / 4096), (2.0 ** -20, 2.0 ** 12)):
        got = host(ops.volume_pyramid(cu(f1 * np.float32(s1)), cu(f2 * np.float32(s2)), 1, "fp32")[0])
        assert np.array_equal(got, base * np.float32(s1 * s2))


def test_avg_pool_kernel_exact(gold_pyramids):
    from prior_flow_b200 import ops
    ga, _ = gold_pyramids
    for l in range(3):
        assert np.array_equal(host(ops.avg_pool2x2(cu(ga[l]))), ga[l + 1])


# ------------------------------------------------------------------ lookups
def _gold_pyr_cuda(gold_pyramids):
    ga, gb = gold_pyramids
    return [cu(x) for x in ga], [cu(x) for x in gb]


@pytest.mark.parametrize("tag,seed", [("rand", 2), ("edge", None)])
def test_dccl_lookup_vs_golden(geo, gold_pyramids, tag, seed):
    from prior_flow_b200 import ops
    pa, pb = _gold_pyr_cuda(gold_pyramids)
    c = cases.coords(seed=seed) if seed is not None else cases.edge_coords()
    own, other = ops.lookup(cu(c), pa, pb, cu(geo["a2b_w2c_8x"]), cu(geo["b2a_8x"]), radius=4)
    g = golden("dccl_" + tag)
    assert own.shape == (1, 324, 16, 32) and own.is_contiguous() and own.dtype == torch.float32
    # same FMA chain as ATen: expected bit-exact; the bar is 1e-5 of max|ref|
    assert rel_to_max(host(own), g["out_own"]) < 1e-5
    assert rel_to_max(host(other), g["out_other"]) < 1e-5
    assert np.array_equal(host(own), g["out_own"])
    assert np.array_equal(host(other), g["out_other"])


def test_dccl_lookup_other_direction(geo, gold_pyramids):
    from prior_flow_b200 import ops
    pa, pb = _gold_pyr_cuda(gold_pyramids)
    own, other = ops.lookup(cu(cases.coords(seed=3)), pb, pa, cu(geo["b2a_w2c_8x"]), cu(geo["a2b_8x"]), radius=4)
    g = golden("dccl_rand_b")
    assert np.array_equal(host(own), g["out_own"]) and np.array_equal(host(other), g["out_other"])


@pytest.mark.parametrize("edge", [False, True])
def test_lookup_coordinates_bit_exact(geo, gold_pyramids, edge):
    """Debug entry point: the unnormalised (ix, iy) of every tap of both branches == the oracle's, bit for bit."""
    from prior_flow_b200 import ops
    ga, gb = gold_pyramids
    pa, pb = _gold_pyr_cuda(gold_pyramids)
    c = cases.edge_coords() if edge else cases.coords(seed=2)
    (_, _), (d_own, d_other) = ops.lookup(cu(c), pa, pb, cu(geo["a2b_w2c_8x"]), cu(geo["b2a_8x"]), radius=4, debug=True)
    _, _, dbg = O.dccl_lookup(c, ga, gb, geo["a2b_w2c_8x"], geo["b2a_8x"], return_debug=True)
    d_own, d_other = host(d_own), host(d_other)
    for lvl in range(4):
        for got, ix, iy in ((d_own, dbg[lvl]["own_ix"], dbg[lvl]["own_iy"]),
                            (d_other, dbg[lvl]["other_ix"], dbg[lvl]["other_iy"])):
            assert np.array_equal(got[:, lvl, :, 0], ix.reshape(-1, 81)), lvl
            assert np.array_equal(got[:, lvl, :, 1], iy.reshape(-1, 81)), lvl
            assert np.array_equal(np.floor(got[:, lvl, :, 0]), np.floor(ix.reshape(-1, 81)))


@pytest.mark.parametrize("edge", [False, True])
def test_lookup_channels_last_and_fused_sum(geo, gold_pyramids, edge):
    """ABI v2 output modes: channels-last storage and `own + other` fused into the rotate kernel (prior_raft.py:187)."""
    from prior_flow_b200 import ops
    pa, pb = _gold_pyr_cuda(gold_pyramids)
    c = cu(cases.edge_coords() if edge else cases.coords(seed=2))
    gw, gc = cu(geo["a2b_w2c_8x"]), cu(geo["b2a_8x"])
    own, other = ops.lookup(c, pa, pb, gw, gc, radius=4)
    own_cl, other_cl = ops.lookup(c, pa, pb, gw, gc, radius=4, channels_last=True)
    assert own_cl.shape == own.shape and own_cl.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(own_cl, own) and torch.equal(other_cl, other)
    want = own + other
    assert torch.equal(ops.lookup(c, pa, pb, gw, gc, radius=4, fuse_sum=True), want)
    fused_cl = ops.lookup(c, pa, pb, gw, gc, radius=4, channels_last=True, fuse_sum=True)
    assert fused_cl.is_contiguous(memory_format=torch.channels_last) and torch.equal(fused_cl, want)
    single_cl = ops.lookup(c, pa, radius=4, cyclic=False, channels_last=True)
    assert torch.equal(single_cl, ops.lookup(c, pa, radius=4, cyclic=False))


def test_corrblock_lookup(gold_pyramids):
    from prior_flow_b200 import ops
    pa, _ = _gold_pyr_cuda(gold_pyramids)
    g = golden("corrblock")
    assert np.array_equal(host(ops.lookup(cu(cases.coords(seed=2)), pa, radius=4, cyclic=False)), g["out"])
    assert np.array_equal(host(ops.lookup(cu(cases.edge_coords()), pa, radius=4, cyclic=False)), g["out_edge"])


def test_lookup_other_radius_and_levels(gold_pyramids):
    from prior_flow_b200 import ops
    ga, _ = gold_pyramids
    c = cases.coords(seed=9)
    got = host(ops.lookup(cu(c), [cu(x) for x in ga[:2]], radius=2, cyclic=False))
    assert np.array_equal(got, O.corrblock_lookup(c, ga[:2], radius=2))


def test_lookup_batched_and_ragged_query_count(geo):
    """B = 2 and a query grid whose size is not a multiple of the 32-query CTA chunk."""
    from prior_flow_b200 import ops
    rs = np.random.RandomState(5)
    B, h, w = 2, 6, 10                       # N = 60
    pa = [rs.randn(B * h * w, 1, h >> l, w >> l).astype(np.float32) for l in range(2)]
    pb = [rs.randn(B * h * w, 1, h >> l, w >> l).astype(np.float32) for l in range(2)]
    R = O.generate_rotation_matrix([0., 0., -np.pi / 2])
    gw = O.generate_samplegrid((B, 3, h, w), R.T)
    gc = O.generate_samplegrid((B, 3, h, w), O.generate_rotation_matrix([0., 0., np.pi / 2]))
    c = cases.coords(seed=8, B=B, h=h, w=w, sigma=2.0)
    own, other = ops.lookup(cu(c), [cu(x) for x in pa], [cu(x) for x in pb], cu(gw), cu(gc), radius=3)
    want_own, want_other = O.dccl_lookup(c, pa, pb, gw, gc, radius=3)
    assert np.array_equal(host(own), want_own)
    assert np.array_equal(host(other), want_other)


def test_warp_groupcorr():
    from prior_flow_b200 import ops
    f1a, f2a, _, _ = cases.fmaps(seed=1)
    got = host(ops.warp_groupcorr(cu(f1a), cu(f2a), cu(cases.coords(seed=4)), 4))
    assert rel_to_max(got, golden("warp_groupcorr")["flaw"]) < 1e-5      # mean over 64 channels: order differs


# ------------------------------------------------------------------ on-the-fly variant
def test_onthefly_equals_materialised(geo):
    """No reference oracle exists (alt_cuda_corr unshipped): must equal the lookup into the materialised pyramid up to
    rounding — 1e-5 of max|ref|."""
    from prior_flow_b200 import ops
    f1a, f2a, f1b, f2b = (cu(x) for x in cases.fmaps(seed=1))
    pa, pb = ops.volume_pyramid(f1a, f2a, 4, "fp32_simt"), ops.volume_pyramid(f1b, f2b, 4, "fp32_simt")
    gw, gc = cu(geo["a2b_w2c_8x"]), cu(geo["b2a_8x"])
    for c in (cases.coords(seed=2), cases.edge_coords()):
        want_own, want_other = ops.lookup(cu(c), pa, pb, gw, gc, radius=4)
        got_own, got_other = ops.lookup_onthefly(
            cu(c), f1a.permute(0, 2, 3, 1).contiguous(), ops.channels_last_pyramid(f2a, 4),
            f1b.permute(0, 2, 3, 1).contiguous(), ops.channels_last_pyramid(f2b, 4), gw, gc, radius=4)
        assert rel_to_max(host(got_own), host(want_own)) < 1e-5
        assert rel_to_max(host(got_other), host(want_other)) < 1e-5


# ------------------------------------------------------------------ backward kernels vs autograd of the torch oracle
def test_backward_kernels_match_autograd(geo):
    from prior_flow_b200 import ops
    from oracle import torch_oracle as TO
    f1a, f2a, f1b, f2b = (cu(x) for x in cases.fmaps(seed=31, h=8, w=32))      # N = 256
    R1, R2 = TO.rotation_matrix([0., 0., -np.pi / 2]), TO.rotation_matrix([0., 0., np.pi / 2])
    gw = TO.generate_samplegrid((1, 3, 8, 32), R1.T.contiguous().cuda())
    gc = TO.generate_samplegrid((1, 3, 8, 32), R2.cuda())
    c = cu(cases.coords(seed=32, h=8, w=32, sigma=3.0))
    wa = torch.randn(1, 243, 8, 32, device="cuda")
    wb = torch.randn(1, 243, 8, 32, device="cuda")
    wf = torch.randn(1, 4, 8, 32, device="cuda")

    def run(use_kernels, fused=False):
        leaves = [t.clone().requires_grad_(True) for t in (f1a, f2a, f1b, f2b)]
        a1, a2, b1, b2 = leaves
        if use_kernels:
            pa = ops.volume_pyramid_autograd(a1, a2, 3, "fp32_simt")
            pb = ops.volume_pyramid_autograd(b1, b2, 3, "fp32_simt")
            if fused:   # channels-last + fused sum: one output, gradient flows to both branches
                both = ops.lookup_autograd(c, pa, pb, gw, gc, 4, channels_last=True, fuse_sum=True)
                own, other = both, torch.zeros_like(both)
            else:
                own, other = ops.lookup_autograd(c, pa, pb, gw, gc, 4)
            flaw = ops.warp_groupcorr_autograd(a1, a2, c, 4)
        else:
            pa, pb = TO.build_pyramid(TO.corr_volume(a1, a2), 3), TO.build_pyramid(TO.corr_volume(b1, b2), 3)
            own, other = TO.dccl_lookup(c, pa, pb, gw, gc, 4)
            flaw = TO.warp_groupcorr(a1, a2, c, 4)
        wb_ = wa if fused else wb
        loss = (own * wa).sum() + (other * wb_).sum() + (flaw * wf).sum()
        loss.backward()
        return [t.grad.detach().cpu().numpy() for t in leaves]

    prev = ops.get_div_mode()
    ops.set_div_mode("aten_cuda")
    try:
        got, want = run(True), run(False)
        got_fused, want_fused = run(True, fused=True), run(False, fused=True)
    finally:
        ops.set_div_mode(prev)
    for g, w_ in list(zip(got, want)) + list(zip(got_fused, want_fused)):
        assert rel_to_max(g, w_) < 2e-5      # atomics + different summation order


def test_remap_backward_matches_autograd(geo):
    from prior_flow_b200 import ops
    from oracle import torch_oracle as TO
    src = cu(cases.image(seed=41, H=16, W=32, ch=5))
    grid = cu(geo["a2b_8x"])
    w_ = torch.randn(1, 5, 16, 32, device="cuda")
    a = src.clone().requires_grad_(True)
    (ops.remap_autograd(a, grid, "B2HW", True) * w_).sum().backward()
    b = src.clone().requires_grad_(True)
    (TO.img_rotate(b, grid) * w_).sum().backward()
    assert rel_to_max(host(a.grad), host(b.grad)) < 1e-5


# ------------------------------------------------------------------ shapes the tensor-core tiling does not cover
@pytest.mark.parametrize("B,C,h,w,L", [(2, 64, 5, 9, 2), (1, 96, 11, 24, 3), (1, 128, 8, 32, 4)])
def test_odd_shapes_volume_pyramid_lookup(B, C, h, w, L):
    """Arbitrary (odd, non-multiple-of-8) feature-map sizes: the host layer falls back to the CUDA-core volume kernel,
    pooling floors like avg_pool2d, lookups handle ragged query counts.  Checked against the numpy oracle."""
    from prior_flow_b200 import ops
    rs = np.random.RandomState(77)
    f1 = (rs.randn(B, C, h, w) * 1.45).astype(np.float32)
    f2 = (rs.randn(B, C, h, w) * 1.45).astype(np.float32)
    want = O.build_pyramid(O.corr_volume(f1, f2), L)
    got = ops.volume_pyramid(cu(f1), cu(f2), L)            # default mode "fp32": tcgen05 when tileable, else CUDA cores
    for l in range(L):
        assert got[l].shape == want[l].shape
        assert rel_to_max(host(got[l]), want[l]) < 1e-5
    c = cases.coords(seed=5, B=B, h=h, w=w, sigma=1.5)
    out = host(ops.lookup(cu(c), [cu(x) for x in want], radius=2, cyclic=True))
    ref_outs = []
    for lvl in range(L):
        Hl, Wl = want[lvl].shape[-2:]
        px, py = O.window_points(c, lvl, 2)
        ix, iy = O.pixel_to_sample_coords(px, py, Hl, Wl, True)
        ref_outs.append(O.bilinear_zeros(want[lvl], ix, iy).reshape(B, h, w, 25))
    assert np.array_equal(out, np.concatenate(ref_outs, -1).transpose(0, 3, 1, 2))


def test_bad_arguments_raise_with_library_message():
    from prior_flow_b200 import _lib, ops
    with pytest.raises(_lib.PriorCorrError, match="pyramid too deep"):
        ops.lookup(torch.zeros(1, 2, 4, 4, device="cuda"), [torch.zeros(16, 1, 4 >> l, 4 >> l, device="cuda") for l in range(4)],
                   radius=4, cyclic=False)
    with pytest.raises(ValueError):
        ops.volume_pyramid(torch.zeros(1, 64, 8, 32, device="cuda"), torch.zeros(1, 64, 8, 16, device="cuda"))
    with pytest.raises(TypeError):
        ops.flo_rotate(torch.zeros(1, 2, 8, 16, device="cuda", dtype=torch.float16), torch.zeros(1, 2, 8, 16, device="cuda"),
                       torch.zeros(1, 2, 8, 16, device="cuda"))
