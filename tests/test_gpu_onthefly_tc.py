"""GPU: the tensor-core on-the-fly lookup (pf_lookup_onthefly_tc: tile boxes -> tcgen05 dots into per-query local planes ->
blend) against the materialised DCCL lookup and against the CUDA-core on-the-fly kernel it replaces.

Bar: 1e-5 of max|ref| (the volume kernel's own bar: the dots are the same fp16 hi/lo three-product contraction).  Scenes:
a smooth flow field (what the network produces: nearly every tile takes the tensor-core path), i.i.d. noise of sigma 5 px per
query (tile boxes overflow 32 x 24 at the fine levels: mixed paths inside one launch), and a constant shift across the ERP
seam (own-view windows wrap: CUDA-core fallback tiles next to tensor-core tiles)."""
import pytest
import torch

from oracle import torch_oracle as TO
from test_gpu_configs import make_scene, rel

pytestmark = pytest.mark.gpu


def smooth_coords(B, h, w, seed, amp):
    g = torch.Generator(device="cuda").manual_seed(seed)
    low = torch.randn(B, 2, max(h // 16, 2), max(w // 16, 2), device="cuda", generator=g) * amp
    return TO.coords_grid(B, h, w, "cuda") + torch.nn.functional.interpolate(low, size=(h, w), mode="bilinear", align_corners=True)


def run_both(ops, coords, fm, gw, gc, dual=True):
    cl = lambda t: t.permute(0, 2, 3, 1).contiguous()
    f1a, f2a, f1b, f2b = cl(fm[0]), ops.channels_last_pyramid(fm[1], 4), cl(fm[2]), ops.channels_last_pyramid(fm[3], 4)
    pa, pb = ops.OnTheFlyPlanes(f1a, f2a), ops.OnTheFlyPlanes(f1b, f2b)
    if dual:
        tc = ops.lookup_onthefly(coords, f1a, f2a, f1b, f2b, gw, gc, 4, planes_own=pa, planes_other=pb)
        cc = ops.lookup_onthefly(coords, f1a, f2a, f1b, f2b, gw, gc, 4)
    else:
        tc = (ops.lookup_onthefly(coords, f1a, f2a, radius=4, planes_own=pa),)
        cc = (ops.lookup_onthefly(coords, f1a, f2a, radius=4),)
    return tc, cc


@pytest.mark.parametrize("B,h,w", [(1, 64, 128), (2, 32, 64), (1, 16, 32)])
@pytest.mark.parametrize("scene", ["smooth", "noise", "seam"])
def test_tensor_core_onthefly_matches_materialised(B, h, w, scene):
    from prior_flow_b200 import ops
    fm, noisy, gw, gc = make_scene(B, h, w, 100 + h)
    coords = {"smooth": smooth_coords(B, h, w, 7, 6.0), "noise": noisy,
              "seam": TO.coords_grid(B, h, w, "cuda") + torch.tensor([w / 2 - 3.3, 1.7], device="cuda").view(1, 2, 1, 1)}[scene]
    pyr_a, pyr_b = ops.volume_pyramid(fm[0], fm[1], 4, "fp32"), ops.volume_pyramid(fm[2], fm[3], 4, "fp32")
    want = ops.lookup(coords, pyr_a, pyr_b, gw, gc, 4)
    tc, cc = run_both(ops, coords, fm, gw, gc)
    for name, got, old, ref in zip(("own", "other"), tc, cc, want):
        e_tc, e_cc = rel(got, ref), rel(old, ref)
        print(f"\n[onthefly tc {scene} B{B} {h}x{w}] {name}: tensor-core {e_tc:.2e}, CUDA-core {e_cc:.2e} of max|ref|")
        assert got.shape == ref.shape
        assert e_tc < 1e-5


def test_small_pool_sends_the_overflow_to_the_cuda_core_path(monkeypatch):
    """A pool that holds only part of the boxes: the tiles past its end take otf_fallback_kernel; same results."""
    from prior_flow_b200 import ops
    B, h, w = 1, 64, 128
    fm, noisy, gw, gc = make_scene(B, h, w, 21)
    pyr_a, pyr_b = ops.volume_pyramid(fm[0], fm[1], 4, "fp32"), ops.volume_pyramid(fm[2], fm[3], 4, "fp32")
    for coords in (smooth_coords(B, h, w, 5, 6.0), noisy):
        want = ops.lookup(coords, pyr_a, pyr_b, gw, gc, 4)
        for per_query in (0.3, 0.02):
            monkeypatch.setattr(ops.OnTheFlyPlanes, "POOL_SEGMENTS_PER_QUERY", per_query)
            tc, _ = run_both(ops, coords, fm, gw, gc)
            work, T = ops._state["otf_work"][:2]
            on_cuda_cores = int(work[0])
            print(f"\n[onthefly tc pool {per_query} segments/query] tiles on the CUDA-core path: {on_cuda_cores} of {T}")
            assert on_cuda_cores > 0
            for got, ref in zip(tc, want):
                assert rel(got, ref) < 1e-5


def test_single_view_and_zero_features():
    from prior_flow_b200 import ops
    B, h, w = 1, 32, 64
    fm, _, gw, gc = make_scene(B, h, w, 5)
    coords = smooth_coords(B, h, w, 9, 4.0)
    (tc,), (cc,) = run_both(ops, coords, fm, gw, gc, dual=False)
    assert rel(tc, cc) < 1e-5
    zero = [torch.zeros_like(t) for t in fm]
    (tz, _), _ = run_both(ops, coords, zero, gw, gc)
    assert float(tz.abs().max()) == 0.0


def test_full_resolution_1024x2048_smooth_flow():
    """BASELINE configs[4] shape (128x256 features): tensor-core on-the-fly vs the CUDA-core kernel."""
    from prior_flow_b200 import ops
    B, h, w = 1, 128, 256
    fm, _, gw, gc = make_scene(B, h, w, 77)
    coords = smooth_coords(B, h, w, 3, 8.0)
    tc, cc = run_both(ops, coords, fm, gw, gc)
    for got, old in zip(tc, cc):
        assert rel(got, old) < 1e-5


def test_model_onthefly_mode_uses_planes_and_matches(monkeypatch):
    from prior_flow_b200 import corr as pcorr
    B, h, w = 1, 32, 64
    fm, _, gw, gc = make_scene(B, h, w, 11)
    coords = smooth_coords(B, h, w, 13, 5.0)
    outs = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("PF_ONTHEFLY_TC", flag)
        d = pcorr.DCCL(4, 4, mode="onthefly")
        with torch.no_grad():
            pa, pb = d.build_pyramid(pcorr.corr(fm[0], fm[1])), d.build_pyramid(pcorr.corr(fm[2], fm[3]))
            assert (pa.planes is not None) == (flag == "1")
            outs[flag] = d(coords, pa, pb, gw, gc)
    for a, b in zip(outs["1"], outs["0"]):
        assert rel(a, b) < 1e-5


@pytest.mark.parametrize("B,h,w", [(1, 64, 128), (2, 16, 32)])
def test_onthefly_lookup_with_fused_first_conv(B, h, w):
    """SURVEY §8 f1 behind the volume-free lookup: both views stay channels-last, pf_dccl_conv rotates, sums, convolves."""
    from prior_flow_b200 import ops
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        fm, _, gw, gc = make_scene(B, h, w, 31 + h)
        coords = smooth_coords(B, h, w, 17, 5.0)
        conv = torch.nn.Conv2d(324, 256, 1).cuda()
        cl = lambda t: t.permute(0, 2, 3, 1).contiguous()
        f1a, f2a, f1b, f2b = cl(fm[0]), ops.channels_last_pyramid(fm[1], 4), cl(fm[2]), ops.channels_last_pyramid(fm[3], 4)
        pa, pb = ops.OnTheFlyPlanes(f1a, f2a), ops.OnTheFlyPlanes(f1b, f2b)
        with torch.no_grad():
            own, other = ops.lookup_onthefly(coords, f1a, f2a, f1b, f2b, gw, gc, 4, planes_own=pa, planes_other=pb)
            want = torch.relu(conv(own + other))
            for chl in (False, True):
                got = ops.lookup_onthefly_conv(coords, f1a, f2a, f1b, f2b, gw, gc, pa, pb, conv.weight, conv.bias, channels_last=chl, fp32=True)
                assert got.shape == want.shape
                e = rel(got, want)
                print(f"\n[onthefly tc + conv1 B{B} {h}x{w} cl={chl}] vs cuDNN fp32 on the un-fused lookup: {e:.2e}")
                assert e < 1e-5
    finally:
        torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("B,h,w", [(1, 8, 16), (3, 24, 48), (1, 40, 112)])
def test_wild_and_non_finite_coordinates_match_the_cuda_core_path(B, h, w):
    """Flows of hundreds of pixels (boxes as large as the plane: pool overflow, CUDA-core tiles), NaN / inf / 1e30 coordinates (the
    sampler maps them to -100: no tap touches the plane), the smallest grid (one query tile), grids that are not powers of two."""
    from prior_flow_b200 import ops
    fm, _, gw, gc = make_scene(B, h, w, 900 + h)
    g = torch.Generator(device="cuda").manual_seed(h)
    coords = TO.coords_grid(B, h, w, "cuda") + torch.randn(B, 2, h, w, device="cuda", generator=g) * 60.0
    flat = coords.view(-1)
    idx = torch.randperm(flat.numel(), device="cuda", generator=g)[:max(8, flat.numel() // 50)]
    flat[idx[0::4]] = float("nan")
    flat[idx[1::4]] = float("inf")
    flat[idx[2::4]] = -float("inf")
    flat[idx[3::4]] = 1e30
    tc, cc = run_both(ops, coords, fm, gw, gc)
    for got, old in zip(tc, cc):
        assert torch.isfinite(got).all()
        assert rel(got, old) < 1e-5
