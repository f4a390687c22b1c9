"""GPU: SURVEY §8 (f2) convex upsampling, (f3) CUDA-graph inference in the product API, (f4) loss / metric geometry — each
against the reference's own functions (vendored copy) or the eager restatement in model.py / train.py.
Tolerances (max-abs error / max-abs reference): 1e-5 — float sums are re-associated, nothing here feeds coordinates."""
import numpy as np
import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def eager_upsample(flow, mask):
    import torch.nn.functional as F
    B, _, h, w = flow.shape
    m = torch.softmax(mask.view(B, 1, 9, 8, 8, h, w), dim=2)
    nb = F.unfold(8 * flow, [3, 3], padding=1).view(B, 2, 9, 1, 1, h, w)
    return torch.sum(m * nb, dim=2).permute(0, 1, 4, 2, 5, 3).reshape(B, 2, 8 * h, 8 * w)


@pytest.mark.parametrize("B,h,w", [(1, 64, 128), (2, 16, 32), (1, 5, 7)])
def test_convex_upsample(B, h, w):
    from prior_flow_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(h)
    flow = torch.randn(B, 2, h, w, device="cuda", generator=g) * 4
    mask = torch.randn(B, 576, h, w, device="cuda", generator=g) * 2
    want = eager_upsample(flow, mask)
    assert rel(ops.convex_upsample(flow, mask), want) < 1e-5
    assert rel(ops.convex_upsample(flow, mask.contiguous(memory_format=torch.channels_last)), want) < 1e-5
    if ref_shim.available():
        ref = ref_shim.load()
        model = ref_shim.make_model(ref)
        assert rel(ops.convex_upsample(flow, mask), model.upsample_flow(flow, mask)) < 1e-5


@pytest.mark.parametrize("B,h,w", [(1, 64, 128), (2, 16, 32), (1, 5, 7)])
@pytest.mark.parametrize("channels_last", [False, True])
def test_convex_upsample_gradients(B, h, w, channels_last):
    """pf_convex_upsample_bwd against autograd of the eager softmax / unfold / mul / sum chain (core/prior_raft.py:58-67)."""
    from prior_flow_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(7 + h)
    flow0 = torch.randn(B, 2, h, w, device="cuda", generator=g) * 4
    mask0 = torch.randn(B, 576, h, w, device="cuda", generator=g) * 2
    up = torch.randn(B, 2, 8 * h, 8 * w, device="cuda", generator=g)
    res = []
    for ours in (False, True):
        flow = flow0.clone().requires_grad_(True)
        mask = mask0.clone()
        if channels_last:
            mask = mask.contiguous(memory_format=torch.channels_last)
        mask.requires_grad_(True)
        out = ops.convex_upsample(flow, mask) if ours else eager_upsample(flow, mask)
        (out * up).sum().backward()
        res.append((out.detach(), flow.grad.clone(), mask.grad.clone()))
    (o_a, df_a, dm_a), (o_b, df_b, dm_b) = res
    assert rel(o_b, o_a) < 1e-5
    assert rel(df_b, df_a) < 1e-5 and rel(dm_b, dm_a) < 1e-5
    if channels_last:
        assert dm_b.is_contiguous(memory_format=torch.channels_last)


def test_uniform_loss_terms_and_gradients():
    from prior_flow_b200 import train as T
    g = torch.Generator(device="cuda").manual_seed(3)
    B, H, W = 2, 64, 128
    gt = torch.randn(B, 2, H, W, device="cuda", generator=g) * 6
    gt[0, :, 5, 7] = 500.0                                     # beyond MAX_FLOW: excluded
    valid = (torch.rand(B, H, W, device="cuda", generator=g) > 0.2).float()
    preds = [(gt + torch.randn(B, 2, H, W, device="cuda", generator=g) * (3 - 0.5 * i)).requires_grad_(True) for i in range(4)]
    wts = T.latitude_weights(H, W, "cuda")
    loss, _ = T.sequence_loss(preds, gt, valid, wts)
    loss.backward()
    got = [p.grad.clone() for p in preds]
    # eager restatement (train_flow.py:60-69)
    mag = torch.sum(gt ** 2, dim=1).sqrt()
    ok = ((valid >= 0.5) & (mag < 400.0)).float()
    ref_preds = [p.detach().clone().requires_grad_(True) for p in preds]
    ref_loss = sum(0.8 ** (4 - i - 1) * torch.sum(ok * wts * torch.sum((p - gt).abs(), dim=1)) for i, p in enumerate(ref_preds))
    ref_loss.backward()
    assert abs(float(loss) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
    for a, p in zip(got, ref_preds):
        assert torch.equal(a, p.grad)
    if ref_shim.available():
        ref = ref_shim.load()
        import importlib
        sph = importlib.import_module("core.utils.spherical")
        mask = torch.from_numpy(sph.spherical_mask(H, W)).cuda()[None]
        assert rel(wts.expand(1, H, W), mask) < 1e-6


def test_great_circle_distance_matches_reference():
    from prior_flow_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    B, H, W = 2, 64, 128
    gt = torch.randn(B, 2, H, W, device="cuda", generator=g) * 8
    pred = gt + torch.randn(B, 2, H, W, device="cuda", generator=g) * 2
    got = ops.great_circle_distance(pred, gt)
    assert got.shape == (B, H, W) and torch.isfinite(got).all()
    if ref_shim.available():
        ref_shim.load()
        import importlib
        sph = importlib.import_module("core.utils.spherical")
        want = sph.calculate_great_circle_distance(pred, gt)
        # asin(sqrt(h)) near h = 0 amplifies rounding: compare absolutely, against the largest distance
        assert float((got - want).abs().max()) < 1e-5 * float(want.abs().max()) + 1e-6
    assert float(ops.great_circle_distance(gt, gt).abs().max()) < 1e-6


def test_graphed_forward_matches_eager_api():
    from prior_flow_b200.model import PriOrRAFT
    import cases
    torch.manual_seed(0)
    m = PriOrRAFT().cuda().eval()
    im1, im2 = (torch.from_numpy(x).cuda() for x in cases.e2e_images())
    with torch.no_grad():
        want = m(im1, im2, iters=4, test_mode=True)
        want_swapped = m(im2, im1, iters=4, test_mode=True)
    run = m.graphed(iters=4)
    got = run(im1, im2).clone()
    got_swapped = run(im2.cpu().pin_memory(), im1.cpu().pin_memory()).clone()     # second call: replay only, host inputs
    epe = lambda a, b: float(torch.sqrt(((a.double() - b.double()) ** 2).sum(1)).mean())
    assert epe(got, want) < 1e-4 and epe(got_swapped, want_swapped) < 1e-4
    assert len(run._graphs) == 1
