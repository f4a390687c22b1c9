"""Seeded synthetic inputs shared by `make_golden.py` (which runs the unmodified reference on them)
and by the tests (which run the oracle / the CUDA path on the same arrays).

numpy's legacy `RandomState` streams are stable across platforms and numpy versions, so only the
reference's *outputs* are stored in the .npz fixtures; the inputs are regenerated from the seed.
Statistics follow SURVEY.md §8(d): fmap ~ N(0, 1.45^2); coords = coords_grid + N(0, 5^2).
"""
import numpy as np

F = np.float32
FH, FW = 16, 32          # 1/8-resolution feature map of a 128x256 ERP image (smallest sane size)
C = 256


def fmaps(seed, B=1, c=C, h=FH, w=FW, n=4):
    rs = np.random.RandomState(seed)
    return [(rs.randn(B, c, h, w) * 1.45).astype(F) for _ in range(n)]


def base_grid(B, h, w):
    ys, xs = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    return np.repeat(np.stack([xs, ys], 0).astype(F)[None], B, 0)


def coords(seed, B=1, h=FH, w=FW, sigma=5.0):
    rs = np.random.RandomState(seed)
    return (base_grid(B, h, w) + rs.randn(B, 2, h, w) * sigma).astype(F)


def flow(seed, B=2, h=FH, w=FW, sigma=5.0):
    rs = np.random.RandomState(seed)
    return (rs.randn(B, 2, h, w) * sigma).astype(F)


def edge_coords(h=FH, w=FW):
    """Edge vectors of SURVEY.md §4 laid out as a [1,2,h,w] coords tensor: x exactly W-1, W-0.5,
    tiny negative (-1e-8 % W == W), far outside, and y in {-1,-0.5,H-1,H-0.5,H}; rest = identity."""
    c = base_grid(1, h, w)
    xs = [w - 1.0, w - 0.5, -1e-8, -0.25, w + 3.75, -w - 2.5, 0.0, 1e-4, w - 1e-3, 2 * w + 0.5, 5.5, 7.25]
    ys = [-1.0, -0.5, h - 1.0, h - 0.5, float(h), -3.5, 0.0, h + 6.0, 0.5, 3.75, h - 1.5, 1e-3]
    k = 0
    for yv in ys:
        for xv in xs:
            i, j = divmod(k, w)
            if i >= h:
                break
            c[0, 0, i, j] = xv
            c[0, 1, i, j] = yv
            k += 1
    return c.astype(F)


def image(seed, B=1, H=64, W=128, ch=6):
    rs = np.random.RandomState(seed)
    return (rs.rand(B, ch, H, W) * 2 - 1).astype(F)


def small_sampler_case(seed=11):
    rs = np.random.RandomState(seed)
    img = rs.randn(2, 5, 6, 8).astype(F)
    pts = np.stack([rs.uniform(-3, 11, size=(2, 7, 9)), rs.uniform(-2, 8, size=(2, 7, 9))], -1).astype(F)
    # plant exact edge values
    pts[0, 0, :6, 0] = [7.0, 7.5, -1e-8, 8.0, 0.0, -0.5]
    pts[0, 1, :6, 1] = [-1.0, -0.5, 5.0, 5.5, 6.0, 0.0]
    return img, pts


def seeded_state_dict(template):
    """Deterministic parameters for the end-to-end golden: every tensor of `template` (a state_dict: name -> tensor
    with the reference's names/shapes) is filled from a numpy stream seeded by crc32(name), so the reference model
    (build container, CPU) and prior_flow_b200.model.PriOrRAFT (GPU box) get bit-identical weights without
    shipping a 33 MB checkpoint.  Conv weights ~ N(0, 1/(3 fan_in)) — the variance of PyTorch's default conv init,
    which keeps the recurrent update well conditioned (a 2/fan_in init makes the random network chaotic: a 1e-6
    perturbation of the volume grows to 2 px after 12 iterations, measured); biases ~ N(0, 0.01^2); norm weights ~ 1."""
    import zlib
    import torch
    out = {}
    for name in sorted(template):
        t = template[name]
        rs = np.random.RandomState(zlib.crc32(name.encode()) & 0x7FFFFFFF)
        shape = tuple(t.shape)
        if name.endswith("num_batches_tracked"):
            out[name] = torch.zeros(shape, dtype=t.dtype)
        elif name.endswith("running_mean"):
            out[name] = torch.from_numpy((rs.randn(*shape) * 0.05).astype(F))
        elif name.endswith("running_var"):
            out[name] = torch.from_numpy((1.0 + 0.1 * rs.rand(*shape)).astype(F))
        elif len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            out[name] = torch.from_numpy((rs.randn(*shape) * np.sqrt(1.0 / (3.0 * fan_in))).astype(F))
        elif name.endswith("weight"):
            out[name] = torch.from_numpy((1.0 + 0.05 * rs.randn(*shape)).astype(F))
        else:
            out[name] = torch.from_numpy((0.01 * rs.randn(*shape)).astype(F))
    return out


def e2e_images(seed=8, H=128, W=256):
    rs = np.random.RandomState(seed)
    return (rs.rand(1, 3, H, W) * 255).astype(F), (rs.rand(1, 3, H, W) * 255).astype(F)
