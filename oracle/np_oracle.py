"""TEST INFRASTRUCTURE ONLY — numpy fp32 restatement of PriOr-RAFT's correlation hot path.

This module is the *checker* for the CUDA path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` leg may import it; the
product (`prior_flow_b200/`) never does and fails loudly without its CUDA
library.

Every function restates one reference function with explicit, individually
rounded fp32 arithmetic (numpy float32 arrays round after every op, exactly
like a chain of eager ATen kernels) and cites the reference lines it follows
(paths relative to `/root/reference/PriOr-RAFT/`).  `F.grid_sample` is written
out (ATen `grid_sampler_2d`, bilinear / zeros padding / align_corners=True), it
is not called.

Parity pinning: the reference ships no tests or golden vectors ("parity
unpinned" upstream).  This oracle is pinned instead against
outputs of the unmodified reference executed in the build container on CPU
(`tests/golden/make_golden.py` → `tests/golden/*.npz`,
`tests/test_oracle_golden.py`).

Two device-dependent details of ATen are modelled explicitly because the
coordinate path must be bit-exact:

* `div_mode` — `tensor / python_scalar` is a true IEEE division on CPU but
  `tensor * (1.0f / scalar)` in ATen's CUDA kernel (`div_true_kernel_cuda`
  takes the reciprocal of a CPU scalar once).  "ieee" reproduces the CPU
  reference (what the golden files hold), "aten_cuda" the reference as it
  runs on a GPU.
* `acc` — ATen's `out_acc += v * w` over the taps nw, ne, sw, se is contracted into an FMA
  chain by the compiler, in the CUDA kernel and (measured: bit-exact on the golden files) in
  the x86 CPU build as well.  "fma" (default) emulates that chain through float64 (exact up
  to rare double rounding); "unfused" rounds every mul and add separately.
"""
from __future__ import annotations

import numpy as np

F = np.float32
_PI = F(np.pi)          # python float -> fp32 at each tensor-scalar op
_TWO_PI = F(2 * np.pi)  # `2 * np.pi` is evaluated in double first, then cast


def _f(x):
    return np.ascontiguousarray(x, dtype=np.float32)


# --------------------------------------------------------------------------- scalar-op helpers
def remainder(x, m):
    """torch.remainder(x, m) for floats: fmod, then `+ m` if the sign differs (may return m itself)."""
    m = F(m)
    r = np.fmod(x, m).astype(F)
    fix = (r != 0) & ((r < 0) != (m < 0))
    return np.where(fix, r + m, r).astype(F)


def div_scalar(a, s, div_mode="ieee"):
    """`tensor / python_scalar` (see module docstring)."""
    s = F(s)
    if div_mode == "ieee":
        with np.errstate(divide="ignore", invalid="ignore"):
            return (a / s).astype(F)
    if div_mode == "aten_cuda":
        with np.errstate(divide="ignore", invalid="ignore"):
            return (a * (F(1.0) / s)).astype(F)
    raise ValueError(div_mode)


def _normalise(p, size, div_mode):
    """core/utils/utils.py:85-86 — `2*p/(size-1) - 1`."""
    return (div_scalar(F(2) * p, size - 1, div_mode) - F(1)).astype(F)


def _unnormalise(g, size):
    """ATen grid_sampler_unnormalize(align_corners=True): ((g + 1) / 2) * (size - 1),
    followed by safe_downgrade_to_int_range (non-finite / huge -> -100)."""
    with np.errstate(invalid="ignore"):
        v = (((g + F(1)) / F(2)) * F(size - 1)).astype(F)
        bad = ~np.isfinite(v) | (v > F(2147483646.0)) | (v < F(-2147483648.0))
    return np.where(bad, F(-100.0), v).astype(F)


def pixel_to_sample_coords(px, py, H, W, cyclic, div_mode="ieee"):
    """Pixel coords -> the unnormalised (ix, iy) ATen finally samples at.

    cyclic=True  : cycle_bilinear_sampler (core/utils/utils.py:78-95) and the module-local
                   sampler used by img_rotate (core/utils/projection_prim_ortho.py:119-135)
    cyclic=False : bilinear_sampler (core/utils/utils.py:61-75)
    `ix` is NOT always equal to `px` — the normalise/unnormalise round trip changes
    low bits, so it is restated step by step.
    """
    px = _f(px)
    py = _f(py)
    if cyclic:
        px = remainder(px, W)
    gx = _normalise(px, W, div_mode)
    gy = _normalise(py, H, div_mode)
    return _unnormalise(gx, W), _unnormalise(gy, H)


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def bilinear_zeros(img, ix, iy, acc="fma"):
    """ATen grid_sampler_2d forward, bilinear, zeros padding, on unnormalised coords.

    img [P,C,H,W]; ix, iy [P,*S]  ->  [P,C,*S].
    Taps outside [0,W-1]x[0,H-1] are dropped — x is NOT wrapped here, so a sample at
    x in (W-1, W) blends column W-1 with zero.
    """
    img = _f(img)
    P, C, H, W = img.shape
    S = ix.shape[1:]
    ix = _f(ix).reshape(P, -1)
    iy = _f(iy).reshape(P, -1)
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    x1 = x0 + F(1)
    y1 = y0 + F(1)
    # ATen grid_sampler_2d: nw = (ix_se - ix) * (iy_se - iy), ne = (ix - ix_sw) * (iy_sw - iy), ...
    w_nw = (x1 - ix) * (y1 - iy)
    w_ne = (ix - x0) * (y1 - iy)
    w_sw = (x1 - ix) * (iy - y0)
    w_se = (ix - x0) * (iy - y0)
    flat = img.reshape(P, C, H * W)

    def tap(xi, yi):
        ok = (xi >= 0) & (xi <= W - 1) & (yi >= 0) & (yi <= H - 1)
        xi_c = np.clip(xi, 0, W - 1).astype(np.int64)
        yi_c = np.clip(yi, 0, H - 1).astype(np.int64)
        idx = np.broadcast_to((yi_c * W + xi_c)[:, None, :], (P, C, xi.shape[1]))
        v = np.take_along_axis(flat, idx, axis=2)
        return np.where(ok[:, None, :], v, F(0)).astype(F)

    v_nw, v_ne, v_sw, v_se = tap(x0, y0), tap(x1, y0), tap(x0, y1), tap(x1, y1)
    w_nw, w_ne, w_sw, w_se = (w[:, None, :].astype(F) for w in (w_nw, w_ne, w_sw, w_se))
    if acc == "unfused":
        out = ((v_nw * w_nw + v_ne * w_ne) + v_sw * w_sw) + v_se * w_se
    elif acc == "fma":
        out = _fma(v_se, np.broadcast_to(w_se, v_se.shape),
                   _fma(v_sw, np.broadcast_to(w_sw, v_sw.shape),
                        _fma(v_ne, np.broadcast_to(w_ne, v_ne.shape),
                             (v_nw * w_nw).astype(F))))
    else:
        raise ValueError(acc)
    return out.astype(F).reshape((P, C) + tuple(S))


# --------------------------------------------------------------------------- samplers
def cycle_bilinear_sampler(img, coords, div_mode="ieee", acc="fma"):
    """core/utils/utils.py:78-95.  img [B,C,H,W], coords [B,Ho,Wo,2] (x,y) pixels -> [B,C,Ho,Wo]."""
    H, W = img.shape[-2:]
    ix, iy = pixel_to_sample_coords(coords[..., 0], coords[..., 1], H, W, True, div_mode)
    return bilinear_zeros(img, ix, iy, acc)


def bilinear_sampler(img, coords, div_mode="ieee", acc="fma"):
    """core/utils/utils.py:61-75 (no x wrap)."""
    H, W = img.shape[-2:]
    ix, iy = pixel_to_sample_coords(coords[..., 0], coords[..., 1], H, W, False, div_mode)
    return bilinear_zeros(img, ix, iy, acc)


def coords_grid(batch, ht, wd):
    """core/utils/utils.py:98-101 — [B,2,ht,wd], channel 0 = x, channel 1 = y."""
    ys, xs = np.meshgrid(np.arange(ht), np.arange(wd), indexing="ij")
    g = np.stack([xs, ys], axis=0).astype(F)
    return np.repeat(g[None], batch, axis=0)


# --------------------------------------------------------------------------- volume + pyramid
def corr_volume(fmap1, fmap2):
    """core/prior_raft.py:69-75 — V[b,n,m] = sum_c f1[b,c,n] f2[b,c,m] / sqrt(C)  -> [B,h,w,h,w]."""
    fmap1, fmap2 = _f(fmap1), _f(fmap2)
    B, C, h, w = fmap1.shape
    a = fmap1.reshape(B, C, h * w).transpose(0, 2, 1)
    v = np.matmul(a, fmap2.reshape(B, C, h * w)).astype(F)
    return (v / np.sqrt(F(C))).astype(F).reshape(B, h, w, h, w)


def avg_pool2x2(x):
    """F.avg_pool2d(x, 2, stride=2): sum in window row-major order, then / 4; odd tails dropped."""
    H2, W2 = x.shape[-2] // 2, x.shape[-1] // 2
    x = x[..., : 2 * H2, : 2 * W2]
    a = x[..., 0::2, 0::2]
    b = x[..., 0::2, 1::2]
    c = x[..., 1::2, 0::2]
    d = x[..., 1::2, 1::2]
    return ((((a + b) + c) + d) / F(4)).astype(F)


def build_pyramid(volume, num_levels=4):
    """core/corr.py:99-111 — list of [B*h*w, 1, h/2^l, w/2^l]."""
    B, h, w, h2, w2 = volume.shape
    lvl = _f(volume).reshape(B * h * w, 1, h2, w2)
    out = [lvl]
    for _ in range(num_levels - 1):
        lvl = avg_pool2x2(lvl)
        out.append(lvl)
    return out


# --------------------------------------------------------------------------- DCCL dual lookup
def window_points(coords, level, radius=4):
    """core/corr.py:120-126 — sample points of the (2r+1)^2 window at one level.

    coords [B,2,h,w] -> px, py [B*h*w, 2r+1, 2r+1]; index [n,a,b] samples at
    (cx/2^l + (a-r), cy/2^l + (b-r))  — the window is x-major."""
    B, _, h, w = coords.shape
    c = _f(coords).transpose(0, 2, 3, 1).reshape(B * h * w, 2)
    c = (c / F(2 ** level)).astype(F)
    d = np.arange(-radius, radius + 1).astype(F)
    k = 2 * radius + 1
    px = np.broadcast_to(c[:, 0, None, None] + d[None, :, None], (B * h * w, k, k)).astype(F)
    py = np.broadcast_to(c[:, 1, None, None] + d[None, None, :], (B * h * w, k, k)).astype(F)
    return px, py


def dccl_lookup(coords, pyr_own, pyr_other, grid_w2c, grid_c2w, radius=4,
                div_mode="ieee", acc="fma", return_debug=False):
    """core/corr.py:113-144 — DCCL.__call__.

    coords [B,2,h,w]; pyramids as from build_pyramid; grid_w2c / grid_c2w [B,2,h,w] (the `_8x`
    sample grids).  Returns (out_own, out_other), each [B, L*(2r+1)^2, h, w] fp32 with channel
    l*81 + a*9 + b.  With return_debug also returns per-level dicts of the unnormalised sample
    coordinates (own: ix,iy ; other: ix,iy of the mapped point and the raw pre-rotation map)."""
    B, _, h, w = coords.shape
    k = 2 * radius + 1
    N = B * h * w
    grid_w2c, grid_c2w = _f(grid_w2c), _f(grid_c2w)
    outs_own, outs_other, dbg = [], [], []
    for lvl in range(len(pyr_own)):
        own, other = _f(pyr_own[lvl]), _f(pyr_other[lvl])
        Hl, Wl = own.shape[-2:]
        px, py = window_points(coords, lvl, radius)
        # own-view branch (corr.py:128-130)
        ix, iy = pixel_to_sample_coords(px, py, Hl, Wl, True, div_mode)
        o = bilinear_zeros(own, ix, iy, acc)                      # [N,1,k,k]
        outs_own.append(o.reshape(B, h, w, k * k))
        # orthogonal branch: map the window through the LEVEL-0 rotation grid (corr.py:132-133)
        gx, gy = pixel_to_sample_coords(px.reshape(B, h * w * k * k), py.reshape(B, h * w * k * k),
                                        h, w, True, div_mode)
        mapped = bilinear_zeros(grid_w2c, gx, gy, acc)           # [B,2,h*w*k*k]
        qx = mapped[:, 0].reshape(N, k, k)
        qy = mapped[:, 1].reshape(N, k, k)
        # ... and index the LEVEL-l volume of the other view with it (corr.py:135-136)
        jx, jy = pixel_to_sample_coords(qx, qy, Hl, Wl, True, div_mode)
        raw = bilinear_zeros(other, jx, jy, acc).reshape(B, h, w, k * k).transpose(0, 3, 1, 2)
        # rotate the [B,81,h,w] map back (corr.py:137-138 -> projection_prim_ortho.py:507-514)
        rot = img_rotate(raw, grid_c2w, div_mode, acc)
        outs_other.append(rot.transpose(0, 2, 3, 1))
        if return_debug:
            dbg.append(dict(own_ix=ix, own_iy=iy, map_ix=gx.reshape(N, k, k), map_iy=gy.reshape(N, k, k),
                            other_px=qx, other_py=qy, other_ix=jx, other_iy=jy, other_raw=raw))
    out_own = np.ascontiguousarray(np.concatenate(outs_own, -1).transpose(0, 3, 1, 2), dtype=F)
    out_other = np.ascontiguousarray(np.concatenate(outs_other, -1).transpose(0, 3, 1, 2), dtype=F)
    if return_debug:
        return out_own, out_other, dbg
    return out_own, out_other


def corrblock_lookup(coords, pyramid, radius=4, div_mode="ieee", acc="fma"):
    """core/corr.py:30-51 — CorrBlock.__call__ (non-cyclic sampler, single view)."""
    B, _, h, w = coords.shape
    k = 2 * radius + 1
    outs = []
    for lvl, vol in enumerate(pyramid):
        Hl, Wl = vol.shape[-2:]
        px, py = window_points(coords, lvl, radius)
        ix, iy = pixel_to_sample_coords(px, py, Hl, Wl, False, div_mode)
        outs.append(bilinear_zeros(_f(vol), ix, iy, acc).reshape(B, h, w, k * k))
    return np.ascontiguousarray(np.concatenate(outs, -1).transpose(0, 3, 1, 2), dtype=F)


def alt_corr_lookup(fmap1, fmap2, coords, num_levels=4, radius=4, div_mode="ieee", acc="fma"):
    """Memory-efficient lookup with no volume.  The reference's `alt_cuda_corr` extension is not
    shipped (core/corr.py:7-11,64-91), so the oracle for it is the identity the survey derives:
    avg-pool is linear, hence lookup(pool_l(V)) == lookup(corr(f1, pool_l(f2))) up to rounding.
    Uses the cyclic sampler (the DCCL flavour)."""
    f2 = _f(fmap2)
    B, C, h, w = fmap1.shape
    k = 2 * radius + 1
    outs = []
    for lvl in range(num_levels):
        vol = corr_volume_rect(_f(fmap1), f2)
        Hl, Wl = f2.shape[-2:]
        px, py = window_points(coords, lvl, radius)
        ix, iy = pixel_to_sample_coords(px, py, Hl, Wl, True, div_mode)
        outs.append(bilinear_zeros(vol.reshape(B * h * w, 1, Hl, Wl), ix, iy, acc).reshape(B, h, w, k * k))
        f2 = avg_pool2x2(f2)
    return np.ascontiguousarray(np.concatenate(outs, -1).transpose(0, 3, 1, 2), dtype=F)


def corr_volume_rect(fmap1, fmap2):
    """corr_volume with different spatial sizes for the two maps -> [B,h1,w1,h2,w2]."""
    B, C, h1, w1 = fmap1.shape
    _, _, h2, w2 = fmap2.shape
    v = np.matmul(fmap1.reshape(B, C, -1).transpose(0, 2, 1), fmap2.reshape(B, C, -1)).astype(F)
    return (v / np.sqrt(F(C))).astype(F).reshape(B, h1, w1, h2, w2)


# --------------------------------------------------------------------------- feature warp + group corr
def groupwise_corr(fea1, fea2, num_groups=4):
    """core/prior_raft.py:77-83."""
    B, C, H, W = fea1.shape
    prod = (_f(fea1) * _f(fea2)).reshape(B, num_groups, C // num_groups, H, W)
    return prod.mean(axis=2, dtype=np.float32).astype(F)


def warp_groupcorr(fmap1, fmap2, coords, num_groups=4, div_mode="ieee", acc="fma"):
    """core/prior_raft.py:173-174 — groupwise_corr(fmap1, cycle_bilinear_sampler(fmap2, coords^T))."""
    warped = cycle_bilinear_sampler(fmap2, _f(coords).transpose(0, 2, 3, 1), div_mode, acc)
    return groupwise_corr(fmap1, warped, num_groups)


# --------------------------------------------------------------------------- ERP geometry
def generate_rotation_matrix(theta_list=None, axis_list=None):
    """core/utils/projection_prim_ortho.py:23-48 — R = prod of axis rotations, fp32 cos/sin of
    fp32(theta).  cos/sin are taken in float64 and rounded (a correctly rounded libm)."""
    axis_list = ["z", "y", "x"] if axis_list is None else axis_list
    theta_list = [0.0, 0.0, 0.0] if theta_list is None else theta_list
    R = np.eye(3, dtype=F)
    for axis, theta in zip(axis_list, theta_list):
        t = np.float64(F(theta))
        c, s = F(np.cos(t)), F(np.sin(t))
        if axis == "x":
            M = np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=F)
        elif axis == "y":
            M = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=F)
        else:
            M = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=F)
        R = (R @ M).astype(F)
    return R


def _trig(fn, *a):
    return fn(*[np.asarray(v, dtype=np.float64) for v in a]).astype(F)


def _diverge_zero(x, eps=1e-6):
    """projection_prim_ortho.py:69-74."""
    near = (np.abs(x) < F(eps)).astype(F)
    return (x + (np.sign(x).astype(F) * near) * F(eps)).astype(F)


def generate_samplegrid(size, R, div_mode="ieee"):
    """projection_prim_ortho.py:432-443 (+ :10-20, :397-411, :77-89, :247-261, :51-74, :413-429).

    size = (B, _, H, W); R [3,3] fp32.  Returns [B,2,H,W] fp32: for every output pixel (m,n) the
    source pixel (m',n') under rotation R.  SURVEY.md §A.4."""
    B, _, H, W = size
    R = _f(R)
    m = np.broadcast_to(np.arange(W, dtype=F)[None, :], (H, W))
    n = np.broadcast_to(np.arange(H, dtype=F)[:, None], (H, W))
    u = div_scalar(m + F(0.5), W, div_mode)
    theta = ((u - F(0.5)) * F(2)) * _PI
    v = div_scalar(n + F(0.5), H, div_mode)
    phi = (F(0.5) - v) * _PI
    cphi = _trig(np.cos, phi)
    x = cphi * _trig(np.cos, theta)
    y = cphi * _trig(np.sin, theta)
    z = _trig(np.sin, phi)
    xyz = np.stack([x, y, z], 0).astype(F)
    rot = np.einsum("ij,jhw->ihw", R.astype(np.float64), xyz.astype(np.float64)).astype(F)
    phi2 = _trig(np.arcsin, rot[2])
    theta2 = _trig(np.arctan2, _diverge_zero(rot[1]), _diverge_zero(rot[0]))
    u2 = div_scalar(theta2, _TWO_PI, div_mode) + F(0.5)
    m2 = u2 * F(W) - F(0.5)
    v2 = F(0.5) - div_scalar(phi2, _PI, div_mode)
    n2 = v2 * F(H) - F(0.5)
    g = np.stack([m2, n2], 0).astype(F)
    return np.ascontiguousarray(np.broadcast_to(g[None], (B, 2, H, W)))


def img_rotate(image, sample_grid, div_mode="ieee", acc="fma"):
    """projection_prim_ortho.py:507-514 -> module-local cyclic bilinear_sampler (:119-135).
    image [B,C,H,W], sample_grid [B,2,H,W] -> [B,C,H,W]."""
    return cycle_bilinear_sampler(image, _f(sample_grid).transpose(0, 2, 3, 1), div_mode, acc)


def flow2endpoint(flow):
    """projection_prim_ortho.py:200-218 (stack=False) with generate_plane_grid (:10-20)."""
    B, _, H, W = flow.shape
    start = coords_grid(B, H, W)
    end = (start + _f(flow)).astype(F)
    ex = remainder(end[:, 0] + F(0.5), W) - F(0.5)
    ey = np.clip(end[:, 1], F(-0.5), F(H - 0.5)).astype(F)
    return np.stack([ex, ey], 1).astype(F)


def cycle_grid_sample(src, grid, is_grid=False):
    """core/utils/my_cycle_sample.py:6-79 (+ adjust_sample_m :82-97).

    src [B,C,H,W]; grid [B,2,Hg,Wg] pixel coords.  True x wrap, y clamp, weights not
    renormalised; with is_grid the m-channel of taps b,c,d is re-centred to within +-W/2 of
    tap a.  (The reference also writes `grid.x % W` back through a view of the caller's tensor;
    the oracle leaves its input alone.)"""
    src, grid = _f(src), _f(grid)
    B, C, H, W = src.shape
    _, _, Hg, Wg = grid.shape
    g = grid.reshape(B, 2, -1)
    gx = remainder(g[:, 0], W)
    gy = g[:, 1]
    fx, fy = np.floor(gx), np.floor(gy)
    xw, yw = (gx - fx).astype(F), (gy - fy).astype(F)
    wa = (F(1) - xw) * (F(1) - yw)
    wb = (F(1) - xw) * yw
    wc = xw * (F(1) - yw)
    wd = xw * yw
    x0i = fx.astype(np.int64)
    y0i = fy.astype(np.int64)
    x0 = np.mod(x0i, W)
    x1 = np.mod(x0i + 1, W)
    y0 = np.clip(y0i, 0, H - 1)
    y1 = np.clip(y0i + 1, 0, H - 1)
    flat = src.reshape(B, C, H * W)

    def gather(yy, xx):
        idx = np.broadcast_to((yy * W + xx)[:, None, :], (B, C, yy.shape[1]))
        return np.take_along_axis(flat, idx, axis=2).astype(F)

    Ia, Ib, Ic, Id = gather(y0, x0), gather(y1, x0), gather(y0, x1), gather(y1, x1)
    if is_grid:
        half = F(W / 2)
        for I in (Ib, Ic, Id):
            t = remainder((I[:, 0] - Ia[:, 0]) + half, W)
            I[:, 0] = (Ia[:, 0] + t) - half
    out = ((wa[:, None] * Ia + wb[:, None] * Ib) + wc[:, None] * Ic) + wd[:, None] * Id
    return np.ascontiguousarray(out.astype(F).reshape(B, C, Hg, Wg))


def u_clip(u, W):
    """projection_prim_ortho.py:234-244."""
    half = F(W / 2)
    return (remainder(u + half, W) - half).astype(F)


def flo_rotate(flow, grid_w2c, grid_c2w):
    """projection_prim_ortho.py:531-546 with both sample grids given (the only form the model uses,
    core/prior_raft.py:165,179).  flow [B,2,H,W] -> [B,2,H,W].  SURVEY.md §A.5."""
    grid_w2c, grid_c2w = _f(grid_w2c), _f(grid_c2w)
    W = flow.shape[-1]
    end_w = flow2endpoint(flow)
    end_c = cycle_grid_sample(grid_w2c, end_w, is_grid=True)
    flow_c = (end_c - grid_w2c).astype(F)
    flow_c[:, 0] = u_clip(flow_c[:, 0], W)
    return cycle_grid_sample(flow_c, grid_c2w, is_grid=False)
