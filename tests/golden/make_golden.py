"""Generates tests/golden/*.npz by running the UNMODIFIED reference on CPU (build container only).

    python tests/golden/make_golden.py            # needs /root/reference (read-only mount)

The reference has no tests or golden vectors of its own, so these files are how the
oracle (`oracle/np_oracle.py`, `oracle/torch_oracle.py`) and through it the CUDA path are pinned to
the reference's behaviour.  Inputs come from `cases.py` (seeded numpy streams) and are not stored;
each .npz holds only what the reference returned, as fp32.  Reference entry points exercised
(paths under /root/reference/PriOr-RAFT):
  core/prior_raft.py:69-83   PriOr_RAFT.corr, groupwise_corr      core/corr.py:13-61,94-144 CorrBlock, DCCL
  core/utils/utils.py:61-101 samplers, coords_grid                 core/utils/my_cycle_sample.py:6-97
  core/utils/projection_prim_ortho.py:23-48,432-443,507-546       rotation matrix, sample grids, img/flo_rotate
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import cases  # noqa: E402
from oracle import ref_shim  # noqa: E402


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def N(t):
    return t.detach().cpu().numpy().astype(np.float32)


def main():
    ref = ref_shim.load()
    torch.set_grad_enabled(False)
    ppo, utils, corr_mod = ref.ppo, ref.utils, ref.corr
    model_cls = ref.prior_raft.PriOr_RAFT
    save = lambda name, **kw: np.savez_compressed(os.path.join(HERE, name), **kw)

    # ---- geometry: rotation matrices + the 8 sample grids of prior_raft.py:115-125 at 128x256
    H, W = cases.FH * 8, cases.FW * 8
    R_a2b = ppo.generate_rotation_metrix(theta_list=[0., 0., -np.pi / 2])
    R_b2a = ppo.generate_rotation_metrix(theta_list=[0., 0., np.pi / 2])
    R_gen = ppo.generate_rotation_metrix(theta_list=[0.3, -0.7, 1.1])
    grids = {}
    for tag, R in (("a2b", R_a2b), ("b2a", R_b2a)):
        grids[f"{tag}"] = ppo.generate_samplegrid([1, 3, H, W], R)
        grids[f"{tag}_8x"] = ppo.generate_samplegrid([1, 3, H // 8, W // 8], R)
        grids[f"{tag}_w2c"] = ppo.generate_samplegrid([1, 3, H, W], R.T)
        grids[f"{tag}_w2c_8x"] = ppo.generate_samplegrid([1, 3, H // 8, W // 8], R.T)
    grids["gen_24x40"] = ppo.generate_samplegrid([2, 3, 24, 40], R_gen)
    grids["a2b_64x128"] = ppo.generate_samplegrid([1, 3, 64, 128], R_a2b)
    save("geometry.npz", R_a2b=N(R_a2b), R_b2a=N(R_b2a), R_gen=N(R_gen), **{k: N(v) for k, v in grids.items()})

    # ---- volume + pyramid (prior_raft.py:69-75, corr.py:99-111)
    f1a, f2a, f1b, f2b = (T(x) for x in cases.fmaps(seed=1))
    corr = lambda a, b: model_cls.corr(None, a, b)
    vol_a, vol_b = corr(f1a, f2a), corr(f1b, f2b)
    dccl = corr_mod.DCCL(num_levels=4, radius=4)
    pyr_a, pyr_b = dccl.build_pyramid(vol_a), dccl.build_pyramid(vol_b)
    save("volume.npz", **{f"a{l}": N(p) for l, p in enumerate(pyr_a)}, **{f"b{l}": N(p) for l, p in enumerate(pyr_b)})

    # ---- DCCL dual lookup (corr.py:113-144), random coords and the edge vectors
    g_w2c, g_c2w = grids["a2b_w2c_8x"], grids["b2a_8x"]
    for tag, c in (("rand", cases.coords(seed=2)), ("edge", cases.edge_coords())):
        # the reference mutates nothing here, but pass clones to be safe
        oa, ob = dccl(T(c), pyr_a, pyr_b, g_w2c.clone(), g_c2w.clone())
        save(f"dccl_{tag}.npz", out_own=N(oa), out_other=N(ob))
    # second call direction of prior_raft.py:186
    oa, ob = dccl(T(cases.coords(seed=3)), pyr_b, pyr_a, grids["b2a_w2c_8x"].clone(), grids["a2b_8x"].clone())
    save("dccl_rand_b.npz", out_own=N(oa), out_other=N(ob))

    # ---- CorrBlock (corr.py:13-61)
    cb = corr_mod.CorrBlock(f1a, f2a, num_levels=4, radius=4)
    save("corrblock.npz", out=N(cb(T(cases.coords(seed=2)))), out_edge=N(cb(T(cases.edge_coords()))))

    # ---- plain samplers (utils.py:61-95)
    img, pts = cases.small_sampler_case()
    save("samplers.npz", cyc=N(utils.cycle_bilinear_sampler(T(img), T(pts))),
         lin=N(utils.bilinear_sampler(T(img), T(pts))), coords_grid=N(utils.coords_grid(2, 5, 7, device="cpu")))

    # ---- feature warp + group-wise correlation (prior_raft.py:77-83,173-174)
    c = T(cases.coords(seed=4))
    warped = utils.cycle_bilinear_sampler(f2a, c.permute(0, 2, 3, 1))
    save("warp_groupcorr.npz", flaw=N(model_cls.groupwise_corr(None, f1a, warped, num_groups=4)))

    # ---- img_rotate (projection_prim_ortho.py:507-514) on a 6-channel 64x128 image
    im = T(cases.image(seed=5))
    save("img_rotate.npz", out=N(ppo.img_rotate(im, sample_grid=grids["a2b_64x128"].clone())))

    # ---- flo_rotate (projection_prim_ortho.py:531-546) incl. cycle_grid_sample pieces
    fl = T(cases.flow(seed=6))
    gw = grids["b2a_w2c_8x"].repeat(2, 1, 1, 1)
    gc = grids["b2a_8x"].repeat(2, 1, 1, 1)
    out = ppo.flo_rotate(fl.clone(), sample_grid_W2C=gw.clone(), sample_grid_C2W=gc.clone())
    big = T(cases.flow(seed=7, sigma=60.0))   # large flows: wrap + clamp paths
    out_big = ppo.flo_rotate(big.clone(), sample_grid_W2C=gw.clone(), sample_grid_C2W=gc.clone())
    end_w = ppo.flow2endpoint(ppo.generate_plane_grid(fl.shape), fl, stack=False)
    end_c = ref.mcs.cycle_grid_sample(gw.clone(), end_w.clone(), is_grid=True)
    save("flo_rotate.npz", out=N(out), out_big=N(out_big), end_w=N(end_w), end_c=N(end_c))
    # ---- end to end: the reference model (prior_raft.py:107-215) with seeded weights, 128x256, 4 iterations
    from argparse import Namespace
    model = model_cls(Namespace(mixed_precision=False, dropout=0.0)).eval()
    model.load_state_dict(cases.seeded_state_dict(model.state_dict()), strict=True)
    im1, im2 = (T(x) for x in cases.e2e_images())
    flow4 = model(im1, im2, iters=4, test_mode=True)
    flow12 = model(im1, im2, iters=12, test_mode=True)
    init = T(cases.flow(seed=9, B=1, sigma=2.0))
    flow_init = model(im1, im2, iters=2, init_flow=init, test_mode=True)
    model.train()  # BatchNorm in cnet uses batch statistics; predictions for every iteration, both views
    model.freeze_bn()
    pa, pb = model(im1, im2, iters=2)
    save("e2e.npz", flow4=N(flow4), flow12=N(flow12), flow_init=N(flow_init), train_A1=N(pa[1]), train_B1=N(pb[1]))
    print("golden files written to", HERE)


if __name__ == "__main__":
    main()
