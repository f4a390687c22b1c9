"""CPU: the oracles (oracle/np_oracle.py, oracle/torch_oracle.py) against golden vectors produced by the
unmodified reference (tests/golden/make_golden.py).  Integer/coordinate paths are bit-exact; values whose
reduction order is the BLAS's / the reduction kernel's own are held to 1e-6 of max|ref|."""
import numpy as np
import pytest
import torch

import cases
from conftest import golden, rel_to_max
from oracle import np_oracle as O
from oracle import torch_oracle as TO

T = lambda a: torch.from_numpy(np.ascontiguousarray(a))


def test_rotation_matrices(geo):
    assert np.array_equal(O.generate_rotation_matrix([0., 0., -np.pi / 2]), geo["R_a2b"])
    assert np.array_equal(O.generate_rotation_matrix([0., 0., np.pi / 2]), geo["R_b2a"])
    assert np.abs(O.generate_rotation_matrix([0.3, -0.7, 1.1]) - geo["R_gen"]).max() < 1e-7
    assert np.array_equal(TO.rotation_matrix([0., 0., -np.pi / 2]).numpy(), geo["R_a2b"])
    assert np.array_equal(TO.rotation_matrix([0.3, -0.7, 1.1]).numpy(), geo["R_gen"])


@pytest.mark.parametrize("tag,Rk,H,W,transpose", [
    ("a2b", "R_a2b", 128, 256, False), ("a2b_8x", "R_a2b", 16, 32, False), ("a2b_w2c", "R_a2b", 128, 256, True),
    ("a2b_w2c_8x", "R_a2b", 16, 32, True), ("b2a", "R_b2a", 128, 256, False), ("b2a_8x", "R_b2a", 16, 32, False),
    ("b2a_w2c", "R_b2a", 128, 256, True), ("b2a_w2c_8x", "R_b2a", 16, 32, True)])
def test_samplegrid(geo, tag, Rk, H, W, transpose):
    R = geo[Rk].T if transpose else geo[Rk]
    # trig functions differ by an ulp between libms: <= 1e-4 px at W=256 (ulp(256) = 3e-5)
    assert np.abs(O.generate_samplegrid((1, 3, H, W), R) - geo[tag]).max() < 1e-4
    got = TO.generate_samplegrid((1, 3, H, W), T(np.ascontiguousarray(R))).numpy()
    assert np.array_equal(got, geo[tag])          # same ATen ops on the same device type: exact


def test_samplegrid_general_rotation(geo):
    assert np.abs(O.generate_samplegrid((2, 3, 24, 40), geo["R_gen"]) - geo["gen_24x40"]).max() < 2e-5
    assert np.array_equal(TO.generate_samplegrid((2, 3, 24, 40), T(geo["R_gen"])).numpy(), geo["gen_24x40"])


def test_volume_and_pyramid(gold_pyramids):
    f1a, f2a, f1b, f2b = cases.fmaps(seed=1)
    for (f1, f2), gold in zip(((f1a, f2a), (f1b, f2b)), gold_pyramids):
        pyr = O.build_pyramid(O.corr_volume(f1, f2))
        tpyr = TO.build_pyramid(TO.corr_volume(T(f1), T(f2)))
        for l in range(4):
            assert pyr[l].shape == gold[l].shape
            assert rel_to_max(pyr[l], gold[l]) < 1e-6
            assert np.array_equal(tpyr[l].numpy(), gold[l])
        # the pooling itself is exact given the finer level
        for l in range(3):
            assert np.array_equal(O.avg_pool2x2(gold[l]), gold[l + 1])


@pytest.mark.parametrize("tag,seed", [("rand", 2), ("edge", None)])
def test_dccl_lookup_bit_exact(geo, gold_pyramids, tag, seed):
    ga, gb = gold_pyramids
    c = cases.coords(seed=seed) if seed is not None else cases.edge_coords()
    g = golden("dccl_" + tag)
    own, other = O.dccl_lookup(c, ga, gb, geo["a2b_w2c_8x"], geo["b2a_8x"])
    assert np.array_equal(own, g["out_own"])
    assert np.array_equal(other, g["out_other"])
    town, tother = TO.dccl_lookup(T(c), [T(x) for x in ga], [T(x) for x in gb], T(geo["a2b_w2c_8x"]), T(geo["b2a_8x"]))
    assert np.array_equal(town.numpy(), g["out_own"]) and np.array_equal(tother.numpy(), g["out_other"])


def test_dccl_lookup_other_direction(geo, gold_pyramids):
    ga, gb = gold_pyramids
    g = golden("dccl_rand_b")
    own, other = O.dccl_lookup(cases.coords(seed=3), gb, ga, geo["b2a_w2c_8x"], geo["a2b_8x"])
    assert np.array_equal(own, g["out_own"]) and np.array_equal(other, g["out_other"])


def test_corrblock_lookup(gold_pyramids):
    ga, _ = gold_pyramids
    g = golden("corrblock")
    assert np.array_equal(O.corrblock_lookup(cases.coords(seed=2), ga), g["out"])
    assert np.array_equal(O.corrblock_lookup(cases.edge_coords(), ga), g["out_edge"])
    assert np.array_equal(TO.corrblock_lookup(T(cases.coords(seed=2)), [T(x) for x in ga]).numpy(), g["out"])


def test_samplers_and_coords_grid():
    img, pts = cases.small_sampler_case()
    g = golden("samplers")
    assert np.array_equal(O.cycle_bilinear_sampler(img, pts), g["cyc"])
    assert np.array_equal(O.bilinear_sampler(img, pts), g["lin"])
    assert np.array_equal(O.coords_grid(2, 5, 7), g["coords_grid"])
    assert np.array_equal(TO.cycle_bilinear_sampler(T(img), T(pts)).numpy(), g["cyc"])
    assert np.array_equal(TO.coords_grid(2, 5, 7, "cpu").numpy(), g["coords_grid"])


def test_cycle_sampler_is_not_a_true_wrap():
    """x in (W-1, W) blends column W-1 with zero, not with column 0."""
    img = np.arange(8, dtype=np.float32).reshape(1, 1, 1, 8).repeat(2, axis=2) + 1
    out = O.cycle_bilinear_sampler(img, np.array([[[[7.5, 0.0]]]], dtype=np.float32))
    assert out.item() == pytest.approx(0.5 * 8.0)
    # tiny negative x: remainder returns exactly W, i.e. column W-1 gets weight 0 and the sample is 0
    assert O.remainder(np.float32(-1e-8), 8) == np.float32(8.0)


def test_warp_groupcorr():
    f1a, f2a, _, _ = cases.fmaps(seed=1)
    g = golden("warp_groupcorr")["flaw"]
    assert rel_to_max(O.warp_groupcorr(f1a, f2a, cases.coords(seed=4)), g) < 1e-6   # mean's reduction order
    assert rel_to_max(TO.warp_groupcorr(T(f1a), T(f2a), T(cases.coords(seed=4))).numpy(), g) < 1e-6


def test_img_rotate(geo):
    g = golden("img_rotate")["out"]
    assert np.array_equal(O.img_rotate(cases.image(seed=5), geo["a2b_64x128"]), g)
    assert np.array_equal(TO.img_rotate(T(cases.image(seed=5)), T(geo["a2b_64x128"])).numpy(), g)


def test_flo_rotate_bit_exact(geo):
    g = golden("flo_rotate")
    gw, gc = np.repeat(geo["b2a_w2c_8x"], 2, 0), np.repeat(geo["b2a_8x"], 2, 0)
    fl = cases.flow(seed=6)
    assert np.array_equal(O.flow2endpoint(fl), g["end_w"])
    assert np.array_equal(O.cycle_grid_sample(gw, O.flow2endpoint(fl), True), g["end_c"])
    assert np.array_equal(O.flo_rotate(fl, gw, gc), g["out"])
    assert np.array_equal(O.flo_rotate(cases.flow(seed=7, sigma=60.0), gw, gc), g["out_big"])
    assert np.array_equal(TO.flo_rotate(T(fl), T(gw), T(gc)).numpy(), g["out"])
    assert np.array_equal(TO.flo_rotate(T(cases.flow(seed=7, sigma=60.0)), T(gw), T(gc)).numpy(), g["out_big"])


def test_alt_corr_equals_materialised_lookup(gold_pyramids):
    """No reference oracle exists for the on-the-fly variant (alt_cuda_corr is not shipped): it must equal the
    lookup into the pooled volume up to fp rounding, because avg-pool is linear."""
    f1a, f2a, _, _ = cases.fmaps(seed=1)
    ga, _ = gold_pyramids
    c = cases.coords(seed=2)
    want = O.corrblock_lookup(c, ga)  # non-cyclic; compare the cyclic flavour against a cyclic lookup instead
    pyr = O.build_pyramid(O.corr_volume(f1a, f2a))
    k = 9
    outs = []
    for lvl in range(4):
        Hl, Wl = pyr[lvl].shape[-2:]
        px, py = O.window_points(c, lvl, 4)
        ix, iy = O.pixel_to_sample_coords(px, py, Hl, Wl, True)
        outs.append(O.bilinear_zeros(pyr[lvl], ix, iy).reshape(1, 16, 32, k * k))
    want = np.concatenate(outs, -1).transpose(0, 3, 1, 2)
    got = O.alt_corr_lookup(f1a, f2a, c)
    assert rel_to_max(got, want) < 2e-6
