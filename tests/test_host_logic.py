"""CPU: host-side logic that needs neither a GPU nor the CUDA library — drop-in installation into the unmodified
reference (container only), state-dict compatibility of the surrounding network, sharding, and a 2-rank gloo run
of the multi-process plumbing bench.py uses."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import cases
from oracle import ref_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_reference = pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not mounted (GPU box)")


@needs_reference
def test_install_patches_every_name_prior_raft_resolves():
    import prior_flow_b200 as pfb
    ref = ref_shim.load()
    orig_dccl, orig_corr = ref.prior_raft.DCCL, ref.prior_raft.PriOr_RAFT.corr
    try:
        counts = pfb.install()
        assert pfb.installed() and sum(counts.values()) >= 19
        assert ref.prior_raft.DCCL is pfb.DCCL and ref.corr.DCCL is pfb.DCCL and ref.corr.CorrBlock is pfb.CorrBlock
        assert ref.prior_raft.cycle_bilinear_sampler is pfb.cycle_bilinear_sampler
        assert ref.ppo.generate_samplegrid is pfb.generate_samplegrid and ref.ppo.flo_rotate is pfb.flo_rotate
        assert ref.ppo.img_rotate is pfb.img_rotate and ref.utils.bilinear_sampler is pfb.bilinear_sampler
        vol = ref.prior_raft.PriOr_RAFT.corr(None, torch.zeros(1, 256, 8, 32), torch.zeros(1, 256, 8, 32))
        assert isinstance(vol, pfb.CostVolume) and tuple(vol.shape) == (1, 8, 32, 8, 32)
        assert pfb.install() == {}                      # idempotent
    finally:
        pfb.uninstall()
    assert ref.prior_raft.DCCL is orig_dccl and ref.prior_raft.PriOr_RAFT.corr is orig_corr and not pfb.installed()


@needs_reference
def test_network_is_state_dict_compatible_with_the_reference():
    from argparse import Namespace
    from prior_flow_b200.model import PriOrRAFT
    ref = ref_shim.load()
    theirs = ref.prior_raft.PriOr_RAFT(Namespace(mixed_precision=False, dropout=0.0)).state_dict()
    mine = PriOrRAFT()
    assert {k: tuple(v.shape) for k, v in theirs.items()} == {k: tuple(v.shape) for k, v in mine.state_dict().items()}
    mine.load_state_dict({"module." + k: v for k, v in theirs.items()} if False else theirs, strict=True)
    assert sum(p.numel() for p in mine.parameters()) == 8337646


def test_eager_network_reproduces_reference_golden_on_cpu():
    """The network around the hot path (oracle/cpu_model.py = model.py's modules + eager ATen hot path) against the
    unmodified reference's end-to-end flows: identical op sequence on the same device type -> bit-exact."""
    from oracle.cpu_model import EagerPriOrRAFT
    g = np.load(os.path.join(ROOT, "tests", "golden", "e2e.npz"))
    m = EagerPriOrRAFT().eval()
    m.load_state_dict(cases.seeded_state_dict(m.state_dict()), strict=True)
    im1, im2 = (torch.from_numpy(x) for x in cases.e2e_images())
    with torch.no_grad():
        flow = m(im1, im2, iters=4, test_mode=True).numpy()
    assert np.array_equal(flow, g["flow4"])


def test_rotation_matrix_host_matches_golden(geo):
    from prior_flow_b200 import geometry
    assert np.array_equal(geometry.rotation_matrix_host([0., 0., -np.pi / 2]).numpy(), geo["R_a2b"])
    assert np.array_equal(geometry.rotation_matrix_host([0.3, -0.7, 1.1]).numpy(), geo["R_gen"])
    assert np.array_equal(geometry.coords_grid(2, 5, 7, "cpu").numpy(), np.load(os.path.join(ROOT, "tests/golden/samplers.npz"))["coords_grid"])


def test_shard_pairs_partitions_the_batch():
    from prior_flow_b200.distributed import shard_pairs
    for total, world in ((64, 8), (10, 4), (3, 8), (0, 2)):
        parts = [shard_pairs(total, r, world) for r in range(world)]
        assert sum(len(p) for p in parts) == total
        assert sorted(i for p in parts for i in p) == list(range(total))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_two_rank_gloo_plumbing():
    """world_size 2 over gloo: rank/shard bookkeeping, max-over-ranks timing reduce, DDP loss scaling."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "gloo_worker.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "rank0 ok" in out.stdout and "rank1 ok" in out.stdout


def test_dccl_auto_mode_decides_by_free_memory_once(monkeypatch):
    """mode="auto": materialise while both pyramids fit in a fraction of the free device memory, else the volume-free
    lookup; decided once per instance and shape so that both views take the same path."""
    from types import SimpleNamespace
    from prior_flow_b200 import corr as pcorr

    fmap = SimpleNamespace(shape=(1, 256, 128, 256), device="cuda:0", is_cuda=True)     # 1024x2048 ERP: 2 x 5.7 GB
    free = {"bytes": 170 << 30}
    monkeypatch.setattr(torch.cuda, "mem_get_info", lambda dev=None: (free["bytes"], 180 << 30))
    monkeypatch.setattr(pcorr, "DRIVER_FREE_TTL_S", -1.0)          # every instance asks the driver (no reuse of the last answer)
    big = pcorr.DCCL(4, 4, mode="auto")
    assert big._use_onthefly(fmap) is False              # fits on a 180 GB B200
    free["bytes"] = 1 << 30
    assert big._use_onthefly(fmap) is False              # the first decision stands for the second view
    small = pcorr.DCCL(4, 4, mode="auto")
    assert small._use_onthefly(fmap) is True             # 1 GiB free: volume-free lookup
    assert pcorr.DCCL(4, 4, mode="onthefly")._use_onthefly(fmap) is True
    assert pcorr.DCCL(4, 4, mode="materialized")._use_onthefly(fmap) is False
    # the driver's answer is reused for DRIVER_FREE_TTL_S: a new DCCL per forward (the reference's pattern) does not query per forward
    calls = {"n": 0}

    def counting(dev=None):
        calls["n"] += 1
        return (170 << 30, 180 << 30)
    monkeypatch.setattr(torch.cuda, "mem_get_info", counting)
    monkeypatch.setattr(pcorr, "DRIVER_FREE_TTL_S", 60.0)
    pcorr._driver_free.clear()
    for _ in range(5):
        assert pcorr.DCCL(4, 4, mode="auto")._use_onthefly(fmap) is False
    assert calls["n"] == 1


def test_onthefly_tensor_core_path_shape_rules_and_scratch_sizes():
    """Host logic of the tensor-core on-the-fly lookup: which feature grids take it, and the scratch sizes the C ABI documents
    (PF_OTF_WORK_INTS in include/priorcorr.h) as the Python layer computes them."""
    import re
    from prior_flow_b200 import ops
    ok = lambda *shape, radius=4: ops.OnTheFlyPlanes.supported(torch.empty(*shape), radius)
    assert ok(1, 64, 128, 256) and ok(2, 8, 16, 128) and ok(1, 128, 256, 512)
    assert not ok(1, 60, 128, 256)          # h % 8
    assert not ok(1, 64, 120, 256)          # w % 16
    assert not ok(1, 64, 128, 192)          # channels % 128
    assert not ok(1, 64, 128, 640)          # channels > 512
    assert not ok(1, 64, 128, 256, radius=3)
    header = open(os.path.join(ROOT, "include", "priorcorr.h")).read()
    m = re.search(r"#define PF_OTF_WORK_INTS\(T, segs\) \((.*)\)\n", header)
    assert m, "PF_OTF_WORK_INTS not found"
    expr = m.group(1).replace("(long long)", "").replace("/", "//")
    for T, segs in ((512, 32768), (4096, 131072), (8, 8)):
        assert eval(expr, {"T": T, "segs": segs}) == 16 + 10 * T + 2 + 2 * (segs // 8 + T)      # ops._otf_scratch
