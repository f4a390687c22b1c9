"""TEST / BASELINE INFRASTRUCTURE ONLY — imports the *unmodified* reference.

Where the reference comes from, in this order:
  1. `$PRIORFLOW_REFERENCE`, if set;
  2. `/root/reference/PriOr-RAFT` — the read-only mount of the build container (absent on the GPU box);
  3. `baseline/_ref/PriOr-RAFT` — the byte-for-byte copy `scripts/vendor_reference.py` makes (git-ignored, travels
     to the GPU box with the snapshot; `baseline/ref_manifest.json` holds the sha256 of every file and `verified()`
     checks them, so "unmodified" is a tested statement, not a promise).

Users: `tests/golden/make_golden.py` (container, freezes golden vectors), the `-m gpu` drop-in tests (unmodified
`PriOr_RAFT` with and without `prior_flow_b200.install()`), and `bench.py`'s reference legs (`--impl reference`,
`gpu_eager_baseline`, `dropin`).  Nothing under `prior_flow_b200/` may import it.

What the shim does:
  * stubs `timm` and `omegaconf` in `sys.modules` — both are imported by the reference (`core/extractor.py:4`,
    `core/__init__.py:3`) but never used by the model;
  * `load(cpu=True)` additionally makes `Tensor.cuda` / `Module.cuda` no-ops, because the reference hard-codes `.cuda()`
    in its geometry helpers (`core/utils/projection_prim_ortho.py:19,29,37,42,47,48,66,411`) — that is how the
    reference runs on host cores (golden vectors, CPU baseline).  `unpatch_cuda()` undoes it.
No reference file is modified.
"""
import hashlib
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_MOUNT = "/root/reference/PriOr-RAFT"
_VENDORED = os.path.join(ROOT, "baseline", "_ref", "PriOr-RAFT")
_MANIFEST = os.path.join(ROOT, "baseline", "ref_manifest.json")


def reference_root() -> str:
    env = os.environ.get("PRIORFLOW_REFERENCE")
    if env:
        return env
    if os.path.isdir(os.path.join(_MOUNT, "core")):
        return _MOUNT
    return _VENDORED


REFERENCE_ROOT = reference_root()


def available() -> bool:
    return os.path.isdir(os.path.join(reference_root(), "core"))


def verified() -> bool:
    """Every file named in baseline/ref_manifest.json exists under the reference root with the recorded sha256."""
    if not available() or not os.path.isfile(_MANIFEST):
        return False
    with open(_MANIFEST) as fh:
        man = json.load(fh)["files"]
    base = reference_root()
    for rel, h in man.items():
        p = os.path.join(base, rel)
        if not os.path.isfile(p):
            return False
        with open(p, "rb") as fh:
            if hashlib.sha256(fh.read()).hexdigest() != h:
                return False
    return True


_orig_cuda = {}


def patch_cuda_noop() -> None:
    import torch
    if not _orig_cuda:
        _orig_cuda["t"], _orig_cuda["m"] = torch.Tensor.cuda, torch.nn.Module.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self


def unpatch_cuda() -> None:
    import torch
    if _orig_cuda:
        torch.Tensor.cuda, torch.nn.Module.cuda = _orig_cuda.pop("t"), _orig_cuda.pop("m")


def load(cpu=None):
    """Returns a namespace with the reference modules (raises if the reference is absent).
    cpu=None: patch `.cuda()` to a no-op only when there is no GPU; cpu=True: always (CPU baseline on a GPU box)."""
    base = reference_root()
    if not available():
        raise RuntimeError(f"reference not found at {base} (run scripts/vendor_reference.py in the build container)")
    import torch

    for name in ("omegaconf", "timm"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["omegaconf"].OmegaConf = object
    sys.modules["omegaconf"].ListConfig = object
    if cpu is True or (cpu is None and not torch.cuda.is_available()):
        patch_cuda_noop()
    if base not in sys.path:
        sys.path.insert(0, base)
    import core.corr as corr
    import core.prior_raft as prior_raft
    import core.update as update
    import core.extractor as extractor
    import core.utils.utils as utils
    import core.utils.projection_prim_ortho as ppo
    import core.utils.my_cycle_sample as mcs

    return types.SimpleNamespace(corr=corr, prior_raft=prior_raft, update=update, extractor=extractor, utils=utils,
                                 ppo=ppo, mcs=mcs, root=base)


def make_model(ref=None, seed: int = 0, mixed_precision: bool = False):
    """`PriOr_RAFT(Namespace(mixed_precision, dropout=0))` with `torch.manual_seed(seed)` — SURVEY.md §8(d)."""
    import argparse
    import torch
    ref = ref or load()
    torch.manual_seed(seed)
    return ref.prior_raft.PriOr_RAFT(argparse.Namespace(mixed_precision=mixed_precision, dropout=0.0))
