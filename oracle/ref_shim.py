"""TEST INFRASTRUCTURE ONLY — imports the *unmodified* reference (container only).

The reference (`/root/reference/PriOr-RAFT`) is a read-only mount that exists in
the build container but NOT on the GPU box.  This shim is what
`tests/golden/make_golden.py` uses to run the reference's own code on CPU and
freeze golden vectors; nothing in the product path, `bench.py`, `smoke()` or
the `-m gpu` tests may import it.

What the shim does:
  * stubs `timm` and `omegaconf` in `sys.modules` — both are imported by the
    reference (`core/extractor.py:4`, `core/__init__.py:3`) but never used by
    the model;
  * makes `Tensor.cuda` / `Module.cuda` no-ops when there is no GPU, because
    the reference hard-codes `.cuda()` in its geometry helpers
    (`core/utils/projection_prim_ortho.py:19,29,37,42,47,48,66,411`).
No reference file is modified or copied.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PRIORFLOW_REFERENCE", "/root/reference/PriOr-RAFT")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "core"))


def load():
    """Returns a namespace with the reference modules (raises if the mount is absent)."""
    if not available():
        raise RuntimeError(f"reference not mounted at {REFERENCE_ROOT}")
    import torch

    for name in ("omegaconf", "timm"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["omegaconf"].OmegaConf = object
    sys.modules["omegaconf"].ListConfig = object
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import core.corr as corr
    import core.prior_raft as prior_raft
    import core.update as update
    import core.extractor as extractor
    import core.utils.utils as utils
    import core.utils.projection_prim_ortho as ppo
    import core.utils.my_cycle_sample as mcs

    return types.SimpleNamespace(corr=corr, prior_raft=prior_raft, update=update,
                                 extractor=extractor, utils=utils, ppo=ppo, mcs=mcs)
