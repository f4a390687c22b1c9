"""prior_flow_b200 — B200 (sm_100a) implementation of PriOr-RAFT's correlation hot path.

Hand-written CUDA behind a plain C ABI (include/priorcorr.h, csrc/), a thin PyTorch host layer
(`ops`), and the host-side mirror of the reference interface (`corr`, `geometry`, `install`).
The repository directory `prior-flow_b200/` is an alias of this package (Python identifiers cannot
contain '-').  There is no CPU fallback: everything raises without the CUDA library / a GPU.
"""
__version__ = "0.1.0"

from . import ops  # noqa: F401
from .corr import DCCL, AlternateCorrBlock, CorrBlock, CostVolume  # noqa: F401
from .corr import corr as corr_volume  # noqa: F401  (the submodule keeps the name `corr`)
from .geometry import (bilinear_sampler, coords_grid, cycle_bilinear_sampler, flo_A2B, flo_B2A, flo_rotate,  # noqa: F401
                       generate_rotation_metrix, generate_samplegrid, img_A2B, img_B2A, img_rotate)
from .install import install, installed, uninstall  # noqa: F401
