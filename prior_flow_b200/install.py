"""Drop-in installation into an *unmodified* checkout of the reference.

`core/prior_raft.py` stays byte-identical: the replacement hooks in through the names that file resolves
(SURVEY.md §8b):

    core.prior_raft.DCCL                          (bound at import, prior_raft.py:7; used :157-159,185-186)
    core.prior_raft.cycle_bilinear_sampler        (prior_raft.py:8; used :173,181)
    core.prior_raft.PriOr_RAFT.corr               (method, prior_raft.py:69; called :151-152)
    core.utils.projection_prim_ortho.{generate_rotation_metrix, generate_samplegrid, img_rotate, flo_rotate,
                                      img_A2B, img_B2A, flo_A2B, flo_B2A}   (module attributes looked up at call time)
    core.corr.{DCCL, CorrBlock, AlternateCorrBlock}, core.utils.utils.{cycle_bilinear_sampler, bilinear_sampler}

Usage (with PriOr-RAFT/ on sys.path):

    import prior_flow_b200
    prior_flow_b200.install()        # before or after `from core.prior_raft import PriOr_RAFT`
    model = PriOr_RAFT(args).cuda()  # runs the sm_100a hot path; fnet/cnet/update blocks stay on cuDNN
"""
from __future__ import annotations

import importlib
from typing import Dict, List, Tuple

_corr = importlib.import_module(".corr", __package__)
_geo = importlib.import_module(".geometry", __package__)

_saved: List[Tuple[object, str, object]] = []


def _patch(obj, name: str, value) -> None:
    _saved.append((obj, name, getattr(obj, name, None)))
    setattr(obj, name, value)


def _corr_method(self, fmap1, fmap2):
    """Replacement for PriOr_RAFT.corr (core/prior_raft.py:69-75): a lazy handle, consumed by DCCL.build_pyramid."""
    return _corr.corr(fmap1, fmap2)


def install(package: str = "core") -> Dict[str, int]:
    """Patches the reference modules (imported from `package`).  Idempotent.  Returns {module: #names patched}."""
    if _saved:
        return {}
    prior_raft = importlib.import_module(f"{package}.prior_raft")
    ref_corr = importlib.import_module(f"{package}.corr")
    utils = importlib.import_module(f"{package}.utils.utils")
    ppo = importlib.import_module(f"{package}.utils.projection_prim_ortho")
    counts = {}

    _patch(prior_raft, "DCCL", _corr.DCCL)
    _patch(prior_raft, "cycle_bilinear_sampler", _geo.cycle_bilinear_sampler)
    _patch(prior_raft, "bilinear_sampler", _geo.bilinear_sampler)
    _patch(prior_raft.PriOr_RAFT, "corr", _corr_method)
    counts[prior_raft.__name__] = 4

    for name in ("DCCL", "CorrBlock", "AlternateCorrBlock"):
        _patch(ref_corr, name, getattr(_corr, name))
    _patch(ref_corr, "cycle_bilinear_sampler", _geo.cycle_bilinear_sampler)
    _patch(ref_corr, "bilinear_sampler", _geo.bilinear_sampler)
    counts[ref_corr.__name__] = 5

    for name in ("cycle_bilinear_sampler", "bilinear_sampler"):
        _patch(utils, name, getattr(_geo, name))
    counts[utils.__name__] = 2

    geo_names = ("generate_rotation_metrix", "generate_samplegrid", "img_rotate", "flo_rotate",
                 "img_A2B", "img_B2A", "flo_A2B", "flo_B2A")
    for name in geo_names:
        _patch(ppo, name, getattr(_geo, name))
    counts[ppo.__name__] = len(geo_names)
    return counts


def uninstall() -> None:
    """Restores every patched name."""
    while _saved:
        obj, name, old = _saved.pop()
        if old is None:
            try:
                delattr(obj, name)
            except AttributeError:
                pass
        else:
            setattr(obj, name, old)


def installed() -> bool:
    return bool(_saved)
