"""Runs the UNMODIFIED reference (`/root/reference/models/*.py`) on the leaf-only diffusers shim.

TEST INFRASTRUCTURE.  Only usable where the reference tree is mounted (this build container); the GPU box has no
`/root/reference`, so `-m gpu` tests use `oracle/restated.py` (asserted equal to this on CPU) and the committed
goldens instead.  parity unpinned (no reference-side golden vectors exist).
"""
from __future__ import annotations

import os
import sys

import torch

REFERENCE_ROOT = os.environ.get("BYA_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "diffusers_shim")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "transformer.py"))


def import_reference():
    """Returns the reference's `models.transformer` module (imports router / audio_model on the way)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (REFERENCE_ROOT, _SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    import models.router  # noqa: F401
    import models.audio_model  # noqa: F401
    import models.transformer as T

    return T


def build_reference_model(cfg, seed: int = 0, dtype=torch.float32):
    """Reference `BindyouravatarTransformer3DModel` with seeded weights (`bya_b200.synth.fill_module`).

    Patches the router's hard-coded 13x45x30 grid for other geometries exactly as SURVEY.md §8c prescribes:
    `router.height := grid_w`, `router.width := grid_h` (the reference's swapped naming, `router.py:312-314`)."""
    import bya_b200  # noqa: F401
    from bya_b200.synth import fill_module

    T = import_reference()
    assert cfg.chars == 2, "the verbatim reference hard-codes two characters (transformer.py:784)"
    m = T.BindyouravatarTransformer3DModel(**cfg.ctor_kwargs()).eval()
    r = m.router
    if (r.frames, r.height, r.width) != (cfg.frames, cfg.grid_w, cfg.grid_h):
        r.frames, r.height, r.width = cfg.frames, cfg.grid_w, cfg.grid_h
        r.pos_emb = r._create_positional_embedding()
    fill_module(m, seed)
    return m.to(dtype)


@torch.no_grad()
def run_reference(model, inputs: dict):
    kw = dict(inputs)
    kw.setdefault("return_dict", False)
    return model(**kw)[0]
