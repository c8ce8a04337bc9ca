"""Leaf-only stand-in for `diffusers==0.34.0.dev0` (TEST INFRASTRUCTURE, not product code).

The reference (`/root/reference/models/{transformer,router,audio_model}.py`) imports a handful of
leaf modules from diffusers, which is not installable in this image.  This package
restates *only* those leaves, from the published diffusers algorithm, so that the reference's own
orchestration code can be imported **unmodified** and executed as the oracle.  parity unpinned: the
reference ships no golden vectors for these leaves.
"""
from .models.modeling_utils import ModelMixin  # noqa: F401
