"""vit_rpe_rope_b200 - the ViT attention hot path of zhengyk19/vit-rpe-rope on B200 (sm_100a).

``vit_rpe_rope_b200.models`` mirrors the reference's ``models`` package; the arithmetic of the hot
path runs in ``lib/libvrr_b200.so`` (C ABI: ``include/vrr.h``).  Importing the package does not
need a GPU; running a model does (CUDA sm_100 only, no fallback).
"""
from . import _lib, models, ops  # noqa: F401
from .models.vit import VisionTransformer  # noqa: F401

__all__ = ["models", "ops", "VisionTransformer"]
