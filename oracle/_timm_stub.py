"""Minimal stand-in for the three ``timm`` symbols the reference imports.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference depends on ``timm>=0.4.12`` (``/root/reference/requirements.txt:3``,
no lock file) for ``timm.models.vision_transformer.{PatchEmbed, Mlp}`` and
``timm.models.layers.DropPath`` (``models/vit.py:9-10``).  ``timm`` is not
installed here and cannot be (no network).  Only ``Mlp`` is ever executed
(``models/vit.py:118``); ``PatchEmbed`` is imported but unused and ``DropPath``
is only instantiated for ``drop_path > 0`` which never happens
(``models/vit.py:115,200``).  This module restates timm's published ``Mlp``
semantics: fc1 -> act_layer() -> Dropout(drop) -> fc2 -> Dropout(drop), with
parameter names ``fc1.*`` / ``fc2.*``.
"""
import sys
import types

import torch.nn as nn


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None,
                 act_layer=nn.GELU, drop=0.0):
        super().__init__()
        hidden_features = hidden_features or in_features
        out_features = out_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop)
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.drop1(self.act(self.fc1(x)))))


class PatchEmbed(nn.Module):  # imported by the reference, never used
    pass


class DropPath(nn.Module):  # never instantiated by the reference (drop_path == 0)
    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        return x


def install():
    """Register the stub under the ``timm`` module names (idempotent)."""
    if "timm" in sys.modules and not getattr(sys.modules["timm"], "_vrr_stub", False):
        return  # a real timm is importable: leave it alone
    timm = types.ModuleType("timm")
    timm._vrr_stub = True
    models = types.ModuleType("timm.models")
    vit = types.ModuleType("timm.models.vision_transformer")
    layers = types.ModuleType("timm.models.layers")
    vit.Mlp, vit.PatchEmbed, layers.DropPath = Mlp, PatchEmbed, DropPath
    timm.models, models.vision_transformer, models.layers = models, vit, layers
    sys.modules.update({
        "timm": timm,
        "timm.models": models,
        "timm.models.vision_transformer": vit,
        "timm.models.layers": layers,
    })
