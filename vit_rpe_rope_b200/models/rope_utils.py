"""RoPE helpers: drop-in for ``models/rope_utils.py`` of the reference.

``apply_rotary_emb`` / ``reshape_for_broadcast`` keep the reference's names, argument meaning and
error behaviour (citations: /root/reference/models/rope_utils.py).  ``apply_rotary_emb`` runs the
library's rotate-half kernel (CUDA sm_100 only).  Inside ``Attention.forward`` neither is called:
the rotation is applied in the epilogue of the QKV projection GEMM.
"""
from .. import ops


def reshape_for_broadcast(x, target_tensor):
    """[seq, dim/2] -> [1, 1, seq, dim/2]; [heads, seq, dim/2] -> [1, heads, seq, dim/2]
    (reference :39-66).  Anything else raises ``ValueError`` like the reference."""
    if target_tensor.ndim == 4 and x.ndim == 3:
        return x.unsqueeze(0)
    if target_tensor.ndim == 4 and x.ndim == 2:
        return x.unsqueeze(0).unsqueeze(0)
    raise ValueError(f"Unexpected tensor shapes: {x.shape} vs {target_tensor.shape}")


def apply_rotary_emb(q, k, cos, sin):
    """Rotate-half RoPE on ``q`` and ``k`` [B, H, N, D] (reference :3-37): channel pairs are
    ``(d, d + D/2)``, out = cat[x1*cos - x2*sin, x1*sin + x2*cos].  ``cos`` / ``sin`` may be
    [N, D/2], [H, N, D/2] or their ``reshape_for_broadcast`` forms.  Differentiable w.r.t. q, k."""
    return ops.RopeApplyFn.apply(q, k, cos, sin)
