"""numpy float64 restatement of the attention core, forward and backward.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Citations are relative
to ``/root/reference``.  The forward follows ``Attention.forward``
(models/vit.py:43-94) between the QKV projection and the output projection;
the backward is the hand-derived gradient of that forward -
the formulas the CUDA backward kernels implement - and is itself checked
against torch autograd of the reference in ``tests/test_oracle_golden.py``.
"""
import numpy as np

from . import tables_np


def rotate_half(x, cos, sin):
    """``apply_rotary_emb`` - models/rope_utils.py:3-37 (one operand).

    Pairs are ``(d, d + D/2)`` ("rotate-half", not interleaved):
    ``out = cat[x1*cos - x2*sin, x1*sin + x2*cos]``.
    ``x`` [..., N', D]; ``cos``/``sin`` broadcastable to [..., N', D/2].
    """
    d2 = x.shape[-1] // 2
    x1, x2 = x[..., :d2], x[..., d2:]
    return np.concatenate([x1 * cos - x2 * sin, x1 * sin + x2 * cos], axis=-1)


def rotate_half_inverse(g, cos, sin):
    """Gradient of :func:`rotate_half` w.r.t. ``x`` (the transpose rotation)."""
    d2 = g.shape[-1] // 2
    g1, g2 = g[..., :d2], g[..., d2:]
    return np.concatenate([g1 * cos + g2 * sin, -g1 * sin + g2 * cos], axis=-1)


def _bcast_cs(c):
    """``reshape_for_broadcast`` - models/rope_utils.py:39-66."""
    if c.ndim == 3:
        return c[None]
    if c.ndim == 2:
        return c[None, None]
    raise ValueError(f"Unexpected tensor shapes: {c.shape}")


def apply_rope_skip_cls(q, k, cos, sin):
    """RoPE branch of ``Attention.forward`` - models/vit.py:51-68.

    Token 0 (cls) is passed through; tokens 1.. are rotated.
    """
    c, s = _bcast_cs(np.asarray(cos, np.float64)), _bcast_cs(np.asarray(sin, np.float64))
    qr = np.concatenate([q[:, :, :1], rotate_half(q[:, :, 1:], c, s)], axis=2)
    kr = np.concatenate([k[:, :, :1], rotate_half(k[:, :, 1:], c, s)], axis=2)
    return qr, kr


def attention_forward(q, k, v, scale, bias=None, cos=None, sin=None):
    """softmax(q k^T * scale + bias) v  - models/vit.py:51-88.

    q, k, v: [B, H, N, D] (any float dtype, promoted to float64).
    bias: [H, N, N] or None (relative / polynomial branch, vit.py:73-81).
    cos, sin: RoPE tables or None (vit.py:51-71; no bias in that branch).
    Returns ``(O [B, N, H*D], P [B, H, N, N], q_rot, k_rot)``.
    """
    q, k, v = (np.asarray(t, np.float64) for t in (q, k, v))
    if cos is not None:
        q, k = apply_rope_skip_cls(q, k, cos, sin)
    s = np.einsum("bhid,bhjd->bhij", q, k) * scale  # scale applied after QK^T, before the bias
    if bias is not None:
        s = s + np.asarray(bias, np.float64)[None]
    s = s - s.max(axis=-1, keepdims=True)
    p = np.exp(s)
    p /= p.sum(axis=-1, keepdims=True)
    o = np.einsum("bhij,bhjd->bhid", p, v)
    b, h, n, d = o.shape
    return o.transpose(0, 2, 1, 3).reshape(b, n, h * d), p, q, k


def attention_backward(d_out, q, k, v, scale, bias=None, cos=None, sin=None, out=None):
    """Gradient of :func:`attention_forward`.

    ``out``: optionally the forward output [B, N, H*D] *as stored* (e.g. rounded to bf16).  The
    row statistic ``delta = rowsum(dO * O)`` is then taken from it, which is what a backward pass
    that re-reads the stored output computes (the reference's autograd does the equivalent with
    its own stored softmax output).

    Returns a dict with ``dq, dk, dv`` [B,H,N,D] (w.r.t. the UN-rotated q, k),
    ``dbias`` [H,N,N] (sum over batch of dS), and for RoPE ``dcos, dsin`` with
    the shape of ``cos`` (sum over batch, and over heads when cos is 2-D).
    """
    q0, k0, v = (np.asarray(t, np.float64) for t in (q, k, v))
    o, p, qr, kr = attention_forward(q0, k0, v, scale, bias, cos, sin)
    b, h, n, d = q0.shape
    do = np.asarray(d_out, np.float64).reshape(b, n, h, d).transpose(0, 2, 1, 3)
    if out is not None:
        o = np.asarray(out, np.float64)
    oh = o.reshape(b, n, h, d).transpose(0, 2, 1, 3)
    dv = np.einsum("bhij,bhid->bhjd", p, do)
    dp = np.einsum("bhid,bhjd->bhij", do, v)
    delta = (do * oh).sum(-1, keepdims=True)
    ds = p * (dp - delta)
    dqr = scale * np.einsum("bhij,bhjd->bhid", ds, kr)
    dkr = scale * np.einsum("bhij,bhid->bhjd", ds, qr)
    out = {"dv": dv, "dbias": ds.sum(0)}
    if cos is None:
        out["dq"], out["dk"] = dqr, dkr
        return out
    c, s = _bcast_cs(np.asarray(cos, np.float64)), _bcast_cs(np.asarray(sin, np.float64))
    d2 = d // 2
    dq = dqr.copy()
    dk = dkr.copy()
    dq[:, :, 1:] = rotate_half_inverse(dqr[:, :, 1:], c, s)
    dk[:, :, 1:] = rotate_half_inverse(dkr[:, :, 1:], c, s)
    dc = np.zeros((b, h, n - 1, d2))
    dsn = np.zeros((b, h, n - 1, d2))
    for x, g in ((q0, dqr), (k0, dkr)):
        x1, x2 = x[:, :, 1:, :d2], x[:, :, 1:, d2:]
        g1, g2 = g[:, :, 1:, :d2], g[:, :, 1:, d2:]
        dc += g1 * x1 + g2 * x2
        dsn += -g1 * x2 + g2 * x1
    if np.asarray(cos).ndim == 3:
        out["dcos"], out["dsin"] = dc.sum(0), dsn.sum(0)
    else:
        out["dcos"], out["dsin"] = dc.sum((0, 1)), dsn.sum((0, 1))
    out["dq"], out["dk"] = dq, dk
    return out


def dtable_from_dbias(dbias):
    """Backward of ``table[:, idx]`` (index_put accumulate): dTable[h, i-j+L-1] += dbias[h,i,j]."""
    h, n, _ = dbias.shape
    idx = tables_np.relative_position_index(n)
    out = np.zeros((h, 2 * n - 1))
    for hh in range(h):
        np.add.at(out[hh], idx, dbias[hh])
    return out


def dcoef_from_dbias(dbias, degree, shared=True):
    """Backward of the polynomial bias: dcoef[k] = sum dS[h, i>=1, j>=1] * dist**k."""
    h, n, _ = dbias.shape
    dist = tables_np.poly_l1_distance(n - 1).astype(np.float64)
    core = dbias[:, 1:, 1:]
    per_head = np.stack([(core * dist[None] ** k).sum((1, 2)) for k in range(degree + 1)], axis=-1)
    return per_head.sum(0) if shared else per_head
