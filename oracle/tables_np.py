"""numpy restatement of the reference's index / coordinate / frequency tables.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Citations are relative
to ``/root/reference``.  Integer tables and the integer-valued fp32 coordinate
tables are bit-exact targets; the fp32 angle tables are products of exactly
representable operands and are bit-exact as well; cos/sin go through libm and
are compared to 1 ulp-level tolerance in the tests.
"""
import math

import numpy as np


def relative_position_index(seq_length: int) -> np.ndarray:
    """``RelativePositionalEncoding.__init__`` - models/positional_encoding.py:67-75.

    1-D Toeplitz index over raster token order (cls token is position 0):
    ``idx[i, j] = i - j + (L - 1)``, clamped to ``[0, 2L-2]`` (the clamp never
    fires).  int64, shape [L, L].
    """
    pos = np.arange(seq_length, dtype=np.int64)
    idx = pos[:, None] - pos[None, :] + (seq_length - 1)
    return np.clip(idx, 0, 2 * seq_length - 2)


def relative_bias(table: np.ndarray, seq_length: int) -> np.ndarray:
    """``RelativePositionalEncoding.get_bias`` - models/positional_encoding.py:82-95.

    ``table`` is [H, 2L-1]; returns ``table[:, idx]`` of shape [H, L, L].
    """
    return table[:, relative_position_index(seq_length)]


def grid_coords(h: int, w: int):
    """``init_t_xy`` - models/positional_encoding.py:198-214 and :292-311.

    ``t_x = t % w`` (column), ``t_y = floor(t / w)`` (row), both fp32.
    """
    t = np.arange(h * w, dtype=np.float32)
    t_x = np.mod(t, np.float32(w)).astype(np.float32)
    t_y = np.floor(t / np.float32(w)).astype(np.float32)
    return t_x, t_y


def axial_inv_freq(head_dim: int, theta: float) -> np.ndarray:
    """``RoPEAxial.__init__`` - models/positional_encoding.py:184-192.

    ``inv_freq[i] = 1 / theta ** (i / (head_dim // 4))`` for i in [0, head_dim//4), fp32.
    """
    q = head_dim // 4
    expo = np.arange(0, q, dtype=np.float32) / np.float32(q)
    return (np.float32(1.0) / np.power(np.float32(theta), expo, dtype=np.float32)).astype(np.float32)


def axial_angles(num_patches: int, inv_freq: np.ndarray) -> np.ndarray:
    """``RoPEAxial.get_freqs_cis`` - models/positional_encoding.py:216-245 (before cos/sin).

    Square grid ``g = int(sqrt(num_patches))``; angles [g*g, head_dim/2] =
    ``cat[outer(t_x, inv_freq), outer(t_y, inv_freq)]`` in fp32.
    """
    g = int(math.sqrt(num_patches))
    t_x, t_y = grid_coords(g, g)
    fx = (t_x[:, None] * inv_freq[None, :]).astype(np.float32)
    fy = (t_y[:, None] * inv_freq[None, :]).astype(np.float32)
    return np.concatenate([fx, fy], axis=-1)


def axial_cos_sin(num_patches: int, inv_freq: np.ndarray):
    a = axial_angles(num_patches, inv_freq)
    return np.cos(a).astype(np.float32), np.sin(a).astype(np.float32)


def mixed_scramble(num_heads: int, num_patches: int):
    """The head/position scramble of ``RoPEMixed.get_freqs_cis``.

    models/positional_encoding.py:337-342: ``t_x[:, None] @ freqs[0][:, None, :]``
    broadcasts to a contiguous [H, N, D/2] tensor, which is then *reinterpreted*
    by ``.view(N, H, -1).permute(1, 0, 2)``.  Output element ``[h', n', :]``
    therefore reads source head ``hs`` and source position ``ps`` with
    ``(hs, ps) = divmod(n' * H + h', N)``.  Returns int64 arrays (hs, ps), each
    of shape [H, N].
    """
    hp = np.arange(num_heads, dtype=np.int64)[:, None]
    n = np.arange(num_patches, dtype=np.int64)[None, :]
    flat = n * num_heads + hp
    return flat // num_patches, flat % num_patches


def mixed_angles(freqs: np.ndarray, num_patches: int) -> np.ndarray:
    """``RoPEMixed.get_freqs_cis`` - models/positional_encoding.py:313-346 (before cos/sin).

    ``freqs`` is the learnable [2, H, D/2] parameter.  Returns fp32 [H, N, D/2]:
    ``phase[h', n', d] = t_x[ps] * freqs[0, hs, d] + t_y[ps] * freqs[1, hs, d]``.
    """
    freqs = np.asarray(freqs, dtype=np.float32)
    num_heads = freqs.shape[1]
    g = int(math.sqrt(num_patches))
    t_x, t_y = grid_coords(g, g)
    hs, ps = mixed_scramble(num_heads, g * g)
    px = (t_x[ps][..., None] * freqs[0][hs]).astype(np.float32)
    py = (t_y[ps][..., None] * freqs[1][hs]).astype(np.float32)
    return (px + py).astype(np.float32)


def mixed_cos_sin(freqs: np.ndarray, num_patches: int):
    a = mixed_angles(freqs, num_patches)
    return np.cos(a).astype(np.float32), np.sin(a).astype(np.float32)


def mixed_init_freqs(head_dim: int, num_heads: int, theta: float, angles: np.ndarray) -> np.ndarray:
    """``RoPEMixed.__init__`` - models/positional_encoding.py:258-290, given the per-head
    random angles (the reference draws ``torch.rand(1) * 2 * pi`` per head).

    Returns [2, H, head_dim/2] float64 (compare with tolerance; torch computes in fp32).
    """
    mag = 1.0 / theta ** (np.arange(0, head_dim, 4)[: head_dim // 4].astype(np.float64) / head_dim)
    fx, fy = [], []
    for a in np.asarray(angles, dtype=np.float64).reshape(-1)[:num_heads]:
        fx.append(np.concatenate([mag * np.cos(a), mag * np.cos(np.pi / 2 + a)]))
        fy.append(np.concatenate([mag * np.sin(a), mag * np.sin(np.pi / 2 + a)]))
    return np.stack([np.stack(fx), np.stack(fy)])


def poly_l1_distance(num_patches: int) -> np.ndarray:
    """``PolynomialRPE.get_bias`` coordinates - models/positional_encoding.py:134-142.

    Patch p has ``y = p % g`` (``arange(g).repeat(g)``) and ``x = p // g``
    (``repeat_interleave``); the feature is the scalar L1 distance
    ``|dy| + |dx|``.  int64 [Np, Np].
    """
    g = int(math.sqrt(num_patches))
    p = np.arange(g * g, dtype=np.int64)
    y, x = p % g, p // g
    return np.abs(y[:, None] - y[None, :]) + np.abs(x[:, None] - x[None, :])


def poly_bias(coefficients: np.ndarray, num_patches: int, num_heads: int) -> np.ndarray:
    """``PolynomialRPE.get_bias`` - models/positional_encoding.py:127-171.

    ``coefficients`` is [deg+1] (shared) or [H, deg+1].  ``bias = sum_k c_k * d**k``
    (``0**0 == 1``), zero-padded with a cls row and column.  fp32 [H, Np+1, Np+1].
    """
    coef = np.asarray(coefficients, dtype=np.float32)
    d = poly_l1_distance(num_patches).astype(np.float32)
    g2 = d.shape[0]
    deg = coef.shape[-1] - 1
    feats = np.stack([np.power(d, np.float32(k), dtype=np.float32) for k in range(deg + 1)], axis=-1)
    if coef.ndim == 1:
        core = np.broadcast_to((feats @ coef).astype(np.float32)[None], (num_heads, g2, g2))
    else:
        core = np.stack([(feats @ coef[h]).astype(np.float32) for h in range(num_heads)])
    out = np.zeros((num_heads, g2 + 1, g2 + 1), dtype=np.float32)
    out[:, 1:, 1:] = core
    return out
