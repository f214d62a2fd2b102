"""Positional-encoding modules: drop-in for ``models/positional_encoding.py`` of the reference.

Same class names, constructor signatures, parameter / buffer names (hence ``state_dict`` keys),
attributes and methods as the reference (citations: /root/reference/models/positional_encoding.py),
re-implemented.  These modules are *parameter holders and table builders*: the learnable tensors
live here, and ``get_bias()`` / ``get_freqs_cis()`` build the small tables the reference's
visualisers read (``pe_similarity_visualizer.py:234,273,338``; ``rope_visualizer.py:163-172``) with
device-agnostic torch ops.  On the accelerated path (``models/vit.py`` of this package) the logit
bias is never materialised - the attention kernel gathers ``relative_position_bias_table`` /
evaluates ``coefficients`` on the fly - and the (cos, sin) tables built here once per forward are
consumed by the fused QKV+RoPE GEMM epilogue.
"""
import math

import torch
import torch.nn as nn


def _grid_positions(h, w, device):
    """Column / row coordinate of every raster position, fp32 (reference ``init_t_xy``, :198-214)."""
    t = torch.arange(h * w, device=device, dtype=torch.float32)
    return (t % w).float(), torch.div(t, w, rounding_mode="floor").float()


class NoPositionalEncoding(nn.Module):
    """Identity placeholder (reference :5-21)."""

    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, x):
        return x

    def get_bias(self):
        return None


class AbsolutePositionalEncoding(nn.Module):
    """Learnable absolute table added to the patch tokens, never to the cls token (reference :23-40).

    ``forward`` keeps the reference's in-place semantics (it mutates and returns ``x``); inside
    ``VisionTransformer`` the add is fused into the patch-embedding kernel instead.
    """

    def __init__(self, d_model, max_len=5000):
        super().__init__()
        self.pos_embed = nn.Parameter(torch.zeros(1, max_len, d_model))
        nn.init.trunc_normal_(self.pos_embed, std=.02)

    def forward(self, x):
        n_patch_tokens = x.size(1) - 1
        x[:, 1:] = x[:, 1:] + self.pos_embed[:, :n_patch_tokens]
        return x


class RelativePositionalEncoding(nn.Module):
    """Relative position bias over raster token order, 2L-1 learnable entries per head
    (reference :42-95).  ``relative_position_index[i, j] = i - j + L - 1`` with the cls token at
    position 0; the attention kernel computes that index arithmetically."""

    def __init__(self, num_patches, num_heads=8):
        super().__init__()
        self.num_patches = num_patches
        self.num_heads = num_heads
        self.seq_length = num_patches + 1
        n_entries = 2 * self.seq_length - 1
        self.relative_position_bias_table = nn.Parameter(torch.zeros(num_heads, n_entries))
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)
        pos = torch.arange(self.seq_length)
        index = (pos[:, None] - pos[None, :]) + (self.seq_length - 1)
        self.register_buffer("relative_position_index", index.clamp_(0, n_entries - 1))

    def forward(self, x):
        return x

    def get_bias(self):
        """[num_heads, L, L] gather of the table (reference :82-95)."""
        return self.relative_position_bias_table[:, self.relative_position_index]


class PolynomialRPE(nn.Module):
    """Bias = polynomial in the L1 distance between patch coordinates (reference :97-171)."""

    def __init__(self, num_patches, degree=3, num_heads=8, shared_across_heads=True):
        super().__init__()
        self.num_patches = num_patches
        self.degree = degree
        self.num_heads = num_heads
        self.shared_across_heads = shared_across_heads
        self.grid_size = int(math.sqrt(num_patches))
        shape = (degree + 1,) if shared_across_heads else (num_heads, degree + 1)
        self.coefficients = nn.Parameter(torch.zeros(*shape))
        nn.init.trunc_normal_(self.coefficients, std=0.02)

    def forward(self, x):
        return x

    def l1_distance(self, device=None):
        """int64 [Np, Np]: |dy| + |dx| with y = p % g, x = p // g (reference :136-142)."""
        g = self.grid_size
        p = torch.arange(g * g, device=device)
        y, x = p % g, torch.div(p, g, rounding_mode="floor")
        return (y[:, None] - y[None, :]).abs() + (x[:, None] - x[None, :]).abs()

    def get_bias(self):
        """[num_heads, Np+1, Np+1]; zero on the cls row and column (reference :127-171)."""
        coef = self.coefficients
        dist = self.l1_distance(coef.device).float()
        powers = torch.stack([dist.pow(k) for k in range(self.degree + 1)], dim=-1)
        if self.shared_across_heads:
            core = (powers @ coef).unsqueeze(0).expand(self.num_heads, -1, -1)
        else:
            core = torch.stack([powers @ coef[h] for h in range(self.num_heads)], dim=0)
        n = self.num_patches + 1
        bias = torch.zeros(self.num_heads, n, n, device=coef.device, dtype=core.dtype)
        bias[:, 1:, 1:] = core
        return bias


class RoPEAxial(nn.Module):
    """Axial 2-D RoPE: fixed frequencies, first quarter-pairs follow x, the rest follow y
    (reference :173-245).  ``inv_freq`` is a buffer - this mode has no learnable parameter."""

    def __init__(self, dim, theta=100.0):
        super().__init__()
        self.dim = dim
        self.theta = theta
        n_freq = dim // 4
        inv_freq = 1.0 / (theta ** (torch.arange(0, n_freq, dtype=torch.float) / n_freq))
        self.register_buffer("inv_freq", inv_freq)

    def forward(self, x):
        return x

    def init_t_xy(self, h, w, device):
        return _grid_positions(h, w, device)

    def get_freqs_cis(self, seq_len, device):
        """(cos, sin), each fp32 [seq_len, dim/2], for a square grid of ``seq_len`` patches."""
        side = int(math.sqrt(seq_len))
        t_x, t_y = self.init_t_xy(side, side, device)
        angles = torch.cat([torch.outer(t_x, self.inv_freq), torch.outer(t_y, self.inv_freq)], dim=-1)
        return torch.cos(angles), torch.sin(angles)


class RoPEMixed(nn.Module):
    """Mixed learnable 2-D RoPE frequencies per head (reference :247-351).

    ``freqs`` is [2, num_heads, dim/2].  ``get_freqs_cis`` reproduces the reference's memory
    reinterpretation exactly (reference :337-342): the [H, N, D/2] outer products are *viewed*
    as [N, H, D/2] and permuted back, so output element ``[h', n']`` carries the phase of source
    head ``hs`` at source position ``ps`` with ``(hs, ps) = divmod(n' * H + h', N)``.
    """

    def __init__(self, dim, num_heads, theta=10.0):
        super().__init__()
        self.dim = dim
        self.num_heads = num_heads
        self.theta = theta
        mag = 1 / (theta ** (torch.arange(0, dim, 4)[: (dim // 4)].float() / dim))
        per_head_x, per_head_y = [], []
        for _ in range(num_heads):
            a = torch.rand(1) * 2 * torch.pi  # one random orientation per head
            per_head_x.append(torch.cat([mag * torch.cos(a), mag * torch.cos(torch.pi / 2 + a)], dim=-1))
            per_head_y.append(torch.cat([mag * torch.sin(a), mag * torch.sin(torch.pi / 2 + a)], dim=-1))
        freqs = torch.stack([torch.stack(per_head_x, dim=0), torch.stack(per_head_y, dim=0)], dim=0)
        self.freqs = nn.Parameter(freqs.clone(), requires_grad=True)

    def forward(self, x):
        return x

    def init_t_xy(self, h, w, device):
        return _grid_positions(h, w, device)

    def get_freqs_cis(self, seq_len, device):
        """(cos, sin), each fp32 [num_heads, seq_len, dim/2]; differentiable w.r.t. ``freqs``."""
        side = int(math.sqrt(seq_len))
        t_x, t_y = self.init_t_xy(side, side, device)
        t_x, t_y = t_x.to(self.freqs.device), t_y.to(self.freqs.device)
        with torch.autocast("cuda", enabled=False):  # the table stays fp32 under CUDA autocast (:334)
            # the reference's [N,1] @ [H,1,D/2] outer products (K = 1: one exact product per element), written as
            # broadcast multiplies so that no library GEMM is launched for them
            phase_x = t_x.view(1, -1, 1) * self.freqs[0].unsqueeze(-2)  # [H, N, D/2]
            phase_y = t_y.view(1, -1, 1) * self.freqs[1].unsqueeze(-2)
            phase_x = phase_x.view(seq_len, self.num_heads, -1).permute(1, 0, 2)
            phase_y = phase_y.view(seq_len, self.num_heads, -1).permute(1, 0, 2)
            # (contiguous: the sum of two permuted views keeps their strides, and every block would re-copy the
            # two tables before handing them to the kernels)
            phase = (phase_x + phase_y).contiguous()
            return torch.cos(phase), torch.sin(phase)
