"""Autograd bindings of the libvrr_b200 kernels (host side of the C ABI).

Each ``torch.autograd.Function`` here is the product implementation of one stretch of the
reference's ``Attention.forward`` / ``forward_features`` (citations: /root/reference):

* :class:`PatchEmbedFn`      - models/vit.py:164,248-258 (conv patch embed, cls concat, absolute add)
* :class:`QkvRopeFn`         - models/vit.py:47-68 + models/rope_utils.py:22-35
* :class:`FusedAttentionFn`  - models/vit.py:71-88

PyTorch is used for device memory, streams and autograd plumbing only; all arithmetic on these
paths runs in the hand-written sm_100a kernels.  CUDA only - CPU tensors raise.
"""
import ctypes
import os
import weakref

import torch

from . import _lib

_DT = {torch.float32: _lib.VRR_F32, torch.bfloat16: _lib.VRR_BF16}


def compute_dtype(x: torch.Tensor) -> torch.dtype:
    """Element type the kernels run in: the autocast dtype inside ``torch.autocast('cuda')``
    (how the build defines "bf16 training"), else the tensor's own dtype."""
    dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
    if dt not in _DT:
        raise TypeError(f"vit_rpe_rope_b200 kernels support float32 and bfloat16, got {dt}")
    return dt


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "vit_rpe_rope_b200 runs on CUDA sm_100 devices only (no CPU fallback); got a "
                f"{t.device} tensor")


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t):
    return None if t is None else t.detach().to(torch.float32).contiguous()


# ---- low-precision weight cache ---------------------------------------------------------------------
# Under autocast the reference re-casts every fp32 weight to bf16 in every forward (~60 small cast kernels per
# ViT-B step).  Here the model asks once per forward: the bf16 copies of all its GEMM weights are refreshed by
# ONE multi-tensor copy, and only when a master weight changed (its version counter moved) - i.e. once per
# optimizer step.  The GEMM Functions take the fp32 master (it receives the fp32 weight gradient straight from
# the dW GEMM) plus this copy.
_LP = {}  # id(param) -> [weakref(param), version at last refresh, low-precision copy]


def refresh_lp_weights(params, dt):
    src, dst = [], []
    capturing = torch.cuda.is_current_stream_capturing()
    for p in params:
        if p.dtype == dt or not p.is_cuda:
            continue
        ent = _LP.get(id(p))
        if ent is None or ent[0]() is not p or ent[2].device != p.device or ent[2].dtype != dt or ent[2].shape != p.shape:
            ent = [weakref.ref(p), -1, torch.empty_like(p, dtype=dt, memory_format=torch.contiguous_format)]
            _LP[id(p)] = ent
        if capturing or ent[1] != p._version:  # a captured step must always contain its own refresh
            src.append(p.detach())
            dst.append(ent[2])
            ent[1] = p._version
    if dst:
        torch._foreach_copy_(dst, src)
    if len(_LP) > 4096:  # drop entries of dead parameters
        for k in [k for k, e in _LP.items() if e[0]() is None]:
            del _LP[k]


def lp_weight(p, dt):
    """The cached ``dt`` copy of parameter ``p`` if :func:`refresh_lp_weights` holds a current one, else None."""
    ent = _LP.get(id(p))
    if ent is None or ent[0]() is not p or ent[2].dtype != dt or ent[1] != p._version or ent[2].device != p.device:
        return None
    return ent[2]


# bench.py sets this to a list to time individual kernel launches with CUDA events recorded on the
# launching (current) stream; None (default) adds no overhead.
PROFILE_EVENTS = None


class _timed:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if PROFILE_EVENTS is not None:
            self.start = torch.cuda.Event(enable_timing=True)
            self.start.record()

    def __exit__(self, *exc):
        if PROFILE_EVENTS is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            PROFILE_EVENTS.append((self.name, self.start, end))
        return False


# ------------------------------------------------------------------------------------------------
class PatchEmbedFn(torch.autograd.Function):
    """tokens[B, Np+1, E] from images: unfold + GEMM + bias (+ absolute pos_embed) + cls row.

    weight / bias carry the conv dtype; cls_token / pos_embed / tokens the token-stream dtype (fp32
    under autocast, like the reference's cat with the fp32 cls token); images may stay fp32 when the
    conv runs in bf16 - the rounding is fused into the kernel's load."""

    @staticmethod
    def forward(ctx, images, weight, bias, cls_token, pos_embed, patch):
        _require_cuda(images, weight, bias, cls_token, pos_embed)
        lib = _lib.load()
        B, C, Hi, Wi = images.shape
        E = weight.shape[0]
        idt, dt, tdt = images.dtype, weight.dtype, cls_token.dtype
        if bias.dtype != dt or (pos_embed is not None and pos_embed.dtype != tdt):
            raise TypeError("patch_embed: weight/bias must share a dtype, cls_token/pos_embed another")
        if idt != dt and not (idt == torch.float32 and dt == torch.bfloat16):
            raise TypeError(f"patch_embed: images {idt} with weights {dt} is not supported")
        images = images.contiguous()
        w = weight.detach().contiguous()
        b = bias.detach().contiguous()
        cls = cls_token.detach().reshape(-1).contiguous()
        Np = (Hi // patch) * (Wi // patch)
        pos = None
        if pos_embed is not None:
            if pos_embed.shape[-2] < Np:
                raise RuntimeError(f"absolute pos_embed has {pos_embed.shape[-2]} rows, need {Np}")
            pos = pos_embed.detach().reshape(-1, E).contiguous()
        tokens = torch.empty(B, Np + 1, E, device=images.device, dtype=tdt)
        ws_bytes = int(lib.vrr_patch_embed_workspace_bytes(B, C, Hi, Wi, patch, E, _DT[dt]))
        ws = torch.empty(ws_bytes, device=images.device, dtype=torch.uint8) if ws_bytes else None
        with torch.cuda.device(images.device), _timed("patch_embed_fwd"):
            _lib.check(lib.vrr_patch_embed_fwd(_ptr(images), _ptr(w), _ptr(b), _ptr(cls), _ptr(pos), _ptr(tokens),
                                               _ptr(ws), ws_bytes, B, C, Hi, Wi, patch, E, _DT[idt], _DT[dt],
                                               _DT[tdt], _stream()), "vrr_patch_embed_fwd")
        if ctx.needs_input_grad[0]:
            ctx.w_for_dimg = w
        # the tcgen05 path's workspace IS unfold(images) [B*Np][C*P*P] bf16: the weight-gradient GEMM reuses it
        ctx.unfolded = ws if (ws is not None and ws_bytes == B * Np * C * patch * patch * 2) else None
        ctx.save_for_backward(images)
        ctx.meta = (B, C, Hi, Wi, patch, E, Np, pos_embed.shape if pos_embed is not None else None,
                    weight.shape, cls_token.shape, idt, dt, tdt)
        return tokens

    @staticmethod
    def backward(ctx, d_tokens):
        lib = _lib.load()
        (images,) = ctx.saved_tensors
        B, C, Hi, Wi, patch, E, Np, pos_shape, w_shape, cls_shape, idt, dt, tdt = ctx.meta
        d_tokens = d_tokens.contiguous()
        dev = images.device
        K = C * patch * patch
        # bf16: the weight gradient is the GEMM d_tokens^T . unfold(images) on the tcgen05 kernel (fp32 result,
        # split over the B*Np rows); fp32: the library's FFMA split-K GEMM with the unfold fused in.
        own_dw = not (dt == torch.bfloat16 and patch % 8 == 0 and Wi % 4 == 0 and E % 8 == 0)
        d_w = torch.empty(E, K, device=dev, dtype=torch.float32) if own_dw else None
        d_b = torch.empty(E, device=dev, dtype=torch.float32)
        d_cls = torch.empty(E, device=dev, dtype=torch.float32)
        d_pos_rows = torch.empty(Np, E, device=dev, dtype=torch.float32) if pos_shape is not None else None
        with torch.cuda.device(dev), _timed("patch_embed_bwd"):
            _lib.check(lib.vrr_patch_embed_bwd(_ptr(images), _ptr(d_tokens), _ptr(d_w), _ptr(d_b), _ptr(d_cls),
                                               _ptr(d_pos_rows), B, C, Hi, Wi, patch, E, _DT[idt], _DT[dt], _DT[tdt],
                                               _stream()), "vrr_patch_embed_bwd")
            if not own_dw:
                if ctx.unfolded is not None:
                    unf = ctx.unfolded.view(torch.bfloat16).view(B * Np, K)
                else:
                    unf = torch.empty(B * Np, K, device=dev, dtype=torch.bfloat16)
                    _lib.check(lib.vrr_patch_unfold(_ptr(images), _ptr(unf), B, C, Hi, Wi, patch, _DT[idt], _stream()),
                               "vrr_patch_unfold")
                # cast first: one strided read + packed bf16 write (reshaping the fp32 slice first copied 154 MB twice)
                g = d_tokens[:, 1:, :].to(torch.bfloat16).reshape(B * Np, E)
                d_w = _gemm(g, unf, True, False, torch.float32, name="patch_dw")
        d_pos = None
        if pos_shape is not None:
            d_pos = torch.zeros(pos_shape, device=dev, dtype=tdt)
            d_pos.view(-1, E)[:Np] = d_pos_rows.to(tdt)
        return (None, d_w.view(w_shape).to(dt), d_b.to(dt), d_cls.view(cls_shape).to(tdt), d_pos, None)


def patch_embed(images, weight, bias, cls_token, pos_embed, patch):
    """Conv weights are cast to the compute dtype (autograd-tracked casts keep fp32 master grads);
    under autocast the token stream (cls, pos, output) stays fp32 as in the reference, and fp32 images
    are passed as they are (rounded to bf16 inside the kernel)."""
    dt = compute_dtype(images)
    tdt = torch.float32 if torch.is_autocast_enabled("cuda") else dt
    if not (images.dtype == torch.float32 and dt == torch.bfloat16):
        images = images.to(dt)
    return PatchEmbedFn.apply(images, weight.to(dt), bias.to(dt), cls_token.to(tdt),
                              None if pos_embed is None else pos_embed.to(tdt), patch)


# ------------------------------------------------------------------------------------------------
def _rope_args(cos, sin, H, N, Dh):
    """Classify the (cos, sin) pair the way ``reshape_for_broadcast`` does (rope_utils.py:39-66)."""
    if cos is None:
        return _lib.ROPE_NONE
    if cos.shape != sin.shape:
        raise RuntimeError(f"cos {tuple(cos.shape)} and sin {tuple(sin.shape)} differ")
    if cos.ndim == 2:
        mode, want = _lib.ROPE_AXIAL, (N - 1, Dh // 2)
    elif cos.ndim == 3:
        mode, want = _lib.ROPE_MIXED, (H, N - 1, Dh // 2)
    else:
        raise ValueError(f"Unexpected tensor shapes: {cos.shape} vs {(1, H, N - 1, Dh)}")
    if tuple(cos.shape) != want:
        raise RuntimeError(f"RoPE table shape {tuple(cos.shape)} does not broadcast against the "
                           f"{N - 1} patch tokens x {H} heads x {Dh} channels (expected {want})")
    return mode


_packed_cache = {}  # device index -> (key, cos32, sin32, packed); the table tensors are held so their storage stays theirs


def _packed_tables(cos32, sin32):
    """``vrr_rope_pack_tables`` of the fp32 tables ([heads][half/2][rows] float4), computed once per table pair: every
    block of a forward pass receives the SAME cos / sin storage (one shared PE module), so the repack runs once per
    step (and once inside a captured graph).  Keyed on storage address + version counter; the cache keeps the tables
    alive, so an address cannot be recycled for other data while its entry exists."""
    dev = cos32.device.index
    try:
        key = (cos32.data_ptr(), sin32.data_ptr(), cos32._version, sin32._version, tuple(cos32.shape))
    except RuntimeError:  # tensors made under torch.inference_mode() carry no version counter: repack every call
        key = None
    hit = _packed_cache.get(dev)
    if key is not None and hit is not None and hit[0] == key:
        return hit[3]
    heads = cos32.shape[0] if cos32.ndim == 3 else 1
    rows, half = cos32.shape[-2], cos32.shape[-1]
    packed = torch.empty(heads, half // 2, rows, 4, device=cos32.device, dtype=torch.float32)
    _lib.check(_lib.load().vrr_rope_pack_tables(_ptr(cos32), _ptr(sin32), _ptr(packed), heads, rows, half, _stream()),
               "vrr_rope_pack_tables")
    if key is not None:
        _packed_cache[dev] = (key, cos32, sin32, packed)
    return packed


class QkvRopeFn(torch.autograd.Function):
    """planes[3, B, H, N, Dh] = split_heads(x @ w_qkv^T), q/k rows 1.. rotated in the GEMM epilogue."""

    @staticmethod
    def forward(ctx, x, w_qkv, cos, sin, num_heads, w_lp=None):
        """``w_qkv`` receives the gradient (fp32 master weight or a tensor already in x's dtype); ``w_lp`` is an
        optional ready-made copy of it in x's dtype (the model's bf16 weight cache)."""
        _require_cuda(x, w_qkv, cos, sin)
        lib = _lib.load()
        B, N, E = x.shape
        H, Dh = num_heads, E // num_heads
        mode = _rope_args(cos, sin, H, N, Dh)
        x = x.contiguous()
        w = (w_lp if w_lp is not None else w_qkv.detach().to(x.dtype)).detach().contiguous()
        ctx.w_grad_dtype = w_qkv.dtype
        cos32, sin32 = _f32c(cos), _f32c(sin)
        planes = torch.empty(3, B, H, N, Dh, device=x.device, dtype=x.dtype)
        with torch.cuda.device(x.device):
            packed = None
            if mode != _lib.ROPE_NONE and x.dtype == torch.bfloat16 and Dh == 64:  # the tcgen05 epilogue's table layout
                packed = _packed_tables(cos32, sin32)
            with _timed("qkv_rope_fwd"):
                _lib.check(lib.vrr_qkv_rope_fwd_packed(_ptr(x), _ptr(w), _ptr(cos32), _ptr(sin32), _ptr(packed), _ptr(planes),
                                                       B, N, E, H, mode, _DT[x.dtype], _stream()), "vrr_qkv_rope_fwd_packed")
        ctx.save_for_backward(x, w, planes, cos32, sin32)
        ctx.meta = (B, N, E, H, mode, None if cos is None else (cos.dtype, sin.dtype))
        return planes

    @staticmethod
    def backward(ctx, d_planes):
        lib = _lib.load()
        x, w, planes, cos32, sin32 = ctx.saved_tensors
        B, N, E, H, mode, cs_dtypes = ctx.meta
        d_planes = d_planes.contiguous()
        dev, dt = x.device, x.dtype
        need_cs = mode != _lib.ROPE_NONE and (ctx.needs_input_grad[2] or ctx.needs_input_grad[3])
        d_cos = d_sin = None
        if need_cs:  # halves of one buffer: zeroed by one memset
            d_cs = torch.empty((2,) + tuple(cos32.shape), device=dev, dtype=torch.float32)
            d_cos, d_sin = d_cs[0], d_cs[1]
        d_qkv = torch.empty(B * N, 3 * E, device=dev, dtype=dt)
        with torch.cuda.device(dev):
            _lib.check(lib.vrr_qkv_rope_bwd(_ptr(d_planes), _ptr(planes), _ptr(cos32), _ptr(sin32), _ptr(d_qkv),
                                            _ptr(d_cos), _ptr(d_sin), B, N, E, H, mode, _DT[dt], _stream()),
                       "vrr_qkv_rope_bwd")
            dx = dw = None
            if ctx.needs_input_grad[0]:
                dx = _gemm(d_qkv, w, False, False, dt, name="qkv_dx").view(B, N, E)            # [BN,3E] . [3E,E]
            if ctx.needs_input_grad[1]:
                dw = _gemm(d_qkv, x.view(B * N, E), True, False, ctx.w_grad_dtype, name="qkv_dw")  # [3E,BN] . [BN,E]
        if need_cs:
            d_cos, d_sin = d_cos.to(cs_dtypes[0]), d_sin.to(cs_dtypes[1])
        return dx, dw, d_cos, d_sin, None, None


def _gemm(a, b, trans_a, trans_b, out_dtype, bias=None, epilogue=_lib.EPI_NONE, out2=False, name="gemm", aux=None):
    """C = op(A) . op(B) in the library's kernels (``vrr_gemm_ex``) - no cuBLAS anywhere on the path.

    op(A) is [M][K] (trans_a: A stored [K][M]); op(B) is [K][N] (trans_b: B stored [N][K]).  bf16 operands
    run the tcgen05 CTA-pair kernel (fp32 accumulation; result in ``out_dtype``: bf16, or fp32 for weight
    gradients), fp32 operands the exact-fp32 FFMA kernel (no TF32).  ``epilogue``: fused ``+ bias`` or
    ``+ bias, GELU`` (``out2``: second output - gelu(C) for EPI_BIAS_GELU; EPI_BIAS_GELU_GRAD returns
    (gelu(h), gelu'(h)) instead; EPI_BIAS_GELU_ACT returns gelu(h) alone); ``aux``: the [M][N] multiplier of EPI_MUL."""
    lib = _lib.load()
    M = a.shape[1] if trans_a else a.shape[0]
    K = a.shape[0] if trans_a else a.shape[1]
    N = b.shape[0] if trans_b else b.shape[1]
    a, b = a.contiguous(), b.contiguous()
    c = torch.empty(M, N, device=a.device, dtype=out_dtype)
    c2 = torch.empty_like(c) if out2 else (aux.contiguous() if aux is not None else None)
    bias32 = _f32c(bias) if bias is not None and epilogue != _lib.EPI_NONE else None
    with _timed(name):
        _lib.check(lib.vrr_gemm_ex(_ptr(a), _ptr(b), _ptr(c), _ptr(c2), _ptr(bias32), M, N, K, int(trans_a), int(trans_b),
                                   _DT[a.dtype], _DT[out_dtype], int(epilogue), 0, _stream()), "vrr_gemm_ex")
    return (c, c2) if out2 else c


def qkv_rope(x, w_qkv, cos, sin, num_heads, w_lp=None):
    dt = compute_dtype(x)
    if w_lp is not None and w_lp.dtype != dt:
        w_lp = None
    return QkvRopeFn.apply(x.to(dt), w_qkv, cos, sin, num_heads, w_lp)


# ------------------------------------------------------------------------------------------------
class FusedAttentionFn(torch.autograd.Function):
    """out[B, N, E] = merge_heads(softmax(q k^T * scale + bias) v) from the qkv planes."""

    @staticmethod
    def forward(ctx, planes, bias_param, bias_mode, bias_grid, scale):
        _require_cuda(planes, bias_param)
        lib = _lib.load()
        _, B, H, N, Dh = planes.shape
        planes = planes.contiguous()
        bp = _f32c(bias_param)
        desc = _bias_desc(bias_mode, bp, bias_grid)
        out = torch.empty(B, N, H * Dh, device=planes.device, dtype=planes.dtype)
        lse = torch.empty(B, H, N, device=planes.device, dtype=torch.float32)
        with torch.cuda.device(planes.device), _timed("attn_fwd"):
            _lib.check(lib.vrr_attn_fwd(_ptr(planes), ctypes.byref(desc), _ptr(out), _ptr(lse), B, H, N, Dh,
                                        float(scale), _DT[planes.dtype], _stream()), "vrr_attn_fwd")
        ctx.save_for_backward(planes, out, lse, bp)
        ctx.meta = (B, H, N, Dh, float(scale), bias_mode, bias_grid,
                    None if bias_param is None else (bias_param.shape, bias_param.dtype))
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        planes, out, lse, bp = ctx.saved_tensors
        B, H, N, Dh, scale, bias_mode, bias_grid, bp_meta = ctx.meta
        d_out = d_out.contiguous()
        dev = planes.device
        desc = _bias_desc(bias_mode, bp, bias_grid)
        d_planes = torch.empty_like(planes)
        d_bp = torch.empty_like(bp) if bp is not None else None
        ws_bytes = int(lib.vrr_attn_bwd_workspace_bytes(B, H, N, Dh, ctypes.byref(desc)))
        ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev), _timed("attn_bwd"):
            _lib.check(lib.vrr_attn_bwd(_ptr(planes), ctypes.byref(desc), _ptr(out), _ptr(d_out), _ptr(lse),
                                        _ptr(d_planes), _ptr(d_bp), _ptr(ws), ws_bytes, B, H, N, Dh, scale,
                                        _DT[planes.dtype], _stream()), "vrr_attn_bwd")
        if d_bp is not None:
            d_bp = d_bp.view(bp_meta[0]).to(bp_meta[1])
        return d_planes, d_bp, None, None, None


def _bias_desc(mode, param32, grid):
    d = _lib.BiasDesc()
    d.mode = mode
    if mode == _lib.BIAS_NONE:
        d.heads = d.len = d.grid = 0
        d.param = None
        return d
    p2 = param32.reshape(-1, param32.shape[-1])
    d.heads, d.len, d.grid = p2.shape[0], p2.shape[1], int(grid)
    d.param = param32.data_ptr()
    return d


def fused_attention(planes, scale, bias_mode=_lib.BIAS_NONE, bias_param=None, bias_grid=0):
    return FusedAttentionFn.apply(planes, bias_param, bias_mode, bias_grid, scale)


# ------------------------------------------------------------------------------------------------
class LayerNormFn(torch.autograd.Function):
    """y = LayerNorm(x) over the last dim, written directly in the consumer's dtype (fp32 statistics)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, out_dtype):
        _require_cuda(x, weight, bias)
        lib = _lib.load()
        E = x.shape[-1]
        x2 = x.contiguous().view(-1, E)
        M = x2.shape[0]
        w32, b32 = _f32c(weight), _f32c(bias)
        y = torch.empty(M, E, device=x.device, dtype=out_dtype)
        mean = torch.empty(M, device=x.device, dtype=torch.float32)
        rstd = torch.empty(M, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device), _timed("layernorm_fwd"):
            _lib.check(lib.vrr_layernorm_fwd(_ptr(x2), _ptr(w32), _ptr(b32), _ptr(y), _ptr(mean), _ptr(rstd), M, E,
                                             float(eps), _DT[x2.dtype], _DT[out_dtype], _stream()),
                       "vrr_layernorm_fwd")
        ctx.save_for_backward(x2, w32, mean, rstd)
        ctx.meta = (x.shape, weight.dtype, bias.dtype, out_dtype)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        lib = _lib.load()
        x2, w32, mean, rstd = ctx.saved_tensors
        shape, w_dtype, b_dtype, out_dtype = ctx.meta
        M, E = x2.shape
        dy2 = dy.contiguous().view(M, E)
        if dy2.dtype != out_dtype:
            dy2 = dy2.to(out_dtype)
        dx = torch.empty_like(x2)
        dgb = torch.empty(2, E, device=x2.device, dtype=torch.float32)  # one buffer: zeroed by one memset
        dg, db = dgb[0], dgb[1]
        with torch.cuda.device(x2.device), _timed("layernorm_bwd"):
            _lib.check(lib.vrr_layernorm_bwd(_ptr(dy2), _ptr(x2), _ptr(w32), _ptr(mean), _ptr(rstd), _ptr(dx), _ptr(dg),
                                             _ptr(db), M, E, _DT[x2.dtype], _DT[out_dtype], _stream()),
                       "vrr_layernorm_bwd")
        return dx.view(shape), dg.to(w_dtype), db.to(b_dtype), None, None


class AddLayerNormFn(torch.autograd.Function):
    """(x_new, y) = (x + branch, LayerNorm(x + branch)): the residual add of one sub-block fused with the
    LayerNorm that opens the next one; the backward emits the residual-stream gradient (fp32) and the
    branch gradient (branch dtype) from a single pass."""

    @staticmethod
    def forward(ctx, x, branch, weight, bias, eps, out_dtype):
        _require_cuda(x, branch, weight, bias)
        lib = _lib.load()
        E = x.shape[-1]
        x2 = x.contiguous().view(-1, E)
        b2 = branch.contiguous().view(-1, E)
        M = x2.shape[0]
        w32, b32 = _f32c(weight), _f32c(bias)
        x_new = torch.empty_like(x2)
        y = torch.empty(M, E, device=x.device, dtype=out_dtype)
        mean = torch.empty(M, device=x.device, dtype=torch.float32)
        rstd = torch.empty(M, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device), _timed("add_layernorm_fwd"):
            _lib.check(lib.vrr_add_layernorm_fwd(_ptr(x2), _ptr(b2), _ptr(x_new), _ptr(w32), _ptr(b32), _ptr(y),
                                                 _ptr(mean), _ptr(rstd), M, E, float(eps), _DT[b2.dtype],
                                                 _DT[out_dtype], _stream()), "vrr_add_layernorm_fwd")
        ctx.save_for_backward(x_new, w32, mean, rstd)
        ctx.meta = (x.shape, weight.dtype, bias.dtype, out_dtype, b2.dtype)
        return x_new.view(x.shape), y.view(x.shape)

    @staticmethod
    def backward(ctx, d_xnew, d_y):
        lib = _lib.load()
        x_new, w32, mean, rstd = ctx.saved_tensors
        shape, w_dtype, b_dtype, out_dtype, br_dtype = ctx.meta
        M, E = x_new.shape
        if d_y is None:
            d_y = torch.zeros(M, E, device=x_new.device, dtype=out_dtype)
        dy2 = d_y.contiguous().view(M, E)
        if dy2.dtype != out_dtype:
            dy2 = dy2.to(out_dtype)
        dr2 = None
        if d_xnew is not None:
            dr2 = d_xnew.contiguous().view(M, E)
            if dr2.dtype != torch.float32:
                dr2 = dr2.float()
        dx = torch.empty_like(x_new)
        d_branch = torch.empty(M, E, device=x_new.device, dtype=br_dtype)
        dgb = torch.empty(2, E, device=x_new.device, dtype=torch.float32)  # one buffer: the library zeroes both rows with one memset
        dg, db = dgb[0], dgb[1]
        with torch.cuda.device(x_new.device), _timed("add_layernorm_bwd"):
            _lib.check(lib.vrr_add_layernorm_bwd(_ptr(dy2), _ptr(dr2), _ptr(x_new), _ptr(w32), _ptr(mean), _ptr(rstd),
                                                 _ptr(dx), _ptr(d_branch), _ptr(dg), _ptr(db), M, E, _DT[br_dtype],
                                                 _DT[out_dtype], _stream()), "vrr_add_layernorm_bwd")
        return dx.view(shape), d_branch.view(shape), dg.to(w_dtype), db.to(b_dtype), None, None


def can_fuse_add_layer_norm(x, norm) -> bool:
    """The fused add+LayerNorm kernels need an fp32 residual stream and a plain affine last-dim LayerNorm."""
    return (x.is_cuda and x.dtype == torch.float32 and type(norm) is torch.nn.LayerNorm and norm.elementwise_affine
            and norm.bias is not None and tuple(norm.normalized_shape) == (x.shape[-1],))


def add_layer_norm(x, branch, norm: torch.nn.LayerNorm):
    """``x_new = x + branch; y = norm(x_new)`` in one kernel (see :class:`AddLayerNormFn`)."""
    out_dtype = compute_dtype(x)
    if branch.dtype not in (torch.float32, out_dtype):
        branch = branch.to(out_dtype)
    return AddLayerNormFn.apply(x, branch, norm.weight, norm.bias, norm.eps, out_dtype)


def layer_norm(x, norm: torch.nn.LayerNorm):
    """``norm(x)`` for an ``nn.LayerNorm`` over the last dimension, output in the compute dtype
    (bf16 under autocast: the reference's LayerNorm-then-cast, fused).  Falls through to the module
    itself for anything that is not a plain affine last-dim LayerNorm (a caller-supplied norm_layer)."""
    if (type(norm) is not torch.nn.LayerNorm or not norm.elementwise_affine or norm.bias is None
            or tuple(norm.normalized_shape) != (x.shape[-1],)):
        return norm(x)
    out_dtype = compute_dtype(x)
    if x.dtype == torch.bfloat16 and out_dtype == torch.float32:
        x = x.float()
    return LayerNormFn.apply(x, norm.weight, norm.bias, norm.eps, out_dtype)


# ------------------------------------------------------------------------------------------------
def _colsum(t2d):
    lib = _lib.load()
    M, C = t2d.shape
    out = torch.empty(C, device=t2d.device, dtype=torch.float32)
    _lib.check(lib.vrr_colsum(_ptr(t2d), _ptr(out), M, C, _DT[t2d.dtype], _stream()), "vrr_colsum")
    return out


class LinearFn(torch.autograd.Function):
    """y = x W^T + b (out-proj / head, models/vit.py:91,285).  bf16: forward GEMM with the bias in its
    epilogue, dX and dW (fp32, straight into the master-weight gradient) on the tcgen05 CTA-pair kernel;
    bias gradient on the column-sum kernel.  ``w`` / ``b`` receive the gradients (fp32 masters or tensors in
    x's dtype); ``w_lp`` is an optional ready-made copy of ``w`` in x's dtype."""

    @staticmethod
    def forward(ctx, x, w, b, w_lp=None):
        _require_cuda(x, w, b)
        dt = x.dtype
        wl = (w_lp if w_lp is not None else w.detach().to(dt)).detach()
        x2 = x.reshape(-1, x.shape[-1])
        y = _gemm(x2, wl, False, True, dt, bias=b, epilogue=_lib.EPI_BIAS if b is not None else _lib.EPI_NONE,
                  name="linear_fwd")
        ctx.save_for_backward(x2, wl)
        ctx.meta = (x.shape, w.dtype, None if b is None else b.dtype)
        return y.view(*x.shape[:-1], wl.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, wl = ctx.saved_tensors
        x_shape, w_dtype, b_dtype = ctx.meta
        dy2 = dy.contiguous().view(-1, dy.shape[-1])
        if dy2.dtype != x2.dtype:
            dy2 = dy2.to(x2.dtype)
        dx = dw = db = None
        with torch.cuda.device(dy.device):
            if ctx.needs_input_grad[0]:
                dx = _gemm(dy2, wl, False, False, x2.dtype, name="linear_dx").view(x_shape)
            if ctx.needs_input_grad[1]:
                dw = _gemm(dy2, x2, True, False, w_dtype, name="linear_dw")
            if b_dtype is not None and ctx.needs_input_grad[2]:
                with _timed("colsum"):
                    db = _colsum(dy2).to(b_dtype)
        return dx, dw, db, None


_FUSE_DB1 = os.environ.get("VRR_FUSE_DB1", "1") != "0"  # A/B switch: fc1's bias gradient inside fc2's dX GEMM


class MlpFn(torch.autograd.Function):
    """fc1 -> exact GELU -> fc2 (timm Mlp, models/vit.py:118).  fc1's GEMM epilogue adds the bias and writes BOTH
    gelu(h) and gelu'(h) (h itself is never stored: the backward only needs the derivative); fc2 adds its bias.
    Backward = four GEMMs on the same kernel (dW in fp32); the GELU backward is the epilogue of fc2's dX GEMM
    (dh = (dy . W2) * gelu'(h)), the two bias gradients are column sums."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, w1_lp=None, w2_lp=None):
        _require_cuda(x, w1, b1, w2, b2)
        dt = x.dtype
        w1l = (w1_lp if w1_lp is not None else w1.detach().to(dt)).detach()
        w2l = (w2_lp if w2_lp is not None else w2.detach().to(dt)).detach()
        x2 = x.reshape(-1, x.shape[-1])
        if not any(ctx.needs_input_grad):  # inference: the activation alone - one output stream instead of two
            a = _gemm(x2, w1l, False, True, dt, bias=b1, epilogue=_lib.EPI_BIAS_GELU_ACT, name="fc1_fwd")
            y = _gemm(a, w2l, False, True, dt, bias=b2, epilogue=_lib.EPI_BIAS, name="fc2_fwd")
            return y.view(*x.shape[:-1], w2l.shape[0])
        a, gp = _gemm(x2, w1l, False, True, dt, bias=b1, epilogue=_lib.EPI_BIAS_GELU_GRAD, out2=True, name="fc1_fwd")
        y = _gemm(a, w2l, False, True, dt, bias=b2, epilogue=_lib.EPI_BIAS, name="fc2_fwd")
        ctx.save_for_backward(x2, gp, a, w1l, w2l)
        ctx.meta = (x.shape, w1.dtype, b1.dtype, w2.dtype, b2.dtype)
        return y.view(*x.shape[:-1], w2l.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, gp, a2, w1l, w2l = ctx.saved_tensors
        x_shape, w1_dt, b1_dt, w2_dt, b2_dt = ctx.meta
        dy2 = dy.contiguous().view(-1, dy.shape[-1])
        if dy2.dtype != x2.dtype:
            dy2 = dy2.to(x2.dtype)
        with torch.cuda.device(dy.device):
            # dh = (dy . W2) * gelu'(h) with fc1's bias gradient (column sums of dh as stored, like autograd's sum
            # over the rounded tensor) accumulated by the same GEMM's epilogue
            if _FUSE_DB1:
                lib = _lib.load()
                M, N, K = dy2.shape[0], w2l.shape[1], dy2.shape[1]
                dh = torch.empty(M, N, device=dy2.device, dtype=x2.dtype)
                db1 = torch.empty(N, device=dy2.device, dtype=torch.float32)
                with _timed("fc2_dx"):
                    _lib.check(lib.vrr_gemm_mul_colsum(_ptr(dy2), _ptr(w2l), _ptr(dh), _ptr(gp), _ptr(db1), M, N, K, 0, 0,
                                                       _DT[x2.dtype], _stream()), "vrr_gemm_mul_colsum")
                db1 = db1.to(b1_dt)
            else:
                dh = _gemm(dy2, w2l, False, False, x2.dtype, epilogue=_lib.EPI_MUL, aux=gp, name="fc2_dx")
                with _timed("colsum"):
                    db1 = _colsum(dh).to(b1_dt)
            dw2 = _gemm(dy2, a2, True, False, w2_dt, name="fc2_dw")
            with _timed("colsum"):
                db2 = _colsum(dy2).to(b2_dt)
            dx = _gemm(dh, w1l, False, False, x2.dtype, name="fc1_dx").view(x_shape)
            dw1 = _gemm(dh, x2, True, False, w1_dt, name="fc1_dw")
        return dx, dw1, db1, dw2, db2, None, None


def linear(x, lin: torch.nn.Linear, w_lp=None):
    """``lin(x)`` through :class:`LinearFn` (CUDA, bias, supported dtype, width % 4 == 0) else the module."""
    if not (x.is_cuda and lin.bias is not None and lin.out_features % 4 == 0 and x.dtype in _DT):
        return lin(x)
    dt = compute_dtype(x)
    if w_lp is not None and w_lp.dtype != dt:
        w_lp = None
    return LinearFn.apply(x.to(dt), lin.weight, lin.bias, w_lp)


def mlp(x, fc1: torch.nn.Linear, fc2: torch.nn.Linear, w1_lp=None, w2_lp=None):
    dt = compute_dtype(x)
    if w1_lp is not None and (w1_lp.dtype != dt or w2_lp is None or w2_lp.dtype != dt):
        w1_lp = w2_lp = None
    return MlpFn.apply(x.to(dt), fc1.weight, fc1.bias, fc2.weight, fc2.bias, w1_lp, w2_lp)


def can_fuse_mlp(x, m) -> bool:
    act = getattr(m, "act", None)
    return (x.is_cuda and x.dtype in _DT and type(act) is torch.nn.GELU and getattr(act, "approximate", "none") == "none"
            and m.fc1.bias is not None and m.fc2.bias is not None and m.fc1.out_features % 4 == 0
            and m.fc2.out_features % 4 == 0 and getattr(m.drop1, "p", 0.0) == 0.0 and getattr(m.drop2, "p", 0.0) == 0.0)


# ------------------------------------------------------------------------------------------------
class RopeApplyFn(torch.autograd.Function):
    """Stand-alone rotate-half of q and k (public ``apply_rotary_emb``, rope_utils.py:3-37)."""

    @staticmethod
    def forward(ctx, q, k, cos, sin):
        _require_cuda(q, k, cos, sin)
        lib = _lib.load()
        B, H, Nr, Dh = q.shape
        cos_b = torch.broadcast_to(cos, (1, cos.shape[1], Nr, Dh // 2))[0] if cos.ndim == 4 else cos
        sin_b = torch.broadcast_to(sin, (1, sin.shape[1], Nr, Dh // 2))[0] if sin.ndim == 4 else sin
        if cos_b.ndim == 3 and cos_b.shape[0] == 1:
            cos_b, sin_b = cos_b[0], sin_b[0]
        mode = _lib.ROPE_AXIAL if cos_b.ndim == 2 else _lib.ROPE_MIXED
        want = (Nr, Dh // 2) if mode == _lib.ROPE_AXIAL else (H, Nr, Dh // 2)
        if tuple(cos_b.shape) != want or tuple(sin_b.shape) != want:
            raise RuntimeError(f"cos/sin {tuple(cos.shape)} do not broadcast against q {tuple(q.shape)}")
        q, k = q.contiguous(), k.contiguous()
        c32, s32 = _f32c(cos_b), _f32c(sin_b)
        qo, ko = torch.empty_like(q), torch.empty_like(k)
        with torch.cuda.device(q.device):
            _lib.check(lib.vrr_rope_apply(_ptr(q), _ptr(k), _ptr(c32), _ptr(s32), _ptr(qo), _ptr(ko), B, H, Nr, Dh,
                                          mode, 0, _DT[q.dtype], _stream()), "vrr_rope_apply")
        need_tab = ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        ctx.save_for_backward(c32, s32, *((q, k) if need_tab else ()))
        ctx.meta = (B, H, Nr, Dh, mode, cos.shape, sin.shape, cos.dtype, sin.dtype)
        return qo, ko

    @staticmethod
    def backward(ctx, dq, dk):
        lib = _lib.load()
        c32, s32 = ctx.saved_tensors[:2]
        B, H, Nr, Dh, mode, cos_shape, sin_shape, cos_dt, sin_dt = ctx.meta
        dq, dk = dq.contiguous(), dk.contiguous()
        gq, gk = torch.empty_like(dq), torch.empty_like(dk)
        d_cos = d_sin = None
        with torch.cuda.device(dq.device):
            _lib.check(lib.vrr_rope_apply(_ptr(dq), _ptr(dk), _ptr(c32), _ptr(s32), _ptr(gq), _ptr(gk), B, H, Nr,
                                          Dh, mode, 1, _DT[dq.dtype], _stream()), "vrr_rope_apply")
            if len(ctx.saved_tensors) == 4:  # the tables are differentiable too (RoPEMixed.freqs feeds them)
                q, k = ctx.saved_tensors[2:]
                d_cos, d_sin = torch.empty_like(c32), torch.empty_like(s32)
                _lib.check(lib.vrr_rope_table_grad(_ptr(q), _ptr(k), _ptr(dq), _ptr(dk), _ptr(d_cos), _ptr(d_sin), B, H,
                                                   Nr, Dh, mode, _DT[dq.dtype], _stream()), "vrr_rope_table_grad")
                # undo the broadcast of reshape_for_broadcast: sum_to_size restores [1, 1|H, Nr, Dh/2] / the raw shape
                d_cos = d_cos.reshape((1,) * (len(cos_shape) - d_cos.ndim) + tuple(d_cos.shape)).sum_to_size(cos_shape).to(cos_dt)
                d_sin = d_sin.reshape((1,) * (len(sin_shape) - d_sin.ndim) + tuple(d_sin.shape)).sum_to_size(sin_shape).to(sin_dt)
        return gq, gk, (d_cos if ctx.needs_input_grad[2] else None), (d_sin if ctx.needs_input_grad[3] else None)
