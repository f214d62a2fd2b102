"""Run the hot-path kernels at ViT-B/16-224 size a few times (for ncu captures; not a benchmark)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_rpe_rope_b200 import _lib, ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which == "new":
    # round-2b kernels: long-sequence forward (1025 tokens), register-resident add+LayerNorm, RoPE backward, column sum
    from vit_rpe_rope_b200.ops import _ptr, _DT, _stream
    lib = _lib.load()
    dev = "cuda:0"
    dt = torch.bfloat16
    pl = (torch.randn(3, 32, 12, 1025, 64) * 0.8).to(dt).to(dev)
    M, E = 256 * 197, 768
    x = torch.randn(M, E, device=dev); br = torch.randn(M, E, device=dev).to(dt)
    gam = torch.randn(E, device=dev); bet = torch.randn(E, device=dev)
    xn = torch.empty_like(x); y = torch.empty(M, E, device=dev, dtype=dt)
    mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
    dy = torch.randn(M, E, device=dev).to(dt); dxn = torch.randn(M, E, device=dev)
    dx = torch.empty_like(x); dbr = torch.empty(M, E, device=dev, dtype=dt)
    dg = torch.empty(E, device=dev); db = torch.empty(E, device=dev)
    planes = torch.randn(3, 256, 12, 197, 64, device=dev).to(dt); d_planes = torch.randn(3, 256, 12, 197, 64, device=dev).to(dt)
    cos = torch.rand(12, 196, 32, device=dev); sin = torch.rand(12, 196, 32, device=dev)
    d_cos = torch.empty_like(cos); d_sin = torch.empty_like(sin)
    d_qkv = torch.empty(M, 3 * E, device=dev, dtype=dt)
    dh = torch.randn(M, 4 * E, device=dev).to(dt); dbias = torch.empty(4 * E, device=dev)
    for it in range(3):
        ops.fused_attention(pl, 0.125)
        _lib.check(lib.vrr_add_layernorm_fwd(_ptr(x), _ptr(br), _ptr(xn), _ptr(gam), _ptr(bet), _ptr(y), _ptr(mean), _ptr(rstd), M, E,
                                             1e-6, _DT[dt], _DT[dt], _stream()), "add_ln_fwd")
        _lib.check(lib.vrr_add_layernorm_bwd(_ptr(dy), _ptr(dxn), _ptr(xn), _ptr(gam), _ptr(mean), _ptr(rstd), _ptr(dx), _ptr(dbr),
                                             _ptr(dg), _ptr(db), M, E, _DT[dt], _DT[dt], _stream()), "add_ln_bwd")
        _lib.check(lib.vrr_qkv_rope_bwd(_ptr(d_planes), _ptr(planes), _ptr(cos), _ptr(sin), _ptr(d_qkv), _ptr(d_cos), _ptr(d_sin),
                                        256, 197, E, 12, 2, _DT[dt], _stream()), "rope_bwd")
        _lib.check(lib.vrr_colsum(_ptr(dh), _ptr(dbias), M, 4 * E, _DT[dt], _stream()), "colsum")
    torch.cuda.synchronize()
    print("done", _lib.launch_count(), "launches")
    sys.exit(0)
B, H, N, D = 256, 12, 197, 64
E = H * D
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
x = torch.randn(B, N, E, generator=g).to(torch.bfloat16).to(dev).requires_grad_(True)
w = (torch.randn(3 * E, E, generator=g) * E ** -0.5).to(torch.bfloat16).to(dev).requires_grad_(True)
ang = torch.rand(H, N - 1, D // 2, generator=g) * 6
cos, sin = torch.cos(ang).to(dev), torch.sin(ang).to(dev)
d_out = torch.randn(B, N, E, generator=g).to(torch.bfloat16).to(dev)
table = (torch.randn(H, 2 * N - 1, generator=g) * 0.5).to(dev)
for it in range(3):
    planes = ops.QkvRopeFn.apply(x, w, cos, sin, H)
    if which in ("all", "rope"):
        o = ops.fused_attention(planes, D ** -0.5)
        o.backward(d_out)
    if which in ("all", "table"):
        pl = planes.detach().requires_grad_(True)
        o = ops.fused_attention(pl, D ** -0.5, _lib.BIAS_TABLE, table.requires_grad_(True), 0)
        o.backward(d_out)
torch.cuda.synchronize()
print("done", _lib.launch_count(), "launches")
