"""Time the hot-path kernels in isolation at a BASELINE geometry (CUDA events, inputs > L2).

    python scripts/kbench.py [vitb|vitl|x512]
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_rpe_rope_b200 import _lib, ops  # noqa: E402

geo = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "vitb"
B, H, N, D = {"vitb": (256, 12, 197, 64), "vitl": (64, 16, 577, 64), "x512": (32, 12, 1025, 64)}[geo]
E = H * D
dev = "cuda:0"
lib = _lib.load()
g = torch.Generator().manual_seed(0)
x = torch.randn(B, N, E, generator=g).to(torch.bfloat16).to(dev)
w = (torch.randn(3 * E, E, generator=g) * E ** -0.5).to(torch.bfloat16).to(dev)
ang = torch.rand(H, N - 1, D // 2, generator=g) * 6
cos, sin = torch.cos(ang).to(dev), torch.sin(ang).to(dev)
d_out = torch.randn(B, N, E, generator=g).to(torch.bfloat16).to(dev)
table = (torch.randn(H, 2 * N - 1, generator=g) * 0.5).to(dev)
grid = int(round((N - 1) ** 0.5))
coef = (torch.randn(4, generator=g) * 0.5 * torch.tensor([float(2 * grid) ** -k for k in range(4)])).to(dev)


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


planes = ops.QkvRopeFn.apply(x, w, cos, sin, H)
fl_fwd, by_fwd = 4.0 * N * N * E * B, (8.0 * N * E + 4.0 * H * N) * B
t = timeit(lambda: ops.QkvRopeFn.apply(x, w, cos, sin, H))
print(f"{geo} qkv_rope_fwd      {t * 1e3:8.1f} us  {2.0 * B * N * E * 3 * E / t / 1e9:7.1f} TFLOP/s")
for name, mode, prm, gr in (("none", _lib.BIAS_NONE, None, 0), ("table", _lib.BIAS_TABLE, table, 0),
                            ("poly", _lib.BIAS_POLY, coef, grid)):
    t = timeit(lambda: ops.fused_attention(planes, D ** -0.5, mode, prm, gr))
    print(f"{geo} attn_fwd[{name:5s}]   {t * 1e3:8.1f} us  {fl_fwd / t / 1e9:7.1f} TFLOP/s  {by_fwd / t / 1e6:7.1f} GB/s (algorithmic)")
    pl = planes.detach().requires_grad_(True)
    pr = None if prm is None else prm.detach().requires_grad_(True)
    o = ops.fused_attention(pl, D ** -0.5, mode, pr, gr)
    t = timeit(lambda: torch.autograd.grad(o, [pl] + ([pr] if pr is not None else []), d_out, retain_graph=True))
    print(f"{geo} attn_bwd[{name:5s}]   {t * 1e3:8.1f} us  {2 * fl_fwd / t / 1e9:7.1f} TFLOP/s  {2 * by_fwd / t / 1e6:7.1f} GB/s (algorithmic)")
dpl = torch.randn_like(planes)
xs, ws = x.detach().requires_grad_(True), w.detach().requires_grad_(True)
cs, sn = cos.detach().requires_grad_(True), sin.detach().requires_grad_(True)
pp = ops.QkvRopeFn.apply(xs, ws, cs, sn, H)
ops.PROFILE_EVENTS = None
t = timeit(lambda: torch.autograd.grad(pp, [xs, ws, cs, sn], dpl, retain_graph=True), iters=10)
print(f"{geo} qkv_rope_bwd(all) {t * 1e3:8.1f} us  (rope un-rotation + d_cos/d_sin + cuBLAS dX, dW)")
