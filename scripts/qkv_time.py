"""QKV projection + RoPE epilogue in isolation, per rope mode, beside the plain GEMM of the same shape.
usage: python scripts/qkv_time.py [B N E H]"""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_rpe_rope_b200 import _lib
from vit_rpe_rope_b200.ops import _ptr, _DT, _stream
a = [int(v) for v in sys.argv[1:]]
B, N, E, H = (a + [256, 197, 768, 12][len(a):])[:4]
D = E // H
lib = _lib.load()
dev = "cuda:0"
dt = torch.bfloat16
x = torch.randn(B * N, E).to(dt).to(dev)
w = (torch.randn(3 * E, E) * E ** -0.5).to(dt).to(dev)
planes = torch.empty(3, B, H, N, D, device=dev, dtype=dt)
c = torch.empty(B * N, 3 * E, device=dev, dtype=dt)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
fl = 2.0 * B * N * 3 * E * E


def timeit(fn):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / 10 * 1e3


us = timeit(lambda: _lib.check(lib.vrr_gemm_ex(_ptr(x), _ptr(w), _ptr(c), None, None, B * N, 3 * E, E, 0, 1, 1, 1, 0, 0, _stream()), "gemm"))
print(f"plain GEMM [{B * N} x {3 * E} x {E}]: {us:7.1f} us {fl / us / 1e6:6.0f} TF")
for mode, name in ((0, "none"), (1, "axial"), (2, "mixed")):
    heads = H if mode == 2 else 1
    cos = torch.rand(heads, N - 1, D // 2, device=dev); sin = torch.rand(heads, N - 1, D // 2, device=dev)
    us = timeit(lambda: _lib.check(lib.vrr_qkv_rope_fwd(_ptr(x), _ptr(w), _ptr(cos), _ptr(sin), _ptr(planes), B, N, E, H, mode,
                                                        _DT[dt], _stream()), "qkv"))
    print(f"qkv_rope_fwd rope={name:5s}: {us:7.1f} us {fl / us / 1e6:6.0f} TF")
    if mode:
        packed = torch.empty(heads, D // 4, N - 1, 4, device=dev)
        _lib.check(lib.vrr_rope_pack_tables(_ptr(cos), _ptr(sin), _ptr(packed), heads, N - 1, D // 2, _stream()), "pack")
        us = timeit(lambda: _lib.check(lib.vrr_qkv_rope_fwd_packed(_ptr(x), _ptr(w), _ptr(cos), _ptr(sin), _ptr(packed), _ptr(planes),
                                                                   B, N, E, H, mode, _DT[dt], _stream()), "qkvp"))
        print(f"   ... with packed tables : {us:7.1f} us {fl / us / 1e6:6.0f} TF")
