"""qkv_rope_bwd (plane gradients -> token-layout d_qkv + cos/sin table gradients) in isolation.
usage: python scripts/rope_bwd_time.py [B] [N] [E] [H]   (default ViT-B step 256 x 197 x 768, 12 heads)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_rpe_rope_b200 import _lib
from vit_rpe_rope_b200.ops import _ptr, _DT, _stream

a = [int(v) for v in sys.argv[1:]]
B, N, E, H = (a + [256, 197, 768, 12][len(a):])[:4]
Dh, hd = E // H, E // H // 2
lib = _lib.load()
dev = torch.device("cuda")
torch.manual_seed(0)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
for dt in (torch.bfloat16, torch.float32):
    d_planes = torch.randn(3, B, H, N, Dh, device=dev).to(dt)
    planes = torch.randn(3, B, H, N, Dh, device=dev).to(dt)
    d_qkv = torch.empty(B * N, 3 * E, device=dev, dtype=dt)
    es = d_planes.element_size()
    for mode, name in ((0, "none"), (1, "axial"), (2, "mixed")):
        heads = H if mode == 2 else 1
        cos = torch.rand(heads, N - 1, hd, device=dev); sin = torch.rand(heads, N - 1, hd, device=dev)
        for cs in (False, True):
            if mode == 0 and cs:
                continue
            d_cos = torch.empty_like(cos) if cs else None
            d_sin = torch.empty_like(sin) if cs else None
            fn = lambda: _lib.check(lib.vrr_qkv_rope_bwd(_ptr(d_planes), _ptr(planes), _ptr(cos), _ptr(sin), _ptr(d_qkv), _ptr(d_cos),
                                                         _ptr(d_sin), B, N, E, H, mode, _DT[dt], _stream()), "rope_bwd")
            for _ in range(3):
                fn()
            tot = 0.0
            for _ in range(20):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            us = tot / 20 * 1e3
            nbytes = B * H * N * Dh * es * (3 + 3 + (2 if cs else 0))
            print(f"{str(dt):15s} rope={name:5s} table_grad={int(cs)}: {us:7.1f} us  {nbytes / us / 1e3:6.0f} GB/s of {nbytes / 1e6:.0f} MB")
