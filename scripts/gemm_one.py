"""Run one GEMM shape through vrr_gemm_ex a few times (for ncu): python scripts/gemm_one.py M N K ta tb [f32|epiN]"""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_rpe_rope_b200 import _lib
lib = _lib.load()
M, N, K, ta, tb = (int(v) for v in sys.argv[1:6])
f32 = len(sys.argv) > 6 and sys.argv[6] == "f32"
epi = int(sys.argv[6][3:]) if len(sys.argv) > 6 and sys.argv[6].startswith("epi") else 0
dev = "cuda:0"
a = torch.randn((K, M) if ta else (M, K)).to(torch.bfloat16).to(dev)
b = torch.randn((N, K) if tb else (K, N)).to(torch.bfloat16).to(dev)
c = torch.empty(M, N, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
c2 = torch.empty_like(c) if epi in (2, 3) else (torch.randn(M, N).to(torch.bfloat16).to(dev) if epi == 4 else None)
bias = torch.randn(N, device=dev) if epi in (1, 2, 3) else None
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
for _ in range(5):
    rc = lib.vrr_gemm_ex(p(a), p(b), p(c), p(c2), p(bias), M, N, K, ta, tb, 1, 0 if f32 else 1, epi, 0, st)
    assert rc == 0, _lib.last_error()
torch.cuda.synchronize()
print("ok", c.float().abs().mean().item())
