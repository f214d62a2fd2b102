"""Time vrr_gemm_ex epilogues on one shape: python scripts/gemm_epi_time.py M N K  (A [M][K], B stored [N][K])"""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_rpe_rope_b200 import _lib
lib = _lib.load()
M, N, K = (int(v) for v in sys.argv[1:4])
dev = "cuda:0"
a = torch.randn(M, K).to(torch.bfloat16).to(dev)
b = (torch.randn(N, K) * K ** -0.5).to(torch.bfloat16).to(dev)
c = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
c2 = torch.randn(M, N).to(torch.bfloat16).to(dev)
bias = torch.randn(N, device=dev)
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
names = {0: "none", 1: "bias", 2: "bias+gelu (h, act)", 3: "bias+gelu (act, act')", 4: "multiply", 5: "bias+gelu (act only)"}
for epi in (0, 1, 2, 3, 4, 5):
    def run():
        rc = lib.vrr_gemm_ex(p(a), p(b), p(c), p(c2) if epi in (2, 3, 4) else None, p(bias) if epi not in (0, 4) else None, M, N, K, 0, 1, 1, 1, epi, 0, st)
        assert rc == 0, _lib.last_error()
    for _ in range(3):
        run()
    tot = 0.0
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    us = tot / 10 * 1e3
    print(f"M={M} N={N} K={K} epilogue {epi} {names[epi]:24s}: {us:7.1f} us  {2.0 * M * N * K / us / 1e6:6.0f} TFLOP/s")
