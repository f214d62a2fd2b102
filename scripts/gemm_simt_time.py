"""fp32 FFMA GEMM (vrr_gemm) against cuBLAS SGEMM at the ViT-Tiny shapes, both tile sizes."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_rpe_rope_b200 import _lib
lib = _lib.load()
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda:0"
p = lambda t: ctypes.c_void_p(t.data_ptr())
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def timeit(fn, iters=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
Mt = 128 * 65
cases = [("fc1 fwd NT", Mt, 768, 192, 0, 1), ("fc2 fwd NT", Mt, 192, 768, 0, 1), ("qkv fwd NT", Mt, 576, 192, 0, 1),
         ("fc1 dX NN", Mt, 192, 768, 0, 0), ("fc2 dX NN", Mt, 768, 192, 0, 0), ("fc1 dW TN", 768, 192, Mt, 1, 0), ("fc2 dW TN", 192, 768, Mt, 1, 0)]
for name, M, N, K, ta, tb in cases:
    a = torch.randn((K, M) if ta else (M, K), device=dev)
    b = torch.randn((N, K) if tb else (K, N), device=dev)
    c = torch.empty(M, N, device=dev)
    res = {}
    for tile in (64, 128):
        _lib.check(lib.vrr_set_option(b"simt_gemm_tile", tile), "opt")
        res[tile] = timeit(lambda: _lib.check(lib.vrr_gemm(p(a), p(b), p(c), M, N, K, ta, tb, 0, 0, st()), "gemm"))
    aa, bb = (a.t() if ta else a), (b.t() if tb else b)
    want = aa @ bb
    err = ((c - want).abs().max() / want.abs().max()).item()
    t_lib = timeit(lambda: torch.matmul(aa, bb))
    fl = 2.0 * M * N * K
    print(f"{name:12s} M={M:5d} N={N:4d} K={K:5d}: 64-tile {res[64]*1e3:7.1f} us {fl/res[64]/1e9:5.1f} TF | 128-tile {res[128]*1e3:7.1f} us {fl/res[128]/1e9:5.1f} TF | cuBLAS {t_lib*1e3:7.1f} us {fl/t_lib/1e9:5.1f} TF | err {err:.1e}")
