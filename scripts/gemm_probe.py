"""Check and time the tcgen05 CTA-pair GEMM (vrr_gemm_ex, bf16) on a B200 (run under `timeout`).

    python scripts/gemm_probe.py check     # every layout / epilogue / ragged shape vs an fp32 torch product
    python scripts/gemm_probe.py time      # ViT-B / ViT-L shapes: TFLOP/s, with torch.matmul (cuBLAS) beside it
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_rpe_rope_b200 import _lib  # noqa: E402

dev = "cuda:0"
lib = _lib.load()
which = sys.argv[1] if len(sys.argv) > 1 else "check"
torch.backends.cuda.matmul.allow_tf32 = False


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def gemm(a, b, ta, tb, c_dtype, bias=None, epi=0, accumulate_into=None, aux=None):
    M = a.shape[1] if ta else a.shape[0]
    K = a.shape[0] if ta else a.shape[1]
    N = b.shape[0] if tb else b.shape[1]
    c = accumulate_into if accumulate_into is not None else torch.empty(M, N, device=dev, dtype=c_dtype)
    c2 = torch.empty_like(c) if epi in (2, 3) else (aux if epi == 4 else None)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.vrr_gemm_ex(ptr(a), ptr(b), ptr(c), ptr(c2), ptr(bias), M, N, K, ta, tb, 1, 0 if c_dtype == torch.float32 else 1,
                         epi, 1 if accumulate_into is not None else 0, st)
    assert rc == 0, _lib.last_error()
    return c, c2


def rel(x, y):
    return ((x.float() - y.float()).abs().max() / y.float().abs().max().clamp_min(1e-6)).item()


if which == "check":
    _lib.set_impl(_lib.IMPL_TCGEN05)
    bad = 0
    g = torch.Generator().manual_seed(0)
    shapes = [(256, 256, 64), (256, 256, 128), (512, 768, 768), (394, 2304, 768), (1000, 104, 200), (8, 8, 8), (7, 128, 32),
              (130, 72, 520), (3000, 1000, 768), (768, 768, 3152), (2304, 768, 1576)]
    for (M, N, K) in shapes:
        for ta, tb in ((0, 1), (0, 0), (1, 0), (1, 1)):
            if ta and M % 8:
                continue
            A = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16)
            Bm = (torch.randn(K, N, generator=g) * 0.5).to(torch.bfloat16)
            a = (A.t().contiguous() if ta else A).to(dev)
            b = (Bm.t().contiguous() if tb else Bm).to(dev)
            want = A.float().to(dev) @ Bm.float().to(dev)
            for cdt in (torch.bfloat16, torch.float32):
                c, _ = gemm(a, b, ta, tb, cdt)
                torch.cuda.synchronize()
                e = rel(c, want)
                ok = e <= (8e-3 if cdt == torch.bfloat16 else 2e-5)
                bad += not ok
                print(f"gemm M={M} N={N} K={K} ta={ta} tb={tb} c={str(cdt)[6:]}: rel err {e:.2e} {'ok' if ok else 'BAD'}", flush=True)
            if (ta, tb) == (0, 1):
                bias = torch.randn(N, generator=g).to(dev)
                bias_r = bias.to(torch.bfloat16).float()
                c, _ = gemm(a, b, ta, tb, torch.bfloat16, bias, 1)
                e1 = rel(c, want + bias_r)
                h, a2 = gemm(a, b, ta, tb, torch.bfloat16, bias, 2)
                torch.cuda.synchronize()
                e2 = rel(h, want + bias_r)
                e3 = rel(a2, torch.nn.functional.gelu(h.float()))
                cf, _ = gemm(a, b, ta, tb, torch.float32, bias, 1)
                e4 = rel(cf, want + bias)
                a3, gp = gemm(a, b, ta, tb, torch.bfloat16, bias, 3)
                torch.cuda.synchronize()
                hf = h.float().requires_grad_(True)
                torch.nn.functional.gelu(hf).sum().backward()
                e5, e6 = rel(a3, torch.nn.functional.gelu(h.float())), rel(gp, hf.grad)
                ok = e1 <= 8e-3 and e2 <= 8e-3 and e3 <= 8e-3 and e4 <= 2e-5 and e5 <= 8e-3 and e6 <= 8e-3
                bad += not ok
                print(f"     bias {e1:.2e}  bias+gelu h {e2:.2e} gelu {e3:.2e}  fp32+bias {e4:.2e}  gelu/grad {e5:.2e} {e6:.2e} "
                      f"{'ok' if ok else 'BAD'}", flush=True)
            if (ta, tb) == (0, 0) and N % 8 == 0:
                mul = torch.randn(M, N, generator=g).to(torch.bfloat16).to(dev)
                c, _ = gemm(a, b, ta, tb, torch.bfloat16, None, 4, aux=mul)
                torch.cuda.synchronize()
                e = rel(c, want * mul.float())
                ok = e <= 8e-3
                bad += not ok
                print(f"     mul epilogue: {e:.2e} {'ok' if ok else 'BAD'}", flush=True)
            if (ta, tb) == (1, 0):
                acc = torch.randn(M, N, generator=g).to(dev)
                base = acc.clone()
                gemm(a, b, ta, tb, torch.float32, accumulate_into=acc)
                torch.cuda.synchronize()
                e = rel(acc, base + want)
                ok = e <= 2e-5
                bad += not ok
                print(f"     accumulate: {e:.2e} {'ok' if ok else 'BAD'}", flush=True)
    print("PROBE gemm", "FAILED" if bad else "PASSED")
    sys.exit(1 if bad else 0)


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


if which == "time":
    g = torch.Generator().manual_seed(0)
    rows = {"vitb": 256 * 197, "vitl": 64 * 577}
    for geo, E in (("vitb", 768), ("vitl", 1024)):
        Mtok = rows[geo]
        cases = [("qkv fwd  NT", Mtok, 3 * E, E, 0, 1, torch.bfloat16, 0), ("proj fwd NT", Mtok, E, E, 0, 1, torch.bfloat16, 1),
                 ("fc1 fwd  NT+gelu", Mtok, 4 * E, E, 0, 1, torch.bfloat16, 2), ("fc1 fwd  NT+gelu+grad", Mtok, 4 * E, E, 0, 1, torch.bfloat16, 3),
                 ("fc2 dX   NN*mul", Mtok, 4 * E, E, 0, 0, torch.bfloat16, 4), ("fc2 fwd  NT", Mtok, E, 4 * E, 0, 1, torch.bfloat16, 1),
                 ("qkv dX   NN", Mtok, E, 3 * E, 0, 0, torch.bfloat16, 0), ("fc1 dX   NN", Mtok, E, 4 * E, 0, 0, torch.bfloat16, 0),
                 ("fc2 dX   NN", Mtok, 4 * E, E, 0, 0, torch.bfloat16, 0),
                 ("qkv dW   TN", 3 * E, E, Mtok, 1, 0, torch.float32, 0), ("proj dW  TN", E, E, Mtok, 1, 0, torch.float32, 0),
                 ("fc1 dW   TN", 4 * E, E, Mtok, 1, 0, torch.float32, 0), ("fc2 dW   TN", E, 4 * E, Mtok, 1, 0, torch.float32, 0)]
        for name, M, N, K, ta, tb, cdt, epi in cases:
            a = (torch.randn((K, M) if ta else (M, K), generator=g) * 0.5).to(torch.bfloat16).to(dev)
            b = (torch.randn((N, K) if tb else (K, N), generator=g) * 0.5).to(torch.bfloat16).to(dev)
            bias = torch.randn(N, generator=g).to(dev) if epi in (1, 2, 3) else None
            aux = torch.randn(M, N, generator=g).to(torch.bfloat16).to(dev) if epi == 4 else None
            t_mine = timeit(lambda: gemm(a, b, ta, tb, cdt, bias, epi, aux=aux))
            aa, bb = (a.t() if ta else a), (b.t() if tb else b)
            t_lib = timeit(lambda: torch.matmul(aa, bb))
            fl = 2.0 * M * N * K
            print(f"{geo} {name:18s} M={M:6d} N={N:5d} K={K:6d}: mine {t_mine * 1e3:8.1f} us {fl / t_mine / 1e9:7.1f} TF | "
                  f"cuBLAS (no epilogue, bf16 out) {t_lib * 1e3:8.1f} us {fl / t_lib / 1e9:7.1f} TF", flush=True)
