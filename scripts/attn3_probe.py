"""Probe of the persistent tcgen05 attention kernels (variant 3) on a B200: parity against the float64 numpy
oracle at ragged / multi-item sizes and every bias mode, then timing against variant 2.

    python scripts/attn3_probe.py [fwd] [bwd] [time]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import attention_np as A  # noqa: E402
from oracle import tables_np as T  # noqa: E402
from vit_rpe_rope_b200 import _lib, ops  # noqa: E402

DEV = "cuda:0"
lib = _lib.load()
what = sys.argv[1:] or ["fwd", "bwd", "time"]


def setopt(name, v):
    _lib.check(lib.vrr_set_option(name.encode(), int(v)), name)


def bias_case(kind, h, n):
    g = torch.Generator().manual_seed(5)
    if kind == "none":
        return _lib.BIAS_NONE, None, 0, None
    if kind == "table":
        tab = torch.randn(h, 2 * n - 1, generator=g) * 0.5
        return _lib.BIAS_TABLE, tab, 0, T.relative_bias(tab.numpy().astype(np.float64), n)
    shared = kind == "poly"
    grid = int(round((n - 1) ** 0.5))
    span = torch.tensor([float(max(2 * grid - 2, 1)) ** -k for k in range(4)])
    coef = (torch.randn(4, generator=g) if shared else torch.randn(h, 4, generator=g)) * 0.7 * span
    return _lib.BIAS_POLY, coef, grid, T.poly_bias(coef.numpy(), n - 1, h).astype(np.float64)


def check(b, h, n, kind, bwd):
    if kind.startswith("poly") and (n < 2 or int(round((n - 1) ** 0.5)) ** 2 != n - 1):
        return True
    d = 64
    mode, param, grid, bias_np = bias_case(kind, h, n)
    g = torch.Generator().manual_seed(n * 7 + b)
    planes = (torch.randn(3, b, h, n, d, generator=g) * 0.8).to(torch.bfloat16).to(DEV).requires_grad_(bwd)
    prm = None if param is None else param.to(DEV).requires_grad_(bwd)
    scale = d ** -0.5
    out = ops.fused_attention(planes, scale, mode, prm, grid)
    torch.cuda.synchronize()
    pl = planes.detach().double().cpu().numpy()
    want, _, _, _ = A.attention_forward(pl[0], pl[1], pl[2], scale, bias_np)
    got = out.detach().double().cpu().numpy()
    err = np.abs(got - want).max() / max(1e-6, np.abs(want).max())
    ok = bool(err < 2e-2) and bool(np.isfinite(got).all())
    msg = f"B={b} H={h} N={n} {kind:12s} fwd err {err:.2e}"
    if bwd:
        d_out = (torch.randn(b, n, h * d, generator=g) * 0.5).to(torch.bfloat16)
        grads = torch.autograd.grad(out, [planes] + ([prm] if prm is not None else []), d_out.to(DEV))
        torch.cuda.synchronize()
        gr = A.attention_backward(d_out.double().numpy(), pl[0], pl[1], pl[2], scale, bias_np, out=got)
        gp = grads[0].double().cpu().numpy()
        for idx, nm in enumerate(("dq", "dk", "dv")):
            ww = gr[nm]
            e = np.abs(gp[idx] - ww).max() / max(1e-6, np.abs(ww).max())
            msg += f" {nm} {e:.2e}"
            ok = ok and bool(e < 2e-2) and bool(np.isfinite(gp[idx]).all())
        if kind == "table":
            ww = A.dtable_from_dbias(gr["dbias"])
            e = np.abs(grads[1].double().cpu().numpy() - ww).max() / max(1e-6, np.abs(ww).max())
            msg += f" dtab {e:.2e}"
            ok = ok and bool(e < 2e-2)
        if kind.startswith("poly"):
            ww = A.dcoef_from_dbias(gr["dbias"], 3, shared=(kind == "poly"))
            e = np.abs(grads[1].double().cpu().numpy() - ww).max() / max(1e-6, np.abs(ww).max())
            msg += f" dcoef {e:.2e}"
            ok = ok and bool(e < 4e-2)
    print(("ok   " if ok else "FAIL ") + msg, flush=True)
    return ok


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


all_ok = True
if "fwd" in what or "bwd" in what:
    bwd = "bwd" in what
    sizes = [(2, 3, 1), (2, 3, 2), (2, 3, 17), (2, 3, 65), (2, 3, 197), (2, 3, 200), (2, 3, 257), (1, 2, 577), (1, 2, 1025),
             (40, 12, 197), (20, 12, 65), (3, 16, 577), (64, 12, 197)]
    for (b, h, n) in sizes:
        for kind in ("none", "table", "poly", "poly_perhead"):
            if b * h > 100 and kind == "poly_perhead":
                continue
            t0 = time.time()
            try:
                all_ok = check(b, h, n, kind, bwd) and all_ok
            except RuntimeError as e:
                print(f"FAIL B={b} H={h} N={n} {kind}: {e}", flush=True)
                all_ok = False
            if time.time() - t0 > 60:
                print("slow case - stopping", flush=True)
                sys.exit(2)
    print("PROBE attn3 " + ("PASSED" if all_ok else "FAILED"), flush=True)

if "time" in what:
    for geo, (B, H, N) in {"vitb": (256, 12, 197), "vitl": (64, 16, 577), "x512": (32, 12, 1025), "tiny64": (512, 4, 65)}.items():
        D = 64
        g = torch.Generator().manual_seed(0)
        planes = (torch.randn(3, B, H, N, D, generator=g) * 0.8).to(torch.bfloat16).to(DEV)
        d_out = torch.randn(B, N, H * D, generator=g).to(torch.bfloat16).to(DEV)
        fl, by = 4.0 * N * N * H * D * B, (8.0 * N * H * D + 4.0 * H * N) * B
        for variant in (2, 3):
            setopt("attn_fwd_variant", variant)
            setopt("attn_bwd_variant", variant)
            t = timeit(lambda: ops.fused_attention(planes, D ** -0.5))
            line = f"{geo} v{variant} fwd {t * 1e3:8.1f} us {fl / t / 1e9:7.1f} TF {by / t / 1e6:7.1f} GB/s"
            pl = planes.detach().requires_grad_(True)
            o = ops.fused_attention(pl, D ** -0.5)
            t = timeit(lambda: torch.autograd.grad(o, [pl], d_out, retain_graph=True))
            line += f" | bwd {t * 1e3:8.1f} us {2 * fl / t / 1e9:7.1f} TF {2 * by / t / 1e6:7.1f} GB/s"
            print(line, flush=True)
    setopt("attn_fwd_variant", 4)
    setopt("attn_bwd_variant", 3)
sys.exit(0 if all_ok else 1)
