"""Long-sequence attention forward (N > 256): parity of the four-CTAs-per-SM kernel (attn_fwd_streams = 4) against the
float64 numpy oracle at ragged sizes and every bias mode, then timing against the two-CTAs-per-SM kernel (= 2).

    python scripts/attn_long_probe.py [check] [time]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import attention_np as A  # noqa: E402
from oracle import tables_np as T  # noqa: E402
from vit_rpe_rope_b200 import _lib, ops  # noqa: E402

DEV = "cuda:0"
lib = _lib.load()
what = sys.argv[1:] or ["check", "time"]


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


def check(b, h, n, kind):
    if kind.startswith("poly") and int(round((n - 1) ** 0.5)) ** 2 != n - 1:
        return True
    d = 64
    mode, param, grid, bias_np = bias_case(kind, h, n)
    g = torch.Generator().manual_seed(n * 7 + b)
    planes = (torch.randn(3, b, h, n, d, generator=g) * 0.8).to(torch.bfloat16).to(DEV)
    prm = None if param is None else param.to(DEV)
    scale = d ** -0.5
    pl = planes.double().cpu().numpy()
    want, _, _, _ = A.attention_forward(pl[0], pl[1], pl[2], scale, bias_np)
    ok = True
    for streams in (4, 2):
        setopt("attn_fwd_streams", streams)
        out = ops.fused_attention(planes, scale, mode, prm, grid)
        torch.cuda.synchronize()
        got = out.double().cpu().numpy()
        err = np.abs(got - want).max() / max(1e-6, np.abs(want).max())
        good = bool(err < 2e-2) and bool(np.isfinite(got).all())
        ok &= good
        print(f"B={b} H={h} N={n} {kind:12s} streams={streams} err {err:.2e} {'ok' if good else 'FAIL'}", flush=True)
    return ok


def timeit(b, h, n, kind="none", reps=10):
    d = 64
    mode, param, grid, _ = bias_case(kind, h, n)
    planes = (torch.randn(3, b, h, n, d) * 0.8).to(torch.bfloat16).to(DEV)
    prm = None if param is None else param.to(DEV)
    flush = torch.empty(256 << 20, device=DEV, dtype=torch.uint8)
    res = {}
    for streams in (2, 4, 40, 41, 43, 44):  # 4x = four CTAs per SM with x of 8 exponential pairs by polynomial
        setopt("attn_fwd_streams", 2 if streams == 2 else 4)
        setopt("attn_fwd_poly_exp", streams - 40 if streams >= 40 else 2)
        for _ in range(3):
            ops.fused_attention(planes, d ** -0.5, mode, prm, grid)
        tot = 0.0
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.fused_attention(planes, d ** -0.5, mode, prm, grid)
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        res[streams] = tot / reps * 1e3
    fl = 4.0 * b * h * n * n * d
    print(f"B={b} H={h} N={n} {kind:6s}: 2-CTA {res[2]:7.1f} us ({fl / res[2] / 1e6:5.0f} TF)   4-CTA {res[4]:7.1f} us "
          f"({fl / res[4] / 1e6:5.0f} TF)   x{res[2] / res[4]:.2f}   poly 0/1/3/4 of 8: "
          + " ".join(f"{res[k]:.1f}" for k in (40, 41, 43, 44)), flush=True)
    setopt("attn_fwd_poly_exp", 2)


ok = True
if "check" in what:
    for (b, h, n) in [(1, 2, 257), (2, 3, 300), (1, 2, 577), (2, 2, 1025), (1, 1, 64 * 5 + 1), (1, 2, 130), (2, 2, 336), (1, 3, 368),
                      (2, 2, 384)]:
        for kind in ("none", "table", "poly", "poly_heads"):
            ok &= check(b, h, n, kind)
    print("PARITY", "OK" if ok else "FAILED")
if "time" in what:
    timeit(64, 16, 577)
    timeit(64, 12, 1025)
    timeit(32, 12, 1025)
    timeit(16, 12, 577, "table")
    timeit(16, 12, 577, "poly")
    timeit(256, 12, 197)  # N <= 256: the whole-sequence kernel serves every column above ...
    setopt("attn_fwd_variant", 2)  # ... unless the long-sequence kernels are forced: 2-CTA / 4-CTA at short sequences
    timeit(256, 12, 197)
    timeit(512, 4, 65)
    setopt("attn_fwd_variant", 4)
setopt("attn_fwd_streams", 4)
sys.exit(0 if ok else 1)
