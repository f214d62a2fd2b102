"""Cross-check the tcgen05 kernels against the SIMT kernels on a B200 (run under `timeout`).

Prints one line per case with the max error relative to max|reference|; exits non-zero if any case
is off.  Used by scripts/gpu_check.sh before the test-suite so that a broken tensor-core kernel is
reported (and the run continues on the SIMT family) instead of hanging the suite."""
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_rpe_rope_b200 import _lib, ops  # noqa: E402

dev = "cuda:0"
which = sys.argv[1] if len(sys.argv) > 1 else "attn"
bad = 0


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-3)).item()


if which == "attn":
    cases = [(2, 3, n, kind) for n in (197, 65, 1, 16, 128, 129, 256, 257, 577, 1025) for kind in ("none", "table", "poly")]
    cases += [(64, 12, 197, "none"), (64, 12, 197, "table"), (64, 12, 197, "poly"), (8, 16, 577, "none")]
    for (b, h, n, kind) in cases:
        g = int(round((n - 1) ** 0.5))
        if kind == "poly" and (g * g != n - 1 or n == 1):
            continue
        gen = torch.Generator().manual_seed(n * 7 + b)
        planes = (torch.randn(3, b, h, n, 64, generator=gen) * 0.8).to(torch.bfloat16).to(dev)
        mode, prm, grid = _lib.BIAS_NONE, None, 0
        if kind == "table":
            mode, prm = _lib.BIAS_TABLE, (torch.randn(h, 2 * n - 1, generator=gen) * 0.5).to(dev)
        elif kind == "poly":
            mode, prm, grid = _lib.BIAS_POLY, (torch.randn(h, 4, generator=gen) * 0.02).to(dev), g
        outs = {}
        for name, impl in (("simt", _lib.IMPL_SIMT), ("tc", _lib.IMPL_TCGEN05)):
            _lib.set_impl(impl)
            o = ops.fused_attention(planes, 0.125, mode, prm, grid)
            torch.cuda.synchronize()
            outs[name] = o
        e = rel(outs["tc"], outs["simt"])
        ok = e < 2e-2 and bool(torch.isfinite(outs["tc"].float()).all())
        bad += (not ok)
        print(f"attn_fwd B={b} H={h} N={n} bias={kind}: rel err tc vs simt {e:.3e} {'ok' if ok else 'BAD'}", flush=True)
elif which == "qkv":
    for (b, n, e_, h, rope) in [(2, 197, 256, 4, "none"), (2, 197, 256, 4, "axial"), (2, 197, 256, 4, "mixed"),
                                (3, 65, 192, 3, "mixed"), (1, 5, 64, 1, "axial"), (16, 197, 768, 12, "mixed"),
                                (4, 577, 1024, 16, "axial"), (2, 1025, 768, 12, "mixed"), (7, 50, 128, 2, "none")]:
        gen = torch.Generator().manual_seed(n + e_)
        x = torch.randn(b, n, e_, generator=gen).to(torch.bfloat16).to(dev)
        w = (torch.randn(3 * e_, e_, generator=gen) * e_ ** -0.5).to(torch.bfloat16).to(dev)
        cos = sin = None
        if rope != "none":
            g = int(round((n - 1) ** 0.5))
            if g * g != n - 1:
                continue
            dh = e_ // h
            shape = (n - 1, dh // 2) if rope == "axial" else (h, n - 1, dh // 2)
            ang = torch.rand(*shape, generator=gen) * 6.0
            cos, sin = torch.cos(ang).to(dev), torch.sin(ang).to(dev)
        outs = {}
        for name, impl in (("simt", _lib.IMPL_SIMT), ("tc", _lib.IMPL_TCGEN05)):
            _lib.set_impl(impl)
            o = ops.QkvRopeFn.apply(x, w, cos, sin, h)
            torch.cuda.synchronize()
            outs[name] = o
        er = rel(outs["tc"], outs["simt"])
        ok = er < 2e-2 and bool(torch.isfinite(outs["tc"].float()).all())
        bad += (not ok)
        print(f"qkv_rope_fwd B={b} N={n} E={e_} H={h} rope={rope}: rel err tc vs simt {er:.3e} {'ok' if ok else 'BAD'}", flush=True)
elif which == "patch":
    for (b, c, hw, pch, e_, absolute, tdt) in [(4, 3, 224, 16, 768, False, torch.float32), (4, 3, 224, 16, 768, True, torch.float32),
                                              (3, 3, 64, 8, 256, True, torch.float32), (2, 3, 64, 16, 128, True, torch.bfloat16),
                                              (2, 1, 64, 8, 64, False, torch.float32), (5, 3, 384, 16, 1024, False, torch.float32)]:
        gen = torch.Generator().manual_seed(hw + e_)
        idt = torch.float32 if tdt == torch.float32 else torch.bfloat16
        img = torch.randn(b, c, hw, hw, generator=gen).to(idt).to(dev)
        w = (torch.randn(e_, c, pch, pch, generator=gen) * 0.05).to(torch.bfloat16).to(dev)
        bias = (torch.randn(e_, generator=gen) * 0.1).to(torch.bfloat16).to(dev)
        cls = (torch.randn(1, 1, e_, generator=gen) * 0.1).to(tdt).to(dev)
        pos = (torch.randn(1, 700, e_, generator=gen) * 0.1).to(tdt).to(dev) if absolute else None
        outs = {}
        for name, impl in (("simt", _lib.IMPL_SIMT), ("tc", _lib.IMPL_TCGEN05)):
            _lib.set_impl(impl)
            o = ops.PatchEmbedFn.apply(img, w, bias, cls, pos, pch)
            torch.cuda.synchronize()
            outs[name] = o
        er = rel(outs["tc"], outs["simt"])
        ok = er < 1e-2 and bool(torch.isfinite(outs["tc"].float()).all())
        bad += (not ok)
        print(f"patch_embed B={b} C={c} img={hw} P={pch} E={e_} abs={absolute} tok={tdt}: rel err tc vs simt {er:.3e} "
              f"{'ok' if ok else 'BAD'}", flush=True)
elif which == "bwd":
    cases = [(2, 3, n, kind) for n in (197, 65, 1, 17, 64, 128, 129, 257, 577, 1025) for kind in ("none", "table", "poly", "polyh")]
    cases += [(32, 12, 197, "none"), (32, 12, 197, "table"), (32, 12, 197, "poly")]
    for (b, h, n, kind) in cases:
        g = int(round((n - 1) ** 0.5))
        if kind.startswith("poly") and (g * g != n - 1 or n == 1):
            continue
        gen = torch.Generator().manual_seed(n * 3 + b)
        planes = (torch.randn(3, b, h, n, 64, generator=gen) * 0.8).to(torch.bfloat16).to(dev)
        d_out = torch.randn(b, n, h * 64, generator=gen).to(torch.bfloat16).to(dev)
        mode, prm, grid = _lib.BIAS_NONE, None, 0
        if kind == "table":
            mode, prm = _lib.BIAS_TABLE, (torch.randn(h, 2 * n - 1, generator=gen) * 0.5).to(dev)
        elif kind.startswith("poly"):
            span = torch.tensor([float(max(2 * g - 2, 1)) ** -k for k in range(4)])
            c = (torch.randn(h, 4, generator=gen) if kind == "polyh" else torch.randn(4, generator=gen)) * 0.7 * span
            mode, prm, grid = _lib.BIAS_POLY, c.to(dev), g
        res = {}
        for name, impl in (("simt", _lib.IMPL_SIMT), ("tc", _lib.IMPL_TCGEN05)):
            _lib.set_impl(_lib.IMPL_SIMT)   # identical forward (same saved out / lse) for both backward families
            pl = planes.clone().requires_grad_(True)
            pr = None if prm is None else prm.clone().requires_grad_(True)
            o = ops.fused_attention(pl, 0.125, mode, pr, grid)
            _lib.set_impl(impl)
            o.backward(d_out)
            torch.cuda.synchronize()
            res[name] = (pl.grad, None if pr is None else pr.grad)
        errs = [rel(res["tc"][0][k], res["simt"][0][k]) for k in range(3)]
        eb = rel(res["tc"][1], res["simt"][1]) if prm is not None else 0.0
        ok = max(errs) < 2e-2 and eb < 2e-2 and bool(torch.isfinite(res["tc"][0].float()).all())
        bad += (not ok)
        print(f"attn_bwd B={b} H={h} N={n} bias={kind}: dq {errs[0]:.2e} dk {errs[1]:.2e} dv {errs[2]:.2e} dbias {eb:.2e} "
              f"{'ok' if ok else 'BAD'}", flush=True)
_lib.set_impl(_lib.IMPL_AUTO)
print("PROBE", which, "FAILED" if bad else "PASSED", flush=True)
sys.exit(1 if bad else 0)
