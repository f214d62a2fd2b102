"""Fused add+LayerNorm kernels: register-resident (ln_reg=1) vs generic multi-pass (ln_reg=0) - bit equality and time.
usage: python scripts/ln_time.py [M] [E]      (default ViT-B step: 256 x 197 rows of 768)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_rpe_rope_b200 import _lib
from vit_rpe_rope_b200.ops import _ptr, _DT, _stream

M = int(sys.argv[1]) if len(sys.argv) > 1 else 256 * 197
E = int(sys.argv[2]) if len(sys.argv) > 2 else 768
lib = _lib.load()
dev = torch.device("cuda")
torch.manual_seed(0)


def opt(name, v):
    _lib.check(lib.vrr_set_option(name.encode(), int(v)), name)


def run(dt, reps=20):
    x = torch.randn(M, E, device=dev)
    br = torch.randn(M, E, device=dev).to(dt)
    g = torch.randn(E, device=dev)
    b = torch.randn(E, device=dev)
    dy = torch.randn(M, E, device=dev).to(dt)
    dxn = torch.randn(M, E, device=dev)
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    outs = {}
    for mode in ((0, 2), (1, 2), (1, 1), (1, 3)):
        opt("ln_reg", mode[0]); opt("ln_bwd_minb", mode[1])
        xn = torch.empty_like(x); y = torch.empty(M, E, device=dev, dtype=dt)
        mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
        dx = torch.empty_like(x); dbr = torch.empty(M, E, device=dev, dtype=dt)
        dg = torch.empty(E, device=dev); db = torch.empty(E, device=dev)
        fwd = lambda: _lib.check(lib.vrr_add_layernorm_fwd(_ptr(x), _ptr(br), _ptr(xn), _ptr(g), _ptr(b), _ptr(y), _ptr(mean),
                                                           _ptr(rstd), M, E, 1e-6, _DT[dt], _DT[dt], _stream()), "fwd")
        bwd = lambda: _lib.check(lib.vrr_add_layernorm_bwd(_ptr(dy), _ptr(dxn), _ptr(xn), _ptr(g), _ptr(mean), _ptr(rstd), _ptr(dx),
                                                           _ptr(dbr), _ptr(dg), _ptr(db), M, E, _DT[dt], _DT[dt], _stream()), "bwd")
        t = {}
        for name, fn in (("fwd", fwd), ("bwd", bwd)):
            for _ in range(3):
                fn()
            tot = 0.0
            for _ in range(reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            t[name] = tot / reps * 1e3
        outs[mode] = [o.clone() for o in (xn, y, mean, rstd, dx, dbr, dg, db)]
        es = 2 if dt == torch.bfloat16 else 4
        fb = M * E * (4 + es + 4 + es); bb = M * E * (es + 4 + 4 + 4 + es)
        print(f"{str(dt):15s} ln_reg={mode[0]} minb={mode[1]}: fwd {t['fwd']:7.1f} us ({fb / t['fwd'] / 1e3:6.0f} GB/s)   "
              f"bwd {t['bwd']:7.1f} us ({bb / t['bwd'] / 1e3:6.0f} GB/s)")
    names = "x_new y mean rstd dx d_branch dgamma dbeta".split()
    ref = outs[(0, 2)]
    for mode in ((1, 2), (1, 1), (1, 3)):
        for n, a, r in zip(names, outs[mode], ref):
            if n in ("dgamma", "dbeta"):  # atomics over a different row partition: compare with a tolerance
                err = float((a - r).abs().max() / r.abs().max())
                assert err < 1e-5, (mode, n, err)
            else:
                assert torch.equal(a, r), (mode, n, float((a.float() - r.float()).abs().max()))
    print(f"{dt}: register-resident == generic (bitwise; dgamma/dbeta to 1e-5)")


for dt in (torch.bfloat16, torch.float32):
    run(dt)
opt("ln_reg", 1); opt("ln_bwd_minb", 3)
