import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_rpe_rope_b200 import _lib, ops
lib = _lib.load()
dev = "cuda:0"
b, h, n = 64, 12, 197
gen = torch.Generator().manual_seed(n * 7 + b)
planes = (torch.randn(3, b, h, n, 64, generator=gen) * 0.8).to(torch.bfloat16).to(dev)
prm = (torch.randn(h, 2 * n - 1, generator=gen) * 0.5).to(dev)
zero = torch.zeros_like(prm)

def run(impl, mode, p):
    _lib.set_impl(impl)
    o = ops.fused_attention(planes, 0.125, mode, p, 0).float().view(b, n, h, 64)
    torch.cuda.synchronize()
    return o

ref_none = run(_lib.IMPL_SIMT, _lib.BIAS_NONE, None)
ref_tab = run(_lib.IMPL_SIMT, _lib.BIAS_TABLE, prm)
for pad in (0, 60):
    lib.vrr_set_option(b"attn_fwd_smem_pad_kb", pad)
    for rep in range(3):
        e0 = (run(_lib.IMPL_TCGEN05, _lib.BIAS_TABLE, zero) - ref_none).abs()
        e1 = (run(_lib.IMPL_TCGEN05, _lib.BIAS_TABLE, prm) - ref_tab).abs()
        e2 = (run(_lib.IMPL_TCGEN05, _lib.BIAS_NONE, None) - ref_none).abs()
        bad1 = (e1.amax(dim=3) > 0.05).nonzero()
        print(f"pad={pad}KB rep={rep}: zero-table err {e0.max():.4f}  table err {e1.max():.4f}  none err {e2.max():.4f}  "
              f"bad rows(table)={bad1.shape[0]} rows>= {bad1[:,1].min().item() if bad1.shape[0] else '-'}")
