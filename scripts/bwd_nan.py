import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_rpe_rope_b200 import _lib, ops
B, H, N, D = (int(v) for v in sys.argv[1:5])
g = torch.Generator().manual_seed(0)
planes = (torch.randn(3, B, H, N, D, generator=g) * 0.8).to(torch.bfloat16).cuda().requires_grad_(True)
d_out = (torch.randn(B, N, H * D, generator=g) * 0.5).to(torch.bfloat16).cuda()
out = ops.fused_attention(planes, D ** -0.5)
(gr,) = torch.autograd.grad(out, [planes], d_out, retain_graph=True)
torch.cuda.synchronize()
lib = _lib.load()
_lib.check(lib.vrr_set_option(b"attn_bwd_variant", 2), "opt")
(ref,) = torch.autograd.grad(out, [planes], d_out)
torch.cuda.synchronize()
for w, nm in enumerate(("dq", "dk", "dv")):
    a, r = gr[w].float(), ref[w].float()
    bad = (~torch.isfinite(a)) | ((a - r).abs() > 0.05 * r.abs().max())
    per = bad.view(B * H, N, D).any(-1)          # [BH, N]
    items = per.any(-1).nonzero().flatten().tolist()
    print(nm, "bad items:", items[:40], "count", len(items))
    if items:
        it = items[0]
        rows = per[it].nonzero().flatten().tolist()
        print("   first bad item", it, "bad rows", rows[:50], "n", len(rows))
        r0 = rows[0]
        print("   got", a.view(B * H, N, D)[it, r0, :8].tolist())
        print("   ref", r.view(B * H, N, D)[it, r0, :8].tolist())
