"""Run the hot-path kernels at ViT-B/16-224 size a few times (for ncu captures; not a benchmark)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_rpe_rope_b200 import _lib, ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
B, H, N, D = 256, 12, 197, 64
E = H * D
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
x = torch.randn(B, N, E, generator=g).to(torch.bfloat16).to(dev).requires_grad_(True)
w = (torch.randn(3 * E, E, generator=g) * E ** -0.5).to(torch.bfloat16).to(dev).requires_grad_(True)
ang = torch.rand(H, N - 1, D // 2, generator=g) * 6
cos, sin = torch.cos(ang).to(dev), torch.sin(ang).to(dev)
d_out = torch.randn(B, N, E, generator=g).to(torch.bfloat16).to(dev)
table = (torch.randn(H, 2 * N - 1, generator=g) * 0.5).to(dev)
for it in range(3):
    planes = ops.QkvRopeFn.apply(x, w, cos, sin, H)
    if which in ("all", "rope"):
        o = ops.fused_attention(planes, D ** -0.5)
        o.backward(d_out)
    if which in ("all", "table"):
        pl = planes.detach().requires_grad_(True)
        o = ops.fused_attention(pl, D ** -0.5, _lib.BIAS_TABLE, table.requires_grad_(True), 0)
        o.backward(d_out)
torch.cuda.synchronize()
print("done", _lib.launch_count(), "launches")
