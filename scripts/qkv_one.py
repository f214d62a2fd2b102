"""One QKV + RoPE projection through the library (for compute-sanitizer / ncu): python scripts/qkv_one.py B N E H"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_rpe_rope_b200 import ops
B, N, E, H = (int(v) for v in sys.argv[1:5])
D = E // H
g = torch.Generator().manual_seed(0)
x = torch.randn(B, N, E, generator=g).to(torch.bfloat16).cuda()
w = (torch.randn(3 * E, E, generator=g) * E ** -0.5).to(torch.bfloat16).cuda()
ang = torch.rand(H, N - 1, D // 2, generator=g) * 6
cos, sin = torch.cos(ang).cuda(), torch.sin(ang).cuda()
for _ in range(3):
    planes = ops.QkvRopeFn.apply(x, w, cos, sin, H)
torch.cuda.synchronize()
q = (x.float() @ w.float().t()).view(B, N, 3, H, D).permute(2, 0, 3, 1, 4)
print("ok", planes.float().abs().mean().item(), "v-plane err", (planes[2].float() - q[2]).abs().max().item())
