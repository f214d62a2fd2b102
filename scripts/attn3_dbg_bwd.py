"""Phase timeline (clock64 stamps) of CTA 0 of the whole-sequence attention backward."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_rpe_rope_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
B, H, N, D = 256, 12, 197, 64
g = torch.Generator().manual_seed(0)
planes = (torch.randn(3, B, H, N, D, generator=g) * 0.8).to(torch.bfloat16).to("cuda:0").requires_grad_(True)
d_out = torch.randn(B, N, H * D, generator=g).to(torch.bfloat16).to("cuda:0")
o = ops.fused_attention(planes, D ** -0.5)
for _ in range(3):
    torch.autograd.grad(o, [planes], d_out, retain_graph=True)
torch.cuda.synchronize()
buf = torch.zeros(1024, dtype=torch.int64, device="cuda:0")
_lib.check(lib.vrr_debug_timestamps(buf.data_ptr()), "dbg")
torch.autograd.grad(o, [planes], d_out, retain_graph=True)
torch.cuda.synchronize()
_lib.check(lib.vrr_debug_timestamps(None), "dbg")
v = buf.cpu().view(4, 256)
t0 = int(v[v > 0].min())
names = ["group 0 per pass (wait_s, s_ready, p_arrived, acc_ready)", "group 1 per pass",
         "issuer per pass (start, S/dP issued, P[0] ready, acc[0] issued, P[1] ready, acc[1] issued)", "group 0 epilogue per lane tile (acc wait done, ld done, arrived, first store done)"]
per = [4, 4, 8, 4]
for reg in range(4):
    print(names[reg])
    row = v[reg]
    for i in range(0, 256, per[reg]):
        chunk = row[i:i + per[reg]]
        if int(chunk.max()) == 0:
            break
        print(f"  {i // per[reg]:3d}: " + " ".join(f"{(int(c) - t0) if int(c) > 0 else -1:7d}" for c in chunk))
