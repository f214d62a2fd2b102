"""Phase timeline (clock64 stamps) of CTA 0 of the persistent attention forward at a given geometry."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_rpe_rope_b200 import _lib, ops  # noqa: E402

lib = _lib.load()
geo = sys.argv[1] if len(sys.argv) > 1 else "vitb"
B, H, N = {"vitb": (256, 12, 197), "vitl": (64, 16, 577), "x512": (32, 12, 1025)}[geo]
D = 64
g = torch.Generator().manual_seed(0)
planes = (torch.randn(3, B, H, N, D, generator=g) * 0.8).to(torch.bfloat16).to("cuda:0")
for _ in range(3):
    ops.fused_attention(planes, D ** -0.5)
torch.cuda.synchronize()
buf = torch.zeros(1024, dtype=torch.int64, device="cuda:0")
_lib.check(lib.vrr_debug_timestamps(buf.data_ptr()), "dbg")
ops.fused_attention(planes, D ** -0.5)
torch.cuda.synchronize()
_lib.check(lib.vrr_debug_timestamps(None), "dbg")
v = buf.cpu().view(4, 256)
t0 = int(v[v > 0].min())
names = ["softmax WG0 (wait_s, s_ready, tile_done, arrived) per tile", "softmax WG1",
         "issuer per tile (start, S(f+1) issued, vfull, P0 ready, PV0 issued, P1 ready, PV1 issued)",
         "producer per tile (kempty passed, vempty passed)"]
per = [4, 4, 8, 2]
for reg in range(4):
    print(names[reg])
    row = v[reg]
    for i in range(0, 256, per[reg]):
        chunk = row[i:i + per[reg]]
        if int(chunk.max()) == 0:
            break
        print(f"  {i // per[reg]:3d}: " + " ".join(f"{(int(c) - t0) if int(c) > 0 else -1:7d}" for c in chunk))
