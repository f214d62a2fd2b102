"""Phase timestamps (clock64) of one CTA of the attention forward kernel: where does a key tile's time go?"""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_rpe_rope_b200 import _lib, ops
lib = _lib.load()
B, H, N, D = (256, 12, 197, 64) if len(sys.argv) < 2 or sys.argv[1] == "vitb" else (64, 16, 577, 64)
kt = int(sys.argv[2]) if len(sys.argv) > 2 else 64
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 2
lib.vrr_set_option(b"attn_fwd_key_tile", kt)
lib.vrr_set_option(b"attn_fwd_variant", variant)
planes = (torch.randn(3, B, H, N, D) * 0.8).to(torch.bfloat16).cuda()
buf = torch.zeros(64, dtype=torch.int64, device="cuda")
for _ in range(3):
    ops.fused_attention(planes, D ** -0.5)
lib.vrr_debug_timestamps(ctypes.c_void_p(buf.data_ptr()))
ops.fused_attention(planes, D ** -0.5)
torch.cuda.synchronize()
lib.vrr_debug_timestamps(None)
t = buf.cpu().view(8, 8)
names = ["start", "S issued", "S ready", "softmax+P stored", "after syncthreads", "PV issued", "O ready", "acc updated"]
if variant == 2:
    names = ["tile start", "S ready", "S loaded", "max done", "exp+P st issued", "st done+arrived"]
    print(f"N={N} variant 2: softmax-thread phase deltas (cycles) for CTA (0, {B*H//2})")
    for i in range(8):
        if t[i, 0] == 0: break
        d = [int(t[i, k] - t[i, k - 1]) for k in range(1, 6)]
        nxt = int(t[i + 1, 0] - t[i, 5]) if i < 7 and t[i + 1, 0] != 0 else 0
        print(f" tile {i}: " + "  ".join(f"{n}:{v}" for n, v in zip(names[1:], d)) + f"  | total {int(t[i,5]-t[i,0])}")
    sys.exit(0)
print(f"N={N} KT={kt}: per-tile phase deltas (cycles) for CTA (0, {B*H//2})")
for i in range(8):
    if t[i, 0] == 0: break
    d = [int(t[i, k] - t[i, k - 1]) for k in range(1, 8)]
    print(f" tile {i}: " + "  ".join(f"{n}:{v}" for n, v in zip(names[1:], d)) + f"  | total {int(t[i,7]-t[i,0])}")
