"""Per-phase SM-clock timeline of attention CTA (0,0,0) (f5_debug_attention_ts)."""
import os, sys
import torch
from f5_tts_mlx_b200 import _lib

lib = _lib.load()
dev = "cuda"
B, N, H = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (2, 937, 16)))
D = H * 64
qkv = torch.randn(B * N, 3 * D, device=dev).bfloat16() * 0.3
out = torch.empty(B * N, D, device=dev, dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    lib.f5_attention_fwd(qkv.data_ptr(), 3 * D, out.data_ptr(), D, B, N, H, 64, None, st)
ts = torch.zeros(3, 64, 8, dtype=torch.int64, device=dev)
lib.f5_debug_attention_ts(ts.data_ptr())
lib.f5_attention_fwd(qkv.data_ptr(), 3 * D, out.data_ptr(), D, B, N, H, 64, None, st)
torch.cuda.synchronize()
lib.f5_debug_attention_ts(None)
t = ts.cpu()
t0 = int(t[t > 0].min())
nkv = (N + 127) // 128
names = ["top", "s_full", "ldtm", "max", "pv_wait", "handoff", "exp", "p_arrive"]
for g in (0, 1):
    print(f"group {g}: cycles since first stamp; per tile: " + " ".join(names))
    for j in range(nkv):
        row = [int(t[g, j, k]) - t0 if t[g, j, k] > 0 else -1 for k in range(8)]
        d = [row[0]] + [row[k] - row[k - 1] for k in range(1, 8)]
        print(f"  j={j}: t={row[0]:6d}  d: " + " ".join(f"{x:5d}" for x in d[1:]) + f"   iter={row[7] - row[0]}")
print("mma: S0 S1 vfull PV0 PV1 (absolute cycles)")
for j in range(nkv):
    print(f"  j={j}: " + " ".join(f"{int(t[2, j, k]) - t0 if t[2, j, k] > 0 else -1:6d}" for k in range(5)))
