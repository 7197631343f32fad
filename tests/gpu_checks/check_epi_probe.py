"""Sub-step stamps of the one-wave GEMM's epilogue (library built with -DF5_EPI_PROBE=n, selected through F5_LIB):
prints slots 3 / 4 and epi_end relative to epi0 (accumulator complete), mean over the CTAs."""
import os, numpy as np, torch
from f5_tts_mlx_b200 import ops
dev = "cuda"
lvl = os.environ.get("F5_PROBE_LEVEL", "?")
WHAT = {"1": ("tmem ld 0 landed", "chunk 0 done"), "2": ("chunk 0 math done", "chunk 0 staged+fenced"),
        "3": ("chunk 0 wait_read done", "chunk 0 barrier passed"), "4": ("chunk 0 stores issued", "chunk 1 math done"),
        "5": ("chunk 1 wait_read done", "chunk 1 stores issued")}.get(lvl, ("slot3", "slot4"))
M, N, K = 1874, 1024, 1024
g = torch.Generator().manual_seed(0)
a = torch.randn(M, K, generator=g).bfloat16().to(dev); w = (torch.randn(N, K, generator=g) / 32).bfloat16().to(dev)
resid = torch.randn(M, N, generator=g).to(dev); gate = torch.randn(1, N, generator=g).to(dev)
bias = torch.randn(N, generator=g).to(dev); ln_scale = (torch.randn(N, generator=g) * 0.1).to(dev)
out = torch.empty(M, N, device=dev); out2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
stats = torch.empty(M, N // 64, 2, device=dev)
for mode in ("fp32 + LN copy", "fp32 only", "bf16 only"):
    kw = dict(bias=bias, resid=resid, gate=gate[0], tile_n=128, variant=1)
    o = out
    if mode == "fp32 + LN copy": kw.update(out2=out2, ln_scale=ln_scale, ln_stats=stats)
    if mode == "bf16 only": o = out2; kw.pop("resid"); kw.pop("gate")
    ts = torch.zeros(4096, 10, dtype=torch.int64, device=dev)
    for _ in range(3): ops.gemm(a, w, o, **kw)
    torch.cuda.synchronize()
    ops.gemm(a, w, o, debug_ts=ts, **kw); torch.cuda.synchronize()
    t = ts.cpu().double().numpy(); t = t[t[:, 0] > 0]
    rel = lambda s: float(np.mean(t[:, s] - t[:, 7])) / 1e3
    print(f"probe {lvl} {mode:15s}: mma_last->epi0 {float(np.mean(t[:, 7] - t[:, 6])) / 1e3:.2f} | {WHAT[0]} {rel(3):.2f}  {WHAT[1]} {rel(4):.2f}  "
          f"epi_end {rel(8):.2f}  exit {rel(9):.2f} us after epi0", flush=True)
