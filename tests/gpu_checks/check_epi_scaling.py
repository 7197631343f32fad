"""Is the GEMM epilogue tail a per-SM chain or a chip-wide store burst?  Same out-proj-shaped kernel (128x128 tiles, fp32
out + residual + gate + fused-LN bf16 copy), M = 128 (8 CTAs) ... 1874 (120 CTAs): per-CTA mma_last -> epi_end."""
import numpy as np, torch
from f5_tts_mlx_b200 import ops
dev = "cuda"
N = K = 1024
g = torch.Generator().manual_seed(0)
for M in (128, 512, 1024, 1874):
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
        epi = (t[:, 8] - t[:, 6]) / 1e3; main = (t[:, 6] - t[:, 5]) / 1e3; pro = (t[:, 5] - t[:, 2]) / 1e3
        print(f"M {M:5d} ctas {t.shape[0]:4d} {mode:15s}: wait->mma0 {pro.mean():.2f}  main {main.mean():.2f}  "
              f"mma_last->epi_end mean {epi.mean():.2f} max {epi.max():.2f} us", flush=True)
