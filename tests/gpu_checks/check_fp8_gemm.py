"""Standalone timing of the GEMM kernels with bf16 vs e4m3 operands (same shapes), plain bf16 output."""
import torch
from f5_tts_mlx_b200 import ops
dev = "cuda"
g = torch.Generator().manual_seed(0)
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for (M, N, K) in [(1874, 3072, 1024), (1874, 2048, 1024), (119936, 2048, 1024), (119936, 3072, 1024), (119936, 1024, 2048)]:
    a = torch.randn(M, K, generator=g).to(dev); w = (torch.randn(N, K, generator=g) * K ** -0.5).to(dev)
    a16, w16 = a.bfloat16(), w.bfloat16()
    a8, w8 = a.clamp(-448, 448).to(torch.float8_e4m3fn), (w * 64).clamp(-448, 448).to(torch.float8_e4m3fn)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for var, tile in ((0, 0), (2, 256), (1, 128)):
        t16 = timeit(lambda: ops.gemm(a16, w16, out, variant=var, tile_n=tile))
        t8 = timeit(lambda: ops.gemm(a8, w8, out, ab_fp8=True, acc_scale=1 / 64, variant=var, tile_n=tile))
        print(f"M{M} N{N} K{K} variant {var} tile {tile}: bf16 {t16 * 1e3:8.1f} us {2 * M * N * K / t16 / 1e9:6.0f} TF | e4m3 {t8 * 1e3:8.1f} us {2 * M * N * K / t8 / 1e9:6.0f} TF", flush=True)
