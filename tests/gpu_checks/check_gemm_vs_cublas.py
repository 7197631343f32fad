"""The block's GEMM shapes (bias epilogue, bf16 out, launcher defaults) against cuBLAS (torch.matmul / F.linear) on the
same box: µs per launch inside a replayed CUDA graph of 20 back-to-back launches (dependent launches, as in a step)."""
import torch
import torch.nn.functional as F
from f5_tts_mlx_b200 import ops
dev = "cuda"
g = torch.Generator().manual_seed(0)

def bench(fn, reps=10, per=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            for _ in range(per): fn()
    for _ in range(2): gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): gr.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * per)

for M in (1874, 119936):
    for name, N, K in (("QKV", 3072, 1024), ("out", 1024, 1024), ("FF1", 2048, 1024), ("FF2", 1024, 2048)):
        a = torch.randn(M, K, generator=g).bfloat16().to(dev); w = (torch.randn(N, K, generator=g) * K ** -0.5).bfloat16().to(dev)
        bias = torch.randn(N, generator=g).to(dev); bias16 = bias.bfloat16()
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        t_ours = bench(lambda: ops.gemm(a, w, out, bias=bias), reps=10 if M < 10000 else 3)
        t_cub = bench(lambda: F.linear(a, w, bias16), reps=10 if M < 10000 else 3)
        fl = 2.0 * M * N * K
        print(f"M {M:6d} {name} N{N} K{K}: ours {t_ours:8.2f} us ({fl / t_ours / 1e6:7.1f} TFLOP/s)   cuBLAS {t_cub:8.2f} us ({fl / t_cub / 1e6:7.1f} TFLOP/s)   ours/cuBLAS {t_ours / t_cub:.2f}", flush=True)
