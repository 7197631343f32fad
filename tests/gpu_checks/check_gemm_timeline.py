"""Debug: per-CTA phase timelines (globaltimer) of the B=1 GEMM shapes, warm and after an L2 flush."""
import torch
from f5_tts_mlx_b200 import ops
dev = "cuda"
NAMES = ["start", "setup", "pdl_wait", "tma0", "tma_last", "mma0", "mma_last", "epi0", "epi_end", "exit"]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
def run(M, N, K, var, tile, cold):
    a = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(N, K, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ts = torch.zeros(4096, 10, dtype=torch.int64, device=dev)
    for _ in range(3): ops.gemm(a, w, out, tile_n=tile, variant=var)
    torch.cuda.synchronize()
    if cold: flush.zero_(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.gemm(a, w, out, tile_n=tile, variant=var, debug_ts=ts); e1.record(); torch.cuda.synchronize()
    t = ts.cpu()
    used = t[:, 0] > 0
    t = t[used].double()
    t0 = t[:, 0].min()
    rel = (t - t0) / 1e3
    rel[t == 0] = float("nan")
    import numpy as np
    r = rel.numpy()
    print(f"M{M} N{N} K{K} v{var}/t{tile} {'cold' if cold else 'warm'}: event {e0.elapsed_time(e1)*1e3:.1f} us, ctas {r.shape[0]}, span {np.nanmax(r):.1f} us")
    print("    " + " ".join(f"{n:>9s}" for n in NAMES))
    print("mean" + " ".join(f"{np.nanmean(r[:, i]):9.2f}" for i in range(10)))
    print("max " + " ".join(f"{np.nanmax(r[:, i]):9.2f}" for i in range(10)))
    print("min " + " ".join(f"{np.nanmin(r[:, i]):9.2f}" for i in range(10)), flush=True)
for cold in (False, True):
    for (M, N, K, var, tile) in [(1874, 1024, 1024, 1, 128), (1874, 3072, 1024, 1, 128), (1874, 3072, 1024, 2, 256), (1874, 1024, 2048, 1, 128),
                                 (1874, 1024, 2048, 2, 128), (1874, 2048, 1024, 1, 128), (1874, 2048, 1024, 2, 256)]:
        run(M, N, K, var, tile, cold)
