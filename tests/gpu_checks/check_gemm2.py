"""GPU bring-up of the persistent CTA-pair GEMM (variant=2) vs fp32 torch, plus timing vs variant 1."""
import sys
import torch
import torch.nn.functional as F
from f5_tts_mlx_b200 import ops
from f5_tts_mlx_b200.dit import rope_table

torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
g = torch.Generator().manual_seed(0)
fails = []
def rnd(*shape, scale=1.0): return (torch.randn(*shape, generator=g) * scale).to(dev)
def report(name, got, ref, tol):
    got = got.float(); err = (got - ref).abs().max().item(); rel = ((got - ref).norm() / (ref.norm() + 1e-30)).item()
    ok = rel < tol and err == err
    print(f"{'OK  ' if ok else 'FAIL'} {name:44s} max_abs={err:.3e} rel_l2={rel:.3e}", flush=True)
    if not ok: fails.append(name)

for (M, N, K, tile) in [(256, 256, 64, 256), (300, 256, 128, 256), (300, 256, 128, 128), (1874, 1024, 1024, 256), (1874, 1024, 1024, 128),
                        (130, 72, 200, 128), (5000, 3072, 1024, 256), (20000, 2048, 1024, 256), (20000, 100, 1024, 128)]:
    a = rnd(M, K).bfloat16(); w = rnd(N, K, scale=K ** -0.5).bfloat16(); bias = rnd(N)
    out = torch.full((M, N), float("nan"), device=dev)
    ops.gemm(a, w, out, bias=bias, tile_n=tile, variant=2)
    torch.cuda.synchronize()
    report(f"v2 plain f32 M{M} N{N} K{K} t{tile}", out, a.float() @ w.float().T + bias, 1e-5)

B, NF, D = 2, 937, 1024
M = B * NF
a = rnd(M, D).bfloat16(); w = rnd(3 * D, D, scale=D ** -0.5).bfloat16(); bias = rnd(3 * D)
rope = rope_table(NF).to(dev)
for tile in (128, 256):
    out = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
    ops.gemm(a, w, out, bias=bias, rope=rope, rope_cols=2 * D, q_scale=0.125, q_cols=D, rows_per_batch=NF, num_batches=B, tile_n=tile, variant=2)
    torch.cuda.synchronize()
    ref = (a.float() @ w.float().T + bias).view(B, NF, 3 * D // 64, 32, 2)
    c, s = rope[None, :, None, :, 0], rope[None, :, None, :, 1]
    rot = torch.stack([ref[..., 0] * c - ref[..., 1] * s, ref[..., 1] * c + ref[..., 0] * s], dim=-1)
    ref2 = ref.clone(); ref2[:, :, : 2 * D // 64] = rot[:, :, : 2 * D // 64]
    ref2 = ref2.reshape(M, 3 * D).clone(); ref2[:, :D] *= 0.125
    report(f"v2 qkv rope t{tile}", out, ref2, 4e-3)

a = rnd(M, 2048).bfloat16(); w = rnd(D, 2048, scale=2048 ** -0.5).bfloat16(); bias = rnd(D)
gate = rnd(B, 6 * D); x = rnd(M, D); x0 = x.clone()
lens = torch.tensor([937, 700], dtype=torch.int32, device=dev)
ops.gemm(a, w, x, bias=bias, resid=x, gate=gate[:, 2 * D:3 * D], row_len=lens, rows_per_batch=NF, num_batches=B, tile_n=128, variant=2)
torch.cuda.synchronize()
ref = (a.float() @ w.float().T + bias).view(B, NF, D)
valid = (torch.arange(NF, device=dev)[None] < lens[:, None]).float()[..., None]
report("v2 gate+mask+resid", x, (x0.view(B, NF, D) + gate[:, None, 2 * D:3 * D] * (ref * valid)).view(M, D), 1e-5)

# dense conv (batched tiles) through v2
Bc, Nc = 2, 500
xx = torch.zeros(Bc * Nc, 128, device=dev, dtype=torch.bfloat16); xx[:, :100] = rnd(Bc * Nc, 100).bfloat16()
wt = rnd(512, 100, 7, scale=700 ** -0.5).bfloat16(); bias = rnd(512)
wp = torch.zeros(512, 7, 128, device=dev, dtype=torch.bfloat16); wp[:, :, :100] = wt.permute(0, 2, 1)
wp = wp.reshape(512, 7 * 128).contiguous()
out = torch.empty(Bc * Nc, 512, device=dev)
ops.gemm(xx, wp, out, n=512, k=128, bias=bias, rows_per_batch=Nc, num_batches=Bc, batched_tiles=True, conv_taps=7, conv_pad=3, tile_n=256, variant=2)
torch.cuda.synchronize()
ref = F.conv1d(xx[:, :100].float().view(Bc, Nc, 100).transpose(1, 2), wt.float(), bias, padding=3)
report("v2 dense conv7", out, ref.transpose(1, 2).reshape(Bc * Nc, 512), 1e-5)

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
def timeit_cold(fn, iters=10):
    tot = 0.0
    for _ in range(iters):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / iters

for (M, N, K) in [(1874, 3072, 1024), (1874, 1024, 1024), (1874, 2048, 1024), (1874, 1024, 2048), (119936, 2048, 1024), (119936, 1024, 2048), (119936, 3072, 1024), (119936, 1024, 1024)]:
    a = rnd(M, K).bfloat16(); w = rnd(N, K).bfloat16(); out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    res = []
    for (var, tile) in [(1, 128), (1, 64), (2, 128), (2, 256)]:
        ms = timeit(lambda: ops.gemm(a, w, out, tile_n=tile, variant=var))
        msc = timeit_cold(lambda: ops.gemm(a, w, out, tile_n=tile, variant=var)) if M < 10000 else float("nan")
        res.append(f"v{var}/t{tile}: {ms*1e3:.1f}us {2*M*N*K/ms/1e9:.0f}TF (cold {msc*1e3:.1f}us)")
    ms_t = timeit(lambda: torch.matmul(a, w.T))
    print(f"time M{M} N{N} K{K}: " + " | ".join(res) + f" | cuBLAS {ms_t*1e3:.1f}us {2*M*N*K/ms_t/1e9:.0f}TF", flush=True)
print("FAILS:", fails); sys.exit(1 if fails else 0)
