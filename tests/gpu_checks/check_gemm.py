"""GPU bring-up check of the tcgen05 GEMM against a plain fp32 torch matmul of the same
bf16-rounded operands (runs on the GPU box only)."""
import sys
import torch
import torch.nn.functional as F

from f5_tts_mlx_b200 import ops

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = "cuda"
g = torch.Generator(device="cpu").manual_seed(0)
fails = []


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, generator=g) * scale).to(dev)


def report(name, got, ref, tol):
    got = got.float()
    err = (got - ref).abs().max().item()
    rel = ((got - ref).norm() / (ref.norm() + 1e-30)).item()
    ok = rel < tol and err == err
    print(f"{'OK  ' if ok else 'FAIL'} {name:38s} max_abs={err:.3e} rel_l2={rel:.3e}", flush=True)
    if not ok:
        fails.append(name)


def gelu_tanh(x):
    return F.gelu(x, approximate="tanh")


# 1. small plain GEMM, partial tiles everywhere, fp32 out
for (M, N, K, tile) in [(300, 256, 128, 0), (128, 128, 64, 128), (130, 72, 200, 64), (1874, 1024, 1024, 0)]:
    a = rnd(M, K).bfloat16(); w = rnd(N, K, scale=K ** -0.5).bfloat16()
    out = torch.full((M, N), float("nan"), device=dev)
    ops.gemm(a, w, out, tile_n=tile)
    torch.cuda.synchronize()
    report(f"plain f32 M{M} N{N} K{K} t{tile}", out, a.float() @ w.float().T, 1e-5)

# 2. QKV-like: bias + RoPE + q scale, bf16 out
B, NF, D = 2, 937, 1024
M = B * NF
a = rnd(M, D).bfloat16(); w = rnd(3 * D, D, scale=D ** -0.5).bfloat16(); bias = rnd(3 * D)
pos = torch.arange(NF, device=dev, dtype=torch.float32)
inv = 1.0 / (10000.0 ** (torch.arange(0, 64, 2, device=dev, dtype=torch.float32) / 64))
ang = pos[:, None] * inv[None, :]
rope = torch.stack([ang.cos(), ang.sin()], dim=-1).contiguous()  # [NF,32,2]
out = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
ops.gemm(a, w, out, bias=bias, rope=rope, rope_cols=2 * D, q_scale=0.125, q_cols=D,
         rows_per_batch=NF, num_batches=B)
torch.cuda.synchronize()
ref = a.float() @ w.float().T + bias
r4 = ref.view(B, NF, 3 * D // 64, 32, 2)
c, s = rope[None, :, None, :, 0], rope[None, :, None, :, 1]
x0, x1 = r4[..., 0], r4[..., 1]
rot = torch.stack([x0 * c - x1 * s, x1 * c + x0 * s], dim=-1)
nh = 2 * D // 64
ref2 = r4.clone(); ref2[:, :, :nh] = rot[:, :, :nh]
ref2 = ref2.reshape(M, 3 * D).clone(); ref2[:, :D] *= 0.125
report("qkv bias+rope+qscale bf16", out, ref2, 4e-3)

# 3. out-proj-like: bias, row mask, gate, residual (fp32, in place)
a = rnd(M, 2048).bfloat16(); w = rnd(D, 2048, scale=2048 ** -0.5).bfloat16(); bias = rnd(D)
gate = rnd(B, 6 * D); x = rnd(M, D); x0 = x.clone()
lens = torch.tensor([937, 700], dtype=torch.int32, device=dev)
ops.gemm(a, w, x, bias=bias, resid=x, gate=gate[:, 2 * D:3 * D], row_len=lens,
         rows_per_batch=NF, num_batches=B)
torch.cuda.synchronize()
ref = (a.float() @ w.float().T + bias).view(B, NF, D)
valid = (torch.arange(NF, device=dev)[None, :] < lens[:, None]).float()[..., None]
ref = x0.view(B, NF, D) + gate[:, None, 2 * D:3 * D] * (ref * valid)
report("out-proj mask+gate+resid f32", x, ref.view(M, D), 1e-5)

# 4. FF1-like GELU-tanh bf16 / GELU-erf / Mish
a = rnd(M, D).bfloat16(); w = rnd(2048, D, scale=D ** -0.5).bfloat16(); bias = rnd(2048)
for act, fn, nm in [(ops.ACT_GELU_TANH, gelu_tanh, "gelu_tanh"), (ops.ACT_GELU_ERF, F.gelu, "gelu_erf"),
                    (ops.ACT_MISH, F.mish, "mish")]:
    out = torch.empty(M, 2048, device=dev, dtype=torch.bfloat16)
    ops.gemm(a, w, out, bias=bias, act=act)
    torch.cuda.synchronize()
    report(f"ff1 {nm} bf16", out, fn(a.float() @ w.float().T + bias), 4e-3)

# 5. proj_out-like N=100 fp32
w = rnd(100, D, scale=D ** -0.5).bfloat16(); bias = rnd(100)
out = torch.full((M, 100), float("nan"), device=dev)
ops.gemm(a, w, out, bias=bias)
torch.cuda.synchronize()
report("proj_out N=100 f32", out, a.float() @ w.float().T + bias, 1e-5)

# 6. grouped conv (ConvPositionEmbedding): k=31, groups of 64 channels, Mish
for (Bc, Nc, Cc) in [(2, 937, 1024), (1, 200, 128)]:
    x = rnd(Bc * Nc, Cc).bfloat16()
    wt = rnd(Cc, 64, 31, scale=(64 * 31) ** -0.5).bfloat16()   # torch layout (O, I/g, K)
    bias = rnd(Cc)
    wp = wt.permute(0, 2, 1).reshape(Cc, 31 * 64).contiguous()  # (O, tap, I/g)
    out = torch.empty(Bc * Nc, Cc, device=dev, dtype=torch.bfloat16)
    ops.gemm(x, wp, out, n=Cc, k=64, bias=bias, act=ops.ACT_MISH, rows_per_batch=Nc, num_batches=Bc,
             batched_tiles=True, conv_taps=31, conv_pad=15, conv_grouped=True)
    torch.cuda.synchronize()
    xr = x.float().view(Bc, Nc, Cc).transpose(1, 2)
    ref = F.mish(F.conv1d(xr, wt.float(), bias, padding=15, groups=Cc // 64)).transpose(1, 2).reshape(Bc * Nc, Cc)
    report(f"grouped conv31 mish B{Bc} N{Nc} C{Cc}", out, ref, 4e-3)

# 7. dense conv (Vocos embed): 100 -> 512, k=7, input padded to 128 channels
Bc, Nc = 1, 500
x = torch.zeros(Bc * Nc, 128, device=dev, dtype=torch.bfloat16); x[:, :100] = rnd(Bc * Nc, 100).bfloat16()
wt = rnd(512, 100, 7, scale=700 ** -0.5).bfloat16(); bias = rnd(512)
wp = torch.zeros(512, 7, 128, device=dev, dtype=torch.bfloat16); wp[:, :, :100] = wt.permute(0, 2, 1)
wp = wp.reshape(512, 7 * 128).contiguous()
out = torch.empty(Bc * Nc, 512, device=dev)
ops.gemm(x, wp, out, n=512, k=128, bias=bias, rows_per_batch=Nc, num_batches=Bc, batched_tiles=True,
         conv_taps=7, conv_pad=3)
torch.cuda.synchronize()
ref = F.conv1d(x[:, :100].float().view(Bc, Nc, 100).transpose(1, 2), wt.float(), bias, padding=3)
report("dense conv7 100->512 f32", out, ref.transpose(1, 2).reshape(Bc * Nc, 512), 1e-5)

# 8. timing
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

for (M, N, K, tile) in [(1874, 3072, 1024, 128), (1874, 1024, 1024, 128), (1874, 1024, 1024, 64), (1874, 2048, 1024, 128),
                        (1874, 1024, 2048, 128), (119936, 2048, 1024, 128), (119936, 1024, 2048, 128), (119936, 3072, 1024, 128)]:
    a = rnd(M, K).bfloat16(); w = rnd(N, K).bfloat16(); out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.gemm(a, w, out, tile_n=tile))
    ms_t = timeit(lambda: torch.matmul(a, w.T))
    print(f"time M{M} N{N} K{K} t{tile}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s   (torch/cuBLAS {ms_t*1e3:.1f} us {2*M*N*K/ms_t/1e9:.1f} TFLOP/s)", flush=True)

print("FAILS:", fails)
sys.exit(1 if fails else 0)
