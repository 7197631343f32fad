"""GPU bring-up check of the tcgen05 flash attention against explicit fp32 softmax attention."""
import ctypes as C, sys, math
import torch
from f5_tts_mlx_b200 import _lib

torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
g = torch.Generator().manual_seed(0)
lib = _lib.load()
fails = []

def run(B, N, H, kv_len=None, scale_in=1.0):
    D = H * 64
    qkv = (torch.randn(B * N, 3 * D, generator=g) * scale_in).to(dev).bfloat16()
    out = torch.full((B * N, D), float("nan"), device=dev, dtype=torch.bfloat16)
    kl = None
    if kv_len is not None:
        kl = torch.tensor(kv_len, dtype=torch.int32, device=dev)
    _lib.check(lib.f5_attention_fwd(qkv.data_ptr(), 3 * D, out.data_ptr(), D, B, N, H, 64,
                                    kl.data_ptr() if kl is not None else None,
                                    torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    q, k, v = [t.float().view(B, N, H, 64).permute(0, 2, 1, 3) for t in qkv.split(D, dim=1)]
    s = q @ k.transpose(-1, -2)            # q is "pre-scaled" by convention
    if kl is not None:
        m = torch.arange(N, device=dev)[None, :] < kl[:, None]
        s = s.masked_fill(~m[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * N, D)
    got = out.float()
    err = (got - ref).abs().max().item(); rel = ((got - ref).norm() / ref.norm()).item()
    ok = rel < 1e-2 and err == err
    print(f"{'OK  ' if ok else 'FAIL'} attn B{B} N{N} H{H} kv={kv_len} max_abs={err:.3e} rel_l2={rel:.3e}", flush=True)
    if not ok: fails.append((B, N, H, kv_len))

run(1, 128, 1)
run(1, 256, 2)
run(1, 100, 1)
run(2, 937, 16, scale_in=0.35)
run(2, 937, 16, kv_len=[937, 500], scale_in=0.35)
run(3, 300, 8, kv_len=[300, 129, 1])
run(1, 1500, 4, scale_in=0.5)

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

for (B, N, H) in [(2, 937, 16), (128, 937, 16), (2, 5625, 16)]:
    D = H * 64
    qkv = torch.randn(B * N, 3 * D, device=dev).bfloat16() * 0.3
    out = torch.empty(B * N, D, device=dev, dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    ms = timeit(lambda: lib.f5_attention_fwd(qkv.data_ptr(), 3 * D, out.data_ptr(), D, B, N, H, 64, None, st))
    fl = 4.0 * B * H * N * N * 64
    q, k, v = [t.view(B, N, H, 64).permute(0, 2, 1, 3) for t in qkv.split(D, dim=1)]
    ms_t = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v, scale=1.0))
    print(f"time attn B{B} N{N} H{H}: {ms*1e3:.1f} us {fl/ms/1e9:.1f} TFLOP/s  (torch sdpa {ms_t*1e3:.1f} us {fl/ms_t/1e9:.1f} TFLOP/s)", flush=True)
print("FAILS:", fails)
sys.exit(1 if fails else 0)
