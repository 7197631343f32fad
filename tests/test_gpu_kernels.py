"""-m gpu: per-kernel parity of the C-ABI entry points against a plain fp32 torch reference of the
same op on the same bf16-rounded operands (tolerances: fp32-output GEMMs 1e-5 rel; anything that
rounds its output to bf16 4e-3 rel = 2^-8)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
dev = "cuda"


@pytest.fixture(autouse=True)
def _strict():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(dev)


def rel(a, b):
    return ((a.float() - b).norm() / (b.norm() + 1e-30)).item()


@pytest.mark.parametrize("M,N,K,tile", [(300, 256, 128, 0), (128, 128, 64, 128), (130, 72, 200, 64), (1, 128, 64, 0),
                                        (1874, 1024, 1024, 0), (257, 100, 1024, 0)])
def test_gemm_plain_fp32(M, N, K, tile):
    from f5_tts_mlx_b200 import ops
    a = rnd(M, K).bfloat16(); w = rnd(N, K, scale=K ** -0.5).bfloat16(); bias = rnd(N)
    out = torch.full((M, N), float("nan"), device=dev)
    ops.gemm(a, w, out, bias=bias, tile_n=tile)
    assert rel(out, a.float() @ w.float().T + bias) < 1e-5


def test_gemm_qkv_rope_epilogue():
    from f5_tts_mlx_b200 import ops
    from f5_tts_mlx_b200.dit import rope_table
    B, NF, D = 2, 937, 1024
    M = B * NF
    a = rnd(M, D).bfloat16(); w = rnd(3 * D, D, scale=D ** -0.5).bfloat16(); bias = rnd(3 * D)
    rope = rope_table(NF).to(dev)
    out = torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
    ops.gemm(a, w, out, bias=bias, rope=rope, rope_cols=2 * D, q_scale=0.125, q_cols=D, rows_per_batch=NF, num_batches=B)
    ref = (a.float() @ w.float().T + bias).view(B, NF, 3 * D // 64, 32, 2)
    c, s = rope[None, :, None, :, 0], rope[None, :, None, :, 1]
    rot = torch.stack([ref[..., 0] * c - ref[..., 1] * s, ref[..., 1] * c + ref[..., 0] * s], dim=-1)
    ref2 = ref.clone(); ref2[:, :, : 2 * D // 64] = rot[:, :, : 2 * D // 64]
    ref2 = ref2.reshape(M, 3 * D).clone(); ref2[:, :D] *= 0.125
    assert rel(out, ref2) < 4e-3


def test_gemm_gate_mask_residual_inplace():
    from f5_tts_mlx_b200 import ops
    B, NF, D = 2, 937, 1024
    M = B * NF
    a = rnd(M, 2048).bfloat16(); w = rnd(D, 2048, scale=2048 ** -0.5).bfloat16(); bias = rnd(D)
    gate = rnd(B, 6 * D); x = rnd(M, D); x0 = x.clone()
    lens = torch.tensor([937, 700], dtype=torch.int32, device=dev)
    ops.gemm(a, w, x, bias=bias, resid=x, gate=gate[:, 2 * D:3 * D], row_len=lens, rows_per_batch=NF, num_batches=B)
    ref = (a.float() @ w.float().T + bias).view(B, NF, D)
    valid = (torch.arange(NF, device=dev)[None] < lens[:, None]).float()[..., None]
    ref = x0.view(B, NF, D) + gate[:, None, 2 * D:3 * D] * (ref * valid)
    assert rel(x, ref.view(M, D)) < 1e-5


@pytest.mark.parametrize("act,fn", [(1, lambda v: F.gelu(v, approximate="tanh")), (2, F.gelu), (3, F.mish)])
def test_gemm_activations(act, fn):
    from f5_tts_mlx_b200 import ops
    a = rnd(500, 512).bfloat16(); w = rnd(1024, 512, scale=512 ** -0.5).bfloat16(); bias = rnd(1024)
    out = torch.empty(500, 1024, device=dev, dtype=torch.bfloat16)
    ops.gemm(a, w, out, bias=bias, act=act)
    assert rel(out, fn(a.float() @ w.float().T + bias)) < 4e-3


@pytest.mark.parametrize("B,N,C", [(2, 937, 1024), (1, 200, 128), (3, 31, 512), (1, 1, 64)])
def test_grouped_conv31_implicit_gemm(B, N, C):
    from f5_tts_mlx_b200 import ops
    x = rnd(B * N, C).bfloat16()
    wt = rnd(C, 64, 31, scale=(64 * 31) ** -0.5).bfloat16(); bias = rnd(C)
    wp = wt.permute(0, 2, 1).reshape(C, 31 * 64).contiguous()
    out = torch.empty(B * N, C, device=dev, dtype=torch.bfloat16)
    ops.gemm(x, wp, out, n=C, k=64, bias=bias, act=3, rows_per_batch=N, num_batches=B, batched_tiles=True,
             conv_taps=31, conv_pad=15, conv_grouped=True)
    ref = F.mish(F.conv1d(x.float().view(B, N, C).transpose(1, 2), wt.float(), bias, padding=15, groups=C // 64))
    assert rel(out, ref.transpose(1, 2).reshape(B * N, C)) < 4e-3


def _attn(B, N, H, kv_len=None, scale_in=1.0, seed=0):
    from f5_tts_mlx_b200 import _lib
    D = H * 64
    qkv = rnd(B * N, 3 * D, scale=scale_in, seed=seed).bfloat16()
    out = torch.full((B * N, D), float("nan"), device=dev, dtype=torch.bfloat16)
    kl = torch.tensor(kv_len, dtype=torch.int32, device=dev) if kv_len is not None else None
    _lib.check(_lib.load().f5_attention_fwd(qkv.data_ptr(), 3 * D, out.data_ptr(), D, B, N, H, 64,
                                            kl.data_ptr() if kl is not None else None,
                                            torch.cuda.current_stream().cuda_stream))
    q, k, v = [t.float().view(B, N, H, 64).permute(0, 2, 1, 3) for t in qkv.split(D, dim=1)]
    s = q @ k.transpose(-1, -2)
    if kl is not None:
        m = torch.arange(N, device=dev)[None] < kl[:, None]
        s = s.masked_fill(~m[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * N, D)
    return out, ref


@pytest.mark.parametrize("B,N,H,kv,sc", [(1, 128, 1, None, 1.0), (1, 100, 1, None, 1.0), (2, 937, 16, None, 0.35),
                                         (2, 937, 16, [937, 500], 0.35), (3, 300, 8, [300, 129, 1], 1.0),
                                         (1, 1500, 4, None, 0.5), (1, 1, 2, None, 1.0)])
def test_attention(B, N, H, kv, sc):
    out, ref = _attn(B, N, H, kv, sc)
    assert torch.isfinite(out.float()).all()
    assert rel(out, ref) < 6e-3


def test_attention_long_sequence_properties():
    """N = 5625 (60 s): rows of softmax sum to one -> attention of constant V returns that constant."""
    from f5_tts_mlx_b200 import _lib
    B, N, H = 1, 5625, 16
    D = H * 64
    qkv = rnd(B * N, 3 * D, scale=0.3).bfloat16()
    qkv[:, 2 * D:] = 0.5
    out = torch.empty(B * N, D, device=dev, dtype=torch.bfloat16)
    _lib.check(_lib.load().f5_attention_fwd(qkv.data_ptr(), 3 * D, out.data_ptr(), D, B, N, H, 64, None,
                                            torch.cuda.current_stream().cuda_stream))
    assert (out.float() - 0.5).abs().max().item() < 4e-3


@pytest.mark.parametrize("D", [512, 1024])
def test_ln_modulate(D):
    from f5_tts_mlx_b200 import _lib
    rows = 777
    x = rnd(rows, D) * 3 + 1; mod = rnd(6 * D)
    y = torch.empty(rows, D, device=dev, dtype=torch.bfloat16)
    _lib.check(_lib.load().f5_ln_modulate(x.data_ptr(), y.data_ptr(), rows, D, 0, mod[D:].data_ptr(), mod.data_ptr(), 0, 1,
                                          torch.cuda.current_stream().cuda_stream))
    ref = F.layer_norm(x, (D,), eps=1e-6) * (1 + mod[D:2 * D]) + mod[:D]
    assert rel(y, ref) < 4e-3


def test_dwconv7_ln_and_grn():
    from f5_tts_mlx_b200 import _lib
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    B, N, Cc = 2, 333, 512
    x = rnd(B, N, Cc); w = rnd(Cc, 7, scale=0.4); wb = rnd(Cc); lw = 1 + 0.1 * rnd(Cc); lb = 0.1 * rnd(Cc)
    y = torch.empty(B * N, Cc, device=dev, dtype=torch.bfloat16)
    wt = w.t().contiguous()
    _lib.check(lib.f5_dwconv7_ln(x.data_ptr(), y.data_ptr(), B, N, Cc, wt.data_ptr(), wb.data_ptr(), lw.data_ptr(),
                                 lb.data_ptr(), st))
    ref = F.conv1d(x.transpose(1, 2), w[:, None, :], wb, padding=3, groups=Cc).transpose(1, 2)
    ref = F.layer_norm(ref, (Cc,), lw, lb, eps=1e-6)
    assert rel(y, ref.reshape(B * N, Cc)) < 4e-3
    Ci = 1024
    h = rnd(B, N, Ci).bfloat16(); gamma = rnd(Ci) * 0.5; beta = rnd(Ci) * 0.5
    out = torch.empty_like(h); nx = torch.empty(B, 1 + (N + 31) // 32, Ci, device=dev)
    _lib.check(lib.f5_grn(h.data_ptr(), out.data_ptr(), nx.data_ptr(), gamma.data_ptr(), beta.data_ptr(), B, N, Ci, st))
    hf = h.float()
    Gx = hf.pow(2).sum(1, keepdim=True).sqrt()
    ref = gamma * (hf * (Gx / (Gx.mean(-1, keepdim=True) + 1e-6))) + beta + hf
    assert rel(out, ref) < 4e-3
