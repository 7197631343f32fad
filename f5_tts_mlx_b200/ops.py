"""Thin torch-tensor wrappers over the C ABI (pointer + shape marshalling only; no arithmetic).

torch is used for device memory and streams.  Every function launches on the current torch CUDA
stream and returns its output tensor; nothing here falls back to a torch kernel.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF, ACT_MISH = 0, 1, 2, 3


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: torch.Tensor | None) -> C.c_void_p | None:
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def _need_cuda(*ts: torch.Tensor | None) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.F5Error("f5_tts_mlx_b200 ops need CUDA tensors (no CPU fallback exists)")


def gemm(
    a: torch.Tensor,            # bf16 [rows, >=k] (row stride = a.stride(0))
    w: torch.Tensor,            # bf16 [n, taps*k_pad]
    out: torch.Tensor,          # bf16 / f32 [rows, >=n]
    *,
    n: int | None = None,
    k: int | None = None,
    bias: torch.Tensor | None = None,
    act: int = ACT_NONE,
    resid: torch.Tensor | None = None,
    gate: torch.Tensor | None = None,      # f32 [num_batches, n] view (row stride honoured)
    row_len: torch.Tensor | None = None,   # i32 [num_batches]
    rope: torch.Tensor | None = None,      # f32 [rows_per_batch, 32, 2]
    rope_cols: int = 0,
    q_scale: float = 1.0,
    q_cols: int = 0,
    rows_per_batch: int = 0,
    num_batches: int = 1,
    batched_tiles: bool = False,
    conv_taps: int = 1,
    conv_pad: int = 0,
    conv_grouped: bool = False,
    tile_n: int = 0,
    variant: int = 0,
    debug_ts: torch.Tensor | None = None,
    out2: torch.Tensor | None = None,          # bf16 [rows, >=n] second copy (or the fused-LN operand, see ln_scale)
    ln_scale: torch.Tensor | None = None,      # f32 [n]: producer mode — out2 = bf16(out * (1 + ln_scale)), ln_stats filled
    ln_stats: torch.Tensor | None = None,      # f32 [rows, n/64, 2] (sum, sum of squares) per 64 columns
    ln_in_stats: torch.Tensor | None = None,   # f32 [rows, k/64, 2]: consumer mode
    ln_tab: torch.Tensor | None = None,        # f32 [4, >=n] rows c1_hi, c1_lo, c2_hi, c2_lo
    ab_fp8: bool = False,                      # a and w are e4m3 bytes (uint8 / float8_e4m3fn tensors), k % 128 == 0
    acc_scale: float = 1.0,                    # multiplies the accumulator in FP8 mode (weight tensor scale)
    out2_fp8: bool = False,                    # out2 is written as e4m3 bytes (uint8 tensor)
    out_fp8: bool = False,                     # out (a uint8 tensor) is written as e4m3 bytes
) -> torch.Tensor:
    _need_cuda(a, w, out, bias, resid, gate, row_len, rope, out2, ln_scale, ln_stats, ln_in_stats, ln_tab)
    if ab_fp8:
        a = a.view(torch.uint8) if a.dtype != torch.uint8 else a
        w = w.view(torch.uint8) if w.dtype != torch.uint8 else w
    else:
        assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    assert a.stride(-1) == 1 and w.stride(-1) == 1 and out.stride(-1) == 1
    m = a.shape[0]
    g = _lib.GemmArgs()
    g.a, g.lda = a.data_ptr(), a.stride(0)
    g.w, g.ldw = w.data_ptr(), w.stride(0)
    g.m = m
    g.n = n if n is not None else w.shape[0]
    g.k = k if k is not None else (64 if conv_grouped else a.shape[1])
    g.rows_per_batch = rows_per_batch
    g.num_batches = num_batches
    g.batched_tiles = int(batched_tiles)
    g.conv_taps, g.conv_pad, g.conv_grouped = conv_taps, conv_pad, int(conv_grouped)
    g.act = act
    g.out_bf16 = int(out.dtype == torch.bfloat16 or out_fp8)
    g.out_fp8 = int(out_fp8)
    assert out.dtype in (torch.bfloat16, torch.float32) or (out_fp8 and out.dtype == torch.uint8)
    g.bias = bias.data_ptr() if bias is not None else None
    g.out, g.ldo = out.data_ptr(), out.stride(0)
    if resid is not None:
        assert resid.dtype == torch.float32
        g.resid, g.ldr = resid.data_ptr(), resid.stride(0)
    if gate is not None:
        assert gate.dtype == torch.float32 and gate.stride(-1) == 1
        g.gate, g.gate_ld = gate.data_ptr(), gate.stride(0) if gate.dim() > 1 else 0
    if row_len is not None:
        assert row_len.dtype == torch.int32
        g.row_len = row_len.data_ptr()
    if rope is not None:
        assert rope.dtype == torch.float32 and rope.is_contiguous()
        g.rope = rope.data_ptr()
    g.rope_cols, g.q_scale, g.q_cols = rope_cols, q_scale, q_cols
    g.tile_n = tile_n
    g.variant = variant
    if debug_ts is not None:
        g.debug_ts = debug_ts.data_ptr()
    g.ab_fp8, g.acc_scale, g.out2_fp8 = int(ab_fp8), float(acc_scale), int(out2_fp8)
    if out2 is not None:
        assert out2.dtype == (torch.uint8 if out2_fp8 else torch.bfloat16) and out2.stride(-1) == 1
        g.out2_bf16, g.ldo2 = out2.data_ptr(), out2.stride(0)
    if ln_scale is not None:
        assert ln_scale.dtype == torch.float32 and ln_stats is not None and ln_stats.dtype == torch.float32
        g.ln_scale, g.ln_stats = ln_scale.data_ptr(), ln_stats.data_ptr()
    if ln_in_stats is not None:
        assert ln_in_stats.dtype == torch.float32 and ln_tab is not None and ln_tab.dtype == torch.float32
        g.ln_in_stats, g.ln_tab, g.ln_tab_ld = ln_in_stats.data_ptr(), ln_tab.data_ptr(), ln_tab.stride(0)
    _lib.check(_lib.load().f5_gemm_bf16(C.byref(g), _stream()))
    return out
