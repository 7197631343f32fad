// Fused GEMM epilogue shared by the single-CTA (gemm_sm100.cuh) and the 2-CTA persistent
// (gemm2_sm100.cuh) kernels: one thread owns one output row and 32 consecutive accumulator columns.
#pragma once
#include "ptx.cuh"

namespace f5 {

enum GemmAct { ACT_NONE = 0, ACT_GELU_TANH = 1, ACT_GELU_ERF = 2, ACT_MISH = 3 };

struct GemmParams {
  int M, N, K;             // logical problem (flat mode: M rows; batched mode: see below)
  // row mapping
  int rows_per_batch;      // frames per utterance (row -> batch = row / rows_per_batch); 0 = M
  int tiles_per_batch;     // >0: M tiles never straddle utterances (batched / conv mode)
  int num_batches;
  // implicit-conv mode
  int conv_taps;           // 1 = plain GEMM
  int conv_pad;            // frames of left padding (k//2)
  int k_per_tap;           // K elements per tap (multiple of 64)
  int conv_grouped;        // 1: A channel base = output column base (block-diagonal groups of 64)
  // epilogue
  const float* bias;       // [N] or null
  void* out;               // bf16 or f32, row-major, ldo elements
  int ldo;
  const float* resid;      // f32 [rows, ldr] or null (may alias out)
  int ldr;
  const float* gate;       // f32 [num_batches, gate_ld] or null
  int gate_ld;
  const int* row_len;      // [num_batches] valid frames per utterance, or null
  const float2* rope;      // [rows_per_batch, 32] (cos, sin), or null
  int rope_cols;
  float q_scale;
  int q_cols;
  __nv_bfloat16* out2;     // optional bf16 copy of the result
  int ldo2;
  unsigned long long* ts;  // debug: per-CTA phase timestamps (globaltimer ns), 10 slots per CTA, or null
};

__device__ __forceinline__ void ts_mark(const GemmParams& p, int cta, int slot) {
  if (p.ts != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.ts[(size_t)cta * 10 + slot] = t;
  }
}

// acc: 32 fp32 accumulator columns [col0, col0+32) of output row `row` (utterance b_idx, frame pos)
template <int ACT, bool OUT_BF16, bool ROPE>
__device__ __forceinline__ void gemm_epilogue_chunk(const uint32_t (&acc)[32], const GemmParams& p,
                                                    int col0, int row, int pos, int b_idx,
                                                    bool row_ok, bool row_valid) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (col0 + j < p.N) {
        float4 bb = *reinterpret_cast<const float4*>(p.bias + col0 + j);
        v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
      }
    }
  }
  if (ACT == ACT_GELU_TANH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_f(v[j]);
  } else if (ACT == ACT_GELU_ERF) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf_f(v[j]);
  } else if (ACT == ACT_MISH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = mish_f(v[j]);
  }
  if (ROPE) {
    if (col0 < p.rope_cols) {
      const float2* rp = p.rope + (size_t)pos * 32 + ((col0 & 63) >> 1);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float2 cs = rp[j];
        float a = v[2 * j], b = v[2 * j + 1];
        v[2 * j] = a * cs.x - b * cs.y;
        v[2 * j + 1] = b * cs.x + a * cs.y;
      }
    }
    if (col0 < p.q_cols) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] *= p.q_scale;
    }
  }
  if (!row_valid) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
  }
  if (p.gate != nullptr) {
    const float* g = p.gate + (size_t)b_idx * p.gate_ld + col0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (col0 + j < p.N) {
        float4 gg = *reinterpret_cast<const float4*>(g + j);
        v[j] *= gg.x; v[j + 1] *= gg.y; v[j + 2] *= gg.z; v[j + 3] *= gg.w;
      }
    }
  }
  if (row_ok) {
    if (p.resid != nullptr) {
      const float* rr = p.resid + (size_t)row * p.ldr + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (col0 + j < p.N) {
          float4 x = *reinterpret_cast<const float4*>(rr + j);
          v[j] += x.x; v[j + 1] += x.y; v[j + 2] += x.z; v[j + 3] += x.w;
        }
      }
    }
    if (p.out2 != nullptr) {
      __nv_bfloat16* o2 = p.out2 + (size_t)row * p.ldo2 + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        if (col0 + j < p.N) {
          uint4 w;
          w.x = pack_bf16x2(v[j], v[j + 1]);
          w.y = pack_bf16x2(v[j + 2], v[j + 3]);
          w.z = pack_bf16x2(v[j + 4], v[j + 5]);
          w.w = pack_bf16x2(v[j + 6], v[j + 7]);
          *reinterpret_cast<uint4*>(o2 + j) = w;
        }
      }
    }
    if (OUT_BF16) {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldo + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        if (col0 + j < p.N) {
          uint4 w;
          w.x = pack_bf16x2(v[j], v[j + 1]);
          w.y = pack_bf16x2(v[j + 2], v[j + 3]);
          w.z = pack_bf16x2(v[j + 4], v[j + 5]);
          w.w = pack_bf16x2(v[j + 6], v[j + 7]);
          *reinterpret_cast<uint4*>(o + j) = w;
        }
      }
    } else {
      float* o = reinterpret_cast<float*>(p.out) + (size_t)row * p.ldo + col0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (col0 + j < p.N) {
          *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
    }
  }
}

}  // namespace f5
