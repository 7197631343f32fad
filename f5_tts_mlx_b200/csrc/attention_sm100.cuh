// Non-causal flash-attention forward on tcgen05 (sm_100a), head_dim 64.
// Replaces mx.fast.scaled_dot_product_attention(q, k, v, scale, mask=key-padding) at dit.py:166
// (fp32 softmax, as MLX) plus the head split / merge transposes at dit.py:141-143,167.
//
// Input  qkv : bf16 [B*N, 3*H*64] — the QKV GEMM output as is; q already rotated and multiplied
//              by the softmax scale, k rotated (gemm_sm100.cuh epilogue).  Head h of q/k/v is the
//              64-column slice at h*64 / H*64+h*64 / 2*H*64+h*64: the head split is just a TMA
//              coordinate, no transpose is ever materialised.
// Output o   : bf16 [B*N, H*64]  (heads merged, ready to be the A operand of the out-projection).
//
// One CTA = one (batch, head, 128-query tile); loop over 128-key tiles:
//   warp 0   TMA: Q once, K/V ring (2 stages each)
//   warp 1   MMA issuer: S = Q K^T (M128 N128 K64) -> TMEM cols [0,128);
//                        O_j = P V  (M128 N64 K128, V as MN-major B) -> TMEM cols [128,192)
//   warps 2-5 softmax (thread = query row): S from TMEM, online max/sum in fp32, P -> bf16 into a
//            128B-swizzled smem tile (the A operand of the PV MMA), O_acc = O_acc*alpha + O_j in
//            registers.
// 112 KB smem + 256 TMEM columns per CTA, so two CTAs share an SM and one's softmax overlaps the
// other's MMAs.
#pragma once
#include "ptx.cuh"

namespace f5 {

struct AttnParams {
  int B, N, H;
  const int* kv_len;       // [B] valid keys per utterance, or null (= N)
  __nv_bfloat16* out;      // [B*N, H*64]
  int ldo;
  unsigned long long* ts;  // debug: [3 roles][64 tiles][8 slots] SM-clock stamps of CTA (0,0,0), or null
  int handoff;             // v2: softmax groups alternate in the exponential loop (F5_ATTN_HANDOFF=0 disables)
  unsigned long long* prof;  // in-graph timing slot (ptx.cuh prof_stamp_*), or null
  int out_fp8;             // v2: `out` receives e4m3 bytes (ldo in bytes) — the A operand of an FP8-mode out-projection
};

struct AttnSmem {
  static constexpr int kQ = 0;                       // 128 x 64 bf16
  static constexpr int kK = 16384;                   // 2 stages x 16 KB
  static constexpr int kV = kK + 2 * 16384;          // 2 stages x 16 KB
  static constexpr int kP = kV + 2 * 16384;          // 2 k-blocks x (128 x 64 bf16)
  static constexpr int kBar = kP + 32768;
  // barriers: q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full, p_full, o_full
  static constexpr int kNumBars = 12;
  static constexpr int kTotal = kBar + kNumBars * 8 + 16;
};

__global__ void __launch_bounds__(192, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tma_qkv, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnSmem::kBar);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;   // [2]
  uint64_t* v_empty = bars + 7;  // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* p_full = bars + 10;
  uint64_t* o_full = bars + 11;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + AttnSmem::kNumBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int HD = p.H * 64;
  int kv_len = p.kv_len ? p.kv_len[b] : p.N;
  kv_len = min(max(kv_len, 1), p.N);
  const int num_kv = (kv_len + 127) >> 7;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_qkv);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(q_full, 16384);
      tma_load_3d(smem + AttnSmem::kQ, &tma_qkv, q_full, h * 64, q0, b);
      for (int j = 0; j < num_kv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_expect_tx(&k_full[s], 16384);
        tma_load_3d(smem + AttnSmem::kK + s * 16384, &tma_qkv, &k_full[s], HD + h * 64, j * 128, b);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_expect_tx(&v_full[s], 16384);
        tma_load_3d(smem + AttnSmem::kV + s * 16384, &tma_qkv, &v_full[s], 2 * HD + h * 64, j * 128, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);  // B (=V) is MN-major
    const uint32_t sQ = smem_u32(smem + AttnSmem::kQ);
    const uint32_t sP = smem_u32(smem + AttnSmem::kP);
    mbar_wait(q_full, 0);
    for (int j = 0; j < num_kv; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      // ---- S = Q K^T ----
      mbar_wait(&k_full[s], ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sK = smem_u32(smem + AttnSmem::kK + s * 16384);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16_ss(tmem_S, umma_desc_sw128(sQ + k * 32, 16, 1024),
                      umma_desc_sw128(sK + k * 32, 16, 1024), idesc_s, k != 0);
        tc_commit(&k_empty[s]);
        tc_commit(s_full);
      }
      __syncwarp();
      // ---- O_j = P V ----
      mbar_wait(&v_full[s], ph);
      mbar_wait(p_full, j & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sV = smem_u32(smem + AttnSmem::kV + s * 16384);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // A: P k-block (k/4) of 64 columns, +32 B per 16-column step inside the swizzle span
          const uint64_t da = umma_desc_sw128(sP + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
          // B: V rows [16k, 16k+16) = two 8-row groups 1024 B apart; 64 d-columns = one span
          const uint64_t db = umma_desc_sw128(sV + k * 2048, 16384, 1024);
          umma_f16_ss(tmem_O, da, db, idesc_o, k != 0);
        }
        tc_commit(&v_empty[s]);
        tc_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax / output =====================
    const int lg = warp & 3;
    const int r = lg * 32 + lane;                       // query row within the tile
    const uint32_t lane_addr = (uint32_t)(lg * 32) << 16;
    uint8_t* sP = smem + AttnSmem::kP;
    constexpr float kLog2e = 1.4426950408889634f;
    float m_run = -INFINITY, l_run = 0.f;
    float o_acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o_acc[i] = 0.f;

    for (int j = 0; j < num_kv; ++j) {
      const uint32_t ph = j & 1;
      mbar_wait(s_full, ph);
      tc_fence_after();
      const int kv0 = j * 128;
      const bool partial = kv0 + 128 > kv_len;
      // pass 1: row max
      float m_new = m_run;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t sv[32];
        tmem_ld32(tmem_S + lane_addr + c * 32, sv);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float x = __uint_as_float(sv[i]);
          if (partial && kv0 + c * 32 + i >= kv_len) x = -INFINITY;
          m_new = fmaxf(m_new, x);
        }
      }
      const float alpha = ex2_approx((m_run - m_new) * kLog2e);
      const float mb = m_new * kLog2e;
      float l_tile = 0.f;
      // pass 2: p = exp2(s*log2e - m*log2e), write bf16 P into the swizzled A-operand tile
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t sv[32];
        tmem_ld32(tmem_S + lane_addr + c * 32, sv);
        tmem_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float x0 = __uint_as_float(sv[i]), x1 = __uint_as_float(sv[i + 1]);
          float p0 = ex2_approx(fmaf(x0, kLog2e, -mb));
          float p1 = ex2_approx(fmaf(x1, kLog2e, -mb));
          if (partial) {
            if (kv0 + c * 32 + i >= kv_len) p0 = 0.f;
            if (kv0 + c * 32 + i + 1 >= kv_len) p1 = 0.f;
          }
          // the row sum uses the bf16-rounded probabilities that the PV MMA will see
          __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
          float2 pr = __bfloat1622float2(pb);
          l_tile += pr.x + pr.y;
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&pb);
        }
        uint8_t* blk = sP + (c >> 1) * 16384 + r * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = ((c & 1) * 4 + q) ^ (r & 7);
          *reinterpret_cast<uint4*>(blk + chunk * 16) =
              make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
        }
      }
      // hand P to the MMA warp: generic-proxy writes -> async proxy, and S fully consumed
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(p_full);
      l_run = l_run * alpha + l_tile;
      m_run = m_new;
      // O_acc = O_acc * alpha + O_j
      mbar_wait(o_full, ph);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t ov[32];
        tmem_ld32(tmem_O + lane_addr + c * 32, ov);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] = fmaf(o_acc[c * 32 + i], alpha, __uint_as_float(ov[i]));
      }
      tc_fence_before();
    }
    // epilogue: normalise, bf16, heads merged
    pdl_launch_dependents();
    const int n = q0 + r;
    if (n < p.N) {
      const float inv = 1.f / l_run;
      __nv_bfloat16* o = p.out + ((size_t)b * p.N + n) * p.ldo + h * 64;
#pragma unroll
      for (int i = 0; i < 64; i += 8) {
        uint4 w;
        w.x = pack_bf16x2(o_acc[i] * inv, o_acc[i + 1] * inv);
        w.y = pack_bf16x2(o_acc[i + 2] * inv, o_acc[i + 3] * inv);
        w.z = pack_bf16x2(o_acc[i + 4] * inv, o_acc[i + 5] * inv);
        w.w = pack_bf16x2(o_acc[i + 6] * inv, o_acc[i + 7] * inv);
        *reinterpret_cast<uint4*>(o + i) = w;
      }
    }
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace f5
