// Flash-attention forward v3 for sm_100a: attention2_sm100.cuh with SIXTEEN softmax warps — every
// query row is shared by two threads (64 key columns each), i.e. 4 softmax warps per scheduler
// instead of 2, to keep the MUFU pipe (16 ex2/clk/SM, the bound of this kernel at head_dim 64) fed
// while other warps sit in their load / max / pack phases.  The halves exchange their partial row
// maxima through shared memory once per tile (named barrier per group).
// --- original header of v2 follows ---
// Flash-attention forward v2 for sm_100a (head_dim 64, non-causal, key-padding mask): two 128-query
// tiles per CTA ping-pong on one tensor-core stream, accumulator O kept in TMEM with lazy rescaling.
// Same contract as attention_sm100.cuh (replaces mx.fast.scaled_dot_product_attention, dit.py:166).
//
//   warp 0      TMA: Q0,Q1 once; K/V 128-key tiles in 2-stage rings
//   warp 1      MMA issuer, order  S0_0 S1_0 | PV0_j S0_{j+1} PV1_j S1_{j+1} | ...
//               S_g = Q_g K^T (M128 N128 K64) -> TMEM S_g;  O_g += P_g V (M128 N64 K128) -> TMEM O_g
//   warps 2-3   idle (complete warpgroup 0, which hands its registers to the softmax warpgroups)
//   warps 4-7   softmax group 0 (thread = query row of tile 0), 224 registers via setmaxnreg
//   warps 8-11  softmax group 1 (tile 1)
// While group 0 exponentiates S0_j the tensor cores compute S1_j / PV1_{j-1}, and vice versa.
// The running max used for exponentiation (m_used) is only advanced — and O_g/l rescaled in TMEM —
// when the true row max has grown by more than 2^8 (any thread of the warp), so the common case
// has no accumulator traffic at all; the final O/l is exact either way.
// TMEM: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384).  smem 160 KB -> one CTA per SM.
#pragma once
#include "attention2_sm100.cuh"

namespace f5 {

struct Attn3Smem {
  static constexpr int kQ = 0;                        // 2 x (128 x 64 bf16)
  static constexpr int kK = 2 * 16384;                // 2 stages
  static constexpr int kV = kK + 2 * 16384;           // 2 stages
  static constexpr int kP = kV + 2 * 16384;           // 2 groups x 32 KB
  static constexpr int kBar = kP + 2 * 32768;
  // q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2], p_full[2], pv_done[2], s_free[2]
  static constexpr int kNumBars = 17;
  static constexpr int kX = kBar + kNumBars * 8 + 16;          // float xch[2 parity][2 groups][2 halves][128]
  static constexpr int kTotal = kX + 2 * 2 * 2 * 128 * 4;
};

__global__ void __launch_bounds__(640, 1)
attn3_fwd_kernel(const __grid_constant__ CUtensorMap tma_qkv, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Attn3Smem::kBar);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2] per group
  uint64_t* p_full = bars + 11;   // [2] per group
  uint64_t* pv_done = bars + 13;  // [2] per group
  uint64_t* s_free = bars + 15;   // [2] per group: the softmax group holds S_j in registers
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + Attn3Smem::kNumBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int HD = p.H * 64;
  int kv_len = p.kv_len ? p.kv_len[b] : p.N;
  kv_len = min(max(kv_len, 1), p.N);
  const int num_kv = (kv_len + 127) >> 7;
  const bool g1_active = q0 + 128 < p.N;   // second query tile has at least one real row

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_qkv);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 256);
      mbar_init(&pv_done[i], 1);
      mbar_init(&s_free[i], 256);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(q_full, g1_active ? 32768 : 16384);
      tma_load_3d(smem + Attn3Smem::kQ, &tma_qkv, q_full, h * 64, q0, b);
      if (g1_active) tma_load_3d(smem + Attn3Smem::kQ + 16384, &tma_qkv, q_full, h * 64, q0 + 128, b);
      for (int j = 0; j < num_kv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_expect_tx(&k_full[s], 16384);
        tma_load_3d(smem + Attn3Smem::kK + s * 16384, &tma_qkv, &k_full[s], HD + h * 64, j * 128, b);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_expect_tx(&v_full[s], 16384);
        tma_load_3d(smem + Attn3Smem::kV + s * 16384, &tma_qkv, &v_full[s], 2 * HD + h * 64, j * 128, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
    const uint32_t sQ = smem_u32(smem + Attn3Smem::kQ);
    const uint32_t sPbase = smem_u32(smem + Attn3Smem::kP);
    const int ngroups = g1_active ? 2 : 1;
    auto issue_S = [&](int g, int stage) {
      const uint32_t sK = smem_u32(smem + Attn3Smem::kK + stage * 16384);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16_ss(tmem_base + g * 128, umma_desc_sw128(sQ + g * 16384 + k * 32, 16, 1024),
                    umma_desc_sw128(sK + k * 32, 16, 1024), idesc_s, k != 0);
    };
    auto issue_PV = [&](int g, int stage, bool acc) {
      const uint32_t sV = smem_u32(smem + Attn3Smem::kV + stage * 16384);
      const uint32_t sP = sPbase + g * 32768;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_f16_ss(tmem_base + 256 + g * 64,
                    umma_desc_sw128(sP + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                    umma_desc_sw128(sV + k * 2048, 16384, 1024), idesc_o, (acc || k != 0) ? 1u : 0u);
    };
    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    if (lane == 0) {
      for (int g = 0; g < ngroups; ++g) {
        issue_S(g, 0);
        tc_commit(&s_full[g]);
      }
      tc_commit(&k_empty[0]);
    }
    __syncwarp();
    for (int j = 0; j < num_kv; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const bool more = j + 1 < num_kv;
      // S_g(j+1) as soon as group g has pulled S_g(j) into registers (s_free) — it is then ready by the
      // time the group finishes exponentiating tile j, instead of being issued behind PV_g(j)
      if (more) {
        mbar_wait(&k_full[s ^ 1], ((j + 1) >> 1) & 1);
        for (int g = 0; g < ngroups; ++g) {
          mbar_wait(&s_free[g], j & 1);
          tc_fence_after();
          if (lane == 0) {
            issue_S(g, s ^ 1);
            tc_commit(&s_full[g]);
            if (g == ngroups - 1) tc_commit(&k_empty[s ^ 1]);
          }
          __syncwarp();
        }
      }
      mbar_wait(&v_full[s], ph);
      for (int g = 0; g < ngroups; ++g) {
        mbar_wait(&p_full[g], j & 1);
        tc_fence_after();
        if (lane == 0) {
          issue_PV(g, s, j > 0);
          tc_commit(&pv_done[g]);
          if (g == ngroups - 1) tc_commit(&v_empty[s]);
        }
        __syncwarp();
      }
    }
  }
  } else {
    // ===================== softmax groups =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;\n");
    const int g = (warp - 4) >> 3;            // 8 warps per query tile
    const int hh = ((warp - 4) >> 2) & 1;     // which 64-column half of the key tile / 32-column half of O
    if (g == 0 || g1_active) {
      const int lg = warp & 3;
      const int r = lg * 32 + lane;
      const uint32_t lane_addr = (uint32_t)(lg * 32) << 16;
      const uint32_t tmem_S = tmem_base + g * 128 + hh * 64 + lane_addr;
      const uint32_t tmem_O = tmem_base + 256 + g * 64 + hh * 32 + lane_addr;
      uint8_t* sP = smem + Attn3Smem::kP + g * 32768 + hh * 16384;
      float* xch = reinterpret_cast<float*>(smem + Attn3Smem::kX);   // [parity][g][hh][row]
      constexpr float kLog2e = 1.4426950408889634f;
      float m_run = -INFINITY;   // true running row max (identical in both threads of a row)
      float m_used = 0.f;        // max used for the exponentials / O / l
      float l_run = 0.f;         // this thread's share of the row sum

      for (int j = 0; j < num_kv; ++j) {
        mbar_wait(&s_full[g], j & 1);
        tc_fence_after();
        uint32_t sv[64];
        tmem_ld32(tmem_S + 0, sv);
        tmem_ld32(tmem_S + 32, sv + 32);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(&s_free[g]);
        const int kv0 = j * 128 + hh * 64;
        if (kv0 + 64 > kv_len) {
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (kv0 + i >= kv_len) sv[i] = __float_as_uint(-INFINITY);
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(sv[i]));
          mx1 = fmaxf(mx1, __uint_as_float(sv[i + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(sv[i + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(sv[i + 3]));
        }
        const float pm = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        float* xs = xch + (((j & 1) * 2 + g) * 2) * 128;
        xs[hh * 128 + r] = pm;
        asm volatile("bar.sync %0, 256;" ::"r"(2 + g) : "memory");
        m_run = fmaxf(m_run, fmaxf(pm, xs[(hh ^ 1) * 128 + r]));
        if (j > 0) mbar_wait(&pv_done[g], (j - 1) & 1);
        bool grow = (j == 0) || ((m_run - m_used) * kLog2e > 8.f);
        if (__any_sync(0xffffffffu, grow)) {
          if (j > 0) {
            tc_fence_after();
            const float sc = ex2_approx((m_used - m_run) * kLog2e);
            uint32_t ov[32];
            tmem_ld32(tmem_O, ov);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * sc);
            tmem_st32(tmem_O, ov);
            tmem_wait_st();
            l_run *= sc;
          }
          m_used = m_run;
        }
        const float mb = m_used * kLog2e;
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
        uint8_t* prow = sP + r * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {   // 8 chunks of 8 probabilities = 16 bytes
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) e[i] = ex2_approx(fmaf(__uint_as_float(sv[c * 8 + i]), kLog2e, -mb));
          l0 += e[0] + e[4]; l1 += e[1] + e[5]; l2 += e[2] + e[6]; l3 += e[3] + e[7];
          *reinterpret_cast<uint4*>(prow + ((c ^ (r & 7)) * 16)) =
              make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]),
                         pack_bf16x2(e[6], e[7]));
        }
        l_run += (l0 + l1) + (l2 + l3);
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(&p_full[g]);
      }
      // epilogue: total row sum = sum of the two halves; each thread stores its 32 columns of O / l
      pdl_launch_dependents();
      float* xs = xch + ((num_kv & 1) * 2 + g) * 2 * 128;
      xs[hh * 128 + r] = l_run;
      asm volatile("bar.sync %0, 256;" ::"r"(2 + g) : "memory");
      const float inv = 1.f / (l_run + xs[(hh ^ 1) * 128 + r]);
      mbar_wait(&pv_done[g], (num_kv - 1) & 1);
      tc_fence_after();
      const int n = q0 + g * 128 + r;
      __nv_bfloat16* o = p.out + ((size_t)b * p.N + (n < p.N ? n : 0)) * p.ldo + h * 64 + hh * 32;
      uint32_t ov[32];
      tmem_ld32(tmem_O, ov);
      tmem_wait_ld();
      if (n < p.N) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(ov[i]) * inv, __uint_as_float(ov[i + 1]) * inv);
          w.y = pack_bf16x2(__uint_as_float(ov[i + 2]) * inv, __uint_as_float(ov[i + 3]) * inv);
          w.z = pack_bf16x2(__uint_as_float(ov[i + 4]) * inv, __uint_as_float(ov[i + 5]) * inv);
          w.w = pack_bf16x2(__uint_as_float(ov[i + 6]) * inv, __uint_as_float(ov[i + 7]) * inv);
          *reinterpret_cast<uint4*>(o + i) = w;
        }
      }
      tc_fence_before();
    }
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace f5
