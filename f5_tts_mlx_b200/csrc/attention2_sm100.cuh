// Flash-attention forward v2 for sm_100a (head_dim 64, non-causal, key-padding mask): two 128-query
// tiles per CTA ping-pong on one tensor-core stream, accumulator O kept in TMEM with lazy rescaling.
// Non-causal flash-attention forward (replaces mx.fast.scaled_dot_product_attention + the head split / merge,
// dit.py:141-143,161-167).  The first version (one query tile per CTA, O in registers) is in the history at 0dd48cd.
//
//   warp 0      TMA: Q0,Q1 once; K/V 128-key tiles in 2-stage rings
//   warp 1, 2   MMA issuers, one per query tile g (order per tile:  S_g(0) | S_g(j+1) PV_g(j) | ...)
//               S_g = Q_g K^T (M128 N128 K64) -> TMEM S_g;  O_g += P_g V (M128 N64 K128) -> TMEM O_g
//   warp 3      idle (completes warpgroup 0, which hands its registers to the softmax warpgroups)
//   warps 4-7   softmax group 0 (thread = query row of tile 0), 224 registers via setmaxnreg
//   warps 8-11  softmax group 1 (tile 1)
// The running max used for exponentiation (m_used) is only advanced — and O_g/l rescaled in TMEM —
// when the true row max has grown by more than 2^8 (any thread of the warp), so the common case
// has no accumulator traffic at all; the final O/l is exact either way.
// TMEM: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384) P0 [384,448) P1 [448,512) (kPT).
//
// What the r01 probes (f5_debug_attention_ts, tests/gpu_checks/check_attn_timeline.py) showed, per 128-key
// tile of both groups, and what was done about it:
//   * one issuing thread for both tiles = a 3600-clk serial chain of 24 MMAs + 6 commits (short bursts cost
//     55-130 clk per MMA to issue, tools/microbench/mma_issue.cu)            -> one issuer warp per tile;
//   * 256 KB of smem traffic (2048 clk at 128 B/clk) next to the 2048-clk MUFU floor
//                                                                          -> P in TMEM (kPT), 128 KB;
//   * both groups in lockstep, XU idle during their TMEM loads / row max    -> hand-off barriers: a group
//     starts its exponentials when the other one is half way through its own.
// B2 N937 H16: 22.1 -> 18.4 us; B128: 1153 -> 990 us (cuDNN SDPA on the same box: 17.3 / 651 us).
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


struct Attn2Smem {
  static constexpr int kQ = 0;                        // 2 x (128 x 64 bf16)
  static constexpr int kK = 2 * 16384;                // 2 stages
  static constexpr int kV = kK + 2 * 16384;           // 2 stages
  static constexpr int kP = kV + 2 * 16384;           // 2 groups x 32 KB
  static constexpr int kBar = kP + 2 * 32768;
  // q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2], p_full[2], pv_done[2], s_free[2]
  static constexpr int kNumBars = 17;
  static constexpr int kTotal = kBar + kNumBars * 8 + 16;
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
      "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
      "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
      "r"(r[30]), "r"(r[31])
      : "memory");
}

// kPT: P goes to TMEM columns [384,448) / [448,512) and the PV MMA reads it from there (A operand in tensor
// memory) instead of a swizzled smem tile.  Timeline probes (r01) showed the SS variant bound by shared-memory
// bandwidth: per key tile the two groups moved 256 KB through smem (K/V fill 32, S operands 64, PV operands
// 96, P stores 64) = 2048 clk at 128 B/clk, on top of the 2048 clk MUFU floor; with P in TMEM it is 128 KB.
// kPoly: pairs per chunk of 8 exponentials evaluated on the FMA/ALU pipes (ex2_poly2) instead of MUFU.
template <bool kPT, int kPoly = 0, bool kOut8 = false>     // kOut8: e4m3 output (FP8 mode), its own instantiation
__global__ void __launch_bounds__(384, 1)
attn2_fwd_kernel(const __grid_constant__ CUtensorMap tma_qkv, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Attn2Smem::kBar);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2] per group
  uint64_t* p_full = bars + 11;   // [2] per group
  uint64_t* pv_done = bars + 13;  // [2] per group
  uint64_t* s_free = bars + 15;   // [2] per group: the softmax group holds S_j in registers
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + Attn2Smem::kNumBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 256;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int HD = p.H * 64;
  int kv_len = p.kv_len ? p.kv_len[b] : p.N;
  kv_len = min(max(kv_len, 1), p.N);
  const int num_kv = (kv_len + 127) >> 7;
  const bool g1_active = q0 + 128 < p.N;   // second query tile has at least one real row
  const bool handoff = g1_active && p.handoff;
  const bool probe = p.ts != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 &&
                     (warp == 1 || warp == 2 || warp == 4 || warp == 8);
  auto stamp = [&](int role, int j, int slot) {
    if (probe && j < 64) p.ts[(role * 64 + j) * 8 + slot] = (unsigned long long)clock64();
  };

  if (warp == 0 && F5_ELECT_LANE()) {
    tma_prefetch_desc(&tma_qkv);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], g1_active ? 2 : 1);   // released by every active issuer
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], g1_active ? 2 : 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&pv_done[i], 1);
      mbar_init(&s_free[i], 128);
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
  if (threadIdx.x == 128) prof_stamp_begin(p.prof);   // a softmax thread, not the producer

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (F5_ELECT_LANE()) {
      mbar_expect_tx(q_full, g1_active ? 32768 : 16384);
      tma_load_3d(smem + Attn2Smem::kQ, &tma_qkv, q_full, h * 64, q0, b);
      if (g1_active) tma_load_3d(smem + Attn2Smem::kQ + 16384, &tma_qkv, q_full, h * 64, q0 + 128, b);
      for (int j = 0; j < num_kv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_expect_tx(&k_full[s], 16384);
        tma_load_3d(smem + Attn2Smem::kK + s * 16384, &tma_qkv, &k_full[s], HD + h * 64, j * 128, b);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_expect_tx(&v_full[s], 16384);
        tma_load_3d(smem + Attn2Smem::kV + s * 16384, &tma_qkv, &v_full[s], 2 * HD + h * 64, j * 128, b);
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ===================== MMA issuers: one warp per query tile =====================
    // (r01 timeline: with ONE issuing thread for both tiles the 24 MMAs + 6 commits of a key tile were a
    // serial chain of ~3600 clk — short MMA bursts cost 55-85 clk each to issue, tools/microbench/mma_issue.cu —
    // and every group's P·V sat behind the other group's barriers.  Two issuers decouple the groups.)
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
    const int g = warp - 1;
    const uint32_t sQ = smem_u32(smem + Attn2Smem::kQ) + g * 16384;
    const uint32_t sP = smem_u32(smem + Attn2Smem::kP) + g * 32768;
    const uint32_t tS = tmem_base + g * 128, tO = tmem_base + 256 + g * 64, tP = tmem_base + 384 + g * 64;
    auto issue_S = [&](int stage) {
      const uint32_t sK = smem_u32(smem + Attn2Smem::kK + stage * 16384);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_f16_ss(tS, umma_desc_sw128(sQ + k * 32, 16, 1024), umma_desc_sw128(sK + k * 32, 16, 1024), idesc_s,
                    k != 0);
    };
    auto issue_PV = [&](int stage, bool acc) {
      const uint32_t sV = smem_u32(smem + Attn2Smem::kV + stage * 16384);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if constexpr (kPT)
          umma_f16_ts(tO, tP + k * 8, umma_desc_sw128(sV + k * 2048, 16384, 1024), idesc_o,
                      (acc || k != 0) ? 1u : 0u);
        else
          umma_f16_ss(tO, umma_desc_sw128(sP + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                      umma_desc_sw128(sV + k * 2048, 16384, 1024), idesc_o, (acc || k != 0) ? 1u : 0u);
      }
    };
    // (a single-thread issuer loop — no warp-wide polling / reconvergence — measured SLOWER here: 12.8 vs 12.2 ms of
    // attention per B=1 step, r02; it helped the GEMM issue loops)
    if (g == 0 || g1_active) {
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      if (F5_ELECT_LANE()) {
        issue_S(0);
        tc_commit(&s_full[g]);
        tc_commit(&k_empty[0]);
      }
      __syncwarp();
      for (int j = 0; j < num_kv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        // S_g(j+1) as soon as the group has pulled S_g(j) into registers (s_free): it is ready by the time
        // the group finishes exponentiating tile j
        if (j + 1 < num_kv) {
          mbar_wait(&k_full[s ^ 1], ((j + 1) >> 1) & 1);
          mbar_wait(&s_free[g], j & 1);
          tc_fence_after();
          if (F5_ELECT_LANE()) {
            stamp(2, j, g);
            issue_S(s ^ 1);
            tc_commit(&s_full[g]);
            tc_commit(&k_empty[s ^ 1]);
          }
          __syncwarp();
        }
        mbar_wait(&v_full[s], ph);
        mbar_wait(&p_full[g], j & 1);
        tc_fence_after();
        if (F5_ELECT_LANE()) {
          stamp(2, j, 3 + g);
          issue_PV(s, j > 0);
          tc_commit(&pv_done[g]);
          tc_commit(&v_empty[s]);
        }
        __syncwarp();
      }
    }
  }
  } else {
    // ===================== softmax groups =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;\n");
    const int g = (warp - 4) >> 2;
    if (g == 0 || g1_active) {
      const int lg = warp & 3;
      const int r = lg * 32 + lane;
      const uint32_t lane_addr = (uint32_t)(lg * 32) << 16;
      const uint32_t tmem_S = tmem_base + g * 128 + lane_addr;
      const uint32_t tmem_O = tmem_base + 256 + g * 64 + lane_addr;
      const uint32_t tmem_P = tmem_base + 384 + g * 64 + lane_addr;
      uint8_t* sP = smem + Attn2Smem::kP + g * 32768;
      constexpr float kLog2e = 1.4426950408889634f;
      float m_run = -INFINITY;   // true running max
      float m_used = 0.f;        // max used for the exponentials / O / l (lags m_run by <= 8 in log2 units)
      float l_run = 0.f;
      if (g == 1 && handoff) asm volatile("bar.arrive 2, 256;\n" ::: "memory");   // group 0 goes first

      for (int j = 0; j < num_kv; ++j) {
        stamp(g, j, 0);
        mbar_wait(&s_full[g], j & 1);
        stamp(g, j, 1);
        tc_fence_after();
        uint32_t sv[128];
        tmem_ld32(tmem_S + 0, sv);
        tmem_ld32(tmem_S + 32, sv + 32);
        tmem_ld32(tmem_S + 64, sv + 64);
        tmem_ld32(tmem_S + 96, sv + 96);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(&s_free[g]);          // S_g may be overwritten by the next tile's scores
        stamp(g, j, 2);
        const int kv0 = j * 128;
        if (kv0 + 128 > kv_len) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (kv0 + i >= kv_len) sv[i] = __float_as_uint(-INFINITY);
        }
        float mx0 = m_run, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 128; i += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(sv[i]));
          mx1 = fmaxf(mx1, __uint_as_float(sv[i + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(sv[i + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(sv[i + 3]));
        }
        m_run = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        stamp(g, j, 3);
        // P_g and O_g are free once PV_g(j-1) has completed.  (Deferring this wait into the exponential loop
        // and splitting the TMEM load around the row max both measured slower: they break the MUFU schedule.)
        if (j > 0) mbar_wait(&pv_done[g], (j - 1) & 1);
        stamp(g, j, 4);
        bool grow = (j == 0) || ((m_run - m_used) * kLog2e > 8.f);
        if (__any_sync(0xffffffffu, grow)) {
          if (j > 0) {
            // rescale O_g and l to the new reference max (warp-uniform branch: tcgen05 ops are .aligned)
            tc_fence_after();
            const float sc = ex2_approx((m_used - m_run) * kLog2e);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t ov[32];
              tmem_ld32(tmem_O + c * 32, ov);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * sc);
              tmem_st32(tmem_O + c * 32, ov);
            }
            tmem_wait_st();
            l_run *= sc;
          }
          m_used = m_run;
        }
        const float mb = m_used * kLog2e;
        // MUFU hand-off: the two groups take turns in the exponential loop (named barriers 2/3, 128 + 128
        // threads), so one group's TMEM loads / row max / barrier traffic run under the other's MUFU work
        // instead of both idling the XU pipe at the same time (they otherwise fall into lockstep behind
        // the single MMA warp)
        if (handoff) {
          if (g == 0) asm volatile("bar.sync 2, 256;\n" ::: "memory");
          else asm volatile("bar.sync 3, 256;\n" ::: "memory");
        }
        // (a packed FFMA2/FADD2 + 25 % polynomial-exp2 variant of this loop measured SLOWER on B200:
        // 24.9 us vs 21.5 us at B2 N937 H16 — kept out until the register-pair moves are understood)
        stamp(g, j, 5);
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
        uint8_t* prow = sP + r * 128;
        uint32_t pk_lo[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int c = 0; c < 16; ++c) {   // 16 chunks of 8 probabilities = 16 bytes
          float e[8];
#pragma unroll
          for (int i = 0; i < 8 - 2 * kPoly; ++i) e[i] = ex2_approx(fmaf(__uint_as_float(sv[c * 8 + i]), kLog2e, -mb));
#pragma unroll
          for (int i = 8 - 2 * kPoly; i < 8; i += 2) {
            const float2 pp = ex2_poly2(make_float2(fmaf(__uint_as_float(sv[c * 8 + i]), kLog2e, -mb),
                                                    fmaf(__uint_as_float(sv[c * 8 + i + 1]), kLog2e, -mb)));
            e[i] = pp.x;
            e[i + 1] = pp.y;
          }
          l0 += e[0] + e[4]; l1 += e[1] + e[5]; l2 += e[2] + e[6]; l3 += e[3] + e[7];
          if (c == 7 && p.handoff == 2 && handoff) {   // early release: the other group may start at half time
            if (g == 0) asm volatile("bar.arrive 3, 256;\n" ::: "memory");
            else if (j + 1 < num_kv) asm volatile("bar.arrive 2, 256;\n" ::: "memory");
          }
          if constexpr (kPT) {
            const uint32_t pk[4] = {pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]),
                                    pack_bf16x2(e[6], e[7])};
            if (c & 1) {
              const uint32_t both[8] = {pk_lo[0], pk_lo[1], pk_lo[2], pk_lo[3], pk[0], pk[1], pk[2], pk[3]};
              tmem_st8(tmem_P + (c >> 1) * 8, both);
            } else {
              pk_lo[0] = pk[0]; pk_lo[1] = pk[1]; pk_lo[2] = pk[2]; pk_lo[3] = pk[3];
            }
          } else {
            const int chunk = (c & 7) ^ (r & 7);
            *reinterpret_cast<uint4*>(prow + (c >> 3) * 16384 + chunk * 16) =
                make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]),
                           pack_bf16x2(e[6], e[7]));
          }
        }
        l_run += (l0 + l1) + (l2 + l3);
        stamp(g, j, 6);
        if (handoff && p.handoff != 2) {
          if (g == 0) asm volatile("bar.arrive 3, 256;\n" ::: "memory");
          else if (j + 1 < num_kv) asm volatile("bar.arrive 2, 256;\n" ::: "memory");
        }
        if constexpr (kPT) {
          tmem_wait_st();
          tc_fence_before();
        } else {
          tc_fence_before();
          fence_proxy_async_smem();
        }
        mbar_arrive(&p_full[g]);
        stamp(g, j, 7);
      }
      // epilogue: O / l
      pdl_launch_dependents();
      mbar_wait(&pv_done[g], (num_kv - 1) & 1);
      tc_fence_after();
      const int n = q0 + g * 128 + r;
      const float inv = 1.f / l_run;
      __nv_bfloat16* o = p.out + ((size_t)b * p.N + (n < p.N ? n : 0)) * p.ldo + h * 64;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t ov[32];
        tmem_ld32(tmem_O + c * 32, ov);
        tmem_wait_ld();
        if (kOut8 && n < p.N) {
          uint8_t* o8 = reinterpret_cast<uint8_t*>(p.out) + ((size_t)b * p.N + n) * p.ldo + h * 64 + c * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 16) {
            uint4 w;
            w.x = pack_e4m3x4(__uint_as_float(ov[i]) * inv, __uint_as_float(ov[i + 1]) * inv, __uint_as_float(ov[i + 2]) * inv, __uint_as_float(ov[i + 3]) * inv);
            w.y = pack_e4m3x4(__uint_as_float(ov[i + 4]) * inv, __uint_as_float(ov[i + 5]) * inv, __uint_as_float(ov[i + 6]) * inv, __uint_as_float(ov[i + 7]) * inv);
            w.z = pack_e4m3x4(__uint_as_float(ov[i + 8]) * inv, __uint_as_float(ov[i + 9]) * inv, __uint_as_float(ov[i + 10]) * inv, __uint_as_float(ov[i + 11]) * inv);
            w.w = pack_e4m3x4(__uint_as_float(ov[i + 12]) * inv, __uint_as_float(ov[i + 13]) * inv, __uint_as_float(ov[i + 14]) * inv, __uint_as_float(ov[i + 15]) * inv);
            *reinterpret_cast<uint4*>(o8 + i) = w;
          }
        } else if (n < p.N) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 w;
            w.x = pack_bf16x2(__uint_as_float(ov[i]) * inv, __uint_as_float(ov[i + 1]) * inv);
            w.y = pack_bf16x2(__uint_as_float(ov[i + 2]) * inv, __uint_as_float(ov[i + 3]) * inv);
            w.z = pack_bf16x2(__uint_as_float(ov[i + 4]) * inv, __uint_as_float(ov[i + 5]) * inv);
            w.w = pack_bf16x2(__uint_as_float(ov[i + 6]) * inv, __uint_as_float(ov[i + 7]) * inv);
            *reinterpret_cast<uint4*>(o + c * 32 + i) = w;
          }
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
  if (threadIdx.x == 0) prof_stamp_end(p.prof);
}

}  // namespace f5
