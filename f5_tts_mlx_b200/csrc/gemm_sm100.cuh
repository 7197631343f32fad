// tcgen05 + TMA GEMM for sm_100a:  D[M,N] = epilogue( A[M,K] (bf16, K-major) · W[N,K]^T (bf16) )
//
// One CTA computes one 128 x BN output tile.  Warp roles (192 or 320 threads):
//   warp 0     TMA producer   (one elected lane): A/B k-blocks -> 128B-swizzled smem ring; the first ring of
//              weight tiles is requested before the PDL wait when the caller marks W as static
//   warp 1     MMA issuer     (one elected lane): tcgen05.mma kind::f16, fp32 accumulator in TMEM
//   warps 2-5  epilogue       (128 threads = 128 TMEM lanes = 128 output rows)
//   warps 6-9  second epilogue group (one-CTA-per-SM instantiation only, see GemmEpi)
// Pipelines: smem full/empty mbarriers (TMA <-> MMA) and one tmem_full mbarrier (MMA -> epilogue).
// With kStages <= 4 (kStages*(16+BN/8) KB of smem) two CTAs are co-resident per SM, so one CTA's epilogue
// overlaps the other's main loop (TMEM: 2 x BN columns <= 512); kStages = 6 is the one-wave variant.
//
// The same kernel runs the 1-D convolutions of the path as implicit GEMMs: the A operand is a
// 3-D tensor map (channels, frames, batch) and k-block kb reads the tile shifted by
// (tap - pad) frames; TMA zero-fills the out-of-range frames, which is exactly the conv's zero
// padding (reference: nn.Conv1d(padding=k//2), dit.py:33-38).
//
// Epilogue (all fp32, fused, per reference op):
//   v = acc + bias[col]                                   Linear bias        (dit.py:136-143 ...)
//   v = act(v)            none | GELU-tanh | GELU-erf | Mish   (dit.py:94-99, convnext_v2.py:41, dit.py:36)
//   RoPE on adjacent column pairs for col < rope_cols     (rope.py:87-107, dit.py:157-158)
//   v *= q_scale for col < q_cols                         softmax scale folded into q (dit.py:166)
//   v = row valid ? v : 0                                 "x * mask" (dit.py:172-173)
//   v *= gate[batch, col]                                 AdaLN-Zero gate (dit.py:319,323)
//   v += resid[row, col]                                  residual (fp32 stream)
//   store fp32 or bf16
#pragma once
#include <type_traits>

#ifndef F5_ISSUE1
#define F5_ISSUE1 1
#endif

#include "ptx.cuh"
#include "gemm_epilogue.cuh"

namespace f5 {

template <int BN, int kStages>
struct GemmSmem {
  static constexpr int kABytes = 128 * 64 * 2;
  static constexpr int kBBytes = BN * 64 * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = kStages * kStageBytes;
  static constexpr int kColsOffset = (kBarOffset + (2 * kStages + 1) * 8 + 16 + 15) & ~15;   // bias_s[BN], gate_s[BN], aux_s[BN] (float4 reads)
  static constexpr int kTotal = kColsOffset + 3 * BN * 4 + 1024;  // + align slack
  // epilogue store staging reuses the (idle) operand ring: fp32/bf16 chunks at [0, 64 KB) (32 KB per group), the
  // bf16 copy of the fused-LN producer mode at [64 KB, 96 KB) (16 KB per group, two alternating 8 KB buffers)
  static_assert(kStages * kStageBytes >= 98304, "operand ring too small for the epilogue staging");
  static_assert(kTotal <= 232448, "single-CTA GEMM: shared memory over the 227 KB limit");
};

// Epilogue groups: 4 warps cover the 128 accumulator rows (TMEM lane quarter = warp % 4).  With 128-column
// tiles a second group of 4 warps drains columns [64,128) while the first drains [0,64): in the one-wave
// B=1 GEMMs the epilogue is a serial tail behind the main loop (in-situ timelines: 3.0-4.2 us of a 10-17 us
// kernel with one group), and its cost is latency (TMEM load -> math -> staged store), not bandwidth.
// (Only for the one-CTA-per-SM variant: with two co-resident CTAs 320 threads would leave 96 registers.)
// (r02: F5_TWO_GROUPS_2CTA=1 gives the two-CTAs-per-SM variant a second group too when its epilogue holds neither a
// RoPE table nor residual tiles — FF1 at batch 1; 2 x 320 threads leave 96 registers, which that epilogue fits without
// spilling.  Measured SLOWER on the same box, 53.0 vs 52.55 ms per step: four epilogue groups per SM compete with the
// co-resident CTA's main loop.  Off.)
#ifndef F5_TWO_GROUPS_2CTA
#define F5_TWO_GROUPS_2CTA 0
#endif
template <int BN, int kStages, bool ROPE = false, bool RESID = true>
struct GemmEpi {
  static constexpr int kGroups = (BN >= 128 && (kStages > 4 || (F5_TWO_GROUPS_2CTA && !ROPE && !RESID))) ? 2 : 1;
  static constexpr int kThreads = 64 + 128 * kGroups;
  static constexpr int kCols = BN / kGroups;          // columns per group
};

// FP8 = false instantiations have every e4m3 feature (ab8 / out_fp8 / out2_fp8 / acc_scale) folded away at compile
// time: the prologue, the issue loop and the epilogue are sensitive to every extra instruction (r02: carrying the
// run-time flags cost the bf16 path 1-3 % of the step), so only the FP8 mode pays for the FP8 mode.
// RESID = false instantiations (no residual input) drop the residual tiles from the epilogue's registers.
template <int BN, int kStages, int ACT, bool OUT_BF16, bool ROPE, bool FP8 = false, bool RESID = true>
__global__ void __launch_bounds__(GemmEpi<BN, kStages, ROPE, RESID>::kThreads, (kStages > 4) ? 1 : 2)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tma_a,
                    const __grid_constant__ CUtensorMap tma_b, const __grid_constant__ CUtensorMap tma_out,
                    const __grid_constant__ CUtensorMap tma_out2, const GemmParams p_arg) {
  GemmParams p = p_arg;
  if constexpr (!FP8) { p.ab8 = 0; p.out_fp8 = 0; p.out2_fp8 = 0; p.acc_scale = 1.f; }
  if constexpr (!RESID) p.resid = nullptr;
  using S = GemmSmem<BN, kStages>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles must sit on 1024-byte boundaries
  // (pointer arithmetic on the __shared__ array, not an integer round trip: the compiler keeps the address space and
  // emits LDS / STS instead of generic LD / ST for everything derived from `smem`)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates ----
  const int n0 = blockIdx.x * BN;
  int batch = 0, m_in_batch0 = 0, row0;
  if (p.tiles_per_batch > 0) {
    batch = blockIdx.y / p.tiles_per_batch;
    m_in_batch0 = (blockIdx.y % p.tiles_per_batch) * 128;
    row0 = batch * p.rows_per_batch + m_in_batch0;
  } else {
    row0 = blockIdx.y * 128;
    m_in_batch0 = row0;
  }
  const int kbe = p.ab8 ? 128 : 64;      // elements per k-block: always 128 bytes per row (one swizzle span)
  const int kb_per_tap = (p.k_per_tap + kbe - 1) / kbe;
  const int num_kb = p.conv_taps * kb_per_tap;

  // ---- one-time setup (overlaps the predecessor kernel under PDL) ----
  const int cta_lin = blockIdx.y * gridDim.x + blockIdx.x;
  if (threadIdx.x == 0) ts_mark(p, cta_lin, 0);
  if (warp == 0 && F5_ELECT_LANE()) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_out);
    if (p.out2 != nullptr) tma_prefetch_desc(&tma_out2);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, BN);
    tmem_relinquish();
  }
  if (warp == 2) prefetch_slice_l2(p, cta_lin, gridDim.x * gridDim.y, lane);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (threadIdx.x == 0) ts_mark(p, cta_lin, 1);
  // weights do not depend on the predecessor kernel: the first ring of B tiles is requested BEFORE the PDL
  // wait, so their (possibly HBM) latency runs under the predecessor's tail
  const int early_b = p.w_static ? min(kStages, num_kb) : 0;
  if (warp == 0 && F5_ELECT_LANE()) {
    for (int kb = 0; kb < early_b; ++kb) {
      mbar_expect_tx(&full_bar[kb], S::kStageBytes);
      tma_load_2d(smem + kb * S::kStageBytes + S::kABytes, &tma_b, &full_bar[kb], kb * kbe, n0);
    }
  }
  pdl_wait();   // predecessor's outputs (our A operand / residual) are complete and visible
  if (threadIdx.x == 64) { ts_mark(p, cta_lin, 2); prof_stamp_begin(p.prof); }   // an epilogue thread: the producer goes straight to its first load

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (F5_ELECT_LANE()) {
      auto produce = [&](auto ab8_tag) {
        constexpr int KBE = decltype(ab8_tag)::value ? 128 : 64;     // elements per k-block, compile-time in the loop
        // incremental stage / phase / tap bookkeeping: no division in the loop (see the issue loop below)
        int s = 0, tap = 0, kc = 0;
        uint32_t ph = 1;
        uint8_t* sa = smem;
        const int a_col0 = p.conv_grouped ? n0 : 0;
        const int a_row0 = m_in_batch0 - p.conv_pad;
        const int a_b = p.tiles_per_batch > 0 ? batch : 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph);
          if (kb >= early_b) mbar_expect_tx(&full_bar[s], S::kStageBytes);
          tma_load_3d(sa, &tma_a, &full_bar[s], a_col0 + kc * KBE, a_row0 + tap, a_b);
          if (kb >= early_b) tma_load_2d(sa + S::kABytes, &tma_b, &full_bar[s], kb * KBE, n0);
#ifndef F5_EPI_PROBE
          if (kb == 0) ts_mark(p, cta_lin, 3);
#endif
          if (++s == kStages) { s = 0; ph ^= 1; sa = smem; } else { sa += S::kStageBytes; }
          if (++kc == kb_per_tap) { kc = 0; ++tap; }
        }
      };
      if (p.ab8) produce(std::true_type{});
      else produce(std::false_type{});
#ifndef F5_EPI_PROBE
      ts_mark(p, cta_lin, 4);
#endif
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The issue loop is the GEMM's critical resource (r02: a run-time branch per MMA between the bf16 and the e4m3
    // instruction cost 15 % of the whole step), so the operand kind is resolved ONCE, outside the loop.
    auto mma_loop = [&](auto ab8_tag) {
      constexpr bool AB8 = decltype(ab8_tag)::value;
      constexpr uint32_t idesc = AB8 ? umma_idesc_e4m3(128, BN) : umma_idesc_bf16(128, BN, 0, 0);
#if F5_ISSUE1
      // one thread runs the whole loop (no warp-wide barrier polling / reconvergence per k-block); the two operand
      // descriptors of a stage differ from stage 0's by a constant, and the four K-steps by 32 bytes (>> 4 = 2)
      if (F5_ELECT_LANE()) {
        // stage index / phase / descriptor advance by increments (no % or / by the non-power-of-two stage count in
        // the loop: every instruction between two MMAs of this thread is time the tensor pipe may sit idle).  One
        // running 32-bit descriptor word: B's tile sits kABytes behind A's, a K-step is 32 bytes (>> 4 = 2).
        constexpr uint32_t kHi = (uint32_t)(umma_desc_sw128(0, 16, 1024) >> 32);
        constexpr uint32_t kStageInc = S::kStageBytes >> 4, kBOff = S::kABytes >> 4;
        uint32_t a_lo = (uint32_t)umma_desc_sw128(smem_u32(smem), 16, 1024);
        uint32_t bar = smem_u32(full_bar);       // full_bar[s]; empty_bar[s] is kStages * 8 bytes behind
        int s = 0;
        uint32_t ph = 0;
        if (num_kb > 0) mbar_wait_u32(bar, 0);   // first stage has landed: stamp outside the loop
        ts_mark(p, cta_lin, 5);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_u32(bar, ph);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_desc_words(a_lo + 2 * k, kHi), db = umma_desc_words(a_lo + kBOff + 2 * k, kHi);
            if constexpr (AB8) umma_f8_ss(tmem_base, da, db, idesc, (kb | k) != 0);
            else umma_f16_ss(tmem_base, da, db, idesc, (kb | k) != 0);
          }
          tc_commit_u32(bar + kStages * 8);
          if (++s == kStages) { s = 0; ph ^= 1; a_lo -= (kStages - 1) * kStageInc; bar -= (kStages - 1) * 8; }
          else { a_lo += kStageInc; bar += 8; }
        }
        tc_commit(tmem_full_bar);     // accumulator complete (commits track every MMA issued before)
        ts_mark(p, cta_lin, 6);
      }
#else
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (lane == 0) {
          if (kb == 0) ts_mark(p, cta_lin, 5);
          const uint32_t sa = smem_u32(smem + s * S::kStageBytes);
          const uint32_t sb = sa + S::kABytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 4 x UMMA_K (16 bf16 / 32 e4m3 = 32 bytes) per 128-byte k-block row
            uint64_t da = umma_desc_sw128(sa + k * 32, 16, 1024);
            uint64_t db = umma_desc_sw128(sb + k * 32, 16, 1024);
            if constexpr (AB8) umma_f8_ss(tmem_base, da, db, idesc, (kb | k) != 0);
            else umma_f16_ss(tmem_base, da, db, idesc, (kb | k) != 0);
          }
          tc_commit(&empty_bar[s]);                        // frees the smem slot when MMAs retire
          if (kb == num_kb - 1) {
            tc_commit(tmem_full_bar);  // accumulator complete
            ts_mark(p, cta_lin, 6);
          }
        }
        __syncwarp();
      }
#endif
    };
    if (p.ab8) mma_loop(std::true_type{});
    else mma_loop(std::false_type{});
  } else {
    // ===================== epilogue =====================
    using E = GemmEpi<BN, kStages, ROPE, RESID>;
    constexpr int BNG = E::kCols;
    const int grp = (warp - 2) >> 2;       // 0: columns [0, BNG), 1: [BNG, BN)
    const int lg = warp & 3;               // TMEM lane group this warp may access
    const int et = ((warp - 2) & 3) * 32 + lane;
    const int n0g = n0 + grp * BNG;
    const int r_in_tile = lg * 32 + lane;
    const int m_in_batch = m_in_batch0 + r_in_tile;
    const int row = row0 + r_in_tile;
    bool row_ok;
    int b_idx, pos;
    if (p.tiles_per_batch > 0) {
      row_ok = m_in_batch < p.rows_per_batch;
      b_idx = batch;
      pos = m_in_batch;
    } else {
      row_ok = row < p.M;
      const int rpb = p.rows_per_batch > 0 ? p.rows_per_batch : p.M;
      b_idx = row / rpb;
      pos = row - b_idx * rpb;
    }
    if (!row_ok) { b_idx = 0; pos = 0; }
    bool row_valid = true;
    if (p.row_len != nullptr) row_valid = pos < p.row_len[b_idx];

    // operand staging, overlapped with the main loop
    float* bias_s = reinterpret_cast<float*>(smem + S::kColsOffset) + grp * BNG;
    float* gate_s = reinterpret_cast<float*>(smem + S::kColsOffset) + BN + grp * BNG;
    float* aux_s = reinterpret_cast<float*>(smem + S::kColsOffset) + 2 * BN + grp * BNG;
    epi_stage_cols<BNG>(p, n0g, et, bias_s, gate_s, aux_s);
    float ln_mu_r, ln_rstd;
    epi_load_ln_row(p, row, row_ok, ln_mu_r, ln_rstd);
    float2 cs[ROPE ? 32 : 1];
    epi_load_rope<ROPE>(p, pos, cs);
    constexpr bool kPreloadAll = (kStages > 4);   // single-wave variant: one CTA per SM
    float4 res_all[kPreloadAll ? BNG / 32 : 1][8];
    if (kPreloadAll) {
#pragma unroll
      for (int c = 0; c < BNG / 32; ++c) epi_load_resid(p, row, n0g + c * 32, row_ok, res_all[c]);
    } else {
      epi_load_resid(p, row, n0g, row_ok, res_all[0]);
    }
    asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");   // bias_s / gate_s visible to the group

    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    // PDL: the main loop is done — let the next kernel of the stream start its prologue (barrier
    // init, TMEM alloc, descriptor prefetch) under this epilogue; it still waits (pdl_wait) for this
    // grid to complete before touching memory.  (Triggering at kernel start measured SLOWER.)
    pdl_launch_dependents();
    if (warp == 2 && lane == 0) ts_mark(p, cta_lin, 7);
    {
      // all MMAs have retired (tmem_full), so the operand ring is idle: its first 32 KB per group stage the stores
      EpiStage stg;
      stg.buf = smem + grp * 32768;
      stg.et = et;
      stg.r = r_in_tile;
      stg.map_out = &tma_out;
      stg.map_out2 = &tma_out2;
      stg.c1 = m_in_batch0;
      stg.c2 = p.tiles_per_batch > 0 ? batch : 0;
      stg.bar_id = 1 + grp;
      stg.probe_cta = cta_lin;
      stg.probe = p.ts != nullptr ? p.ts + (size_t)cta_lin * 10 : nullptr;
      stg.buf2 = smem + 65536 + grp * 16384;
      stg.buf2_par = 8192;
      stg.mu_r = ln_mu_r; stg.rstd = ln_rstd;
      stg.out_fp8 = p.out_fp8;
      const uint32_t tacc = tmem_base + grp * BNG + ((uint32_t)(lg * 32) << 16);
      if constexpr (kPreloadAll) {
        epi_drain_tile_preloaded<BNG, ACT, OUT_BF16, ROPE>(tacc, bias_s, gate_s, aux_s, cs, res_all, p, n0g, row, b_idx,
                                                           row_ok, row_valid, stg);
      } else {
        epi_drain_tile<BNG, ACT, OUT_BF16, ROPE>(tacc, bias_s, gate_s, aux_s, cs, res_all[0], p, n0g, row, b_idx, row_ok,
                                                 row_valid, stg, 0, 1);
      }
    }
    if (et == 0) tma_store_wait_read<0>();   // the staging buffers must outlive the TMA unit's reads; grid completion
                                             // makes the global writes visible to the dependent kernel
    tc_fence_before();
    if (warp == 2 + 4 * (GemmEpi<BN, kStages, ROPE, RESID>::kGroups - 1) && lane == 0) ts_mark(p, cta_lin, 8);
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
  if (threadIdx.x == 0) { ts_mark(p, cta_lin, 9); prof_stamp_end(p.prof); }
}

}  // namespace f5
