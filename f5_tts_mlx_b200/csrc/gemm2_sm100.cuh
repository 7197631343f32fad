// Persistent 2-CTA (cta_group::2) tcgen05 GEMM for sm_100a — the large-problem variant of
// gemm_sm100.cuh (same math, same fused epilogue, same implicit-conv mode).
//
// A cluster of two CTAs (one SM pair / TPC) computes 256 x BN output tiles with ONE
// `tcgen05.mma.cta_group::2` stream issued by the leader CTA: CTA r holds rows [128r, 128r+128) of
// A and HALF of the B tile (BN/2 weight rows) in its shared memory, the tensor cores of both SMs
// read both halves, and each CTA's TMEM receives its 128 accumulator rows.  Per k-block a CTA
// pulls 16 KB (A) + BN/2*128 B (B) into shared memory for 128 x BN x 64 MACs — twice the arithmetic
// intensity of the single-CTA 128 x 128 tile.  That is what matters: shared memory moves 128 B/clk per SM and
// carries both the TMA fill and the tensor core's operand reads (tools/microbench/); at BN = 256 the pair tile
// needs exactly that, the single-CTA tile twice as much.
//
// Persistent: grid = min(74 clusters, tiles); tiles are walked n-fastest.  Warp roles (384 threads):
//   warp 0  TMA producer (both CTAs; transaction bytes of both land on the LEADER's full barrier); the first
//           ring of weight tiles is requested before the PDL wait when the caller marks W as static
//   warp 1  MMA issuer (leader only); tcgen05.commit multicast frees the stage in both CTAs
//   warp 2  TMEM allocator (2 accumulator stages x BN columns), warp 3 L2 prefetch of the next GEMM's weights
//   warps 4-7, 8-11  two epilogue groups (both CTAs) interleaving the tile's 64-column units: they drain
//           accumulator stage i while the MMAs fill stage i^1, then arrive on the leader's tmem_empty barrier;
//           setmaxnreg hands registers from warps 0-3 (40) to the epilogue warps (232)
#pragma once
#include <type_traits>

#include "gemm_epilogue.cuh"

namespace f5 {

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-pair rank bit of a shared::cluster address

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                   smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void umma_f8_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (count 1) on the barrier at the same smem offset in every CTA of `mask` when all prior
// tcgen05.mma of this thread have completed
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;\n" ::"r"(smem_u32(bar)),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2sm_u32(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;\n" ::"r"(bar),
      "h"(mask)
      : "memory");
}
// TMA loads whose completion bytes are credited to the LEADER CTA's mbarrier
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                                int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];\n" ::"r"(smem_u32(smem_dst)),
      "l"(m), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                                int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(smem_u32(smem_dst)),
      "l"(m), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(
                   smem_u32(bar) & kPeerBitMask)
               : "memory");
}

// LNO: the instantiation supports a second (bf16) output — the fused-LN producer mode — (fp32 output, 256-wide tiles):
// one more 8 KB staging buffer per epilogue group
template <int BN, int kStages, bool LNO = false>
struct Gemm2Smem {
  static constexpr int kABytes = 128 * 64 * 2;
  static constexpr int kBBytes = (BN / 2) * 64 * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = kStages * kStageBytes;
  static constexpr int kNumBars = 2 * kStages + 4;
  static constexpr int kColVecs = 3;
  // column vectors (bias_s[BN], gate_s[BN], aux_s[BN]): one copy per epilogue group and accumulator stage — the groups
  // run tiles at their own pace, so a shared copy could be restaged by one group while the other still reads the
  // previous tile's values (compute-sanitizer racecheck, r02)
  static constexpr int kColsOffset = (kBarOffset + kNumBars * 8 + 16 + 15) & ~15;
  static constexpr int kStageOutOffset = (kColsOffset + 2 * 2 * kColVecs * BN * 4 + 1023) & ~1023;   // 2 x 16 KB store staging per group
  static constexpr int kStage2Offset = kStageOutOffset + 2 * 32768;                  // 8 KB per group (LNO)
  static constexpr int kTotal = kStage2Offset + (LNO ? 2 * 8192 : 0) + 1024;  // + align slack
  static_assert(kTotal <= 232448, "CTA-pair GEMM: shared memory over the 227 KB limit");
};

// Tile walk of one cluster: tiles t = first + i * stride, i < count, with t = m_tile * n_tiles + n_tile.
// Default: first = cluster, stride = clusters (n-fastest round robin).  "Row-sticky" walk when there are more tiles
// than clusters but no more row blocks than clusters and it costs no extra round (QKV at batch 1: 8 row blocks x 16
// column tiles on 74 clusters): G = clusters / m_tiles clusters share one row block and split its column tiles, so a
// cluster's tiles all have the SAME rows — the epilogue loads the rows' RoPE table / LayerNorm statistics once.
struct TileWalk {
  int first, stride, count;
  bool sticky;
};
__device__ __forceinline__ TileWalk tile_walk(int cluster_id, int num_clusters, int n_tiles, int total_tiles) {
  TileWalk w;
  const int m_tiles = total_tiles / n_tiles;
  const int G = m_tiles > 0 ? num_clusters / m_tiles : 0;
  w.sticky = total_tiles > num_clusters && G >= 1 &&
             (n_tiles + G - 1) / G <= (total_tiles + num_clusters - 1) / num_clusters;
  if (w.sticky) {
    const int m = cluster_id / G, j = cluster_id - m * G;
    w.first = m * n_tiles + j;
    w.stride = G;
    w.count = (m < m_tiles && j < n_tiles) ? (n_tiles - j + G - 1) / G : 0;
  } else {
    w.first = cluster_id;
    w.stride = num_clusters;
    w.count = cluster_id < total_tiles ? (total_tiles - cluster_id + num_clusters - 1) / num_clusters : 0;
  }
  return w;
}

template <int BN, bool OUT_BF16>
struct Gemm2Lno {
  static constexpr bool value = !OUT_BF16 && BN == 256;
};

template <int BN, int kStages, int ACT, bool OUT_BF16, bool ROPE, bool FP8 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
gemm2_bf16_tn_kernel(const __grid_constant__ CUtensorMap tma_a,
                     const __grid_constant__ CUtensorMap tma_b, const __grid_constant__ CUtensorMap tma_out,
                     const __grid_constant__ CUtensorMap tma_out2, const GemmParams p_arg,
                     const int n_tiles, const int total_tiles) {
  GemmParams p = p_arg;      // FP8 = false: the e4m3 features are compile-time off (see gemm_sm100.cuh)
  if constexpr (!FP8) { p.ab8 = 0; p.out_fp8 = 0; p.out2_fp8 = 0; p.acc_scale = 1.f; }
  constexpr bool LNO = Gemm2Lno<BN, OUT_BF16>::value;
  using S = Gemm2Smem<BN, kStages, LNO>;
  extern __shared__ uint8_t smem_raw[];
  // (pointer arithmetic on the __shared__ array, not an integer round trip: the compiler keeps the address space and
  // emits LDS / STS instead of generic LD / ST for everything derived from `smem`)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2] (leader's copy is the live one)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int kbe = p.ab8 ? 128 : 64;      // elements per k-block: always 128 bytes per row (one swizzle span)
  const int kb_per_tap = (p.k_per_tap + kbe - 1) / kbe;
  const int num_kb = p.conv_taps * kb_per_tap;
  const int pair_tiles_per_batch = p.tiles_per_batch;   // in units of 256-row pair tiles
  const TileWalk walk = tile_walk(cluster_id, num_clusters, n_tiles, total_tiles);

  if (threadIdx.x == 0) ts_mark(p, blockIdx.x, 0);
  if (warp == 0 && F5_ELECT_LANE()) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_out);
    if (p.out2 != nullptr) tma_prefetch_desc(&tma_out2);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 16);   // 8 epilogue warps x 2 CTAs
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_ptr_smem, BN > 128 ? 512 : 256);   // 2 accumulator stages, power of two
    tmem_relinquish_2sm();
  }
  if (warp == 3) prefetch_slice_l2(p, blockIdx.x, gridDim.x, lane);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (threadIdx.x == 0) ts_mark(p, blockIdx.x, 1);
  // first ring of B (weight) tiles of this cluster's first tile: requested before the PDL wait (gemm_sm100.cuh)
  const int early_b = (p.w_static && walk.count > 0) ? min(kStages, num_kb) : 0;
  if (warp == 0 && F5_ELECT_LANE()) {
    const int n0e = (walk.first % n_tiles) * BN;
    for (int kb = 0; kb < early_b; ++kb) {
      if (rank == 0) mbar_expect_tx(&full_bar[kb], 2 * S::kStageBytes);
      tma_load_2d_2sm(smem + kb * S::kStageBytes + S::kABytes, &tma_b, &full_bar[kb], kb * kbe,
                      n0e + (int)rank * (BN / 2));
    }
  }
  pdl_wait();
  if (threadIdx.x == 128) { ts_mark(p, blockIdx.x, 2); prof_stamp_begin(p.prof); }   // an epilogue thread, not the producer

  // register hand-over (384 threads cap every thread at 168): warps 0-3 need few, the epilogue warps hold a
  // row's RoPE table, residual and accumulator chunk.  128 x 40 + 256 x 232 = 64 512 = exactly the 384 x 168 registers
  // the CTA owns: a larger small budget (tried: 48) makes setmaxnreg.inc wait forever — the kernel hangs.
  // (setmaxnreg must sit INSIDE the role branches, or ptxas applies the small budget to everything.)
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (F5_ELECT_LANE()) {
      auto produce = [&](auto ab8_tag) {
        constexpr int KBE = decltype(ab8_tag)::value ? 128 : 64;     // elements per k-block, compile-time in the loop
        // incremental stage / phase / tap bookkeeping: no division in the k loop (see gemm_sm100.cuh)
        int kcount = 0, s = 0;
        uint32_t ph = 1;
        uint8_t* sa = smem;
        const int b_row = (int)rank * (BN / 2);
        for (int i = 0, t = walk.first; i < walk.count; ++i, t += walk.stride) {
          const int n_tile = t % n_tiles, m_tile = t / n_tiles;
          const int n0 = n_tile * BN;
          int batch = 0, m_in_batch0;
          if (pair_tiles_per_batch > 0) {
            batch = m_tile / pair_tiles_per_batch;
            m_in_batch0 = (m_tile % pair_tiles_per_batch) * 256 + (int)rank * 128;
          } else {
            m_in_batch0 = m_tile * 256 + (int)rank * 128;
          }
          const int a_col0 = p.conv_grouped ? n0 : 0;
          const int a_row0 = m_in_batch0 - p.conv_pad;
          int tap = 0, kc = 0;
          for (int kb = 0; kb < num_kb; ++kb, ++kcount) {
            mbar_wait(&empty_bar[s], ph);
            const bool early = kcount < early_b;   // B tile and expect_tx already issued before the PDL wait
            if (rank == 0 && !early) mbar_expect_tx(&full_bar[s], 2 * S::kStageBytes);
            tma_load_3d_2sm(sa, &tma_a, &full_bar[s], a_col0 + kc * KBE, a_row0 + tap, batch);
            if (!early) tma_load_2d_2sm(sa + S::kABytes, &tma_b, &full_bar[s], kb * KBE, n0 + b_row);
            if (kcount == 0) ts_mark(p, blockIdx.x, 3);
            if (++s == kStages) { s = 0; ph ^= 1; sa = smem; } else { sa += S::kStageBytes; }
            if (++kc == kb_per_tap) { kc = 0; ++tap; }
          }
        }
      };
      if (p.ab8) produce(std::true_type{});
      else produce(std::false_type{});
      ts_mark(p, blockIdx.x, 4);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      // operand kind resolved once, outside the issue loop (see gemm_sm100.cuh)
      auto mma_loop = [&](auto ab8_tag) {
        constexpr bool AB8 = decltype(ab8_tag)::value;
        constexpr uint32_t idesc = AB8 ? umma_idesc_e4m3(256, BN) : umma_idesc_bf16(256, BN, 0, 0);
#if F5_ISSUE1
        if (F5_ELECT_LANE()) {
          // one running 32-bit descriptor word and one running barrier address (gemm_sm100.cuh); this warp lives on 40
          // registers, and a spilled loop variable is a local-memory round trip between two MMAs
          constexpr uint32_t kHi = (uint32_t)(umma_desc_sw128(0, 16, 1024) >> 32);
          constexpr uint32_t kStageInc = S::kStageBytes >> 4, kBOff = S::kABytes >> 4;
          uint32_t a_lo = (uint32_t)umma_desc_sw128(smem_u32(smem), 16, 1024);
          uint32_t bar = smem_u32(full_bar);       // full_bar[s]; empty_bar[s] is kStages * 8 bytes behind
          const uint32_t acc_bar = smem_u32(tmem_full_bar);   // tmem_full_bar[as]; tmem_empty_bar[as] 16 bytes behind
          int s = 0;
          uint32_t ph = 0;
          if (walk.count > 0 && num_kb > 0) {      // first stage has landed: stamp outside the loops
            mbar_wait_u32(bar, 0);
            ts_mark(p, blockIdx.x, 5);
          }
          for (int i = 0; i < walk.count; ++i) {
            const uint32_t as = i & 1;
            mbar_wait_u32(acc_bar + 16 + as * 8, ((i >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + as * BN;
            for (int kb = 0; kb < num_kb; ++kb) {
              mbar_wait_u32(bar, ph);
              tc_fence_after();
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da = umma_desc_words(a_lo + 2 * k, kHi);
                const uint64_t db = umma_desc_words(a_lo + kBOff + 2 * k, kHi);
                if constexpr (AB8) umma_f8_ss_2sm(tmem_acc, da, db, idesc, (kb | k) != 0);
                else umma_f16_ss_2sm(tmem_acc, da, db, idesc, (kb | k) != 0);
              }
              tc_commit_2sm_u32(bar + kStages * 8, 3);
              if (++s == kStages) { s = 0; ph ^= 1; a_lo -= (kStages - 1) * kStageInc; bar -= (kStages - 1) * 8; }
              else { a_lo += kStageInc; bar += 8; }
            }
            tc_commit_2sm_u32(acc_bar + as * 8, 3);
          }
        }
#else
        int kcount = 0, acount = 0;
        for (int i = 0; i < walk.count; ++i, ++acount) {
          const int as = acount & 1;
          const uint32_t aph = (acount >> 1) & 1;
          mbar_wait(&tmem_empty_bar[as], aph ^ 1);
          tc_fence_after();
          const uint32_t tmem_acc = tmem_base + as * BN;
          for (int kb = 0; kb < num_kb; ++kb, ++kcount) {
            const int s = kcount % kStages;
            const uint32_t ph = (kcount / kStages) & 1;
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            if (lane == 0) {
              if (kcount == 0) ts_mark(p, blockIdx.x, 5);
              const uint32_t sa = smem_u32(smem + s * S::kStageBytes);
              const uint32_t sb = sa + S::kABytes;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                uint64_t da = umma_desc_sw128(sa + k * 32, 16, 1024);
                uint64_t db = umma_desc_sw128(sb + k * 32, 16, 1024);
                if constexpr (AB8) umma_f8_ss_2sm(tmem_acc, da, db, idesc, (kb | k) != 0);
                else umma_f16_ss_2sm(tmem_acc, da, db, idesc, (kb | k) != 0);
              }
              tc_commit_2sm(&empty_bar[s], 3);
              if (kb == num_kb - 1) tc_commit_2sm(&tmem_full_bar[as], 3);
            }
            __syncwarp();
          }
        }
#endif
      };
      if (p.ab8) mma_loop(std::true_type{});
      else mma_loop(std::false_type{});
      if (F5_ELECT_LANE()) ts_mark(p, blockIdx.x, 6);
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
    // ===================== epilogue (both CTAs) =====================
    // two groups of 4 warps interleave the tile's 64-column units (see gemm_sm100.cuh: the epilogue is a
    // latency chain; with K = 1024 one group needed longer per tile than the main loop)
    const int grp = (warp - 4) >> 2;
    const int et = ((warp - 4) & 3) * 32 + lane;
    const int lg = warp & 3;
    const int r_in_tile = (int)rank * 128 + lg * 32 + lane;   // row inside the 256-row pair tile
    int acount = 0;
    constexpr bool kHalves = ((BN / 64) % 2 == 1);   // odd unit count: the groups split every unit instead (gemm_epilogue.cuh)
    float2 cs[ROPE ? 32 : 1];
    float ln_mu_r = 0.f, ln_rstd = 1.f;
    for (int i = 0, t = walk.first; i < walk.count; ++i, t += walk.stride, ++acount) {
      const int n_tile = t % n_tiles, m_tile = t / n_tiles;
      const int n0 = n_tile * BN;
      const int as = acount & 1;
      const uint32_t aph = (acount >> 1) & 1;
      int row, pos, b_idx;
      bool row_ok;
      const int b_idx_tile = pair_tiles_per_batch > 0 ? m_tile / pair_tiles_per_batch : 0;
      if (pair_tiles_per_batch > 0) {
        b_idx = m_tile / pair_tiles_per_batch;
        pos = (m_tile % pair_tiles_per_batch) * 256 + r_in_tile;
        row_ok = pos < p.rows_per_batch;
        row = b_idx * p.rows_per_batch + pos;
      } else {
        row = m_tile * 256 + r_in_tile;
        row_ok = row < p.M;
        const int rpb = p.rows_per_batch > 0 ? p.rows_per_batch : p.M;
        b_idx = row / rpb;
        pos = row - b_idx * rpb;
      }
      if (!row_ok) { b_idx = 0; pos = 0; }
      bool row_valid = true;
      if (p.row_len != nullptr) row_valid = pos < p.row_len[b_idx];

      // operand staging for this tile (overlaps the MMAs still filling the accumulator)
      float* bias_s = reinterpret_cast<float*>(smem + S::kColsOffset) + (grp * 2 + as) * S::kColVecs * BN;
      float* gate_s = bias_s + BN;
      float* aux_s = gate_s + BN;
      epi_stage_cols<BN>(p, n0, et, bias_s, gate_s, aux_s);   // this group's own copy
      float4 res0[8];
      if (!walk.sticky || i == 0) {     // row-sticky walk: every tile of this cluster has the same rows
        epi_load_ln_row(p, row, row_ok, ln_mu_r, ln_rstd);
        if constexpr (kHalves) epi_load_rope_half<ROPE>(p, pos, cs, grp);
        else epi_load_rope<ROPE>(p, pos, cs);
      }
      if constexpr (!kHalves) epi_load_resid(p, row, n0 + grp * 64, row_ok, res0);
      asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");

      mbar_wait(&tmem_full_bar[as], aph);
      tc_fence_after();
      if (i + 1 == walk.count) pdl_launch_dependents();   // last tile of this CTA: see gemm_sm100.cuh
      if (warp == 4 && lane == 0 && acount == 0) ts_mark(p, blockIdx.x, 7);
      EpiStage stg;
      stg.buf = smem + S::kStageOutOffset + grp * 32768;
      stg.et = et;
      stg.bar_id = 1 + grp;
      stg.probe_cta = blockIdx.x;
      stg.probe = p.ts != nullptr ? p.ts + (size_t)blockIdx.x * 10 : nullptr;
      stg.r = lg * 32 + lane;
      stg.buf2 = LNO ? smem + S::kStage2Offset + grp * 8192 : nullptr;
      stg.buf2_par = 0;
      stg.mu_r = ln_mu_r; stg.rstd = ln_rstd;
      stg.out_fp8 = p.out_fp8;
      stg.map_out = &tma_out;
      stg.map_out2 = &tma_out2;
      if (pair_tiles_per_batch > 0) {
        stg.c1 = (m_tile % pair_tiles_per_batch) * 256 + (int)rank * 128;
        stg.c2 = b_idx_tile;
      } else {
        stg.c1 = m_tile * 256 + (int)rank * 128;
        stg.c2 = 0;
      }
      const uint32_t tacc = tmem_base + as * BN + ((uint32_t)(lg * 32) << 16);
      if constexpr (kHalves) {
        if (grp == 0)
          epi_drain_tile_half<BN, ACT, OUT_BF16, ROPE, 0>(tacc, bias_s, gate_s, aux_s, cs, p, n0, row, b_idx, row_ok, row_valid, stg);
        else
          epi_drain_tile_half<BN, ACT, OUT_BF16, ROPE, 1>(tacc, bias_s, gate_s, aux_s, cs, p, n0, row, b_idx, row_ok, row_valid, stg);
      } else {
        epi_drain_tile<BN, ACT, OUT_BF16, ROPE>(tacc, bias_s, gate_s, aux_s, cs, res0, p, n0, row, b_idx, row_ok, row_valid,
                                                stg, grp, 2);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[as]);
    }
    if (et == 0) tma_store_wait_read<0>();   // the staging buffers must outlive the TMA unit's reads; grid completion
                                             // makes the global writes visible to the dependent kernel
    if (warp == 4 && lane == 0) ts_mark(p, blockIdx.x, 8);
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, BN > 128 ? 512 : 256);
  }
  if (threadIdx.x == 0) { ts_mark(p, blockIdx.x, 9); prof_stamp_end(p.prof); }
}

}  // namespace f5
