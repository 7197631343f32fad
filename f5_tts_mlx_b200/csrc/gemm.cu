// Host launcher + C-ABI entry for the tcgen05 GEMM (see gemm_sm100.cuh and include/f5_b200.h).
#include "gemm_sm100.cuh"
#include "host_common.h"

namespace f5 {

template <int BN, int kStages, int ACT, bool OUT_BF16, bool ROPE>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                       dim3 grid, cudaStream_t stream) {
  using S = GemmSmem<BN, kStages>;
  auto kern = gemm_bf16_tn_kernel<BN, kStages, ACT, OUT_BF16, ROPE>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    F5_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       S::kTotal));
    attr_set = true;
  }
  const double taps = p.conv_taps;
  ProfScope ps(PROF_GEMM, 2.0 * p.M * (double)p.N * (double)p.k_per_tap * taps,
               2.0 * ((double)p.M * p.k_per_tap + (double)p.N * p.k_per_tap * taps) +
                   (double)p.M * p.N * (OUT_BF16 ? 2.0 : 4.0),
               stream);
  kern<<<grid, 192, S::kTotal, stream>>>(ta, tb, p);
  F5_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int BN, int kStages>
static int dispatch_epi(int act, bool out_bf16, bool rope, const CUtensorMap& ta,
                        const CUtensorMap& tb, const GemmParams& p, dim3 grid,
                        cudaStream_t stream) {
#define F5_CASE(A, O, R) \
  if (act == A && out_bf16 == O && rope == R) return launch_gemm<BN, kStages, A, O, R>(ta, tb, p, grid, stream);
  F5_CASE(ACT_NONE, true, true)
  F5_CASE(ACT_NONE, true, false)
  F5_CASE(ACT_NONE, false, false)
  F5_CASE(ACT_GELU_TANH, true, false)
  F5_CASE(ACT_GELU_ERF, true, false)
  F5_CASE(ACT_MISH, true, false)
  F5_CASE(ACT_MISH, false, false)
#undef F5_CASE
  return set_error(F5_ERR_INVALID, "f5_gemm_bf16: unsupported epilogue act=%d out_bf16=%d rope=%d",
                   act, (int)out_bf16, (int)rope);
}

}  // namespace f5

extern "C" int f5_gemm_bf16(const f5_gemm_args* a, void* stream_) {
  using namespace f5;
  if (int e = device_check()) return e;
  F5_REQUIRE(a != nullptr, "f5_gemm_bf16: null args");
  F5_REQUIRE(a->a && a->w && a->out, "f5_gemm_bf16: null operand pointer");
  F5_REQUIRE(a->m > 0 && a->n > 0 && a->k > 0, "f5_gemm_bf16: bad shape m=%d n=%d k=%d", a->m,
             a->n, a->k);
  F5_REQUIRE(a->lda % 8 == 0 && a->ldw % 8 == 0, "f5_gemm_bf16: lda/ldw must be multiples of 8");
  F5_REQUIRE(a->n % (a->out_bf16 ? 8 : 4) == 0, "f5_gemm_bf16: n=%d not vector aligned", a->n);
  F5_REQUIRE(a->ldo % (a->out_bf16 ? 8 : 4) == 0, "f5_gemm_bf16: ldo not vector aligned");
  const int taps = a->conv_taps > 0 ? a->conv_taps : 1;
  const int nb = a->num_batches > 0 ? a->num_batches : 1;
  const int rpb = a->rows_per_batch > 0 ? a->rows_per_batch : a->m;
  const bool batched = a->batched_tiles != 0;
  F5_REQUIRE(taps == 1 || batched, "f5_gemm_bf16: conv mode requires batched_tiles");
  F5_REQUIRE(!a->conv_grouped || a->k == 64, "f5_gemm_bf16: grouped conv needs k == 64");
  F5_REQUIRE((int64_t)nb * rpb == a->m, "f5_gemm_bf16: m=%d != num_batches*rows_per_batch=%d*%d",
             a->m, nb, rpb);
  if (a->rope) {
    F5_REQUIRE(a->rope_cols % 64 == 0 && a->q_cols % 32 == 0, "f5_gemm_bf16: rope_cols/q_cols");
    F5_REQUIRE(a->act == F5_ACT_NONE && a->out_bf16, "f5_gemm_bf16: rope epilogue is bf16/no-act");
  }
  if (a->resid) F5_REQUIRE(a->ldr % 4 == 0, "f5_gemm_bf16: ldr not multiple of 4");
  if (a->gate) F5_REQUIRE(a->gate_ld % 4 == 0, "f5_gemm_bf16: gate_ld not multiple of 4");

  int bn = a->tile_n;
  if (a->conv_grouped) bn = 64;
  if (bn == 0) {
    // fill the 148 SMs: prefer 128-wide tiles unless that leaves most SMs idle
    const int mt = batched ? nb * cdiv(rpb, 128) : cdiv(a->m, 128);
    bn = (mt * cdiv(a->n, 128) >= 120 || a->n <= 64) ? 128 : 64;
    if (a->n <= 64) bn = 64;
  }
  F5_REQUIRE(bn == 64 || bn == 128, "f5_gemm_bf16: tile_n must be 64 or 128");

  GemmParams p;
  p.M = a->m; p.N = a->n; p.K = a->k;
  p.rows_per_batch = rpb;
  p.tiles_per_batch = batched ? cdiv(rpb, 128) : 0;
  p.num_batches = nb;
  p.conv_taps = taps;
  p.conv_pad = a->conv_pad;
  p.k_per_tap = a->k;
  p.conv_grouped = a->conv_grouped;
  p.bias = a->bias;
  p.out = a->out; p.ldo = (int)a->ldo;
  p.resid = a->resid; p.ldr = (int)a->ldr;
  p.gate = a->gate; p.gate_ld = (int)a->gate_ld;
  p.row_len = a->row_len;
  p.rope = reinterpret_cast<const float2*>(a->rope);
  p.rope_cols = a->rope_cols;
  p.q_scale = a->q_scale;
  p.q_cols = a->q_cols;
  p.out2 = reinterpret_cast<__nv_bfloat16*>(a->out2_bf16);
  p.ldo2 = (int)a->ldo2;
  if (a->out2_bf16) F5_REQUIRE(a->ldo2 % 8 == 0 && a->n % 8 == 0, "f5_gemm_bf16: out2 alignment");

  // A: (channels, frames, utterances); flat mode is one "utterance" of m rows
  CUtensorMap ta, tb;
  {
    const int kcols = a->conv_grouped ? a->n : a->k;  // grouped: channel axis spans all groups
    uint64_t dims[3] = {(uint64_t)kcols, (uint64_t)(batched ? rpb : a->m),
                        (uint64_t)(batched ? nb : 1)};
    uint64_t str[2] = {(uint64_t)a->lda * 2, (uint64_t)a->lda * 2 * (uint64_t)rpb};
    uint32_t box[3] = {64, 128, 1};
    if (int e = make_tmap_bf16(&ta, a->a, 3, dims, str, box)) return e;
  }
  {
    const int kpad = cdiv(a->k, 64) * 64;
    uint64_t dims[2] = {(uint64_t)(taps == 1 ? a->k : taps * kpad), (uint64_t)a->n};
    uint64_t str[1] = {(uint64_t)a->ldw * 2};
    uint32_t box[2] = {64, (uint32_t)bn};
    if (int e = make_tmap_bf16(&tb, a->w, 2, dims, str, box)) return e;
  }
  dim3 grid(cdiv(a->n, bn), batched ? nb * cdiv(rpb, 128) : cdiv(a->m, 128), 1);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const bool rope = a->rope != nullptr;
  if (bn == 128) return dispatch_epi<128, 3>(a->act, a->out_bf16 != 0, rope, ta, tb, p, grid, stream);
  return dispatch_epi<64, 4>(a->act, a->out_bf16 != 0, rope, ta, tb, p, grid, stream);
}
