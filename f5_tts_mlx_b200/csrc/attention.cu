// Host launcher + C-ABI entry for the tcgen05 flash-attention forward (attention_sm100.cuh).
#include <stdlib.h>

#include "attention_sm100.cuh"
#include "attention2_sm100.cuh"
#include "attention3_sm100.cuh"
#include "host_common.h"

static unsigned long long* g_attn_ts = nullptr;
extern "C" int f5_debug_attention_ts(void* base) {
  g_attn_ts = reinterpret_cast<unsigned long long*>(base);
  return 0;
}

extern "C" int f5_attention_fwd(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out,
                                int32_t batch, int32_t frames, int32_t heads, int32_t head_dim,
                                const int32_t* kv_len, void* stream_) {
  using namespace f5;
  if (int e = device_check()) return e;
  F5_REQUIRE(qkv && out, "f5_attention_fwd: null pointer");
  F5_REQUIRE(head_dim == 64, "f5_attention_fwd: head_dim %d unsupported (only 64)", head_dim);
  F5_REQUIRE(batch > 0 && frames > 0 && heads > 0, "f5_attention_fwd: bad shape");
  F5_REQUIRE(ld_qkv % 8 == 0 && ld_qkv >= 3 * heads * 64, "f5_attention_fwd: bad ld_qkv");
  F5_REQUIRE(ld_out % 8 == 0 && ld_out >= heads * 64, "f5_attention_fwd: bad ld_out");
  CUtensorMap tm;
  uint64_t dims[3] = {(uint64_t)3 * heads * 64, (uint64_t)frames, (uint64_t)batch};
  uint64_t str[2] = {(uint64_t)ld_qkv * 2, (uint64_t)ld_qkv * 2 * (uint64_t)frames};
  uint32_t box[3] = {64, 128, 1};
  if (int e = make_tmap_bf16(&tm, qkv, 3, dims, str, box)) return e;
  AttnParams p;
  p.B = batch; p.N = frames; p.H = heads;
  p.kv_len = kv_len;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ldo = (int)ld_out;
  p.handoff = 1;
  static bool attr_set = false;
  static int poly = 0;      // pairs per 8 exponentials moved to the FMA pipe (F5_ATTN_POLY=0|1|2)
  static int handoff = 2;   // 0 off, 1 strict alternation of the exponential loops, 2 release at half time
  static int variant = 4;   // 2: two query tiles per CTA, O in TMEM (attention2_sm100.cuh); 1: v1
  if (!attr_set) {
    F5_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       AttnSmem::kTotal));
    F5_CHECK_CUDA(cudaFuncSetAttribute(attn2_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Attn2Smem::kTotal));
    F5_CHECK_CUDA(cudaFuncSetAttribute(attn2_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Attn2Smem::kTotal));
    F5_CHECK_CUDA((cudaFuncSetAttribute(attn2_fwd_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Attn2Smem::kTotal)));
    F5_CHECK_CUDA((cudaFuncSetAttribute(attn2_fwd_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Attn2Smem::kTotal)));
    const char* po = getenv("F5_ATTN_POLY");
    if (po) poly = atoi(po);
    F5_CHECK_CUDA(cudaFuncSetAttribute(attn3_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Attn3Smem::kTotal));
    const char* v = getenv("F5_ATTN_VARIANT");
    if (v && v[0] == '1') variant = 1;
    if (v && v[0] == '2') variant = 2;
    if (v && v[0] == '3') variant = 3;
    if (v && v[0] == '4') variant = 4;
    const char* ho = getenv("F5_ATTN_HANDOFF");
    if (ho && ho[0] == '0') handoff = 0;
    if (ho && ho[0] == '1') handoff = 1;
    attr_set = true;
  }
  p.handoff = handoff;
  p.ts = g_attn_ts;
  if (variant == 3) {
    dim3 grid3(cdiv(frames, 256), heads, batch);
    ProfScope ps3(PROF_ATTN, 4.0 * batch * heads * (double)frames * frames * 64.0,
                  2.0 * batch * (double)frames * heads * 64.0 * 4.0, reinterpret_cast<cudaStream_t>(stream_));
    F5_CHECK_CUDA(launch_kernel(attn3_fwd_kernel, grid3, dim3(640), Attn3Smem::kTotal,
                                reinterpret_cast<cudaStream_t>(stream_), tm, p));
    return 0;
  }
  if (variant == 2 || variant == 4) {
    dim3 grid2(cdiv(frames, 256), heads, batch);
    ProfScope ps2(PROF_ATTN, 4.0 * batch * heads * (double)frames * frames * 64.0,
                  2.0 * batch * (double)frames * heads * 64.0 * 4.0, reinterpret_cast<cudaStream_t>(stream_));
    if (variant == 4 && poly == 1)
      F5_CHECK_CUDA((launch_kernel(attn2_fwd_kernel<true, 1>, dim3(grid2), dim3(384), Attn2Smem::kTotal,
                                   reinterpret_cast<cudaStream_t>(stream_), tm, p)));
    else if (variant == 4 && poly == 2)
      F5_CHECK_CUDA((launch_kernel(attn2_fwd_kernel<true, 2>, dim3(grid2), dim3(384), Attn2Smem::kTotal,
                                   reinterpret_cast<cudaStream_t>(stream_), tm, p)));
    else if (variant == 4)
      F5_CHECK_CUDA(launch_kernel(attn2_fwd_kernel<true>, dim3(grid2), dim3(384), Attn2Smem::kTotal,
                                  reinterpret_cast<cudaStream_t>(stream_), tm, p));
    else
      F5_CHECK_CUDA(launch_kernel(attn2_fwd_kernel<false>, dim3(grid2), dim3(384), Attn2Smem::kTotal,
                                  reinterpret_cast<cudaStream_t>(stream_), tm, p));
    F5_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  dim3 grid(cdiv(frames, 128), heads, batch);
  ProfScope ps(PROF_ATTN, 4.0 * batch * heads * (double)frames * frames * 64.0,
               2.0 * batch * (double)frames * heads * 64.0 * 4.0, reinterpret_cast<cudaStream_t>(stream_));
  F5_CHECK_CUDA(launch_kernel(attn_fwd_kernel, dim3(grid), dim3(192), AttnSmem::kTotal, reinterpret_cast<cudaStream_t>(stream_), tm, p));
  F5_CHECK_CUDA(cudaGetLastError());
  return 0;
}
