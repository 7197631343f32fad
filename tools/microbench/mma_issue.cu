// Microbenchmark: cost of ISSUING tcgen05.mma from one thread vs the tensor pipe's execution time.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I f5_tts_mlx_b200/csrc tools/microbench/mma_issue.cu -o mma_issue -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace f5;

// MODE 0: constant descriptors; 1: descriptors advance per MMA (k & 3) * 32 B; 2: as 1, issued by the whole
// warp converged under elect.sync; 3: as 1 with a commit after every 4 MMAs
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) issue_kernel(long long* out, int nmma) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_slot;
  const uint32_t sA = smem_u32(smem), sB = smem_u32(smem + 32768);
  constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
  if (threadIdx.x < 32) {
    long long t0 = 0, t1 = 0, t2 = 0;
    if (MODE == 2) {
      __syncwarp();
      t0 = clock64();
      for (int k = 0; k < nmma; ++k) {
        const uint64_t da = umma_desc_sw128(sA + (k & 3) * 32, 16, 1024);
        const uint64_t db = umma_desc_sw128(sB + (k & 3) * 32, 16, 1024);
        if (elect_one()) umma_f16_ss(tm, da, db, idesc, 1);
        __syncwarp();
      }
      t1 = clock64();
      if (elect_one()) tc_commit(&bar[0]);
      __syncwarp();
      mbar_wait(&bar[0], 0);
      t2 = clock64();
    } else if (threadIdx.x == 0) {
      t0 = clock64();
      for (int k = 0; k < nmma; ++k) {
        const int kk = MODE == 0 ? 0 : (k & 3);
        umma_f16_ss(tm, umma_desc_sw128(sA + kk * 32, 16, 1024), umma_desc_sw128(sB + kk * 32, 16, 1024), idesc, 1);
        if (MODE == 3 && (k & 3) == 3) tc_commit(&bar[1]);
      }
      t1 = clock64();
      tc_commit(&bar[0]);
      mbar_wait(&bar[0], 0);
      t2 = clock64();
    }
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int N, int MODE>
void run(const char* name, long long* d, int nmma) {
  cudaFuncSetAttribute(issue_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  long long h[2];
  for (int rep = 0; rep < 2; ++rep) {
    issue_kernel<N, MODE><<<1, 128, 65536>>>(d, nmma);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-44s nmma=%3d issue %6.1f clk/mma   issue+drain %6.1f clk/mma\n", name, nmma, (double)h[0] / nmma,
         (double)h[1] / nmma);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  for (int nmma : {8, 64, 256}) {
    run<128, 0>("M128 N128 K16 const desc (lane0 branch)", d, nmma);
    run<128, 1>("M128 N128 K16 varying desc (lane0 branch)", d, nmma);
    run<128, 2>("M128 N128 K16 varying desc (elect, converged)", d, nmma);
    run<128, 3>("M128 N128 K16 varying desc + commit every 4", d, nmma);
    run<64, 1>("M128 N64  K16 varying desc", d, nmma);
    run<256, 1>("M128 N256 K16 varying desc", d, nmma);
  }
  return 0;
}
