// Microbenchmark: aggregate L2 -> shared-memory TMA bandwidth of all SMs streaming GEMM-like operand tiles from
// an L2-resident matrix, plain vs cluster multicast of the shared operand.  Decides whether the small-M GEMMs
// (bound at ~94 GB/s per SM in the r01 timelines) are limited by L2 read bandwidth (multicast helps) or by the
// SM's own ingest rate (it does not).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I f5_tts_mlx_b200/csrc tools/microbench/tma_ingest.cu -o tools/microbench/tma_ingest -lcuda
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace f5;

constexpr int kK = 1024;           // columns of both operand matrices (bf16)
constexpr int kKBlocks = kK / 64;  // 16 k-blocks of 64 columns = 128 B rows (one swizzle span)

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.aligned;\nbarrier.cluster.wait.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;\n" ::"r"(smem_u32(smem_dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  uint32_t local = smem_u32(bar), remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(cta));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

// Each CTA streams, per k-block, an A tile (128 rows x 64 cols = 16 KB) and a B tile (16 KB).
// C = cluster size along "N": the C CTAs of a cluster use the SAME A tile; with MC each loads 128/C rows of it
// and multicasts them to all peers, and loads its own B tile in full.
template <int C, bool MC, int kStages>
__global__ void __launch_bounds__(128)
ingest_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAslice,
              const __grid_constant__ CUtensorMap tmB, int iters, int a_tiles, int b_tiles, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kStages], empty[kStages];
  const uint32_t rank = C > 1 ? cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / C;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], C); }
    fence_mbar_init();
  }
  __syncthreads();
  if (C > 1) cluster_sync_all();
  const int a_tile = cluster_id % a_tiles, b_tile = blockIdx.x % b_tiles;
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    // producer
    int it = 0;
    for (int i = 0; i < iters; ++i) {
      for (int kb = 0; kb < kKBlocks; ++kb, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], 32768);
        uint8_t* dst = smem + s * 32768;
        if (MC) {
          constexpr int rows = 128 / C;
          tma_load_2d_mc(dst + rank * rows * 128, &tmAslice, &full[s], kb * 64, a_tile * 128 + rank * rows,
                         (uint16_t)((1u << C) - 1));
        } else {
          tma_load_2d(dst, &tmA, &full[s], kb * 64, a_tile * 128);
        }
        tma_load_2d(dst + 16384, &tmB, &full[s], kb * 64, b_tile * 128);
      }
    }
  } else if (threadIdx.x == 32) {
    // consumer: observe the stage, release it in every CTA of the cluster
    int it = 0;
    for (int i = 0; i < iters; ++i) {
      for (int kb = 0; kb < kKBlocks; ++kb, ++it) {
        const int s = it % kStages;
        mbar_wait(&full[s], (it / kStages) & 1);
        if (C > 1) {
          for (uint32_t c = 0; c < C; ++c) mbar_arrive_remote(&empty[s], c);
        } else {
          mbar_arrive(&empty[s]);
        }
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (C > 1) cluster_sync_all();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn get_encode() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  return (EncodeFn)fn;
}
static CUtensorMap make_map(void* p, int rows, int box_rows) {
  CUtensorMap m; cuuint64_t dims[2] = {(cuuint64_t)kK, (cuuint64_t)rows}; cuuint64_t str[1] = {(cuuint64_t)kK * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("tensormap encode failed %d\n", (int)r); exit(1); }
  return m;
}

template <int C, bool MC, int kStages = 6>
void run(const char* name, void* A, void* B, int a_tiles, int b_tiles, int grid, long long* d_out) {
  CUtensorMap tA = make_map(A, a_tiles * 128, 128), tAs = make_map(A, a_tiles * 128, 128 / C), tB = make_map(B, b_tiles * 128, 128);
  auto kern = ingest_kernel<C, MC, kStages>;
  const int smem = kStages * 32768 + 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  const int iters = 8;
  float best = 1e9f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tA, tAs, tB, iters, a_tiles, b_tiles, d_out);
    cudaEventRecord(e1);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e != cudaSuccess || e2 != cudaSuccess) { printf("%s: launch %s / sync %s\n", name, cudaGetErrorString(e), cudaGetErrorString(e2)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  const double delivered = (double)grid * iters * kKBlocks * 32768.0;               // bytes landing in smem
  const double l2read = MC ? (double)grid * iters * kKBlocks * (16384.0 / C + 16384.0) : delivered;
  printf("%-40s grid %3d: %7.1f us  delivered %6.2f TB/s (%5.1f GB/s per SM)  L2 reads %6.2f TB/s\n", name, grid,
         best * 1e3, delivered / best / 1e9, delivered / best / 1e6 / grid, l2read / best / 1e9);
}

int main() {
  const int a_tiles = 16, b_tiles = 24;                 // 16 x 128 rows of A (4 MB), 24 x 128 rows of B (6 MB): L2 resident
  void *A, *B; long long* d_out;
  cudaMalloc(&A, (size_t)a_tiles * 128 * kK * 2); cudaMalloc(&B, (size_t)b_tiles * 128 * kK * 2); cudaMalloc(&d_out, 4096 * 8);
  cudaMemset(A, 0, (size_t)a_tiles * 128 * kK * 2); cudaMemset(B, 0, (size_t)b_tiles * 128 * kK * 2);
  // ring depth sweep (bytes in flight per CTA = stages x 32 KB) and two CTAs per SM (2 x 3 stages)
  run<1, false, 2>("plain, 2 stages (64 KB in flight)", A, B, a_tiles, b_tiles, 148, d_out);
  run<1, false, 3>("plain, 3 stages (96 KB)", A, B, a_tiles, b_tiles, 148, d_out);
  run<1, false, 4>("plain, 4 stages (128 KB)", A, B, a_tiles, b_tiles, 148, d_out);
  run<1, false, 6>("plain, 6 stages (192 KB)", A, B, a_tiles, b_tiles, 148, d_out);
  run<1, false, 3>("plain, 3 stages, 2 CTAs per SM", A, B, a_tiles, b_tiles, 296, d_out);
  run<1, false, 2>("plain, 2 stages, 3 CTAs per SM", A, B, a_tiles, b_tiles, 444, d_out);
  for (int grid : {1, 120}) {
    run<1, false>("plain, no cluster", A, B, a_tiles, b_tiles, grid, d_out);
    if (grid % 2 == 0) run<2, false>("cluster 2, no multicast", A, B, a_tiles, b_tiles, grid, d_out);
    if (grid % 2 == 0) run<2, true>("cluster 2, A multicast", A, B, a_tiles, b_tiles, grid, d_out);
    if (grid % 4 == 0) run<4, true>("cluster 4, A multicast", A, B, a_tiles, b_tiles, grid, d_out);
    if (grid % 8 == 0) run<8, true>("cluster 8, A multicast", A, B, a_tiles, b_tiles, grid, d_out);
  }
  return 0;
}
