// development probe: cycles per tcgen05.mma (M = 128, K = 16, fp16, SS mode) by operand majorness and N.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I singlecellhaystack_b200/csrc tools/mma_rate.cu -o /tmp/mma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "hs_ptx.cuh"
using namespace hs_ptx;

__device__ uint64_t desc_mn(uint32_t a) {   // MN-major SW128 (hs_tiles.cu)
  return (uint64_t)((a >> 4) & 0x3FFFu) | ((uint64_t)(1024 >> 4) << 16) | ((uint64_t)(2048 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ uint64_t desc_k(uint32_t a) {    // K-major SW128: 8-row groups 1024 B apart
  return (uint64_t)((a >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(128, 1) k_rate(int a_k, int b_k, int npad, int reps, int ctas_active, long long* out) {
  extern __shared__ __align__(1024) unsigned char raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = smem_u32(smem);
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;   // fp16 1.0
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0 && (int)blockIdx.x < ctas_active) {
    const uint32_t idesc = (1u << 4) | ((a_k ? 0u : 1u) << 15) | ((b_k ? 0u : 1u) << 16) | ((uint32_t)(npad >> 3) << 17) | (8u << 24);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int k16 = 0; k16 < 4; ++k16) {
        // MN-major: 16 cells further = 2 * SBO = 4096 B; K-major: 32 B inside the 128 B row
        const uint32_t ao = a_k ? (uint32_t)k16 * 32u : (uint32_t)k16 * 4096u;
        const uint32_t bo = b_k ? (uint32_t)k16 * 32u : (uint32_t)k16 * 4096u;
        const uint64_t da = a_k ? desc_k(base + ao) : desc_mn(base + ao);
        const uint64_t db = b_k ? desc_k(base + 32768 + bo) : desc_mn(base + 32768 + bo);
        umma_f16(tm, da, db, idesc, 1u);
        umma_f16(tm + 128, da, db, idesc, 1u);
        umma_f16(tm + 128, da, db, idesc, 1u);
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 64);
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
  const int reps = 64;
  for (int ctas = 1; ctas <= 148; ctas += 147)
    for (int npad : {112, 128})
      for (int ak = 0; ak < 2; ++ak)
        for (int bk = 0; bk < 2; ++bk) {
          out[0] = 0;
          k_rate<<<148, 128, 70 * 1024>>>(ak, bk, npad, reps, ctas, out);
          cudaError_t e = cudaDeviceSynchronize();
          printf("ctas %3d N %3d A %s B %s : %7.1f clk per MMA  (%s)\n", ctas, npad, ak ? "K " : "MN", bk ? "K " : "MN",
                 (double)out[0] / (reps * 12.0), cudaGetErrorString(e));
        }
  return 0;
}
