// micro_umma.cu -- issue rate of tcgen05.mma (kind::f16, cta_group::1, operands in shared memory, SWIZZLE_128B K-major)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../aaltoasr_b200/csrc -o micro_umma micro_umma.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace akugpu::tc;

// mode bit0: alternate two accumulators; bit1: every MMA reads a different A chunk / B chunk (else the same)
template <int N>
__global__ void __launch_bounds__(128, 1) k_umma(int n_mma, int reps, int mode, long long *out)
{
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;   // fp16 1.0
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;
  constexpr uint32_t IDESC = umma_idesc(128, N, false);
  if (warp == 1 && lane == 0) {
    const uint32_t a_base = smem_u32(smem), b_base = a_base + 32768;
    long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
      for (int i = 0; i < n_mma; i++) {
        const uint32_t off = (mode & 2) ? (uint32_t)((i & 3) * 32 + ((i >> 2) & 1) * 16384) : 0u;
        const uint32_t d = tmem_base + ((mode & 1) ? (i & 1) * N : 0);
        umma_f16(d, umma_desc(a_base + off), umma_desc(b_base + off), IDESC, i > 1 ? 1u : 0u);
      }
      umma_commit(&bar);
      mbar_wait(&bar, r & 1);
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

int main()
{
  long long *d_out, h;
  cudaMalloc(&d_out, 8);
  const size_t smem = 65 * 1024 + 1024;
  cudaFuncSetAttribute(k_umma<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_umma<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int grid : {1, 148})
    for (int mode = 0; mode < 4; mode++)
      for (int n_mma : {1, 4, 15, 60}) {
        const int reps = 200;
        k_umma<128><<<grid, 128, smem>>>(n_mma, reps, mode, d_out);
        cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
        double c128 = (double)h / reps;
        k_umma<256><<<grid, 128, smem>>>(n_mma, reps, mode, d_out);
        cudaError_t e = cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
        double c256 = (double)h / reps;
        printf("grid %3d mode %d n_mma %2d: N=128 %.0f clk/batch (%.1f per MMA)   N=256 %.0f clk/batch (%.1f per MMA) %s\n", grid, mode, n_mma,
               c128, c128 / n_mma, c256, c256 / n_mma, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
