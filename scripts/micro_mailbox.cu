// micro_mailbox.cu -- host <-> persistent-kernel round trip through pinned, mapped memory: which load flavour sees a
// host write promptly, what a poll costs with 1 / 148 polling CTAs, what the echo store needs to leave the GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/micro_mailbox scripts/micro_mailbox.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <chrono>
#include <atomic>
#include <immintrin.h>

__device__ __forceinline__ unsigned ld_flavour(const unsigned *p, int mode)
{
  unsigned v;
  switch (mode) {
    case 0: asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); break;
    case 1: asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); break;
    case 2: asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); break;
    default: asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); break;
  }
  return v;
}

// every CTA polls cmd[0]; on change, `payload` threads of it fetch one float each (coalesced or not is the hardware's
// business), store `nstore` floats of results, fence, count; the last CTA echoes the sequence number.
__global__ void mailbox_kernel(const unsigned *cmd, const float *x, float *out, volatile unsigned *done, unsigned *cnt, int mode,
                               int payload, int nstore, int fence_after, unsigned n_calls)
{
  __shared__ unsigned s_seq;
  __shared__ float s_sum;
  unsigned last = 0;
  for (unsigned c = 0; c < n_calls; c++) {
    if (threadIdx.x == 0) {
      unsigned v;
      do { v = ld_flavour(cmd, mode); } while (v == last);
      s_seq = v;
      s_sum = 0.f;
    }
    __syncthreads();
    last = s_seq;
    if (last == 0xffffffffu) return;
    float v = 0.f;
    if ((int)threadIdx.x < payload) asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(x + threadIdx.x) : "memory");
    if ((int)threadIdx.x < payload) atomicAdd(&s_sum, v);
    __syncthreads();
    if ((int)threadIdx.x < nstore) out[blockIdx.x * nstore + threadIdx.x] = s_sum + (float)threadIdx.x;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      if (atomicAdd(cnt, 1u) == gridDim.x - 1) {
        *cnt = 0u;
        __threadfence_system();
        done[0] = last;
        if (fence_after) __threadfence_system();
      }
    }
  }
}

int main()
{
  unsigned char *h;
  cudaHostAlloc(&h, 1 << 20, cudaHostAllocMapped);
  unsigned char *d;
  cudaHostGetDevicePointer(&d, h, 0);
  unsigned *cnt;
  cudaMalloc(&cnt, 16);
  cudaStream_t st;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  volatile unsigned *h_done = (volatile unsigned *)h, *h_cmd = (volatile unsigned *)(h + 64);
  float *h_x = (float *)(h + 256);
  for (int i = 0; i < 1024; i++) h_x[i] = 1.f;
  const unsigned N = 2000;
  printf("mode: 0 ld.relaxed.sys 1 ld.volatile 2 ld.cv 3 ld.acquire.sys\n");
  for (int grid : {1, 148})
    for (int mode = 0; mode < 4; mode++)
      for (int cfg = 0; cfg < 4; cfg++) {
        const int payload = cfg == 0 ? 0 : 39, nstore = cfg <= 1 ? 0 : (grid == 1 ? 128 : 34), fence_after = cfg == 3 ? 0 : 1;
        memset(h, 0, 256);
        cudaMemset(cnt, 0, 16);
        cudaDeviceSynchronize();
        mailbox_kernel<<<grid, 128, 0, st>>>((const unsigned *)(d + 64), (const float *)(d + 256), (float *)(d + 8192), (volatile unsigned *)d, cnt,
                                             mode, payload, nstore, fence_after, N);
        double worst = 0;
        auto t0 = std::chrono::steady_clock::now();
        bool ok = true;
        for (unsigned c = 1; c <= N && ok; c++) {
          auto a = std::chrono::steady_clock::now();
          h_cmd[0] = c;
          uint64_t spins = 0;
          while (h_done[0] != c) {
            _mm_pause();
            if (++spins > 400000000ull) { ok = false; break; }
          }
          double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - a).count();
          if (c > 10 && us > worst) worst = us;
        }
        double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / N;
        if (!ok) { h_cmd[0] = 0xffffffffu; }
        cudaError_t e = cudaStreamSynchronize(st);
        printf("grid %3d mode %d payload %2d floats, %3d result floats per CTA, fence after flag %d: %8.2f us per round trip (worst %.1f)%s %s\n", grid, mode,
               payload, nstore, fence_after, us, worst, ok ? "" : "  TIMED OUT", e == cudaSuccess ? "" : cudaGetErrorString(e));
        fflush(stdout);
      }
  return 0;
}
