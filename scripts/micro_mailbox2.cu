// micro_mailbox2.cu -- second round: does every CTA polling ITS OWN host line scale (148 CTAs), with the payload in
// self-validating 32-byte sectors of the same packet; and which completion is cheapest:
//   0  results to host from every CTA, fence.sys per CTA, count, last CTA publishes
//   1  results to host from every CTA, fence.gpu per CTA, count, ONE fence.sys by the last CTA, publish
//   2  results to device memory, fence.gpu, count, CTA 0 carries all rows to the host, fence.sys, publish
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <chrono>
#include <immintrin.h>

constexpr int PKT = 256;   // bytes of a CTA's packet: 8 sectors of {7 floats, tag}

__global__ void mailbox2_kernel(const unsigned *pk, float *out, float *gres, volatile unsigned *done, unsigned *cnt, int own_line,
                                int completion, int nres, unsigned n_calls, unsigned *errs)
{
  __shared__ unsigned s_seq;
  __shared__ float s_x[64];
  const unsigned *mine = pk + (own_line ? (size_t)blockIdx.x * (PKT / 4) : 0);
  unsigned last = 0;
  for (unsigned c = 0; c < n_calls; c++) {
    if (threadIdx.x < 32) {
      // lanes 0..15: 16 bytes each of the first 6+ sectors (lane 2k, 2k+1 = sector k; the tag is the last word of the odd lane)
      uint4 v = make_uint4(0, 0, 0, 0);
      unsigned ok = 0;
      const int lane = threadIdx.x;
      for (;;) {
        if (lane < 12) asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(mine + lane * 4) : "memory");
        const unsigned tag = __shfl_sync(0xffffffffu, v.w, lane | 1);
        const unsigned t0 = __shfl_sync(0xffffffffu, tag, 1);
        ok = __all_sync(0xffffffffu, lane >= 12 || (tag == t0 && tag != last));
        if (ok) { if (lane == 0) s_seq = t0; break; }
      }
      if (lane < 12) { s_x[lane * 4] = __uint_as_float(v.x); s_x[lane * 4 + 1] = __uint_as_float(v.y); s_x[lane * 4 + 2] = __uint_as_float(v.z); }
      if (lane < 12) {   // payload word j of a sector must be (float)(tag + j)
        const int j0 = (lane & 1) ? 4 : 0;
        const float want = (float)(s_seq + j0);
        if (__uint_as_float(v.x) != want || __uint_as_float(v.y) != want + 1.f || __uint_as_float(v.z) != want + 2.f) atomicAdd(errs, 1u);
      }
    }
    __syncthreads();
    last = s_seq;
    if (last == 0xffffffffu) return;
    const float r = s_x[threadIdx.x & 31] + (float)threadIdx.x;
    float *dst = completion == 2 ? gres : out;
    if ((int)threadIdx.x < nres) dst[blockIdx.x * nres + threadIdx.x] = r;
    __syncthreads();
    if (completion == 0) {
      if (threadIdx.x == 0) {
        __threadfence_system();
        if (atomicAdd(cnt, 1u) == gridDim.x - 1) { *cnt = 0u; __threadfence_system(); done[0] = last; __threadfence_system(); }
      }
    } else if (completion == 1) {
      if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(cnt, 1u) == gridDim.x - 1) { *cnt = 0u; __threadfence_system(); done[0] = last; __threadfence_system(); }
      }
    } else {
      if (threadIdx.x == 0) { __threadfence(); atomicAdd(cnt, 1u); }
      if (blockIdx.x == 0) {
        if (threadIdx.x == 0) { unsigned v; do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory"); } while (v != gridDim.x); }
        __syncthreads();
        const int total = gridDim.x * nres;
        for (int i = threadIdx.x; i < total / 4; i += blockDim.x) reinterpret_cast<uint4 *>(out)[i] = __ldcg(reinterpret_cast<const uint4 *>(gres) + i);
        __syncthreads();
        if (threadIdx.x == 0) { *cnt = 0u; __threadfence_system(); done[0] = last; __threadfence_system(); }
      }
    }
  }
}

int main()
{
  unsigned char *h;
  cudaHostAlloc(&h, 4 << 20, cudaHostAllocMapped);
  unsigned char *d;
  cudaHostGetDevicePointer(&d, h, 0);
  unsigned *cnt;
  float *gres;
  cudaMalloc(&cnt, 16);
  unsigned *errs;
  cudaMalloc(&errs, 16);
  cudaMalloc(&gres, 1 << 20);
  cudaStream_t st;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  volatile unsigned *h_done = (volatile unsigned *)h;
  unsigned *h_pk = (unsigned *)(h + 4096);
  const unsigned N = 3000;
  for (int grid : {1, 16, 148})
    for (int own = 0; own < 2; own++)
      for (int wmode = 0; wmode < 3; wmode++) {
        const int completion = 1;
        if (!own) continue;
        const int nres = 36;                     // 148 x 36 floats ~ 21 KB of results
        memset(h, 0, 4096 + 148 * PKT);
        cudaMemset(cnt, 0, 16);
        cudaMemset(errs, 0, 16);
        cudaDeviceSynchronize();
        mailbox2_kernel<<<grid, 128, 0, st>>>((const unsigned *)(d + 4096), (float *)(d + (1 << 20)), gres, (volatile unsigned *)d, cnt, own, completion, nres, N, errs);
        bool ok = true;
        double worst = 0, host_write = 0;
        auto t0 = std::chrono::steady_clock::now();
        for (unsigned c = 1; c <= N && ok; c++) {
          auto a = std::chrono::steady_clock::now();
          const int npk = own ? grid : 1;
          alignas(64) unsigned tmpl[64];
          for (int s = 0; s < 6; s++) { for (int k = 0; k < 7; k++) { float f = (float)(c + k); memcpy(&tmpl[s * 8 + k], &f, 4); } tmpl[s * 8 + 7] = c; }
          for (int b = 0; b < npk; b++) {
            unsigned *p = h_pk + (size_t)b * (PKT / 4);
            if (wmode == 0) {
              for (int s = 0; s < 6; s++) {
                memcpy(p + s * 8, tmpl + s * 8, 28);
                __asm__ __volatile__("" ::: "memory");
                ((volatile unsigned *)p)[s * 8 + 7] = c;                  // the sector's tag, after its payload (x86: stores stay in order)
              }
            } else if (wmode == 1) {   // two 16-byte stores per sector, the tag in the second
              for (int s = 0; s < 12; s++) _mm_store_si128((__m128i *)(p + s * 4), _mm_load_si128((const __m128i *)(tmpl + s * 4)));
            } else {                   // non-temporal: full 64-byte lines through the write-combining buffers
              for (int s = 0; s < 12; s++) _mm_stream_si128((__m128i *)(p + s * 4), _mm_load_si128((const __m128i *)(tmpl + s * 4)));
            }
          }
          if (wmode == 2) _mm_sfence();
          auto w = std::chrono::steady_clock::now();
          uint64_t spins = 0;
          while (h_done[0] != c) { _mm_pause(); if (++spins > 200000000ull) { ok = false; break; } }
          auto e = std::chrono::steady_clock::now();
          host_write += std::chrono::duration<double, std::micro>(w - a).count();
          double us = std::chrono::duration<double, std::micro>(e - a).count();
          if (c > 10 && us > worst) worst = us;
        }
        double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / N;
        if (!ok) { for (int b = 0; b < 148; b++) for (int s = 0; s < 6; s++) h_pk[(size_t)b * (PKT / 4) + s * 8 + 7] = 0xffffffffu; }
        cudaError_t e = cudaStreamSynchronize(st);
        unsigned herr = 0;
        cudaMemcpy(&herr, errs, 4, cudaMemcpyDeviceToHost);
        printf("write mode %d (0 scalar + tag, 1 SSE 16 B, 2 non-temporal) payload errors %u | ", wmode, herr);
        printf("grid %3d %s completion %d: %7.2f us per round trip (host packet writes %.2f us, worst %.1f)%s %s\n", grid, own ? "own packet per CTA" : "one shared packet  ",
               completion, us, host_write / N, worst, ok ? "" : "  TIMED OUT", e == cudaSuccess ? "" : cudaGetErrorString(e));
        fflush(stdout);
      }
  return 0;
}
