// Replica of the scorer's main loop (v4: 8 warps, 128 comps per stage, frame-pair FFMA2) with optional
// per-stage overheads, to find what separates the kernel (56 TFLOP/s) from the bare loop (67).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NPAIR = 64, TC = 128, DP = 20, GR = 16;

// STAGED 0: bare loop; 1: re-init + min-sink; 2: re-init + atomic + full sum; 3: full sum only (no re-init);
// 5: atomic + min-sink (no re-init)
// (old) STAGED = 0: one endless loop over the same smem;  1: re-init accumulators from smem every 20 dp (a "stage")
// 2: + warp-level atomic + fence at each stage end (the re-arm protocol without TMA)
template <int STAGED>
__global__ void __launch_bounds__(256, 2) k(float *out, int stages, int stagger_ns)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4 *xs = reinterpret_cast<float4 *>(smem_raw);                    // [DP][NPAIR]
  float4 *ps0 = reinterpret_cast<float4 *>(smem_raw + DP * NPAIR * 16);  // [DP][TC]
  float *cs = reinterpret_cast<float *>(ps0 + DP * TC);
  __shared__ int done_cnt;
  for (int i = threadIdx.x; i < DP * NPAIR * 4 + DP * TC * 4 + TC; i += 256) reinterpret_cast<float *>(smem_raw)[i] = 0.001f * (i % 97);
  if (threadIdx.x == 0) done_cnt = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float4 *ps = ps0 + warp * GR;
  float2 acc[2][GR];
  for (int c = 0; c < GR; c++) { acc[0][c] = make_float2(0.f, 0.f); acc[1][c] = make_float2(0.f, 0.f); }
  float r = 0.f;
  if (stagger_ns > 0) {
    int slot = (warp >> 2) + 2 * (blockIdx.x & 1);   // the 4 warps that share an SM sub-partition
    for (int q = 0; q < slot; q++) __nanosleep(stagger_ns);
  }
  for (int t = 0; t < stages; t++) {
    if (STAGED == 1 || STAGED == 2 || STAGED == 4) {
#pragma unroll
      for (int c = 0; c < GR; c++) { float nc = cs[warp * GR + c]; acc[0][c] = make_float2(nc, nc); acc[1][c] = make_float2(nc, nc); }
    }
#pragma unroll 2
    for (int dp = 0; dp < DP; dp++) {
      const float4 xa = xs[dp * NPAIR + lane], xb = xs[dp * NPAIR + 32 + lane];
#pragma unroll
      for (int c = 0; c < GR; c++) {
        const float4 p = ps[dp * TC + c];
        float2 t0 = __ffma2_rn(make_float2(xa.x, xa.y), make_float2(p.x, p.x), make_float2(p.z, p.z));
        float2 u0 = __ffma2_rn(make_float2(xb.x, xb.y), make_float2(p.x, p.x), make_float2(p.z, p.z));
        float2 t1 = __ffma2_rn(make_float2(xa.z, xa.w), make_float2(p.y, p.y), make_float2(p.w, p.w));
        float2 u1 = __ffma2_rn(make_float2(xb.z, xb.w), make_float2(p.y, p.y), make_float2(p.w, p.w));
        acc[0][c] = __ffma2_rn(t0, t0, acc[0][c]);
        acc[1][c] = __ffma2_rn(u0, u0, acc[1][c]);
        acc[0][c] = __ffma2_rn(t1, t1, acc[0][c]);
        acc[1][c] = __ffma2_rn(u1, u1, acc[1][c]);
      }
    }
    if (STAGED == 2 || STAGED == 5) {
      __syncwarp();
      if (lane == 0) {
        int old = atomicAdd(&done_cnt, 1);
        if (old == 7) { done_cnt = 0; __threadfence_block(); }
      }
    }
    if (STAGED == 2 || STAGED == 3) {
#pragma unroll
      for (int c = 0; c < GR; c++) r += acc[0][c].x + acc[0][c].y + acc[1][c].x + acc[1][c].y;
    }
    if (STAGED == 1 || STAGED == 4 || STAGED == 5) {   // cheap sink that keeps every accumulator alive
      float2 m = acc[0][0];
#pragma unroll
      for (int c = 0; c < GR; c++) { m.x = fminf(m.x, fminf(acc[0][c].x, acc[1][c].x)); m.y = fminf(m.y, fminf(acc[0][c].y, acc[1][c].y)); }
      r += m.x + m.y;
    }
  }
  for (int c = 0; c < GR; c++) r += acc[0][c].x + acc[0][c].y + acc[1][c].x + acc[1][c].y;
  if (r == 1.2345f) out[0] = r;
}

template <int STAGED>
void run(int sms, int stagger_ns)
{
  float *d; cudaMalloc(&d, 16);
  int smem = DP * NPAIR * 16 + 2 * (DP * TC * 16 + TC * 4 + 32) + 8192;    // same footprint as the scorer
  cudaFuncSetAttribute(k<STAGED>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int stages = 625;
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    k<STAGED><<<sms * 2, 256, smem>>>(d, stages, stagger_ns);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  double fma = (double)sms * 2 * 256 * stages * DP * 128.0 * 2;
  printf("staged %d stagger %5d ns : %.2f ms  %.3e lane-FMA/s = %.1f TFLOP/s (%s)\n", STAGED, stagger_ns, best, fma / (best * 1e-3), 2 * fma / (best * 1e-3) / 1e12,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int n = p.multiProcessorCount;
  run<0>(n, 0); run<1>(n, 0); run<2>(n, 0); run<3>(n, 0); run<5>(n, 0);
  return 0;
}
