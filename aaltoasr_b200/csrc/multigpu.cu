// multigpu.cu -- what the utterance-sharded run over several GPUs needs from the library:
//   * per-utterance, order-sensitive checksums of LNA records, computed on the device -- by the rank that produced the
//     records and again by the sink that received them (1-GPU-vs-N-GPU equality, checksum sink of the LNA gather);
//   * device buffers that can be mapped into another process of the same node (CUDA IPC), so that the LNA kernel of a
//     rank stores its records straight into the writer rank's rotating buffer over NVLink (fused epilogue + gather:
//     no send buffer, no separate copy kernel).
// The reference has no counterpart: its processes (`phone_probs -B N -I i`, aku/phone_probs.cc:78-79,135-139) share
// nothing and write their own files.
#include "ctx.hpp"
#include "kernels.hpp"

namespace akugpu {

// row(f) = sum_j (2j+1) w_j,  w_j = little-endian 32-bit word j of the frame's record (zero padded to whole words)
// utt(u) = sum_i (2i+1) row(first frame of u + i)          -- all modulo 2^64
// One warp per frame; the owning utterance by binary search in the frame-offset table.
__global__ void __launch_bounds__(256)
utt_checksum_kernel(const uint8_t *__restrict__ rec, int64_t first_frame, int64_t nf, int64_t rec_bytes,
                    const int64_t *__restrict__ fo, int n_utts, unsigned long long *__restrict__ acc)
{
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool aligned = ((uintptr_t)rec & 3) == 0 && (rec_bytes & 3) == 0;
  const int64_t nw = rec_bytes >> 2;
  for (int64_t f = warp0; f < nf; f += nwarps) {
    const uint8_t *row = rec + f * rec_bytes;
    unsigned long long h = 0;
    if (aligned) {
      const uint32_t *w = reinterpret_cast<const uint32_t *>(row);
      for (int64_t j = lane; j < nw; j += 32) h += (unsigned long long)(2 * j + 1) * __ldcs(w + j);
    } else {
      for (int64_t j = lane; j * 4 < rec_bytes; j += 32) {
        uint32_t v = 0;
        for (int b = 0; b < 4; b++)
          if (j * 4 + b < rec_bytes) v |= (uint32_t)row[j * 4 + b] << (8 * b);
        h += (unsigned long long)(2 * j + 1) * v;
      }
    }
    for (int o = 16; o > 0; o >>= 1) h += __shfl_down_sync(0xffffffffu, h, o);
    if (lane == 0) {
      const int64_t g = first_frame + f;
      int lo = 0, hi = n_utts;                  // largest u with fo[u] <= g
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (fo[mid] <= g) lo = mid; else hi = mid; }
      atomicAdd(acc + lo, h * (unsigned long long)(2 * (g - fo[lo]) + 1));
    }
  }
}

void checksum_begin(akugpu_ctx *ctx, const int64_t *frame_offsets, int n_utts, int64_t rec_bytes)
{
  if (n_utts < 0 || rec_bytes <= 0 || (n_utts && !frame_offsets)) throw Error(AKUGPU_E_ARG, "checksum: bad n_utts / rec_bytes / NULL frame_offsets");
  for (int u = 0; u < n_utts; u++)
    if (frame_offsets[u + 1] < frame_offsets[u]) throw Error(AKUGPU_E_ARG, "checksum: frame_offsets must be non-decreasing");
  ctx->chk_n_utts = n_utts;
  ctx->chk_rec = rec_bytes;
  ctx->chk_frames = n_utts ? frame_offsets[n_utts] : 0;
  ctx->chk_first = n_utts ? frame_offsets[0] : 0;
  ctx->d_chk_fo.reserve((size_t)(n_utts + 1) * sizeof(int64_t));
  ctx->d_chk_acc.reserve(std::max<size_t>(8, (size_t)n_utts * 8));
  if (n_utts) AKU_CUDA(cudaMemcpyAsync(ctx->d_chk_fo.p, frame_offsets, (size_t)(n_utts + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
  AKU_CUDA(cudaMemsetAsync(ctx->d_chk_acc.p, 0, std::max<size_t>(8, (size_t)n_utts * 8), ctx->stream));
  AKU_CUDA(cudaStreamSynchronize(ctx->stream));     // frame_offsets may be a temporary of the caller
  ctx->chk_open = true;
}

void checksum_update(akugpu_ctx *ctx, const uint8_t *records, int64_t first_frame, int64_t n_frames)
{
  if (!ctx->chk_open) throw Error(AKUGPU_E_STATE, "akugpu_checksum_update without akugpu_checksum_begin");
  if (n_frames <= 0) return;
  if (!records || !is_device_ptr(records)) throw Error(AKUGPU_E_ARG, "checksum: records must be a device buffer");
  if (first_frame < ctx->chk_first || first_frame + n_frames > ctx->chk_frames)
    throw Error(AKUGPU_E_ARG, "checksum: frame range outside the frame-offset table");
  const int64_t warps = std::min<int64_t>(n_frames, (int64_t)ctx->sm_count * 8 * 8);
  utt_checksum_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, ctx->stream>>>(records, first_frame, n_frames, ctx->chk_rec,
                                                                          ctx->d_chk_fo.as<int64_t>(), ctx->chk_n_utts,
                                                                          ctx->d_chk_acc.as<unsigned long long>());
  AKU_CUDA(cudaGetLastError());
  ctx->launches++;
}

void checksum_end(akugpu_ctx *ctx, uint64_t *out)
{
  if (!ctx->chk_open) throw Error(AKUGPU_E_STATE, "akugpu_checksum_end without akugpu_checksum_begin");
  if (ctx->chk_n_utts && !out) throw Error(AKUGPU_E_ARG, "checksum: out is NULL");
  if (ctx->chk_n_utts)
    AKU_CUDA(cudaMemcpyAsync(out, ctx->d_chk_acc.p, (size_t)ctx->chk_n_utts * 8, cudaMemcpyDeviceToHost, ctx->stream));
  AKU_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->chk_open = false;
}

}  // namespace akugpu

using namespace akugpu;

#define API_BEGIN                                             \
  if (!ctx) return AKUGPU_E_ARG;                              \
  try {                                                       \
    AKU_CUDA(cudaSetDevice(ctx->device));
#define API_END                                               \
    return AKUGPU_OK;                                         \
  } catch (const Error &e) {                                  \
    ctx->err = e.msg;                                         \
    return e.code;                                            \
  } catch (const std::exception &e) {                         \
    ctx->err = e.what();                                      \
    return AKUGPU_E_ARG;                                      \
  }

extern "C" {

int akugpu_checksum_begin(akugpu_ctx *ctx, const int64_t *frame_offsets, int n_utts, int64_t rec_bytes)
{
  API_BEGIN
  checksum_begin(ctx, frame_offsets, n_utts, rec_bytes);
  API_END
}

int akugpu_checksum_update(akugpu_ctx *ctx, const uint8_t *records, int64_t first_frame, int64_t n_frames)
{
  API_BEGIN
  checksum_update(ctx, records, first_frame, n_frames);
  API_END
}

int akugpu_checksum_end(akugpu_ctx *ctx, uint64_t *utt_checksums)
{
  API_BEGIN
  checksum_end(ctx, utt_checksums);
  API_END
}

// ---- device buffers shared between the processes of one node (one process per GPU) ----
int akugpu_shared_alloc(akugpu_ctx *ctx, size_t bytes, void **dev_ptr, unsigned char handle[64])
{
  API_BEGIN
  if (!dev_ptr || !handle || bytes == 0) throw Error(AKUGPU_E_ARG, "shared_alloc: NULL argument / zero size");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI carries the IPC handle as 64 opaque bytes");
  void *p = nullptr;
  AKU_CUDA(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); cudaGetLastError(); throw Error(AKUGPU_E_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
  memcpy(handle, &h, 64);
  *dev_ptr = p;
  ctx->shared_own.push_back(p);
  API_END
}

int akugpu_shared_open(akugpu_ctx *ctx, const unsigned char handle[64], void **dev_ptr)
{
  API_BEGIN
  if (!dev_ptr || !handle) throw Error(AKUGPU_E_ARG, "shared_open: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  void *p = nullptr;
  AKU_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));      // maps the peer's buffer, enables peer access
  cudaGetLastError();   // "peer access already enabled" (NCCL may have done it) is not an error of this call
  *dev_ptr = p;
  ctx->shared_peer.push_back(p);
  API_END
}

// Asynchronous device-to-device copy on the context's stream (local slot -> the writer's mapped buffer: a copy engine
// moves the records over NVLink while the SMs score the next sub-batch).
int akugpu_copy_async(akugpu_ctx *ctx, void *dst, const void *src, size_t bytes)
{
  API_BEGIN
  if (bytes && (!dst || !src)) throw Error(AKUGPU_E_ARG, "copy_async: NULL pointer");
  if (bytes) AKU_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
  API_END
}

int akugpu_shared_release(akugpu_ctx *ctx, void *dev_ptr)
{
  API_BEGIN
  if (!dev_ptr) return AKUGPU_OK;
  AKU_CUDA(cudaStreamSynchronize(ctx->stream));
  for (size_t i = 0; i < ctx->shared_own.size(); i++)
    if (ctx->shared_own[i] == dev_ptr) {
      ctx->shared_own.erase(ctx->shared_own.begin() + i);
      AKU_CUDA(cudaFree(dev_ptr));
      return AKUGPU_OK;
    }
  for (size_t i = 0; i < ctx->shared_peer.size(); i++)
    if (ctx->shared_peer[i] == dev_ptr) {
      ctx->shared_peer.erase(ctx->shared_peer.begin() + i);
      AKU_CUDA(cudaIpcCloseMemHandle(dev_ptr));
      return AKUGPU_OK;
    }
  throw Error(AKUGPU_E_ARG, "shared_release: not a buffer of this context");
  API_END
}

}  // extern "C"
