// stream_common.cuh -- constants and host helpers shared by the streaming-regime scorers (gmm_stream.cu: one launch per
// call; gmm_resident.cu: a resident kernel that keeps the parameter image in shared memory between calls).
#pragma once
#include "ctx.hpp"
#include "tc_common.cuh"

namespace akugpu {

namespace tcs {
constexpr int BM = 128, BK = 64, GR = 16, SLOTS = BM / GR;
constexpr uint32_t B_BLOCK = BM * BK * 2;          // 16 KB: one k-block of a component tile of B'
constexpr int THREADS = 384;                       // warp 0 TMA, warp 1 MMA, warps 4-11 epilogue (two groups of four)
constexpr int EPI_THREADS = 256, GROUP_THREADS = 128;
constexpr int MAX_TSLOTS = 4;
constexpr float LO_INV = 1.f / 2048.f, LO_SCALE = 2048.f;
constexpr int XS_DIM = 40;                         // centred features of up to NF frames x 40 dims travel as kernel parameters
}  // namespace tcs

// pinned, mapped host block of a context's streaming calls:
//   bytes   0.. 15  device -> host: word 0 = sequence number of the last completed call, word 1 = fp16-range flag
//   bytes  16.. 31  device -> host: timings of the resident scorer's last call (akugpu_stream_stats)
//   bytes 256..     centred features [STREAM_MAX_FRAMES][64] floats (launch-per-call scorer with wide features; calls the
//                   resident scorer relays through CTA 0), then results [STREAM_MAX_FRAMES][S] floats
// (the resident scorer's commands travel in a block of their own: one 512-byte packet per CTA, gmm_resident.cu)
constexpr int STREAM_X_BYTE = 256;
void stream_buffers(akugpu_ctx *ctx, int S);
void stream_map_ready(akugpu_ctx *ctx);

}  // namespace akugpu
