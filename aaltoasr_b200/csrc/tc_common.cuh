// tc_common.cuh -- PTX wrappers shared by the tensor-core scorers (gmm_tc.cu, gmm_tc16.cu): mbarriers, TMA
// bulk-tensor loads, tcgen05 MMA / commit, shared-memory matrix descriptors.  sm_100a only.
#pragma once
#include <cuda.h>
#include <stdint.h>
#include "ctx.hpp"

namespace akugpu {
namespace tc {

constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nTC_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra TC_DONE;\nbra TC_WAIT;\nTC_DONE:\n}\n" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// The same load delivered to the same shared-memory offset of every CTA of the cluster named in `mask` (each destination's
// barrier at the same offset receives the bytes).
__device__ __forceinline__ void tma_load_2d_mc(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4, LBO = 1,
// SBO = 1024 B (8 rows x 128 B) >> 4, version 1 (Blackwell), layout type 2.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// The same for SWIZZLE_64B: rows of 64 B (32 fp16), SBO = 512 B (8 rows x 64 B) >> 4, layout type 4.
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bit 4), A/B format at bits 7/10 (0 = F16, 1 = BF16),
// K-major both, N>>3 at bit 17, M>>4 at bit 24.
constexpr uint32_t umma_idesc(int m, int n, bool bf16) {
  return (1u << 4) | (bf16 ? ((1u << 7) | (1u << 10)) : 0u) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// the arrival delivered to the barrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Exactly one lane of the (converged) calling warp returns true.  elect.sync tells the compiler that the branch is
// single-threaded, so tcgen05.mma / commit / TMA are emitted straight (a plain `lane == 0` test makes it wrap every one
// of them in a uniformisation loop: measured ~200 clk per MMA issue instead of the pipe's 64-87).
__device__ __forceinline__ bool elect_one()
{
  uint32_t p;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(p));
  return p != 0;
}

#define AKU_TMEM_LD16(r, taddr)                                                                                              \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"    \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), \
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                    \
               : "r"(taddr)                                                                                                  \
               : "memory")
#define AKU_TMEM_LD_WAIT() asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory")

}  // namespace tc

// host helpers implemented in gmm_tc.cu
// boxes of 128 rows x 64 columns with SWIZZLE_128B (default) or 128 x 32 with SWIZZLE_64B
void tc_make_map(CUtensorMap *map, void *base, uint64_t rows, uint64_t cols, bool fp16, int box_cols = 64, int box_rows = 128);
double tc_expanded_params(const HostModel &hm, bool full, int L, std::vector<double> &cen, std::vector<double> &theta,
                          std::vector<double> &gconst, std::vector<double> *q_of_gauss);
// Largest cancelling magnitude q the expanded form is trusted with: predicted log-likelihood error 4e-7 * q <= 8e-5.
constexpr double TC_Q_MAX = 200.0;
void tc_build_slots(const HostModel &hm, int group, std::vector<int> &slot_state, std::vector<int> &slot_k0, std::vector<int> &slot_flags,
                    const std::vector<char> *skip_state);
const int *tc_tile_ranges(akugpu_ctx *ctx, int n_tiles, const std::vector<char> &clean,
                          std::map<int, std::pair<int, std::shared_ptr<DevBuf>>> &ranges, int want, int &got);
}  // namespace akugpu
