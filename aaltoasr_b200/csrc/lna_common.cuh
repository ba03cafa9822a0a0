// lna_common.cuh -- float-cast emulation shared by the LNA epilogue and the scorer's fused normaliser pass.
#pragma once
namespace akugpu {
constexpr float LN_2M150 = -103.97207708399179f;   // ln 2^-150: (float)x == 0 at or below this
constexpr float LN_2M126 = -87.33654475055310f;    // ln 2^-126: smallest normal float
constexpr float LN_2P149 = 103.27892990343184f;    // 149 ln 2
constexpr float LP_FLOOR = -115.12925464970229f;   // (float) log(1e-50)

// ln( (float) exp(v) ), -inf when the float is zero (aku/phone_probs.cc:228: obs_log_probs is vector<float>).
__device__ __forceinline__ float log_of_float_cast(float v)
{
  if (v >= LN_2M126) return v;
  if (v <= LN_2M150) return -INFINITY;
  float q = rintf(expf(v + LN_2P149));
  if (q < 1.f) return -INFINITY;
  return logf(q) - LN_2P149;
}
}  // namespace akugpu
