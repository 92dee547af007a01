// Noise streams of the Bayesian-network kernels (bnn.cuh: inference, layered.cuh: training) -- shared
// device helpers only, no kernels (both translation units include this file).
#pragma once
#include "common.cuh"

namespace bgm {
namespace bnn {

constexpr uint32_t NOISE_BNN_W = 5, NOISE_BNN_SIGN = 6;
enum { NET_G = 0, NET_F = 1, NET_H = 2, NET_E = 3 };

// 32 sign bits starting at bit `o` of the (net, layer) sign stream of this row / call
__device__ __forceinline__ uint32_t sign_bits32(uint64_t seed, int64_t grow, uint32_t call, int net_id, int l, int o) {
  const int w0 = o >> 5, sh = o & 31;
  const uint32_t jb = ((uint32_t)net_id << 8) | ((uint32_t)l << 4);
  const uint4 b0 = noise_block(seed, grow, call, NOISE_BNN_SIGN, jb | (uint32_t)(w0 >> 2));
  const int i0 = w0 & 3;
  const uint32_t lo = i0 == 0 ? b0.x : (i0 == 1 ? b0.y : (i0 == 2 ? b0.z : b0.w));
  if (sh == 0) return lo;
  uint32_t hi;
  if (i0 < 3) {
    hi = i0 == 0 ? b0.y : (i0 == 1 ? b0.z : b0.w);
  } else {
    hi = noise_block(seed, grow, call, NOISE_BNN_SIGN, jb | (uint32_t)((w0 + 1) >> 2)).x;
  }
  return (lo >> sh) | (hi << (32 - sh));
}

}  // namespace bnn
}  // namespace bgm
