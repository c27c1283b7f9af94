// Counter-based random bits for the stochastic training mode (dropout masks, goal masking, multinomial routing).
//
// The reference draws them from torch's global generator (nn.Dropout, SDPA dropout_p, torch.bernoulli,
// torch.multinomial: modedit.py:149, :254, :389-390, :888). Those streams cannot be reproduced outside torch, so the
// engine defines its own: every random decision is a pure function of (seed, step, stream, layer, logical coordinates)
// — no state, no ordering dependence, identical in the forward and the backward kernels, and restated in numpy
// (oracle/mode_rng.py) so the reference can be run with exactly the same masks (tests/golden/make_train_goldens.py).
//
//   bits(key, idx) = lowbias32(idx ^ key)                 (a 32-bit bijection with full avalanche)
//   key(seed, step, stream, layer) = lowbias32(lowbias32(lowbias32(seed_lo ^ 0x9e3779b9) + seed_hi) + step) ... below
//
// A dropout decision uses 16 bits: keep iff bits16 >= round(p * 65536); one 32-bit word serves two neighbouring
// elements (low half: even element, high half: odd element).
#pragma once
#include <stdint.h>

namespace mode {

enum RngStream : uint32_t { RNG_GOAL = 1, RNG_ATTN = 2, RNG_MLP = 3, RNG_ROUTE = 4, RNG_EMBED = 5 };

__host__ __device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}

__host__ __device__ __forceinline__ uint32_t rng_key(uint64_t seed, uint32_t step, uint32_t stream, uint32_t layer) {
  uint32_t k = lowbias32(static_cast<uint32_t>(seed) ^ 0x9e3779b9U);
  k = lowbias32(k + static_cast<uint32_t>(seed >> 32));
  k = lowbias32(k + step);
  k = lowbias32(k + stream * 0x85ebca6bU + layer * 0xc2b2ae35U);
  return k;
}

__host__ __device__ __forceinline__ uint32_t rng_bits(uint32_t key, uint32_t idx) { return lowbias32(idx ^ key); }

// 16-bit dropout threshold: an element is DROPPED iff its 16 random bits are < thr
__host__ __device__ __forceinline__ uint32_t drop_threshold(float p) {
  const float t = p * 65536.0f + 0.5f;
  return t <= 0.f ? 0u : (t >= 65535.f ? 65535u : static_cast<uint32_t>(t));
}

// uniform in (0, 1): 23 random bits, (bits + 0.5) / 2^23 is exact in fp32 and never 0 or 1
__host__ __device__ __forceinline__ float rng_uniform(uint32_t bits) {
  return (static_cast<float>(bits >> 9) + 0.5f) * (1.0f / 8388608.0f);
}

struct DropoutSpec {
  uint32_t key;   // rng_key(seed, step, stream, layer)
  uint32_t thr;   // drop_threshold(p); 0 = no dropout
  float scale;    // 1 / (1 - p)
};

}  // namespace mode
