// Counter-based dropout masks shared by the attention and glue kernels (inverted dropout, train mode of
// memory_attention.py:58-99 and transformer.py:304-306).  A mask bit is a pure function of
// (per-call 64-bit seed, site id, element index), so the backward kernels regenerate exactly the forward's mask and
// no mask tensor is ever stored.  keep  <=>  mix32(index ^ key(seed, site)) >= p * 2^32.
// The seed is read from DEVICE memory, so a CUDA graph replays with fresh masks.
#pragma once

#include <stdint.h>

namespace sam2b200 {

struct Dropout {
  const unsigned long long* seed;   // nullptr = dropout off
  uint32_t site;                    // which dropout of the stack (layer * 8 + site)
  uint32_t thresh;                  // round(p * 2^32); an element is kept iff its 32 random bits >= thresh
  float inv_keep;                   // 1 / (1 - p)
};

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {   // "lowbias32" integer hash (full avalanche)
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t dropout_key(unsigned long long seed, uint32_t site) {
  return mix32(mix32((uint32_t)seed ^ (site * 0x9e3779b9u)) ^ (uint32_t)(seed >> 32));
}
__device__ __forceinline__ bool dropout_keep(uint32_t key, uint32_t index, uint32_t thresh) {
  return mix32(index ^ key) >= thresh;
}
inline Dropout make_dropout(const unsigned long long* seed, uint32_t site, float p) {
  Dropout d{nullptr, site, 0u, 1.0f};
  if (seed != nullptr && p > 0.f) {
    d.seed = seed;
    const double t = (double)p * 4294967296.0;
    d.thresh = t >= 4294967295.0 ? 0xffffffffu : (uint32_t)(t + 0.5);
    d.inv_keep = 1.0f / (1.0f - p);
  }
  return d;
}

}  // namespace sam2b200
