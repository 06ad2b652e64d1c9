// Counter-based dropout masks for the head training step (train.py runs the head in model.train() mode:
// init_dropout on the encoder output and the TransformerEncoderLayer's own dropout on the attention weights, after
// the attention block, inside the FFN and after the FFN; lib/models.py:291-319). A mask bit is a pure function of
// (seed, site, element index), so the forward and every backward pass regenerate it instead of storing it, and a
// test can rebuild the same masks on the host:
//     keep(idx) = lowbias32(idx ^ key) >= thresh,   key = lowbias32(seed * 0x9E3779B9 + site),   thresh = p * 2^32
// (lowbias32: the two-round multiply-xorshift integer hash). Kept values are scaled by 1 / (1 - p).
#pragma once
#include <stdint.h>

namespace w2v {

struct DropSite {
  uint32_t key;       // per (seed, site)
  uint32_t thresh;    // 0 = no dropout
  float inv_keep;     // 1 / (1 - p)
};

__host__ __device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ DropSite make_drop_site(float p, uint32_t seed, uint32_t site) {
  DropSite d;
  d.key = lowbias32(seed * 0x9E3779B9U + site);
  d.thresh = p > 0.f ? (uint32_t)((double)p * 4294967296.0) : 0U;
  d.inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  return d;
}
// factor applied to element idx: 0 (dropped) or 1 / (1 - p)
__device__ __forceinline__ float drop_factor(const DropSite& d, uint32_t idx) {
  return lowbias32(idx ^ d.key) >= d.thresh ? d.inv_keep : 0.f;
}

}  // namespace w2v
