// Exact FAST-9/16 corner score of the pixel at `c` inside a u8 tile with row pitch PITCH
// (SURVEY.md App. A3):  score = max over the 16 arcs of 9 contiguous circle pixels of
// max(min_k (I(c) - I(p_k)), min_k (I(p_k) - I(c))) - 1.
// Written on the non-negative pixel values (min over an arc of I(c)-I(p) = I(c) - max I(p)),
// sliding-window max/min by doubling: 2, 4, 8, then +1.
#pragma once
#include <stdint.h>

namespace psl {

template <int PITCH>
__device__ __forceinline__ int fast_score_at(const uint8_t* c) {
  unsigned p[16];
  p[0] = c[3 * PITCH];       p[1] = c[3 * PITCH + 1];   p[2] = c[2 * PITCH + 2];   p[3] = c[PITCH + 3];
  p[4] = c[3];               p[5] = c[-PITCH + 3];      p[6] = c[-2 * PITCH + 2];  p[7] = c[-3 * PITCH + 1];
  p[8] = c[-3 * PITCH];      p[9] = c[-3 * PITCH - 1];  p[10] = c[-2 * PITCH - 2]; p[11] = c[-PITCH - 3];
  p[12] = c[-3];             p[13] = c[PITCH - 3];      p[14] = c[2 * PITCH - 2];  p[15] = c[3 * PITCH - 1];
  const int v = c[0];
  unsigned h2[16], h4[16], h8[16], l2[16], l4[16], l8[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { h2[i] = max(p[i], p[(i + 1) & 15]); l2[i] = min(p[i], p[(i + 1) & 15]); }
#pragma unroll
  for (int i = 0; i < 16; ++i) { h4[i] = max(h2[i], h2[(i + 2) & 15]); l4[i] = min(l2[i], l2[(i + 2) & 15]); }
#pragma unroll
  for (int i = 0; i < 16; ++i) { h8[i] = max(h4[i], h4[(i + 4) & 15]); l8[i] = min(l4[i], l4[(i + 4) & 15]); }
  int best = -255;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int hi = (int)max(h8[i], p[(i + 8) & 15]);  // brightest pixel of the arc starting at i
    const int lo = (int)min(l8[i], p[(i + 8) & 15]);  // darkest
    best = max(best, max(v - hi, lo - v));
  }
  return best - 1;
}

}  // namespace psl
