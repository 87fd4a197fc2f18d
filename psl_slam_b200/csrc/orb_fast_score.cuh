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

// Two pixels at once on packed u16x2 lanes (VIMNMX3.U16x2 on sm_100a): `ca`, `cb` point at the two
// centres.  3-input max/min give the 9-wide sliding windows in two steps (3, then 3 of 3).
// Returns score(a) | score(b) << 16 with scores clamped at 0 (a score <= 0 is never a corner).
template <int PITCH>
__device__ __forceinline__ unsigned fast_score_pair(const uint8_t* ca, const uint8_t* cb) {
  unsigned p[16];
#define PSL_P2(k, off) p[k] = (unsigned)ca[off] | ((unsigned)cb[off] << 16);
  PSL_P2(0, 3 * PITCH) PSL_P2(1, 3 * PITCH + 1) PSL_P2(2, 2 * PITCH + 2) PSL_P2(3, PITCH + 3)
  PSL_P2(4, 3) PSL_P2(5, -PITCH + 3) PSL_P2(6, -2 * PITCH + 2) PSL_P2(7, -3 * PITCH + 1)
  PSL_P2(8, -3 * PITCH) PSL_P2(9, -3 * PITCH - 1) PSL_P2(10, -2 * PITCH - 2) PSL_P2(11, -PITCH - 3)
  PSL_P2(12, -3) PSL_P2(13, PITCH - 3) PSL_P2(14, 2 * PITCH - 2) PSL_P2(15, 3 * PITCH - 1)
#undef PSL_P2
  unsigned h3[16], l3[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    h3[i] = __vimax3_u16x2(p[i], p[(i + 1) & 15], p[(i + 2) & 15]);
    l3[i] = __vimin3_u16x2(p[i], p[(i + 1) & 15], p[(i + 2) & 15]);
  }
  unsigned H = 0xFFFFFFFFu, L = 0u;  // min over arcs of the arc's brightest pixel / max of the darkest
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    H = __vminu2(H, __vimax3_u16x2(h3[i], h3[(i + 3) & 15], h3[(i + 6) & 15]));
    L = __vmaxu2(L, __vimin3_u16x2(l3[i], l3[(i + 3) & 15], l3[(i + 6) & 15]));
  }
  const int va = ca[0], vb = cb[0];
  const int sa = max(va - (int)(H & 0xFFFFu), (int)(L & 0xFFFFu) - va) - 1;
  const int sb = max(vb - (int)(H >> 16), (int)(L >> 16) - vb) - 1;
  return (unsigned)max(sa, 0) | ((unsigned)max(sb, 0) << 16);
}

}  // namespace psl
