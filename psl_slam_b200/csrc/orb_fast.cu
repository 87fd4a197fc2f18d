// K2: the cell loop of ORBextractor::ComputeKeyPointsOctTree (ORBextractor.cc:789-829).
//
// The reference runs cv::FAST(threshold, nonmax=true) on every ~30x30 cell sub-image (with a
// 6-px overlap) and retries the cell with minThFAST when iniThFAST finds nothing.  Facts used
// (re-verified against cv2 in tests/golden):
//   * a cell's evaluated interior is its sub-image minus a 3-px rim; interiors tile the level
//     window without overlap, and the 7x7 support of an interior pixel lies inside the cell;
//   * score(p) = max over the 16 arcs of 9 contiguous circle pixels of max(min d, min -d) - 1
//     and p is a corner at threshold t  <=>  score(p) >= t   (threshold independent);
//   * NMS keeps p iff score(p) > score(q) for the 8 neighbours q inside the interior that are
//     corners; since sub-threshold neighbours have score < t <= score(p), the NMS verdict does
//     not depend on t either.
// One CTA per cell: (A) a two-pair compass reject compacts the pixels that can be corners,
// (B) their exact scores are computed two pixels per thread on packed u16x2 min/max,
// (C) NMS runs over the survivors only and marks two bitmaps (>= iniTh, >= minTh), and
// (D) the ini/min fallback is a count, the raster-ordered output a popc prefix sum over the
// bitmap words.  Cells are stitched in row-major order later by K3.
#include "orb_fast_score.cuh"
#include "orb_kernels.cuh"

namespace psl {

constexpr int kFastThreads = 128;
constexpr int kTilePitch = 72;     // bytes per tile row (>= kMaxCellDim + 3 alignment slack, multiple of 4)
constexpr int kTileWords = kTilePitch / 4;
constexpr int kInteriorMax = 60;   // kMaxCellDim - 6
constexpr int kMaxWords = (kInteriorMax * kInteriorMax + 31) / 32;  // bitmap words (113 <= threads)

template <bool ALIGNED>
__global__ void __launch_bounds__(kFastThreads)
    fast_cells_kernel(const OrbGeometry* __restrict__ geo, ImgBatch in0, int ini_th, int min_th,
                      uint32_t* __restrict__ pool, int pool_cap, uint32_t* __restrict__ pool_count,
                      uint2* __restrict__ cell_tab, uint32_t* __restrict__ status) {
  __shared__ __align__(16) uint8_t s_tile[kMaxCellDim][kTilePitch];
  __shared__ __align__(16) uint8_t s_score[kInteriorMax + 2][kInteriorMax + 4];  // 1-px zero rim for the NMS
  __shared__ uint16_t s_list[kInteriorMax * kInteriorMax];
  __shared__ uint32_t s_keep[2][kMaxWords];
  __shared__ int s_nlist;
  __shared__ int s_warp[2][kFastThreads / 32];
  __shared__ uint32_t s_base;

  const int cell = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  int lvl = 0;
  const int nl = geo->nlevels;
  for (int l = 1; l < nl; ++l)
    if (cell >= geo->grid[l].first_cell) lvl = l;
  const CellGrid g = geo->grid[lvl];
  const int local = cell - g.first_cell;
  const int ci = local / g.n_cols, cj = local - ci * g.n_cols;
  uint2* tab = cell_tab + (size_t)b * geo->total_cells + cell;

  // :789-806 cell window and skip tests
  const int iniY = kMinBorder + ci * g.h_cell, iniX = kMinBorder + cj * g.w_cell;
  int maxY = iniY + g.h_cell + 6, maxX = iniX + g.w_cell + 6;
  if (iniY >= g.max_by - 3 || iniX >= g.max_bx - 6) {
    if (tid == 0) *tab = make_uint2(0u, 0u);
    return;
  }
  maxY = min(maxY, g.max_by);
  maxX = min(maxX, g.max_bx);
  const int tw = maxX - iniX, th = maxY - iniY;  // sub-image
  const int iw = tw - 6, ih = th - 6;            // evaluated interior
  if (iw <= 0 || ih <= 0) {
    if (tid == 0) *tab = make_uint2(0u, 0u);
    return;
  }
  const int npx = iw * ih;
  // i / d for i < 2^13, d <= 72 without an integer divide: the true quotient (i + 0.5) / d is at least
  // 0.5 / d away from an integer, far more than the fp32 rounding error of the product.
  const float rcp_iw = __frcp_rn((float)iw);
#define PSL_DIV(i, rcp) ((int)(((float)(i) + 0.5f) * (rcp)))

  const uint8_t* __restrict__ img;
  int pitch;
  if (lvl == 0) {
    img = in0.ptr + (size_t)b * in0.frame_stride;
    pitch = in0.pitch;
  } else {
    img = geo->level[lvl].ptr + (size_t)b * geo->level[lvl].frame_stride;
    pitch = geo->level[lvl].pitch;
  }
  // Tile load.  ALIGNED: rows are 4-byte aligned -> whole words starting at iniX rounded down; the tile
  // keeps that left slack (ax) so that shared and global addresses stay word-congruent.
  const int ax = ALIGNED ? (iniX & 3) : 0;
  if (ALIGNED) {
    const int nw = (tw + ax + 3) >> 2;
    const float rcp_nw = __frcp_rn((float)nw);
    const uint8_t* src = img + (size_t)iniY * pitch + (iniX - ax);
    for (int i = tid; i < nw * th; i += kFastThreads) {
      const int y = PSL_DIV(i, rcp_nw), x = i - y * nw;
      reinterpret_cast<uint32_t*>(&s_tile[y][0])[x] = __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)y * pitch) + x);
    }
  } else {
    const float rcp_tw = __frcp_rn((float)tw);
    for (int i = tid; i < tw * th; i += kFastThreads) {
      const int y = PSL_DIV(i, rcp_tw), x = i - y * tw;
      s_tile[y][x] = __ldg(img + (size_t)(iniY + y) * pitch + iniX + x);
    }
  }
  for (int i = tid; i < (kInteriorMax + 2) * (kInteriorMax + 4) / 4; i += kFastThreads)
    reinterpret_cast<uint32_t*>(&s_score[0][0])[i] = 0u;
  if (tid < kMaxWords) { s_keep[0][tid] = 0u; s_keep[1][tid] = 0u; }
  if (tid == 0) s_nlist = 0;
  __syncthreads();

  // The exact scores are the expensive part and only the pixels that pass the compass test need one.  At
  // iniThFAST far fewer pixels pass than at minThFAST, and a textured cell almost always has a corner at
  // iniThFAST, so the cell is first done at iniThFAST alone; only a cell that comes out empty is redone at
  // minThFAST (:812-816).  Scores and the NMS verdict do not depend on the threshold (see the header), so the
  // second pass simply extends the first.
  const int lane = tid & 31, wid = tid >> 5;
  const int nwords = (npx + 31) >> 5;
  uint32_t keep = 0;
  int total = 0, pos = 0;
  for (int pass = 0; pass < 2; ++pass) {
    const int t_cur = pass == 0 ? ini_th : min_th;
    if (pass == 1) {
      if (min_th >= ini_th) break;  // nothing new can appear
      if (tid == 0) s_nlist = 0;
      __syncthreads();
    }
    // ---- A: compass reject.  Any 9-arc contains one pixel of each antipodal pair, so a corner needs
    // (p0 or p8) and (p4 or p12) outside [v-t, v+t].  Compared on pixel values, never on differences
    // (nvcc 12.9 mis-packs min/max/abs of u8 differences for sm_100a).
    for (int i = tid; i < npx; i += kFastThreads) {
      const int y = PSL_DIV(i, rcp_iw), x = i - y * iw;
      const uint8_t* c = &s_tile[y + 3][x + 3 + ax];
      const int v = *c, hi = v + t_cur, lo = v - t_cur;
      const int q0 = c[3 * kTilePitch], q8 = c[-3 * kTilePitch];
      if (q0 >= lo && q0 <= hi && q8 >= lo && q8 <= hi) continue;
      const int q4 = c[3], q12 = c[-3];
      if (q4 >= lo && q4 <= hi && q12 >= lo && q12 <= hi) continue;
      s_list[atomicAdd(&s_nlist, 1)] = (uint16_t)i;
    }
    __syncthreads();

    // ---- B: exact scores, two survivors per thread on u16x2 lanes ---------------------------------
    const int nlist = s_nlist;
    for (int k = 2 * tid; k < nlist; k += 2 * kFastThreads) {
      const int ia = s_list[k], ib = s_list[min(k + 1, nlist - 1)];
      const int ya = PSL_DIV(ia, rcp_iw), xa = ia - ya * iw;
      const int yb = PSL_DIV(ib, rcp_iw), xb = ib - yb * iw;
      const unsigned sc = fast_score_pair<kTilePitch>(&s_tile[ya + 3][xa + 3 + ax], &s_tile[yb + 3][xb + 3 + ax]);
      const unsigned sa = sc & 0xFFFFu, sb = sc >> 16;
      s_score[ya + 1][xa + 1] = (uint8_t)(sa >= (unsigned)t_cur ? sa : 0u);
      s_score[yb + 1][xb + 1] = (uint8_t)(sb >= (unsigned)t_cur ? sb : 0u);
    }
    __syncthreads();

    // ---- C: NMS over the survivors -> keep-bitmap ---------------------------------------------------
    for (int k = tid; k < nlist; k += kFastThreads) {
      const int i = s_list[k];
      const int y = PSL_DIV(i, rcp_iw), x = i - y * iw;
      const unsigned sc = s_score[y + 1][x + 1];
      if (!sc) continue;
      const uint8_t* r0 = &s_score[y][x];
      const uint8_t* r1 = &s_score[y + 1][x];
      const uint8_t* r2 = &s_score[y + 2][x];
      const unsigned m = max(max(max((unsigned)r0[0], (unsigned)r0[1]), max((unsigned)r0[2], (unsigned)r1[0])),
                             max(max((unsigned)r1[2], (unsigned)r2[0]), max((unsigned)r2[1], (unsigned)r2[2])));
      if (sc > m) atomicOr(&s_keep[0][i >> 5], 1u << (i & 31));
    }
    __syncthreads();

    // ---- D: count + raster-ordered positions (thread t owns bitmap word t) -----------------------
    keep = tid < nwords ? s_keep[0][tid] : 0u;
    const int cnt = __popc(keep);
    int inc = cnt;
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
      const int a = __shfl_up_sync(0xffffffffu, inc, dlt);
      if (lane >= dlt) inc += a;
    }
    if (lane == 31) s_warp[pass][wid] = inc;
    __syncthreads();
    int base_w = 0;
    total = 0;
#pragma unroll
    for (int w = 0; w < kFastThreads / 32; ++w) {
      if (w < wid) base_w += s_warp[pass][w];
      total += s_warp[pass][w];
    }
    pos = base_w + inc - cnt;
    if (total > 0) break;  // :812-816 retry with minThFAST only when the cell is empty
  }
  if (tid == 0) {
    uint32_t base = 0;
    if (total) base = atomicAdd(pool_count + b, (uint32_t)total);
    s_base = base;
    *tab = make_uint2(base, (uint32_t)total);
    if (base + total > (uint32_t)pool_cap) atomicOr(status, kStatCandOverflow);
  }
  __syncthreads();
  const uint32_t base = s_base;
  if (!total || base + total > (uint32_t)pool_cap) return;
  uint32_t* out = pool + (size_t)b * pool_cap + base;
  while (keep) {
    const int bit = __ffs(keep) - 1;
    keep &= keep - 1;
    const int p = tid * 32 + bit;
    const int y = PSL_DIV(p, rcp_iw), x = p - y * iw;
    // cell-local + (j*wCell, i*hCell) (:820-825) == level coordinate - minBorder
    out[pos++] = pack_cand(x + 3 + cj * g.w_cell, y + 3 + ci * g.h_cell, s_score[y + 1][x + 1]);
  }
}

#undef PSL_DIV

void launch_fast_cells(const OrbGeometry* d_geo, const OrbGeometry& geo, ImgBatch in0, int ini_th, int min_th,
                       uint32_t* pool, int pool_cap, uint32_t* pool_count, uint2* cell_tab, uint32_t* status, int B,
                       cudaStream_t st) {
  dim3 grid(geo.total_cells, B);
  // levels >= 1 are ours (128-byte pitch); level 0 is the caller's image
  const bool aligned = ((uintptr_t)in0.ptr & 3) == 0 && (in0.pitch & 3) == 0 && (in0.frame_stride & 3) == 0;
  if (aligned)
    fast_cells_kernel<true><<<grid, kFastThreads, 0, st>>>(d_geo, in0, ini_th, min_th, pool, pool_cap, pool_count,
                                                           cell_tab, status);
  else
    fast_cells_kernel<false><<<grid, kFastThreads, 0, st>>>(d_geo, in0, ini_th, min_th, pool, pool_cap, pool_count,
                                                            cell_tab, status);
}

}  // namespace psl
