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
#include <cuda.h>

#include "orb_fast_score.cuh"
#include "orb_kernels.cuh"

namespace psl {

constexpr int kFastThreads = 128;
constexpr int kTilePitch = 72;     // bytes per tile row (>= kMaxCellDim + 3 alignment slack, multiple of 4)
constexpr int kTileWords = kTilePitch / 4;
constexpr int kInteriorMax = 60;   // kMaxCellDim - 6
constexpr int kMaxWords = (kInteriorMax * kInteriorMax + 31) / 32;  // bitmap words (113 <= threads)

// The per-cell form (one CTA per cell: tile -> smem, compass reject, exact scores, NMS, ordered compaction).  It
// serves the cells the dense path below hands over: a cell that has no corner at iniThFAST is redone here at
// minThFAST (first_pass = 1), and the whole job when a threshold is above the byte-SIMD range (first_pass = 0).
template <bool ALIGNED>
__device__ __forceinline__ void fast_cell_body(const OrbGeometry* __restrict__ geo, const ImgBatch& in0, int ini_th,
                                               int min_th, int first_pass, int cell, int b,
                                               uint32_t* __restrict__ pool, int pool_cap,
                                               uint32_t* __restrict__ pool_count, uint2* __restrict__ cell_tab,
                                               uint32_t* __restrict__ status) {
  __shared__ __align__(16) uint8_t s_tile[kMaxCellDim][kTilePitch];
  __shared__ __align__(16) uint8_t s_score[kInteriorMax + 2][kInteriorMax + 4];  // 1-px zero rim for the NMS
  __shared__ uint16_t s_list[kInteriorMax * kInteriorMax];
  __shared__ uint32_t s_keep[2][kMaxWords];
  __shared__ int s_nlist;
  __shared__ int s_warp[2][kFastThreads / 32];
  __shared__ uint32_t s_base;

  const int tid = threadIdx.x;
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
  for (int pass = first_pass; pass < 2; ++pass) {
    const int t_cur = pass == 0 ? ini_th : min_th;
    if (pass == 1 && first_pass == 0) {
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


// Work list form: item = b * total_cells + cell, either the fallback list the collect kernel wrote (n_items on
// the device) or, with list == nullptr, every cell of every frame.
template <bool ALIGNED>
__global__ void __launch_bounds__(kFastThreads)
    fast_cells_kernel(const OrbGeometry* __restrict__ geo, ImgBatch in0, int ini_th, int min_th, int first_pass,
                      const uint32_t* __restrict__ list, const uint32_t* __restrict__ n_items, uint32_t n_all,
                      uint32_t* __restrict__ pool, int pool_cap, uint32_t* __restrict__ pool_count,
                      uint2* __restrict__ cell_tab, uint32_t* __restrict__ status) {
  const uint32_t n = list ? *n_items : n_all;
  const uint32_t cells = (uint32_t)geo->total_cells;
  for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
    const uint32_t item = list ? list[i] : i;
    const int b = (int)(item / cells), cell = (int)(item - (uint32_t)b * cells);
    fast_cell_body<ALIGNED>(geo, in0, ini_th, min_th, first_pass, cell, b, pool, pool_cap, pool_count, cell_tab,
                            status);
    __syncthreads();  // the shared tiles are reused by the next item
  }
}

// ---------------------------------------------------------------------------------------------
// Dense path (thresholds <= 127).
//
// score kernel: a CTA owns a 112 x 30 tile of a level's detection window.  The 144 x 38 pixel box the tile's
// scores depend on is staged in shared memory by one TMA tensor copy (out-of-image bytes arrive as zeros); the
// compass reject runs on 4 adjacent pixels per thread with byte-SIMD (VABSDIFF4 + a carry-free ">" on packed
// bytes); the survivors are compacted, get exact scores two at a time on u16x2 lanes, and the per-cell NMS is
// applied (a neighbour in another cell counts as 0, exactly as in the cell sub-images of the reference).
// Output: the level's score map, u8 = score of a corner at iniThFAST that survives the NMS, else 0.  The map
// lives in the level's blur buffer, which is not written until the blur stage that follows FAST and the octree.
// collect kernel: a warp per cell turns the non-zero bytes of the cell's interior into the raster-ordered
// candidate slice of the pool (popc prefix sums); a cell without any goes to the fallback list, which the
// per-cell kernel above redoes at minThFAST (:812-816).
// ---------------------------------------------------------------------------------------------
constexpr int kTW = 112, kTH = 30;          // core tile (pixels whose map bytes this CTA writes); the TMA box must
                                            // start on a 16-byte boundary of the row, hence a multiple of 16
constexpr int kSW = 128, kSH = 32;          // scored region: core + 4 columns / 1 row of rim on each side (120 columns used)
constexpr int kPR = kSH + 6;                // pixel rows staged
constexpr int kPWords = 36, kPP = 144;      // pixel words per row used / row pitch in bytes (= TMA box width)
constexpr int kScoreThreads = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool ALIGNED>
__global__ void __launch_bounds__(kScoreThreads)
    fast_score_tiles_kernel(const OrbGeometry* __restrict__ geo, ImgBatch in0, const __grid_constant__ FastMaps maps,
                            int ini_th) {
  __shared__ __align__(128) uint8_t s_px[kPR][kPP];
  __shared__ __align__(16) uint8_t s_sc[kSH][kSW];
  __shared__ uint16_t s_list[kSH * kSW];
  __shared__ uint16_t s_list2[kTH * kTW];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ int s_n, s_n2;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, b = blockIdx.y;
  const uint32_t te = __ldg(geo->fast_tab + blockIdx.x);
  const int lvl = (int)(te >> 24), ty = (int)((te >> 12) & 0xFFFu), tx = (int)(te & 0xFFFu);
  const int cx0 = kMinBorder + kTW * tx, cy0 = kMinBorder + 3 + kTH * ty;
  const int gx0 = cx0 - 16, gy0 = cy0 - 4;  // level coordinates of shared pixel (0, 0)
  const bool tma = (maps.valid >> lvl) & 1u;
  if (tma && tid == 0) {
    const uint32_t bar = smem_u32(&s_bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kPR * kPP) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(&s_px[0][0])), "l"(reinterpret_cast<const void*>(&maps.m[lvl][0])), "r"(gx0), "r"(gy0), "r"(b),
        "r"(bar)
        : "memory");
  }
  const CellGrid g = geo->grid[lvl];
  const int x_lo = kMinBorder + 3, x_hi = g.max_bx - 3, y_lo = kMinBorder + 3, y_hi = g.max_by - 3;
  uint8_t* __restrict__ map = geo->blur[lvl].ptr + (size_t)b * geo->blur[lvl].frame_stride;
  const int mp = geo->blur[lvl].pitch;

  if (!tma) {  // the layout has no descriptor (caller's image with odd strides): plain loads
    const uint8_t* __restrict__ img;
    int pitch, h, w;
    if (lvl == 0) {
      img = in0.ptr + (size_t)b * in0.frame_stride;
      pitch = in0.pitch; w = in0.w; h = in0.h;
    } else {
      img = geo->level[lvl].ptr + (size_t)b * geo->level[lvl].frame_stride;
      pitch = geo->level[lvl].pitch; w = geo->level[lvl].w; h = geo->level[lvl].h;
    }
    for (int i = tid; i < kPR * kPWords; i += kScoreThreads) {
      const int r = i / kPWords, wc = i - r * kPWords;
      const int y = gy0 + r, x = gx0 + 4 * wc;
      uint32_t v = 0;
      if (y < h) {
        const uint8_t* p = img + (size_t)y * pitch + x;
        if (ALIGNED) {
          if (x + 4 <= pitch) v = __ldg(reinterpret_cast<const uint32_t*>(p));
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (x + k < w) v |= (uint32_t)__ldg(p + k) << (8 * k);
        }
      }
      reinterpret_cast<uint32_t*>(&s_px[r][0])[wc] = v;
    }
  }
  reinterpret_cast<uint4*>(&s_sc[0][0])[tid] = make_uint4(0u, 0u, 0u, 0u);  // 32 x 128 B = 256 x 16 B
  if (tid == 0) { s_n = 0; s_n2 = 0; }
  // the core of the map starts as zeros; the corners that survive are stored after the barriers below
  for (int i = tid; i < kTH * (kTW / 8); i += kScoreThreads) {
    const int r = i / (kTW / 8), c8 = i - r * (kTW / 8);
    const int y = cy0 + r, x = cx0 + 8 * c8;
    if (y < y_hi && x < x_hi) *reinterpret_cast<uint2*>(map + (size_t)y * mp + x) = make_uint2(0u, 0u);
  }
  __syncthreads();
  if (tma) {
    const uint32_t bar = smem_u32(&s_bar);
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}"
          : "=r"(done) : "r"(bar) : "memory");
    }
  }

  // ---- compass reject, 4 pixels per thread-row: a corner needs (up or down) and (left or right) further than
  // t from the centre.  d > t on packed bytes: ((d & 0x7f) + (127 - t)) | d has bit 7 set.
  {
    const int x = cx0 - 4 + 4 * lane;
    uint32_t vm = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (x + k >= x_lo && x + k < x_hi && lane < (kTW + 8) / 4) vm |= 0x80u << (8 * k);
    const uint32_t K = (uint32_t)(127 - ini_th) * 0x01010101u;
    uint32_t M = 0;  // bit 8 j + k: pixel j of this thread's group in its row k survives
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = wid * 4 + k, y = cy0 - 1 + r;
      const uint32_t* rc = reinterpret_cast<const uint32_t*>(&s_px[r + 3][0]) + 3 + lane;
      const uint32_t wc = rc[0], wl = rc[-1], wr = rc[1];
      const uint32_t wu = rc[-3 * (kPP / 4)], wd = rc[3 * (kPP / 4)];
      const uint32_t left = __funnelshift_r(wl, wc, 8), right = __funnelshift_r(wc, wr, 24);  // x-3 / x+3
      const uint32_t du = __vabsdiffu4(wc, wu), dd = __vabsdiffu4(wc, wd);
      const uint32_t dl = __vabsdiffu4(wc, left), dr = __vabsdiffu4(wc, right);
      const uint32_t ud = ((du & 0x7f7f7f7fu) + K) | ((dd & 0x7f7f7f7fu) + K) | du | dd;
      const uint32_t lr = ((dl & 0x7f7f7f7fu) + K) | ((dr & 0x7f7f7f7fu) + K) | dl | dr;
      const uint32_t m = (y >= y_lo && y < y_hi) ? (ud & lr & vm) : 0u;
      M |= m >> (7 - k);
    }
    const int cnt = __popc(M);
    int inc = cnt;
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
      const int a = __shfl_up_sync(0xffffffffu, inc, dlt);
      if (lane >= dlt) inc += a;
    }
    int base = 0;
    if (lane == 31 && inc) base = atomicAdd(&s_n, inc);
    int pos = __shfl_sync(0xffffffffu, base, 31) + inc - cnt;
    const uint32_t e0 = ((uint32_t)(wid * 4) << 7) | (uint32_t)(4 * lane);  // entry = scored row << 7 | scored column
    while (M) {
      const uint32_t bit = (uint32_t)__ffs(M) - 1u;
      M &= M - 1;
      s_list[pos++] = (uint16_t)(e0 + ((bit & 7u) << 7) + (bit >> 3));
    }
  }
  __syncthreads();

  // ---- exact scores of the survivors, two per thread (scored (r, c) = shared pixel (r + 3, c + 12)) -------------
  const int n = s_n;
  for (int k = 2 * tid; k < n; k += 2 * kScoreThreads) {
    const int ea = s_list[k], eb = s_list[min(k + 1, n - 1)];
    const int ra = ea >> 7, ca = ea & 127, rb = eb >> 7, cb = eb & 127;
    const unsigned sc = fast_score_pair<kPP>(&s_px[ra + 3][ca + 12], &s_px[rb + 3][cb + 12]);
    const unsigned sa = sc & 0xFFFFu, sb = sc >> 16;
    if (sa >= (unsigned)ini_th) {
      s_sc[ra][ca] = (uint8_t)sa;
      if (ra >= 1 && ra <= kTH && ca >= 4 && ca < 4 + kTW) s_list2[atomicAdd(&s_n2, 1)] = (uint16_t)ea;
    }
    if (k + 1 < n && sb >= (unsigned)ini_th) {
      s_sc[rb][cb] = (uint8_t)sb;
      if (rb >= 1 && rb <= kTH && cb >= 4 && cb < 4 + kTW) s_list2[atomicAdd(&s_n2, 1)] = (uint16_t)eb;
    }
  }
  __syncthreads();

  // ---- NMS inside the cell the pixel belongs to ----------------------------------------------------------------
  const int n2 = s_n2;
  const float rcp_wc = __frcp_rn((float)g.w_cell), rcp_hc = __frcp_rn((float)g.h_cell);
  for (int k = tid; k < n2; k += kScoreThreads) {
    const int e = s_list2[k], r = e >> 7, c = e & 127;
    const int x = cx0 - 4 + c, y = cy0 - 1 + r;
    // position inside the cell: (x - x_lo) mod w_cell without an integer divide (quotient of values < 2^13 by
    // a divisor <= 60: the true quotient + 0.5 / d is far from an integer compared with the fp32 error)
    const int lx = (x - x_lo) - (int)(((float)(x - x_lo) + 0.5f) * rcp_wc) * g.w_cell;
    const int ly = (y - y_lo) - (int)(((float)(y - y_lo) + 0.5f) * rcp_hc) * g.h_cell;
    const unsigned ml = lx != 0 ? 0xFFu : 0u, mr = lx != g.w_cell - 1 ? 0xFFu : 0u;
    const unsigned mu = ly != 0 ? 0xFFu : 0u, md = ly != g.h_cell - 1 ? 0xFFu : 0u;
    const uint8_t* r0 = &s_sc[r - 1][c];
    const uint8_t* r1 = &s_sc[r][c];
    const uint8_t* r2 = &s_sc[r + 1][c];
    const unsigned up = max(max(r0[-1] & ml, (unsigned)r0[0]), r0[1] & mr) & mu;
    const unsigned dn = max(max(r2[-1] & ml, (unsigned)r2[0]), r2[1] & mr) & md;
    const unsigned mid = max(r1[-1] & ml, r1[1] & mr);
    const unsigned sc = r1[0];
    if (sc > max(max(up, dn), mid)) map[(size_t)y * mp + x] = (uint8_t)sc;
  }
}

constexpr int kCollectWarps = 8;
constexpr int kCollectCap = (kInteriorMax * kInteriorMax + 3) / 4;  // NMS survivors of a cell: at most one per 2x2 block

// A warp per cell, one pass: the interior is read as whole words, `rpi` rows per warp step (2 when a row has at
// most 16 words; a lane always holds the same word column, so its byte mask is fixed), five steps in flight.
// Corners go to a per-warp shared list in raster order (steps ascending, lanes ascending = row then word, bytes
// ascending); the list is copied to the cell's slice of the pool once its length is known.
__global__ void __launch_bounds__(kCollectWarps * 32)
    fast_collect_kernel(const OrbGeometry* __restrict__ geo, int redo_empty, uint32_t* __restrict__ pool, int pool_cap,
                        uint32_t* __restrict__ pool_count, uint2* __restrict__ cell_tab,
                        uint32_t* __restrict__ fb_list, uint32_t* __restrict__ fb_count,
                        uint32_t* __restrict__ status) {
  __shared__ uint32_t s_out[kCollectWarps][kCollectCap];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, b = blockIdx.y;
  const int cell = blockIdx.x * kCollectWarps + wid;
  const int cells = geo->total_cells;
  if (cell >= cells) return;
  const uint32_t ce = __ldg(geo->fast_tab + geo->n_tiles + cell);
  const int lvl = (int)(ce >> 24), ci = (int)((ce >> 12) & 0xFFFu), cj = (int)(ce & 0xFFFu);
  const CellGrid g = geo->grid[lvl];
  uint2* tab = cell_tab + (size_t)b * cells + cell;
  // interior of the cell (:789-806): level x in [xa, xa + iw), y in [ya, ya + ih)
  const int xa = kMinBorder + 3 + cj * g.w_cell, ya = kMinBorder + 3 + ci * g.h_cell;
  const int iw = min(g.w_cell, g.max_bx - 3 - xa), ih = min(g.h_cell, g.max_by - 3 - ya);
  if (iw <= 0 || ih <= 0) {
    if (lane == 0) *tab = make_uint2(0u, 0u);
    return;
  }
  const int mp = geo->blur[lvl].pitch;
  const int w0 = xa >> 2, nw = ((xa + iw + 3) >> 2) - w0;
  const int two = nw <= 16, wi = two ? (lane & 15) : lane, sub = two ? (lane >> 4) : 0, rpi = two ? 2 : 1;
  // bytes of this lane's word column that belong to the interior
  uint32_t lmask = wi < nw ? 0xFFFFFFFFu : 0u;
  if (wi == 0) lmask &= 0xFFFFFFFFu << (8 * (xa & 3));
  if (wi == nw - 1 && ((xa + iw) & 3)) lmask &= ~(0xFFFFFFFFu << (8 * ((xa + iw) & 3)));
  const uint32_t* __restrict__ col = reinterpret_cast<const uint32_t*>(
      geo->blur[lvl].ptr + (size_t)b * geo->blur[lvl].frame_stride + (size_t)(ya + sub) * mp) + w0 + min(wi, nw - 1);
  const int stride = rpi * (mp >> 2), nit = (ih + rpi - 1) / rpi;
  const uint32_t xy0 = pack_cand(4 * (w0 + wi) - kMinBorder, ya + sub - kMinBorder, 0);  // level coordinate - minBorder
  uint32_t* list = s_out[wid];
  int run = 0;
  constexpr int kFlight = 5;
  for (int it0 = 0; it0 < nit; it0 += kFlight) {
    uint32_t v[kFlight];
#pragma unroll
    for (int k = 0; k < kFlight; ++k) {
      const int it = it0 + k;
      v[k] = (it * rpi + sub < ih) ? __ldg(col + (size_t)it * stride) : 0u;
    }
#pragma unroll
    for (int k = 0; k < kFlight; ++k) {
      const uint32_t vv = v[k] & lmask;
      uint32_t nz = (((vv & 0x7f7f7f7fu) + 0x7f7f7f7fu) | vv) & 0x80808080u;
      if (!__any_sync(0xffffffffu, nz != 0u)) continue;
      const int c = __popc(nz);
      int inc = c;
#pragma unroll
      for (int dlt = 1; dlt < 32; dlt <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, inc, dlt);
        if (lane >= dlt) inc += a;
      }
      int pos = run + inc - c;
      run += __shfl_sync(0xffffffffu, inc, 31);
      const uint32_t xy = xy0 + ((uint32_t)((it0 + k) * rpi) << 8);
      while (nz) {
        const int j = (__ffs(nz) - 1) >> 3;
        nz &= nz - 1;
        if (pos < kCollectCap) list[pos] = xy + ((uint32_t)j << 20) + ((vv >> (8 * j)) & 0xFFu);
        ++pos;
      }
    }
  }
  const int total = run;
  if (total == 0) {
    if (lane == 0) {
      if (redo_empty) {
        *tab = make_uint2(0xFFFFFFFFu, 0u);
        fb_list[atomicAdd(fb_count, 1u)] = (uint32_t)b * (uint32_t)cells + (uint32_t)cell;
      } else {
        *tab = make_uint2(0u, 0u);
      }
    }
    return;
  }
  uint32_t base = 0;
  if (lane == 0) {
    base = atomicAdd(pool_count + b, (uint32_t)total);
    *tab = make_uint2(base, (uint32_t)total);
    if (base + total > (uint32_t)pool_cap) atomicOr(status, kStatCandOverflow);
    if (total > kCollectCap) atomicOr(status, kStatNodeOverflow);  // cannot happen: NMS survivors are never adjacent
  }
  base = __shfl_sync(0xffffffffu, base, 0);
  if (base + total > (uint32_t)pool_cap || total > kCollectCap) return;
  __syncwarp();
  uint32_t* out = pool + (size_t)b * pool_cap + base;
  for (int i = lane; i < total; i += 32) out[i] = list[i];
}

#undef PSL_DIV

int fast_tile_count(const OrbGeometry& geo) {
  int n = 0;
  for (int l = 0; l < geo.nlevels; ++l) {
    const CellGrid& cg = geo.grid[l];
    const int ntx = (cg.max_bx - 3 - kMinBorder + kTW - 1) / kTW, nty = (cg.max_by - 3 - (kMinBorder + 3) + kTH - 1) / kTH;
    n += (ntx > 0 && nty > 0) ? ntx * nty : 0;
  }
  return n;
}

void fast_build_tab(const OrbGeometry& geo, uint32_t* tab) {
  int n = 0;
  for (int l = 0; l < geo.nlevels; ++l) {
    const CellGrid& cg = geo.grid[l];
    const int ntx = (cg.max_bx - 3 - kMinBorder + kTW - 1) / kTW, nty = (cg.max_by - 3 - (kMinBorder + 3) + kTH - 1) / kTH;
    for (int ty = 0; ty < nty; ++ty)
      for (int tx = 0; tx < ntx; ++tx) tab[n++] = ((uint32_t)l << 24) | ((uint32_t)ty << 12) | (uint32_t)tx;
  }
  for (int l = 0; l < geo.nlevels; ++l) {
    const CellGrid& cg = geo.grid[l];
    for (int ci = 0; ci < cg.n_rows; ++ci)
      for (int cj = 0; cj < cg.n_cols; ++cj) tab[n++] = ((uint32_t)l << 24) | ((uint32_t)ci << 12) | (uint32_t)cj;
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool fast_encode_map(FastMaps& maps, int level, const void* ptr, int w, int h, int pitch, int64_t frame_stride,
                     int frames) {
  static_assert(sizeof(CUtensorMap) == 128, "FastMaps slot size");
  static EncodeTiledFn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      fn = nullptr;
    return reinterpret_cast<EncodeTiledFn>(fn);
  }();
  maps.valid &= ~(1u << level);
  if (!encode || ((uintptr_t)ptr & 15) || (pitch & 15) || (frame_stride & 15) || frames < 1) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)frames};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frame_stride};
  const cuuint32_t box[3] = {(cuuint32_t)kPP, (cuuint32_t)kPR, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = encode(reinterpret_cast<CUtensorMap*>(&maps.m[level][0]), CU_TENSOR_MAP_DATA_TYPE_UINT8, 3,
                            const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  maps.valid |= 1u << level;
  return true;
}

void launch_fast_cells(const OrbGeometry* d_geo, const OrbGeometry& geo, ImgBatch in0, int ini_th, int min_th,
                       uint32_t* pool, int pool_cap, uint32_t* pool_count, uint2* cell_tab, uint32_t* fb_list,
                       uint32_t* fb_count, FastMaps& maps, uint32_t* status, int B, cudaStream_t st) {
  // levels >= 1 are ours (128-byte pitch); level 0 is the caller's image
  const bool aligned = ((uintptr_t)in0.ptr & 3) == 0 && (in0.pitch & 3) == 0 && (in0.frame_stride & 3) == 0;
  const uint32_t n_all = (uint32_t)geo.total_cells * (uint32_t)B;
  const int list_grid = 148 * 8;
  if (ini_th > 127) {  // outside the byte-SIMD compare: the per-cell kernel does both passes for every cell
    const int grid = (int)(n_all < 65535u * 16u ? n_all : 65535u * 16u);
    if (aligned)
      fast_cells_kernel<true><<<grid, kFastThreads, 0, st>>>(d_geo, in0, ini_th, min_th, 0, nullptr, nullptr, n_all, pool,
                                                             pool_cap, pool_count, cell_tab, status);
    else
      fast_cells_kernel<false><<<grid, kFastThreads, 0, st>>>(d_geo, in0, ini_th, min_th, 0, nullptr, nullptr, n_all,
                                                              pool, pool_cap, pool_count, cell_tab, status);
    return;
  }
  fast_encode_map(maps, 0, in0.ptr, in0.w, in0.h, in0.pitch, in0.frame_stride, B);
  const int redo = min_th < ini_th;
  if (redo) cudaMemsetAsync(fb_count, 0, sizeof(uint32_t), st);
  dim3 tgrid(geo.n_tiles, B);
  if (aligned) fast_score_tiles_kernel<true><<<tgrid, kScoreThreads, 0, st>>>(d_geo, in0, maps, ini_th);
  else fast_score_tiles_kernel<false><<<tgrid, kScoreThreads, 0, st>>>(d_geo, in0, maps, ini_th);
  dim3 cgrid((geo.total_cells + kCollectWarps - 1) / kCollectWarps, B);
  fast_collect_kernel<<<cgrid, kCollectWarps * 32, 0, st>>>(d_geo, redo, pool, pool_cap, pool_count, cell_tab, fb_list,
                                                            fb_count, status);
  if (redo) {
    if (aligned)
      fast_cells_kernel<true><<<list_grid, kFastThreads, 0, st>>>(d_geo, in0, ini_th, min_th, 1, fb_list, fb_count, 0u,
                                                                  pool, pool_cap, pool_count, cell_tab, status);
    else
      fast_cells_kernel<false><<<list_grid, kFastThreads, 0, st>>>(d_geo, in0, ini_th, min_th, 1, fb_list, fb_count, 0u,
                                                                   pool, pool_cap, pool_count, cell_tab, status);
  }
}

}  // namespace psl
