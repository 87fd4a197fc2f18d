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
    // a cell that does not fit keeps no candidates and says so in its table entry: the octree gathers through the
    // table, and a level whose total still fits would otherwise read past the pool
    const bool fits = base + total <= (uint32_t)pool_cap;
    *tab = make_uint2(fits ? base : 0u, fits ? (uint32_t)total : 0u);
    if (!fits) { atomicOr(status, kStatCandOverflow); atomicMax(status + 1, (uint32_t)b + 1u); }
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

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#undef PSL_DIV

// ---------------------------------------------------------------------------------------------
// Fused dense path (thresholds <= 127): one kernel from pixels to the raster-ordered candidate slices.
//
// A CTA owns up to four neighbouring cells of one cell row.  Cell interiors tile the detection window without overlap
// and the reference's non-maximum suppression never looks across a cell seam (every cv::FAST call sees its own
// sub-image), so a tile made of whole cells needs no halo of scores at all — only the 3-pixel ring FAST itself reads.
//   1. the (h_cell + 6) x 160 pixel box arrives by ONE TMA tensor copy (per-level descriptor; out-of-image bytes = 0);
//   2. compass reject, 4 pixels per thread-row with byte SIMD; survivors compacted into a shared list;
//   3. exact scores two at a time on u16x2 lanes -> shared score tile + list of corners at iniThFAST;
//   4. NMS over the corners only (neighbours in another cell or outside the interior count as 0) -> keep bitmap;
//   5. a warp per cell turns its rows of the bitmap into the cell's slice of the pool: popc prefix over the rows, one
//      atomicAdd for the slice, entries written in raster order.  A cell without any corner goes to the fallback list
//      that the per-cell kernel redoes at minThFAST (:812-816).
// No score map in HBM, no second kernel reading it back.
// ---------------------------------------------------------------------------------------------
constexpr int kRW = 128;                  // scored columns per tile (4 per lane)
constexpr int kRP = 160;                  // pixel row pitch = TMA box width
constexpr int kRHmax = kInteriorMax;      // interior rows of a cell
constexpr int kRowsThreads = 128;   // (256: 10.6, 128: 7.9, 64: 8.5 ms per 4096 frames)

__host__ __device__ inline int fast_group_cells(int w_cell) { return w_cell >= 63 ? 1 : (kRW - 3) / w_cell > 4 ? 4 : (kRW - 3) / w_cell; }

template <bool ALIGNED>
__global__ void __launch_bounds__(kRowsThreads, 1024 / kRowsThreads)
    fast_rows_kernel(const OrbGeometry* __restrict__ geo, ImgBatch in0, const __grid_constant__ FastMaps maps, int ini_th,
                     int redo_empty, uint32_t* __restrict__ pool, int pool_cap, uint32_t* __restrict__ pool_count,
                     uint2* __restrict__ cell_tab, uint32_t* __restrict__ fb_list, uint32_t* __restrict__ fb_count,
                     uint32_t* __restrict__ status, int rh) {
  // tiles sized by the tallest cell of this geometry (rh rows, 30..37 for the usual level sizes; kRHmax is the bound of
  // the cell rule): fewer bytes per CTA, more CTAs per SM
  extern __shared__ __align__(128) unsigned char rows_smem[];
  uint8_t (*s_px)[kRP] = reinterpret_cast<uint8_t (*)[kRP]>(rows_smem);                                   // [rh + 6][kRP]
  uint8_t (*s_sc)[kRW] = reinterpret_cast<uint8_t (*)[kRW]>(rows_smem + (size_t)(rh + 6) * kRP);           // [rh][kRW]
  uint32_t (*s_keep)[kRW / 32] = reinterpret_cast<uint32_t (*)[kRW / 32]>(&s_sc[rh][0]);                   // [rh][kRW / 32]
  uint16_t* s_list = reinterpret_cast<uint16_t*>(&s_keep[rh][0]);                                          // [rh * kRW]
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ int s_n, s_n2;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, b = blockIdx.y;
  const uint32_t te = __ldg(geo->fast_tab + blockIdx.x);
  const int lvl = (int)(te >> 24), ci = (int)((te >> 14) & 0x3FFu), cj0 = (int)((te >> 4) & 0x3FFu), nc = (int)(te & 0xFu);
  const CellGrid g = geo->grid[lvl];
  // interiors: cell cj covers level x in [19 + cj * w_cell, ...) clipped at max_bx - 3; rows [ya, ya + ih)
  const int xa0 = kMinBorder + 3 + cj0 * g.w_cell, ya = kMinBorder + 3 + ci * g.h_cell;
  const int Wt = min(nc * g.w_cell, g.max_bx - 3 - xa0), ih = min(g.h_cell, g.max_by - 3 - ya);
  const int X0 = (xa0 - 7) & ~15, Y0 = ya - 3;      // level coordinates of shared pixel (0, 0); X0 is the TMA box origin
  const int c0 = xa0 - X0, c_lo = c0 & ~3;          // first interior column, first scored column (>= 4, a word boundary)
  const int cofs = c0 - c_lo;                       // interior starts at scored column cofs (0..3)
  const bool tma = (maps.valid >> lvl) & 1u;
  const int rows_px = g.h_cell + 6;
  if (tma && tid == 0) {
    const uint32_t bar = smem_u32(&s_bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(rows_px * kRP) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(&s_px[0][0])), "l"(reinterpret_cast<const void*>(&maps.m[lvl][0])), "r"(X0), "r"(Y0), "r"(b),
        "r"(bar)
        : "memory");
  }
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
    for (int i = tid; i < rows_px * (kRP / 4); i += kRowsThreads) {
      const int r = i / (kRP / 4), wc = i - r * (kRP / 4);
      const int y = Y0 + r, x = X0 + 4 * wc;
      uint32_t v = 0;
      if (y >= 0 && y < h) {
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
  for (int i = tid; i < ih * (kRW / 16); i += kRowsThreads) reinterpret_cast<uint4*>(&s_sc[0][0])[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < ih * (kRW / 32); i += kRowsThreads) (&s_keep[0][0])[i] = 0u;
  if (tid == 0) { s_n = 0; s_n2 = 0; }
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

  // ---- compass reject, 4 pixels per thread-row: a corner needs (up or down) and (left or right) further than t from
  // the centre.  d > t on packed bytes: ((d & 0x7f) + (127 - t)) | d has bit 7 set.
  {
    uint32_t vm = 0;   // bytes of this lane's word that are interior columns
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int cc = 4 * lane + k - cofs;
      if (cc >= 0 && cc < Wt) vm |= 0x80u << (8 * k);
    }
    const uint32_t K = (uint32_t)(127 - ini_th) * 0x01010101u;
    constexpr int kRT = 4;   // rows per thread and pass (5 measured slower: 5.32 against 5.19 ms per 2048 frames)
    for (int r0 = 0; r0 < ih; r0 += (kRowsThreads / 32) * kRT) {
      uint32_t M = 0;  // bit 8 j + k: pixel j of this thread's group in its row k survives
#pragma unroll
      for (int k = 0; k < kRT; ++k) {
        const int r = r0 + wid * kRT + k;
        if (r < ih && vm) {
          const uint32_t* rc = reinterpret_cast<const uint32_t*>(&s_px[r + 3][0]) + (c_lo >> 2) + lane;
          const uint32_t wc = rc[0], wl = rc[-1], wr = rc[1];
          const uint32_t wu = rc[-3 * (kRP / 4)], wd = rc[3 * (kRP / 4)];
          const uint32_t left = __funnelshift_r(wl, wc, 8), right = __funnelshift_r(wc, wr, 24);  // x-3 / x+3
          const uint32_t du = __vabsdiffu4(wc, wu), dd = __vabsdiffu4(wc, wd);
          const uint32_t dl = __vabsdiffu4(wc, left), dr = __vabsdiffu4(wc, right);
          const uint32_t ud = ((du & 0x7f7f7f7fu) + K) | ((dd & 0x7f7f7f7fu) + K) | du | dd;
          const uint32_t lr = ((dl & 0x7f7f7f7fu) + K) | ((dr & 0x7f7f7f7fu) + K) | dl | dr;
          M |= (ud & lr & vm) >> (7 - k);
        }
      }
      // the order of the survivors does not matter (scores are per pixel, the NMS and the output work off bitmaps): one
      // shared atomic per thread that has any
      int pos = M ? atomicAdd(&s_n, __popc(M)) : 0;
      const uint32_t e0 = ((uint32_t)(r0 + wid * kRT) << 7) | (uint32_t)(4 * lane);  // entry = row << 7 | scored column
      while (M) {
        const uint32_t bit = (uint32_t)__ffs(M) - 1u;
        M &= M - 1;
        s_list[pos++] = (uint16_t)(e0 + ((bit & 7u) << 7) + (bit >> 3));
      }
    }
  }
  __syncthreads();

  // ---- exact scores of the survivors, two per thread (scored (r, c) = shared pixel (r + 3, c_lo + c)) ---------------
  const int n = s_n;
  const int kHalf = rh * kRW / 2;
  const bool two_lists = n <= kHalf;
  for (int k = 2 * tid; k < n; k += 2 * kRowsThreads) {
    const int ea = s_list[k], eb = s_list[min(k + 1, n - 1)];
    const int ra = ea >> 7, ca = ea & 127, rb = eb >> 7, cb = eb & 127;
    const unsigned sc = fast_score_pair<kRP>(&s_px[ra + 3][c_lo + ca], &s_px[rb + 3][c_lo + cb]);
    const unsigned sa = sc & 0xFFFFu, sb = sc >> 16;
    const bool ka = sa >= (unsigned)ini_th, kb = k + 1 < n && sb >= (unsigned)ini_th;
    if (ka) s_sc[ra][ca] = (uint8_t)sa;
    if (kb) s_sc[rb][cb] = (uint8_t)sb;
    if (two_lists && (ka || kb)) {   // the corners, for the NMS pass (upper half of the list array, free when n is small)
      const int q = atomicAdd(&s_n2, (int)ka + (int)kb);
      if (ka) s_list[kHalf + q] = (uint16_t)ea;
      if (kb) s_list[kHalf + q + (int)ka] = (uint16_t)eb;
    }
  }
  __syncthreads();

  // ---- NMS inside the cell the corner belongs to -> keep bitmap -------------------------------------------------------
  const float rcp_wc = __frcp_rn((float)g.w_cell);
  const int n_nms = two_lists ? s_n2 : n;
  const uint16_t* nms_list = two_lists ? s_list + kHalf : s_list;
  for (int k = tid; k < n_nms; k += kRowsThreads) {
    const int e = nms_list[k], r = e >> 7, c = e & 127;
    if (!two_lists && s_sc[r][c] == 0) continue;
    const int xt = c - cofs;                                  // column inside the tile's interior
    // xt mod w_cell without an integer divide (values < 2^8, divisor <= 60: the quotient + 0.5 / d is far from an
    // integer compared with the fp32 error)
    const int cell = (int)(((float)xt + 0.5f) * rcp_wc), lx = xt - cell * g.w_cell;
    const int iwk = min(g.w_cell, Wt - cell * g.w_cell);
    const unsigned ml = lx != 0 ? 0xFFu : 0u, mr = lx != iwk - 1 ? 0xFFu : 0u;
    const unsigned mu = r != 0 ? 0xFFu : 0u, md = r != ih - 1 ? 0xFFu : 0u;
    const uint8_t* r0 = &s_sc[max(r - 1, 0)][c];
    const uint8_t* r1 = &s_sc[r][c];
    const uint8_t* r2 = &s_sc[min(r + 1, ih - 1)][c];
    const int cm = max(c - 1, 0) - c, cp = min(c + 1, kRW - 1) - c;
    const unsigned up = max(max(r0[cm] & ml, (unsigned)r0[0]), r0[cp] & mr) & mu;
    const unsigned dn = max(max(r2[cm] & ml, (unsigned)r2[0]), r2[cp] & mr) & md;
    const unsigned mid = max(r1[cm] & ml, r1[cp] & mr);
    if ((unsigned)r1[0] > max(max(up, dn), mid)) atomicOr(&s_keep[r][c >> 5], 1u << (c & 31));
  }
  __syncthreads();

  // ---- a warp per cell: the cell's rows of the bitmap -> its slice of the pool, raster order -------------------------
  for (int cw = wid; cw < nc; cw += kRowsThreads / 32) {
    const int cell = g.first_cell + ci * g.n_cols + cj0 + cw;
    const int cells = geo->total_cells;
    uint2* tab = cell_tab + (size_t)b * cells + cell;
    const int cs = cofs + cw * g.w_cell;                     // first scored column of the cell
    const int iwk = min(g.w_cell, Wt - cw * g.w_cell);
    if (iwk <= 0 || ih <= 0) {
      if (lane == 0) *tab = make_uint2(0u, 0u);
      continue;
    }
    // bits of row r that belong to the cell, as a 64-bit mask starting at the cell's first column (iwk <= 60)
    auto row_bits = [&](int r) -> unsigned long long {
      const int w0 = cs >> 5, sh = cs & 31;
      const unsigned long long lo = (unsigned long long)s_keep[r][w0] | ((unsigned long long)(w0 + 1 < kRW / 32 ? s_keep[r][w0 + 1] : 0u) << 32);
      unsigned long long v = lo >> sh;
      if (sh && w0 + 2 < kRW / 32) v |= (unsigned long long)s_keep[r][w0 + 2] << (64 - sh);
      return v & ((1ull << iwk) - 1ull);
    };
    const unsigned long long m0 = lane < ih ? row_bits(lane) : 0ull;
    const unsigned long long m1 = lane + 32 < ih ? row_bits(lane + 32) : 0ull;
    const int c0n = __popcll(m0), c1n = __popcll(m1);
    int i0 = c0n, i1 = c1n;
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
      const int a0 = __shfl_up_sync(0xffffffffu, i0, dlt), a1 = __shfl_up_sync(0xffffffffu, i1, dlt);
      if (lane >= dlt) { i0 += a0; i1 += a1; }
    }
    const int t0 = __shfl_sync(0xffffffffu, i0, 31), total = t0 + __shfl_sync(0xffffffffu, i1, 31);
    if (total == 0) {
      if (lane == 0) {
        if (redo_empty) {
          *tab = make_uint2(0xFFFFFFFFu, 0u);
          fb_list[atomicAdd(fb_count, 1u)] = (uint32_t)b * (uint32_t)cells + (uint32_t)cell;
        } else {
          *tab = make_uint2(0u, 0u);
        }
      }
      continue;
    }
    uint32_t base = 0;
    if (lane == 0) {
      base = atomicAdd(pool_count + b, (uint32_t)total);
      const bool fits = base + total <= (uint32_t)pool_cap;   // (see fast_cell_body)
      *tab = make_uint2(fits ? base : 0u, fits ? (uint32_t)total : 0u);
      if (!fits) { atomicOr(status, kStatCandOverflow); atomicMax(status + 1, (uint32_t)b + 1u); }
    }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base + total > (uint32_t)pool_cap) continue;
    uint32_t* out = pool + (size_t)b * pool_cap + base;
    // level coordinate - minBorder of the cell's first interior pixel
    const uint32_t x_rel = (uint32_t)(xa0 + cw * g.w_cell - kMinBorder), y_rel = (uint32_t)(ya - kMinBorder);
    auto emit = [&](unsigned long long m, int r, int pos) {
      while (m) {
        const int x = __ffsll((long long)m) - 1;
        m &= m - 1;
        out[pos++] = pack_cand((int)x_rel + x, (int)y_rel + r, s_sc[r][cs + x]);
      }
    };
    emit(m0, lane, i0 - c0n);
    emit(m1, lane + 32, t0 + i1 - c1n);
  }
}

int fast_tile_count(const OrbGeometry& geo) {
  int n = 0;
  for (int l = 0; l < geo.nlevels; ++l) {
    const CellGrid& cg = geo.grid[l];
    const int G = fast_group_cells(cg.w_cell);
    n += cg.n_rows * ((cg.n_cols + G - 1) / G);
  }
  return n;
}

void fast_build_tab(const OrbGeometry& geo, uint32_t* tab) {
  int n = 0;
  for (int l = 0; l < geo.nlevels; ++l) {
    const CellGrid& cg = geo.grid[l];
    const int G = fast_group_cells(cg.w_cell);
    for (int ci = 0; ci < cg.n_rows; ++ci)
      for (int cj = 0; cj < cg.n_cols; cj += G)
        tab[n++] = ((uint32_t)l << 24) | ((uint32_t)ci << 14) | ((uint32_t)cj << 4) | (uint32_t)std::min(G, cg.n_cols - cj);
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
                     int frames, int box_rows) {
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
  if (!encode || ((uintptr_t)ptr & 15) || (pitch & 15) || (frame_stride & 15) || frames < 1 || box_rows < 1 ||
      box_rows > kRHmax + 6)
    return false;
  const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)frames};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frame_stride};
  const cuuint32_t box[3] = {(cuuint32_t)kRP, (cuuint32_t)box_rows, 1u};
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
  fast_encode_map(maps, 0, in0.ptr, in0.w, in0.h, in0.pitch, in0.frame_stride, B, geo.grid[0].h_cell + 6);
  const int redo = min_th < ini_th;
  if (redo) cudaMemsetAsync(fb_count, 0, sizeof(uint32_t), st);
  dim3 tgrid(geo.n_tiles, B);
  int rh = 1;
  for (int l = 0; l < geo.nlevels; ++l) rh = std::max(rh, geo.grid[l].h_cell);
  rh = std::min(rh, kRHmax);
  // pixels (+ 3 rows above / below) | scores | NMS bitmap | survivor list
  const size_t smem = (size_t)(rh + 6) * kRP + (size_t)rh * kRW + (size_t)rh * (kRW / 32) * 4 + (size_t)rh * kRW * 2;
  if (smem > 48 * 1024) {
    cudaFuncSetAttribute(fast_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(fast_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  if (aligned)
    fast_rows_kernel<true><<<tgrid, kRowsThreads, smem, st>>>(d_geo, in0, maps, ini_th, redo, pool, pool_cap, pool_count,
                                                             cell_tab, fb_list, fb_count, status, rh);
  else
    fast_rows_kernel<false><<<tgrid, kRowsThreads, smem, st>>>(d_geo, in0, maps, ini_th, redo, pool, pool_cap, pool_count,
                                                              cell_tab, fb_list, fb_count, status, rh);
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
