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
// So one CTA per cell computes the score map once, takes the NMS verdict once, and applies the
// ini/min fallback by counting.  Output order is raster inside the cell (prefix sum, no
// atomics inside a cell); cells are stitched in row-major order later by K3.
#include "orb_fast_score.cuh"
#include "orb_kernels.cuh"

namespace psl {

constexpr int kFastThreads = 128;
constexpr int kTilePitch = 72;      // >= kMaxCellDim, multiple of 4
constexpr int kInteriorMax = 60;    // kMaxCellDim - 6

__device__ __forceinline__ bool has_run9(uint32_t m16) {
  uint32_t m = m16 | (m16 << 16);
  uint32_t r = m & (m >> 1);
  r &= r >> 2;
  r &= r >> 4;       // runs of 8
  r &= m >> 8;       // runs of 9
  return (r & 0xFFFFu) != 0;
}

__global__ void __launch_bounds__(kFastThreads)
    fast_cells_kernel(const OrbGeometry* __restrict__ geo, ImgBatch in0, int ini_th, int min_th,
                      uint32_t* __restrict__ pool, int pool_cap, uint32_t* __restrict__ pool_count,
                      uint2* __restrict__ cell_tab, uint32_t* __restrict__ status) {
  __shared__ __align__(16) uint8_t s_tile[kMaxCellDim][kTilePitch];
  __shared__ __align__(16) uint8_t s_score[kInteriorMax + 2][kInteriorMax + 4];  // 1-px zero rim for the NMS
  __shared__ uint16_t s_list[kInteriorMax * kInteriorMax];
  __shared__ int s_nlist;
  __shared__ int s_warp[2][kFastThreads / 32];
  __shared__ uint32_t s_base;

  const int cell = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  // locate the level of this cell
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

  const uint8_t* __restrict__ img;
  int pitch;
  if (lvl == 0) {
    img = in0.ptr + (size_t)b * in0.frame_stride;
    pitch = in0.pitch;
  } else {
    img = geo->level[lvl].ptr + (size_t)b * geo->level[lvl].frame_stride;
    pitch = geo->level[lvl].pitch;
  }

  for (int i = tid; i < tw * th; i += kFastThreads) {
    const int y = i / tw, x = i - y * tw;
    s_tile[y][x] = __ldg(img + (size_t)(iniY + y) * pitch + iniX + x);
  }
  for (int i = tid; i < (kInteriorMax + 2) * (kInteriorMax + 4) / 4; i += kFastThreads)
    reinterpret_cast<uint32_t*>(&s_score[0][0])[i] = 0u;
  if (tid == 0) s_nlist = 0;
  __syncthreads();

  // ---- pass A: corner test at the lower threshold, compact the few survivors ----------------
  const int t_lo = min(ini_th, min_th);
  for (int i = tid; i < iw * ih; i += kFastThreads) {
    const int y = i / iw, x = i - y * iw;
    const uint8_t* c = &s_tile[y + 3][x + 3];
    const int v = *c;
    // Any 9-arc contains one pixel of each antipodal pair: two cheap rejects.  (Compared on the
    // pixel values, not on differences: nvcc 12.9 packs min/max/abs of u8 differences into
    // unsigned 16-bit SIMD for sm_100a and gets negative differences wrong — see DESIGN.md.)
    const int hi = v + t_lo, lo = v - t_lo;
    const int q0 = c[3 * kTilePitch], q8 = c[-3 * kTilePitch];
    if (q0 >= lo && q0 <= hi && q8 >= lo && q8 <= hi) continue;
    const int q4 = c[3], q12 = c[-3];
    if (q4 >= lo && q4 <= hi && q12 >= lo && q12 <= hi) continue;
    uint32_t mb = 0, md = 0;  // circle pixels darker than v - t (d > t)  /  brighter than v + t
#define PSL_CIRC(k, dx, dy)                                  \
  {                                                          \
    const int pv = c[(dy) * kTilePitch + (dx)];              \
    mb |= (pv < lo ? 1u : 0u) << (k);                        \
    md |= (pv > hi ? 1u : 0u) << (k);                        \
  }
    PSL_CIRC(0, 0, 3) PSL_CIRC(1, 1, 3) PSL_CIRC(2, 2, 2) PSL_CIRC(3, 3, 1) PSL_CIRC(4, 3, 0) PSL_CIRC(5, 3, -1)
    PSL_CIRC(6, 2, -2) PSL_CIRC(7, 1, -3) PSL_CIRC(8, 0, -3) PSL_CIRC(9, -1, -3) PSL_CIRC(10, -2, -2)
    PSL_CIRC(11, -3, -1) PSL_CIRC(12, -3, 0) PSL_CIRC(13, -3, 1) PSL_CIRC(14, -2, 2) PSL_CIRC(15, -1, 3)
#undef PSL_CIRC
    if (has_run9(mb) || has_run9(md)) s_list[atomicAdd(&s_nlist, 1)] = (uint16_t)i;
  }
  __syncthreads();

  // ---- pass B: exact scores of the survivors (dense, no divergence) -------------------------
  const int nlist = s_nlist;
  for (int k = tid; k < nlist; k += kFastThreads) {
    const int i = s_list[k];
    const int y = i / iw, x = i - y * iw;
    const uint8_t* c = &s_tile[y + 3][x + 3];
    s_score[y + 1][x + 1] = (uint8_t)fast_score_at<kTilePitch>(c);
  }
  __syncthreads();

  // ---- NMS + threshold fallback + raster-ordered compaction ---------------------------------
  const int npx = iw * ih;
  const int run = (npx + kFastThreads - 1) / kFastThreads;  // <= 29
  const int p0 = tid * run, p1 = min(p0 + run, npx);
  uint32_t keep_ini = 0, keep_lo = 0;
  {
    int y = p0 / iw, x = p0 - y * iw;
    for (int p = p0; p < p1; ++p) {
      const int s = s_score[y + 1][x + 1];
      if (s) {
        const uint8_t* r0 = &s_score[y][x];
        const uint8_t* r1 = &s_score[y + 1][x];
        const uint8_t* r2 = &s_score[y + 2][x];
        const int m = max(max(max(r0[0], r0[1]), max(r0[2], r1[0])), max(max(r1[2], r2[0]), max(r2[1], r2[2])));
        if (s > m) {
          if (s >= ini_th) keep_ini |= 1u << (p - p0);
          if (s >= min_th) keep_lo |= 1u << (p - p0);
        }
      }
      if (++x == iw) { x = 0; ++y; }
    }
  }
  // block exclusive scan of both counts
  const int lane = tid & 31, wid = tid >> 5;
  int c_ini = __popc(keep_ini), c_lo = __popc(keep_lo);
  int i_ini = c_ini, i_lo = c_lo;
#pragma unroll
  for (int dlt = 1; dlt < 32; dlt <<= 1) {
    const int a = __shfl_up_sync(0xffffffffu, i_ini, dlt), bb = __shfl_up_sync(0xffffffffu, i_lo, dlt);
    if (lane >= dlt) { i_ini += a; i_lo += bb; }
  }
  if (lane == 31) { s_warp[0][wid] = i_ini; s_warp[1][wid] = i_lo; }
  __syncthreads();
  int base_ini = 0, base_lo = 0, tot_ini = 0, tot_lo = 0;
#pragma unroll
  for (int w = 0; w < kFastThreads / 32; ++w) {
    if (w < wid) { base_ini += s_warp[0][w]; base_lo += s_warp[1][w]; }
    tot_ini += s_warp[0][w];
    tot_lo += s_warp[1][w];
  }
  const bool use_ini = tot_ini > 0;  // :812-816 retry with minThFAST only when the cell is empty
  const int total = use_ini ? tot_ini : tot_lo;
  uint32_t keep = use_ini ? keep_ini : keep_lo;
  int pos = use_ini ? base_ini + i_ini - c_ini : base_lo + i_lo - c_lo;
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
    const int p = p0 + bit;
    const int y = p / iw, x = p - y * iw;
    // cell-local + (j*wCell, i*hCell) (:820-825) == level coordinate - minBorder
    out[pos++] = pack_cand(x + 3 + cj * g.w_cell, y + 3 + ci * g.h_cell, s_score[y + 1][x + 1]);
  }
}

void launch_fast_cells(const OrbGeometry* d_geo, const OrbGeometry& geo, ImgBatch in0, int ini_th, int min_th,
                       uint32_t* pool, int pool_cap, uint32_t* pool_count, uint2* cell_tab, uint32_t* status, int B,
                       cudaStream_t st) {
  dim3 grid(geo.total_cells, B);
  fast_cells_kernel<<<grid, kFastThreads, 0, st>>>(d_geo, in0, ini_th, min_th, pool, pool_cap, pool_count, cell_tab,
                                                   status);
}

}  // namespace psl
