// K3: DistributeOctTree for every (frame, level) of a batch — see orb_octree_core.cuh for the
// algorithm.  This file only gathers the level's candidates in the reference's order
// (cells row-major, raster inside a cell; ORBextractor.cc:789-829) and runs the selection.
#include "orb_kernels.cuh"
#include "orb_octree_core.cuh"

namespace psl {

constexpr int kOctThreads = 256;

template <int NODE_CAP>
__global__ void __launch_bounds__(kOctThreads)
    octree_kernel(const OrbGeometry* __restrict__ geo, const uint32_t* __restrict__ pool, int pool_cap,
                  const uint2* __restrict__ cell_tab, uint32_t* __restrict__ key_scratch,
                  uint16_t* __restrict__ node_scratch, uint32_t* __restrict__ sel, int32_t* __restrict__ sel_count,
                  uint32_t* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using Sh = octree::Shared<kOctThreads, NODE_CAP>;
  Sh& sh = *reinterpret_cast<Sh*>(smem_raw);
  __shared__ int s_red[2][kOctThreads / 32];

  const int lvl = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const CellGrid g = geo->grid[lvl];
  const int ncell = g.n_cols * g.n_rows;
  const uint2* tab = cell_tab + (size_t)b * geo->total_cells;

  // candidates of this level, and of all lower levels (offset of our scratch segment)
  int mine = 0, before = 0;
  for (int c = tid; c < g.first_cell + ncell; c += kOctThreads) {
    const int cnt = (int)tab[c].y;
    if (c < g.first_cell) before += cnt; else mine += cnt;
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    mine += __shfl_xor_sync(0xffffffffu, mine, d);
    before += __shfl_xor_sync(0xffffffffu, before, d);
  }
  if ((tid & 31) == 0) { s_red[0][tid >> 5] = mine; s_red[1][tid >> 5] = before; }
  __syncthreads();
  int n = 0, off = 0;
#pragma unroll
  for (int w = 0; w < kOctThreads / 32; ++w) { n += s_red[0][w]; off += s_red[1][w]; }

  int32_t* out_count = sel_count + (size_t)b * geo->nlevels + lvl;
  if (n == 0 || off + n > pool_cap || n > 65535) {
    // n > 65535 would overflow the packed 16-bit scan counters; reported as pool overflow
    if (tid == 0) {
      *out_count = 0;
      if (n > 65535) atomicOr(status, kStatCandOverflow);
    }
    return;
  }
  uint32_t* keys0 = key_scratch + ((size_t)b * 2 + 0) * pool_cap + off;
  uint32_t* keys1 = key_scratch + ((size_t)b * 2 + 1) * pool_cap + off;
  uint16_t* kn0 = node_scratch + ((size_t)b * 2 + 0) * pool_cap + off;
  uint16_t* kn1 = node_scratch + ((size_t)b * 2 + 1) * pool_cap + off;
  const uint32_t* fpool = pool + (size_t)b * pool_cap;

  // gather: exclusive scan of the cell counts in row-major cell order, one thread per cell
  if (tid == 0) sh.error = 0;
  int carry = 0;
  for (int base = 0; base < ncell; base += kOctThreads) {
    const int c = base + tid;
    uint2 e = make_uint2(0u, 0u);
    if (c < ncell) e = tab[g.first_cell + c];
    sh.iscan_tmp[tid] = 0;
    int32_t* stage = reinterpret_cast<int32_t*>(sh.strip_base);
    stage[tid] = (int)e.y;
    __syncthreads();
    const int tot = octree::scan_i32<kOctThreads>(stage, sh.iscan_tmp);
    const int dst = carry + stage[tid];
    for (uint32_t k = 0; k < e.y; ++k) keys0[dst + k] = fpool[e.x + k];
    carry += tot;
    __syncthreads();
  }
  __syncthreads();

  const int N = geo->quota[lvl];
  uint32_t* out = sel + (size_t)b * geo->total_sel + geo->sel_off[lvl];
  const int L = octree::select<kOctThreads, NODE_CAP>(sh, n, keys0, keys1, kn0, kn1, geo->n_ini[lvl], geo->hx[lvl],
                                                      g.max_bx - kMinBorder, g.max_by - kMinBorder, N, out,
                                                      geo->sel_cap[lvl]);
  if (tid == 0) {
    if (L < 0 || L > geo->sel_cap[lvl]) {
      atomicOr(status, kStatNodeOverflow);
      *out_count = 0;
    } else {
      *out_count = L;
    }
    if (sh.error == 2) atomicOr(status, kStatBadRoot);
  }
}

size_t octree_smem_bytes(int node_cap) {
  if (node_cap <= 256) return sizeof(octree::Shared<kOctThreads, 256>);
  if (node_cap <= 512) return sizeof(octree::Shared<kOctThreads, 512>);
  if (node_cap <= 1024) return sizeof(octree::Shared<kOctThreads, 1024>);
  return sizeof(octree::Shared<kOctThreads, 2048>);
}

template <int NODE_CAP>
static void launch_octree_t(const OrbGeometry* d_geo, const OrbGeometry& geo, const uint32_t* pool, int pool_cap,
                            const uint2* cell_tab, uint32_t* key_scratch, uint16_t* node_scratch, uint32_t* sel,
                            int32_t* sel_count, uint32_t* status, int B, cudaStream_t st) {
  const size_t smem = sizeof(octree::Shared<kOctThreads, NODE_CAP>);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(octree_kernel<NODE_CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  dim3 grid(geo.nlevels, B);
  octree_kernel<NODE_CAP><<<grid, kOctThreads, smem, st>>>(d_geo, pool, pool_cap, cell_tab, key_scratch,
                                                           node_scratch, sel, sel_count, status);
}

void launch_octree(const OrbGeometry* d_geo, const OrbGeometry& geo, const uint32_t* pool, int pool_cap,
                   const uint2* cell_tab, uint32_t* key_scratch, uint16_t* node_scratch, uint32_t* sel,
                   int32_t* sel_count, uint32_t* status, int B, cudaStream_t st) {
  int need = 0;
  for (int l = 0; l < geo.nlevels; ++l) need = need > geo.sel_cap[l] ? need : geo.sel_cap[l];
#define PSL_GO(CAP) \
  launch_octree_t<CAP>(d_geo, geo, pool, pool_cap, cell_tab, key_scratch, node_scratch, sel, sel_count, status, B, st)
  if (need <= 256) PSL_GO(256);
  else if (need <= 512) PSL_GO(512);
  else if (need <= 1024) PSL_GO(1024);
  else PSL_GO(2048);
#undef PSL_GO
}

}  // namespace psl
