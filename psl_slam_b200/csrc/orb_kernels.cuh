// Launchers of the ORB extractor kernels (all asynchronous on `st`).
#pragma once
#include "psl_common.cuh"

namespace psl {

// per-axis resize tables of one level (built on the host)
struct ResizeTables {
  const short4* xt;  // [dw]  (sx, sx+1 clamped, a0, a1)
  const short4* yt;  // [dh]  (sy, sy+1 clamped, b0, b1)
  // the x table again, per group of 4 adjacent outputs (null when a group's taps span more than 3 source words):
  const uint4* xw;      // [ceil(dw/4)] weights a0 | a1 << 16 of the 4 outputs
  const uint32_t* xo;   // [ceil(dw/4)] first source word | byte offsets of the 4 left taps from it, 4 bits each, << 16
};
// host side of xw / xo; false when the scale is too large for the 3-word window
bool resize_group_tables(const short4* xt, int dw, uint4* xw, uint32_t* xo);

// K1: level l from level l-1 (ORBextractor.cc:1120)
void launch_resize(const ImgBatch& src, const ImgBatchMut& dst, const ResizeTables& t, int B, cudaStream_t st);

// Dense FAST path: the host side of its lookup tables (OrbGeometry::fast_tab) and of the TMA descriptors the
// score kernel stages its pixel tiles with (one 3-D u8 tensor [frames][h][pitch] per level).
struct alignas(64) FastMaps {
  unsigned char m[kMaxLevels][128];  // CUtensorMap, opaque here
  uint32_t valid;                    // bit l: level l has a descriptor
};
int fast_tile_count(const OrbGeometry& geo);                       // tiles per frame
void fast_build_tab(const OrbGeometry& geo, uint32_t* host_tab);   // n_tiles + total_cells words
// descriptor of one level; false when the layout cannot be described (base or strides not 16-byte aligned)
bool fast_encode_map(FastMaps& maps, int level, const void* ptr, int w, int h, int pitch, int64_t frame_stride,
                     int frames, int box_rows);
constexpr int kFastLaunches = 2;  // fused rows kernel + per-cell fallback (the counter memset is not a kernel)

// K2: per-cell FAST + NMS + threshold fallback + ordered compaction (ORBextractor.cc:789-829).
// fb_list: [B * total_cells] u32 scratch (cells to redo at minThFAST), fb_count: one u32.
// maps: descriptors of levels >= 1 (level 0 is encoded here from in0).
void launch_fast_cells(const OrbGeometry* d_geo, const OrbGeometry& geo, ImgBatch in0, int ini_th, int min_th,
                       uint32_t* pool, int pool_cap, uint32_t* pool_count, uint2* cell_tab, uint32_t* fb_list,
                       uint32_t* fb_count, FastMaps& maps, uint32_t* status, int B, cudaStream_t st);

// K3: DistributeOctTree per (frame, level) (ORBextractor.cc:539-763)
void launch_octree(const OrbGeometry* d_geo, const OrbGeometry& geo, const uint32_t* pool, int pool_cap,
                   const uint2* cell_tab, uint32_t* key_scratch, uint16_t* node_scratch, uint32_t* sel,
                   int32_t* sel_count, uint32_t* status, int B, cudaStream_t st);

constexpr int kBlurLaunches = 2;  // interior word columns + edge columns (word-aligned source)
// generic separable 7-tap Q8 blur of one image batch (taps t0,t1,t2,t3=centre; dst pitch multiple of 4)
void launch_blur7(const ImgBatch& src, const ImgBatchMut& dst, int t0, int t1, int t2, int t3, int B, cudaStream_t st);

// K5: 7x7 sigma-2 Gaussian blur of every level (ORBextractor.cc:1085-1086)
void launch_gauss7(const OrbGeometry& geo, ImgBatch in0, int B, cudaStream_t st);

// K4+K6: orientation, rBRIEF, final keypoint records (ORBextractor.cc:77-147,837-852,1095-1103)
void launch_describe(const OrbGeometry* d_geo, const OrbGeometry& geo, ImgBatch in0, const uint32_t* sel,
                     const int32_t* sel_count, psl_keypoint* kps, uint8_t* desc, int cap, int32_t* n_out,
                     uint32_t* status, int B, cudaStream_t st);

size_t octree_smem_bytes(int node_cap);

}  // namespace psl
