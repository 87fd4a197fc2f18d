// Shared device/host definitions for the sm_100a front-end kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/psl_frontend.h"

namespace psl {

constexpr int kEdge = 19;        // EDGE_THRESHOLD  (ORBextractor.cc:74)
constexpr int kMinBorder = 16;   // EDGE_THRESHOLD-3 (ORBextractor.cc:773)
constexpr int kHalfPatch = 15;   // HALF_PATCH_SIZE (ORBextractor.cc:73)
constexpr int kPatch = 31;       // PATCH_SIZE      (ORBextractor.cc:72)
constexpr int kMaxLevels = 16;
constexpr int kCellW = 30;       // W in ComputeKeyPointsOctTree (ORBextractor.cc:769)
constexpr int kMaxCellDim = 66;  // cell sub-image side incl. the 6-px overlap (wCell+6 <= 66)

// A batch of same-sized u8 images: frame b, row y at ptr + b*frame_stride + y*pitch.
struct ImgBatch {
  const uint8_t* ptr;
  int32_t pitch;
  int64_t frame_stride;
  int32_t w, h;
};
struct ImgBatchMut {
  uint8_t* ptr;
  int32_t pitch;
  int64_t frame_stride;
  int32_t w, h;
};

// Per-level FAST cell grid (ORBextractor.cc:781-787)
struct CellGrid {
  int32_t n_cols, n_rows;  // cells kept after the skip tests (:794,803) are the leading ones
  int32_t w_cell, h_cell;
  int32_t max_bx, max_by;  // maxBorderX/Y (exclusive)
  int32_t first_cell;      // index of this level's first cell in the per-frame cell table
};

// Everything the ORB kernels need to know about the pyramid of one (w,h).
struct OrbGeometry {
  int32_t nlevels;
  int32_t total_cells;  // cells of all levels (per frame)
  int32_t total_sel;    // sum of sel_cap
  ImgBatchMut level[kMaxLevels];  // level 0 aliases the input (ptr filled per call)
  ImgBatchMut blur[kMaxLevels];
  CellGrid grid[kMaxLevels];
  int32_t quota[kMaxLevels];    // mnFeaturesPerLevel
  int32_t n_ini[kMaxLevels];    // octree roots (ORBextractor.cc:543)
  float hx[kMaxLevels];         // root width hX (ORBextractor.cc:545)
  int32_t sel_cap[kMaxLevels];  // capacity of the selected list of a level
  int32_t sel_off[kMaxLevels];  // offset of a level's selected list inside a frame's block
  float scale[kMaxLevels];      // mvScaleFactor
  float kp_size[kMaxLevels];    // (int)(31*scale)
  // lookup tables of the dense FAST path (device memory): tile -> lvl<<24 | ty<<12 | tx for the n_tiles score
  // tiles of a frame, then cell -> lvl<<24 | ci<<12 | cj for its total_cells cells
  const uint32_t* fast_tab;
  int32_t n_tiles;
};

// FAST candidate packing: x-16 (12 bit) | y-16 (12 bit) | score (8 bit)
__host__ __device__ inline uint32_t pack_cand(int xr, int yr, int score) {
  return ((uint32_t)xr << 20) | ((uint32_t)yr << 8) | (uint32_t)score;
}
__host__ __device__ inline int cand_x(uint32_t c) { return (int)(c >> 20); }
__host__ __device__ inline int cand_y(uint32_t c) { return (int)((c >> 8) & 0xFFFu); }
__host__ __device__ inline int cand_score(uint32_t c) { return (int)(c & 0xFFu); }

// device status word bits (sticky until read by the host)
enum : uint32_t {
  kStatCandOverflow = 1u,   // candidate pool of a frame too small
  kStatOutOverflow = 2u,    // caller's kps/desc capacity too small
  kStatNodeOverflow = 4u,   // octree node table too small (internal bound violated)
  kStatBadRoot = 8u,        // candidate outside every octree root (aspect not supported)
  kStatLineNeighbours = 16u,  // a segment has more merge neighbours than line::kNbCap
  kStatLineRaw = 32u,       // more raw LSD segments than psl_config.line_max_raw
};

}  // namespace psl
