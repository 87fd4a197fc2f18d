// Launchers of the line front end (LSD + merge + LBD); all asynchronous on `st`, batch-first.
#pragma once
#include "line_core.cuh"
#include "psl_common.cuh"

namespace psl {

// geometry + per-chunk device buffers of the line path for one frame size
struct LineBuffers {
  int w, h, pitch;     // input size; pitch of the u8 work images (multiple of 128)
  int Ws, Hs;          // LSD working size (0.8x)
  int min_reg_size;    // LSD minimal region size (-logNT / log10(p))
  int raw_cap;         // raw LSD segments kept per frame
  const short2* xtab;  // [Ws] exact-resize taps (src index, Q8 weight of the right tap or -1)
  const short2* ytab;  // [Hs]
  const uint4* xw4;     // xtab per group of 4 outputs for the word kernel (weights 256-f | f << 16), null if unusable
  const uint32_t* xo4;  // first source word | 4-bit byte offsets of the 4 left taps << 16
  uint8_t* blur;       // [C][h][pitch]   7x7 sigma 0.75 (LSD) and, later, 5x5 sigma 1 (LBD)
  uint8_t* scaled;     // [C][Hs][Ws]
  const float4* lut;   // [1021*1021] pixel record of every integer gradient (gx, gy)
  const float2* seed_lut;  // [1021*1021] (float)cos / (float)sin of the fp64 angle: the sums a region starts with
  float4* pix;         // [C][Hs*Ws]  (angle in degrees | cos | sin | int bits: integer gradient + "not available" flag)
  uint32_t* reg;       // [C][Hs*Ws]
  int32_t* max_n2;     // [C]
  int32_t* row_cnt;    // [C][Hs]  defined pixels per row -> exclusive offsets
  int32_t* n_def;      // [C]
  uint16_t* key_in;    // [C][Hs*Ws]  gradient bin of every seed-capable pixel, raster order
  uint32_t* val_in;    // [C][Hs*Ws]  its pixel index
  uint32_t* val_out;   // [C][Hs*Ws]  the pixel indices in seed order (bins descending, raster order inside a bin)
  float* raw;          // [C][raw_cap][4]
  int32_t* n_raw;      // [C]
  line::Seg* t1;       // [C][raw_cap]
  line::Seg* t2;       // [C][raw_cap]
  float* m_angles;     // merge scratch, [C][raw_cap] each
  float* m_length;
  float* m_sangles;
  uint16_t* m_order;
  uint16_t* m_tmp16;
  uint16_t* m_nb;      // [C][raw_cap][kNbCap]
  uint16_t* m_fw;      // [C][raw_cap][kNbCap]
  line::ScanRec* m_scan;  // [C][raw_cap]
  int32_t* m_cnt;      // [C][2]  lines of the current merge pass | neighbour-list overflow
  uint16_t* m_nb_cnt;
  int16_t* m_code;
  uint16_t* m_check;
  uint16_t* m_loc;
  uint8_t* m_flag;
  short2* gxy;         // [C][h][w] Sobel (dx, dy) of the sigma-1 blurred image
};

size_t lsd_lut_bytes();
size_t lsd_seed_lut_bytes();
void launch_lsd_lut(float4* lut, float2* seed_lut, cudaStream_t st);

// LSD for `nb` frames (cv::LineSegmentDetector behind LineExtractor.cpp:336-337), three stages:
// blur + 0.8x resize + gradient + seed keys (6 launches); stable seed ordering (one counting pass, 1 launch);
// the sequential region-growing core -> raw segments + counts (1 launch)
void launch_lsd_prologue(const LineBuffers& L, ImgBatch in, int nb, cudaStream_t st);
void launch_lsd_order(const LineBuffers& L, int nb, cudaStream_t st);
void launch_lsd_core(const LineBuffers& L, int nb, uint32_t* status, cudaStream_t st);
// clamp + merge + top-N + keylines + line equations (LineExtractor.cpp:338-363); outputs are [nb][cap] blocks
constexpr int kLinePostLaunches = 5;  // prepare, pair scan, finish + prepare, pair scan, finish + key lines
void launch_line_post(const LineBuffers& L, int nb, int nfeatures, psl_keyline* kl, double* lineeq, int cap,
                      int32_t* n_out, uint32_t* status, cudaStream_t st);
// LBD descriptors of the keylines (BinaryDescriptor::compute, LineExtractor.cpp:349-350); lbd72 optional.
// 3 launches: 5x5 sigma-1 blur, Sobel, descriptor.
void launch_lbd(const LineBuffers& L, ImgBatch in, int nb, int nfeatures, const psl_keyline* kl, const int32_t* n_kl,
                int cap, uint8_t* ldesc, float* lbd72, cudaStream_t st);
// LBD weight tables (BinaryDescriptor ctor, binary_descriptor_custom.cpp:219-261), computed once on the host
void upload_lbd_tables();

}  // namespace psl
