// K4 + K6: one warp per selected keypoint computes the intensity-centroid orientation
// (IC_Angle, ORBextractor.cc:77-104) on the unblurred level, the 256-bit steered BRIEF
// descriptor (computeOrbDescriptor, :108-147) on the blurred level, and writes the final
// cv::KeyPoint-compatible record in the reference's output order (levels ascending, list
// order inside a level; :1076-1104).
//
// Float parity: every fp32 operation is an explicit round-to-nearest intrinsic (no FMA
// contraction); cos/sin are evaluated in fp64 and rounded to fp32 (H2).
#include "orb_kernels.cuh"

namespace psl {

__device__ const int8_t g_pattern[1024] = {
#include "orb_pattern.inc"
};
// umax of the r=15 disc, ORBextractor.cc:454-469
__device__ const int8_t g_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

// cv::fastAtan2, degrees
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
  const float sc = (float)(180.0 / 3.141592653589793238462643383279502884);
  const float p1 = __fmul_rn(0.9997878412794807f, sc), p3 = __fmul_rn(-0.3258083974640975f, sc),
              p5 = __fmul_rn(0.1555786518463281f, sc), p7 = __fmul_rn(-0.04432655554792128f, sc);
  const float eps = (float)2.2204460492503131e-16;
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = __fdiv_rn(ay, __fadd_rn(ax, eps));
    c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = __fdiv_rn(ax, __fadd_rn(ay, eps));
    c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0.f) a = __fsub_rn(180.f, a);
  if (y < 0.f) a = __fsub_rn(360.f, a);
  return a;
}

constexpr int kDescWarps = 8;
constexpr int kWinPitch = 44;  // staged window row: 11 words cover x-19 .. x+19 from a word-aligned start

__global__ void __launch_bounds__(kDescWarps * 32)
    describe_kernel(const OrbGeometry* __restrict__ geo, ImgBatch in0, const uint32_t* __restrict__ sel,
                    const int32_t* __restrict__ sel_count, psl_keypoint* __restrict__ kps,
                    uint8_t* __restrict__ desc, int cap, int32_t* __restrict__ n_out, uint32_t* __restrict__ status) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int i = blockIdx.x * kDescWarps + warp;
  const int nl = geo->nlevels;
  // offsets of the levels inside the frame's output block
  const int cnt = lane < nl ? sel_count[(size_t)b * nl + lane] : 0;
  int inc = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  const int off = inc - cnt;
  const int n = __shfl_sync(0xffffffffu, inc, 31);
  if (i == 0 && lane == 0) {
    n_out[b] = n < cap ? n : cap;
    if (n > cap) { atomicOr(status, kStatOutOverflow); atomicMax(status + 1, (uint32_t)b + 1u); }
  }
  if (i >= n || i >= cap) return;
  const uint32_t m = __ballot_sync(0xffffffffu, lane < nl && off <= i);
  const int lvl = 31 - __clz(m);
  const int j = i - __shfl_sync(0xffffffffu, off, lvl);
  const uint32_t key = sel[(size_t)b * geo->total_sel + geo->sel_off[lvl] + j];
  const int x = cand_x(key) + kMinBorder, y = cand_y(key) + kMinBorder;  // :843-844

  const uint8_t* img;
  int pitch;
  if (lvl == 0) {
    img = in0.ptr + (size_t)b * in0.frame_stride;
    pitch = in0.pitch;
  } else {
    img = geo->level[lvl].ptr + (size_t)b * geo->level[lvl].frame_stride;
    pitch = geo->level[lvl].pitch;
  }
  // Both stages gather single bytes around the keypoint (31 patch rows, then 512 rotated pattern points within
  // 19 px): straight from HBM/L2 that is one 32-byte sector per byte.  The warp first copies the two windows
  // into shared memory with row-contiguous word loads (~80 sectors instead of ~540) and gathers from there.
  __shared__ __align__(16) uint8_t s_blur[kDescWarps][39][kWinPitch];
  const uint8_t* blur_img = geo->blur[lvl].ptr + (size_t)b * geo->blur[lvl].frame_stride;
  const int bp = geo->blur[lvl].pitch;
  const bool aligned = (((uintptr_t)img | (uintptr_t)blur_img) & 3) == 0 && ((pitch | bp) & 3) == 0;
  const int xs = (x - kEdge) & ~3;            // first staged column (word aligned), x - 19 >= 0
  const int nw = ((x + kEdge - xs) >> 2) + 1;  // words per row (<= 11)
  // two rows per step, a word column per lane (no index division): 20 steps for the 39 rows.  All loads of the warp —
  // these and the patch rows of IC_Angle below — are issued before anything is consumed: the kernel is bound by the
  // latency of these L2 / HBM reads, not by instructions.
  const int sc_ = lane & 15, rs = lane >> 4;
  uint32_t bv[20];
  if (aligned && sc_ < nw) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(blur_img + (size_t)(y - kEdge + rs) * bp + xs) + sc_;
#pragma unroll
    for (int k = 0; k < 20; ++k)
      if (rs + 2 * k < 39) bv[k] = __ldg(src + (size_t)k * (bp >> 1));
  }
  const int xo = x - xs;  // column of the keypoint inside the staged rows
  // ---- IC_Angle: lane r owns patch row v = r - 15 -------------------------------------------
  int m10 = 0, m01 = 0;
  if (lane < 31) {
    const int v = lane - kHalfPatch;
    const int d = g_umax[v < 0 ? -v : v];
    int sum = 0;
    if (aligned) {
      // the row's 2 d + 1 pixels straight from the level image as whole words; sum and first moment of a word are two
      // byte dot products (pixels u8, offsets u = column - x as s8), bytes outside [-d, d] masked out
      const int xr = (x - kHalfPatch) & ~3;                       // first word of the widest row
      const uint32_t* row = reinterpret_cast<const uint32_t*>(img + (size_t)(y + v) * pitch + xr);
      const uint32_t dd = (uint32_t)d * 0x01010101u, lim = (uint32_t)(2 * d) * 0x01010101u;
      const uint32_t uw0 = __vadd4((uint32_t)((xr - x) & 0xFF) * 0x01010101u, 0x03020100u);   // u of the word's four bytes
      uint32_t px[9], okm[9];
      uint32_t uw = uw0;
#pragma unroll
      for (int w = 0; w < 9; ++w) {   // x - 15 .. x + 15 spans at most 9 aligned words
        okm[w] = __vcmpleu4(__vadd4(uw, dd), lim);      // 0xFF where -d <= u <= d
        px[w] = okm[w] ? __ldg(row + w) : 0u;
        uw = __vadd4(uw, 0x04040404u);
      }
      // (the staged blur window goes to shared memory while those loads are in flight)
      if (sc_ < nw) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(&s_blur[warp][rs][0]) + sc_;
#pragma unroll
        for (int k = 0; k < 20; ++k)
          if (rs + 2 * k < 39) dst[k * 2 * (kWinPitch / 4)] = bv[k];
      }
      uw = uw0;
#pragma unroll
      for (int w = 0; w < 9; ++w) {
        const uint32_t p = px[w] & okm[w];
        int t;
        asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(t) : "r"(p), "r"(0x01010101), "r"(0));
        sum += t;
        asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(m10) : "r"(p), "r"(uw), "r"(m10));
        uw = __vadd4(uw, 0x04040404u);
      }
    } else {
      const uint8_t* row = img + (size_t)(y + v) * pitch + x;
      for (int u = -d; u <= d; ++u) {
        const int val = __ldg(row + u);
        sum += val;
        m10 += u * val;
      }
    }
    m01 = v * sum;
  } else if (aligned && sc_ < nw) {   // lane 31 has no patch row, only its share of the staging
    uint32_t* dst = reinterpret_cast<uint32_t*>(&s_blur[warp][rs][0]) + sc_;
#pragma unroll
    for (int k = 0; k < 20; ++k)
      if (rs + 2 * k < 39) dst[k * 2 * (kWinPitch / 4)] = bv[k];
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    m10 += __shfl_xor_sync(0xffffffffu, m10, d);
    m01 += __shfl_xor_sync(0xffffffffu, m01, d);
  }
  __syncwarp();   // the staged blur window is complete
  const float angle = fast_atan2_deg((float)m01, (float)m10);

  // ---- steered BRIEF: lane k makes descriptor byte k ----------------------------------------
  const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);  // :107
  const float rad = __fmul_rn(angle, factorPI);
  double sd, cd;
  sincos((double)rad, &sd, &cd);   // same values as sin() / cos(), one range reduction
  const float a = (float)cd, bsin = (float)sd;
  const uint8_t* bl = aligned ? &s_blur[warp][kEdge][xo] : blur_img + (size_t)y * bp + x;
  const int bpp = aligned ? kWinPitch : bp;
  const char4* pat = reinterpret_cast<const char4*>(g_pattern) + lane * 8;
  uint32_t byte = 0;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const char4 p = pat[t];
    const float x0 = (float)p.x, y0 = (float)p.y, x1 = (float)p.z, y1 = (float)p.w;
    const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, bsin), __fmul_rn(y0, a)));
    const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, bsin)));
    const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, bsin), __fmul_rn(y1, a)));
    const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, bsin)));
    const int t0 = bl[r0 * bpp + c0], t1 = bl[r1 * bpp + c1];
    byte |= (t0 < t1 ? 1u : 0u) << t;
  }
  desc[((size_t)b * cap + i) * 32 + lane] = (uint8_t)byte;

  if (lane == 0) {
    psl_keypoint k;
    const float sc = geo->scale[lvl];
    k.x = lvl ? __fmul_rn((float)x, sc) : (float)x;  // :1095-1101
    k.y = lvl ? __fmul_rn((float)y, sc) : (float)y;
    k.size = geo->kp_size[lvl];
    k.angle = angle;
    k.response = (float)cand_score(key);
    k.octave = lvl;
    k.class_id = -1;
    kps[(size_t)b * cap + i] = k;
  }
}

void launch_describe(const OrbGeometry* d_geo, const OrbGeometry& geo, ImgBatch in0, const uint32_t* sel,
                     const int32_t* sel_count, psl_keypoint* kps, uint8_t* desc, int cap, int32_t* n_out,
                     uint32_t* status, int B, cudaStream_t st) {
  const int slots = geo.total_sel < cap ? geo.total_sel : cap;
  dim3 grid((slots + kDescWarps - 1) / kDescWarps, B);
  if (grid.x == 0) grid.x = 1;
  describe_kernel<<<grid, kDescWarps * 32, 0, st>>>(d_geo, in0, sel, sel_count, kps, desc, cap, n_out, status);
}

}  // namespace psl
