// K1 pyramid resize and K5 Gaussian blur: the two pure streaming stages of the ORB extractor.
// Integer arithmetic only, bit-exact with cv::resize(INTER_LINEAR) / cv::GaussianBlur on CV_8U
// (SURVEY.md App. A1/A2).  Batch-first: blockIdx.z / .y selects the frame.
#include "orb_kernels.cuh"

namespace psl {

// ---------------------------------------------------------------------------------------------
// K1: level l from level l-1, ORBextractor.cc:1120.  One thread makes 4 consecutive output
// pixels of a row and stores them as one uchar4 (rows are 128-byte aligned).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) resize_kernel(ImgBatch src, ImgBatchMut dst, const short4* __restrict__ xt,
                                                      const short4* __restrict__ yt) {
  const int n4 = (dst.w + 3) >> 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n4 * dst.h) return;
  const int y = idx / n4, x4 = (idx - y * n4) << 2;
  const int b = blockIdx.y;
  const short4 ty = __ldg(yt + y);
  const uint8_t* __restrict__ S0 = src.ptr + (size_t)b * src.frame_stride + (size_t)ty.x * src.pitch;
  const uint8_t* __restrict__ S1 = src.ptr + (size_t)b * src.frame_stride + (size_t)ty.y * src.pitch;
  uint32_t packed = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = x4 + k;
    if (x < dst.w) {
      const short4 tx = __ldg(xt + x);
      const int r0 = (int)__ldg(S0 + tx.x) * tx.z + (int)__ldg(S0 + tx.y) * tx.w;  // Q11 row pass
      const int r1 = (int)__ldg(S1 + tx.x) * tx.z + (int)__ldg(S1 + tx.y) * tx.w;
      const int v = ((((int)ty.z * (r0 >> 4)) >> 16) + (((int)ty.w * (r1 >> 4)) >> 16) + 2) >> 2;
      packed |= (uint32_t)(v & 0xFF) << (8 * k);
    }
  }
  *reinterpret_cast<uint32_t*>(dst.ptr + (size_t)b * dst.frame_stride + (size_t)y * dst.pitch + x4) = packed;
}

// The same arithmetic with whole-word loads: a thread owns 4 adjacent output columns and walks down kRBand output
// rows.  Per source row it loads the 3 aligned words that hold the taps of its 4 outputs, shifts each tap pair
// (S[sx], S[sx+1]) into the low half of a register and makes the Q11 row pass with one IDP.2A (u16 weights x u8
// pixels).  sx+1 is the right tap wherever its weight is non-zero (the clamped ends have a1 = 0), and the row
// pass of a source row that serves two consecutive output rows (every row but one in six at 1.2x) is kept.
constexpr int kRBand = 8;

__global__ void __launch_bounds__(256) resize_words_kernel(ImgBatch src, ImgBatchMut dst, const uint4* __restrict__ xw,
                                                           const uint32_t* __restrict__ xo,
                                                           const short4* __restrict__ yt) {
  const int n4 = (dst.w + 3) >> 2, nbands = (dst.h + kRBand - 1) / kRBand;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n4 * nbands) return;
  const int band = idx / n4, g = idx - band * n4, b = blockIdx.y;
  const uint4 wt = __ldg(xw + g);
  const uint32_t xo_g = __ldg(xo + g);
  const int wb = (int)(xo_g & 0xFFFFu);
  const int nwords = src.pitch >> 2;
  const bool in1 = wb + 1 < nwords, in2 = wb + 2 < nwords;  // stay inside the row (the taps there have weight 0)
  const uint32_t* __restrict__ S = reinterpret_cast<const uint32_t*>(src.ptr + (size_t)b * src.frame_stride) + wb;
  uint8_t* __restrict__ D = dst.ptr + (size_t)b * dst.frame_stride + 4 * g;
  const uint32_t w4[4] = {wt.x, wt.y, wt.z, wt.w};
  int kept_row = -1;
  int kept[4] = {0, 0, 0, 0};
  const int y_end = min((band + 1) * kRBand, dst.h);
  for (int y = band * kRBand; y < y_end; ++y) {
    const short4 ty = __ldg(yt + y);
    int r0[4], r1[4];
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int sy = pass == 0 ? ty.x : ty.y;
      int* r = pass == 0 ? r0 : r1;
      if (pass == 0 && sy == kept_row) {
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = kept[k];
        continue;
      }
      if (pass == 1 && sy == ty.x) {
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = r0[k];
        continue;
      }
      const uint32_t* row = S + (size_t)sy * nwords;
      const uint32_t W0 = __ldg(row), W1 = in1 ? __ldg(row + 1) : 0u, W2 = in2 ? __ldg(row + 2) : 0u;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t off = (xo_g >> (16 + 4 * k)) & 0xFu;
        const uint32_t lo = off < 4u ? W0 : W1, hi = off < 4u ? W1 : W2;
        const uint32_t pair = __funnelshift_r(lo, hi, 8u * (off & 3u));
        r[k] = (int)__dp2a_lo(w4[k], pair, 0u);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) kept[k] = r1[k];
    kept_row = ty.y;
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int v = ((((int)ty.z * (r0[k] >> 4)) >> 16) + (((int)ty.w * (r1[k] >> 4)) >> 16) + 2) >> 2;
      packed |= (uint32_t)(v & 0xFF) << (8 * k);
    }
    *reinterpret_cast<uint32_t*>(D + (size_t)y * dst.pitch) = packed;
  }
}

bool resize_group_tables(const short4* xt, int dw, uint4* xw, uint32_t* xo) {
  const int n4 = (dw + 3) >> 2;
  for (int g = 0; g < n4; ++g) {
    const int wb = xt[4 * g].x >> 2;
    uint32_t w[4] = {0, 0, 0, 0}, offs = 0;
    for (int k = 0; k < 4; ++k) {
      const int x = 4 * g + k;
      if (x >= dw) continue;
      const short4 t = xt[x];
      if (t.y != t.x + 1 && t.w != 0) return false;  // the right tap is not the next pixel
      const int off = t.x - 4 * wb;
      if (off < 0 || off + 1 > 11 || t.z < 0 || t.w < 0) return false;
      w[k] = (uint32_t)(uint16_t)t.z | ((uint32_t)(uint16_t)t.w << 16);
      offs |= (uint32_t)off << (4 * k);
    }
    if (wb > 0xFFFF) return false;
    xw[g] = make_uint4(w[0], w[1], w[2], w[3]);
    xo[g] = (uint32_t)wb | (offs << 16);
  }
  return true;
}

void launch_resize(const ImgBatch& src, const ImgBatchMut& dst, const ResizeTables& t, int B, cudaStream_t st) {
  const int n4 = (dst.w + 3) >> 2;
  const bool aligned = ((uintptr_t)src.ptr & 3) == 0 && (src.pitch & 3) == 0 && (src.frame_stride & 3) == 0;
  if (t.xw && aligned) {
    const int nbands = (dst.h + kRBand - 1) / kRBand;
    dim3 grid((n4 * nbands + 255) / 256, B);
    resize_words_kernel<<<grid, 256, 0, st>>>(src, dst, t.xw, t.xo, t.yt);
    return;
  }
  dim3 grid((n4 * dst.h + 255) / 256, B);
  resize_kernel<<<grid, 256, 0, st>>>(src, dst, t.xt, t.yt);
}

// ---------------------------------------------------------------------------------------------
// K5: separable 7-tap Q8 blur of CV_8U (cv::GaussianBlur bit-exact path), REFLECT_101, dst = (v + 2^15) >> 16.
// Taps (t0,t1,t2,t3 = centre): [18,34,48,56] = 7x7 sigma 2 (ORB), [0,4,56,136] = 7x7 sigma 0.75 (LSD prologue),
// [0,14,62,104] = 5x5 sigma 1 (LBD).
// No shared memory: a thread owns 4 adjacent columns (one output word) and walks down a band of
// rows with the seven row-pass sums of its columns in a register ring.  Per input row it loads
// three aligned words (coalesced across the warp, neighbours hit L1), makes the 4 horizontal
// sums, and emits one uchar4 of the row 3 above.
// ---------------------------------------------------------------------------------------------
constexpr int kGBand = 28;  // output rows per thread (4 ring turns of 7)
constexpr int kGEdgeBand = 7;  // the edge columns are few and latency-bound: shorter bands, four times the threads

__device__ __forceinline__ int reflect101(int i, int n) {
  // valid for -n < i < 2n-1, which holds for a 3-px halo on any level the extractor accepts
  i = i < 0 ? -i : i;
  return i >= n ? 2 * (n - 1) - i : i;
}

// MODE 0: any layout; a warp covers 32 word columns, the edge columns (and every column of an image whose rows are
//         not word aligned) read single bytes through the reflected index table.
// MODE 1: the interior word columns [1, n_int] of a word-aligned image: 3 aligned words per row, the row pass of a
//         pixel is two IDP.4A (taps as packed bytes) on the two funnel-shifted words that hold its 7 pixels.
// MODE 2: the remaining (edge) word columns of such an image, one (band, column) per thread.
template <int MODE>
__global__ void __launch_bounds__(128) gauss7_kernel(ImgBatch src, ImgBatchMut dst, int t0, int t1, int t2, int t3,
                                                     int n_int) {
  const int lane = threadIdx.x, b = blockIdx.z;
  const int w = src.w, h = src.h;
  constexpr int kBand = MODE == 2 ? kGEdgeBand : kGBand;
  int xw, y0;
  if (MODE == 2) {
    const int ne = ((w + 3) >> 2) - n_int, nbands = (h + kBand - 1) / kBand;
    const int t = blockIdx.x * 128 + threadIdx.y * 32 + lane;
    if (t >= ne * nbands) return;
    const int band = t / ne, e = t - band * ne;
    xw = e == 0 ? 0 : n_int + e;
    y0 = band * kBand;
  } else {
    xw = blockIdx.x * 32 + lane + (MODE == 1 ? 1 : 0);  // output word (4 pixels)
    y0 = (blockIdx.y * 4 + threadIdx.y) * kGBand;
    if (MODE == 1 && xw > n_int) return;
  }
  if (y0 >= h || xw * 4 >= w) return;
  const int x = xw * 4;
  const bool interior = MODE == 1;
  const uint8_t* __restrict__ S = src.ptr + (size_t)b * src.frame_stride;
  uint8_t* __restrict__ D = dst.ptr + (size_t)b * dst.frame_stride;
  int xi[10];
  if (!interior) {
#pragma unroll
    for (int k = 0; k < 10; ++k) xi[k] = min(reflect101(x - 3 + k, w), w - 1);
  }
  const uint32_t Ta = (uint32_t)t0 | ((uint32_t)t1 << 8) | ((uint32_t)t2 << 16) | ((uint32_t)t3 << 24);  // x-3 .. x
  const uint32_t Tb = (uint32_t)t2 | ((uint32_t)t1 << 8) | ((uint32_t)t0 << 16);                         // x+1 .. x+3
  int ring[7][4];
#pragma unroll
  for (int j = 0; j < 7; ++j)
#pragma unroll
    for (int k = 0; k < 4; ++k) ring[j][k] = 0;
  const int y_end = min(y0 + kBand, h);
  for (int yb = y0 - 3; yb < y_end + 3; yb += 7) {
    // issue the loads of seven rows before touching any of them (memory-level parallelism)
    uint32_t ld[7][3];
    if (interior) {
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const int yy = min(yb + j, y_end + 2);
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(S + (size_t)reflect101(yy, h) * src.pitch) + xw;
        ld[j][0] = __ldg(rw - 1);
        ld[j][1] = __ldg(rw);
        ld[j][2] = __ldg(rw + 1);
      }
    }
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int yy = yb + j;
      if (yy >= y_end + 3) break;
      if (interior) {
        const uint32_t w0 = ld[j][0], w1 = ld[j][1], w2 = ld[j][2];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // pixels x+k-3 .. x+k and x+k+1 .. x+k+4 (the tap of the last one is 0)
          const uint32_t A = k == 3 ? w1 : __funnelshift_r(w0, w1, 8 * (k + 1));
          const uint32_t Bw = k == 3 ? w2 : __funnelshift_r(w1, w2, 8 * (k + 1));
          ring[j][k] = (int)__dp4a(A, Ta, __dp4a(Bw, Tb, 0u));
        }
      } else {
        int p[10];
        const uint8_t* row = S + (size_t)reflect101(yy, h) * src.pitch;
#pragma unroll
        for (int k = 0; k < 10; ++k) p[k] = __ldg(row + xi[k]);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          ring[j][k] = t0 * (p[k] + p[k + 6]) + t1 * (p[k + 1] + p[k + 5]) + t2 * (p[k + 2] + p[k + 4]) + t3 * p[k + 3];
      }
      const int yo = yy - 3;
      if (yo >= y0) {
        // newest row is slot j; the 7 rows yo-3..yo+3 sit in slots (j+1)%7 .. j
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t v = (uint32_t)t0 * (uint32_t)(ring[(j + 1) % 7][k] + ring[j][k]) +
                             (uint32_t)t1 * (uint32_t)(ring[(j + 2) % 7][k] + ring[(j + 6) % 7][k]) +
                             (uint32_t)t2 * (uint32_t)(ring[(j + 3) % 7][k] + ring[(j + 5) % 7][k]) +
                             (uint32_t)t3 * (uint32_t)ring[(j + 4) % 7][k];
          packed |= ((v + 32768u) >> 16) << (8 * k);
        }
        *reinterpret_cast<uint32_t*>(D + (size_t)yo * dst.pitch + x) = packed;
      }
    }
  }
}

void launch_blur7(const ImgBatch& src, const ImgBatchMut& dst, int t0, int t1, int t2, int t3, int B, cudaStream_t st) {
  const int nwords = (dst.w + 3) / 4, nbands = (dst.h + kGBand - 1) / kGBand;
  const bool aligned = ((uintptr_t)src.ptr & 3) == 0 && (src.pitch & 3) == 0 && (src.frame_stride & 3) == 0;
  const int n_int = dst.w >= 12 ? (dst.w - 8) / 4 : 0;  // word columns xw in [1, n_int]: 4 xw + 7 < w
  dim3 block(32, 4);
  if (aligned && n_int > 0) {
    dim3 grid((n_int + 31) / 32, (nbands + 3) / 4, B);
    gauss7_kernel<1><<<grid, block, 0, st>>>(src, dst, t0, t1, t2, t3, n_int);
    const int ebands = (dst.h + kGEdgeBand - 1) / kGEdgeBand;
    dim3 egrid(((nwords - n_int) * ebands + 127) / 128, 1, B);
    gauss7_kernel<2><<<egrid, block, 0, st>>>(src, dst, t0, t1, t2, t3, n_int);
  } else {
    dim3 grid((nwords + 31) / 32, (nbands + 3) / 4, B);
    gauss7_kernel<0><<<grid, block, 0, st>>>(src, dst, t0, t1, t2, t3, 0);
  }
}

void launch_gauss7(const OrbGeometry& geo, ImgBatch in0, int B, cudaStream_t st) {
  for (int l = 0; l < geo.nlevels; ++l) {
    ImgBatch src = l == 0 ? in0
                          : ImgBatch{geo.level[l].ptr, geo.level[l].pitch, geo.level[l].frame_stride, geo.level[l].w,
                                     geo.level[l].h};
    launch_blur7(src, geo.blur[l], 18, 34, 48, 56, B, st);
  }
}

}  // namespace psl
