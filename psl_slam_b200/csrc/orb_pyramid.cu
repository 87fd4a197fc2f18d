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

void launch_resize(const ImgBatch& src, const ImgBatchMut& dst, const ResizeTables& t, int B, cudaStream_t st) {
  const int n4 = (dst.w + 3) >> 2;
  dim3 grid((n4 * dst.h + 255) / 256, B);
  resize_kernel<<<grid, 256, 0, st>>>(src, dst, t.xt, t.yt);
}

// ---------------------------------------------------------------------------------------------
// K5: 7x7 sigma=2 blur, Q8 taps [18,34,48,56,48,34,18], REFLECT_101, dst = (v + 2^15) >> 16.
// Tile 128x16 outputs per CTA; input tile with 3-px halo staged in shared memory, row pass to
// u16 in shared memory, column pass straight to a uchar4 store.
// ---------------------------------------------------------------------------------------------
constexpr int kGTW = 128, kGTH = 16, kGR = 3;

__device__ __forceinline__ int reflect101(int i, int n) {
  // valid for -n < i < 2n-1, which holds for a 3-px halo on any level the extractor accepts
  i = i < 0 ? -i : i;
  return i >= n ? 2 * (n - 1) - i : i;
}

__global__ void __launch_bounds__(256) gauss7_kernel(ImgBatch src, ImgBatchMut dst) {
  __shared__ uint8_t s_in[kGTH + 2 * kGR][kGTW + 2 * kGR + 2];
  __shared__ uint16_t s_row[kGTH + 2 * kGR][kGTW];
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * kGTW, y0 = blockIdx.y * kGTH;
  const uint8_t* __restrict__ S = src.ptr + (size_t)b * src.frame_stride;
  constexpr int IW = kGTW + 2 * kGR, IH = kGTH + 2 * kGR;
  for (int i = threadIdx.x; i < IW * IH; i += 256) {
    const int ty = i / IW, tx = i - ty * IW;
    const int gx = reflect101(x0 + tx - kGR, src.w), gy = reflect101(y0 + ty - kGR, src.h);
    // tiles hanging over the right/bottom edge read (harmless) reflected pixels
    s_in[ty][tx] = __ldg(S + (size_t)min(max(gy, 0), src.h - 1) * src.pitch + min(max(gx, 0), src.w - 1));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < IH * kGTW; i += 256) {
    const int ty = i / kGTW, tx = i - ty * kGTW;
    const uint8_t* p = &s_in[ty][tx];
    s_row[ty][tx] = (uint16_t)(18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kGTH * (kGTW / 4); i += 256) {
    const int ty = i / (kGTW / 4), tx = (i - ty * (kGTW / 4)) * 4;
    const int gy = y0 + ty, gx = x0 + tx;
    if (gy >= src.h || gx >= src.w) continue;
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t v = 18u * (s_row[ty][tx + k] + s_row[ty + 6][tx + k]) +
                         34u * (s_row[ty + 1][tx + k] + s_row[ty + 5][tx + k]) +
                         48u * (s_row[ty + 2][tx + k] + s_row[ty + 4][tx + k]) + 56u * s_row[ty + 3][tx + k];
      packed |= ((v + 32768u) >> 16) << (8 * k);
    }
    *reinterpret_cast<uint32_t*>(dst.ptr + (size_t)b * dst.frame_stride + (size_t)gy * dst.pitch + gx) = packed;
  }
}

void launch_gauss7(const OrbGeometry& geo, ImgBatch in0, int B, cudaStream_t st) {
  for (int l = 0; l < geo.nlevels; ++l) {
    ImgBatch src = l == 0 ? in0
                          : ImgBatch{geo.level[l].ptr, geo.level[l].pitch, geo.level[l].frame_stride, geo.level[l].w,
                                     geo.level[l].h};
    const ImgBatchMut& dst = geo.blur[l];
    dim3 grid((dst.w + kGTW - 1) / kGTW, (dst.h + kGTH - 1) / kGTH, B);
    gauss7_kernel<<<grid, 256, 0, st>>>(src, dst);
  }
}

}  // namespace psl
