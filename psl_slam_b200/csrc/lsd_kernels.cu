// LSD on the device: data-parallel prologue (7x7 sigma-0.75 blur, exact 0.8x resize, 2x2 gradient,
// stable bin ordering of the seed pixels) and the sequential core (lsd_core.cuh), one frame per warp.
#include <cub/device/device_segmented_radix_sort.cuh>

#include "line_kernels.cuh"
#include "lsd_core.cuh"
#include "orb_kernels.cuh"

namespace psl {

// cv::resize(fx = fy = 0.8, INTER_LINEAR_EXACT) on CV_8U: Q8 weights, Q8.8 row interpolation,
// (v + 2^15) >> 16 (tables built on the host from scale*(d+0.5)-0.5 in fp64).
__global__ void __launch_bounds__(256)
    resize_exact_kernel(const uint8_t* __restrict__ src, int pitch, int64_t fs, uint8_t* __restrict__ dst, int Ws, int Hs,
                        const short2* __restrict__ xtab, const short2* __restrict__ ytab) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  if (x >= Ws) return;
  const short2 tx = __ldg(xtab + x), ty = __ldg(ytab + y);
  const uint8_t* r0 = src + (size_t)b * fs + (size_t)ty.x * pitch;
  uint32_t h0 = tx.y >= 0 ? (uint32_t)r0[tx.x] * (256 - tx.y) + (uint32_t)r0[tx.x + 1] * tx.y : (uint32_t)r0[tx.x] * 256;
  uint32_t v;
  if (ty.y >= 0) {
    const uint8_t* r1 = r0 + pitch;
    const uint32_t h1 = tx.y >= 0 ? (uint32_t)r1[tx.x] * (256 - tx.y) + (uint32_t)r1[tx.x + 1] * tx.y : (uint32_t)r1[tx.x] * 256;
    v = h0 * (256 - ty.y) + h1 * ty.y;
  } else {
    v = h0 * 256;
  }
  dst[((size_t)b * Hs + y) * Ws + x] = (uint8_t)((v + 32768u) >> 16);
}

// ll_angle: gradient on the 2x2 stencil, angle in degrees (fastAtan2), squared norm, per-frame maximum
// and the number of seed-capable pixels per row.  One warp per row.
__global__ void __launch_bounds__(128)
    lsd_gradient_kernel(const uint8_t* __restrict__ scaled, int Ws, int Hs, float* __restrict__ deg,
                        int32_t* __restrict__ n2, uint8_t* __restrict__ used, int32_t* __restrict__ max_n2,
                        int32_t* __restrict__ row_cnt, double rho) {
  const int lane = threadIdx.x & 31, y = blockIdx.x * 4 + (threadIdx.x >> 5), b = blockIdx.y;
  if (y >= Hs) return;
  const uint8_t* r0 = scaled + ((size_t)b * Hs + y) * Ws;
  const uint8_t* r1 = r0 + Ws;
  const size_t base = ((size_t)b * Hs + y) * Ws;
  int cnt = 0, mx = -1;
  for (int x = lane; x < Ws; x += 32) {
    float d = lsd::kNotDefDeg;
    int q = 0;
    if (y < Hs - 1 && x < Ws - 1) {
      const int DA = (int)r1[x + 1] - (int)r0[x], BC = (int)r0[x + 1] - (int)r1[x];
      const int gx = DA + BC, gy = DA - BC;
      q = gx * gx + gy * gy;
      if (!(sqrt((double)q / 4.0) <= rho)) {
        d = lsd::fast_atan2((float)gx, (float)-gy);
        ++cnt;
        mx = max(mx, q);
      }
    }
    deg[base + x] = d;
    n2[base + x] = q;
    used[base + x] = 0;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if (lane == 0) {
    row_cnt[(size_t)b * Hs + y] = cnt;
    if (mx >= 0) atomicMax(max_n2 + b, mx);
  }
}

// exclusive scan of the per-row counts of every frame (one warp per frame)
__global__ void __launch_bounds__(32)
    lsd_row_scan_kernel(int32_t* __restrict__ row_cnt, int Hs, int npx, int32_t* __restrict__ n_def,
                        int32_t* __restrict__ seg_begin, int32_t* __restrict__ seg_end) {
  const int lane = threadIdx.x, b = blockIdx.x;
  int32_t* rc = row_cnt + (size_t)b * Hs;
  int carry = 0;
  for (int base = 0; base < Hs; base += 32) {
    const int v = base + lane < Hs ? rc[base + lane] : 0;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    if (base + lane < Hs) rc[base + lane] = carry + inc - v;
    carry += __shfl_sync(0xffffffffu, inc, 31);
  }
  if (lane == 0) {
    n_def[b] = carry;
    seg_begin[b] = b * npx;
    seg_end[b] = b * npx + carry;
  }
}

// (bin, pixel) pairs of the seed-capable pixels in raster order (ordered compaction, one warp per row)
__global__ void __launch_bounds__(128)
    lsd_keys_kernel(const float* __restrict__ deg, const int32_t* __restrict__ n2, int Ws, int Hs,
                    const int32_t* __restrict__ max_n2, const int32_t* __restrict__ row_off,
                    uint16_t* __restrict__ key, uint32_t* __restrict__ val) {
  const int lane = threadIdx.x & 31, y = blockIdx.x * 4 + (threadIdx.x >> 5), b = blockIdx.y;
  if (y >= Hs - 1) return;
  const size_t npx = (size_t)Ws * Hs, base = (size_t)b * npx + (size_t)y * Ws;
  const int m = max_n2[b];
  const double max_grad = m >= 0 ? sqrt((double)m / 4.0) : -1.0;
  const double bin_coef = max_grad > 0 ? 1023.0 / max_grad : 0.0;
  int pos = row_off[(size_t)b * Hs + y];
  for (int x0 = 0; x0 < Ws - 1; x0 += 32) {
    const int x = x0 + lane;
    const bool def = x < Ws - 1 && deg[base + x] != lsd::kNotDefDeg;
    const unsigned bal = __ballot_sync(0xffffffffu, def);
    if (def) {
      const int p = pos + __popc(bal & ((1u << lane) - 1u));
      key[(size_t)b * npx + p] = (uint16_t)(int)(sqrt((double)n2[base + x] / 4.0) * bin_coef);
      val[(size_t)b * npx + p] = (uint32_t)(y * Ws + x);
    }
    pos += __popc(bal);
  }
}

// the sequential core: one frame per warp, lane 0 walks the seed list
__global__ void __launch_bounds__(32)
    lsd_core_kernel(LineBuffers L, uint32_t* __restrict__ status) {
  if (threadIdx.x != 0) return;
  const int b = blockIdx.x;
  const size_t npx = (size_t)L.Ws * L.Hs;
  lsd::Frame f;
  f.W = L.Ws;
  f.H = L.Hs;
  f.deg = L.deg + b * npx;
  f.n2 = L.n2 + b * npx;
  f.used = L.used + b * npx;
  f.reg = L.reg + b * npx;
  f.seeds = L.val_out + b * npx;
  f.n_seeds = L.n_def[b];
  f.min_reg_size = L.min_reg_size;
  f.out = L.raw + (size_t)b * L.raw_cap * 4;
  f.cap = L.raw_cap;
  const int n = lsd::detect(f);
  L.n_raw[b] = n < L.raw_cap ? n : L.raw_cap;
  if (n > L.raw_cap) atomicOr(status, kStatLineRaw);
}

size_t lsd_sort_temp_bytes(int items_per_frame, int frames) {
  size_t bytes = 0;
  cub::DeviceSegmentedRadixSort::SortPairsDescending(nullptr, bytes, (const uint16_t*)nullptr, (uint16_t*)nullptr,
                                                     (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                                     (int64_t)items_per_frame * frames, frames, (const int32_t*)nullptr,
                                                     (const int32_t*)nullptr, 0, 10);
  return bytes;
}

void launch_lsd_prologue(const LineBuffers& L, ImgBatch in, int nb, cudaStream_t st) {
  const int npx = L.Ws * L.Hs;
  ImgBatchMut bl{L.blur, L.pitch, (int64_t)L.pitch * L.h, L.w, L.h};
  launch_blur7(in, bl, 0, 4, 56, 136, nb, st);  // GaussianBlur(7x7, sigma = 0.6 / 0.8)
  {
    dim3 grid((L.Ws + 255) / 256, L.Hs, nb);
    resize_exact_kernel<<<grid, 256, 0, st>>>(L.blur, L.pitch, (int64_t)L.pitch * L.h, L.scaled, L.Ws, L.Hs, L.xtab,
                                              L.ytab);
  }
  cudaMemsetAsync(L.max_n2, 0xFF, (size_t)nb * sizeof(int32_t), st);  // -1
  const double rho = 2.0 / sin(lsd::kPi * lsd::kAngTh / 180);
  dim3 rows((L.Hs + 3) / 4, nb);
  lsd_gradient_kernel<<<rows, 128, 0, st>>>(L.scaled, L.Ws, L.Hs, L.deg, L.n2, L.used, L.max_n2, L.row_cnt, rho);
  lsd_row_scan_kernel<<<nb, 32, 0, st>>>(L.row_cnt, L.Hs, npx, L.n_def, L.seg_begin, L.seg_end);
  lsd_keys_kernel<<<rows, 128, 0, st>>>(L.deg, L.n2, L.Ws, L.Hs, L.max_n2, L.row_cnt, L.key_in, L.val_in);
}

// stable: bins descending, raster order inside a bin (identical to cv2 4.13 on every golden)
void launch_lsd_order(const LineBuffers& L, int nb, cudaStream_t st) {
  const int npx = L.Ws * L.Hs;
  size_t tmp = L.sort_tmp_bytes;
  cub::DeviceSegmentedRadixSort::SortPairsDescending(L.sort_tmp, tmp, L.key_in, L.key_out, L.val_in, L.val_out,
                                                     (int64_t)npx * nb, nb, L.seg_begin, L.seg_end, 0, 10, st);
}

void launch_lsd_core(const LineBuffers& L, int nb, uint32_t* status, cudaStream_t st) {
  lsd_core_kernel<<<nb, 32, 0, st>>>(L, status);
}

}  // namespace psl
