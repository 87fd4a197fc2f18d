// LSD on the device: data-parallel prologue (7x7 sigma-0.75 blur, exact 0.8x resize, 2x2 gradient,
// stable bin ordering of the seed pixels) and the sequential core (lsd_core.cuh), one frame per warp.
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

// The same arithmetic with whole-word loads (the pyramid's resize_words_kernel, orb_pyramid.cu, with the exact
// variant's Q8 weights and rounding): a thread owns 4 adjacent outputs and walks kXBand output rows; per source row
// the 3 aligned words that hold its taps, one IDP.2A per output for the row pass; a source row that serves two
// consecutive output rows (three in four at 0.8x) is interpolated once.
constexpr int kXBand = 8;

__global__ void __launch_bounds__(256)
    resize_exact_words_kernel(const uint8_t* __restrict__ src, int pitch, int64_t fs, uint8_t* __restrict__ dst, int Ws,
                              int Hs, const uint4* __restrict__ xw, const uint32_t* __restrict__ xo,
                              const short2* __restrict__ ytab) {
  const int n4 = Ws >> 2, nbands = (Hs + kXBand - 1) / kXBand;   // Ws is a multiple of 4 here
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n4 * nbands) return;
  const int band = idx / n4, g = idx - band * n4, b = blockIdx.y;
  const uint4 wt = __ldg(xw + g);
  const uint32_t xo_g = __ldg(xo + g);
  const int wb = (int)(xo_g & 0xFFFFu), nwords = pitch >> 2;
  const bool in1 = wb + 1 < nwords, in2 = wb + 2 < nwords;
  const uint32_t* __restrict__ S = reinterpret_cast<const uint32_t*>(src + (size_t)b * fs) + wb;
  uint8_t* __restrict__ D = dst + (size_t)b * Hs * Ws + 4 * g;
  const uint32_t w4[4] = {wt.x, wt.y, wt.z, wt.w};
  int kept_row = -1;
  uint32_t kept[4] = {0, 0, 0, 0};
  const int y_end = min((band + 1) * kXBand, Hs);
  for (int y = band * kXBand; y < y_end; ++y) {
    const short2 ty = __ldg(ytab + y);
    uint32_t h0[4], h1[4];
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      uint32_t* hh = pass == 0 ? h0 : h1;
      const int sy = ty.x + pass;
      if (pass == 1 && ty.y < 0) break;       // no lower tap: v = h0 * 256
      if (pass == 0 && sy == kept_row) {
#pragma unroll
        for (int k = 0; k < 4; ++k) hh[k] = kept[k];
        continue;
      }
      const uint32_t* row = S + (size_t)sy * nwords;
      const uint32_t W0 = __ldg(row), W1 = in1 ? __ldg(row + 1) : 0u, W2 = in2 ? __ldg(row + 2) : 0u;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t off = (xo_g >> (16 + 4 * k)) & 0xFu;
        const uint32_t lo = off < 4u ? W0 : W1, hi = off < 4u ? W1 : W2;
        hh[k] = __dp2a_lo(w4[k], __funnelshift_r(lo, hi, 8u * (off & 3u)), 0u);
      }
    }
    uint32_t packed = 0;
    if (ty.y >= 0) {
      const uint32_t fy = (uint32_t)ty.y;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        packed |= ((h0[k] * (256u - fy) + h1[k] * fy + 32768u) >> 16) << (8 * k);
        kept[k] = h1[k];
      }
      kept_row = ty.x + 1;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        packed |= ((h0[k] * 256u + 32768u) >> 16) << (8 * k);
        kept[k] = h0[k];
      }
      kept_row = ty.x;
    }
    *reinterpret_cast<uint32_t*>(D + (size_t)y * Ws) = packed;
  }
}

constexpr int lsdw_kGMask = 0x000FFFFF;   // the integer gradient of the pixel, (gy + 510) << 10 | (gx + 510)
constexpr int lsdw_kUnavail = (int)0x80000000;  // same word: not available to region growing (NOTDEF from the start, or USED)
__host__ __device__ __forceinline__ int lsdw_pack_g(int gx, int gy) { return ((gy + 510) << 10) | (gx + 510); }
// gx^2 + gy^2 of a record's int word (modgrad = sqrt(that / 4.0))
__device__ __forceinline__ int lsdw_n2(int w) {
  const int gx = (w & 1023) - 510, gy = ((w >> 10) & 1023) - 510;
  return gx * gx + gy * gy;
}

// ll_angle: gradient on the 2x2 stencil, angle in degrees (fastAtan2), squared norm, per-frame maximum
// and the number of seed-capable pixels per row.  One warp per row.  Each pixel gets one 16-byte record
// (angle in degrees | cos | sin | integer gradient + flag "not available": NOTDEF or USED) so that region growing needs a single LDG.128 per
// neighbour: cos / sin are the fp32 values region_grow adds to its running sums, (float)cos((double)(float)angle)
// (the reference calls cos(float) -> pinned to fp64 evaluation), computed here in parallel instead
// of inside the sequential loop.
// The record of a pixel depends on its integer gradient (gx, gy) only, |gx|, |gy| <= 510: all 1021^2 records are
// tabulated once per context (16.7 MB, L2-resident) so that the per-frame gradient kernel is a gather instead of
// fastAtan2 + fp64 cos / sin per pixel.
constexpr int kLutSide = 1021, kLutOff = 510;

__device__ __forceinline__ float4 lsd_record(int gx, int gy, double rho, float2& seed) {
  float4 rec = make_float4(lsd::kNotDefDeg, 0.f, 0.f, 0.f);
  const int q = gx * gx + gy * gy;
  rec.w = __int_as_float(lsdw_pack_g(gx, gy) | lsdw_kUnavail);
  seed = make_float2(0.f, 0.f);
  if (!(sqrt((double)q / 4.0) <= rho)) {
    rec.w = __int_as_float(lsdw_pack_g(gx, gy));
    rec.x = lsd::fast_atan2((float)gx, (float)-gy);
    const double a = (double)(float)((double)rec.x * lsd::kDegToRad);
    rec.y = (float)cos(a);
    rec.z = (float)sin(a);
    // what region_grow starts its running sums with when this pixel is the seed: cos / sin of the fp64 angle
    double sd, cd;
    sincos((double)rec.x * lsd::kDegToRad, &sd, &cd);
    seed = make_float2((float)cd, (float)sd);
  }
  return rec;
}

__global__ void __launch_bounds__(256) lsd_lut_kernel(float4* __restrict__ lut, float2* __restrict__ seed_lut, double rho) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kLutSide * kLutSide) return;
  const int gy = i / kLutSide - kLutOff, gx = i - (gy + kLutOff) * kLutSide - kLutOff;
  float2 sd;
  lut[i] = lsd_record(gx, gy, rho, sd);
  seed_lut[i] = sd;
}

// pass 1 over the 0.8x image: seed-capable pixels per row and the frame's largest squared gradient
__global__ void __launch_bounds__(128)
    lsd_count_kernel(const uint8_t* __restrict__ scaled, int Ws, int Hs, int32_t* __restrict__ max_n2,
                     int32_t* __restrict__ row_cnt, int q_undef) {
  const int lane = threadIdx.x & 31, y = blockIdx.x * 4 + (threadIdx.x >> 5), b = blockIdx.y;
  if (y >= Hs) return;
  const uint8_t* r0 = scaled + ((size_t)b * Hs + y) * Ws;
  const uint8_t* r1 = r0 + Ws;
  int cnt = 0, mx = -1;
  auto px = [&](int a, int b1, int c, int d) {   // a = r0[x], b1 = r0[x + 1], c = r1[x], d = r1[x + 1]
    const int DA = d - a, BC = b1 - c;
    const int gx = DA + BC, gy = DA - BC;
    const int q = gx * gx + gy * gy;
    if (q > q_undef) {
      ++cnt;
      mx = max(mx, q);
    }
  };
  if (y < Hs - 1) {
    if ((Ws & 3) == 0 && ((reinterpret_cast<uintptr_t>(scaled)) & 3) == 0) {
      // whole words: a lane owns 4 adjacent pixels, the fifth byte of its window is the first of the next word
      for (int x = 4 * lane; x < Ws; x += 128) {
        const uint32_t a = *reinterpret_cast<const uint32_t*>(r0 + x), c = *reinterpret_cast<const uint32_t*>(r1 + x);
        const bool last = x + 4 >= Ws;
        const uint32_t a4 = last ? 0u : (uint32_t)r0[x + 4], c4 = last ? 0u : (uint32_t)r1[x + 4];
        const int A[5] = {(int)(a & 255u), (int)((a >> 8) & 255u), (int)((a >> 16) & 255u), (int)(a >> 24), (int)a4};
        const int Cc[5] = {(int)(c & 255u), (int)((c >> 8) & 255u), (int)((c >> 16) & 255u), (int)(c >> 24), (int)c4};
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (!(last && k == 3)) px(A[k], A[k + 1], Cc[k], Cc[k + 1]);   // the last column of a row has no gradient
      }
    } else {
      for (int x = lane; x < Ws - 1; x += 32) px((int)r0[x], (int)r0[x + 1], (int)r1[x], (int)r1[x + 1]);
    }
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

// pass 2: the pixel records (a gather from the table) and, with the row offsets and the frame maximum known, the
// (bin, pixel) pairs of the seed-capable pixels in raster order (ordered compaction, one warp per row)
__global__ void __launch_bounds__(128)
    lsd_gradient_kernel(const uint8_t* __restrict__ scaled, int Ws, int Hs, const float4* __restrict__ lut,
                        float4* __restrict__ pix, const int32_t* __restrict__ max_n2,
                        const int32_t* __restrict__ row_off, uint16_t* __restrict__ key, uint32_t* __restrict__ val,
                        int q_undef) {
  const int lane = threadIdx.x & 31, y = blockIdx.x * 4 + (threadIdx.x >> 5), b = blockIdx.y;
  if (y >= Hs) return;
  const uint8_t* r0 = scaled + ((size_t)b * Hs + y) * Ws;
  const uint8_t* r1 = r0 + Ws;
  const size_t npx = (size_t)Ws * Hs, base = (size_t)b * npx + (size_t)y * Ws;
  const int m = max_n2[b];
  const double max_grad = m >= 0 ? sqrt((double)m / 4.0) : -1.0;
  const double bin_coef = max_grad > 0 ? 1023.0 / max_grad : 0.0;
  int pos = row_off[(size_t)b * Hs + y];
  for (int x0 = 0; x0 < Ws; x0 += 32) {
    const int x = x0 + lane;
    float4 rec = make_float4(lsd::kNotDefDeg, 0.f, 0.f, __int_as_float(lsdw_pack_g(0, 0) | lsdw_kUnavail));
    bool def = false;
    int q = 0;
    if (y < Hs - 1 && x < Ws - 1) {
      const int DA = (int)r1[x + 1] - (int)r0[x], BC = (int)r0[x + 1] - (int)r1[x];
      const int gx = DA + BC, gy = DA - BC;
      q = gx * gx + gy * gy;
      rec.w = __int_as_float(lsdw_pack_g(gx, gy) | lsdw_kUnavail);
      if (q > q_undef) {  // sqrt(q / 4) > rho  <=>  q > q_undef (largest q with sqrt(q / 4.0) <= rho, found on the host)
        rec = __ldg(lut + (gy + kLutOff) * kLutSide + (gx + kLutOff));
        def = true;
      }
    }
    if (x < Ws) pix[base + x] = rec;
    const unsigned bal = __ballot_sync(0xffffffffu, def);
    if (def) {
      const int p = pos + __popc(bal & ((1u << lane) - 1u));
      key[(size_t)b * npx + p] = (uint16_t)(int)(sqrt((double)q / 4.0) * bin_coef);
      val[(size_t)b * npx + p] = (uint32_t)(y * Ws + x);
    }
    pos += __popc(bal);
  }
}

void launch_lsd_lut(float4* lut, float2* seed_lut, cudaStream_t st) {
  const double rho = 2.0 / sin(lsd::kPi * lsd::kAngTh / 180);
  lsd_lut_kernel<<<(kLutSide * kLutSide + 255) / 256, 256, 0, st>>>(lut, seed_lut, rho);
}
size_t lsd_lut_bytes() { return (size_t)kLutSide * kLutSide * sizeof(float4); }
size_t lsd_seed_lut_bytes() { return (size_t)kLutSide * kLutSide * sizeof(float2); }

// exclusive scan of the per-row counts of every frame (one warp per frame)
__global__ void __launch_bounds__(32)
    lsd_row_scan_kernel(int32_t* __restrict__ row_cnt, int Hs, int32_t* __restrict__ n_def) {
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
  if (lane == 0) n_def[b] = carry;
}

// ---------------------------------------------------------------------------------------------------
// The sequential core, one frame per warp.  Region growing is order dependent (every accepted pixel moves
// the region angle the next neighbour is tested against), so the warp does not split the region; it
// evaluates the next <= 32 neighbour tests of the reference's loop at once (the 8 neighbours of up to
// four consecutive region points, lane order = loop order) and commits a whole step when that is provably
// what the scalar loop (lsd_core.cuh, which the CPU suite checks against the oracle) would have done; any
// other step replays the scalar loop.  The fp64 running sums of region2rect / refine are added in the
// reference's order, the extents are exact min/max reductions.
//
// The kernel is bound by instruction issue and instruction fetch (one warp per frame, about 27 warps per SM at
// different places of the code), so it is written for few and compact instructions: every device function below has
// exactly one inlined call site (the refine / reduce stages loop back into the same region_grow and region2rect),
// and what runs rarely — the scalar replay of a step, the preparation of a re-grow, the radius reduction, the exact
// seed terms — sits in __noinline__ functions outside the hot loops.
// ---------------------------------------------------------------------------------------------------
#ifdef PSL_LSD_STATS
__device__ unsigned long long g_lsd_stats[16];
#define LSD_STAT(i, v) do { const unsigned long long v__ = (unsigned long long)(v); if (lane == 0) atomicAdd(&g_lsd_stats[i], v__); } while (0)
#define LSD_T0(t) long long t = clock64()
#define LSD_T1(i, t) do { LSD_STAT(i, clock64() - t); } while (0)
#else
#define LSD_STAT(i, v) do { } while (0)
#define LSD_T0(t) do { } while (0)
#define LSD_T1(i, t) do { } while (0)
#endif

namespace lsdw {

using lsd::kDegToRad;
using lsd::kM2Pi;
using lsd::kM32Pi;
using lsd::kNotDefDeg;
using lsd::kPi;

constexpr int kRing = 512;   // region points kept in shared memory (the BFS frontier and small regions)
// a refuted guess leaves up to 32 stale entries behind the end of the list, i.e. over the oldest entries of the ring
constexpr int kRingValid = kRing - 32;
constexpr unsigned kFull = 0xffffffffu;

struct Frame {
  int W;
  float4* pix;      // records; W + 1 unavailable records precede the frame (the last row of the previous frame or the pad)
  uint32_t* reg;    // region points in HBM, packed y << 16 | x
  uint32_t* ring;   // the last kRing of them in shared memory
  double* terms;    // [3][32] staging of the fp64 terms that are summed in list order
};

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(kFull, v, src); }
__device__ __forceinline__ int* flags_of(const Frame& f, int idx) { return reinterpret_cast<int*>(f.pix + idx) + 3; }
// region point i of a region of (current) size n
__device__ __forceinline__ uint32_t reg_at(const Frame& f, int i, int n) {
  if (n <= kRingValid) return f.ring[i & (kRing - 1)];   // the whole region is in the ring (uniform test, nearly always)
  return (n - i <= kRingValid) ? f.ring[i & (kRing - 1)] : f.reg[i];
}
__device__ __forceinline__ int lin_of(const Frame& f, uint32_t c) { return (int)(c >> 16) * f.W + (int)(c & 0xFFFFu); }

__device__ __forceinline__ bool aligned_deg(float deg, double theta, double prec) {
  double n_theta = theta - (double)deg * kDegToRad;
  if (n_theta < 0) n_theta = -n_theta;
  if (n_theta > kM32Pi) {
    n_theta -= kM2Pi;
    if (n_theta < 0) n_theta = -n_theta;
  }
  return n_theta <= prec;
}

// One neighbour test: the pixel `off` away from region point c.  A region point has a defined angle, so it is not in
// the last row / column (ll_angle leaves them NOTDEF) and its neighbours can leave the image only at the top / left:
// x = -1 lands on the last pixel of the row above and y = -1 on the row before the frame, all of them unavailable
// records (NOTDEF carries the same flag as USED), so no bounds test is needed.
struct Nbr {
  int lin;        // pixel index
  uint32_t npk;   // packed y << 16 | x (meaningless for a neighbour outside the image, which is never accepted)
  float4 rec;     // pixel record; rec.w < 0 (as an int): not available (USED, no defined angle, or an idle lane)
};
// Whether the pixel was available when it was loaded.  Deliberately NOT evaluated inside load_nbr: the first use of a
// loaded register is where the warp waits for the load, and the step's verification is meant to run before that.
__device__ __forceinline__ bool avail(const Nbr& b) { return __float_as_int(b.rec.w) >= 0; }

__device__ __forceinline__ Nbr load_nbr(const Frame& f, bool active, uint32_t c, int off, uint32_t offpk) {
  Nbr b;
  b.lin = lin_of(f, c) + off;
  b.npk = c + offpk;
  b.rec = make_float4(0.f, 0.f, 0.f, __int_as_float(lsdw_kUnavail));
  if (active) b.rec = f.pix[b.lin];   // (ld.global.L2::128B / ::256B measured: no difference)
  return b;
}

// thresholds of the quick alignment test for a precision `prec`.
// The reference tests |angle_i - reg_angle| <= prec with reg_angle = fastAtan2(sums).  fastAtan2 is within 0.0096 deg
// of the true angle (measured over 2e7 directions) and the fp32 dot product of a record's (cos, sin) with the sums is
// good to 1e-6 relative, i.e. 2e-4 deg at 22.5 deg; with a band of +/- 0.03 deg around prec the outcome of the
// reference's test is certain outside the band, and only a neighbour inside it takes the exact path.
struct Quick {
  bool quick;
  float chi2, clo2;
};

__device__ __noinline__ Quick make_quick(double prec) {
  const double band = 0.03 * kDegToRad;
  Quick q;
  q.quick = prec + band < 1.5 && prec > band;   // both cosines positive: the test can be made on squares
  const float chi = q.quick ? (float)cos(prec - band) : 2.f;   // dot >= chi * |sum|: aligned for sure
  const float clo = q.quick ? (float)cos(prec + band) : 0.f;   // dot <= clo * |sum|: not aligned for sure
  q.chi2 = chi * chi;
  q.clo2 = clo * clo;
  return q;
}

// The reference seeds the running sums with (float)cos(reg_angle), (float)sin(reg_angle) of the seed's fp64 angle.
// Those values (`sterm`, from the seed table of the integer gradient) only matter once a second pixel joins; until then
// the quick test runs on the record's cos / sin (same direction to 1e-7).
struct Sums {
  float x, y;
  int n;
};

// The scalar loop over the lanes of one step (the reference's order of tests), for the steps the batched decision cannot
// prove.  Accepted pixels get their flag, list and ring entries here.  `exact` says whether the sums already are the
// reference's (false until the first accept of the region).
__device__ __noinline__
Sums step_sequential(Frame f, int lin, uint32_t npk, float4 rec, bool cand, Sums s, bool exact,
                                             float seed_deg, float2 sterm, double prec, Quick qk, int lane) {
  unsigned mask = __ballot_sync(kFull, cand);
  while (mask) {
    const float s2 = s.x * s.x + s.y * s.y;
    const float dot = rec.y * s.x + rec.z * s.y, dot2 = dot * dot;
    const bool sure = qk.quick && s2 > 1e-6f;   // opposing gradients cancelled: no direction to test against
    bool ok = cand && sure && dot > 0.f && dot2 >= qk.chi2 * s2;
    const bool maybe = cand && !ok && (!sure || (dot > 0.f && dot2 > qk.clo2 * s2));
    if (__ballot_sync(kFull, maybe) & mask) {
      const double reg_angle = exact ? (double)lsd::fast_atan2(s.y, s.x) * kDegToRad : (double)seed_deg * kDegToRad;
      if (maybe) ok = aligned_deg(rec.x, reg_angle, prec);
    }
    const unsigned am = __ballot_sync(kFull, ok) & mask;
    if (!am) break;
    const int j = __ffs(am) - 1;
    if (!exact) {
      s.x = sterm.x;
      s.y = sterm.y;
      exact = true;
    }
    if (lane == j) {
      *flags_of(f, lin) = __float_as_int(rec.w) | lsdw_kUnavail;
      f.reg[s.n] = npk;
      f.ring[s.n & (kRing - 1)] = npk;
    }
    s.n += 1;
    s.x += __shfl_sync(kFull, rec.y, j);
    s.y += __shfl_sync(kFull, rec.z, j);
    const int alin = __shfl_sync(kFull, lin, j);
    if (lin == alin) cand = false;            // the same pixel seen from another centre is now USED
    mask &= ~((2u << j) - 1u);                // tests before j were made (and failed) with the older sums
  }
  return s;
}

// region_grow; returns the region size, reg_angle out (defined only for regions of at least min_n points; smaller
// ones are dropped by every caller).
//
// The serial chain of the reference is  accept -> sums -> reg_angle = fastAtan2(sums) -> test next neighbour.
//  * The alignment test is first made on the sums themselves (cos_i * sumdx + sin_i * sumdy against
//    cos(prec -/+ band) * |sum|, see make_quick); only a neighbour inside the band needs the exact path.
//  * A whole step (<= 32 tests) is decided at once when that is provably what the scalar loop would do: guess the
//    accepted set A from the sums before the step, give every lane the sums it would see in the scalar loop (the
//    prefix over the lanes of A before it) and re-test; if every lane's outcome under its own prefix is certain and
//    reproduces A, then A is the scalar loop's result by induction over the lanes.  The prefix only feeds the
//    certain / uncertain classification; the running sums themselves are advanced by the accepted terms in lane
//    order, i.e. with the reference's fp32 rounding.
//  * The guess is committed speculatively (flags, ring and list entries) and the records of the next step's
//    neighbours are requested BEFORE the guess is verified, so the verification runs under the load latency and the
//    next step needs no second look at what this step accepted (its loads already see the flags).  A refuted guess
//    (a few percent of the steps) takes the flags back and replays the step with the scalar loop.
__device__ __forceinline__ int region_grow(const Frame& f, int seed_lin, uint32_t c0, float4 srec, float2 sterm,
                                           double& reg_angle, double prec, const Quick qk, int min_n, int lane) {
  // four region points per step, eight lanes each: the centre of a 3x3 neighbourhood is the region point itself
  // (USED, never a candidate), so the loop's nine tests are the eight below in the same order
  const int p = lane >> 3, k8 = lane & 7, k = k8 + (k8 >= 4);
  const int oy = k / 3 - 1, ox = k - 3 * (k / 3) - 1;
  const int off = oy * f.W + ox;
  const uint32_t offpk = (uint32_t)(oy * 65536 + ox);
  const unsigned lt = (1u << lane) - 1u;
  Sums s{srec.y, srec.z, 1};
  bool exact = false;
  if (lane == 0) {
    f.reg[0] = c0;
    f.ring[0] = c0;
    *flags_of(f, seed_lin) = __float_as_int(srec.w) | lsdw_kUnavail;
  }
  Nbr cur = load_nbr(f, p == 0, c0, off, offpk);   // none of the eight is the seed itself
  __syncwarp();
  int i = 0, m = 1;
  for (;;) {
    LSD_STAT(0, 1);
    LSD_STAT(1, m);
    const int i_next = i + m;   // first region point of the next step
    const float s2 = s.x * s.x + s.y * s.y;
    const bool fast = qk.quick && s2 > 1e-6f;
    const bool cand = avail(cur);
    const float dot = cur.rec.y * s.x + cur.rec.z * s.y, dot2 = dot * dot;
    const bool pos = cand && dot > 0.f;
    const bool yes0 = fast && pos && dot2 >= qk.chi2 * s2;
    const bool unsure0 = cand && !yes0 && (!fast || (pos && dot2 > qk.clo2 * s2));
    const unsigned A0 = __ballot_sync(kFull, yes0), U0 = __ballot_sync(kFull, unsure0);
    // guess: the lanes that pass under the sums before the step, first lane of every repeated pixel (the same pixel can
    // only be seen from two centres; all its lanes hold the same record, so they pass together)
    unsigned A = A0;
    bool first = true;
    if (A0 != 0u && m > 1) {
      if (yes0) first = (__match_any_sync(A0, cur.lin) & lt) == 0u;
      A = __ballot_sync(kFull, yes0 && first);
    }
    const bool inA = (A >> lane) & 1u;
    const int at = s.n + __popc(A & lt), n_spec = s.n + __popc(A);
    if (inA) {   // speculative commit
      *flags_of(f, cur.lin) = __float_as_int(cur.rec.w) | lsdw_kUnavail;
      f.ring[at & (kRing - 1)] = cur.npk;
      f.reg[at] = cur.npk;
    }
    __syncwarp();
    const int m2 = min(4, n_spec - i_next);   // <= 0: the list ends with this step
    Nbr nxt;
    {
      uint32_t c = f.ring[(i_next + p) & (kRing - 1)];
      if (n_spec - i_next > kRingValid) c = f.reg[min(i_next + p, n_spec - 1)];   // a frontier longer than the ring (rare, uniform)
      nxt = load_nbr(f, p < m2, c, off, offpk);
    }
    // verification, under the latency of those loads
    bool proven;
    float nx = s.x, ny = s.y;
    if (A == 0u) {
      proven = U0 == 0u;   // nobody passes under the current sums; certain for all only if no lane sits in the band
    } else {
      // the sums every lane would see in the scalar loop: the running sums advanced by the lanes of A before it, added
      // in lane order with the reference's fp32 rounding (lane 31 ends with the sums after the step)
      float Px = s.x, Py = s.y;
      if (!exact) {
        Px = sterm.x;
        Py = sterm.y;
      }
      for (unsigned a = A; a; a &= a - 1) {
        const int j = __ffs(a) - 1;
        const float vx = __shfl_sync(kFull, cur.rec.y, j), vy = __shfl_sync(kFull, cur.rec.z, j);
        if (lane > j) { Px += vx; Py += vy; }
      }
      const float d1 = cur.rec.y * Px + cur.rec.z * Py, q1 = Px * Px + Py * Py;
      const bool sure1 = q1 > 1e-6f;
      const bool yes1 = sure1 && d1 > 0.f && d1 * d1 >= qk.chi2 * q1;
      const bool no1 = sure1 && (d1 <= 0.f || d1 * d1 <= qk.clo2 * q1);
      // an earlier lane of A holds this pixel: USED by the time the scalar loop gets here
      const bool taken = yes0 && !first;
      const bool lane_ok = !cand || taken || (inA ? yes1 : no1);
      proven = __all_sync(kFull, lane_ok);
      // lane 31's prefix misses its own term when it is accepted itself (the last addition of the step)
      const bool last = (A >> 31) != 0u;
      nx = __shfl_sync(kFull, last ? Px + cur.rec.y : Px, 31);
      ny = __shfl_sync(kFull, last ? Py + cur.rec.z : Py, 31);
    }
    if (proven) {
      if (A != 0u) {
        s.x = nx;
        s.y = ny;
        s.n = n_spec;
        exact = true;
      }
      i = i_next;
      if (m2 <= 0) break;
      m = m2;
      cur = nxt;
    } else {
      LSD_STAT(7, 1);
      if (inA) *flags_of(f, cur.lin) = __float_as_int(cur.rec.w);   // take the guess back
      __syncwarp();
      s = step_sequential(f, cur.lin, cur.npk, cur.rec, cand, s, exact, srec.x, sterm, prec, qk, lane);
      exact = exact || s.n > 1;
      i = i_next;
      if (i >= s.n) break;
      __syncwarp();
      m = min(4, s.n - i);
      uint32_t c = 0;
      if (p < m) c = reg_at(f, i + p, s.n);
      cur = load_nbr(f, p < m, c, off, offpk);
    }
  }
  if (s.n >= min_n)
    reg_angle = exact ? (double)lsd::fast_atan2(s.y, s.x) * kDegToRad : (double)srec.x * kDegToRad;
  __syncwarp();
  return s.n;
}

struct Rect { double x1, y1, x2, y2, width, x, y, theta, dx, dy; };

__device__ __forceinline__ double modgrad(const Frame& f, int idx) {
  return sqrt((double)lsdw_n2(*flags_of(f, idx)) / 4.0);
}

// one ordered pass: the 32 terms of three running sums go through shared memory and lanes 0, 1, 2 each add one of
// them in list order (a broadcast read + one add per term)
__device__ __forceinline__ void ordered_add3(const Frame& f, double t0, double t1, double t2, int cnt, double& acc, int lane) {
  __syncwarp();
  f.terms[lane] = t0; f.terms[32 + lane] = t1; f.terms[64 + lane] = t2;
  __syncwarp();
  const double* chain = f.terms + 32 * (lane < 3 ? lane : 0);
  int k = 0;
  for (; k + 4 <= cnt; k += 4) {
    const double2 a = *reinterpret_cast<const double2*>(chain + k), b = *reinterpret_cast<const double2*>(chain + k + 2);
    acc += a.x; acc += a.y; acc += b.x; acc += b.y;
  }
  for (; k < cnt; ++k) acc += chain[k];
}

__device__ __forceinline__ void region2rect(const Frame& f, int n, double reg_angle, double prec, Rect& rec, int lane) {
  // weighted centroid: x += px * w ... in list order
  double acc = 0;
  double w0 = 0, w1 = 0;   // the weights of the first 64 points (most regions are shorter) for the second pass
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    double tx = 0, ty = 0, w = 0;
    if (i < n) {
      const uint32_t c = reg_at(f, i, n);
      const int px = (int)(c & 0xFFFFu), py = (int)(c >> 16);
      w = modgrad(f, py * f.W + px);
      tx = (double)px * w;
      ty = (double)py * w;
    }
    if (base == 0) w0 = w;
    if (base == 32) w1 = w;
    ordered_add3(f, tx, ty, w, min(32, n - base), acc, lane);
  }
  double x = shfl_d(acc, 0), y = shfl_d(acc, 1);
  const double sum = shfl_d(acc, 2);
  x /= sum;
  y /= sum;
  // get_theta: inertia matrix in list order
  acc = 0;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    double t1 = 0, t2 = 0, t3 = 0;
    if (i < n) {
      const uint32_t c = reg_at(f, i, n);
      const int px = (int)(c & 0xFFFFu), py = (int)(c >> 16);
      const double w = base == 0 ? w0 : base == 32 ? w1 : modgrad(f, py * f.W + px);
      const double dx = (double)px - x, dy = (double)py - y;
      t1 = dy * dy * w;
      t2 = dx * dx * w;
      t3 = -(dx * dy * w);   // Ixy -= term  ==  Ixy += -term
    }
    ordered_add3(f, t1, t2, t3, min(32, n - base), acc, lane);
  }
  const double Ixx = shfl_d(acc, 0), Iyy = shfl_d(acc, 1), Ixy = shfl_d(acc, 2);
  const double lambda = 0.5 * (Ixx + Iyy - sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
  double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)lsd::fast_atan2((float)(lambda - Ixx), (float)Ixy)
                                         : (double)lsd::fast_atan2((float)Ixy, (float)(lambda - Iyy));
  theta *= kDegToRad;
  double d = lsd::angle_diff_signed(theta, reg_angle);
  if (d < 0) d = -d;
  if (d > prec) theta += kPi;
  double dx, dy;
  sincos(theta, &dy, &dx);
  // extents: `if (l > l_max) l_max = l; else if (l < l_min) l_min = l;` with both starting at 0 is an
  // independent max and min (a value above the running max is positive, so it cannot lower the min)
  double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
  for (int i = lane; i < n; i += 32) {
    const uint32_t c = reg_at(f, i, n);
    const double rdx = (double)(int)(c & 0xFFFFu) - x, rdy = (double)(int)(c >> 16) - y;
    const double l = rdx * dx + rdy * dy, w = -rdx * dy + rdy * dx;
    l_max = l > l_max ? l : l_max;
    l_min = l < l_min ? l : l_min;
    w_max = w > w_max ? w : w_max;
    w_min = w < w_min ? w : w_min;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    double v = __shfl_xor_sync(kFull, l_max, o); l_max = v > l_max ? v : l_max;
    v = __shfl_xor_sync(kFull, l_min, o); l_min = v < l_min ? v : l_min;
    v = __shfl_xor_sync(kFull, w_max, o); w_max = v > w_max ? v : w_max;
    v = __shfl_xor_sync(kFull, w_min, o); w_min = v < w_min ? v : w_min;
  }
  rec.x1 = x + l_min * dx; rec.y1 = y + l_min * dy;
  rec.x2 = x + l_max * dx; rec.y2 = y + l_max * dy;
  rec.width = w_max - w_min;
  rec.x = x; rec.y = y; rec.theta = theta; rec.dx = dx; rec.dy = dy;
  if (rec.width < 1.0) rec.width = 1.0;
}

__device__ __forceinline__ double density_of(int n, const Rect& rec) {
  return (double)n / (sqrt(lsd::dist_sq(rec.x1, rec.y1, rec.x2, rec.y2)) * rec.width);
}

// refine, first part (rare: the density of a region's first rectangle is below the threshold): release the region's
// pixels and derive the tolerance tau of the re-grow from the angles near the seed.
__device__ __noinline__ double refine_tolerance(Frame f, int n, double width, int lane) {
  const uint32_t c0 = f.reg[0];
  const int sx = (int)(c0 & 0xFFFFu), sy = (int)(c0 >> 16);
  const double xc = (double)sx, yc = (double)sy, ang_c = (double)f.pix[sy * f.W + sx].x * kDegToRad;
  const double* chain = f.terms + 32 * (lane < 2 ? lane : 0);   // lane 0: sum, lane 1: s_sum
  double acc = 0;
  int cnt = 0;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    double a = 0, a2 = 0;   // points outside the radius add +0.0, which leaves the sums unchanged
    bool in = false;
    if (i < n) {
      const uint32_t c = reg_at(f, i, n);
      const int px = (int)(c & 0xFFFFu), py = (int)(c >> 16);
      const float4 r = f.pix[py * f.W + px];
      *flags_of(f, py * f.W + px) = __float_as_int(r.w) & 0x7FFFFFFF;
      if (sqrt(lsd::dist_sq(xc, yc, (double)px, (double)py)) < width) {
        a = lsd::angle_diff_signed((double)r.x * kDegToRad, ang_c);
        a2 = a * a;
        in = true;
      }
    }
    cnt += __popc(__ballot_sync(kFull, in));
    const int c32 = min(32, n - base);
    __syncwarp();
    f.terms[lane] = a; f.terms[32 + lane] = a2;
    __syncwarp();
    for (int k = 0; k < c32; ++k) acc += chain[k];
  }
  __syncwarp();
  const double sum = shfl_d(acc, 0), s_sum = shfl_d(acc, 1);
  const double mean_angle = sum / (double)cnt;
  return 2.0 * sqrt((s_sum - 2.0 * mean_angle * sum) / (double)cnt + mean_angle * mean_angle);
}

// reduce_region_radius, one round (rare): the swap-with-last removal defines the order of the surviving points (and
// with it the rounding of the next region2rect), so lane 0 replays it on the HBM list.  Returns the new size.
__device__ __noinline__ int reduce_once(Frame f, int n, double xc, double yc, double radSq, int lane) {
  if (lane == 0) {
    int m = n;
    for (int i = 0; i < m; ++i) {
      const uint32_t c = f.reg[i];
      const int px = (int)(c & 0xFFFFu), py = (int)(c >> 16);
      if (lsd::dist_sq(xc, yc, (double)px, (double)py) > radSq) {
        *flags_of(f, py * f.W + px) &= 0x7FFFFFFF;
        f.reg[i] = f.reg[m - 1];
        f.reg[m - 1] = c;
        --m;
        --i;
      }
    }
    n = m;
  }
  n = __shfl_sync(kFull, n, 0);
  __syncwarp();
  // the reordered tail goes back to the shared mirror (reg_at reads the last kRing points from it)
  for (int i = max(0, n - kRingValid) + lane; i < n; i += 32) f.ring[i & (kRing - 1)] = f.reg[i];
  __syncwarp();
  return n;
}

}  // namespace lsdw

constexpr int kCoreWarps = 4;  // frames per CTA (one per warp; the warps never synchronise with each other)

__global__ void __launch_bounds__(kCoreWarps * 32, 7)
    lsd_core_kernel(const __grid_constant__ LineBuffers L, int nb, uint32_t stride, uint32_t* __restrict__ status) {
  __shared__ uint32_t ring[kCoreWarps][lsdw::kRing];
  __shared__ __align__(16) double terms[kCoreWarps][96];
  // read once: under register pressure the compiler otherwise re-reads %tid (a slow special-register move) and rebuilds
  // the shared-memory and frame pointers inside the step loop
  unsigned tid_;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid_));
  const int wid = (int)(tid_ >> 5), slot = blockIdx.x * kCoreWarps + wid, lane = (int)(tid_ & 31u);
  if (slot >= nb) return;
  // frame of this warp: a fixed permutation of the batch (stride coprime to nb), so that the warps of one SM hold
  // frames from all over the sequence — neighbouring frames cost about the same, and an SM that got only expensive
  // ones would finish last
  const int b = (int)(((uint64_t)slot * stride) % (uint32_t)nb);
  const size_t npx = (size_t)L.Ws * L.Hs;
  float4* pix_ = L.pix + b * npx;
  uint32_t* reg_ = L.reg + b * npx;
  uint32_t* ring_ = ring[wid];
  asm volatile("" : "+l"(pix_), "+l"(reg_), "+l"(ring_));
  const lsdw::Frame f{L.Ws, pix_, reg_, ring_, terms[wid]};
  const uint32_t* seeds = L.val_out + b * npx;
  const int n_seeds = L.n_def[b];
  float* out = L.raw + (size_t)b * L.raw_cap * 4;
  const double prec0 = lsd::kPi * lsd::kAngTh / 180;
  const lsdw::Quick qk0 = lsdw::make_quick(prec0);
  const float inv_w = 1.0f / (float)L.Ws;
  int nseg = 0;
  LSD_T0(t_all);
  for (int s0 = 0; s0 < n_seeds; s0 += 32) {
    const int my = s0 + lane < n_seeds ? (int)seeds[s0 + lane] : -1;
    // a pixel that is USED now stays USED (only the pixels of the region being refined are ever released)
    const bool free_now = my >= 0 && *lsdw::flags_of(f, my >= 0 ? my : 0) >= 0;
    unsigned todo = __ballot_sync(lsdw::kFull, free_now);
    LSD_STAT(4, __popc(todo));
    // The eight neighbours of a seed that is about to be grown: the flag load brought in the sector of the seed's own
    // record; ask for the rest now, so that all but the first region of the chunk find their first neighbourhood on
    // the way (-1.5 % of the kernel).  x - 1 / y - 1 of a border seed land in the pad or the neighbouring row.
    if (free_now) {
      const float4* c = f.pix + my;
      asm volatile("prefetch.global.L1 [%0];" :: "l"(c - f.W - 1));
      asm volatile("prefetch.global.L1 [%0];" :: "l"(c - f.W + 1));
      asm volatile("prefetch.global.L1 [%0];" :: "l"(c - 1));
      asm volatile("prefetch.global.L1 [%0];" :: "l"(c + 1));
      asm volatile("prefetch.global.L1 [%0];" :: "l"(c + f.W - 1));
      asm volatile("prefetch.global.L1 [%0];" :: "l"(c + f.W + 1));
    }
    while (todo) {
      const int l = __ffs(todo) - 1;
      todo &= todo - 1;
      const int seed = __shfl_sync(lsdw::kFull, my, l);
      const float4 srec = f.pix[seed];
      const int sw = __float_as_int(srec.w);
      if (sw < 0) { LSD_STAT(5, 1); continue; }
      // the sums a region of this seed starts with once a second pixel joins (used by the first accept, far below)
      const float2 sterm = __ldg(L.seed_lut + ((sw >> 10) & 1023) * kLutSide + (sw & 1023));
      // seed / W without the integer division: seed < 2^24 and W <= 4096, so the float quotient is off by at most one
      int sy = (int)(((float)seed + 0.5f) * inv_w), sx = seed - sy * f.W;
      if (sx < 0) { --sy; sx += f.W; }
      else if (sx >= f.W) { ++sy; sx -= f.W; }
      const uint32_t c0 = ((uint32_t)sy << 16) | (uint32_t)sx;
      // LineSegmentDetectorImpl::flsd for one seed, with refine (LSD_REFINE_STD) and reduce_region_radius folded into two
      // loops around ONE region_grow and ONE region2rect:
      //   stage 0  grow with the global tolerance; too small -> next seed; rectangle; dense enough -> segment
      //   stage 1  (refine) release the pixels, re-grow with tolerance tau; < 2 points -> next seed; rectangle; dense -> segment
      //   stage 2+ (reduce_region_radius) shrink the radius by 0.75, drop the points outside; < 2 -> next seed; rectangle; ...
      double prec = prec0, reg_angle = 0, radSq = 0;
      lsdw::Quick qk = qk0;
      int stage = 0, n = 0;
      bool emit = false;
      for (;;) {
        if (stage < 2) {
          LSD_T0(t_g);
          n = lsdw::region_grow(f, seed, c0, srec, sterm, reg_angle, prec, qk, stage == 0 ? L.min_reg_size : 2, lane);
          LSD_T1(10, t_g);
          if (stage == 0) {
            LSD_STAT(13, 1);
            if (n == 1) LSD_STAT(6, 1);
            if (n < L.min_reg_size) { LSD_STAT(9, n); break; }
            LSD_STAT(2, 1);
            LSD_STAT(3, n);
          } else if (n < 2) {
            break;
          }
        }
        lsdw::Rect rec;
        LSD_T0(t_r);
        lsdw::region2rect(f, n, reg_angle, prec0, rec, lane);
        LSD_T1(11, t_r);
        const double density = lsdw::density_of(n, rec);
        if (density >= lsd::kDensityTh) {
          rec.x1 += 0.5; rec.y1 += 0.5; rec.x2 += 0.5; rec.y2 += 0.5;
          rec.x1 /= lsd::kScale; rec.y1 /= lsd::kScale; rec.x2 /= lsd::kScale; rec.y2 /= lsd::kScale;
          if (lane == 0 && nseg < L.raw_cap) {
            out[4 * nseg] = (float)rec.x1; out[4 * nseg + 1] = (float)rec.y1;
            out[4 * nseg + 2] = (float)rec.x2; out[4 * nseg + 3] = (float)rec.y2;
          }
          emit = true;
          break;
        }
        LSD_T0(t_f);
        if (stage == 0) {
          prec = lsdw::refine_tolerance(f, n, rec.width, lane);
          qk = lsdw::make_quick(prec);
          stage = 1;
        } else {
          const double xc = (double)sx, yc = (double)sy;
          if (stage == 1) {
            const double r1 = lsd::dist_sq(xc, yc, rec.x1, rec.y1), r2 = lsd::dist_sq(xc, yc, rec.x2, rec.y2);
            radSq = r1 > r2 ? r1 : r2;
            stage = 2;
          }
          radSq *= 0.75 * 0.75;
          n = lsdw::reduce_once(f, n, xc, yc, radSq, lane);
          if (n < 2) { LSD_T1(12, t_f); break; }
        }
        LSD_T1(12, t_f);
      }
      if (emit) ++nseg;
    }
  }
  LSD_T1(14, t_all);
#ifdef PSL_LSD_STATS
  if (lane == 0) atomicMax(&g_lsd_stats[15], (unsigned long long)(clock64() - t_all));   // slowest frame of the launch
#endif
  if (lane == 0) {
    L.n_raw[b] = nseg < L.raw_cap ? nseg : L.raw_cap;
    if (nseg > L.raw_cap) { atomicOr(status, kStatLineRaw); atomicMax(status + 1, (uint32_t)b + 1u); }
  }
}

#ifdef PSL_LSD_STATS
extern "C" void psl_lsd_stats(unsigned long long* out) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_lsd_stats, sizeof(g_lsd_stats));
  unsigned long long z[16] = {0};
  cudaMemcpyToSymbol(g_lsd_stats, z, sizeof(z));
}
#endif

void launch_lsd_prologue(const LineBuffers& L, ImgBatch in, int nb, cudaStream_t st) {
  const int npx = L.Ws * L.Hs;
  ImgBatchMut bl{L.blur, L.pitch, (int64_t)L.pitch * L.h, L.w, L.h};
  launch_blur7(in, bl, 0, 4, 56, 136, nb, st);  // GaussianBlur(7x7, sigma = 0.6 / 0.8)
  if (L.xw4 && (L.Ws & 3) == 0 && (L.pitch & 3) == 0) {
    const int nbands = (L.Hs + kXBand - 1) / kXBand;
    dim3 grid(((L.Ws >> 2) * nbands + 255) / 256, nb);
    resize_exact_words_kernel<<<grid, 256, 0, st>>>(L.blur, L.pitch, (int64_t)L.pitch * L.h, L.scaled, L.Ws, L.Hs, L.xw4,
                                                    L.xo4, L.ytab);
  } else {
    dim3 grid((L.Ws + 255) / 256, L.Hs, nb);
    resize_exact_kernel<<<grid, 256, 0, st>>>(L.blur, L.pitch, (int64_t)L.pitch * L.h, L.scaled, L.Ws, L.Hs, L.xtab,
                                              L.ytab);
  }
  cudaMemsetAsync(L.max_n2, 0xFF, (size_t)nb * sizeof(int32_t), st);  // -1
  // largest q = gx^2 + gy^2 whose modulus sqrt(q / 4.0) is still <= rho (the same fp64 comparison as ll_angle)
  const double rho = 2.0 / sin(lsd::kPi * lsd::kAngTh / 180);
  int q_undef = 0;
  while (sqrt((double)(q_undef + 1) / 4.0) <= rho) ++q_undef;
  dim3 rows((L.Hs + 3) / 4, nb);
  lsd_count_kernel<<<rows, 128, 0, st>>>(L.scaled, L.Ws, L.Hs, L.max_n2, L.row_cnt, q_undef);
  lsd_row_scan_kernel<<<nb, 32, 0, st>>>(L.row_cnt, L.Hs, L.n_def);
  lsd_gradient_kernel<<<rows, 128, 0, st>>>(L.scaled, L.Ws, L.Hs, L.lut, L.pix, L.max_n2, L.row_cnt, L.key_in, L.val_in,
                                            q_undef);
}

// ---------------------------------------------------------------------------------------------------
// Seed order: the 1024-bin ordering of the seed-capable pixels (ll_angle's bucket lists), bins descending, raster
// order inside a bin (stable: identical to cv2 4.13 on every golden).  One counting pass per frame, one CTA per frame:
// the (bin, pixel) pairs arrive in raster order; warp w owns the w-th contiguous eighth of them.
//   1. every warp counts its segment into its own 1024 counters (shared memory);
//   2. a block scan turns them into start offsets, bins descending, segments in order inside a bin;
//   3. every warp walks its segment again, 32 pairs per step: lanes with the same bin form a group (match.any), the
//      group takes its slots from the warp's counter of that bin with one add, ranks inside a group are lane order.
// Steps of a warp are sequential and segments are ordered, so equal bins keep the raster order.
// ---------------------------------------------------------------------------------------------------
constexpr int kOrdUnroll = 4;   // steps whose loads are in flight together (2: 3.53, 4: 3.55, 8: 3.63, none: 3.73 ms per 4096 frames)
constexpr int kOrdWarps = 32, kOrdBins = 1024;   // 128 KB of counters: one frame per SM, whose output (0.2 MB) stays in L2 while it fills

__global__ void __launch_bounds__(kOrdWarps * 32)
    lsd_seed_order_kernel(const uint16_t* __restrict__ key, const uint32_t* __restrict__ val, const int32_t* __restrict__ n_def,
                          int npx, uint32_t* __restrict__ out) {
  extern __shared__ uint32_t ord_smem[];
  uint32_t (*hist)[kOrdBins] = reinterpret_cast<uint32_t (*)[kOrdBins]>(ord_smem);
  __shared__ uint32_t wsum[kOrdWarps];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, b = blockIdx.x;
  const int n = n_def[b];
  const uint16_t* K = key + (size_t)b * npx;
  const uint32_t* V = val + (size_t)b * npx;
  uint32_t* O = out + (size_t)b * npx;
  for (int i = tid; i < kOrdWarps * kOrdBins; i += kOrdWarps * 32) (&hist[0][0])[i] = 0u;
  const int seg = (((n + kOrdWarps - 1) / kOrdWarps) + 31) & ~31;
  const int lo = min(w * seg, n), hi = min(lo + seg, n);
  __syncthreads();
  // (the loads of kOrdUnroll steps are issued together: a step is otherwise one DRAM round trip long)
  for (int k0 = lo; k0 < hi; k0 += 32 * kOrdUnroll) {   // (low-gradient bins are crowded: one add per group of equal bins, not per lane)
    unsigned kks[kOrdUnroll];
#pragma unroll
    for (int u = 0; u < kOrdUnroll; ++u) {
      const int k = k0 + 32 * u + lane;
      kks[u] = k < hi ? (unsigned)K[k] : 0xFFFFu;
    }
#pragma unroll
    for (int u = 0; u < kOrdUnroll; ++u) {
      const int k = k0 + 32 * u + lane;
      const unsigned kk = kks[u];
      const unsigned grp = __match_any_sync(0xffffffffu, kk);
      if (k < hi && lane == __ffs(grp) - 1) atomicAdd(&hist[w][kk], (uint32_t)__popc(grp));
    }
  }
  __syncthreads();
  {  // start offsets: position p = 1023 - bin; thread t owns p = kPer t .. kPer t + kPer - 1
    constexpr int kPer = kOrdBins / (kOrdWarps * 32);
    uint32_t tot = 0;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      const int bin = kOrdBins - 1 - (kPer * tid + j);
#pragma unroll
      for (int ww = 0; ww < kOrdWarps; ++ww) tot += hist[ww][bin];
    }
    uint32_t inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    uint32_t run = inc - tot;
    for (int ww = 0; ww < w; ++ww) run += wsum[ww];
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      const int bin = kOrdBins - 1 - (kPer * tid + j);
#pragma unroll
      for (int ww = 0; ww < kOrdWarps; ++ww) {
        const uint32_t c = hist[ww][bin];
        hist[ww][bin] = run;
        run += c;
      }
    }
  }
  __syncthreads();
  const unsigned lt = (1u << lane) - 1u;
  for (int k0 = lo; k0 < hi; k0 += 32 * kOrdUnroll) {
    unsigned kks[kOrdUnroll];
    uint32_t vs[kOrdUnroll];
#pragma unroll
    for (int u = 0; u < kOrdUnroll; ++u) {
      const int k = k0 + 32 * u + lane;
      kks[u] = k < hi ? (unsigned)K[k] : 0xFFFFu;   // the idle lanes of the last steps form a group of their own
      vs[u] = k < hi ? V[k] : 0u;
    }
#pragma unroll
    for (int u = 0; u < kOrdUnroll; ++u) {
      const bool act = k0 + 32 * u + lane < hi;
      const unsigned kk = kks[u];
      const unsigned grp = __match_any_sync(0xffffffffu, kk);
      const int leader = __ffs(grp) - 1;
      uint32_t base = 0;
      if (act && lane == leader) base = atomicAdd(&hist[w][kk], (uint32_t)__popc(grp));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (act) O[base + __popc(grp & lt)] = vs[u];
    }
  }
}

void launch_lsd_order(const LineBuffers& L, int nb, cudaStream_t st) {
  constexpr int kSmem = kOrdWarps * kOrdBins * (int)sizeof(uint32_t);
  static bool once = [] {
    cudaFuncSetAttribute(lsd_seed_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    return true;
  }();
  (void)once;
  lsd_seed_order_kernel<<<nb, kOrdWarps * 32, kSmem, st>>>(L.key_in, L.val_in, L.n_def, L.Ws * L.Hs, L.val_out);
}

void launch_lsd_core(const LineBuffers& L, int nb, uint32_t* status, cudaStream_t st) {
  auto gcd = [](uint32_t a, uint32_t b) { while (b) { const uint32_t t = a % b; a = b; b = t; } return a; };
  uint32_t stride = (uint32_t)(0.6180339887 * nb) | 1u;   // golden-ratio stride, made coprime to nb
  while (stride > 1 && gcd(stride, (uint32_t)nb) != 1) stride += 2;
  if (nb < 3) stride = 1;
  lsd_core_kernel<<<(nb + kCoreWarps - 1) / kCoreWarps, kCoreWarps * 32, 0, st>>>(L, nb, stride, status);
}

}  // namespace psl
