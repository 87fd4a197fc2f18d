// Line front end after LSD: the per-frame merge / KeyLine stage (line_core.cuh) and the LBD descriptor
// (opencv_contrib BinaryDescriptor::compute as called at add_src/LineExtractor.cpp:349-350; source-level spec
// vendored at Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp:219-261, 351-413, 1027-1373).
#include <math.h>

#include "line_kernels.cuh"
#include "orb_kernels.cuh"

namespace psl {

// ---------------------------------------------------------------------------------------------------
// clamp + optimizeAndMergeLines_lsd + top-N + KeyLines + line equations: one frame per warp, lane 0
// walks the reference's control flow (line_core.cuh explains why); the batch is the parallel axis.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
    line_post_kernel(LineBuffers L, int nfeatures, psl_keyline* __restrict__ kl, double* __restrict__ lineeq, int cap,
                     int32_t* __restrict__ n_out, uint32_t* __restrict__ status) {
  if (threadIdx.x != 0) return;
  const int b = blockIdx.x;
  const size_t rc = (size_t)L.raw_cap, o = (size_t)b * rc;
  line::MergeScratch S{L.raw_cap,      L.m_angles + o, L.m_length + o, L.m_order + o, L.m_tmp16 + o, L.m_nb + o * line::kNbCap,
                       L.m_nb_cnt + o, L.m_code + o,   L.m_check + o,  L.m_loc + o,   L.m_flag + o,  0};
  line::Seg* raw = reinterpret_cast<line::Seg*>(L.raw) + o;
  const int n = line::frame_lines(raw, L.n_raw[b], L.t1 + o, L.t2 + o, L.w, L.h, nfeatures, S, kl + (size_t)b * cap,
                                  lineeq + (size_t)b * cap * 3, cap);
  if (S.overflow) atomicOr(status, kStatLineNeighbours);
  if (n < 0) atomicOr(status, kStatOutOverflow);
  n_out[b] = n < 0 ? 0 : n;
}

void launch_line_post(const LineBuffers& L, int nb, int nfeatures, psl_keyline* kl, double* lineeq, int cap,
                      int32_t* n_out, uint32_t* status, cudaStream_t st) {
  line_post_kernel<<<nb, 32, 0, st>>>(L, nfeatures, kl, lineeq, cap, n_out, status);
}

// ---------------------------------------------------------------------------------------------------
// LBD
// ---------------------------------------------------------------------------------------------------
constexpr int kBandW = 7, kBands = 9, kLspH = kBandW * kBands;  // widthOfBand_, NUM_OF_BANDS, heightOfLSP

__constant__ float c_gaussG[kLspH];      // gaussCoefG_ as float (the reference casts at every use)
__constant__ float c_gaussL[3 * kBandW];  // gaussCoefL_

void upload_lbd_tables() {
  // BinaryDescriptor ctor (:219-261): the integer divisions are the reference's
  double gL[3 * kBandW], gG[kLspH];
  double u = (kBandW * 3 - 1) / 2, sigma = (kBandW * 2 + 1) / 2, inv = -1 / (2 * sigma * sigma);
  for (int i = 0; i < 3 * kBandW; ++i) gL[i] = exp((i - u) * (i - u) * inv);
  u = (kBands * kBandW - 1) / 2;
  sigma = u;
  inv = -1 / (2 * sigma * sigma);
  for (int i = 0; i < kLspH; ++i) gG[i] = exp((i - u) * (i - u) * inv);
  float fL[3 * kBandW], fG[kLspH];
  for (int i = 0; i < 3 * kBandW; ++i) fL[i] = (float)gL[i];
  for (int i = 0; i < kLspH; ++i) fG[i] = (float)gG[i];
  cudaMemcpyToSymbol(c_gaussL, fL, sizeof(fL));
  cudaMemcpyToSymbol(c_gaussG, fG, sizeof(fG));
}

__device__ __forceinline__ int reflect101_1(int i, int n) {  // one reflection is enough for a 1-px halo, n >= 2
  if (n == 1) return 0;
  return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i);
}

// cv::Sobel(CV_16SC1, 3x3, BORDER_REFLECT_101) of the blurred image, dx and dy interleaved (:374-399)
__global__ void __launch_bounds__(256)
    sobel_kernel(const uint8_t* __restrict__ img, int pitch, int64_t fs, int w, int h, short2* __restrict__ gxy) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  if (x >= w) return;
  const uint8_t* base = img + (size_t)b * fs;
  const uint8_t* r0 = base + (size_t)reflect101_1(y - 1, h) * pitch;
  const uint8_t* r1 = base + (size_t)y * pitch;
  const uint8_t* r2 = base + (size_t)reflect101_1(y + 1, h) * pitch;
  const int xm = reflect101_1(x - 1, w), xp = reflect101_1(x + 1, w);
  const int dx = ((int)r0[xp] - (int)r0[xm]) + 2 * ((int)r1[xp] - (int)r1[xm]) + ((int)r2[xp] - (int)r2[xm]);
  const int dy = ((int)r2[xm] + 2 * (int)r2[x] + (int)r2[xp]) - ((int)r0[xm] + 2 * (int)r0[x] + (int)r0[xp]);
  gxy[((size_t)b * h + y) * w + x] = make_short2((short)dx, (short)dy);
}

// computeLBD for one line (:1074-1330) + binaryConversion (:402-413, :655-668).  Block = one line:
// thread = one row of the line support region for the fp32 row sums (sequential along the line, as in the
// reference, so the sums round identically), then thread = band for the 8 band statistics (rows in hID
// order), then thread 0 for the two normalisations.
__global__ void __launch_bounds__(64)
    lbd_kernel(const short2* __restrict__ gxy, int w, int h, const psl_keyline* __restrict__ kl,
               const int32_t* __restrict__ n_kl, int cap, uint8_t* __restrict__ ldesc, float* __restrict__ lbd72) {
  const int li = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  if (li >= n_kl[b]) return;
  const psl_keyline k = kl[(size_t)b * cap + li];
  const short2* g = gxy + (size_t)b * w * h;
  __shared__ float rows[kLspH][4];
  __shared__ float des[8 * kBands];
  const float dL0 = (float)cos((double)k.angle), dL1 = (float)sin((double)k.angle);
  const float dO0 = -dL1, dO1 = dL0;
  if (t < kLspH) {
    const short lengthOfLSP = (short)k.num_pixels;
    const short halfHeight = (kLspH - 1) / 2, halfWidth = (short)((lengthOfLSP - 1) / 2);
    const short imageWidth = (short)(w - 1), imageHeight = (short)(h - 1);
    const float midX = (float)(0.5 * (double)(k.s_oct_x + k.e_oct_x)), midY = (float)(0.5 * (double)(k.s_oct_y + k.e_oct_y));
    float sCorX0 = -dL0 * (float)halfWidth + dL1 * (float)halfHeight + midX;
    float sCorY0 = -dL1 * (float)halfWidth - dL0 * (float)halfHeight + midY;
    for (int r = 0; r < t; ++r) {  // the reference steps the row origin incrementally in fp32
      sCorX0 -= dL1;
      sCorY0 += dL0;
    }
    float sCorX = sCorX0, sCorY = sCorY0;
    float pgdL = 0, ngdL = 0, pgdO = 0, ngdO = 0;
    for (short wID = 0; wID < lengthOfLSP; ++wID) {
      short q = (short)round((double)sCorX);
      const short xCor = (q < 0) ? (short)0 : (q > imageWidth) ? imageWidth : q;
      q = (short)round((double)sCorY);
      const short yCor = (q < 0) ? (short)0 : (q > imageHeight) ? imageHeight : q;
      const short2 gg = __ldg(g + (int)yCor * w + (int)xCor);
      const float gDL = (float)gg.x * dL0 + (float)gg.y * dL1, gDO = (float)gg.x * dO0 + (float)gg.y * dO1;
      if (gDL > 0) pgdL += gDL; else ngdL -= gDL;
      if (gDO > 0) pgdO += gDO; else ngdO -= gDO;
      sCorX += dL0;
      sCorY += dL1;
    }
    const float coef = c_gaussG[t];
    rows[t][0] = coef * pgdL;
    rows[t][1] = coef * ngdL;
    rows[t][2] = coef * pgdO;
    rows[t][3] = coef * ngdO;
  }
  __syncthreads();
  if (t < kBands) {
    float pL = 0, nL = 0, pL2 = 0, nL2 = 0, pO = 0, nO = 0, pO2 = 0, nO2 = 0;
    const int h0 = max(0, kBandW * (t - 1)), h1 = min(kLspH, kBandW * (t + 2));
    for (int hID = h0; hID < h1; ++hID) {
      const int band = hID / kBandW;
      // a row adds to its own band with gaussL[r+7], to band-1 with gaussL[r+14], to band+1 with gaussL[r]
      const float c = c_gaussL[hID % kBandW + (band == t ? kBandW : band == t + 1 ? 2 * kBandW : 0)];
      const float pgdL = rows[hID][0], ngdL = rows[hID][1], pgdO = rows[hID][2], ngdO = rows[hID][3];
      const float pgdL2 = pgdL * pgdL, ngdL2 = ngdL * ngdL, pgdO2 = pgdO * pgdO, ngdO2 = ngdO * ngdO;
      pL += c * pgdL; nL += c * ngdL;
      pL2 += c * c * pgdL2; nL2 += c * c * ngdL2;
      pO += c * pgdO; nO += c * ngdO;
      pO2 += c * c * pgdO2; nO2 += c * c * ngdO2;
    }
    const float invN = (t == 0 || t == kBands - 1) ? (float)(1.0 / (kBandW * 2.0)) : (float)(1.0 / (kBandW * 3.0));
    float* d = des + t * 8;
    float q = pL * invN; d[0] = q; d[4] = (float)sqrt((double)(pL2 * invN - q * q));
    q = nL * invN; d[1] = q; d[5] = (float)sqrt((double)(nL2 * invN - q * q));
    q = pO * invN; d[2] = q; d[6] = (float)sqrt((double)(pO2 * invN - q * q));
    q = nO * invN; d[3] = q; d[7] = (float)sqrt((double)(nO2 * invN - q * q));
  }
  __syncthreads();
  if (t == 0) {
    float tempM = 0, tempS = 0;
    for (int bd = 0; bd < kBands; ++bd) {
      const float* d = des + 8 * bd;
      tempM += d[0] * d[0]; tempM += d[1] * d[1]; tempM += d[2] * d[2]; tempM += d[3] * d[3];
      tempS += d[4] * d[4]; tempS += d[5] * d[5]; tempS += d[6] * d[6]; tempS += d[7] * d[7];
    }
    tempM = (float)(1 / sqrt((double)tempM));
    tempS = (float)(1 / sqrt((double)tempS));
    for (int bd = 0; bd < kBands; ++bd) {
      float* d = des + 8 * bd;
      d[0] *= tempM; d[1] *= tempM; d[2] *= tempM; d[3] *= tempM;
      d[4] *= tempS; d[5] *= tempS; d[6] *= tempS; d[7] *= tempS;
    }
    for (int i = 0; i < 72; ++i)
      if ((double)des[i] > 0.4) des[i] = (float)0.4;
    float temp = 0;
    for (int i = 0; i < 72; ++i) temp += des[i] * des[i];
    temp = (float)(1 / sqrt((double)temp));
    for (int i = 0; i < 72; ++i) des[i] = des[i] * temp;
  }
  __syncthreads();
  const size_t row = (size_t)b * cap + li;
  if (lbd72) {
    lbd72[row * 72 + t] = des[t];
    if (t < 8) lbd72[row * 72 + 64 + t] = des[64 + t];
  }
  if (t < 32) {
    // the 32 band pairs (i, j), i < j, of `combinations` (:76-109) in its order
    constexpr unsigned char ci[32] = {0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 6, 6, 7};
    constexpr unsigned char cj[32] = {1, 2, 3, 4, 5, 6, 2, 3, 4, 5, 6, 3, 4, 5, 6, 7, 8, 4, 5, 6, 7, 8, 5, 6, 7, 8, 6, 7, 8, 7, 8, 8};
    const float* f1 = des + 8 * ci[t];
    const float* f2 = des + 8 * cj[t];
    unsigned r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (f1[i] > f2[i]) r += 1u << i;
    ldesc[row * 32 + t] = (uint8_t)r;
  }
}

void launch_lbd(const LineBuffers& L, ImgBatch in, int nb, int nfeatures, const psl_keyline* kl, const int32_t* n_kl,
                int cap, uint8_t* ldesc, float* lbd72, cudaStream_t st) {
  ImgBatchMut bl{L.blur, L.pitch, (int64_t)L.pitch * L.h, L.w, L.h};
  launch_blur7(in, bl, 0, 14, 62, 104, nb, st);  // GaussianBlur(5x5, sigma 1), computeGaussianPyramid :351-371
  dim3 g1((L.w + 255) / 256, L.h, nb);
  sobel_kernel<<<g1, 256, 0, st>>>(L.blur, L.pitch, (int64_t)L.pitch * L.h, L.w, L.h, L.gxy);
  dim3 g2(nfeatures < cap ? nfeatures : cap, nb);
  lbd_kernel<<<g2, 64, 0, st>>>(L.gxy, L.w, L.h, kl, n_kl, cap, ldesc, lbd72);
}

}  // namespace psl
