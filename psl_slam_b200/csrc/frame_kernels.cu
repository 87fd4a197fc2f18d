// Frame bookkeeping between extraction and matching on the RGB-D path, kept on the device so the
// batched pipeline never leaves HBM: depth scaling (Tracking.cc:234-235), ComputeStereoFromRGBD and
// UnprojectStereo (Frame.cc:1342-1381), UpdatePoseMatrices, and the per-point prologue of
// SearchByProjection(Current, Last, th, bMono) (ORBmatcher.cc:1339-1393).
// Arithmetic: fp32 with explicit round-to-nearest ops; 3x3 products accumulate in fp64 and round once,
// as cv::Mat (CV_32F) products do.
#include "frame_kernels.cuh"

namespace psl {

__global__ void stereo_kernel(const psl_keypoint* __restrict__ kps, const int32_t* __restrict__ n, int cap,
                              const uint16_t* __restrict__ depth, int stride_px, int64_t frame_stride_px,
                              float depth_factor, float bf, float* __restrict__ u_right, float* __restrict__ z) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n[b]) return;
  const psl_keypoint kp = kps[(size_t)b * cap + i];
  const float d = __fmul_rn((float)depth[(size_t)b * frame_stride_px + (size_t)(int)kp.y * stride_px + (int)kp.x],
                            depth_factor);
  float ur = -1.f, zz = -1.f;
  if (d > 0.f) {
    zz = d;
    ur = __fsub_rn(kp.x, __fdiv_rn(bf, d));
  }
  u_right[(size_t)b * cap + i] = ur;
  z[(size_t)b * cap + i] = zz;
}

void launch_stereo(const psl_keypoint* kps, const int32_t* n, int cap, const uint16_t* depth, int stride_px,
                   int64_t frame_stride_px, float depth_factor, float bf, float* u_right, float* z, int B,
                   cudaStream_t st) {
  dim3 grid((cap + 255) / 256, B);
  stereo_kernel<<<grid, 256, 0, st>>>(kps, n, cap, depth, stride_px, frame_stride_px, depth_factor, bf, u_right, z);
}

// float( sum_k R[i][k]*x[k] in double + t[i] )
__device__ __forceinline__ void affine(const float* R, int rs, const float* x, const float* t, float* out) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double s = __dadd_rn(__dadd_rn(__dmul_rn((double)R[i * rs], (double)x[0]),
                                         __dmul_rn((double)R[i * rs + 1], (double)x[1])),
                               __dmul_rn((double)R[i * rs + 2], (double)x[2]));
    out[i] = (float)__dadd_rn(s, t ? (double)t[i] : 0.0);
  }
}

// Queries of frame gb = first + blockIdx.y from the keypoints of frame gb-1 (same [B][cap] arrays).
__global__ void query_build_kernel(const psl_keypoint* __restrict__ kps, const float* __restrict__ z,
                                   const int32_t* __restrict__ n, int cap, const float* __restrict__ Tcw, int first,
                                   psl_camera cam, QueryBuildParams prm, psl_proj_query* __restrict__ q,
                                   int32_t* __restrict__ nq) {
  const int b = blockIdx.y, gb = first + b, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (gb == 0) {
    if (i == 0) nq[b] = 0;
    return;
  }
  const int nl = n[gb - 1];
  if (i == 0) nq[b] = nl;
  if (i >= nl) return;
  const float* Tl = Tcw + (size_t)(gb - 1) * 12;
  const float* Tc = Tcw + (size_t)gb * 12;
  const float fx = cam.fx, fy = cam.fy, cx = cam.cx, cy = cam.cy, bf = cam.bf;
  const float invfx = __fdiv_rn(1.0f, fx), invfy = __fdiv_rn(1.0f, fy), mb = __fdiv_rn(bf, fx);
  float Rwc[9], nRwc[9], Ow[3], nRcT[9], twc[3], tlc[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      Rwc[r * 3 + k] = Tl[k * 4 + r];
      nRwc[r * 3 + k] = -Tl[k * 4 + r];
      nRcT[r * 3 + k] = -Tc[k * 4 + r];
    }
  const float tl[3] = {Tl[3], Tl[7], Tl[11]}, tc[3] = {Tc[3], Tc[7], Tc[11]};
  affine(nRwc, 3, tl, nullptr, Ow);   // mOw = -Rcw^T tcw (Frame::UpdatePoseMatrices)
  affine(nRcT, 3, tc, nullptr, twc);  // :1345
  affine(Tl, 4, twc, tl, tlc);        // :1350
  const bool fwd = tlc[2] > mb && !prm.mono, bwd = -tlc[2] > mb && !prm.mono;

  psl_proj_query Q;
  Q.u = Q.v = Q.radius = Q.u_right = Q.angle = 0.f;
  Q.min_level = Q.max_level = 0;
  Q.flags = 0;
  const psl_keypoint kp = kps[(size_t)(gb - 1) * cap + i];
  const float zz = z[(size_t)(gb - 1) * cap + i];
  if (zz > 0.f) {
    const float xc[3] = {__fmul_rn(__fmul_rn(__fsub_rn(kp.x, cx), zz), invfx),
                         __fmul_rn(__fmul_rn(__fsub_rn(kp.y, cy), zz), invfy), zz};
    float pw[3], pc[3];
    affine(Rwc, 3, xc, Ow, pw);  // UnprojectStereo
    affine(Tc, 4, pw, tc, pc);   // :1363
    const float invz = (float)__ddiv_rn(1.0, (double)pc[2]);
    if (!(invz < 0.f)) {
      const float u = __fadd_rn(__fmul_rn(__fmul_rn(fx, pc[0]), invz), cx);
      const float v = __fadd_rn(__fmul_rn(__fmul_rn(fy, pc[1]), invz), cy);
      if (!(u < prm.min_x || u > prm.max_x || v < prm.min_y || v > prm.max_y)) {
        const int o = kp.octave;
        Q.u = u;
        Q.v = v;
        Q.radius = __fmul_rn(prm.th, prm.scale[o]);
        if (fwd) { Q.min_level = o; Q.max_level = -1; }
        else if (bwd) { Q.min_level = 0; Q.max_level = o; }
        else { Q.min_level = o - 1; Q.max_level = o + 1; }
        Q.u_right = __fsub_rn(u, __fmul_rn(bf, invz));
        Q.angle = kp.angle;
        Q.flags = PSL_Q_VALID | PSL_Q_CLAIMS;
      }
    }
  }
  q[(size_t)b * cap + i] = Q;
}

void launch_query_build(const psl_keypoint* kps, const float* z, const int32_t* n, int cap, const float* Tcw, int first,
                        const psl_camera& cam, const QueryBuildParams& prm, psl_proj_query* q, int32_t* nq, int B,
                        cudaStream_t st) {
  dim3 grid((cap + 127) / 128, B);
  query_build_kernel<<<grid, 128, 0, st>>>(kps, z, n, cap, Tcw, first, cam, prm, q, nq);
}

// The correspondences PoseOptimization gets after SearchByProjection(Current, Last) in TrackWithMotionModel
// (Tracking.cc:1193-1214): keypoint i of frame b matched to keypoint j = assign[b][i] of frame b-1 observes the
// MapPoint that frame b-1 created at j, i.e. Last.UnprojectStereo(j) (Frame.cc:1367-1381) — the world point of the
// synthetic sequence's "map".  One psl_pose_point per keypoint (flags = 0 without a match), plus the 4x4 prior pose.
__global__ void pose_points_kernel(const psl_keypoint* __restrict__ kps, const float* __restrict__ u_right,
                                   const float* __restrict__ z, const int32_t* __restrict__ assign,
                                   const int32_t* __restrict__ n, int cap, const float* __restrict__ Tcw, psl_camera cam,
                                   const float* __restrict__ inv_sigma2, psl_pose_point* __restrict__ pts,
                                   float* __restrict__ T44) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 16) {
    const int r = i >> 2, c = i & 3;
    T44[(size_t)b * 16 + i] = r < 3 ? Tcw[(size_t)b * 12 + i] : (c == 3 ? 1.f : 0.f);
  }
  if (i >= n[b]) return;
  const psl_keypoint kp = kps[(size_t)b * cap + i];
  psl_pose_point P;
  P.u = kp.x; P.v = kp.y;
  P.u_right = u_right[(size_t)b * cap + i];
  P.inv_sigma2 = inv_sigma2[kp.octave];
  P.xw = P.yw = P.zw = 0.f;
  P.flags = 0;
  const int j = b > 0 ? assign[(size_t)b * cap + i] : -1;
  if (j >= 0) {
    const float* Tl = Tcw + (size_t)(b - 1) * 12;
    float Rwc[9], nRwc[9], Ow[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        Rwc[r * 3 + k] = Tl[k * 4 + r];
        nRwc[r * 3 + k] = -Tl[k * 4 + r];
      }
    const float tl[3] = {Tl[3], Tl[7], Tl[11]};
    affine(nRwc, 3, tl, nullptr, Ow);   // mOw = -Rcw^T tcw
    const psl_keypoint kl = kps[(size_t)(b - 1) * cap + j];
    const float zz = z[(size_t)(b - 1) * cap + j];
    if (zz > 0.f) {
      const float invfx = __fdiv_rn(1.0f, cam.fx), invfy = __fdiv_rn(1.0f, cam.fy);
      const float xc[3] = {__fmul_rn(__fmul_rn(__fsub_rn(kl.x, cam.cx), zz), invfx),
                           __fmul_rn(__fmul_rn(__fsub_rn(kl.y, cam.cy), zz), invfy), zz};
      float pw[3];
      affine(Rwc, 3, xc, Ow, pw);  // UnprojectStereo
      P.xw = pw[0]; P.yw = pw[1]; P.zw = pw[2];
      P.flags = 1;
    }
  }
  pts[(size_t)b * cap + i] = P;
}

void launch_pose_points(const psl_keypoint* kps, const float* u_right, const float* z, const int32_t* assign,
                        const int32_t* n, int cap, const float* Tcw, const psl_camera& cam, const float* inv_sigma2,
                        psl_pose_point* pts, float* T44, int B, cudaStream_t st) {
  dim3 grid((cap + 127) / 128, B);
  pose_points_kernel<<<grid, 128, 0, st>>>(kps, u_right, z, assign, n, cap, Tcw, cam, inv_sigma2, pts, T44);
}

// ---------------------------------------------------------------------------------------------
// K0: input conversion of Tracking::GrabImageRGBD (Tracking.cc:219-235).  Pure streaming: a thread
// converts 4 adjacent pixels (12 or 16 colour bytes in, one gray word out).
// ---------------------------------------------------------------------------------------------
template <int CH>
__global__ void __launch_bounds__(256)
    color_to_gray_kernel(const uint8_t* __restrict__ color, int r_first, int color_stride, int64_t color_fs,
                         uint8_t* __restrict__ gray, int gray_stride, int64_t gray_fs, int w, int h) {
  // a thread owns 4 adjacent pixels; the groups of all rows are numbered through, so that no thread of a CTA idles on a
  // row that is not a multiple of 1024 pixels wide
  const int gpr = (w + 3) >> 2, idx = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  const int y = idx / gpr, x = (idx - y * gpr) * 4;
  if (y >= h) return;
  const uint8_t* src = color + (size_t)b * color_fs + (size_t)y * color_stride + (size_t)x * CH;
  uint8_t* dst = gray + (size_t)b * gray_fs + (size_t)y * gray_stride + x;
  const int n = min(4, w - x);
  uint8_t px[4 * CH];
  const bool vec = n == 4 && ((uintptr_t)src & 3) == 0;
  if (vec) {
#pragma unroll
    for (int k = 0; k < CH; ++k) reinterpret_cast<uint32_t*>(px)[k] = __ldg(reinterpret_cast<const uint32_t*>(src) + k);
  } else {
    for (int k = 0; k < n * CH; ++k) px[k] = __ldg(src + k);
  }
  uint32_t out = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k < n) {
      const uint32_t c0 = px[k * CH], c1 = px[k * CH + 1], c2 = px[k * CH + 2];
      const uint32_t R = r_first ? c0 : c2, B = r_first ? c2 : c0;
      out |= ((R * 9798u + c1 * 19235u + B * 3735u + 16384u) >> 15) << (8 * k);
    }
  }
  if (n == 4 && ((uintptr_t)dst & 3) == 0) *reinterpret_cast<uint32_t*>(dst) = out;
  else
    for (int k = 0; k < n; ++k) dst[k] = (uint8_t)(out >> (8 * k));
}

// The same for 3-channel images whose rows are whole 16-pixel groups on 16-byte boundaries: a thread owns 16 pixels,
// three 16-byte loads in flight, one 16-byte store.
__global__ void __launch_bounds__(256)
    color3_to_gray16_kernel(const uint8_t* __restrict__ color, int r_first, int color_stride, int64_t color_fs,
                            uint8_t* __restrict__ gray, int gray_stride, int64_t gray_fs, int w, int h) {
  const int gpr = w >> 4, idx = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  const int y = idx / gpr, x = (idx - y * gpr) * 16;
  if (y >= h) return;
  const uint4* src = reinterpret_cast<const uint4*>(color + (size_t)b * color_fs + (size_t)y * color_stride + (size_t)x * 3);
  const uint4 v0 = __ldg(src), v1 = __ldg(src + 1), v2 = __ldg(src + 2);
  const uint32_t wds[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
  uint32_t out[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {   // 4 pixels = 12 bytes = words 3g .. 3g + 2
    uint32_t o = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int byte0 = 12 * g + 3 * k;
      auto byte_at = [&](int i) -> uint32_t { return (wds[i >> 2] >> (8 * (i & 3))) & 255u; };
      const uint32_t c0 = byte_at(byte0), c1 = byte_at(byte0 + 1), c2 = byte_at(byte0 + 2);
      const uint32_t R = r_first ? c0 : c2, B = r_first ? c2 : c0;
      o |= ((R * 9798u + c1 * 19235u + B * 3735u + 16384u) >> 15) << (8 * k);
    }
    out[g] = o;
  }
  *reinterpret_cast<uint4*>(gray + (size_t)b * gray_fs + (size_t)y * gray_stride + x) = make_uint4(out[0], out[1], out[2], out[3]);
}

void launch_color_to_gray(const uint8_t* color, int channels, int rgb_order, int color_stride, int64_t color_fs,
                          uint8_t* gray, int gray_stride, int64_t gray_fs, int B, int w, int h, cudaStream_t st) {
  if (channels == 3 && (w & 15) == 0 && ((uintptr_t)color & 15) == 0 && (color_stride & 15) == 0 && (color_fs & 15) == 0 &&
      ((uintptr_t)gray & 15) == 0 && (gray_stride & 15) == 0 && (gray_fs & 15) == 0) {
    dim3 grid16(((w >> 4) * h + 255) / 256, B);
    color3_to_gray16_kernel<<<grid16, 256, 0, st>>>(color, rgb_order, color_stride, color_fs, gray, gray_stride, gray_fs, w, h);
    return;
  }
  dim3 grid(((w + 3) / 4 * h + 255) / 256, B);
  if (channels == 3)
    color_to_gray_kernel<3><<<grid, 256, 0, st>>>(color, rgb_order, color_stride, color_fs, gray, gray_stride, gray_fs, w, h);
  else
    color_to_gray_kernel<4><<<grid, 256, 0, st>>>(color, rgb_order, color_stride, color_fs, gray, gray_stride, gray_fs, w, h);
}

__global__ void __launch_bounds__(256)
    depth_to_float_kernel(const uint16_t* __restrict__ in, int stride_px, int64_t fs_px, float factor,
                          float* __restrict__ out, int w, int h) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
  if (x >= w) return;
  // cv::Mat::convertTo(CV_32F, alpha): saturate_cast<float>(src * alpha) evaluated in fp32 for 16U sources
  out[((size_t)b * h + y) * w + x] = __fmul_rn((float)__ldg(in + (size_t)b * fs_px + (size_t)y * stride_px + x), factor);
}

void launch_depth_to_float(const uint16_t* in, int stride_px, int64_t fs_px, float factor, float* out, int B, int w,
                           int h, cudaStream_t st) {
  dim3 grid((w + 255) / 256, h, B);
  depth_to_float_kernel<<<grid, 256, 0, st>>>(in, stride_px, fs_px, factor, out, w, h);
}

// cv::undistortPoints(src, dst, K, D, noArray(), K) for one point, OpenCV 4.x cvUndistortPointsInternal with its
// default criteria (5 iterations, no epsilon test): everything in double from the float inputs, no contraction.
// The rational terms k4..k6, the thin-prism terms and the tilt are zero for the 4/5-coefficient models PSL-SLAM
// reads (Tracking.cc:66-77); their additions of +0.0 are kept out, which leaves every value unchanged.
__device__ __forceinline__ float2 undistort_point(float u_in, float v_in, const psl_distortion& c) {
  const double fx = (double)c.fx, fy = (double)c.fy, cx = (double)c.cx, cy = (double)c.cy;
  const double k1 = (double)c.k1, k2 = (double)c.k2, p1 = (double)c.p1, p2 = (double)c.p2, k3 = (double)c.k3;
  const double ifx = __ddiv_rn(1.0, fx), ify = __ddiv_rn(1.0, fy);
  const double u = (double)u_in, v = (double)v_in;
  double x = __dmul_rn(__dsub_rn(u, cx), ifx), y = __dmul_rn(__dsub_rn(v, cy), ify);
  const double x0 = x, y0 = y;
  for (int j = 0; j < 5; ++j) {
    const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
    const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k3, r2), k2), r2), k1), r2));
    const double icdist = __ddiv_rn(1.0, den);
    if (icdist < 0) {
      x = x0;
      y = y0;
      break;
    }
    // deltaX = 2*p1*x*y + p2*(r2 + 2*x*x);  deltaY = p1*(r2 + 2*y*y) + 2*p2*x*y   (left to right)
    const double dX = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, p1), x), y),
                                __dmul_rn(p2, __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x))));
    const double dY = __dadd_rn(__dmul_rn(p1, __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))),
                                __dmul_rn(__dmul_rn(__dmul_rn(2.0, p2), x), y));
    x = __dmul_rn(__dsub_rn(x0, dX), icdist);
    y = __dmul_rn(__dsub_rn(y0, dY), icdist);
  }
  // RR = P * I: xx = fx*x + 0*y + cx, yy = 0*x + fy*y + cy, ww = 1 / (0*x + 0*y + 1)
  const double xx = __dadd_rn(__dadd_rn(__dmul_rn(fx, x), __dmul_rn(0.0, y)), cx);
  const double yy = __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(fy, y)), cy);
  const double ww = __ddiv_rn(1.0, __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(0.0, y)), 1.0));
  return make_float2((float)__dmul_rn(xx, ww), (float)__dmul_rn(yy, ww));
}

__global__ void __launch_bounds__(128)
    undistort_kernel(const psl_keypoint* __restrict__ kps, const int32_t* __restrict__ n, int cap, psl_distortion cam,
                     psl_keypoint* __restrict__ kps_un) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n[b]) return;
  psl_keypoint k = kps[(size_t)b * cap + i];
  if (cam.k1 != 0.f) {  // Frame.cc:1064
    const float2 p = undistort_point(k.x, k.y, cam);
    k.x = p.x;
    k.y = p.y;
  }
  kps_un[(size_t)b * cap + i] = k;
}

void launch_undistort(const psl_keypoint* kps, const int32_t* n, int cap, const psl_distortion& cam, psl_keypoint* kps_un,
                      int B, cudaStream_t st) {
  if (B <= 0 || cap <= 0) return;
  dim3 grid((cap + 127) / 128, B);
  undistort_kernel<<<grid, 128, 0, st>>>(kps, n, cap, cam, kps_un);
}

}  // namespace psl
