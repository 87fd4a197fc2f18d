// TEST INFRASTRUCTURE — CPU oracle for the Frame bookkeeping between extraction and matching on the
// RGB-D path: depth scaling (src/Tracking.cc:234-235), Frame::ComputeStereoFromRGBD and
// UnprojectStereo (src/Frame.cc:1342-1381), Frame::UpdatePoseMatrices (mRwc, mOw) and the per-point
// prologue of ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) (src/ORBmatcher.cc:1339-1393).
// cv::Mat products of CV_32F matrices accumulate in double and round once (OpenCV gemm).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "psl_oracle.h"

namespace {
// float( sum_k R[i][k]*x[k] (double) + t[i] ), R row-major 3x3 taken from a 3x4 [R|t] block
inline void affine(const float* R, int rs, const float* x, const float* t, float* out) {
  for (int i = 0; i < 3; ++i) {
    double s = (double)R[i * rs] * x[0] + (double)R[i * rs + 1] * x[1] + (double)R[i * rs + 2] * x[2];
    out[i] = (float)(s + (t ? (double)t[i] : 0.0));
  }
}
}  // namespace

extern "C" {

// cv::undistortPoints(src, dst, K, D, noArray(), K) of one point (OpenCV 4.x cvUndistortPointsInternal, default
// criteria = 5 iterations): double arithmetic on the float inputs; k4..k6, thin prism and tilt terms are zero here.
static void undistort_point(float u_in, float v_in, const psl_distortion& c, float* out) {
  const double fx = c.fx, fy = c.fy, cx = c.cx, cy = c.cy, k1 = c.k1, k2 = c.k2, p1 = c.p1, p2 = c.p2, k3 = c.k3;
  const double ifx = 1. / fx, ify = 1. / fy;
  const double u = u_in, v = v_in;
  double x = (u - cx) * ifx, y = (v - cy) * ify;
  const double x0 = x, y0 = y;
  for (int j = 0; j < 5; ++j) {
    const double r2 = x * x + y * y;
    const double icdist = 1. / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
    if (icdist < 0) {
      x = (u - cx) * ifx;
      y = (v - cy) * ify;
      break;
    }
    const double deltaX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
    const double deltaY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
    x = (x0 - deltaX) * icdist;
    y = (y0 - deltaY) * icdist;
  }
  const double xx = fx * x + 0. * y + cx, yy = 0. * x + fy * y + cy, ww = 1. / (0. * x + 0. * y + 1.);
  out[0] = (float)(xx * ww);
  out[1] = (float)(yy * ww);
}

// Frame::UndistortKeyPoints, Frame.cc:1062-1092
void orc_undistort_keypoints(const psl_keypoint* kps, int n, const psl_distortion* cam, psl_keypoint* kps_un) {
  for (int i = 0; i < n; ++i) {
    psl_keypoint k = kps[i];
    if (cam->k1 != 0.f) {  // :1064
      float p[2];
      undistort_point(k.x, k.y, *cam, p);
      k.x = p[0];
      k.y = p[1];
    }
    kps_un[i] = k;
  }
}

// Frame::ComputeImageBounds, Frame.cc:1135-1163; bounds = mnMinX, mnMinY, mnMaxX, mnMaxY
void orc_image_bounds(int cols, int rows, const psl_distortion* cam, float* bounds) {
  if (cam->k1 == 0.f) {
    bounds[0] = 0.f; bounds[1] = 0.f; bounds[2] = (float)cols; bounds[3] = (float)rows;
    return;
  }
  const float in[4][2] = {{0.f, 0.f}, {(float)cols, 0.f}, {0.f, (float)rows}, {(float)cols, (float)rows}};
  float m[4][2];
  for (int i = 0; i < 4; ++i) undistort_point(in[i][0], in[i][1], *cam, m[i]);
  bounds[0] = std::min(m[0][0], m[2][0]);
  bounds[2] = std::max(m[1][0], m[3][0]);
  bounds[1] = std::min(m[0][1], m[1][1]);
  bounds[3] = std::max(m[2][1], m[3][1]);
}

// mvuRight / mvDepth for every keypoint (zero distortion: mvKeysUn == mvKeys)
void orc_stereo_from_rgbd(const psl_keypoint* kps, int n, const uint16_t* depth, int w, int h, int stride_px,
                          float depth_factor, float bf, float* u_right, float* z) {
  (void)h;
  (void)w;
  for (int i = 0; i < n; ++i) {
    u_right[i] = -1.f;
    z[i] = -1.f;
    const float d = (float)depth[(size_t)(int)kps[i].y * stride_px + (int)kps[i].x] * depth_factor;
    if (d > 0) {
      z[i] = d;
      u_right[i] = kps[i].x - bf / d;
    }
  }
}

// Queries for SearchByProjection(Current, Last): Tcw_* are row-major 3x4 [R|t] float.
// cam = fx, fy, cx, cy, bf.  valid_in (may be NULL) marks keypoints whose MapPoint exists and is no outlier.
void orc_queries_from_last_frame(const psl_keypoint* kps_last, const float* z_last, const uint8_t* valid_in,
                                 const uint8_t* claims_in, int n, const float* Tcw_last, const float* Tcw_cur,
                                 const float* cam, const float* scale_factors, float th, int mono, float min_x,
                                 float min_y, float max_x, float max_y, psl_proj_query* q) {
  const float fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3], bf = cam[4];
  const float invfx = 1.0f / fx, invfy = 1.0f / fy, mb = bf / fx;
  // Frame::UpdatePoseMatrices of the last frame: mRwc = Rcw^T, mOw = -Rcw^T * tcw
  float Rwc[9], Ow[3], nRwc[9];
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k) { Rwc[i * 3 + k] = Tcw_last[k * 4 + i]; nRwc[i * 3 + k] = -Rwc[i * 3 + k]; }
  const float tl[3] = {Tcw_last[3], Tcw_last[7], Tcw_last[11]};
  affine(nRwc, 3, tl, nullptr, Ow);
  // twc of the current frame and tlc (:1342-1352)
  float nRcT[9], twc[3], tlc[3];
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k) nRcT[i * 3 + k] = -Tcw_cur[k * 4 + i];
  const float tc[3] = {Tcw_cur[3], Tcw_cur[7], Tcw_cur[11]};
  affine(nRcT, 3, tc, nullptr, twc);
  affine(Tcw_last, 4, twc, tl, tlc);
  const bool fwd = tlc[2] > mb && !mono, bwd = -tlc[2] > mb && !mono;
  for (int i = 0; i < n; ++i) {
    std::memset(&q[i], 0, sizeof(q[i]));
    if (valid_in && !valid_in[i]) continue;
    const float z = z_last[i];
    if (!(z > 0)) continue;
    const float u0 = kps_last[i].x, v0 = kps_last[i].y;
    const float xc[3] = {(u0 - cx) * z * invfx, (v0 - cy) * z * invfy, z};  // UnprojectStereo
    float pw[3], pc[3];
    affine(Rwc, 3, xc, Ow, pw);
    affine(Tcw_cur, 4, pw, tc, pc);
    const float invz = (float)(1.0 / pc[2]);
    if (invz < 0) continue;
    const float u = fx * pc[0] * invz + cx, v = fy * pc[1] * invz + cy;
    if (u < min_x || u > max_x) continue;
    if (v < min_y || v > max_y) continue;
    const int o = kps_last[i].octave;
    q[i].u = u;
    q[i].v = v;
    q[i].radius = th * scale_factors[o];
    if (fwd) { q[i].min_level = o; q[i].max_level = -1; }
    else if (bwd) { q[i].min_level = 0; q[i].max_level = o; }
    else { q[i].min_level = o - 1; q[i].max_level = o + 1; }
    q[i].u_right = u - bf * invz;
    q[i].angle = kps_last[i].angle;
    q[i].flags = PSL_Q_VALID | ((!claims_in || claims_in[i]) ? PSL_Q_CLAIMS : 0u);
  }
}

}  // extern "C"
