// Optimizer::PoseOptimization ("next" row N4): the pose-only Levenberg optimisation of src/Optimizer.cc:239-1023 over what
// it calls in the vendored g2o (EdgeSE3ProjectXYZOnlyPose / EdgeStereoSE3ProjectXYZOnlyPose, SE3Quat, RobustKernelHuber,
// BaseUnaryEdge / BaseBinaryEdge::constructQuadraticForm, OptimizationAlgorithmLevenberg::solve, SparseOptimizer::optimize)
// with the point edges and the structural-line edges EdgeLILSE3ProjectXYZ (add_inc/EdgeLIL.h:210-374; built at
// Optimizer.cc:619-693, classified at :976-1007).  One CTA per frame: the edges are spread over the
// threads, the 6x6 normal equations and the robust chi2 are block reductions (fixed order: per-thread partial sums in
// index order, shuffle tree, warps in order), thread 0 runs the Levenberg control flow, the 6x6 LDL^T solve and the SE3
// update.  fp64 throughout; the sums run in another order than g2o's edge loop, so the contract is a tolerance on the
// pose and identical outlier flags away from the chi2 thresholds, not bit-exactness.
#include <cfloat>
#include <math.h>

#include <algorithm>

#include "psl_ctx.cuh"

namespace psl {
namespace {

constexpr int kPoseThreads = 256;

struct Quat { double w, x, y, z; };
struct Pose { Quat q; double t[3]; };
struct Cam { double fx, fy, cx, cy, bf; };

__device__ void normalize_rotation(Quat& q) {
  if (q.w < 0) { q.w = -q.w; q.x = -q.x; q.y = -q.y; q.z = -q.z; }
  const double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  q.w /= n; q.x /= n; q.y /= n; q.z /= n;
}
__device__ Quat quat_from_matrix(const double m[3][3]) {
  Quat q;
  double t = m[0][0] + m[1][1] + m[2][2];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (m[2][1] - m[1][2]) * t; q.y = (m[0][2] - m[2][0]) * t; q.z = (m[1][0] - m[0][1]) * t;
  } else {
    int i = 0;
    if (m[1][1] > m[0][0]) i = 1;
    if (m[2][2] > m[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
    double v[3];
    v[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (m[k][j] - m[j][k]) * t;
    v[j] = (m[j][i] + m[i][j]) * t;
    v[k] = (m[k][i] + m[i][k]) * t;
    q.x = v[0]; q.y = v[1]; q.z = v[2];
  }
  return q;
}
__device__ __forceinline__ void quat_rotate(const Quat& q, const double v[3], double out[3]) {
  const double ux = 2 * (q.y * v[2] - q.z * v[1]), uy = 2 * (q.z * v[0] - q.x * v[2]), uz = 2 * (q.x * v[1] - q.y * v[0]);
  out[0] = v[0] + q.w * ux + (q.y * uz - q.z * uy);
  out[1] = v[1] + q.w * uy + (q.z * ux - q.x * uz);
  out[2] = v[2] + q.w * uz + (q.x * uy - q.y * ux);
}
__device__ Quat quat_mul(const Quat& a, const Quat& b) {
  return Quat{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
              a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
__device__ void quat_to_matrix(const Quat& q, double R[3][3]) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y,
               tyz = tz * q.y, tzz = tz * q.z;
  R[0][0] = 1 - (tyy + tzz); R[0][1] = txy - twz; R[0][2] = txz + twy;
  R[1][0] = txy + twz; R[1][1] = 1 - (txx + tzz); R[1][2] = tyz - twx;
  R[2][0] = txz - twy; R[2][1] = tyz + twx; R[2][2] = 1 - (txx + tyy);
}
__device__ Pose pose_exp_times(const double u[6], const Pose& est) {  // SE3Quat::exp(update) * estimate
  const double om[3] = {u[0], u[1], u[2]}, up[3] = {u[3], u[4], u[5]};
  const double theta = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
  const double O[3][3] = {{0, -om[2], om[1]}, {om[2], 0, -om[0]}, {-om[1], om[0], 0}};
  double O2[3][3], R[3][3], V[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) O2[i][j] = O[i][0] * O[0][j] + O[i][1] * O[1][j] + O[i][2] * O[2][j];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      const double I = i == j ? 1.0 : 0.0;
      if (theta < 0.00001) {
        R[i][j] = I + O[i][j] + O2[i][j];
        V[i][j] = R[i][j];
      } else {
        R[i][j] = I + sin(theta) / theta * O[i][j] + (1 - cos(theta)) / (theta * theta) * O2[i][j];
        V[i][j] = I + (1 - cos(theta)) / (theta * theta) * O[i][j] + (theta - sin(theta)) / pow(theta, 3.0) * O2[i][j];
      }
    }
  Pose e;
  e.q = quat_from_matrix(R);
  normalize_rotation(e.q);
  for (int i = 0; i < 3; ++i) e.t[i] = V[i][0] * up[0] + V[i][1] * up[1] + V[i][2] * up[2];
  Pose r = e;
  double rt[3];
  quat_rotate(e.q, est.t, rt);
  for (int i = 0; i < 3; ++i) r.t[i] += rt[i];
  r.q = quat_mul(e.q, est.q);
  normalize_rotation(r.q);
  return r;
}
__device__ __forceinline__ void huber(double e2, double delta, double& rho0, double& rho1) {
  const double dsqr = delta * delta;
  if (e2 <= dsqr) { rho0 = e2; rho1 = 1.; }
  else { const double sqrte = sqrt(e2); rho0 = 2 * sqrte * delta - dsqr; rho1 = delta / sqrte; }
}
__device__ bool solve6(const double* Hin, const double* b, double* x) {  // LDL^T; fails unless positive definite
  double L[6][6], D[6];
  for (int j = 0; j < 6; ++j) {
    double d = Hin[j * 6 + j];
    for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k] * D[k];
    if (!(d > 0)) return false;
    D[j] = d;
    L[j][j] = 1;
    for (int i = j + 1; i < 6; ++i) {
      double s = Hin[i * 6 + j];
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k] * D[k];
      L[i][j] = s / d;
    }
  }
  double y[6];
  for (int i = 0; i < 6; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= L[i][k] * y[k]; y[i] = s; }
  for (int i = 0; i < 6; ++i) y[i] /= D[i];
  for (int i = 5; i >= 0; --i) { double s = y[i]; for (int k = i + 1; k < 6; ++k) s -= L[k][i] * x[k]; x[i] = s; }
  return true;
}

// block sum of K doubles per thread into out[K] (thread 0's view is complete after the trailing __syncthreads)
template <int K>
__device__ void block_sum(double (&v)[K], double* s_part /*[8][K]*/, double* s_out /*[K]*/) {
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int d = 16; d; d >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], d);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
    for (int k = 0; k < K; ++k) s_part[warp * K + k] = v[k];
  __syncthreads();
  if (threadIdx.x < K) {
    double s = 0;
    for (int w = 0; w < kPoseThreads / 32; ++w) s += s_part[w * K + threadIdx.x];
    s_out[threadIdx.x] = s;
  }
  __syncthreads();
}

// the five points of a LIL vertex seen from the pose (segment 1 start / end, segment 2 start / end, cross point)
__device__ __forceinline__ void lil_map(const Pose& p, const psl_pose_lil& l, int k, double X[3]) {
  const double* src = k < 2 ? l.line1 + 3 * k : k < 4 ? l.line2 + 3 * (k - 2) : l.cross;
  const double xw[3] = {src[0], src[1], src[2]};
  quat_rotate(p.q, xw, X);
  X[0] += p.t[0]; X[1] += p.t[1]; X[2] += p.t[2];
}
// EdgeLILSE3ProjectXYZ::computeError (EdgeLIL.h:220-259)
__device__ void lil_error(const Pose& p, const psl_pose_lil& l, const Cam& cam, double* e) {
  double u[5], v[5];
  for (int k = 0; k < 5; ++k) {
    double X[3];
    lil_map(p, l, k, X);
    u[k] = X[0] / X[2] * cam.fx + cam.cx;
    v[k] = X[1] / X[2] * cam.fy + cam.cy;
  }
  e[0] = u[0] * l.obs1[0] + v[0] * l.obs1[1] + 1.0 * l.obs1[2];
  e[1] = u[1] * l.obs1[0] + v[1] * l.obs1[1] + 1.0 * l.obs1[2];
  e[2] = u[2] * l.obs2[0] + v[2] * l.obs2[1] + 1.0 * l.obs2[2];
  e[3] = u[3] * l.obs2[0] + v[3] * l.obs2[1] + 1.0 * l.obs2[2];
  e[4] = l.ins[0] - u[4];
  e[5] = l.ins[1] - v[4];
}
// row r of _jacobianOplusXj (EdgeLIL.h:338-374); rows 2 and 3 both take the END point of the second segment, as the
// reference does (it reads estimate().segment<3>(9) for the start as well, :276-279)
__device__ void lil_jacobian_row(const Pose& p, const psl_pose_lil& l, const Cam& cam, int r, double* J) {
  double X[3];
  lil_map(p, l, r == 0 ? 0 : r == 1 ? 1 : r < 4 ? 3 : 4, X);
  const double x = X[0], y = X[1], invz = 1.0 / X[2], invz_2 = invz * invz;
  if (r < 4) {
    const double l0 = r < 2 ? l.obs1[0] : l.obs2[0], l1 = r < 2 ? l.obs1[1] : l.obs2[1];
    J[0] = -cam.fx * x * y * invz_2 * l0 - cam.fy * (1 + y * y * invz_2) * l1;
    J[1] = cam.fx * (1 + x * x * invz_2) * l0 + cam.fy * x * y * invz_2 * l1;
    J[2] = -cam.fx * y * invz * l0 + cam.fy * x * invz * l1;
    J[3] = cam.fx * invz * l0;
    J[4] = cam.fy * invz * l1;
    J[5] = (-cam.fx * x * l0 - cam.fy * y * l1) * invz_2;
  } else if (r == 4) {
    J[0] = x * y * invz_2 * cam.fx; J[1] = -(1 + (x * x * invz_2)) * cam.fx; J[2] = y * invz * cam.fx;
    J[3] = -cam.fx * invz; J[4] = 0; J[5] = x * invz_2 * cam.fx;
  } else {
    J[0] = (1 + y * y * invz_2) * cam.fy; J[1] = -cam.fy * x * y * invz_2; J[2] = -cam.fy * x * invz;
    J[3] = 0; J[4] = -cam.fy * invz; J[5] = cam.fy * y * invz_2;
  }
}

__global__ void __launch_bounds__(kPoseThreads)
    pose_opt_kernel(const float* __restrict__ Tcw_in, const psl_pose_point* __restrict__ pts, const int32_t* __restrict__ n_pts,
                    int cap, const psl_pose_lil* __restrict__ lils, const int32_t* __restrict__ n_lils, int lil_cap, Cam cam,
                    double* __restrict__ err_scratch, uint8_t* __restrict__ level_scratch, double* __restrict__ lil_err_scratch,
                    uint8_t* __restrict__ lil_level_scratch, float* __restrict__ Tcw_out, uint8_t* __restrict__ outlier,
                    uint8_t* __restrict__ lil_outlier, int32_t* __restrict__ n_inliers) {
  __shared__ double s_part[(kPoseThreads / 32) * 27];
  __shared__ double s_sum[27];
  __shared__ Pose s_est, s_init, s_backup;
  __shared__ double s_lambda, s_ni, s_cur, s_ini, s_rho, s_x[6], s_H[36], s_b[6];
  __shared__ int s_flag, s_qmax, s_bad_steps, s_ok2;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = min(n_pts[b], cap);
  const psl_pose_point* P = pts + (size_t)b * cap;
  double* err = err_scratch + (size_t)b * cap * 3;
  uint8_t* level = level_scratch + (size_t)b * cap;
  uint8_t* out = outlier + (size_t)b * cap;
  const float* Tin = Tcw_in + (size_t)b * 16;
  float* Tout = Tcw_out + (size_t)b * 16;
  const double deltaMono = (double)(float)sqrt(5.991), deltaStereo = (double)(float)sqrt(7.815);   // const float delta = sqrt(5.991)
  const double deltaLil = (double)(float)sqrt(11.07);   // float deltaLJL = sqrt(11.07), Optimizer.cc:629
  // the structural-line edges of this frame (none when lils is NULL); they sit on the LAST threads of the block
  const int nl = lils ? min(n_lils[b], lil_cap) : 0;
  const psl_pose_lil* LL = lils + (size_t)b * lil_cap;
  double* lerr = lil_err_scratch + (size_t)b * lil_cap * 6;
  uint8_t* llevel = lil_level_scratch + (size_t)b * lil_cap;
  uint8_t* lout = lil_outlier + (size_t)b * lil_cap;
  const int ltid = kPoseThreads - 1 - tid;

  double cnt[1] = {0};
  for (int i = tid; i < n; i += kPoseThreads) {
    out[i] = 0;
    level[i] = (P[i].flags & 1u) ? 0 : 2;   // 2: no MapPoint, not an edge
    cnt[0] += (P[i].flags & 1u) ? 1.0 : 0.0;
  }
  for (int i = ltid; i < nl; i += kPoseThreads) {
    lout[i] = 0;
    llevel[i] = (LL[i].flags & 1u) ? 0 : 2;
    cnt[0] += (LL[i].flags & 1u) ? 1.0 : 0.0;
  }
  block_sum<1>(cnt, s_part, s_sum);
  const int n_initial = (int)s_sum[0];
  if (tid < 16) Tout[tid] = Tin[tid];
  if (n_initial < 3) {
    if (tid == 0) n_inliers[b] = 0;
    return;
  }
  if (tid == 0) {
    double R0[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R0[i][j] = Tin[4 * i + j];
    Pose p;
    p.q = quat_from_matrix(R0);
    normalize_rotation(p.q);
    for (int i = 0; i < 3; ++i) p.t[i] = Tin[4 * i + 3];
    s_init = p;
  }
  __syncthreads();

  auto compute_errors = [&](const Pose& p, int lvl) {   // errors of the edges at level `lvl`
    for (int i = tid; i < n; i += kPoseThreads) {
      if (level[i] != lvl) continue;
      const psl_pose_point pt = P[i];
      const double Xw[3] = {pt.xw, pt.yw, pt.zw};
      double X[3];
      quat_rotate(p.q, Xw, X);
      X[0] += p.t[0]; X[1] += p.t[1]; X[2] += p.t[2];
      if (!(pt.u_right < 0)) {
        const float invz = (float)(1.0 / X[2]);
        const double r0 = X[0] * invz * cam.fx + cam.cx, r1 = X[1] * invz * cam.fy + cam.cy;
        err[3 * i] = (double)pt.u - r0; err[3 * i + 1] = (double)pt.v - r1; err[3 * i + 2] = (double)pt.u_right - (r0 - cam.bf * invz);
      } else {
        const double px = X[0] / X[2], py = X[1] / X[2];
        err[3 * i] = (double)pt.u - (px * cam.fx + cam.cx); err[3 * i + 1] = (double)pt.v - (py * cam.fy + cam.cy); err[3 * i + 2] = 0;
      }
    }
    for (int i = ltid; i < nl; i += kPoseThreads)
      if (llevel[i] == lvl) lil_error(p, LL[i], cam, lerr + 6 * i);
  };
  auto lil_chi2_of = [&](int i) {
    double s2 = 0;
    for (int k = 0; k < 6; ++k) s2 += lerr[6 * i + k] * lerr[6 * i + k];
    return s2;   // information = identity
  };
  auto chi2_of = [&](int i) {
    return (err[3 * i] * err[3 * i] + err[3 * i + 1] * err[3 * i + 1] + err[3 * i + 2] * err[3 * i + 2]) * (double)P[i].inv_sigma2;
  };
  auto robust_chi2 = [&](bool robust) {   // result in s_sum[0]
    double c[1] = {0};
    for (int i = tid; i < n; i += kPoseThreads) {
      if (level[i] != 0) continue;
      const double e2 = chi2_of(i);
      if (robust) { double r0, r1; huber(e2, P[i].u_right < 0 ? deltaMono : deltaStereo, r0, r1); c[0] += r0; }
      else c[0] += e2;
    }
    for (int i = ltid; i < nl; i += kPoseThreads) {
      if (llevel[i] != 0) continue;
      const double e2 = lil_chi2_of(i);
      if (robust) { double r0, r1; huber(e2, deltaLil, r0, r1); c[0] += r0; }
      else c[0] += e2;
    }
    block_sum<1>(c, s_part, s_sum);
  };

  int n_bad = 0;
  for (int it = 0; it < 4; ++it) {
    const bool robust = it < 3;   // the kernels are dropped after the third round (:806, :836)
    if (tid == 0) { s_est = s_init; s_bad_steps = 0; }
    __syncthreads();
    for (int iter = 0; iter < 10; ++iter) {
      compute_errors(s_est, 0);
      __syncthreads();
      robust_chi2(robust);
      if (tid == 0) { s_cur = s_sum[0]; s_ini = s_sum[0]; }
      // H (upper triangle, 21) and b (6)
      double acc[27];
#pragma unroll
      for (int k = 0; k < 27; ++k) acc[k] = 0;
      const Pose p = s_est;
      for (int i = tid; i < n; i += kPoseThreads) {
        if (level[i] != 0) continue;
        const psl_pose_point pt = P[i];
        const bool stereo = !(pt.u_right < 0);
        const double Xw[3] = {pt.xw, pt.yw, pt.zw};
        double X[3];
        quat_rotate(p.q, Xw, X);
        const double x = X[0] + p.t[0], y = X[1] + p.t[1], invz = 1.0 / (X[2] + p.t[2]), invz_2 = invz * invz;
        double J[3][6];
        J[0][0] = x * y * invz_2 * cam.fx; J[0][1] = -(1 + (x * x * invz_2)) * cam.fx; J[0][2] = y * invz * cam.fx;
        J[0][3] = -invz * cam.fx; J[0][4] = 0; J[0][5] = x * invz_2 * cam.fx;
        J[1][0] = (1 + y * y * invz_2) * cam.fy; J[1][1] = -x * y * invz_2 * cam.fy; J[1][2] = -x * invz * cam.fy;
        J[1][3] = 0; J[1][4] = -invz * cam.fy; J[1][5] = y * invz_2 * cam.fy;
        if (stereo) {
          J[2][0] = J[0][0] - cam.bf * y * invz_2; J[2][1] = J[0][1] + cam.bf * x * invz_2; J[2][2] = J[0][2];
          J[2][3] = J[0][3]; J[2][4] = 0; J[2][5] = J[0][5] - cam.bf * invz_2;
        } else {
          for (int k = 0; k < 6; ++k) J[2][k] = 0;
        }
        double w = 1.0, r0;
        if (robust) huber(chi2_of(i), stereo ? deltaStereo : deltaMono, r0, w);
        const double info = (double)pt.inv_sigma2;
        const double e0 = err[3 * i], e1 = err[3 * i + 1], e2 = err[3 * i + 2];
        int k = 0;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
#pragma unroll
          for (int q = r; q < 6; ++q) acc[k++] += (J[0][r] * J[0][q] + J[1][r] * J[1][q] + J[2][r] * J[2][q]) * (w * info);
        }
#pragma unroll
        for (int r = 0; r < 6; ++r) acc[21 + r] -= w * ((J[0][r] * e0 + J[1][r] * e1 + J[2][r] * e2) * info);
      }
      // BaseBinaryEdge::constructQuadraticForm with the LIL vertex fixed: only the pose block (base_binary_edge.hpp:58-120)
      for (int i = ltid; i < nl; i += kPoseThreads) {
        if (llevel[i] != 0) continue;
        double w = 1.0, r0;
        if (robust) huber(lil_chi2_of(i), deltaLil, r0, w);
        for (int d = 0; d < 6; ++d) {
          double J[6];
          lil_jacobian_row(p, LL[i], cam, d, J);
          const double ed = lerr[6 * i + d];
          int k = 0;
#pragma unroll
          for (int r = 0; r < 6; ++r) {
#pragma unroll
            for (int q = r; q < 6; ++q) acc[k++] += J[r] * w * J[q];
          }
#pragma unroll
          for (int r = 0; r < 6; ++r) acc[21 + r] -= w * (J[r] * ed);
        }
      }
      block_sum<27>(acc, s_part, s_sum);
      if (tid == 0) {
        int k = 0;
        for (int r = 0; r < 6; ++r)
          for (int q = r; q < 6; ++q) { s_H[r * 6 + q] = s_sum[k]; s_H[q * 6 + r] = s_sum[k]; ++k; }
        for (int r = 0; r < 6; ++r) s_b[r] = s_sum[21 + r];
        if (iter == 0) {
          double maxDiag = 0;
          for (int j = 0; j < 6; ++j) maxDiag = fmax(fabs(s_H[j * 6 + j]), maxDiag);
          s_lambda = 1e-5 * maxDiag;
          s_ni = 2;
          s_bad_steps = 0;
        }
        s_qmax = 0;
        s_rho = 0;
      }
      __syncthreads();
      while (true) {   // the Levenberg trials (optimization_algorithm_levenberg.cpp:100-140)
        if (tid == 0) {
          s_backup = s_est;
          double Hl[36];
          for (int j = 0; j < 36; ++j) Hl[j] = s_H[j];
          for (int j = 0; j < 6; ++j) Hl[j * 6 + j] += s_lambda;
          for (int j = 0; j < 6; ++j) s_x[j] = 0;
          s_ok2 = solve6(Hl, s_b, s_x) ? 1 : 0;
          s_est = pose_exp_times(s_x, s_est);
        }
        __syncthreads();
        compute_errors(s_est, 0);
        __syncthreads();
        robust_chi2(robust);
        if (tid == 0) {
          double tempChi = s_sum[0];
          if (!s_ok2) tempChi = DBL_MAX;
          double rho = s_cur - tempChi;
          double scale = 0;
          for (int j = 0; j < 6; ++j) scale += s_x[j] * (s_lambda * s_x[j] + s_b[j]);
          scale += 1e-3;
          rho /= scale;
          if (rho > 0 && isfinite(tempChi)) {
            double alpha = 1. - pow((2 * rho - 1), 3.0);
            alpha = fmin(alpha, 2. / 3.);
            s_lambda *= fmax(1. / 3., alpha);
            s_ni = 2;
            s_cur = tempChi;
          } else {
            s_lambda *= s_ni;
            s_ni *= 2;
            s_est = s_backup;   // pop(): the edges keep the errors of the rejected trial
          }
          s_rho = rho;
          s_qmax += 1;
          s_flag = (rho < 0 && s_qmax < 10) ? 1 : 0;
        }
        __syncthreads();
        if (!s_flag) break;
      }
      __syncthreads();   // everybody has read the trial flag before thread 0 reuses it
      if (tid == 0) {
        int stop = 0;
        if (s_qmax == 10 || s_rho == 0) stop = 1;
        else {
          if ((s_ini - s_cur) * 1e3 < s_ini) s_bad_steps += 1; else s_bad_steps = 0;
          if (s_bad_steps >= 3) stop = 1;
        }
        s_flag = stop;
      }
      __syncthreads();
      if (s_flag) break;
    }
    __syncthreads();
    // outlier classification (Optimizer.cc:782-838): level-1 edges are re-evaluated at the estimate, the others keep
    // the errors of the last trial
    compute_errors(s_est, 1);
    __syncthreads();
    double nb[1] = {0};
    for (int i = tid; i < n; i += kPoseThreads) {
      if (level[i] > 1) continue;
      const float c2 = (float)chi2_of(i);
      const bool bad = c2 > (P[i].u_right < 0 ? 5.991f : 7.815f);
      out[i] = bad ? 1 : 0;
      level[i] = bad ? 1 : 0;
      nb[0] += bad ? 1.0 : 0.0;
    }
    for (int i = ltid; i < nl; i += kPoseThreads) {   // :976-1007; not counted in nBad (nLineBad stays 0)
      if (llevel[i] > 1) continue;
      const bool bad = (float)lil_chi2_of(i) > 11.07f;
      lout[i] = bad ? 1 : 0;
      llevel[i] = bad ? 1 : 0;
    }
    block_sum<1>(nb, s_part, s_sum);
    n_bad = (int)s_sum[0];
    if (n_initial < 10) break;
    __syncthreads();
  }
  if (tid == 0) {
    double R[3][3];
    quat_to_matrix(s_est.q, R);
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) Tout[4 * i + j] = (float)R[i][j];
      Tout[4 * i + 3] = (float)s_est.t[i];
    }
    Tout[12] = Tout[13] = Tout[14] = 0.f;
    Tout[15] = 1.f;
    n_inliers[b] = n_initial - n_bad;
  }
}

}  // namespace
}  // namespace psl

using namespace psl;

extern "C" {

int psl_pose_optimization_lil_dev(psl_ctx* ctx, const float* d_Tcw_in, const psl_pose_point* d_pts, const int32_t* d_n,
                                  int32_t cap, const psl_pose_lil* d_lils, const int32_t* d_n_lil, int32_t lil_cap, int32_t B,
                                  float fx, float fy, float cx, float cy, float bf, float* d_Tcw_out, uint8_t* d_outlier,
                                  uint8_t* d_lil_outlier, int32_t* d_n_inliers) {
  if (!ctx) return PSL_E_INVALID;
  if (B < 0 || cap < 1 || lil_cap < 0 ||
      (B > 0 && (!d_Tcw_in || !d_pts || !d_n || !d_Tcw_out || !d_outlier || !d_n_inliers)) ||
      (B > 0 && lil_cap > 0 && (!d_lils || !d_n_lil || !d_lil_outlier)))
    return fail(ctx, PSL_E_INVALID, "bad argument");
  if (B == 0) return PSL_OK;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  int rc;
  if ((rc = ensure(ctx, ctx->m_misc[10], (size_t)B * cap * 24))) return rc;
  if ((rc = ensure(ctx, ctx->m_misc[11], (size_t)B * cap))) return rc;
  if ((rc = ensure(ctx, ctx->m_misc[12], (size_t)B * std::max(lil_cap, 1) * 48))) return rc;
  if ((rc = ensure(ctx, ctx->m_misc[13], (size_t)B * std::max(lil_cap, 1)))) return rc;
  size_t e = prof_mark(ctx);
  pose_opt_kernel<<<B, kPoseThreads, 0, ctx->stream>>>(d_Tcw_in, d_pts, d_n, cap, lil_cap > 0 ? d_lils : nullptr, d_n_lil, lil_cap,
                                                       Cam{fx, fy, cx, cy, bf}, ctx->m_misc[10].as<double>(),
                                                       ctx->m_misc[11].as<uint8_t>(), ctx->m_misc[12].as<double>(),
                                                       ctx->m_misc[13].as<uint8_t>(), d_Tcw_out, d_outlier, d_lil_outlier,
                                                       d_n_inliers);
  prof_span(ctx, 15, e, 1);
  PSL_CK(cudaGetLastError());
  return PSL_OK;
}

int psl_pose_optimization_dev(psl_ctx* ctx, const float* d_Tcw_in, const psl_pose_point* d_pts, const int32_t* d_n,
                              int32_t cap, int32_t B, float fx, float fy, float cx, float cy, float bf, float* d_Tcw_out,
                              uint8_t* d_outlier, int32_t* d_n_inliers) {
  return psl_pose_optimization_lil_dev(ctx, d_Tcw_in, d_pts, d_n, cap, nullptr, nullptr, 0, B, fx, fy, cx, cy, bf, d_Tcw_out,
                                       d_outlier, nullptr, d_n_inliers);
}

int psl_pose_optimization_lil(psl_ctx* ctx, const float* Tcw_in, const psl_pose_point* pts, int32_t n,
                              const psl_pose_lil* lils, int32_t n_lil, float fx, float fy, float cx, float cy, float bf,
                              float* Tcw_out, uint8_t* outlier, uint8_t* lil_outlier, int32_t* n_inliers) {
  if (!ctx) return PSL_E_INVALID;
  if (!Tcw_in || !Tcw_out || !n_inliers || n < 0 || n_lil < 0 || (n > 0 && (!pts || !outlier)) ||
      (n_lil > 0 && (!lils || !lil_outlier)))
    return fail(ctx, PSL_E_INVALID, "bad argument");
  *n_inliers = 0;
  for (int i = 0; i < 16; ++i) Tcw_out[i] = Tcw_in[i];
  for (int i = 0; i < n; ++i) outlier[i] = 0;
  for (int i = 0; i < n_lil; ++i) lil_outlier[i] = 0;
  if (n + n_lil == 0) return PSL_OK;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  DevBuf* M = ctx->m_misc;
  int rc;
#define PSL_UPP(buf, src, nbytes)                                                                              \
  do {                                                                                                         \
    if ((rc = ensure(ctx, buf, (nbytes)))) return rc;                                                          \
    if ((nbytes) > 0) PSL_CK(cudaMemcpyAsync((buf).p, (src), (nbytes), cudaMemcpyHostToDevice, ctx->stream));  \
  } while (0)
  const int cap = std::max(n, 1);
  PSL_UPP(M[0], Tcw_in, 64);
  PSL_UPP(M[1], pts, (size_t)n * sizeof(psl_pose_point));
  PSL_UPP(M[4], lils, (size_t)n_lil * sizeof(psl_pose_lil));
  const int32_t nn[2] = {n, n_lil};
  PSL_UPP(ctx->m_n, nn, sizeof(nn));
  if ((rc = ensure(ctx, M[2], 64))) return rc;
  if ((rc = ensure(ctx, M[3], (size_t)cap))) return rc;
  if ((rc = ensure(ctx, M[5], (size_t)std::max(n_lil, 1)))) return rc;
  if ((rc = ensure(ctx, ctx->m_nm, 4))) return rc;
  rc = psl_pose_optimization_lil_dev(ctx, M[0].as<float>(), M[1].as<psl_pose_point>(), ctx->m_n.as<int32_t>(), cap,
                                     M[4].as<psl_pose_lil>(), ctx->m_n.as<int32_t>() + 1, n_lil, 1, fx, fy, cx, cy, bf,
                                     M[2].as<float>(), M[3].as<uint8_t>(), M[5].as<uint8_t>(), ctx->m_nm.as<int32_t>());
  if (rc) return rc;
  PSL_CK(cudaMemcpyAsync(Tcw_out, M[2].p, 64, cudaMemcpyDeviceToHost, ctx->stream));
  if (n > 0) PSL_CK(cudaMemcpyAsync(outlier, M[3].p, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  if (n_lil > 0) PSL_CK(cudaMemcpyAsync(lil_outlier, M[5].p, (size_t)n_lil, cudaMemcpyDeviceToHost, ctx->stream));
  PSL_CK(cudaMemcpyAsync(n_inliers, ctx->m_nm.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  return check_status(ctx);
#undef PSL_UPP
}

int psl_pose_optimization(psl_ctx* ctx, const float* Tcw_in, const psl_pose_point* pts, int32_t n, float fx, float fy,
                          float cx, float cy, float bf, float* Tcw_out, uint8_t* outlier, int32_t* n_inliers) {
  return psl_pose_optimization_lil(ctx, Tcw_in, pts, n, nullptr, 0, fx, fy, cx, cy, bf, Tcw_out, outlier, nullptr, n_inliers);
}

}  // extern "C"
