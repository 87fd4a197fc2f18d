// TEST INFRASTRUCTURE — CPU oracle of Optimizer::PoseOptimization ("next" row N4): point edges and the structural-line
// (LIL) edges.  Restates /root/reference/src/Optimizer.cc:239-1023 (the live LIL block is :619-693, its classification
// :976-1007; the long blocks before them are commented-out code) and add_inc/EdgeLIL.h:210-374 over the vendored g2o it calls: EdgeSE3ProjectXYZOnlyPose / EdgeStereoSE3ProjectXYZOnlyPose
// (Thirdparty/g2o/g2o/types/types_six_dof_expmap.{h,cpp}), SE3Quat (types/se3quat.h), RobustKernelHuber
// (core/robust_kernel_impl.cpp:78-98), BaseUnaryEdge::constructQuadraticForm (core/base_unary_edge.hpp:43-72),
// OptimizationAlgorithmLevenberg::solve (core/optimization_algorithm_levenberg.cpp:61-190), SparseOptimizer::optimize
// (core/sparse_optimizer.cpp:354-419).  g2o needs Eigen, which is not in this image, so this restatement is UNPINNED:
// the 6x6 solve is a plain LDL^T (Eigen's pivoted LDLT differs in rounding only), the sums run in edge order.
// Nothing here is used by the product.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "psl_oracle.h"

namespace {

struct Quat { double w, x, y, z; };
struct Pose { Quat q; double t[3]; };

void normalize_rotation(Quat& q) {  // SE3Quat::normalizeRotation
  if (q.w < 0) { q.w = -q.w; q.x = -q.x; q.y = -q.y; q.z = -q.z; }
  const double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  q.w /= n; q.x /= n; q.y /= n; q.z /= n;
}

Quat quat_from_matrix(const double m[3][3]) {  // Eigen::Quaterniond(Matrix3d)
  Quat q;
  double t = m[0][0] + m[1][1] + m[2][2];
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (m[2][1] - m[1][2]) * t;
    q.y = (m[0][2] - m[2][0]) * t;
    q.z = (m[1][0] - m[0][1]) * t;
  } else {
    int i = 0;
    if (m[1][1] > m[0][0]) i = 1;
    if (m[2][2] > m[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
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

void quat_rotate(const Quat& q, const double v[3], double out[3]) {  // Eigen: v + w * uv + cross(q.vec, uv), uv = 2 cross(q.vec, v)
  const double ux = 2 * (q.y * v[2] - q.z * v[1]), uy = 2 * (q.z * v[0] - q.x * v[2]), uz = 2 * (q.x * v[1] - q.y * v[0]);
  out[0] = v[0] + q.w * ux + (q.y * uz - q.z * uy);
  out[1] = v[1] + q.w * uy + (q.z * ux - q.x * uz);
  out[2] = v[2] + q.w * uz + (q.x * uy - q.y * ux);
}

Quat quat_mul(const Quat& a, const Quat& b) {
  return Quat{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
              a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}

void quat_to_matrix(const Quat& q, double R[3][3]) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y,
               tyz = tz * q.y, tzz = tz * q.z;
  R[0][0] = 1 - (tyy + tzz); R[0][1] = txy - twz; R[0][2] = txz + twy;
  R[1][0] = txy + twz; R[1][1] = 1 - (txx + tzz); R[1][2] = tyz - twx;
  R[2][0] = txz - twy; R[2][1] = tyz + twx; R[2][2] = 1 - (txx + tyy);
}

Pose pose_exp_times(const double u[6], const Pose& est) {  // VertexSE3Expmap::oplusImpl: exp(update) * estimate
  const double om[3] = {u[0], u[1], u[2]}, up[3] = {u[3], u[4], u[5]};
  const double theta = std::sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
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
        R[i][j] = I + std::sin(theta) / theta * O[i][j] + (1 - std::cos(theta)) / (theta * theta) * O2[i][j];
        V[i][j] = I + (1 - std::cos(theta)) / (theta * theta) * O[i][j] + (theta - std::sin(theta)) / std::pow(theta, 3) * O2[i][j];
      }
    }
  Pose e;
  e.q = quat_from_matrix(R);
  normalize_rotation(e.q);
  for (int i = 0; i < 3; ++i) e.t[i] = V[i][0] * up[0] + V[i][1] * up[1] + V[i][2] * up[2];
  Pose r = e;  // SE3Quat::operator*
  double rt[3];
  quat_rotate(e.q, est.t, rt);
  for (int i = 0; i < 3; ++i) r.t[i] += rt[i];
  r.q = quat_mul(e.q, est.q);
  normalize_rotation(r.q);
  return r;
}

struct Edge { double obs[3]; double Xw[3]; double info; bool stereo; double delta; int level; bool robust; double err[3]; };

struct Cam { double fx, fy, cx, cy, bf; };

void compute_error(Edge& e, const Pose& p, const Cam& c) {
  double X[3];
  quat_rotate(p.q, e.Xw, X);
  for (int i = 0; i < 3; ++i) X[i] += p.t[i];
  if (e.stereo) {  // EdgeStereoSE3ProjectXYZOnlyPose::cam_project: invz is a float
    const float invz = (float)(1.0 / X[2]);  // `const float invz = 1.0f / z`: double division, narrowed to float
    const double r0 = X[0] * invz * c.fx + c.cx, r1 = X[1] * invz * c.fy + c.cy;
    e.err[0] = e.obs[0] - r0; e.err[1] = e.obs[1] - r1; e.err[2] = e.obs[2] - (r0 - c.bf * invz);
  } else {         // project2d + intrinsics
    const double px = X[0] / X[2], py = X[1] / X[2];
    e.err[0] = e.obs[0] - (px * c.fx + c.cx); e.err[1] = e.obs[1] - (py * c.fy + c.cy); e.err[2] = 0;
  }
}
double chi2(const Edge& e) { return (e.err[0] * e.err[0] + e.err[1] * e.err[1] + (e.stereo ? e.err[2] * e.err[2] : 0.0)) * e.info; }
void huber(double e2, double delta, double rho[3]) {  // RobustKernelHuber::robustify
  const double dsqr = delta * delta;
  if (e2 <= dsqr) { rho[0] = e2; rho[1] = 1.; rho[2] = 0.; }
  else { const double sqrte = std::sqrt(e2); rho[0] = 2 * sqrte * delta - dsqr; rho[1] = delta / sqrte; rho[2] = -0.5 * rho[1] / e2; }
}
double robust_chi2(const std::vector<Edge>& E) {
  double s = 0;
  for (const Edge& e : E) {
    if (e.level != 0) continue;
    const double c = chi2(e);
    if (e.robust) { double rho[3]; huber(c, e.delta, rho); s += rho[0]; } else s += c;
  }
  return s;
}
void jacobian(const Edge& e, const Pose& p, const Cam& c, double J[3][6]) {
  double X[3];
  quat_rotate(p.q, e.Xw, X);
  for (int i = 0; i < 3; ++i) X[i] += p.t[i];
  const double x = X[0], y = X[1], invz = 1.0 / X[2], invz_2 = invz * invz;
  J[0][0] = x * y * invz_2 * c.fx; J[0][1] = -(1 + (x * x * invz_2)) * c.fx; J[0][2] = y * invz * c.fx;
  J[0][3] = -invz * c.fx; J[0][4] = 0; J[0][5] = x * invz_2 * c.fx;
  J[1][0] = (1 + y * y * invz_2) * c.fy; J[1][1] = -x * y * invz_2 * c.fy; J[1][2] = -x * invz * c.fy;
  J[1][3] = 0; J[1][4] = -invz * c.fy; J[1][5] = y * invz_2 * c.fy;
  if (e.stereo) {
    J[2][0] = J[0][0] - c.bf * y * invz_2; J[2][1] = J[0][1] + c.bf * x * invz_2; J[2][2] = J[0][2];
    J[2][3] = J[0][3]; J[2][4] = 0; J[2][5] = J[0][5] - c.bf * invz_2;
  } else {
    for (int k = 0; k < 6; ++k) J[2][k] = 0;
  }
}
void build_system(const std::vector<Edge>& E, const Pose& p, const Cam& c, double H[6][6], double b[6]) {
  std::memset(H, 0, 36 * sizeof(double));
  std::memset(b, 0, 6 * sizeof(double));
  for (const Edge& e : E) {
    if (e.level != 0) continue;
    double J[3][6];
    jacobian(e, p, c, J);
    double w = 1.0;  // rho[1]
    if (e.robust) { double rho[3]; huber(chi2(e), e.delta, rho); w = rho[1]; }
    const int D = e.stereo ? 3 : 2;
    for (int r = 0; r < 6; ++r) {
      double s = 0;
      for (int d = 0; d < D; ++d) s += J[d][r] * e.info * e.err[d];
      b[r] -= w * s;
      for (int q = 0; q < 6; ++q) {
        double h = 0;
        for (int d = 0; d < D; ++d) h += J[d][r] * (w * e.info) * J[d][q];
        H[r][q] += h;
      }
    }
  }
}
// ---- EdgeLILSE3ProjectXYZ (add_inc/EdgeLIL.h:210-374): a fixed VertexLIL (two 3-D segments + their cross point) seen from
// the pose; 6-dim error = (distance of the projected end points of each segment to the observed 2-D line, 2 + 2, and
// the reprojection error of the cross point, 2), information = identity (invSigma = 1, Optimizer.cc:241, :668)
struct LilEdge { double Xw[15]; double l1[3], l2[3], ins[2]; int level; bool robust; double err[6]; };

void map_point(const Pose& p, const double* xw, double X[3]) {  // SE3Quat::map
  quat_rotate(p.q, xw, X);
  for (int i = 0; i < 3; ++i) X[i] += p.t[i];
}
void lil_error(LilEdge& e, const Pose& p, const Cam& c) {  // computeError, EdgeLIL.h:220-259
  double uv[5][2];
  for (int k = 0; k < 5; ++k) {
    double X[3];
    map_point(p, e.Xw + 3 * k, X);
    uv[k][0] = X[0] / X[2] * c.fx + c.cx;   // cam_project = project2d, then the intrinsics
    uv[k][1] = X[1] / X[2] * c.fy + c.cy;
  }
  e.err[0] = uv[0][0] * e.l1[0] + uv[0][1] * e.l1[1] + 1.0 * e.l1[2];
  e.err[1] = uv[1][0] * e.l1[0] + uv[1][1] * e.l1[1] + 1.0 * e.l1[2];
  e.err[2] = uv[2][0] * e.l2[0] + uv[2][1] * e.l2[1] + 1.0 * e.l2[2];
  e.err[3] = uv[3][0] * e.l2[0] + uv[3][1] * e.l2[1] + 1.0 * e.l2[2];
  e.err[4] = e.ins[0] - uv[4][0];
  e.err[5] = e.ins[1] - uv[4][1];
}
double lil_chi2(const LilEdge& e) {
  double s = 0;
  for (int k = 0; k < 6; ++k) s += e.err[k] * e.err[k];
  return s;
}
// linearizeOplus, the pose block _jacobianOplusXj (:338-374).  The reference reads BOTH end points of the second segment
// from estimate().segment<3>(9) (:276-279), so rows 2 and 3 are the derivative at the segment's END point; kept.
void lil_jacobian(const LilEdge& e, const Pose& p, const Cam& c, double J[6][6]) {
  const int src[4] = {0, 3, 9, 9};
  for (int r = 0; r < 4; ++r) {
    double X[3];
    map_point(p, e.Xw + src[r], X);
    const double x = X[0], y = X[1], invz = 1.0 / X[2], invz_2 = invz * invz;
    const double l0 = r < 2 ? e.l1[0] : e.l2[0], l1 = r < 2 ? e.l1[1] : e.l2[1];
    J[r][0] = -c.fx * x * y * invz_2 * l0 - c.fy * (1 + y * y * invz_2) * l1;
    J[r][1] = c.fx * (1 + x * x * invz_2) * l0 + c.fy * x * y * invz_2 * l1;
    J[r][2] = -c.fx * y * invz * l0 + c.fy * x * invz * l1;
    J[r][3] = c.fx * invz * l0;
    J[r][4] = c.fy * invz * l1;
    J[r][5] = (-c.fx * x * l0 - c.fy * y * l1) * invz_2;
  }
  double X[3];
  map_point(p, e.Xw + 12, X);
  const double x = X[0], y = X[1], invz = 1.0 / X[2], invz_2 = invz * invz;
  J[4][0] = x * y * invz_2 * c.fx; J[4][1] = -(1 + (x * x * invz_2)) * c.fx; J[4][2] = y * invz * c.fx;
  J[4][3] = -c.fx * invz; J[4][4] = 0; J[4][5] = x * invz_2 * c.fx;
  J[5][0] = (1 + y * y * invz_2) * c.fy; J[5][1] = -c.fy * x * y * invz_2; J[5][2] = -c.fy * x * invz;
  J[5][3] = 0; J[5][4] = -c.fy * invz; J[5][5] = c.fy * y * invz_2;
}
// BaseBinaryEdge::constructQuadraticForm with the LIL vertex fixed (core/base_binary_edge.hpp:58-120): only the pose block
void add_lil_edges(const std::vector<LilEdge>& E, const Pose& p, const Cam& c, double delta, double H[6][6], double b[6]) {
  for (const LilEdge& e : E) {
    if (e.level != 0) continue;
    double J[6][6];
    lil_jacobian(e, p, c, J);
    double w = 1.0;
    if (e.robust) { double rho[3]; huber(lil_chi2(e), delta, rho); w = rho[1]; }
    for (int r = 0; r < 6; ++r) {
      double s = 0;
      for (int d = 0; d < 6; ++d) s += J[d][r] * e.err[d];
      b[r] -= w * s;
      for (int q = 0; q < 6; ++q) {
        double h = 0;
        for (int d = 0; d < 6; ++d) h += J[d][r] * w * J[d][q];
        H[r][q] += h;
      }
    }
  }
}
double lil_robust_chi2(const std::vector<LilEdge>& E, double delta) {
  double s = 0;
  for (const LilEdge& e : E) {
    if (e.level != 0) continue;
    const double c = lil_chi2(e);
    if (e.robust) { double rho[3]; huber(c, delta, rho); s += rho[0]; } else s += c;
  }
  return s;
}

bool solve6(const double Hin[6][6], const double b[6], double x[6]) {  // LDL^T, fails unless positive (LinearSolverDense)
  double L[6][6] = {{0}}, D[6];
  for (int j = 0; j < 6; ++j) {
    double d = Hin[j][j];
    for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k] * D[k];
    if (!(d > 0)) return false;
    D[j] = d;
    L[j][j] = 1;
    for (int i = j + 1; i < 6; ++i) {
      double s = Hin[i][j];
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

}  // namespace

extern "C" {

// pts: n records (u, v, u_right (< 0: monocular observation), inv_sigma2, Xw[3], valid) = the keypoints with a MapPoint.
// Tcw: 4x4 row-major float in / out.  outlier[n] = mvbOutlier.  Returns nInitialCorrespondences - nBad (0 if < 3).
// lils: n_lil records, one per structural line of the frame (flags & 1: it holds a map InsectLine that is not bad);
// lil_outlier[n_lil] = mvbOutlier_Insec.  The return value does not subtract the LIL outliers (nLineBad stays 0, :712,1021).
int orc_pose_optimization_lil(const float* Tcw_in, const psl_pose_point* pts, int n, const psl_pose_lil* lils, int n_lil,
                              float fx, float fy, float cx, float cy, float bf, float* Tcw_out, uint8_t* outlier,
                              uint8_t* lil_outlier) {
  const Cam cam{fx, fy, cx, cy, bf};
  std::vector<LilEdge> L;
  std::vector<int> lidx;
  const double deltaLJL = (double)(float)std::sqrt(11.07);  // float deltaLJL = sqrt(11.07), :629
  for (int i = 0; i < n_lil; ++i) {
    lil_outlier[i] = 0;
    if (!(lils[i].flags & 1u)) continue;
    LilEdge e{};
    for (int k = 0; k < 6; ++k) { e.Xw[k] = lils[i].line1[k]; e.Xw[6 + k] = lils[i].line2[k]; }
    for (int k = 0; k < 3; ++k) { e.Xw[12 + k] = lils[i].cross[k]; e.l1[k] = lils[i].obs1[k]; e.l2[k] = lils[i].obs2[k]; }
    e.ins[0] = lils[i].ins[0]; e.ins[1] = lils[i].ins[1];
    e.level = 0; e.robust = true;
    L.push_back(e); lidx.push_back(i);
  }
  std::vector<Edge> E;
  std::vector<int> idx;
  const float deltaMono = (float)std::sqrt(5.991), deltaStereo = (float)std::sqrt(7.815);  // :276-277
  for (int i = 0; i < n; ++i) {
    outlier[i] = 0;
    if (!(pts[i].flags & 1u)) continue;
    Edge e{};
    e.stereo = !(pts[i].u_right < 0);
    e.obs[0] = pts[i].u; e.obs[1] = pts[i].v; e.obs[2] = e.stereo ? pts[i].u_right : 0;
    e.Xw[0] = pts[i].xw; e.Xw[1] = pts[i].yw; e.Xw[2] = pts[i].zw;
    e.info = pts[i].inv_sigma2;
    e.delta = e.stereo ? deltaStereo : deltaMono;
    e.level = 0; e.robust = true;
    E.push_back(e); idx.push_back(i);
  }
  std::memcpy(Tcw_out, Tcw_in, 16 * sizeof(float));
  const int nInitial = (int)E.size() + (int)L.size();
  if (nInitial < 3) return 0;
  double R0[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R0[i][j] = Tcw_in[4 * i + j];
  Pose init;
  init.q = quat_from_matrix(R0);
  normalize_rotation(init.q);
  for (int i = 0; i < 3; ++i) init.t[i] = Tcw_in[4 * i + 3];
  Pose est = init;
  const float chi2Mono = 5.991f, chi2Stereo = 7.815f;
  int nBad = 0;
  for (int it = 0; it < 4; ++it) {
    est = init;  // vSE3->setEstimate(toSE3Quat(pFrame->mTcw))
    double lambda = 0, ni = 2;
    int nBadSteps = 0;
    for (int iter = 0; iter < 10; ++iter) {
      for (Edge& e : E) if (e.level == 0) compute_error(e, est, cam);
      for (LilEdge& e : L) if (e.level == 0) lil_error(e, est, cam);
      double currentChi = robust_chi2(E) + lil_robust_chi2(L, deltaLJL);
      const double iniChi = currentChi;
      double H[6][6], b[6];
      build_system(E, est, cam, H, b);
      add_lil_edges(L, est, cam, deltaLJL, H, b);
      if (iter == 0) {
        double maxDiag = 0;
        for (int j = 0; j < 6; ++j) maxDiag = std::max(std::fabs(H[j][j]), maxDiag);
        lambda = 1e-5 * maxDiag;
        ni = 2;
        nBadSteps = 0;
      }
      double rho = 0;
      int qmax = 0;
      do {
        const Pose backup = est;
        double Hl[6][6];
        std::memcpy(Hl, H, sizeof(Hl));
        for (int j = 0; j < 6; ++j) Hl[j][j] += lambda;
        double x[6] = {0, 0, 0, 0, 0, 0};
        const bool ok2 = solve6(Hl, b, x);
        est = pose_exp_times(x, est);
        for (Edge& e : E) if (e.level == 0) compute_error(e, est, cam);
        for (LilEdge& e : L) if (e.level == 0) lil_error(e, est, cam);
        double tempChi = robust_chi2(E) + lil_robust_chi2(L, deltaLJL);
        if (!ok2) tempChi = std::numeric_limits<double>::max();
        rho = currentChi - tempChi;
        double scale = 0;
        for (int j = 0; j < 6; ++j) scale += x[j] * (lambda * x[j] + b[j]);
        scale += 1e-3;
        rho /= scale;
        if (rho > 0 && std::isfinite(tempChi)) {
          double alpha = 1. - std::pow((2 * rho - 1), 3);
          alpha = std::min(alpha, 2. / 3.);
          const double scaleFactor = std::max(1. / 3., alpha);
          lambda *= scaleFactor;
          ni = 2;
          currentChi = tempChi;
        } else {
          lambda *= ni;
          ni *= 2;
          est = backup;  // pop(): the vertex goes back, the edges keep the errors of the rejected trial
        }
        ++qmax;
      } while (rho < 0 && qmax < 10);
      if (qmax == 10 || rho == 0) break;  // Terminate
      if ((iniChi - currentChi) * 1e3 < iniChi) ++nBadSteps; else nBadSteps = 0;
      if (nBadSteps >= 3) break;
    }
    nBad = 0;
    for (size_t k = 0; k < E.size(); ++k) {  // :782-838
      Edge& e = E[k];
      if (outlier[idx[k]]) compute_error(e, est, cam);  // level-1 edges were not touched by the optimiser
      const float c2 = (float)chi2(e);
      if (c2 > (e.stereo ? chi2Stereo : chi2Mono)) { outlier[idx[k]] = 1; e.level = 1; ++nBad; }
      else { outlier[idx[k]] = 0; e.level = 0; }
      if (it == 2) e.robust = false;
    }
    for (size_t k = 0; k < L.size(); ++k) {  // :976-1007
      LilEdge& e = L[k];
      if (lil_outlier[lidx[k]]) lil_error(e, est, cam);
      const float c2 = (float)lil_chi2(e);
      if (c2 > 11.07f) { lil_outlier[lidx[k]] = 1; e.level = 1; }
      else { lil_outlier[lidx[k]] = 0; e.level = 0; }
      if (it == 2) e.robust = false;
    }
    if (E.size() + L.size() < 10) break;  // optimizer.edges().size() < 10
  }
  double R[3][3];
  quat_to_matrix(est.q, R);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) Tcw_out[4 * i + j] = (float)R[i][j];
    Tcw_out[4 * i + 3] = (float)est.t[i];
  }
  Tcw_out[12] = Tcw_out[13] = Tcw_out[14] = 0.f;
  Tcw_out[15] = 1.f;
  return nInitial - nBad;
}

int orc_pose_optimization(const float* Tcw_in, const psl_pose_point* pts, int n, float fx, float fy, float cx, float cy,
                          float bf, float* Tcw_out, uint8_t* outlier) {
  return orc_pose_optimization_lil(Tcw_in, pts, n, nullptr, 0, fx, fy, cx, cy, bf, Tcw_out, outlier, nullptr);
}

}  // extern "C"
