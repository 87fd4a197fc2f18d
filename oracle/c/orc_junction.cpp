// TEST INFRASTRUCTURE — CPU oracle of the line-junction detection of Frame::ExtractLSD ("next" row N2; see psl_oracle.h).
// Sequential restatement of CPartiallyRecoverConnectivity's constructor (/root/reference/add_src/
// PartiallyRecoverConnectivity.cpp:14-133 with ptsDropInRotatedRect :152-170, isPtInRotatedRect :135-150,
// intersectionOfLines :226-247) and of Frame::convertFansToKeyLines / Frame_shortestDistance (src/Frame.cc:380-472).
// Nothing here is used by the product.
//
// Pinned choices: the cv::MatExpr of ptsDropInRotatedRect evaluates as ONE cv::addWeighted in double with a
// single rounding (cv2 4.13; H4/H8); sinf / cosf / tanf are the correctly rounded values (H2); cv::determinant of a 2x2
// CV_32F works in double; Frame_shortestDistance's 2x2 solve follows Eigen 3.3's ColPivHouseholderQR step by step
// (unpinned: Eigen is not available here) and its missing `return` (UB) reads "no point".
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <vector>

#include "psl_oracle.h"

namespace {

const double kPi = 3.1415926535897932384626433832795;  // CV_PI

struct Rect { float cx, cy, hw, hh, dsin, dcos, arc; };

Rect make_rect(const float* p, float radius) {
  Rect R;
  R.cx = (p[0] + p[2]) / 2;
  R.cy = (p[1] + p[3]) / 2;
  const float dy = p[3] - p[1], dx = p[2] - p[0];
  const float degAng = orc_fast_atan2(dy, dx);
  const float arcAng = (float)(degAng / 180 * kPi);
  const float length = std::fabs((float)std::tan((double)arcAng)) > 1 ? std::fabs(dy) : std::fabs(dx);
  const int height = (int)(radius * 2);              // CvSize: int
  const int width = (int)(length + 2 * radius);
  R.hw = (float)width / 2;
  R.hh = (float)height / 2;
  const float angle = (float)(degAng * kPi / 180);
  R.dsin = (float)std::sin((double)angle);
  R.dcos = (float)std::cos((double)angle);
  R.arc = arcAng;
  return R;
}

bool pt_in_rect(float x, float y, const Rect& R) {  // isPtInRotatedRect: scalar float arithmetic
  const float fposx = R.dcos * (x - R.cx) + R.dsin * (y - R.cy);
  const float fposy = R.dsin * (x - R.cx) - R.dcos * (y - R.cy);
  return -R.hw <= fposx && fposx < R.hw && -R.hh <= fposy && fposy < R.hh;
}

double det2(float a, float b, float c, float d) { return (double)a * d - (double)b * c; }  // cv::determinant, 2x2 CV_32F

void intersection(const float* p, const float* q, float& X, float& Y) {
  const float A1 = p[1] - p[3], B1 = p[2] - p[0], C1 = p[3] * p[0] - p[1] * p[2];
  const float A2 = q[1] - q[3], B2 = q[2] - q[0], C2 = q[3] * q[0] - q[1] * q[2];
  const float D = (float)det2(A1, B1, A2, B2);
  X = (float)(det2(-C1, B1, -C2, B2) / D);
  Y = (float)(det2(A1, -C1, A2, -C2) / D);
}

double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// x = A.colPivHouseholderQr().solve(b) for a 2x2 A (row-major a[4]); Eigen 3.3 ColPivHouseholderQR::computeInPlace +
// _solve_impl
void qr_solve2(const double* a, const double* b, double* x) {
  double m[2][2] = {{a[0], a[1]}, {a[2], a[3]}};  // m[row][col]
  int perm[2] = {0, 1};
  double dir[2], upd[2];
  for (int k = 0; k < 2; ++k) dir[k] = upd[k] = std::sqrt(m[0][k] * m[0][k] + m[1][k] * m[1][k]);
  const double eps = DBL_EPSILON;
  const double mx = upd[0] >= upd[1] ? upd[0] : upd[1];
  const double thr_helper = (mx * eps) * (mx * eps) / 2.0;
  const double downdate_thr = std::sqrt(eps);
  int nonzero = 2;
  // k = 0
  int idx = upd[1] > upd[0] ? 1 : 0;
  if (upd[idx] * upd[idx] < thr_helper * 2.0) nonzero = 0;
  if (idx != 0) {
    for (int r = 0; r < 2; ++r) { const double t = m[r][0]; m[r][0] = m[r][1]; m[r][1] = t; }
    double t = upd[0]; upd[0] = upd[1]; upd[1] = t;
    t = dir[0]; dir[0] = dir[1]; dir[1] = t;
    perm[0] = 1; perm[1] = 0;
  }
  double tau0, ess0, beta;
  {
    const double c0 = m[0][0], tailSq = m[1][0] * m[1][0];
    if (tailSq <= DBL_MIN) { tau0 = 0; beta = c0; ess0 = 0; }
    else {
      beta = std::sqrt(c0 * c0 + tailSq);
      if (c0 >= 0) beta = -beta;
      ess0 = m[1][0] / (c0 - beta);
      tau0 = (beta - c0) / beta;
    }
    m[0][0] = beta;
    if (tau0 != 0) {
      double tmp = ess0 * m[1][1];
      tmp += m[0][1];
      m[0][1] -= tau0 * tmp;
      m[1][1] -= (tau0 * ess0) * tmp;
    }
    if (upd[1] != 0) {
      double t = std::fabs(m[0][1]) / upd[1];
      t = (1.0 + t) * (1.0 - t);
      t = t < 0 ? 0 : t;
      const double r = upd[1] / dir[1];
      const double t2 = t * (r * r);
      if (t2 <= downdate_thr) { dir[1] = std::fabs(m[1][1]); upd[1] = dir[1]; }
      else upd[1] *= std::sqrt(t);
    }
  }
  // k = 1: a 1-vector, tau = 0, beta = m[1][1]
  if (nonzero == 2 && upd[1] * upd[1] < thr_helper * 1.0) nonzero = 1;
  x[0] = x[1] = 0;
  if (nonzero == 0) return;
  double c[2] = {b[0], b[1]};
  if (tau0 != 0) {  // Q^T b
    double tmp = ess0 * c[1];
    tmp += c[0];
    c[0] -= tau0 * tmp;
    c[1] -= (tau0 * ess0) * tmp;
  }
  if (nonzero == 2) {
    c[1] = c[1] / m[1][1];
    c[0] = (c[0] - m[0][1] * c[1]) / m[0][0];
    x[perm[0]] = c[0];
    x[perm[1]] = c[1];
  } else {
    x[perm[0]] = c[0] / m[0][0];
  }
}

// Frame::Frame_shortestDistance, Frame.cc:380-424
bool shortest_distance(const double* L1, const double* L2, double* cross) {
  double d1[3], d2[3], w[3];
  for (int k = 0; k < 3; ++k) { d1[k] = L1[3 + k] - L1[k]; d2[k] = L2[3 + k] - L2[k]; w[k] = L1[k] - L2[k]; }
  const double d11 = dot3(d1, d1), d12 = dot3(d1, d2), d22 = dot3(d2, d2), w1 = dot3(w, d1), w2 = dot3(w, d2);
  const double A[4] = {d11, -d12, d12, -d22}, b[2] = {-w1, -w2};
  if (A[0] * A[3] - A[1] * A[2] == 0) return false;
  double x[2];
  qr_solve2(A, b, x);
  double mm[3];
  for (int k = 0; k < 3; ++k) {
    const double r1 = L1[k] + x[0] * d1[k], r2 = L2[k] + x[1] * d2[k];
    cross[k] = (r1 + r2) * 0.5;
    mm[k] = (L1[k] + L2[k]) * 0.5 - (L1[3 + k] + L2[3 + k]) * 0.5;
  }
  const double distmid = std::sqrt(dot3(mm, mm)) * 2;
  double n1 = 0, n2 = 0;
  for (int k = 0; k < 6; ++k) { n1 += L1[k] * L1[k]; n2 += L2[k] * L2[k]; }
  return distmid < std::sqrt(n1) + std::sqrt(n2);  // otherwise the reference returns nothing (UB): no point
}

}  // namespace

extern "C" {

// fans [cap*4] = (x, y, i, j) rows after the duplicate removal; junctions (may be null, needs lines3d) = the entries of
// Frame::intersection_lines_plane.  The counts may exceed cap (only the first cap rows are stored).
int orc_line_junctions(const psl_keyline* kl_un, const double* lines3d, int n, int img_w, int img_h, float radius,
                       float fan_thr, float* fans, psl_line_junction* junctions, int cap, int32_t* n_fans,
                       int32_t* n_junctions) {
  std::vector<float> L((size_t)n * 4);
  for (int i = 0; i < n; ++i) {  // Frame::keyLinesToMat, Frame.cc:355-374
    L[4 * i] = kl_un[i].start_x; L[4 * i + 1] = kl_un[i].start_y;
    L[4 * i + 2] = kl_un[i].end_x; L[4 * i + 3] = kl_un[i].end_y;
  }
  std::vector<Rect> R((size_t)n);
  for (int i = 0; i < n; ++i) R[i] = make_rect(&L[4 * i], radius);
  struct Fan { float x, y; int i, j; };
  std::vector<Fan> raw;
  for (int i = 0; i < n; ++i) {
    const Rect& r = R[i];
    const double a = r.dcos, b = r.dsin;
    const double gx = (-(double)r.cx) * a + (-(double)r.cy) * b, gy = (-(double)r.cx) * b - (-(double)r.cy) * a;
    for (int j = 0; j < 2 * n; ++j) {  // mPts: the start points, then the end points (:24-25)
      const int cur = j >= n ? j - n : j;
      const double px = j >= n ? L[4 * cur + 2] : L[4 * cur], py = j >= n ? L[4 * cur + 3] : L[4 * cur + 1];
      const float fposx = (float)(px * a + py * b + gx);      // addWeighted(X, dcos, Y, dsin, gamma)
      const float fposy = (float)(px * b + py * (-a) + gy);   // addWeighted(X, dsin, Y, -dcos, gamma)
      if (!(-r.hw <= fposx && fposx < r.hw && -r.hh <= fposy && fposy < r.hh)) continue;
      if (cur == i) continue;
      const float pi_f = (float)kPi;
      const float tmpa = std::fmod(std::fabs(r.arc - R[cur].arc), pi_f);
      if (tmpa < fan_thr || kPi - tmpa < fan_thr) continue;
      float X, Y;
      intersection(&L[4 * i], &L[4 * cur], X, Y);
      if (pt_in_rect(X, Y, r) && (X >= 4 && X < img_w - 4 && Y >= 4 && Y < img_h - 4)) raw.push_back({X, Y, i, cur});
    }
  }
  int nf = 0, nj = 0;
  for (size_t i = 0; i < raw.size(); ++i) {  // :111-132: a row survives unless a later row joins the same two lines
    bool flag = true;
    for (size_t j = i + 1; j < raw.size(); ++j)
      if ((raw[i].i == raw[j].i && raw[i].j == raw[j].j) || (raw[i].i == raw[j].j && raw[i].j == raw[j].i)) { flag = false; break; }
    if (!flag) continue;
    if (nf < cap) { fans[4 * nf] = raw[i].x; fans[4 * nf + 1] = raw[i].y; fans[4 * nf + 2] = (float)raw[i].i; fans[4 * nf + 3] = (float)raw[i].j; }
    ++nf;
    if (junctions && lines3d) {  // Frame::convertFansToKeyLines, Frame.cc:426-472
      double cross[3];
      if (shortest_distance(lines3d + 6 * raw[i].i, lines3d + 6 * raw[i].j, cross) &&
          std::sqrt(dot3(cross, cross)) > DBL_EPSILON) {
        if (nj < cap) {
          psl_line_junction& J = junctions[nj];
          J.l1 = raw[i].i; J.l2 = raw[i].j; J.cross2d_x = raw[i].x; J.cross2d_y = raw[i].y;
          J.cross3d[0] = cross[0]; J.cross3d[1] = cross[1]; J.cross3d[2] = cross[2];
        }
        ++nj;
      }
    }
  }
  *n_fans = nf;
  *n_junctions = nj;
  return 0;
}

}  // extern "C"
