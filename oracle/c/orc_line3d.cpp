// TEST INFRASTRUCTURE — CPU oracle for the 3-D line fit of every KeyLine, Frame::isLineGood (src/Frame.cc:662-750)
// with LINEextractor::compPt3dCov, extract3dline_mahdist, verify3dLine, computeLine3d_svd, mah_dist3d_pt_line
// (add_src/LineExtractor.cpp:27-323).  The two cv::SVD calls are restated with OpenCV's one-sided Jacobi
// (JacobiSVDImpl_, modules/core/src/lapack.cpp; same sweep order, rotation formulas, descending sort and therefore
// the same signs), pinned against the real cv2.SVDecomp through tests/golden/lines3d_*.npz.
// H6: rand() of random_unique is the ANSI C example generator, re-seeded per line (see oracle/pyref/line3d_py.py).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "psl_oracle.h"

namespace {

struct P3 { double x, y, z; };
inline P3 operator-(P3 a, P3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline P3 operator+(P3 a, P3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline P3 operator*(P3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline double dot(P3 a, P3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline double norm(P3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }

// One-sided Jacobi on the n rows (length m) of At; Vt (n x n) accumulates the rotations.  On return the rows are
// sorted by singular value, descending; At is NOT normalised (callers that need U divide by W).
void jacobi_svd(double* At, int astep, double* W, double* Vt, int m, int n) {
  const double eps = 2.220446049250313e-16 * 10;
  for (int i = 0; i < n; ++i) {
    double sd = 0;
    for (int k = 0; k < m; ++k) sd += At[i * astep + k] * At[i * astep + k];
    W[i] = sd;
    for (int k = 0; k < n; ++k) Vt[i * n + k] = 0;
    Vt[i * n + i] = 1;
  }
  const int max_iter = std::max(m, 30);
  for (int iter = 0; iter < max_iter; ++iter) {
    bool changed = false;
    for (int i = 0; i < n - 1; ++i)
      for (int j = i + 1; j < n; ++j) {
        double* Ai = At + i * astep;
        double* Aj = At + j * astep;
        double a = W[i], p = 0, b = W[j];
        for (int k = 0; k < m; ++k) p += Ai[k] * Aj[k];
        if (std::abs(p) <= eps * std::sqrt(a * b)) continue;
        p *= 2;
        const double beta = a - b, gamma = std::hypot(p, beta);
        double c, s;
        if (beta < 0) {
          const double delta = (gamma - beta) * 0.5;
          s = std::sqrt(delta / gamma);
          c = p / (gamma * s * 2);
        } else {
          c = std::sqrt((gamma + beta) / (gamma * 2));
          s = p / (gamma * c * 2);
        }
        a = b = 0;
        for (int k = 0; k < m; ++k) {
          const double t0 = c * Ai[k] + s * Aj[k], t1 = -s * Ai[k] + c * Aj[k];
          Ai[k] = t0;
          Aj[k] = t1;
          a += t0 * t0;
          b += t1 * t1;
        }
        W[i] = a;
        W[j] = b;
        changed = true;
        double* Vi = Vt + i * n;
        double* Vj = Vt + j * n;
        for (int k = 0; k < n; ++k) {
          const double t0 = c * Vi[k] + s * Vj[k], t1 = -s * Vi[k] + c * Vj[k];
          Vi[k] = t0;
          Vj[k] = t1;
        }
      }
    if (!changed) break;
  }
  for (int i = 0; i < n; ++i) {
    double sd = 0;
    for (int k = 0; k < m; ++k) sd += At[i * astep + k] * At[i * astep + k];
    W[i] = std::sqrt(sd);
  }
  for (int i = 0; i < n - 1; ++i) {
    int j = i;
    for (int k = i + 1; k < n; ++k)
      if (W[j] < W[k]) j = k;
    if (i != j) {
      std::swap(W[i], W[j]);
      for (int k = 0; k < m; ++k) std::swap(At[i * astep + k], At[j * astep + k]);
      for (int k = 0; k < n; ++k) std::swap(Vt[i * n + k], Vt[j * n + k]);
    }
  }
}

struct RPt { P3 pos; double DU[9]; };

inline double depth_std(double d) { return 0.00273 * d * d + 0.00074 * d + -0.00058; }

void mat3(const double a[3][3], const double b[3][3], double c[3][3]) {  // cv::gemm, k in order
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c[i][j] = a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j];
}

// compPt3dCov, LineExtractor.cpp:40-95: cov0 = J0 cov_g J0^T, cv::SVD, DU = diag(1/sqrt(w)) U^T
RPt comp_pt3d_cov(P3 pt, double f) {
  RPt rp;
  rp.pos = pt;
  const double J[3][3] = {{pt.z / f, 0, pt.x / pt.z}, {0, pt.z / f, pt.y / pt.z}, {0, 0, 1}};
  const double sd = depth_std(pt.z);
  const double G[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, sd * sd}};
  double Jt[3][3], JG[3][3], cov[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Jt[i][j] = J[j][i];
  mat3(J, G, JG);
  mat3(JG, Jt, cov);
  // cv::SVD(cov): m = n = 3, temp_a = cov^T, u = transpose(rows of temp_a / w)
  double At[9], W[3], Vt[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) At[i * 3 + j] = cov[j][i];
  jacobi_svd(At, 3, W, Vt, 3, 3);
  double U[3][3];  // U[r][c] = u.at(r, c) = normalised At row c, element r
  for (int c = 0; c < 3; ++c) {
    const double s = W[c] > 2.2250738585072014e-308 ? 1 / W[c] : 0.;
    for (int r = 0; r < 3; ++r) U[r][c] = At[c * 3 + r] * s;
  }
  double ws[3];
  for (int i = 0; i < 3; ++i) ws[i] = std::sqrt(W[i]);
  const double D[3][3] = {{1 / ws[0], 0, 0}, {0, 1 / ws[1], 0}, {0, 0, 1 / ws[2]}};
  double Ut[3][3], du[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Ut[i][j] = U[j][i];
  mat3(D, Ut, du);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) rp.DU[i * 3 + j] = du[i][j];
  return rp;
}

// mah_dist3d_pt_line, LineExtractor.cpp:174-204 (the common sub-expressions of term1..term3 are the same products)
double mah_dist(const RPt& pt, P3 q1, P3 q2) {
  const double xa = q1.x, ya = q1.y, za = q1.z, xb = q2.x, yb = q2.y, zb = q2.z;
  const double *c = pt.DU, x1 = pt.pos.x, x2 = pt.pos.y, x3 = pt.pos.z;
  const double A1 = c[0] * (x1 - xa) + c[1] * (x2 - ya) + c[2] * (x3 - za);
  const double A2 = c[3] * (x1 - xa) + c[4] * (x2 - ya) + c[5] * (x3 - za);
  const double A3 = c[6] * (x1 - xa) + c[7] * (x2 - ya) + c[8] * (x3 - za);
  const double B1 = c[0] * (x1 - xb) + c[1] * (x2 - yb) + c[2] * (x3 - zb);
  const double B2 = c[3] * (x1 - xb) + c[4] * (x2 - yb) + c[5] * (x3 - zb);
  const double B3 = c[6] * (x1 - xb) + c[7] * (x2 - yb) + c[8] * (x3 - zb);
  const double t1 = A1 * B2 - A2 * B1, t2 = A1 * B3 - A3 * B1, t3 = A2 * B3 - A3 * B2;
  const double t4 = c[0] * (x1 - xa) - c[0] * (x1 - xb) + c[1] * (x2 - ya) - c[1] * (x2 - yb) + c[2] * (x3 - za) - c[2] * (x3 - zb);
  const double t5 = c[3] * (x1 - xa) - c[3] * (x1 - xb) + c[4] * (x2 - ya) - c[4] * (x2 - yb) + c[5] * (x3 - za) - c[5] * (x3 - zb);
  const double t6 = c[6] * (x1 - xa) - c[6] * (x1 - xb) + c[7] * (x2 - ya) - c[7] * (x2 - yb) + c[8] * (x3 - za) - c[8] * (x3 - zb);
  return std::sqrt((t1 * t1 + t2 * t2 + t3 * t3) / (t4 * t4 + t5 * t5 + t6 * t6));
}

P3 proj_pt(P3 P, P3 mid, P3 drct) {  // projPt3d2Ln3d, LineExtractor.h:199-207
  const P3 A = mid, B = mid + drct, AB = B - A, AP = P - A;
  return A + AB * (dot(AB, AP) / dot(AB, AB));
}

bool verify3d(const std::vector<RPt>& pts, const std::vector<int>& set, P3 A, P3 B) {  // LineExtractor.cpp:97-160
  const int nCells = 10;
  int cells[10] = {0};
  double minv = 100, maxv = -100;
  int idx1 = 0, idx2 = 0;
  const int nPts = (int)set.size();
  for (int i = 0; i < nPts; ++i) {
    const double v = dot(pts[set[i]].pos - A, B - A);
    if (v < minv) { minv = v; idx1 = i; }
    if (v > maxv) { maxv = v; idx2 = i; }
  }
  const P3 C = proj_pt(pts[set[idx1]].pos, (A + B) * 0.5, B - A), D = proj_pt(pts[set[idx2]].pos, (A + B) * 0.5, B - A);
  const double cd = norm(D - C);
  if (cd < 0.0000000001) return false;
  for (int i = 0; i < nPts; ++i) {
    const double lambda = std::abs(dot(pts[set[i]].pos - C, D - C) / cd / cd);
    if (lambda >= 1) cells[nCells - 1] += 1;
    else cells[(unsigned)std::floor(lambda * 10)] += 1;
  }
  double sum = 0;
  for (int i = 0; i < nCells; ++i)
    if (cells[i] > 0) sum = sum + 1;
  return sum / nCells > 0.7;
}

// computeLine3d_svd, LineExtractor.cpp:162-181: mean and vt.row(0) of cv::SVD(P.t(), MODIFY_A)
void line3d_svd(const std::vector<RPt>& pts, const std::vector<int>& idx, P3& mean, P3& drct) {
  const int n = (int)idx.size();
  mean = {0, 0, 0};
  for (int i = 0; i < n; ++i) mean = mean + pts[idx[i]].pos;
  mean = mean * (1.0 / n);
  std::vector<double> Pt((size_t)n * 3);  // P.t(): n x 3
  for (int i = 0; i < n; ++i) {
    Pt[3 * i] = pts[idx[i]].pos.x - mean.x;
    Pt[3 * i + 1] = pts[idx[i]].pos.y - mean.y;
    Pt[3 * i + 2] = pts[idx[i]].pos.z - mean.z;
  }
  double W[3], Vt[9];
  if (n >= 3) {  // m = n_pts >= n = 3: temp_a = src^T (3 rows of length n), vt = accumulated rotations
    std::vector<double> At((size_t)3 * n);
    for (int i = 0; i < n; ++i)
      for (int c = 0; c < 3; ++c) At[(size_t)c * n + i] = Pt[3 * i + c];
    jacobi_svd(At.data(), n, W, Vt, n, 3);
    drct = {Vt[0], Vt[1], Vt[2]};
  } else {  // m < n: the roles swap (at = true): temp_a = src (n_pts rows of length 3), vt = its normalised rows
    std::vector<double> At(Pt);
    double V2[4];
    jacobi_svd(At.data(), 3, W, V2, 3, n);
    const double s = W[0] > 2.2250738585072014e-308 ? 1 / W[0] : 0.;
    drct = {At[0] * s, At[1] * s, At[2] * s};
  }
}

struct Rand {
  uint32_t s;
  int operator()() { s = s * 1103515245u + 12345u; return (int)((s >> 16) & 0x7FFFu); }
};

// extract3dline_mahdist, LineExtractor.cpp:206-323
void extract3dline(const std::vector<RPt>& pts, Rand rnd, P3& A_out, P3& B_out) {
  const int n = (int)pts.size();
  const int maxIterNo = std::min(10, int(n * (n - 1) * 0.5));
  const double distThresh = 3.0;
  std::vector<int> indexes(n), maxInlierSet;
  for (int i = 0; i < n; ++i) indexes[i] = i;
  P3 bestA{0, 0, 0}, bestB{0, 0, 0};
  for (int iter = 0; iter < maxIterNo; ++iter) {
    std::vector<int> inlierSet;
    int left = n;
    for (int b = 0; b < 2; ++b) {  // random_unique(begin, end, 2), LineExtractor.h:25-37
      std::swap(indexes[b], indexes[b + rnd() % left]);
      --left;
    }
    const P3 A = pts[indexes[0]].pos, B = pts[indexes[1]].pos;
    if (norm(B - A) < 0.0000000001) continue;
    for (int i = 0; i < n; ++i)
      if (mah_dist(pts[i], A, B) < distThresh) inlierSet.push_back(i);
    if (inlierSet.size() > maxInlierSet.size() && verify3d(pts, inlierSet, A, B)) {
      maxInlierSet = inlierSet;
      bestA = A;
      bestB = B;
    }
    if (maxInlierSet.size() > n * 0.6) break;
  }
  A_out = {0, 0, 0};
  B_out = {0, 0, 0};
  if (maxInlierSet.size() >= 2) {
    P3 m = (bestA + bestB) * 0.5, d = bestB - bestA;
    while (true) {
      std::vector<int> tmp;
      P3 tm, td;
      line3d_svd(pts, maxInlierSet, tm, td);
      for (int i = 0; i < n; ++i)
        if (mah_dist(pts[i], tm, tm + td) < distThresh) tmp.push_back(i);
      if (tmp.size() > maxInlierSet.size()) {
        maxInlierSet = tmp;
        m = tm;
        d = td;
      } else {
        break;
      }
    }
    double minv = 100, maxv = -100;
    int e1 = 0, e2 = 0;
    for (size_t i = 0; i < maxInlierSet.size(); ++i) {
      const double dp = dot(pts[maxInlierSet[i]].pos - m, d);
      if (dp < minv) { minv = dp; e1 = (int)i; }
      if (dp > maxv) { maxv = dp; e2 = (int)i; }
    }
    A_out = pts[maxInlierSet[e1]].pos;
    B_out = pts[maxInlierSet[e2]].pos;
  }
}

}  // namespace

extern "C" {

// Frame::isLineGood, Frame.cc:662-750.  depth: CV_32F metres [h][stride]; cam: fx, fy, cx, cy (Frame's statics and
// mK(0,0)); lines3d [n*6] = mvLines3D (first, second), line_eq [n*3] = mvLineEq.
void orc_lines_3d(const psl_keyline* kl, int n, const float* depth, int w, int h, int stride, const float* cam,
                  uint32_t seed, double* lines3d, float* line_eq) {
  const float fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3];
  const float invfx = 1.0f / fx, invfy = 1.0f / fy;
  for (int i = 0; i < n; ++i) {
    for (int k = 0; k < 6; ++k) lines3d[6 * i + k] = 0.0;
    for (int k = 0; k < 3; ++k) line_eq[3 * i + k] = -1.0f;
    const float sx = kl[i].start_x, sy = kl[i].start_y, ex = kl[i].end_x, ey = kl[i].end_y;
    const float ddx = sx - ex, ddy = sy - ey;
    const double len = std::sqrt((double)ddx * ddx + (double)ddy * ddy);  // cv::norm(Point2f)
    const double numSmp = (double)std::min((int)len, 20);
    if (numSmp < 1) continue;  // 0 / 0 sample positions in the reference: undefined, no line
    std::vector<RPt> pts;
    for (int j = 0; j <= numSmp; ++j) {
      // Point2f * double -> Point2f (saturate_cast<float> of the double product), Point2f + Point2f, -> Point2d
      const double a = 1 - j / numSmp, b = j / numSmp;
      const double px = (double)((float)(sx * a) + (float)(ex * b)), py = (double)((float)(sy * a) + (float)(ey * b));
      if (px < 0 || py < 0 || px >= w || py >= h) continue;
      int row, col;
      if (std::floor(px) == px && std::floor(py) == py) {
        col = std::max(int(px - 1), 0);
        row = std::max(int(py - 1), 0);
      } else {
        col = int(px);
        row = int(py);
      }
      const float d = depth[(size_t)row * stride + col];
      if (d <= 0.01) continue;
      P3 p;
      p.z = d;
      p.x = (col - cx) * p.z * invfx;
      p.y = (row - cy) * p.z * invfy;
      pts.push_back(comp_pt3d_cov(p, (double)fx));
    }
    if (pts.size() < 5) continue;
    P3 A, B;
    extract3dline(pts, Rand{seed * 1000003u + (uint32_t)i + 1u}, A, B);
    if (norm(A - B) > 0.02) {
      lines3d[6 * i] = A.x; lines3d[6 * i + 1] = A.y; lines3d[6 * i + 2] = A.z;
      lines3d[6 * i + 3] = B.x; lines3d[6 * i + 4] = B.y; lines3d[6 * i + 5] = B.z;
      const float l0 = (float)(B.x - A.x), l1 = (float)(B.y - A.y), l2 = (float)(B.z - A.z);
      const float magn = sqrtf(l0 * l0 + l1 * l1 + l2 * l2);
      line_eq[3 * i] = l0 / magn; line_eq[3 * i + 1] = l1 / magn; line_eq[3 * i + 2] = l2 / magn;
    }
  }
}

}  // extern "C"
