// 3-D line of every KeyLine from the depth image: Frame::isLineGood (src/Frame.cc:662-750) with
// LINEextractor::compPt3dCov, extract3dline_mahdist, verify3dLine, computeLine3d_svd and mah_dist3d_pt_line
// (add_src/LineExtractor.cpp:27-323).  SURVEY "next" row N2.
//
// One thread per line: at most 21 depth samples, their covariances whitened through the SVD of a 3x3 matrix, a RANSAC
// of at most 10 draws under the Mahalanobis distance, an SVD refit loop, and the two extreme inliers as end points.
// Everything is fp64 as in the reference, with explicit round-to-nearest operations (no contraction).  The two
// cv::SVD calls are OpenCV's one-sided Jacobi (JacobiSVDImpl_): same sweep order, rotations and descending sort,
// hence the same singular-vector signs (the sign of vt.row(0) decides which end point is `first`).
// Pinned choice H6: rand() of random_unique is the ANSI C example generator, re-seeded per line.
#include "line_match_kernels.cuh"

namespace psl {
namespace {

constexpr int kMaxSmp = 21;  // min((int)len, 20) + 1 samples

struct P3 { double x, y, z; };
__device__ __forceinline__ P3 sub3(P3 a, P3 b) { return {__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y), __dsub_rn(a.z, b.z)}; }
__device__ __forceinline__ P3 add3(P3 a, P3 b) { return {__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y), __dadd_rn(a.z, b.z)}; }
__device__ __forceinline__ P3 mul3(P3 a, double s) { return {__dmul_rn(a.x, s), __dmul_rn(a.y, s), __dmul_rn(a.z, s)}; }
__device__ __forceinline__ double dot3(P3 a, P3 b) {
  return __dadd_rn(__dadd_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)), __dmul_rn(a.z, b.z));
}
__device__ __forceinline__ double norm3(P3 a) { return __dsqrt_rn(dot3(a, a)); }

// One-sided Jacobi on the n rows (length m, row stride astep) of At; Vt (n x n) accumulates the rotations; rows come
// back sorted by singular value, descending; At is not normalised.
__device__ void jacobi_svd(double* At, int astep, double* W, double* Vt, int m, int n) {
  const double eps = 2.220446049250313e-16 * 10;
  for (int i = 0; i < n; ++i) {
    double sd = 0;
    for (int k = 0; k < m; ++k) sd = __dadd_rn(sd, __dmul_rn(At[i * astep + k], At[i * astep + k]));
    W[i] = sd;
    for (int k = 0; k < n; ++k) Vt[i * n + k] = 0;
    Vt[i * n + i] = 1;
  }
  const int max_iter = max(m, 30);
  for (int iter = 0; iter < max_iter; ++iter) {
    bool changed = false;
    for (int i = 0; i < n - 1; ++i)
      for (int j = i + 1; j < n; ++j) {
        double* Ai = At + i * astep;
        double* Aj = At + j * astep;
        double a = W[i], p = 0, b = W[j];
        for (int k = 0; k < m; ++k) p = __dadd_rn(p, __dmul_rn(Ai[k], Aj[k]));
        if (fabs(p) <= __dmul_rn(eps, __dsqrt_rn(__dmul_rn(a, b)))) continue;
        p = __dmul_rn(p, 2.0);
        const double beta = __dsub_rn(a, b), gamma = hypot(p, beta);
        double c, s;
        if (beta < 0) {
          const double delta = __dmul_rn(__dsub_rn(gamma, beta), 0.5);
          s = __dsqrt_rn(__ddiv_rn(delta, gamma));
          c = __ddiv_rn(p, __dmul_rn(__dmul_rn(gamma, s), 2.0));
        } else {
          c = __dsqrt_rn(__ddiv_rn(__dadd_rn(gamma, beta), __dmul_rn(gamma, 2.0)));
          s = __ddiv_rn(p, __dmul_rn(__dmul_rn(gamma, c), 2.0));
        }
        a = b = 0;
        for (int k = 0; k < m; ++k) {
          const double t0 = __dadd_rn(__dmul_rn(c, Ai[k]), __dmul_rn(s, Aj[k]));
          const double t1 = __dadd_rn(__dmul_rn(-s, Ai[k]), __dmul_rn(c, Aj[k]));
          Ai[k] = t0;
          Aj[k] = t1;
          a = __dadd_rn(a, __dmul_rn(t0, t0));
          b = __dadd_rn(b, __dmul_rn(t1, t1));
        }
        W[i] = a;
        W[j] = b;
        changed = true;
        double* Vi = Vt + i * n;
        double* Vj = Vt + j * n;
        for (int k = 0; k < n; ++k) {
          const double t0 = __dadd_rn(__dmul_rn(c, Vi[k]), __dmul_rn(s, Vj[k]));
          const double t1 = __dadd_rn(__dmul_rn(-s, Vi[k]), __dmul_rn(c, Vj[k]));
          Vi[k] = t0;
          Vj[k] = t1;
        }
      }
    if (!changed) break;
  }
  for (int i = 0; i < n; ++i) {
    double sd = 0;
    for (int k = 0; k < m; ++k) sd = __dadd_rn(sd, __dmul_rn(At[i * astep + k], At[i * astep + k]));
    W[i] = __dsqrt_rn(sd);
  }
  for (int i = 0; i < n - 1; ++i) {
    int j = i;
    for (int k = i + 1; k < n; ++k)
      if (W[j] < W[k]) j = k;
    if (i != j) {
      double t = W[i]; W[i] = W[j]; W[j] = t;
      for (int k = 0; k < m; ++k) { t = At[i * astep + k]; At[i * astep + k] = At[j * astep + k]; At[j * astep + k] = t; }
      for (int k = 0; k < n; ++k) { t = Vt[i * n + k]; Vt[i * n + k] = Vt[j * n + k]; Vt[j * n + k] = t; }
    }
  }
}

__device__ __forceinline__ void mat3(const double a[3][3], const double b[3][3], double c[3][3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      c[i][j] = __dadd_rn(__dadd_rn(__dmul_rn(a[i][0], b[0][j]), __dmul_rn(a[i][1], b[1][j])), __dmul_rn(a[i][2], b[2][j]));
}

// compPt3dCov (LineExtractor.cpp:40-95): DU = diag(1 / sqrt(w)) U^T of cov0 = J0 diag(1, 1, sigma_z^2) J0^T
__device__ void comp_pt3d_cov(P3 pt, double f, double* DU) {
  const double J[3][3] = {{__ddiv_rn(pt.z, f), 0, __ddiv_rn(pt.x, pt.z)}, {0, __ddiv_rn(pt.z, f), __ddiv_rn(pt.y, pt.z)}, {0, 0, 1}};
  // depthStdDev: c1 * d * d + c2 * d + c3
  const double sd = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(0.00273, pt.z), pt.z), __dmul_rn(0.00074, pt.z)), -0.00058);
  const double G[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, __dmul_rn(sd, sd)}};
  double Jt[3][3], JG[3][3], cov[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Jt[i][j] = J[j][i];
  mat3(J, G, JG);
  mat3(JG, Jt, cov);
  double At[9], W[3], Vt[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) At[i * 3 + j] = cov[j][i];
  jacobi_svd(At, 3, W, Vt, 3, 3);
  double Ut[3][3];  // Ut[c][r] = u.at(r, c) = normalised At row c, element r
  for (int c = 0; c < 3; ++c) {
    const double s = W[c] > 2.2250738585072014e-308 ? __ddiv_rn(1.0, W[c]) : 0.;
    for (int r = 0; r < 3; ++r) Ut[c][r] = __dmul_rn(At[c * 3 + r], s);
  }
  const double D[3][3] = {{__ddiv_rn(1.0, __dsqrt_rn(W[0])), 0, 0}, {0, __ddiv_rn(1.0, __dsqrt_rn(W[1])), 0},
                          {0, 0, __ddiv_rn(1.0, __dsqrt_rn(W[2]))}};
  double du[3][3];
  mat3(D, Ut, du);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) DU[i * 3 + j] = du[i][j];
}

// mah_dist3d_pt_line (LineExtractor.cpp:174-204)
__device__ double mah_dist(P3 pos, const double* c, P3 q1, P3 q2) {
  const double ax = __dsub_rn(pos.x, q1.x), ay = __dsub_rn(pos.y, q1.y), az = __dsub_rn(pos.z, q1.z);
  const double bx = __dsub_rn(pos.x, q2.x), by = __dsub_rn(pos.y, q2.y), bz = __dsub_rn(pos.z, q2.z);
#define PSL_ROW(r, x, y, z) __dadd_rn(__dadd_rn(__dmul_rn(c[3 * r], x), __dmul_rn(c[3 * r + 1], y)), __dmul_rn(c[3 * r + 2], z))
  const double A1 = PSL_ROW(0, ax, ay, az), A2 = PSL_ROW(1, ax, ay, az), A3 = PSL_ROW(2, ax, ay, az);
  const double B1 = PSL_ROW(0, bx, by, bz), B2 = PSL_ROW(1, bx, by, bz), B3 = PSL_ROW(2, bx, by, bz);
#undef PSL_ROW
  const double t1 = __dsub_rn(__dmul_rn(A1, B2), __dmul_rn(A2, B1));
  const double t2 = __dsub_rn(__dmul_rn(A1, B3), __dmul_rn(A3, B1));
  const double t3 = __dsub_rn(__dmul_rn(A2, B3), __dmul_rn(A3, B2));
  // term4..6: c1*(x1-xa) - c1*(x1-xb) + c2*(x2-ya) - c2*(x2-yb) + c3*(x3-za) - c3*(x3-zb), left to right
#define PSL_T(r)                                                                                                          \
  __dsub_rn(__dadd_rn(__dsub_rn(__dadd_rn(__dsub_rn(__dmul_rn(c[3 * r], ax), __dmul_rn(c[3 * r], bx)), __dmul_rn(c[3 * r + 1], ay)), \
                                __dmul_rn(c[3 * r + 1], by)),                                                             \
                      __dmul_rn(c[3 * r + 2], az)),                                                                       \
            __dmul_rn(c[3 * r + 2], bz))
  const double t4 = PSL_T(0), t5 = PSL_T(1), t6 = PSL_T(2);
#undef PSL_T
  const double num = __dadd_rn(__dadd_rn(__dmul_rn(t1, t1), __dmul_rn(t2, t2)), __dmul_rn(t3, t3));
  const double den = __dadd_rn(__dadd_rn(__dmul_rn(t4, t4), __dmul_rn(t5, t5)), __dmul_rn(t6, t6));
  return __dsqrt_rn(__ddiv_rn(num, den));
}

__device__ __forceinline__ P3 proj_pt(P3 P, P3 mid, P3 drct) {  // projPt3d2Ln3d
  const P3 A = mid, B = add3(mid, drct), AB = sub3(B, A), AP = sub3(P, A);
  return add3(A, mul3(AB, __ddiv_rn(dot3(AB, AP), dot3(AB, AB))));
}

__device__ bool verify3d(const P3* pos, const int* set, int ns, P3 A, P3 B) {  // LineExtractor.cpp:97-160
  int cells[10];
  for (int i = 0; i < 10; ++i) cells[i] = 0;
  double minv = 100, maxv = -100;
  int idx1 = 0, idx2 = 0;
  const P3 BA = sub3(B, A);
  for (int i = 0; i < ns; ++i) {
    const double v = dot3(sub3(pos[set[i]], A), BA);
    if (v < minv) { minv = v; idx1 = i; }
    if (v > maxv) { maxv = v; idx2 = i; }
  }
  const P3 mid = mul3(add3(A, B), 0.5);
  const P3 C = proj_pt(pos[set[idx1]], mid, BA), D = proj_pt(pos[set[idx2]], mid, BA);
  const P3 DC = sub3(D, C);
  const double cd = norm3(DC);
  if (cd < 0.0000000001) return false;
  for (int i = 0; i < ns; ++i) {
    const double lambda = fabs(__ddiv_rn(__ddiv_rn(dot3(sub3(pos[set[i]], C), DC), cd), cd));
    if (lambda >= 1) cells[9] += 1;
    else cells[(unsigned)floor(__dmul_rn(lambda, 10.0))] += 1;
  }
  int sum = 0;
  for (int i = 0; i < 10; ++i)
    if (cells[i] > 0) ++sum;
  return __ddiv_rn((double)sum, 10.0) > 0.7;
}

// computeLine3d_svd (LineExtractor.cpp:162-181): mean and vt.row(0) of cv::SVD(P.t(), MODIFY_A)
__device__ void line3d_svd(const P3* pos, const int* idx, int n, P3& mean, P3& drct) {
  mean = {0, 0, 0};
  for (int i = 0; i < n; ++i) mean = add3(mean, pos[idx[i]]);
  mean = mul3(mean, __ddiv_rn(1.0, (double)n));
  double At[3 * kMaxSmp], W[3], Vt[9];
  if (n >= 3) {  // temp_a = src^T: the x, y, z coordinate rows, each of length n
    for (int i = 0; i < n; ++i) {
      const P3 d = sub3(pos[idx[i]], mean);
      At[i] = d.x; At[n + i] = d.y; At[2 * n + i] = d.z;
    }
    jacobi_svd(At, n, W, Vt, n, 3);
    drct = {Vt[0], Vt[1], Vt[2]};
  } else {  // fewer points than columns: cv::SVD swaps the roles, vt = the normalised rows of the input
    for (int i = 0; i < n; ++i) {
      const P3 d = sub3(pos[idx[i]], mean);
      At[3 * i] = d.x; At[3 * i + 1] = d.y; At[3 * i + 2] = d.z;
    }
    jacobi_svd(At, 3, W, Vt, 3, n);
    const double s = W[0] > 2.2250738585072014e-308 ? __ddiv_rn(1.0, W[0]) : 0.;
    drct = {__dmul_rn(At[0], s), __dmul_rn(At[1], s), __dmul_rn(At[2], s)};
  }
}

__global__ void __launch_bounds__(64)
    lines3d_kernel(const psl_keyline* __restrict__ kl, const int32_t* __restrict__ n_lines, int cap,
                   const float* __restrict__ depth, int w, int h, int stride, int64_t frame_stride, float fx, float fy,
                   float cx, float cy, uint32_t seed, double* __restrict__ lines3d, float* __restrict__ line_eq) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_lines[b]) return;
  const size_t row_out = (size_t)b * cap + i;
  double* L = lines3d + 6 * row_out;
  float* E = line_eq + 3 * row_out;
  for (int k = 0; k < 6; ++k) L[k] = 0.0;
  for (int k = 0; k < 3; ++k) E[k] = -1.0f;
  const psl_keyline K = kl[row_out];
  const float* dimg = depth + (size_t)b * frame_stride;
  const float invfx = __fdiv_rn(1.0f, fx), invfy = __fdiv_rn(1.0f, fy);
  const float ddx = __fsub_rn(K.start_x, K.end_x), ddy = __fsub_rn(K.start_y, K.end_y);
  const double len = __dsqrt_rn(__dadd_rn(__dmul_rn((double)ddx, (double)ddx), __dmul_rn((double)ddy, (double)ddy)));
  const int nsmp = min((int)len, 20);
  if (nsmp < 1) return;  // 0 / 0 sample positions in the reference: undefined, no line
  const double numSmp = (double)nsmp;
  P3 pos[kMaxSmp];
  double DU[kMaxSmp][9];
  int n = 0;
  for (int j = 0; j <= nsmp; ++j) {
    const double bq = __ddiv_rn((double)j, numSmp), aq = __dsub_rn(1.0, bq);
    const double px = (double)__fadd_rn((float)__dmul_rn((double)K.start_x, aq), (float)__dmul_rn((double)K.end_x, bq));
    const double py = (double)__fadd_rn((float)__dmul_rn((double)K.start_y, aq), (float)__dmul_rn((double)K.end_y, bq));
    if (px < 0 || py < 0 || px >= w || py >= h) continue;
    int row, col;
    if (floor(px) == px && floor(py) == py) {
      col = max((int)(px - 1), 0);
      row = max((int)(py - 1), 0);
    } else {
      col = (int)px;
      row = (int)py;
    }
    const float d = dimg[(size_t)row * stride + col];
    if ((double)d <= 0.01) continue;
    P3 p;
    p.z = (double)d;
    p.x = __dmul_rn(__dmul_rn((double)__fsub_rn((float)col, cx), p.z), (double)invfx);
    p.y = __dmul_rn(__dmul_rn((double)__fsub_rn((float)row, cy), p.z), (double)invfy);
    pos[n] = p;
    comp_pt3d_cov(p, (double)fx, DU[n]);
    ++n;
  }
  if (n < 5) return;

  // extract3dline_mahdist (LineExtractor.cpp:206-323)
  uint32_t rs = seed * 1000003u + (uint32_t)i + 1u;
  int indexes[kMaxSmp], best[kMaxSmp], cur[kMaxSmp];
  for (int k = 0; k < n; ++k) indexes[k] = k;
  int nbest = 0;
  P3 bestA{0, 0, 0}, bestB{0, 0, 0};
  const int maxIterNo = min(10, (int)__dmul_rn((double)(n * (n - 1)), 0.5));
  for (int iter = 0; iter < maxIterNo; ++iter) {
    int left = n;
    for (int q = 0; q < 2; ++q) {  // random_unique(begin, end, 2)
      rs = rs * 1103515245u + 12345u;
      const int r = q + (int)((rs >> 16) & 0x7FFFu) % left;
      const int t = indexes[q]; indexes[q] = indexes[r]; indexes[r] = t;
      --left;
    }
    const P3 A = pos[indexes[0]], B = pos[indexes[1]];
    if (norm3(sub3(B, A)) < 0.0000000001) continue;
    int nc = 0;
    for (int k = 0; k < n; ++k)
      if (mah_dist(pos[k], DU[k], A, B) < 3.0) cur[nc++] = k;
    if (nc > nbest && verify3d(pos, cur, nc, A, B)) {
      nbest = nc;
      for (int k = 0; k < nc; ++k) best[k] = cur[k];
      bestA = A;
      bestB = B;
    }
    if ((double)nbest > __dmul_rn((double)n, 0.6)) break;
  }
  if (nbest < 2) return;  // rl.A = rl.B = (0,0,0): norm 0, no line
  P3 m = mul3(add3(bestA, bestB), 0.5), d = sub3(bestB, bestA);
  while (true) {
    P3 tm, td;
    line3d_svd(pos, best, nbest, tm, td);
    int nc = 0;
    const P3 q2 = add3(tm, td);
    for (int k = 0; k < n; ++k)
      if (mah_dist(pos[k], DU[k], tm, q2) < 3.0) cur[nc++] = k;
    if (nc > nbest) {
      nbest = nc;
      for (int k = 0; k < nc; ++k) best[k] = cur[k];
      m = tm;
      d = td;
    } else {
      break;
    }
  }
  double minv = 100, maxv = -100;
  int e1 = 0, e2 = 0;
  for (int k = 0; k < nbest; ++k) {
    const double dp = dot3(sub3(pos[best[k]], m), d);
    if (dp < minv) { minv = dp; e1 = k; }
    if (dp > maxv) { maxv = dp; e2 = k; }
  }
  const P3 A = pos[best[e1]], B = pos[best[e2]];
  if (norm3(sub3(A, B)) > 0.02) {  // Frame.cc:734-748
    L[0] = A.x; L[1] = A.y; L[2] = A.z; L[3] = B.x; L[4] = B.y; L[5] = B.z;
    const float l0 = (float)__dsub_rn(B.x, A.x), l1 = (float)__dsub_rn(B.y, A.y), l2 = (float)__dsub_rn(B.z, A.z);
    const float magn = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(l0, l0), __fmul_rn(l1, l1)), __fmul_rn(l2, l2)));
    E[0] = __fdiv_rn(l0, magn); E[1] = __fdiv_rn(l1, magn); E[2] = __fdiv_rn(l2, magn);
  }
}

}  // namespace

void launch_lines3d(const psl_keyline* kl, const int32_t* n_lines, int cap, const float* depth, int w, int h, int stride,
                    int64_t frame_stride, float fx, float fy, float cx, float cy, uint32_t seed, double* lines3d,
                    float* line_eq, int B, cudaStream_t st) {
  if (B <= 0 || cap <= 0) return;
  dim3 grid((cap + 63) / 64, B);
  lines3d_kernel<<<grid, 64, 0, st>>>(kl, n_lines, cap, depth, w, h, stride, frame_stride, fx, fy, cx, cy, seed, lines3d,
                                      line_eq);
}

}  // namespace psl
