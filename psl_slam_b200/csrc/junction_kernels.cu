// Line-junction detection of Frame::ExtractLSD ("next" row N2): CPartiallyRecoverConnectivity's constructor
// (add_src/PartiallyRecoverConnectivity.cpp:14-133, called at src/Frame.cc:504-505) and the 3-D cross points of
// Frame::convertFansToKeyLines / Frame_shortestDistance (src/Frame.cc:380-472), i.e. the step between psl_lines_3d
// and psl_plane_hypotheses.  One CTA per frame, one thread per line.
//
// The reference appends a "fan" (x, y, i, j) for every end point j of another line that drops into line i's expanded
// rectangle, if the two lines are not nearly parallel and their intersection lies in that rectangle and in the image;
// a second pass deletes every row that a LATER row repeats (same unordered pair).  Whether row (i, j) survives can be
// decided from the geometry alone: the later rows that can repeat it are (i, other end point of the same line) and,
// if that line comes after i, its own two rows towards i.  So every thread counts its surviving rows, a scan gives the
// offsets, and a second pass writes them in the reference's order — no raw list, no sort.
//
// Arithmetic: float statements as written in the reference (nvcc -fmad=false: no contraction); the cv::MatExpr of
// ptsDropInRotatedRect (:154-160) is one cv::addWeighted evaluated in double and rounded once (cv2 4.13);
// cv::determinant of a 2x2 CV_32F in double; Eigen's 2x2 ColPivHouseholderQR solve restated step by step in fp64.
#include <cfloat>
#include <math.h>

#include <algorithm>

#include "lsd_core.cuh"
#include "psl_ctx.cuh"

namespace psl {
namespace {

constexpr double kCvPi = 3.1415926535897932384626433832795;
constexpr int kJThreads = 256;

struct JLine { float x1, y1, x2, y2, cx, cy, hw, hh, dsin, dcos, arc; };

__device__ JLine make_jline(const psl_keyline& k, float radius) {
  JLine R;
  R.x1 = k.start_x; R.y1 = k.start_y; R.x2 = k.end_x; R.y2 = k.end_y;   // Frame::keyLinesToMat, Frame.cc:355-374
  R.cx = (R.x1 + R.x2) / 2;
  R.cy = (R.y1 + R.y2) / 2;
  const float dy = R.y2 - R.y1, dx = R.x2 - R.x1;
  const float degAng = lsd::fast_atan2(dy, dx);
  const float arcAng = (float)((double)(degAng / 180) * kCvPi);
  const float length = fabsf((float)tan((double)arcAng)) > 1 ? fabsf(dy) : fabsf(dx);
  const int height = (int)(radius * 2);                                  // CvSize holds ints
  const int width = (int)(length + 2 * radius);
  R.hw = (float)width / 2;
  R.hh = (float)height / 2;
  const float angle = (float)((double)degAng * kCvPi / 180);
  R.dsin = (float)sin((double)angle);
  R.dcos = (float)cos((double)angle);
  R.arc = arcAng;
  return R;
}

// ptsDropInRotatedRect for one point: the two addWeighted results
__device__ bool drop_in(const JLine& r, float pxf, float pyf) {
  const double a = r.dcos, b = r.dsin, px = pxf, py = pyf;
  const double gx = (-(double)r.cx) * a + (-(double)r.cy) * b, gy = (-(double)r.cx) * b - (-(double)r.cy) * a;
  const float fposx = (float)(px * a + py * b + gx);
  const float fposy = (float)(px * b + py * (-a) + gy);
  return -r.hw <= fposx && fposx < r.hw && -r.hh <= fposy && fposy < r.hh;
}

__device__ bool pt_in_rect(float x, float y, const JLine& R) {            // isPtInRotatedRect
  const float fposx = R.dcos * (x - R.cx) + R.dsin * (y - R.cy);
  const float fposy = R.dsin * (x - R.cx) - R.dcos * (y - R.cy);
  return -R.hw <= fposx && fposx < R.hw && -R.hh <= fposy && fposy < R.hh;
}

__device__ double det2(float a, float b, float c, float d) { return (double)a * d - (double)b * c; }

__device__ void intersect(const JLine& p, const JLine& q, float& X, float& Y) {   // intersectionOfLines
  const float A1 = p.y1 - p.y2, B1 = p.x2 - p.x1, C1 = p.y2 * p.x1 - p.y1 * p.x2;
  const float A2 = q.y1 - q.y2, B2 = q.x2 - q.x1, C2 = q.y2 * q.x1 - q.y1 * q.x2;
  const float D = (float)det2(A1, B1, A2, B2);
  X = (float)(det2(-C1, B1, -C2, B2) / D);
  Y = (float)(det2(A1, -C1, A2, -C2) / D);
}

// does the loop body of :46-108 append a row for (line i, point j)?  j < n: start point of line j, else end point of j-n
__device__ bool accept(const JLine* R, int n, int i, int j, float fan_thr, int img_w, int img_h, float& X, float& Y) {
  const int cur = j >= n ? j - n : j;
  const JLine& r = R[i];
  const JLine& c = R[cur];
  if (!drop_in(r, j >= n ? c.x2 : c.x1, j >= n ? c.y2 : c.y1)) return false;
  if (cur == i) return false;
  const float tmpa = fmodf(fabsf(r.arc - c.arc), (float)kCvPi);
  if (tmpa < fan_thr || kCvPi - (double)tmpa < (double)fan_thr) return false;
  intersect(r, c, X, Y);
  return pt_in_rect(X, Y, r) && X >= 4 && X < (float)(img_w - 4) && Y >= 4 && Y < (float)(img_h - 4);
}

// ... and does that row survive the duplicate removal (:111-132)?
__device__ bool keeps(const JLine* R, int n, int i, int j, float fan_thr, int img_w, int img_h, float& X, float& Y) {
  if (!accept(R, n, i, j, fan_thr, img_w, img_h, X, Y)) return false;
  float tx, ty;
  if (j < n && accept(R, n, i, j + n, fan_thr, img_w, img_h, tx, ty)) return false;   // same pair, later in i's rows
  const int cur = j >= n ? j - n : j;
  if (cur > i && (accept(R, n, cur, i, fan_thr, img_w, img_h, tx, ty) ||
                  accept(R, n, cur, i + n, fan_thr, img_w, img_h, tx, ty)))
    return false;                                                                     // the pair again in cur's rows
  return true;
}

__device__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// x = A.colPivHouseholderQr().solve(b), A 2x2 row-major (Eigen 3.3: computeInPlace + _solve_impl)
__device__ void qr_solve2(const double* a, const double* b, double* x) {
  double m00 = a[0], m01 = a[1], m10 = a[2], m11 = a[3];
  double dir0 = sqrt(m00 * m00 + m10 * m10), dir1 = sqrt(m01 * m01 + m11 * m11);
  double upd0 = dir0, upd1 = dir1;
  const double eps = DBL_EPSILON;
  const double mx = upd0 >= upd1 ? upd0 : upd1;
  const double thr_helper = (mx * eps) * (mx * eps) / 2.0;
  const double downdate_thr = sqrt(eps);
  int nonzero = 2;
  const bool swap = upd1 > upd0;
  if ((swap ? upd1 * upd1 : upd0 * upd0) < thr_helper * 2.0) nonzero = 0;
  if (swap) {
    double t = m00; m00 = m01; m01 = t;
    t = m10; m10 = m11; m11 = t;
    t = upd0; upd0 = upd1; upd1 = t;
    t = dir0; dir0 = dir1; dir1 = t;
  }
  double tau0, ess0, beta;
  const double c0 = m00, tailSq = m10 * m10;
  if (tailSq <= DBL_MIN) { tau0 = 0; beta = c0; ess0 = 0; }
  else {
    beta = sqrt(c0 * c0 + tailSq);
    if (c0 >= 0) beta = -beta;
    ess0 = m10 / (c0 - beta);
    tau0 = (beta - c0) / beta;
  }
  m00 = beta;
  if (tau0 != 0) {
    double tmp = ess0 * m11;
    tmp += m01;
    m01 -= tau0 * tmp;
    m11 -= (tau0 * ess0) * tmp;
  }
  if (upd1 != 0) {
    double t = fabs(m01) / upd1;
    t = (1.0 + t) * (1.0 - t);
    t = t < 0 ? 0 : t;
    const double r = upd1 / dir1;
    const double t2 = t * (r * r);
    if (t2 <= downdate_thr) { dir1 = fabs(m11); upd1 = dir1; }
    else upd1 *= sqrt(t);
  }
  if (nonzero == 2 && upd1 * upd1 < thr_helper * 1.0) nonzero = 1;
  x[0] = x[1] = 0;
  if (nonzero == 0) return;
  double c[2] = {b[0], b[1]};
  if (tau0 != 0) {
    double tmp = ess0 * c[1];
    tmp += c[0];
    c[0] -= tau0 * tmp;
    c[1] -= (tau0 * ess0) * tmp;
  }
  const int p0 = swap ? 1 : 0;
  if (nonzero == 2) {
    c[1] = c[1] / m11;
    c[0] = (c[0] - m01 * c[1]) / m00;
    x[p0] = c[0];
    x[1 - p0] = c[1];
  } else {
    x[p0] = c[0] / m00;
  }
}

// Frame::Frame_shortestDistance (Frame.cc:380-424); its missing `return` reads "no point"
__device__ bool shortest_distance(const double* L1, const double* L2, double* cross) {
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
  const double distmid = sqrt(dot3(mm, mm)) * 2;
  double n1 = 0, n2 = 0;
  for (int k = 0; k < 6; ++k) { n1 += L1[k] * L1[k]; n2 += L2[k] * L2[k]; }
  return distmid < sqrt(n1) + sqrt(n2);
}

__global__ void __launch_bounds__(kJThreads)
    junction_kernel(const psl_keyline* __restrict__ kl, const int32_t* __restrict__ n_lines, int line_cap,
                    const double* __restrict__ lines3d, int img_w, int img_h, float radius, float fan_thr,
                    float* __restrict__ fans, psl_line_junction* __restrict__ js, int cap, int32_t* __restrict__ n_fans,
                    int32_t* __restrict__ n_js) {
  extern __shared__ unsigned char smem[];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = min(n_lines[b], line_cap);
  JLine* R = reinterpret_cast<JLine*>(smem);
  int* offs = reinterpret_cast<int*>(R + line_cap);   // [line_cap + 1]
  const psl_keyline* k = kl + (size_t)b * line_cap;
  float* out = fans + (size_t)b * cap * 4;
  for (int i = tid; i < n; i += kJThreads) R[i] = make_jline(k[i], radius);
  __syncthreads();
  for (int i = tid; i < n; i += kJThreads) {
    int c = 0;
    float X, Y;
    for (int j = 0; j < 2 * n; ++j) c += keeps(R, n, i, j, fan_thr, img_w, img_h, X, Y) ? 1 : 0;
    offs[i + 1] = c;
  }
  __syncthreads();
  if (tid == 0) {
    offs[0] = 0;
    for (int i = 0; i < n; ++i) offs[i + 1] += offs[i];
    n_fans[b] = offs[n];
  }
  __syncthreads();
  const int total = offs[n], stored = min(total, cap);
  for (int i = tid; i < n; i += kJThreads) {
    int at = offs[i];
    if (at == offs[i + 1]) continue;
    for (int j = 0; j < 2 * n; ++j) {
      float X, Y;
      if (!keeps(R, n, i, j, fan_thr, img_w, img_h, X, Y)) continue;
      if (at < cap) { out[4 * at] = X; out[4 * at + 1] = Y; out[4 * at + 2] = (float)i; out[4 * at + 3] = (float)(j >= n ? j - n : j); }
      ++at;
    }
  }
  __syncthreads();
  // Frame::convertFansToKeyLines (Frame.cc:426-472): the fans whose two 3-D lines have a cross point, in order
  if (tid >= 32) return;
  int nj = 0;
  if (js && lines3d) {
    const double* L3 = lines3d + (size_t)b * line_cap * 6;
    psl_line_junction* oj = js + (size_t)b * cap;
    for (int base = 0; base < stored; base += 32) {
      const int f = base + tid;
      bool ok = false;
      double cross[3] = {0, 0, 0};
      int l1 = 0, l2 = 0;
      if (f < stored) {
        l1 = (int)out[4 * f + 2];
        l2 = (int)out[4 * f + 3];
        ok = shortest_distance(L3 + 6 * l1, L3 + 6 * l2, cross) && sqrt(dot3(cross, cross)) > DBL_EPSILON;
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        psl_line_junction J;
        J.l1 = l1; J.l2 = l2; J.cross2d_x = out[4 * f]; J.cross2d_y = out[4 * f + 1];
        J.cross3d[0] = cross[0]; J.cross3d[1] = cross[1]; J.cross3d[2] = cross[2];
        oj[nj + __popc(m & ((1u << tid) - 1u))] = J;
      }
      nj += __popc(m);
    }
  }
  if (tid == 0) n_js[b] = nj;
}

size_t junction_smem(int line_cap) { return (size_t)line_cap * sizeof(JLine) + (size_t)(line_cap + 1) * sizeof(int); }

}  // namespace

int launch_junctions(psl_ctx* ctx, const psl_keyline* kl, const int32_t* n_lines, int line_cap, int B,
                     const double* lines3d, int img_w, int img_h, float radius, float fan_thr, float* fans,
                     psl_line_junction* js, int cap, int32_t* n_fans, int32_t* n_js, cudaStream_t st) {
  const size_t smem = junction_smem(line_cap);
  if (smem > 200 * 1024) return fail(ctx, PSL_E_INVALID, "more than 4096 lines per frame");
  if (smem > 48 * 1024)
    PSL_CK(cudaFuncSetAttribute(junction_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  junction_kernel<<<B, kJThreads, smem, st>>>(kl, n_lines, line_cap, lines3d, img_w, img_h, radius, fan_thr, fans, js, cap,
                                              n_fans, n_js);
  return PSL_OK;
}

}  // namespace psl

using namespace psl;

extern "C" {

int psl_line_junctions_dev(psl_ctx* ctx, const psl_keyline* d_kl, const int32_t* d_n, int32_t line_cap, int32_t B,
                           const double* d_lines3d, int32_t img_w, int32_t img_h, float radius, float fan_thr,
                           float* d_fans, psl_line_junction* d_junctions, int32_t cap, int32_t* d_n_fans,
                           int32_t* d_n_junctions) {
  if (!ctx) return PSL_E_INVALID;
  if (B < 0 || line_cap < 1 || cap < 1 || img_w <= 8 || img_h <= 8 || !(radius >= 0.f) ||
      (B > 0 && (!d_kl || !d_n || !d_fans || !d_n_fans || !d_n_junctions)) || ((d_junctions != nullptr) != (d_lines3d != nullptr)))
    return fail(ctx, PSL_E_INVALID, "bad argument");
  if (B == 0) return PSL_OK;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  size_t e = prof_mark(ctx);
  int rc = launch_junctions(ctx, d_kl, d_n, line_cap, B, d_lines3d, img_w, img_h, radius, fan_thr, d_fans, d_junctions,
                            cap, d_n_fans, d_n_junctions, ctx->stream);
  if (rc) return rc;
  prof_span(ctx, 15, e, 1);
  PSL_CK(cudaGetLastError());
  return PSL_OK;
}

int psl_line_junctions(psl_ctx* ctx, const psl_keyline* kl_un, const double* lines3d, int32_t n, int32_t img_w,
                       int32_t img_h, float radius, float fan_thr, float* fans, psl_line_junction* junctions,
                       int32_t cap, int32_t* n_fans, int32_t* n_junctions) {
  if (!ctx) return PSL_E_INVALID;
  if (n < 0 || cap < 0 || !n_fans || !n_junctions || img_w <= 8 || img_h <= 8 || !(radius >= 0.f) ||
      (n > 0 && !kl_un) || (cap > 0 && !fans) || (junctions && !lines3d))
    return fail(ctx, PSL_E_INVALID, "bad argument");
  *n_fans = 0;
  *n_junctions = 0;
  if (n == 0) return PSL_OK;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  const int ocap = std::max(cap, 1);
  const bool want3d = junctions != nullptr;
  DevBuf* M = ctx->m_misc;
  int rc;
#define PSL_UPJ(buf, src, nbytes)                                                                              \
  do {                                                                                                         \
    if ((rc = ensure(ctx, buf, (nbytes)))) return rc;                                                          \
    PSL_CK(cudaMemcpyAsync((buf).p, (src), (nbytes), cudaMemcpyHostToDevice, ctx->stream));                    \
  } while (0)
  PSL_UPJ(M[0], kl_un, (size_t)n * sizeof(psl_keyline));
  if (want3d) PSL_UPJ(M[1], lines3d, (size_t)n * 48);
  const int32_t nn[1] = {n};
  PSL_UPJ(ctx->m_n, nn, sizeof(nn));
  if ((rc = ensure(ctx, M[2], (size_t)ocap * 16))) return rc;
  if ((rc = ensure(ctx, M[3], (size_t)ocap * sizeof(psl_line_junction)))) return rc;
  if ((rc = ensure(ctx, ctx->m_nm, 8))) return rc;
  size_t e = prof_mark(ctx);
  rc = launch_junctions(ctx, M[0].as<psl_keyline>(), ctx->m_n.as<int32_t>(), n, 1, want3d ? M[1].as<double>() : nullptr,
                        img_w, img_h, radius, fan_thr, M[2].as<float>(), want3d ? M[3].as<psl_line_junction>() : nullptr,
                        ocap, ctx->m_nm.as<int32_t>(), ctx->m_nm.as<int32_t>() + 1, ctx->stream);
  if (rc) return rc;
  prof_span(ctx, 15, e, 1);
  PSL_CK(cudaGetLastError());
  int32_t cnt[2] = {0, 0};
  PSL_CK(cudaMemcpyAsync(cnt, ctx->m_nm.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  if ((rc = check_status(ctx))) return rc;   // synchronises
  *n_fans = cnt[0];
  *n_junctions = cnt[1];
  const int nf = std::min(cnt[0], cap);
  if (nf > 0) PSL_CK(cudaMemcpyAsync(fans, M[2].p, (size_t)nf * 16, cudaMemcpyDeviceToHost, ctx->stream));
  if (want3d && cnt[1] > 0)
    PSL_CK(cudaMemcpyAsync(junctions, M[3].p, (size_t)cnt[1] * sizeof(psl_line_junction), cudaMemcpyDeviceToHost, ctx->stream));
  PSL_CK(cudaStreamSynchronize(ctx->stream));
  if (cnt[0] > cap) return fail(ctx, PSL_E_CAPACITY, "more junctions than `cap`");
  return PSL_OK;
#undef PSL_UPJ
}

}  // extern "C"
