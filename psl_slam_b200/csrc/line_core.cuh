// Sequential per-frame part of LINEextractor::operator() after LSD (add_src/LineExtractor.cpp:336-363):
// checkLineExtremes of the LSD wrapper (Thirdparty/line_descriptor/src/LSDDetector_custom.cpp:112-138),
// optimizeAndMergeLines_lsd (add_src/uselongline.cpp:24-351, 449-485), convertVec4fToKeyLine (:411-447),
// the top-N filter by response and the 2-D line equations (LineExtractor.cpp:342-363).
//
// A few hundred to ~1500 short records per frame, pointer-chasing control flow (angle-sorted pair scan with
// early break, BFS clustering through std::set, sub-clusters, folded two-line merges): one thread per
// frame walks it exactly as the reference does and the batch supplies the parallelism.  The same source
// builds for the host (PSL_HOST_EMU) so the CPU suite can check it against the oracle.
//
// Pinned choices: index sorts are stable; unqualified atan/atan2/sin/cos/sqrt evaluate in
// fp64 and round once; no FMA contraction.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/psl_frontend.h"

#ifdef PSL_HOST_EMU
#define PSL_LN_HD inline
#define PSL_LN_FMUL(a, b) ((a) * (b))
#define PSL_LN_FADD(a, b) ((a) + (b))
#define PSL_LN_FSUB(a, b) ((a) - (b))
#define PSL_LN_FDIV(a, b) ((a) / (b))
#else
#define PSL_LN_HD __device__ __forceinline__
#define PSL_LN_FMUL(a, b) __fmul_rn((a), (b))
#define PSL_LN_FADD(a, b) __fadd_rn((a), (b))
#define PSL_LN_FSUB(a, b) __fsub_rn((a), (b))
#define PSL_LN_FDIV(a, b) __fdiv_rn((a), (b))
#endif

namespace psl {
namespace line {

constexpr int kNbCap = 32;  // neighbours kept per segment in MergeLines
constexpr double kPi = 3.14159265358979323846;

struct Seg { float v[4]; };

// Scratch of one frame (all arrays sized for `cap` segments).
// pair-scan record of the warp-cooperative kernel: segment | angle | denominator of PointLineDistance
struct alignas(16) ScanRec { Seg s; float angle, tden; double den; };  // tden: (float)(distance threshold * den)

struct MergeScratch {
  int cap;
  float* angles;     // [cap]
  float* length;     // [cap]
  uint16_t* order;   // [cap] indices sorted by angle
  uint16_t* tmp16;   // [cap] merge-sort buffer / cluster member list
  uint16_t* nb;      // [cap][kNbCap] neighbour lists
  uint16_t* nb_cnt;  // [cap]
  int16_t* code;     // [cap] cluster code
  uint16_t* check;   // [cap] BFS frontier
  uint16_t* loc;     // [cap] position inside the sorted cluster
  uint8_t* flag;     // [cap] frontier bitmap / `clustered`
  int overflow;      // set when a neighbour list overflows
  float* sangles;    // [cap] angles in scan order (warp-cooperative kernel only)
  uint16_t* fw;      // [cap][kNbCap] a row's own partners (warp-cooperative kernel only)
  uint32_t* sort_cnt;  // bucket counters of the rank sort, shared memory (warp-cooperative kernel only)
  ScanRec* scan;       // [cap] lines in scan order (warp-cooperative kernel only)
};

PSL_LN_HD float point_line_distance(const Seg& l, float x0, float y0) {  // uselongline.cpp:5-15
  const float x1 = l.v[0], y1 = l.v[1], x2 = l.v[2], y2 = l.v[3];
  const float num = fabsf(PSL_LN_FADD(PSL_LN_FADD(PSL_LN_FMUL(PSL_LN_FSUB(y2, y1), x0), PSL_LN_FMUL(PSL_LN_FSUB(x1, x2), y0)),
                                      PSL_LN_FSUB(PSL_LN_FMUL(x2, y1), PSL_LN_FMUL(x1, y2))));
  const double a = (double)PSL_LN_FSUB(y2, y1), b = (double)PSL_LN_FSUB(x1, x2);
  return (float)((double)num / sqrt(a * a + b * b));
}

PSL_LN_HD float angle_diff(float a1, float a2) {  // uselongline.cpp:17-22
  const float c1 = fabsf(PSL_LN_FSUB(a2, a1));
  const float mn = a1 < a2 ? a1 : a2, mx = a1 < a2 ? a2 : a1;  // std::min / std::max
  const float c2 = (float)(kPi + (double)mn - (double)mx);
  return c2 < c1 ? c2 : c1;
}

PSL_LN_HD Seg merge_two(const Seg& l1, const Seg& l2) {  // uselongline.cpp:266-334
  const float ax = l1.v[0], ay = l1.v[1], bx = l1.v[2], by = l1.v[3];
  const float cx = l2.v[0], cy = l2.v[1], dx = l2.v[2], dy = l2.v[3];
  const float dlix = PSL_LN_FSUB(bx, ax), dliy = PSL_LN_FSUB(by, ay), dljx = PSL_LN_FSUB(dx, cx), dljy = PSL_LN_FSUB(dy, cy);
  const double li = sqrt((double)PSL_LN_FMUL(dlix, dlix) + (double)PSL_LN_FMUL(dliy, dliy));
  const double lj = sqrt((double)PSL_LN_FMUL(dljx, dljx) + (double)PSL_LN_FMUL(dljy, dljy));
  const double xg = (li * (double)PSL_LN_FADD(ax, bx) + lj * (double)PSL_LN_FADD(cx, dx)) / (2.0 * (li + lj));
  const double yg = (li * (double)PSL_LN_FADD(ay, by) + lj * (double)PSL_LN_FADD(cy, dy)) / (2.0 * (li + lj));
  const double thi = dlix == 0.0f ? kPi / 2.0 : atan((double)PSL_LN_FDIV(dliy, dlix));
  const double thj = dljx == 0.0f ? kPi / 2.0 : atan((double)PSL_LN_FDIV(dljy, dljx));
  double thr;
  if (fabs(thi - thj) <= kPi / 2.0) thr = (li * thi + lj * thj) / (li + lj);
  else {
    const double tmp = thj - kPi * (thj / fabs(thj));
    thr = li * thi + lj * tmp;
    thr /= (li + lj);
  }
  const double s = sin(thr), c = cos(thr);
  const double axg = ((double)ay - yg) * s + ((double)ax - xg) * c;
  const double bxg = ((double)by - yg) * s + ((double)bx - xg) * c;
  const double cxg = ((double)cy - yg) * s + ((double)cx - xg) * c;
  const double dxg = ((double)dy - yg) * s + ((double)dx - xg) * c;
  double d1 = cxg < dxg ? cxg : dxg, d2 = cxg < dxg ? dxg : cxg;
  d1 = bxg < d1 ? bxg : d1; d2 = bxg > d2 ? bxg : d2;
  d1 = axg < d1 ? axg : d1; d2 = axg > d2 ? axg : d2;
  Seg r;
  r.v[0] = (float)(d1 * c + xg); r.v[1] = (float)(d1 * s + yg);
  r.v[2] = (float)(d2 * c + xg); r.v[3] = (float)(d2 * s + yg);
  return r;
}

// stable bottom-up merge sort of idx[0..n) by key ascending (desc = false) or descending (desc = true)
PSL_LN_HD void stable_sort_idx(uint16_t* idx, uint16_t* tmp, int n, const float* key, bool desc) {
  for (int width = 1; width < n; width <<= 1) {
    for (int lo = 0; lo < n; lo += 2 * width) {
      const int mid = lo + width < n ? lo + width : n, hi = lo + 2 * width < n ? lo + 2 * width : n;
      int a = lo, b = mid, o = lo;
      while (a < mid && b < hi) {
        const float ka = key[idx[a]], kb = key[idx[b]];
        const bool take_b = desc ? (kb > ka) : (kb < ka);  // strict: ties keep the left run first
        tmp[o++] = take_b ? idx[b++] : idx[a++];
      }
      while (a < mid) tmp[o++] = idx[a++];
      while (b < hi) tmp[o++] = idx[b++];
    }
    for (int i = 0; i < n; ++i) idx[i] = tmp[i];
  }
}

// A segment with its end points ordered along the scan axis of a pair test: x when the row's line is closer to
// horizontal than to vertical, else y (uselongline.cpp:84-99).
struct AxisSeg { float ax, ay, bx, by; };  // a = first end along the axis, b = last
PSL_LN_HD AxisSeg along_axis(const Seg& s, bool horiz) {
  const bool flip = horiz ? (s.v[2] < s.v[0]) : (s.v[3] < s.v[1]);
  AxisSeg r;
  r.ax = flip ? s.v[2] : s.v[0]; r.ay = flip ? s.v[3] : s.v[1];
  r.bx = flip ? s.v[0] : s.v[2]; r.by = flip ? s.v[1] : s.v[3];
  return r;
}
// Two ordered segments overlap along the axis, or the gap between their facing end points is shorter than
// sqrt(gap_sq_thr) (uselongline.cpp:123-143): the one that ends first supplies the tail, the other one the head.
PSL_LN_HD bool ends_meet(const AxisSeg& p, const AxisSeg& q, bool horiz, float gap_sq_thr) {
  const bool q_first = horiz ? (p.bx > q.bx) : (p.by > q.by);
  const float tx = q_first ? q.bx : p.bx, ty = q_first ? q.by : p.by;
  const float hx = q_first ? p.ax : q.ax, hy = q_first ? p.ay : q.ay;
  if (horiz ? (tx >= hx) : (ty >= hy)) return true;
  const float gx = PSL_LN_FSUB(hx, tx), gy = PSL_LN_FSUB(hy, ty);
  return PSL_LN_FADD(PSL_LN_FMUL(gx, gx), PSL_LN_FMUL(gy, gy)) < gap_sq_thr;
}

// MergeLines (uselongline.cpp:24-264); returns the number of lines written to dst
PSL_LN_HD int merge_lines(const Seg* src, int n, Seg* dst, float angle_thr, float distance_thr, float endpoint_threshold,
                          MergeScratch& S) {
  if (n <= 0) return 0;
  for (int i = 0; i < n; ++i) {
    const float dx = PSL_LN_FSUB(src[i].v[2], src[i].v[0]), dy = PSL_LN_FSUB(src[i].v[3], src[i].v[1]);
    S.angles[i] = (float)atan((double)PSL_LN_FDIV(dy, dx));
    S.length[i] = sqrtf(PSL_LN_FADD(PSL_LN_FMUL(dx, dx), PSL_LN_FMUL(dy, dy)));
    S.order[i] = (uint16_t)i;
    S.nb_cnt[i] = 0;
    S.code[i] = -1;
  }
  stable_sort_idx(S.order, S.tmp16, n, S.angles, false);
  const float gap_sq_thr = PSL_LN_FMUL(endpoint_threshold, endpoint_threshold);
  const float quarter_turn = (float)(kPi / 4.0);
  for (int i = 0; i < n; ++i) {
    const int idx1 = S.order[i];
    const float angle1 = S.angles[idx1];
    const bool horiz = fabsf(angle1) < quarter_turn;
    const AxisSeg p = along_axis(src[idx1], horiz);
    for (int j = i + 1; j < n; ++j) {
      const int idx2 = S.order[j];
      const float d_angle = angle_diff(angle1, S.angles[idx2]);
      if (d_angle > angle_thr) {
        if ((double)fabsf(angle1) < (kPi / 2 - (double)angle_thr)) break;
        else continue;
      }
      const float mx1 = (float)(0.5 * (double)PSL_LN_FADD(src[idx1].v[0], src[idx1].v[2]));
      const float my1 = (float)(0.5 * (double)PSL_LN_FADD(src[idx1].v[1], src[idx1].v[3]));
      const float mx2 = (float)(0.5 * (double)PSL_LN_FADD(src[idx2].v[0], src[idx2].v[2]));
      const float my2 = (float)(0.5 * (double)PSL_LN_FADD(src[idx2].v[1], src[idx2].v[3]));
      if (point_line_distance(src[idx2], mx1, my1) > distance_thr && point_line_distance(src[idx1], mx2, my2) > distance_thr)
        continue;
      if (ends_meet(p, along_axis(src[idx2], horiz), horiz, gap_sq_thr)) {
        if (S.nb_cnt[idx1] < kNbCap && S.nb_cnt[idx2] < kNbCap) {
          S.nb[idx1 * kNbCap + S.nb_cnt[idx1]++] = (uint16_t)idx2;
          S.nb[idx2 * kNbCap + S.nb_cnt[idx2]++] = (uint16_t)idx1;
        } else {
          S.overflow = 1;
        }
      }
    }
  }
  int nd = 0;
  // connected components (:153-190), then sub-clusters (:193-229) and the folded merge (:231-262) per component
  for (int i = 0; i < n; ++i) {
    if (S.code[i] >= 0) continue;
    S.code[i] = 1;
    uint16_t* cl = S.tmp16;  // members of this component
    int cs = 0, ncheck = 0;
    cl[cs++] = (uint16_t)i;
    for (int k = 0; k < S.nb_cnt[i]; ++k) S.check[ncheck++] = S.nb[i * kNbCap + k];
    while (ncheck > 0) {
      // std::set<size_t> tmp: unique, ascending
      int lo = n, hi = -1;
      for (int c = 0; c < ncheck; ++c) {
        const int j = S.check[c];
        if (S.code[j] < 0) { S.code[j] = 1; cl[cs++] = (uint16_t)j; }
        for (int k = 0; k < S.nb_cnt[j]; ++k) {
          const int q = S.nb[j * kNbCap + k];
          if (S.code[q] < 0) { S.flag[q] = 1; lo = q < lo ? q : lo; hi = q > hi ? q : hi; }
        }
      }
      ncheck = 0;
      for (int q = lo; q <= hi; ++q)
        if (S.flag[q]) {
          S.flag[q] = 0;
          if (S.code[q] < 0) S.check[ncheck++] = (uint16_t)q;
        }
    }
    if (cs <= 2) {
      Seg nl = src[cl[0]];
      for (int k = 0; k < cs; ++k) nl = merge_two(nl, src[cl[k]]);
      dst[nd++] = nl;
      continue;
    }
    stable_sort_idx(cl, S.check, cs, S.length, true);
    for (int k = 0; k < cs; ++k) { S.loc[cl[k]] = (uint16_t)k; S.flag[k] = 0; }  // flag = `clustered`
    for (int j = 0; j < cs; ++j) {
      if (S.flag[j]) continue;
      const int li = cl[j];
      Seg nl = merge_two(src[li], src[li]);
      for (int k = 0; k < S.nb_cnt[li]; ++k) {
        const int q = S.nb[li * kNbCap + k];
        S.flag[S.loc[q]] = 1;
        nl = merge_two(nl, src[q]);
      }
      dst[nd++] = nl;
    }
    for (int k = 0; k < cs; ++k) S.flag[k] = 0;
  }
  return nd;
}

PSL_LN_HD int filter_short(Seg* lines, int n, float length_thr) {  // uselongline.cpp:338-351
  const float thr2 = PSL_LN_FMUL(length_thr, length_thr);
  int m = 0;
  for (int i = 0; i < n; ++i) {
    const float dx = PSL_LN_FSUB(lines[i].v[2], lines[i].v[0]), dy = PSL_LN_FSUB(lines[i].v[3], lines[i].v[1]);
    if (PSL_LN_FADD(PSL_LN_FMUL(dx, dx), PSL_LN_FMUL(dy, dy)) > thr2) lines[m++] = lines[i];
  }
  return m;
}

PSL_LN_HD void clamp_segment(Seg& e, int w, int h) {  // LSDDetector_custom.cpp:112-138
  if (e.v[0] < 0) e.v[0] = 0;
  if (e.v[0] >= w) e.v[0] = (float)w - 1.0f;
  if (e.v[2] < 0) e.v[2] = 0;
  if (e.v[2] >= w) e.v[2] = (float)w - 1.0f;
  if (e.v[1] < 0) e.v[1] = 0;
  if (e.v[1] >= h) e.v[1] = (float)h - 1.0f;
  if (e.v[3] < 0) e.v[3] = 0;
  if (e.v[3] >= h) e.v[3] = (float)h - 1.0f;
}

// cv::LineIterator(img, Point2f, Point2f).count, 8-connected (cvRound + cv::clipLine on 64-bit points)
PSL_LN_HD int line_iterator_count(int w, int h, float x1f, float y1f, float x2f, float y2f) {
#ifdef PSL_HOST_EMU
  long long x1 = lrintf(x1f), y1 = lrintf(y1f), x2 = lrintf(x2f), y2 = lrintf(y2f);
#else
  long long x1 = __float2int_rn(x1f), y1 = __float2int_rn(y1f), x2 = __float2int_rn(x2f), y2 = __float2int_rn(y2f);
#endif
  const long long right = w - 1, bottom = h - 1;
  if (w <= 0 || h <= 0) return 0;
  int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
  int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
  if ((c1 & c2) == 0 && (c1 | c2) != 0) {
    long long a;
    if (c1 & 12) {
      a = c1 < 8 ? 0 : bottom;
      x1 += (long long)((double)(a - y1) * (double)(x2 - x1) / (double)(y2 - y1));
      y1 = a;
      c1 = (x1 < 0) + (x1 > right) * 2;
    }
    if (c2 & 12) {
      a = c2 < 8 ? 0 : bottom;
      x2 += (long long)((double)(a - y2) * (double)(x2 - x1) / (double)(y2 - y1));
      y2 = a;
      c2 = (x2 < 0) + (x2 > right) * 2;
    }
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      if (c1) {
        a = c1 == 1 ? 0 : right;
        y1 += (long long)((double)(a - x1) * (double)(y2 - y1) / (double)(x2 - x1));
        x1 = a;
        c1 = 0;
      }
      if (c2) {
        a = c2 == 1 ? 0 : right;
        y2 += (long long)((double)(a - x2) * (double)(y2 - y1) / (double)(x2 - x1));
        x2 = a;
        c2 = 0;
      }
    }
  }
  if ((c1 | c2) != 0) return 0;
  const long long dx = x2 > x1 ? x2 - x1 : x1 - x2, dy = y2 > y1 ? y2 - y1 : y1 - y2;
  return (int)((dx > dy ? dx : dy) + 1);
}

PSL_LN_HD void make_keyline(const Seg& l, int i, int w, int h, psl_keyline& k) {  // uselongline.cpp:411-447
  k.start_x = l.v[0]; k.start_y = l.v[1]; k.end_x = l.v[2]; k.end_y = l.v[3];
  k.s_oct_x = l.v[0]; k.s_oct_y = l.v[1]; k.e_oct_x = l.v[2]; k.e_oct_y = l.v[3];
  const double ddx = (double)PSL_LN_FSUB(l.v[0], l.v[2]), ddy = (double)PSL_LN_FSUB(l.v[1], l.v[3]);
  k.line_length = (float)sqrt(ddx * ddx + ddy * ddy);
  k.angle = (float)atan2((double)PSL_LN_FSUB(k.end_y, k.start_y), (double)PSL_LN_FSUB(k.end_x, k.start_x));
  k.class_id = i;
  k.octave = 0;
  k.size = PSL_LN_FMUL(PSL_LN_FSUB(k.end_x, k.start_x), PSL_LN_FSUB(k.end_y, k.start_y));
  k.pt_x = PSL_LN_FDIV(PSL_LN_FADD(k.end_x, k.start_x), 2.f);
  k.pt_y = PSL_LN_FDIV(PSL_LN_FADD(k.end_y, k.start_y), 2.f);
  k.response = PSL_LN_FDIV(k.line_length, (float)(w > h ? w : h));
  k.num_pixels = line_iterator_count(w, h, l.v[0], l.v[1], l.v[2], l.v[3]);
}

// Everything between the raw LSD segments and the LBD stage for one frame.  raw/t1/t2: Seg[cap] buffers.
// Writes up to `kl_cap` keylines (top-N by response when more than nfeatures) and their line equations.
// Returns the number of keylines, or -1 when kl_cap is too small.
PSL_LN_HD int frame_lines(Seg* raw, int n_raw, Seg* t1, Seg* t2, int w, int h, int nfeatures, MergeScratch& S,
                          psl_keyline* kl, double* lineeq, int kl_cap) {
  for (int i = 0; i < n_raw; ++i) clamp_segment(raw[i], w, h);
  int n1 = merge_lines(raw, n_raw, t1, 0.05f, 5.f, 15.f, S);   // uselongline.cpp:458
  n1 = filter_short(t1, n1, 30.f);
  int n2 = merge_lines(t1, n1, t2, 0.03f, 3.f, 30.f, S);       // :464
  n2 = filter_short(t2, n2, 50.f);
  int n = n2;
  if (n2 > nfeatures) {  // LineExtractor.cpp:342-348 (stable by response, descending)
    for (int i = 0; i < n2; ++i) {
      const double ddx = (double)PSL_LN_FSUB(t2[i].v[0], t2[i].v[2]), ddy = (double)PSL_LN_FSUB(t2[i].v[1], t2[i].v[3]);
      S.length[i] = PSL_LN_FDIV((float)sqrt(ddx * ddx + ddy * ddy), (float)(w > h ? w : h));  // response
      S.order[i] = (uint16_t)i;
    }
    stable_sort_idx(S.order, S.tmp16, n2, S.length, true);
    n = nfeatures;
  } else {
    for (int i = 0; i < n2; ++i) S.order[i] = (uint16_t)i;
  }
  if (n > kl_cap) return -1;
  for (int i = 0; i < n; ++i) {
    psl_keyline k;
    make_keyline(t2[S.order[i]], n2 > nfeatures ? i : (int)S.order[i], w, h, k);
    kl[i] = k;
    // sp x ep normalised by its first two components (LineExtractor.cpp:352-363), fp64
    const double sx = k.start_x, sy = k.start_y, ex = k.end_x, ey = k.end_y;
    const double l0 = sy * 1.0 - 1.0 * ey, l1 = 1.0 * ex - sx * 1.0, l2 = sx * ey - sy * ex;
    const double nrm = sqrt(l0 * l0 + l1 * l1);
    lineeq[3 * i] = l0 / nrm; lineeq[3 * i + 1] = l1 / nrm; lineeq[3 * i + 2] = l2 / nrm;
  }
  return n;
}

}  // namespace line
}  // namespace psl
