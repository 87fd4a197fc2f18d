// TEST INFRASTRUCTURE — CPU oracle for the line front end after LSD: the long-line merge
// (add_src/uselongline.cpp), the KeyLine fields (uselongline.cpp:411-447), the top-N filter and line
// equations (add_src/LineExtractor.cpp:342-363) and the LBD descriptor of opencv_contrib's
// BinaryDescriptor::compute as vendored in Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp
// (:219-261 weights, :351-399 blur+Sobel, :1027-1373 computeLBD, :402-413 + :76-109 binarisation).
//
// Pinned choices (the reference leaves them to the toolchain): index sorts are stable (std::sort there);
// unqualified cos/sin/atan/atan2/sqrt on floats evaluate in double and round once (no `using namespace std`
// in those translation units); no FMA contraction (-ffp-contract=off).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <set>
#include <unordered_map>
#include <vector>

#include "psl_oracle.h"

namespace {

struct V4 { float v[4]; };

// PointLineDistance, uselongline.cpp:5-15
float point_line_distance(const V4& l, float x0, float y0) {
  const float x1 = l.v[0], y1 = l.v[1], x2 = l.v[2], y2 = l.v[3];
  const float num = std::fabs((y2 - y1) * x0 + (x1 - x2) * y0 + ((x2 * y1) - (x1 * y2)));
  const double den = std::sqrt(std::pow((double)(y2 - y1), 2) + std::pow((double)(x1 - x2), 2));
  return (float)(num / den);
}

// AngleDiff, uselongline.cpp:17-22
float angle_diff(float a1, float a2) {
  const float c1 = std::abs(a2 - a1);
  const float c2 = (float)(M_PI + std::min(a1, a2) - std::max(a1, a2));
  return std::min(c1, c2);
}

// MergeTwoLines, uselongline.cpp:266-334
V4 merge_two(const V4& l1, const V4& l2) {
  const float ax = l1.v[0], ay = l1.v[1], bx = l1.v[2], by = l1.v[3];
  const float cx = l2.v[0], cy = l2.v[1], dx = l2.v[2], dy = l2.v[3];
  const float dlix = bx - ax, dliy = by - ay, dljx = dx - cx, dljy = dy - cy;
  const double li = sqrt((double)(dlix * dlix) + (double)(dliy * dliy));
  const double lj = sqrt((double)(dljx * dljx) + (double)(dljy * dljy));
  const double xg = (li * (double)(ax + bx) + lj * (double)(cx + dx)) / (double)(2.0 * (li + lj));
  const double yg = (li * (double)(ay + by) + lj * (double)(cy + dy)) / (double)(2.0 * (li + lj));
  const double thi = dlix == 0.0f ? M_PI / 2.0 : atan((double)(dliy / dlix));
  const double thj = dljx == 0.0f ? M_PI / 2.0 : atan((double)(dljy / dljx));
  double thr;
  if (fabs(thi - thj) <= M_PI / 2.0) thr = (li * thi + lj * thj) / (li + lj);
  else {
    const double tmp = thj - M_PI * (thj / fabs(thj));
    thr = li * thi + lj * tmp;
    thr /= (li + lj);
  }
  const double s = sin(thr), c = cos(thr);
  const double axg = ((double)ay - yg) * s + ((double)ax - xg) * c;
  const double bxg = ((double)by - yg) * s + ((double)bx - xg) * c;
  const double cxg = ((double)cy - yg) * s + ((double)cx - xg) * c;
  const double dxg = ((double)dy - yg) * s + ((double)dx - xg) * c;
  const double d1 = std::min(axg, std::min(bxg, std::min(cxg, dxg)));
  const double d2 = std::max(axg, std::max(bxg, std::max(cxg, dxg)));
  V4 r;
  r.v[0] = (float)(d1 * c + xg); r.v[1] = (float)(d1 * s + yg);
  r.v[2] = (float)(d2 * c + xg); r.v[3] = (float)(d2 * s + yg);
  return r;
}

// MergeLines, uselongline.cpp:24-264 (empty input -> empty output; the reference dereferences [0])
void merge_lines(const std::vector<V4>& src, std::vector<V4>& dst, float angle_thr, float distance_thr,
                 float endpoint_threshold) {
  dst.clear();
  const size_t n = src.size();
  if (!n) return;
  std::vector<float> angles(n), length(n);
  for (size_t i = 0; i < n; ++i) {
    const float dx = src[i].v[2] - src[i].v[0], dy = src[i].v[3] - src[i].v[1];
    angles[i] = (float)atan((double)(dy / dx));  // (dy/dx).atan() on an Eigen float array
    length[i] = sqrtf(dx * dx + dy * dy);
  }
  std::vector<size_t> indices(n);
  for (size_t a = 0; a < n; ++a) indices[a] = a;
  std::stable_sort(indices.begin(), indices.end(), [&](size_t i1, size_t i2) { return angles[i1] < angles[i2]; });
  const float ep_thr = endpoint_threshold * endpoint_threshold;
  const float quater_PI = (float)(M_PI / 4.0);
  std::vector<std::vector<size_t>> neighbors(n);
  for (size_t i = 0; i < n; ++i) {
    const size_t idx1 = indices[i];
    float x11 = src[idx1].v[0], y11 = src[idx1].v[1], x12 = src[idx1].v[2], y12 = src[idx1].v[3];
    const float angle1 = angles[idx1];
    const bool to_sort_x = std::abs(angle1) < quater_PI;
    if ((to_sort_x && (x12 < x11)) || ((!to_sort_x) && y12 < y11)) { std::swap(x11, x12); std::swap(y11, y12); }
    for (size_t j = i + 1; j < n; ++j) {
      const size_t idx2 = indices[j];
      float x21 = src[idx2].v[0], y21 = src[idx2].v[1], x22 = src[idx2].v[2], y22 = src[idx2].v[3];
      if ((to_sort_x && (x22 < x21)) || ((!to_sort_x) && y22 < y21)) { std::swap(x21, x22); std::swap(y21, y22); }
      const float angle2 = angles[idx2];
      const float d_angle = angle_diff(angle1, angle2);
      if (d_angle > angle_thr) {
        if (std::abs(angle1) < (M_PI_2 - angle_thr)) break;
        else continue;
      }
      const float mid_x1 = (float)(0.5 * (src[idx1].v[0] + src[idx1].v[2])), mid_y1 = (float)(0.5 * (src[idx1].v[1] + src[idx1].v[3]));
      const float mid_x2 = (float)(0.5 * (src[idx2].v[0] + src[idx2].v[2])), mid_y2 = (float)(0.5 * (src[idx2].v[1] + src[idx2].v[3]));
      const float m1 = point_line_distance(src[idx2], mid_x1, mid_y1);
      const float m2 = point_line_distance(src[idx1], mid_x2, mid_y2);
      if (m1 > distance_thr && m2 > distance_thr) continue;
      float cx12, cy12, cx21, cy21;
      if ((to_sort_x && x12 > x22) || (!to_sort_x && y12 > y22)) { cx12 = x22; cy12 = y22; cx21 = x11; cy21 = y11; }
      else { cx12 = x12; cy12 = y12; cx21 = x21; cy21 = y21; }
      bool to_merge = ((to_sort_x && cx12 >= cx21) || (!to_sort_x && cy12 >= cy21));
      if (!to_merge) {
        const float d_ep = (cx21 - cx12) * (cx21 - cx12) + (cy21 - cy12) * (cy21 - cy12);
        to_merge = d_ep < ep_thr;
      }
      if (to_merge) { neighbors[idx1].push_back(idx2); neighbors[idx2].push_back(idx1); }
    }
  }
  // connected components (:153-190)
  std::vector<int> codes(n, -1);
  std::vector<std::vector<size_t>> clusters;
  for (size_t i = 0; i < n; ++i) {
    if (codes[i] >= 0) continue;
    const int code = (int)clusters.size();
    codes[i] = code;
    std::vector<size_t> to_check = neighbors[i], cluster{i};
    while (!to_check.empty()) {
      std::set<size_t> tmp;
      for (size_t j : to_check) {
        if (codes[j] < 0) { codes[j] = code; cluster.push_back(j); }
        for (size_t k : neighbors[j]) if (codes[k] < 0) tmp.insert(k);
      }
      to_check.assign(tmp.begin(), tmp.end());
    }
    clusters.push_back(cluster);
  }
  // sub-clusters (:193-229)
  std::vector<std::vector<size_t>> subs;
  for (auto& cluster : clusters) {
    const size_t cs = cluster.size();
    if (cs <= 2) { subs.push_back(cluster); continue; }
    std::stable_sort(cluster.begin(), cluster.end(), [&](size_t a, size_t b) { return length[a] > length[b]; });
    std::unordered_map<size_t, size_t> loc;
    for (size_t i = 0; i < cs; ++i) loc[cluster[i]] = i;
    std::vector<bool> clustered(cs, false);
    for (size_t j = 0; j < cs; ++j) {
      if (clustered[j]) continue;
      std::vector<size_t> sub{cluster[j]};
      for (size_t k : neighbors[cluster[j]]) { clustered[loc[k]] = true; sub.push_back(k); }
      subs.push_back(sub);
    }
  }
  // fold every sub-cluster with MergeTwoLines, starting with the head merged with itself (:243-255)
  dst.reserve(subs.size());
  for (auto& c : subs) {
    V4 nl = src[c[0]];
    for (size_t i = 0; i < c.size(); ++i) nl = merge_two(nl, src[c[i]]);
    dst.push_back(nl);
  }
}

// FilterShortLines, uselongline.cpp:338-351
void filter_short(std::vector<V4>& lines, float length_thr) {
  const float thr2 = length_thr * length_thr;
  size_t m = 0;
  for (size_t i = 0; i < lines.size(); ++i) {
    const float dx = lines[i].v[2] - lines[i].v[0], dy = lines[i].v[3] - lines[i].v[1];
    if (dx * dx + dy * dy > thr2) lines[m++] = lines[i];
  }
  lines.resize(m);
}

// cv::clipLine on 64-bit points (OpenCV imgproc drawing.cpp), needed by cv::LineIterator
bool clip_line(int64_t w, int64_t h, int64_t& x1, int64_t& y1, int64_t& x2, int64_t& y2) {
  const int64_t right = w - 1, bottom = h - 1;
  if (w <= 0 || h <= 0) return false;
  int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
  int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
  if ((c1 & c2) == 0 && (c1 | c2) != 0) {
    int64_t a;
    if (c1 & 12) {
      a = c1 < 8 ? 0 : bottom;
      x1 += (int64_t)((double)(a - y1) * (x2 - x1) / (y2 - y1));
      y1 = a;
      c1 = (x1 < 0) + (x1 > right) * 2;
    }
    if (c2 & 12) {
      a = c2 < 8 ? 0 : bottom;
      x2 += (int64_t)((double)(a - y2) * (x2 - x1) / (y2 - y1));
      y2 = a;
      c2 = (x2 < 0) + (x2 > right) * 2;
    }
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      if (c1) {
        a = c1 == 1 ? 0 : right;
        y1 += (int64_t)((double)(a - x1) * (y2 - y1) / (x2 - x1));
        x1 = a;
        c1 = 0;
      }
      if (c2) {
        a = c2 == 1 ? 0 : right;
        y2 += (int64_t)((double)(a - x2) * (y2 - y1) / (x2 - x1));
        x2 = a;
        c2 = 0;
      }
    }
  }
  return (c1 | c2) == 0;
}

const int combinations[32][2] = {{0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6},
                                 {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7}, {2, 8}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8},
                                 {4, 5}, {4, 6}, {4, 7}, {4, 8}, {5, 6}, {5, 7}, {5, 8}, {6, 7}, {6, 8}, {7, 8}};

inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

}  // namespace

extern "C" {

// cv::LineIterator(img, Point2f, Point2f).count (8-connected): endpoints rounded (cvRound), clipped to the image
int orc_line_iterator_count(int w, int h, float x1f, float y1f, float x2f, float y2f) {
  int64_t x1 = lrintf(x1f), y1 = lrintf(y1f), x2 = lrintf(x2f), y2 = lrintf(y2f);
  if (!clip_line(w, h, x1, y1, x2, y2)) return 0;
  const int64_t dx = std::llabs(x2 - x1), dy = std::llabs(y2 - y1);
  return (int)(std::max(dx, dy) + 1);
}

// optimizeAndMergeLines_lsd, uselongline.cpp:449-485 (segments only; KeyLine fields by orc_make_keylines)
int orc_merge_lines_lsd(const float* lines, int n, float* out, int cap) {
  std::vector<V4> src(n), t1, t2;
  for (int i = 0; i < n; ++i) std::memcpy(src[i].v, lines + 4 * i, 16);
  merge_lines(src, t1, 0.05f, 5.f, 15.f);
  filter_short(t1, 30.f);
  merge_lines(t1, t2, 0.03f, 3.f, 30.f);
  filter_short(t2, 50.f);
  for (size_t i = 0; i < t2.size() && (int)i < cap; ++i) std::memcpy(out + 4 * i, t2[i].v, 16);
  return (int)t2.size();
}

// checkLineExtremes of the LSD wrapper (LSDDetector_custom.cpp:112-138): what LINEextractor hands to the merge
void orc_clamp_segments(float* lines, int n, int w, int h) {
  for (int i = 0; i < n; ++i) {
    float* e = lines + 4 * i;
    if (e[0] < 0) e[0] = 0;
    if (e[0] >= w) e[0] = (float)w - 1.0f;
    if (e[2] < 0) e[2] = 0;
    if (e[2] >= w) e[2] = (float)w - 1.0f;
    if (e[1] < 0) e[1] = 0;
    if (e[1] >= h) e[1] = (float)h - 1.0f;
    if (e[3] < 0) e[3] = 0;
    if (e[3] >= h) e[3] = (float)h - 1.0f;
  }
}

// convertVec4fToKeyLine (uselongline.cpp:411-447) + top-N by response (LineExtractor.cpp:342-348)
int orc_make_keylines(const float* lines, int n, int w, int h, int nfeatures, psl_keyline* kl) {
  std::vector<psl_keyline> all(n);
  for (int i = 0; i < n; ++i) {
    const float* l = lines + 4 * i;
    psl_keyline& k = all[i];
    k.start_x = l[0]; k.start_y = l[1]; k.end_x = l[2]; k.end_y = l[3];
    k.s_oct_x = l[0]; k.s_oct_y = l[1]; k.e_oct_x = l[2]; k.e_oct_y = l[3];
    k.line_length = (float)sqrt(pow((double)(l[0] - l[2]), 2) + pow((double)(l[1] - l[3]), 2));
    k.angle = (float)atan2((double)(k.end_y - k.start_y), (double)(k.end_x - k.start_x));
    k.class_id = i;
    k.octave = 0;
    k.size = (k.end_x - k.start_x) * (k.end_y - k.start_y);
    k.pt_x = (k.end_x + k.start_x) / 2;
    k.pt_y = (k.end_y + k.start_y) / 2;
    k.response = k.line_length / std::max(w, h);
    k.num_pixels = orc_line_iterator_count(w, h, l[0], l[1], l[2], l[3]);
  }
  if (n > nfeatures) {
    std::stable_sort(all.begin(), all.end(), [](const psl_keyline& a, const psl_keyline& b) { return a.response > b.response; });
    all.resize(nfeatures);
    for (int i = 0; i < nfeatures; ++i) all[i].class_id = i;
  }
  for (size_t i = 0; i < all.size(); ++i) kl[i] = all[i];
  return (int)all.size();
}

// 3x3 Sobel to CV_16S with BORDER_REFLECT_101 on the 5x5 sigma-1 blurred image (:351-399)
void orc_lbd_gradients(const uint8_t* img, int w, int h, int stride, int16_t* dx, int16_t* dy) {
  std::vector<uint8_t> bl((size_t)w * h);
  orc_gauss_blur_u8(img, w, h, stride, bl.data(), w, 5);
  for (int y = 0; y < h; ++y) {
    const uint8_t* r0 = &bl[(size_t)reflect101(y - 1, h) * w];
    const uint8_t* r1 = &bl[(size_t)y * w];
    const uint8_t* r2 = &bl[(size_t)reflect101(y + 1, h) * w];
    for (int x = 0; x < w; ++x) {
      const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
      dx[(size_t)y * w + x] = (int16_t)((r0[xp] - r0[xm]) + 2 * (r1[xp] - r1[xm]) + (r2[xp] - r2[xm]));
      dy[(size_t)y * w + x] = (int16_t)((r2[xm] + 2 * r2[x] + r2[xp]) - (r0[xm] + 2 * r0[x] + r0[xp]));
    }
  }
}

// computeLBD for one line (:1074-1330): 72 floats
void orc_lbd_one(const int16_t* dxi, const int16_t* dyi, int w, int h, const psl_keyline* kl, float* des) {
  const int widthOfBand = 7, nBands = 9;
  static double gaussL[21], gaussG[63];
  static bool init = false;
  if (!init) {  // BinaryDescriptor ctor (:219-261): integer divisions are the reference's
    double u = (widthOfBand * 3 - 1) / 2, sigma = (widthOfBand * 2 + 1) / 2, inv = -1 / (2 * sigma * sigma);
    for (int i = 0; i < 21; ++i) gaussL[i] = exp((i - u) * (i - u) * inv);
    u = (nBands * widthOfBand - 1) / 2;
    sigma = u;
    inv = -1 / (2 * sigma * sigma);
    for (int i = 0; i < 63; ++i) gaussG[i] = exp((i - u) * (i - u) * inv);
    init = true;
  }
  const short heightOfLSP = (short)(widthOfBand * nBands);
  float pL[9] = {0}, nL[9] = {0}, pL2[9] = {0}, nL2[9] = {0}, pO[9] = {0}, nO[9] = {0}, pO2[9] = {0}, nO2[9] = {0};
  const short lengthOfLSP = (short)kl->num_pixels;
  const short halfHeight = (heightOfLSP - 1) / 2, halfWidth = (lengthOfLSP - 1) / 2;
  const short realWidth = (short)w, imageWidth = realWidth - 1, imageHeight = (short)(h - 1);
  const float midX = (float)(0.5 * (kl->s_oct_x + kl->e_oct_x)), midY = (float)(0.5 * (kl->s_oct_y + kl->e_oct_y));
  float dL[2], dO[2];
  dL[0] = (float)cos((double)kl->angle);
  dL[1] = (float)sin((double)kl->angle);
  dO[0] = -dL[1];
  dO[1] = dL[0];
  float sCorX0 = -dL[0] * halfWidth + dL[1] * halfHeight + midX;
  float sCorY0 = -dL[1] * halfWidth - dL[0] * halfHeight + midY;
  for (short hID = 0; hID < heightOfLSP; ++hID) {
    float sCorX = sCorX0, sCorY = sCorY0;
    float pgdL = 0, ngdL = 0, pgdO = 0, ngdO = 0;
    for (short wID = 0; wID < lengthOfLSP; ++wID) {
      short t = (short)round(sCorX);
      const short xCor = (t < 0) ? 0 : (t > imageWidth) ? imageWidth : t;
      t = (short)round(sCorY);
      const short yCor = (t < 0) ? 0 : (t > imageHeight) ? imageHeight : t;
      const short gx = dxi[yCor * realWidth + xCor], gy = dyi[yCor * realWidth + xCor];
      const float gDL = gx * dL[0] + gy * dL[1], gDO = gx * dO[0] + gy * dO[1];
      if (gDL > 0) pgdL += gDL; else ngdL -= gDL;
      if (gDO > 0) pgdO += gDO; else ngdO -= gDO;
      sCorX += dL[0];
      sCorY += dL[1];
    }
    sCorX0 -= dL[1];
    sCorY0 += dL[0];
    float coef = (float)gaussG[hID];
    pgdL = coef * pgdL; ngdL = coef * ngdL;
    const float pgdL2 = pgdL * pgdL, ngdL2 = ngdL * ngdL;
    pgdO = coef * pgdO; ngdO = coef * ngdO;
    const float pgdO2 = pgdO * pgdO, ngdO2 = ngdO * ngdO;
    auto add = [&](int band, float c) {
      pL[band] += c * pgdL; nL[band] += c * ngdL;
      pL2[band] += c * c * pgdL2; nL2[band] += c * c * ngdL2;
      pO[band] += c * pgdO; nO[band] += c * ngdO;
      pO2[band] += c * c * pgdO2; nO2[band] += c * c * ngdO2;
    };
    short band = (short)(hID / widthOfBand);
    add(band, (float)gaussL[hID % widthOfBand + widthOfBand]);
    band--;
    if (band >= 0) add(band, (float)gaussL[hID % widthOfBand + 2 * widthOfBand]);
    band = band + 2;
    if (band < nBands) add(band, (float)gaussL[hID % widthOfBand]);
  }
  const float invN2 = (float)(1.0 / (widthOfBand * 2.0)), invN3 = (float)(1.0 / (widthOfBand * 3.0));
  for (int b = 0; b < nBands; ++b) {
    const float invN = (b == 0 || b == nBands - 1) ? invN2 : invN3;
    float* d = des + b * 8;
    float t = pL[b] * invN; d[0] = t; d[4] = (float)sqrt((double)(pL2[b] * invN - t * t));
    t = nL[b] * invN; d[1] = t; d[5] = (float)sqrt((double)(nL2[b] * invN - t * t));
    t = pO[b] * invN; d[2] = t; d[6] = (float)sqrt((double)(pO2[b] * invN - t * t));
    t = nO[b] * invN; d[3] = t; d[7] = (float)sqrt((double)(nO2[b] * invN - t * t));
  }
  float tempM = 0, tempS = 0;
  for (int b = 0; b < nBands; ++b) {
    const float* d = des + 8 * b;
    tempM += d[0] * d[0]; tempM += d[1] * d[1]; tempM += d[2] * d[2]; tempM += d[3] * d[3];
    tempS += d[4] * d[4]; tempS += d[5] * d[5]; tempS += d[6] * d[6]; tempS += d[7] * d[7];
  }
  tempM = (float)(1 / sqrt((double)tempM));
  tempS = (float)(1 / sqrt((double)tempS));
  for (int b = 0; b < nBands; ++b) {
    float* d = des + 8 * b;
    d[0] *= tempM; d[1] *= tempM; d[2] *= tempM; d[3] *= tempM;
    d[4] *= tempS; d[5] *= tempS; d[6] *= tempS; d[7] *= tempS;
  }
  for (int i = 0; i < 72; ++i)
    if (des[i] > 0.4) des[i] = (float)0.4;
  float temp = 0;
  for (int i = 0; i < 72; ++i) temp += des[i] * des[i];
  temp = (float)(1 / sqrt((double)temp));
  for (int i = 0; i < 72; ++i) des[i] = des[i] * temp;
}

// binaryConversion over the 32 band pairs (:402-413, :655-668)
void orc_lbd_binarise(const float* des72, uint8_t* out32) {
  for (int c = 0; c < 32; ++c) {
    const float* f1 = des72 + 8 * combinations[c][0];
    const float* f2 = des72 + 8 * combinations[c][1];
    unsigned r = 0;
    for (int i = 0; i < 8; ++i)
      if (f1[i] > f2[i]) r += 1u << i;
    out32[c] = (uint8_t)r;
  }
}

// LINEextractor::operator(), LineExtractor.cpp:325-366.  lineeq: n x 3 doubles, lbd72 optional (n x 72).
int orc_line_extract(const uint8_t* gray, int w, int h, int stride, int nfeatures, psl_keyline* kl, uint8_t* ldesc,
                     double* lineeq, float* lbd72, int cap, int* n_out) {
  *n_out = 0;
  if (w <= 0 || h <= 0) return 0;
  const int kCap = 65536;
  std::vector<float> raw(4 * (size_t)kCap), merged(4 * (size_t)kCap);
  int nr = orc_lsd_detect(gray, w, h, stride, 1, raw.data(), kCap);
  if (nr > kCap) return PSL_E_CAPACITY;
  orc_clamp_segments(raw.data(), nr, w, h);
  int nm = orc_merge_lines_lsd(raw.data(), nr, merged.data(), kCap);
  std::vector<psl_keyline> k(std::max(nm, 1));
  int n = orc_make_keylines(merged.data(), nm, w, h, nfeatures, k.data());
  if (n > cap) return PSL_E_CAPACITY;
  if (n > 0) {
    std::vector<int16_t> dx((size_t)w * h), dy((size_t)w * h);
    orc_lbd_gradients(gray, w, h, stride, dx.data(), dy.data());
    float des[72];
    for (int i = 0; i < n; ++i) {
      kl[i] = k[i];
      orc_lbd_one(dx.data(), dy.data(), w, h, &k[i], des);
      if (lbd72) std::memcpy(lbd72 + 72 * (size_t)i, des, sizeof(des));
      orc_lbd_binarise(des, ldesc + 32 * (size_t)i);
      // line equation: sp x ep normalised by its first two components (:352-363)
      const double sx = k[i].start_x, sy = k[i].start_y, ex = k[i].end_x, ey = k[i].end_y;
      double l0 = sy * 1.0 - 1.0 * ey, l1 = 1.0 * ex - sx * 1.0, l2 = sx * ey - sy * ex;
      const double nrm = sqrt(l0 * l0 + l1 * l1);
      lineeq[3 * i] = l0 / nrm; lineeq[3 * i + 1] = l1 / nrm; lineeq[3 * i + 2] = l2 / nrm;
    }
  }
  *n_out = n;
  return 0;
}

}  // extern "C"
