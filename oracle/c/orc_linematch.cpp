// TEST INFRASTRUCTURE — CPU oracle for the line matchers (see psl_oracle.h): a sequential restatement of
// add_src/LSDmatcher.cpp (match/matchNNR :354-413, SearchByGeomNApearance :36-110, SearchByProjection :112-215 and
// :260-352, FrameBFMatch :492-516, SearchDouble :462-490, lineDescriptorMAD :660-685), the Frame helpers they call
// (src/Frame.cc:286-309 AssignFeaturesToGridForLine, :752-826 GetFeaturesInAreaForLine, add_src/lineIterator.cpp:33-77)
// and the structural-line association (add_src/InsectlineMatch.cpp:9-60, src/Map.cc:204-272) on plain arrays.
// cv::BFMatcher::knnMatch is orc_hamming_knn2 (pinned against cv2.BFMatcher, tests/test_oracle_match.py).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <unordered_set>
#include <vector>

#include "psl_oracle.h"

namespace {

int desc_dist(const uint8_t* a, const uint8_t* b) {  // LSDmatcher::DescriptorDistance, LSDmatcher.cpp:687-703
  int d = 0;
  orc_descriptor_distance(a, b, 1, &d);
  return d;
}

// ORB_SLAM2::LineIterator (add_src/lineIterator.cpp:33-77): Bresenham over grid coordinates
struct LineIt {
  double x1, y1, x2, y2, dx, dy, error;
  bool steep;
  int ystep, x, y, maxX;
  LineIt(double x1_, double y1_, double x2_, double y2_)
      : x1(x1_), y1(y1_), x2(x2_), y2(y2_), steep(std::abs(y2_ - y1_) > std::abs(x2_ - x1_)) {
    if (steep) { std::swap(x1, y1); std::swap(x2, y2); }
    if (x1 > x2) { std::swap(x1, x2); std::swap(y1, y2); }
    dx = x2 - x1;
    dy = std::abs(y2 - y1);
    error = dx / 2.0;
    ystep = (y1 < y2) ? 1 : -1;
    x = static_cast<int>(x1);
    y = static_cast<int>(y1);
    maxX = static_cast<int>(x2);
  }
  bool next(int& px, int& py) {
    if (x > maxX) return false;
    if (steep) { px = y; py = x; } else { px = x; py = y; }
    error -= dy;
    if (error < 0) { y += ystep; error += dx; }
    x++;
    return true;
  }
};

struct LineGrid {  // Frame::AssignFeaturesToGridForLine, Frame.cc:286-309
  std::vector<int> cell[PSL_GRID_COLS][PSL_GRID_ROWS];
  explicit LineGrid(const psl_line_frame_view& f) {
    for (int i = 0; i < f.n; ++i) {
      const psl_keyline& kl = f.kl_un[i];
      LineIt it(kl.start_x * f.grid_w_inv, kl.start_y * f.grid_h_inv, kl.end_x * f.grid_w_inv, kl.end_y * f.grid_h_inv);
      int px, py;
      while (it.next(px, py))
        if (px >= 0 && px < PSL_GRID_COLS && py >= 0 && py < PSL_GRID_ROWS) cell[px][py].push_back(i);
    }
  }
};

// Frame::GetFeaturesInAreaForLine, Frame.cc:752-826 (minLevel / maxLevel are ignored by the reference)
void lines_in_area(const psl_line_frame_view& f, const LineGrid& g, float x1, float y1, float x2, float y2, float r,
                   float TH, std::vector<int>& out) {
  out.clear();
  std::unordered_set<int> seen;
  const float x[3] = {x1, (float)((x1 + x2) / 2.0), x2};
  const float y[3] = {y1, (float)((y1 + y2) / 2.0), y2};
  float d1x = x1 - x2, d1y = y1 - y2;
  const float n1 = (float)sqrt((double)(d1x * d1x + d1y * d1y));
  d1x /= n1;
  d1y /= n1;
  for (int i = 0; i < 3; ++i) {
    const int nMinCellX = std::max(0, (int)floor((double)((x[i] - f.min_x - r) * f.grid_w_inv)));
    if (nMinCellX >= PSL_GRID_COLS) continue;
    const int nMaxCellX = std::min(PSL_GRID_COLS - 1, (int)ceil((double)((x[i] - f.min_x + r) * f.grid_w_inv)));
    if (nMaxCellX < 0) continue;
    const int nMinCellY = std::max(0, (int)floor((double)((y[i] - f.min_y - r) * f.grid_h_inv)));
    if (nMinCellY >= PSL_GRID_ROWS) continue;
    const int nMaxCellY = std::min(PSL_GRID_ROWS - 1, (int)ceil((double)((y[i] - f.min_y + r) * f.grid_h_inv)));
    if (nMaxCellY < 0) continue;
    for (int ix = nMinCellX; ix <= nMaxCellX; ++ix)
      for (int iy = nMinCellY; iy <= nMaxCellY; ++iy)
        for (int j : g.cell[ix][iy]) {
          if (seen.count(j)) continue;
          const psl_keyline& kl = f.kl_un[j];
          float d2x = kl.start_x - kl.end_x, d2y = kl.start_y - kl.end_y;
          const float n2 = (float)sqrt((double)(d2x * d2x + d2y * d2y));
          d2x /= n2;
          d2y /= n2;
          const float cs = std::abs(d1x * d2x + d1y * d2y);
          if (cs < TH) continue;
          const double* L = f.lineeq + 3 * j;
          const float dist = (float)(L[0] * x[i] + L[1] * y[i] + L[2]);
          if (fabs(dist) < r) {
            out.push_back(j);
            seen.insert(j);
          }
        }
  }
}

// LSDmatcher::computeAngle2D, LSDmatcher.cpp:19-34 on (e - s) float differences stored as doubles
double angle2d(double ax, double ay, double bx, double by) {
  const double dot = ax * bx + ay * by;
  const double ma = std::sqrt(ax * ax + ay * ay), mb = std::sqrt(bx * bx + by * by);
  return std::abs(dot / (ma * mb));
}

// lineDescriptorMAD, LSDmatcher.cpp:660-685 (only the two medians are used by the caller)
void mad(const std::vector<float>& d0, const std::vector<float>& d1, double& nn_mad, double& nn12_mad) {
  const size_t n = d0.size();
  std::vector<float> a(d0);
  std::sort(a.begin(), a.end());
  const double med = a[n / 2];
  for (size_t i = 0; i < n; ++i) a[i] = fabsf((float)(d0[i] - med));
  std::sort(a.begin(), a.end());
  nn_mad = 1.4826 * a[n / 2];
  std::vector<float> g(n);
  for (size_t i = 0; i < n; ++i) g[i] = d1[i] - d0[i];
  std::sort(g.begin(), g.end());
  const double med12 = g[n / 2];
  for (size_t i = 0; i < n; ++i) a[i] = fabsf((float)(d1[i] - d0[i] - med12));
  std::sort(a.begin(), a.end());
  nn12_mad = 1.4826 * a[n / 2];
}

}  // namespace

extern "C" {

int orc_line_grid_cells(const psl_line_frame_view* f, int line, int32_t* cells, int cap) {
  const psl_keyline& kl = f->kl_un[line];
  LineIt it(kl.start_x * f->grid_w_inv, kl.start_y * f->grid_h_inv, kl.end_x * f->grid_w_inv, kl.end_y * f->grid_h_inv);
  int px, py, n = 0;
  while (it.next(px, py))
    if (px >= 0 && px < PSL_GRID_COLS && py >= 0 && py < PSL_GRID_ROWS) {
      if (n < cap) cells[n] = px * PSL_GRID_ROWS + py;
      ++n;
    }
  return n;
}

int orc_lines_in_area(const psl_line_frame_view* f, float x1, float y1, float x2, float y2, float r, float TH,
                      int32_t* out, int cap) {
  LineGrid g(*f);
  std::vector<int> v;
  lines_in_area(*f, g, x1, y1, x2, y2, r, TH, v);
  for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = v[i];
  return (int)v.size();
}

int orc_line_match_nnr(const uint8_t* desc1, int n1, const uint8_t* desc2, int n2, float nnr, int32_t* matches12) {
  int matches = 0;
  for (int i = 0; i < n1; ++i) matches12[i] = -1;
  if (n1 <= 0 || n2 < 2) return 0;
  std::vector<int32_t> idx(2 * (size_t)n1), dist(2 * (size_t)n1);
  orc_hamming_knn2(desc1, n1, desc2, n2, idx.data(), dist.data());
  for (int i = 0; i < n1; ++i)
    if ((float)dist[2 * i] < (float)dist[2 * i + 1] * nnr) {
      matches12[i] = idx[2 * i];
      ++matches;
    }
  return matches;
}

int orc_line_search_geom(const psl_keyline* kl_last, const uint8_t* desc_last, const uint8_t* has_ml, int n_last,
                         const psl_keyline* kl_cur, const uint8_t* desc_cur, int n_cur, const float* bounds,
                         float desc_th, int32_t* assign_cur) {
  for (int i = 0; i < n_cur; ++i) assign_cur[i] = -1;
  if (n_cur == 0) return 0;  // CurrentFrame.mLdesc.empty()
  std::vector<int32_t> m12(std::max(n_last, 1));
  orc_line_match_nnr(desc_last, n_last, desc_cur, n_cur, desc_th, m12.data());
  const double deltaWidth = (bounds[2] - bounds[0]) * 0.1, deltaHeight = (bounds[3] - bounds[1]) * 0.1;
  const double cos_th = std::cos(20.0 / 180.0 * M_PI);
  int lmatches = 0;
  for (int i1 = 0; i1 < n_last; ++i1) {
    if (!has_ml[i1]) continue;
    const int i2 = m12[i1];
    if (i2 < 0) continue;
    const psl_keyline& c = kl_cur[i2];
    const psl_keyline& l = kl_last[i1];
    if (c.start_x == 0) continue;
    const double ang = angle2d((double)(c.e_oct_x - c.s_oct_x), (double)(c.e_oct_y - c.s_oct_y),
                               (double)(l.e_oct_x - l.s_oct_x), (double)(l.e_oct_y - l.s_oct_y));
    if (ang < cos_th) continue;
    if ((fabs(c.s_oct_x - l.s_oct_x) > deltaWidth || fabs(c.s_oct_y - l.s_oct_y) > deltaHeight) &&
        (fabs(c.e_oct_x - l.e_oct_x) > deltaWidth || fabs(c.e_oct_y - l.e_oct_y) > deltaHeight))
      continue;
    assign_cur[i2] = i1;
    ++lmatches;
  }
  return lmatches;
}

void orc_line_frame_bf_match(const uint8_t* desc1, int n1, const uint8_t* desc2, int n2, float nn_ratio, float th,
                             int32_t* matches) {
  for (int i = 0; i < n1; ++i) matches[i] = -1;
  if (n1 <= 0 || n2 < 2) return;
  std::vector<int32_t> idx(2 * (size_t)n1), dist(2 * (size_t)n1);
  orc_hamming_knn2(desc1, n1, desc2, n2, idx.data(), dist.data());
  std::vector<float> d0(n1), d1(n1);
  for (int i = 0; i < n1; ++i) { d0[i] = (float)dist[2 * i]; d1[i] = (float)dist[2 * i + 1]; }
  double nn_th, nn12_th;
  mad(d0, d1, nn_th, nn12_th);
  nn12_th = nn12_th * 0.5;
  for (int i = 0; i < n1; ++i) {
    const double dist_12 = d1[i] - d0[i];
    if (dist_12 > nn12_th && d0[i] < th && d0[i] < nn_ratio * d1[i]) matches[i] = idx[2 * i];
  }
}

int orc_line_search_double(const uint8_t* desc1, int n1, const uint8_t* desc2, int n2, float nn_ratio, float th,
                           int32_t* matches12) {
  for (int i = 0; i < n1; ++i) matches12[i] = -1;
  if (n1 == 0 || n2 == 0) return 0;
  std::vector<int32_t> m21(n2);
  orc_line_frame_bf_match(desc1, n1, desc2, n2, nn_ratio, th, matches12);
  orc_line_frame_bf_match(desc2, n2, desc1, n1, nn_ratio, th, m21.data());
  int nm = 0;
  for (int i = 0; i < n1; ++i) {
    const int j = matches12[i];
    if (j >= 0) {
      if (m21[j] != i) matches12[i] = -1;
      else ++nm;
    }
  }
  return nm;
}

// LSDmatcher::SearchForTriangulation, LSDmatcher.cpp:705-779 (pairs form: th = TH_LOW, is_double = 1)
int orc_line_search_triangulation(const uint8_t* desc1, const uint8_t* has_ml1, int n1, const uint8_t* desc2,
                                  const uint8_t* has_ml2, int n2, float nn_ratio, float th, int is_double,
                                  int32_t* matches12) {
  for (int i = 0; i < n1; ++i) matches12[i] = -1;
  if (n1 == 0 || n2 == 0) return 0;
  std::vector<int32_t> m12(n1), m21(n2);
  orc_line_frame_bf_match(desc1, n1, desc2, n2, nn_ratio, th, m12.data());
  orc_line_frame_bf_match(desc2, n2, desc1, n1, nn_ratio, th, m21.data());
  int nm = 0;
  for (int i = 0; i < n1; ++i) {
    const int j = m12[i];
    if (j < 0) continue;
    if (is_double && m21[j] != i) continue;
    if (has_ml1[i] || has_ml2[j]) continue;
    matches12[i] = j;
    ++nm;
  }
  return nm;
}

// ---- the epipolar-overlap variant: LSDmatcher::FrameBFMatchNew / mutualOverlap / SearchForTriangulationNew
// (LSDmatcher.cpp:518-658, 783-824; nothing in the reference calls it) ----
namespace {
struct V3f { float v[3]; };

V3f matvec3(const float* F, float x, float y) {  // cv::Mat(3x3 CV_32F) * (x, y, 1): double accumulation, one rounding
  V3f r;
  for (int k = 0; k < 3; ++k) r.v[k] = (float)((double)F[3 * k] * x + (double)F[3 * k + 1] * y + (double)F[3 * k + 2] * 1.0);
  return r;
}
V3f cross3(const V3f& a, const V3f& b) {  // Mat::cross, CV_32F
  return V3f{{a.v[1] * b.v[2] - a.v[2] * b.v[1], a.v[2] * b.v[0] - a.v[0] * b.v[2], a.v[0] * b.v[1] - a.v[1] * b.v[0]}};
}
double norm_diff(const V3f& a, const V3f& b) {  // cv::norm(a - b): float differences, squares summed in double
  const float d0 = a.v[0] - b.v[0], d1 = a.v[1] - b.v[1], d2 = a.v[2] - b.v[2];
  return std::sqrt((double)d0 * d0 + (double)d1 * d1 + (double)d2 * d2);
}
float mutual_overlap(const V3f* p) {  // :583-658
  float max_dist = 0.0f;
  int outer1 = 0, outer2 = 3;
  for (int i = 0; i < 3; ++i)
    for (int j = i + 1; j < 4; ++j) {
      const float dist = (float)norm_diff(p[i], p[j]);
      if (dist > max_dist) { max_dist = dist; outer1 = i; outer2 = j; }
    }
  if (max_dist < 1.0f) return 0.0f;
  int inner[2], c = 0;
  for (int k = 0; k < 4; ++k)
    if (k != outer1 && k != outer2) inner[c++] = k;
  return (float)(norm_diff(p[inner[0]], p[inner[1]]) / max_dist);  // double norm / float
}
void frame_bf_match_new(const uint8_t* d1, const psl_keyline* kl1, int n1, const uint8_t* d2, const psl_keyline* kl2,
                        const double* func2, int n2, const float* F, float th, float nn_ratio, int32_t* out) {  // :518-581
  for (int i = 0; i < n1; ++i) out[i] = -1;
  if (n1 == 0 || n2 < 2) return;  // with one train row knnMatch returns one entry and the loop `j < size() - 1` is empty
  std::vector<int32_t> idx((size_t)n1 * 2), dist((size_t)n1 * 2);
  orc_hamming_knn2(d1, n1, d2, n2, idx.data(), dist.data());
  for (int q = 0; q < n1; ++q) {
    const int t = idx[2 * q];
    const V3f e1 = matvec3(F, kl1[q].start_x, kl1[q].start_y), e2 = matvec3(F, kl1[q].end_x, kl1[q].end_y);
    const V3f l2{{(float)func2[3 * t], (float)func2[3 * t + 1], (float)func2[3 * t + 2]}};
    V3f p1 = cross3(l2, e1), p2 = cross3(l2, e2);
    if (!(std::fabs(p1.v[2]) > 1e-12 && std::fabs(p2.v[2]) > 1e-12)) continue;
    const float s1 = (float)(1.0 / (double)p1.v[2]), s2 = (float)(1.0 / (double)p2.v[2]);  // Mat /= s -> convertTo(alpha = 1/s)
    for (int k = 0; k < 3; ++k) { p1.v[k] = p1.v[k] * s1; p2.v[k] = p2.v[k] * s2; }
    const V3f pts[4] = {p1, p2, V3f{{kl2[t].start_x, kl2[t].start_y, 1.0f}}, V3f{{kl2[t].end_x, kl2[t].end_y, 1.0f}}};
    const float score = mutual_overlap(pts);
    const float d0 = (float)dist[2 * q], dd1 = (float)dist[2 * q + 1];
    if (d0 < th && score > 0.8 && d0 < nn_ratio * dd1) out[q] = t;
  }
}
}  // namespace

int orc_line_search_triangulation_new(const psl_keyline* kl1, const uint8_t* desc1, const double* func1,
                                      const uint8_t* has_ml1, int n1, const psl_keyline* kl2, const uint8_t* desc2,
                                      const double* func2, const uint8_t* has_ml2, int n2, const float* F21,
                                      const float* F12, float nn_ratio, float th, int is_double, int32_t* pairs) {
  for (int i = 0; i < n1; ++i) pairs[i] = -1;
  if (n1 == 0 || n2 == 0) return 0;
  std::vector<int32_t> m12((size_t)n1), m21((size_t)n2);
  frame_bf_match_new(desc1, kl1, n1, desc2, kl2, func2, n2, F21, th, nn_ratio, m12.data());
  frame_bf_match_new(desc2, kl2, n2, desc1, kl1, func1, n1, F12, th, nn_ratio, m21.data());
  int nm = 0;
  for (int i = 0; i < n1; ++i) {
    const int j = m12[i];
    if (j < 0) continue;
    if (is_double && m21[j] != i) continue;
    if (has_ml1[i] || has_ml2[j]) continue;
    pairs[i] = j;
    ++nm;
  }
  return nm;
}

// KeyFrame::GetLinesInArea, KeyFrame.cc:857-891 (TH defaults to 0.998, KeyFrame.h:144): all KeyLines, index order.
static std::vector<int> lines_in_area(const psl_keyline* kl, int n, float x1, float y1, float x2, float y2, float r,
                                      float TH) {
  std::vector<int> out;
  float delta1x = x1 - x2, delta1y = y1 - y2;                                   // :863-864
  const float norm_delta1 = sqrtf(delta1x * delta1x + delta1y * delta1y);        // :865
  delta1x /= norm_delta1;                                                       // :866-867
  delta1y /= norm_delta1;
  for (int i = 0; i < n; ++i) {
    const psl_keyline& k = kl[i];
    // :873 the 0.5 literal makes the offsets and their squares double; the sum is stored in a float
    const float distance = (float)((0.5 * (x1 + x2) - k.pt_x) * (0.5 * (x1 + x2) - k.pt_x) +
                                   (0.5 * (y1 + y2) - k.pt_y) * (0.5 * (y1 + y2) - k.pt_y));
    if (distance > r * r) continue;                                             // :874-875
    float delta2x = k.start_x - k.end_x, delta2y = k.start_y - k.end_y;          // :877-878
    const float norm_delta2 = sqrtf(delta2x * delta2x + delta2y * delta2y);
    delta2x /= norm_delta2;
    delta2y /= norm_delta2;
    const float CosSita = fabsf(delta1x * delta2x + delta1y * delta2y);          // :882
    if (CosSita < TH) continue;                                                 // :884-885 (NaN passes)
    out.push_back(i);
  }
  return out;
}

// The window search of LSDmatcher::Fuse, LSDmatcher.cpp:916-953 (the caller did :862-914).  The descriptor rows it
// compares against are pKF->mDescriptors.row(idx) (:938), passed as kf_desc.
int orc_line_fuse(const psl_keyline* kl, int n_lines, const uint8_t* kf_desc, const psl_line_fuse_query* qs,
                  const uint8_t* qdesc, int nq, float th_cos, int th_low, int32_t* best_idx, int32_t* best_dist) {
  int fused = 0;
  for (int q = 0; q < nq; ++q) {
    best_idx[q] = -1;
    if (best_dist) best_dist[q] = 256;
    const psl_line_fuse_query& Q = qs[q];
    if (!(Q.flags & PSL_Q_VALID)) continue;
    const std::vector<int> cand = lines_in_area(kl, n_lines, Q.u1, Q.v1, Q.u2, Q.v2, Q.radius, th_cos);
    int bestDist = 256, bestIdx = -1;                                           // :926-927
    for (int idx : cand) {
      const int kpLevel = kl[idx].octave;
      if (kpLevel < Q.pred_level - 1 || kpLevel > Q.pred_level) continue;        // :936-937
      const int dist = desc_dist(qdesc + 32 * (size_t)q, kf_desc + 32 * (size_t)idx);
      if (dist < bestDist) {                                                    // :945-949
        bestDist = dist;
        bestIdx = idx;
      }
    }
    if (best_dist) best_dist[q] = bestDist;
    if (bestDist <= th_low && bestIdx >= 0) {                                    // :952
      best_idx[q] = bestIdx;
      ++fused;
    }
  }
  return fused;
}

int orc_line_match_projection(const psl_line_frame_view* f, const psl_line_query* qs, const uint8_t* qdesc, int nq,
                              const uint8_t* claimed_in, int mode, float nn_ratio, int32_t* assign) {
  LineGrid g(*f);
  std::vector<uint8_t> claimed(std::max(f->n, 1), 0);
  for (int i = 0; i < f->n; ++i) {
    assign[i] = -1;
    if (claimed_in) claimed[i] = claimed_in[i];
  }
  const double cos10 = std::cos(10.0 / 180.0 * M_PI), cos15 = std::cos(15.0 / 180.0 * M_PI);
  int nmatches = 0;
  std::vector<int> cand;
  for (int q = 0; q < nq; ++q) {
    const psl_line_query& Q = qs[q];
    if (!(Q.flags & PSL_Q_VALID)) continue;
    lines_in_area(*f, g, Q.x1, Q.y1, Q.x2, Q.y2, Q.radius, mode == 0 ? 0.96f : 0.998f, cand);
    if (cand.empty()) continue;
    const uint8_t* dq = qdesc + 32 * (size_t)q;
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    for (int i2 : cand) {
      if (claimed[i2]) continue;
      const psl_keyline& c = f->kl_un[i2];
      if (mode == 0) {
        const double ang = angle2d((double)(c.e_oct_x - c.s_oct_x), (double)(c.e_oct_y - c.s_oct_y),
                                   (double)(Q.ex - Q.sx), (double)(Q.ey - Q.sy));
        if (ang < cos10) continue;
        const int dist = desc_dist(dq, f->ldesc + 32 * (size_t)i2);
        const float mx = std::max(Q.length, c.line_length), mn = std::min(Q.length, c.line_length);
        if (mn / mx < 0.75) continue;
        if (dist < bestDist) { bestDist = dist; bestIdx = i2; }
      } else {
        const double* p = f->lines3d + 6 * (size_t)i2;
        const double vx = p[0] - p[3], vy = p[1] - p[4], vz = p[2] - p[5];
        const float dot = (float)(vx * Q.normal[0] + vy * Q.normal[1] + vz * Q.normal[2]);
        const float mag_f = (float)std::sqrt(vx * vx + vy * vy + vz * vz);
        const float mag_ml = (float)std::sqrt(Q.normal[0] * Q.normal[0] + Q.normal[1] * Q.normal[1] + Q.normal[2] * Q.normal[2]);
        const float angle = std::abs(dot / (mag_f * mag_ml));
        if (angle < cos15) continue;
        const int dist = desc_dist(dq, f->ldesc + 32 * (size_t)i2);
        if (dist < bestDist) {
          bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = c.octave; bestIdx = i2;
        } else if (dist < bestDist2) {
          bestLevel2 = c.octave; bestDist2 = dist;
        }
      }
    }
    if (bestDist <= 95) {
      if (mode == 1 && bestLevel == bestLevel2 && bestDist > nn_ratio * bestDist2) continue;
      assign[bestIdx] = q;
      if (Q.flags & PSL_Q_CLAIMS) claimed[bestIdx] = 1;
      ++nmatches;
    }
  }
  return nmatches;
}

// The plane hypotheses of Frame::ExtractLSD, Frame.cc:512-645 (+ Frame::OldPlane :474-487).  Returns the number of
// hypotheses (which may exceed cap; only the first cap are stored).
int orc_plane_hypotheses(const psl_keyline* kl_un, const float* line_eq, const double* lines3d, int n_lines,
                         const psl_line_junction* js, int nj, double* le_l, float* planes, double* normals,
                         int32_t* junction_of, int cap) {
  (void)n_lines;
  std::vector<float> kept;  // mvPlanes, 4 floats each (all of them, also beyond cap: OldPlane reads them)
  int np = 0;
  auto is_zero3 = [](const double* v) {  // Eigen isZero(): |x| <= 1e-12 for every coefficient
    return std::fabs(v[0]) <= 1e-12 && std::fabs(v[1]) <= 1e-12 && std::fabs(v[2]) <= 1e-12;
  };
  for (int i = 0; i < nj; ++i) {
    const int l1 = js[i].l1, l2 = js[i].l2;
    for (int s = 0; s < 2; ++s) {  // :522-528 le_l = sp x ep / sqrt(le0^2 + le1^2), homogeneous points (x, y, 1)
      const psl_keyline& k = kl_un[s == 0 ? l1 : l2];
      const double ax = k.start_x, ay = k.start_y, bx = k.end_x, by = k.end_y;
      const double c0 = ay * 1.0 - 1.0 * by, c1 = 1.0 * bx - ax * 1.0, c2 = ax * by - ay * bx;
      const double nrm = std::sqrt(c0 * c0 + c1 * c1);
      le_l[6 * i + 3 * s] = c0 / nrm;
      le_l[6 * i + 3 * s + 1] = c1 / nrm;
      le_l[6 * i + 3 * s + 2] = c2 / nrm;
    }
    const float* e1 = line_eq + 3 * l1;
    const float* e2 = line_eq + 3 * l2;
    if (e1[0] == 0 && e1[1] == 0 && e1[2] == 0) continue;                     // :531-534
    if (e2[0] == 0 && e2[1] == 0 && e2[2] == 0) continue;
    const double* L1 = lines3d + 6 * l1;
    const double* L2 = lines3d + 6 * l2;
    if (is_zero3(L1) && is_zero3(L1 + 3)) continue;                            // :537-540
    if (is_zero3(L2) && is_zero3(L2 + 3)) continue;
    float pn[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};  // :554
    const float norm = sqrtf(pn[0] * pn[0] + pn[1] * pn[1] + pn[2] * pn[2]);  // :572-573
    pn[0] = pn[0] / norm;
    pn[1] = pn[1] / norm;
    pn[2] = pn[2] / norm;
    double n_[3] = {pn[0], pn[1], pn[2]};                                      // :588
    const double* P[5] = {L1, L1 + 3, L2, L2 + 3, js[i].cross3d};
    float d[5];
    for (int k = 0; k < 5; ++k) d[k] = (float)(n_[0] * P[k][0] + n_[1] * P[k][1] + n_[2] * P[k][2]);  // :597-602
    float dmin = 10000, dmax = -10000;                                         // :609-622
    for (int k = 0; k < 5; ++k) {
      dmin = dmin < d[k] ? dmin : d[k];
      dmax = dmax > d[k] ? dmax : d[k];
    }
    if (dmax - dmin > 0.05) continue;                                          // :628
    const float planeDis = -(d[0] + d[1] + d[2] + d[3] + d[4]) / 5;            // :632
    float pl[4] = {(float)n_[0], (float)n_[1], (float)n_[2], planeDis};
    if (pl[3] < 0) {                                                           // :641-645
      for (int k = 0; k < 4; ++k) pl[k] = -pl[k];
      for (int k = 0; k < 3; ++k) n_[k] = -n_[k];
    }
    bool old = false;                                                          // OldPlane :474-487
    for (size_t m = 0; m < kept.size(); m += 4) {
      const float dd = pl[3] - kept[m + 3];
      const float angle = pl[0] * kept[m] + pl[1] * kept[m + 1] + pl[2] * kept[m + 2];
      if (dd > 0.2 || dd < -0.2) continue;
      if (angle < 0.9397 && angle > -0.9397) continue;
      old = true;
      break;
    }
    if (old) continue;
    kept.insert(kept.end(), pl, pl + 4);
    if (np < cap) {
      for (int k = 0; k < 4; ++k) planes[4 * np + k] = pl[k];
      for (int k = 0; k < 3; ++k) normals[3 * np + k] = n_[k];
      junction_of[np] = i;
    }
    ++np;
  }
  return np;
}

int orc_plane_assoc(const float* planes_cam, const double* pts, int n_ljl, const float* Tcw, const float* map_planes,
                    const uint8_t* map_bad, int n_map, float d_th, float a_th, int mode, int32_t* assign) {
  int nmatches = 0;
  float dTh = d_th;  // Map.cc:251 overwrites the function-level threshold (mode 1)
  for (int i = 0; i < n_ljl; ++i) {
    assign[i] = -1;
    float pM[4];  // Frame::ComputeWorldPlane: Tcw^T * plane, cv::Mat CV_32F product (double accumulation, one rounding)
    for (int k = 0; k < 4; ++k) {
      double acc = 0;
      for (int r = 0; r < 4; ++r) acc += (double)Tcw[4 * r + k] * (double)planes_cam[4 * i + r];
      pM[k] = (float)acc;
    }
    const double* P = pts + 15 * (size_t)i;
    float ldTh = d_th;
    bool found = false;
    for (int m = 0; m < n_map; ++m) {
      if (mode == 0 && map_bad && map_bad[m]) continue;
      float pW[4] = {map_planes[4 * m], map_planes[4 * m + 1], map_planes[4 * m + 2], map_planes[4 * m + 3]};
      if (mode == 1 && pW[3] < 0)
        for (float& v : pW) v = -v;
      const float angle = pM[0] * pW[0] + pM[1] * pW[1] + pM[2] * pW[2];
      if (angle > a_th || angle < -a_th) {
        float d[5];
        for (int k = 0; k < 5; ++k) d[k] = (float)(pW[0] * P[3 * k] + pW[1] * P[3 * k + 1] + pW[2] * P[3 * k + 2] + pW[3]);
        const float dis = (d[0] + d[1] + d[2] + d[3] + d[4]) / 5;
        float& thr = mode == 0 ? ldTh : dTh;
        if (std::abs(dis) < thr) {
          thr = dis;  // signed: the reference's quirk (InsectlineMatch.cpp:46-47, Map.cc:250-251)
          assign[i] = m;
          found = true;
          if (mode == 1) ++nmatches;
        }
      }
    }
    if (mode == 0 && found) ++nmatches;
  }
  return nmatches;
}

}  // extern "C"
