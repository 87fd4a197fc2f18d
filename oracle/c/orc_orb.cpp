// TEST INFRASTRUCTURE — CPU oracle for the ORB extractor (see psl_oracle.h).
// Scalar, sequential restatement of /root/reference/src/ORBextractor.cc with the
// OpenCV primitives it calls re-derived from their published arithmetic
// (each verified against cv2 4.13).  Build with
// -ffp-contract=off (pinned choice H3).  Nothing here is used by the product.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <list>
#include <vector>

#include "psl_oracle.h"

namespace {

const int kEdge = 19;       // EDGE_THRESHOLD, ORBextractor.cc:74
const int kHalfPatch = 15;  // HALF_PATCH_SIZE :73
const int kPatch = 31;      // PATCH_SIZE :72

const int8_t kPattern[1024] = {
#include "../../psl_slam_b200/csrc/orb_pattern.inc"
};

inline int cv_round(float v) { return (int)lrintf(v); }    // cvRound: half-to-even (App. A7)
inline int cv_round_d(double v) { return (int)lrint(v); }

inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

// One axis of cv::resize INTER_LINEAR for 8U (App. A2): source index + Q11 weights.
void resize_axis(int sn, int dn, std::vector<int>& ofs, std::vector<int>& a0, std::vector<int>& a1, bool zero_f) {
  ofs.resize(dn); a0.resize(dn); a1.resize(dn);
  double scale = (double)sn / dn;
  for (int d = 0; d < dn; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= s;
    if (zero_f) {  // horizontal: weights are zeroed when the tap is clamped
      if (s < 0) { s = 0; f = 0.f; }
      if (s >= sn - 1) { s = sn - 1; f = 0.f; }
    }
    ofs[d] = s;
    a0[d] = cv_round((1.f - f) * 2048.f);
    a1[d] = cv_round(f * 2048.f);
  }
}

}  // namespace

extern "C" {

void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride) {
  std::vector<int> xo, xa0, xa1, yo, yb0, yb1;
  resize_axis(sw, dw, xo, xa0, xa1, true);
  resize_axis(sh, dh, yo, yb0, yb1, false);
  std::vector<int> r0(dw), r1(dw);
  for (int y = 0; y < dh; ++y) {
    int sy0 = std::min(std::max(yo[y], 0), sh - 1), sy1 = std::min(std::max(yo[y] + 1, 0), sh - 1);
    const uint8_t* S0 = src + (size_t)sy0 * sstride;
    const uint8_t* S1 = src + (size_t)sy1 * sstride;
    for (int x = 0; x < dw; ++x) {
      int sx = xo[x], sx1 = std::min(sx + 1, sw - 1);
      r0[x] = S0[sx] * xa0[x] + S0[sx1] * xa1[x];
      r1[x] = S1[sx] * xa0[x] + S1[sx1] * xa1[x];
    }
    uint8_t* D = dst + (size_t)y * dstride;
    for (int x = 0; x < dw; ++x)
      D[x] = (uint8_t)((((yb0[y] * (r0[x] >> 4)) >> 16) + ((yb1[y] * (r1[x] >> 4)) >> 16) + 2) >> 2);
  }
}

// cv::GaussianBlur CV_8U bit-exact path (App. A1): Q8 kernel, Q8.8 row pass, (v+2^15)>>16.
void orc_gauss_blur_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride, int ksize) {
  static const int k7[7] = {18, 34, 48, 56, 48, 34, 18};  // sigma 2
  static const int k5[5] = {14, 62, 104, 62, 14};         // sigma 1
  const int* k = ksize == 7 ? k7 : k5;
  const int r = ksize / 2;
  static thread_local std::vector<uint16_t> H;
  H.resize((size_t)w * h);
  for (int y = 0; y < h; ++y) {
    const uint8_t* S = src + (size_t)y * sstride;
    uint16_t* Hr = &H[(size_t)y * w];
    const int xa = std::min(r, w), xb = std::max(w - r, xa);
    for (int x = 0; x < xa; ++x) {
      int acc = 0;
      for (int i = -r; i <= r; ++i) acc += k[i + r] * S[reflect101(x + i, w)];
      Hr[x] = (uint16_t)acc;
    }
    if (ksize == 7)
      for (int x = xa; x < xb; ++x)
        Hr[x] = (uint16_t)(18 * (S[x - 3] + S[x + 3]) + 34 * (S[x - 2] + S[x + 2]) + 48 * (S[x - 1] + S[x + 1]) + 56 * S[x]);
    else
      for (int x = xa; x < xb; ++x)
        Hr[x] = (uint16_t)(14 * (S[x - 2] + S[x + 2]) + 62 * (S[x - 1] + S[x + 1]) + 104 * S[x]);
    for (int x = xb; x < w; ++x) {
      int acc = 0;
      for (int i = -r; i <= r; ++i) acc += k[i + r] * S[reflect101(x + i, w)];
      Hr[x] = (uint16_t)acc;
    }
  }
  for (int y = 0; y < h; ++y) {
    const uint16_t* rows[7];
    for (int j = -r; j <= r; ++j) rows[j + r] = &H[(size_t)reflect101(y + j, h) * w];
    uint8_t* D = dst + (size_t)y * dstride;
    if (ksize == 7) {
      const uint16_t *r0 = rows[0], *r1 = rows[1], *r2 = rows[2], *r3 = rows[3], *r4 = rows[4], *r5 = rows[5], *r6 = rows[6];
      for (int x = 0; x < w; ++x) {
        uint32_t acc = 18u * ((uint32_t)r0[x] + r6[x]) + 34u * ((uint32_t)r1[x] + r5[x]) + 48u * ((uint32_t)r2[x] + r4[x]) + 56u * r3[x];
        D[x] = (uint8_t)((acc + 32768u) >> 16);
      }
    } else {
      for (int x = 0; x < w; ++x) {
        uint32_t acc = 0;
        for (int j = 0; j < ksize; ++j) acc += (uint32_t)k[j] * rows[j][x];
        D[x] = (uint8_t)((acc + 32768u) >> 16);
      }
    }
  }
}

// cv::fastAtan2 scalar (App. A4), degrees.
float orc_fast_atan2(float y, float x) {
  const float sc = (float)(180.0 / 3.141592653589793238462643383279502884);
  const float p1 = 0.9997878412794807f * sc, p3 = -0.3258083974640975f * sc, p5 = 0.1555786518463281f * sc,
              p7 = -0.04432655554792128f * sc;
  float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
  if (ax >= ay) {
    c = ay / (ax + (float)DBL_EPSILON);
    c2 = c * c;
    a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  } else {
    c = ax / (ay + (float)DBL_EPSILON);
    c2 = c * c;
    a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// FAST-9/16 corner score (App. A3): max over the 16 arcs of 9 contiguous circle pixels of
// max(min d, min -d), minus 1; reported only when >= th, else 0.
void orc_fast_score(const uint8_t* img, int w, int h, int stride, int th, uint8_t* score) {
  static const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
  static const int dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
  std::memset(score, 0, (size_t)w * h);
  int off[16];
  for (int k = 0; k < 16; ++k) off[k] = dy[k] * stride + dx[k];
  for (int y = 3; y < h - 3; ++y)
    for (int x = 3; x < w - 3; ++x) {
      const uint8_t* p = img + (size_t)y * stride + x;
      int c = *p;
      // any 9-arc contains one pixel of each antipodal pair: quick rejects
      int d0 = c - p[off[0]], d8 = c - p[off[8]];
      if (std::abs(d0) <= th && std::abs(d8) <= th) continue;
      int d4 = c - p[off[4]], d12 = c - p[off[12]];
      if (std::abs(d4) <= th && std::abs(d12) <= th) continue;
      int d[25];
      unsigned mb = 0, md = 0;
      for (int k = 0; k < 16; ++k) {
        d[k] = c - p[off[k]];
        mb |= (unsigned)(d[k] > th) << k;
        md |= (unsigned)(d[k] < -th) << k;
      }
      auto run9 = [](unsigned m16) {
        unsigned m = m16 | (m16 << 16), r = m & (m >> 1);
        r &= r >> 2;
        r &= r >> 4;
        r &= m >> 8;
        return (r & 0xFFFFu) != 0;
      };
      if (!run9(mb) && !run9(md)) continue;  // not a corner at th: its score would be < th
      for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
      int best = 0;
      for (int s = 0; s < 16; ++s) {
        int mn = d[s], mx = d[s];
        for (int i = 1; i < 9; ++i) { mn = std::min(mn, d[s + i]); mx = std::max(mx, d[s + i]); }
        best = std::max(best, std::max(mn, -mx));
      }
      int sc = best - 1;
      if (sc >= th) score[(size_t)y * w + x] = (uint8_t)sc;
    }
}

void orc_orb_tables(const orc_orb_params* p, float* scale, float* inv_scale, int32_t* quota, int32_t* umax16) {
  int L = p->nlevels;
  double sf = (double)p->scale_factor;  // float ctor arg kept in a double member (ORBextractor.h:99)
  std::vector<float> s(L);
  s[0] = 1.0f;
  for (int i = 1; i < L; ++i) s[i] = (float)(s[i - 1] * sf);  // :421
  for (int i = 0; i < L; ++i) {
    if (scale) scale[i] = s[i];
    if (inv_scale) inv_scale[i] = 1.0f / s[i];
  }
  if (quota) {  // :436-446
    float factor = (float)(1.0f / sf);
    float nd = p->nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)L));
    int sum = 0;
    for (int l = 0; l < L - 1; ++l) {
      quota[l] = cv_round(nd);
      sum += quota[l];
      nd *= factor;
    }
    quota[L - 1] = std::max(p->nfeatures - sum, 0);
  }
  if (umax16) {  // :454-469
    int um[kHalfPatch + 2] = {0};
    int vmax = (int)floor(kHalfPatch * sqrtf(2.f) / 2 + 1);
    int vmin = (int)ceil(kHalfPatch * sqrtf(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (int v = 0; v <= vmax; ++v) um[v] = cv_round_d(sqrt(hp2 - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
      while (um[v0] == um[v0 + 1]) ++v0;
      um[v] = v0;
      ++v0;
    }
    for (int v = 0; v <= kHalfPatch; ++v) umax16[v] = um[v];
  }
}

void orc_orb_level_size(const orc_orb_params* p, int w, int h, int level, int* lw, int* lh) {
  std::vector<float> inv(p->nlevels);
  orc_orb_tables(p, nullptr, inv.data(), nullptr, nullptr);
  *lw = cv_round((float)w * inv[level]);  // :1111-1112
  *lh = cv_round((float)h * inv[level]);
}

// Cell loop of ComputeKeyPointsOctTree, ORBextractor.cc:765-829.  One FAST score map per
// threshold replaces the 815 cv::FAST calls: the score does not depend on the cell, only
// the NMS neighbourhood (clipped to the cell's evaluated interior) and the fallback do.
int orc_fast_cells(const uint8_t* img, int w, int h, int stride, int ini_th, int min_th, float* xyr, int cap) {
  const int minBX = kEdge - 3, minBY = minBX, maxBX = w - kEdge + 3, maxBY = h - kEdge + 3;
  const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
  const int nCols = (int)(width / 30.f), nRows = (int)(height / 30.f);
  if (nCols <= 0 || nRows <= 0) return 0;
  const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
  // The score is threshold independent (App. A3), so cv::FAST(th)'s score image is the raw
  // score map with values below th zeroed: one map at the lower threshold serves both passes.
  static thread_local std::vector<uint8_t> raw;
  raw.resize((size_t)w * h);
  orc_fast_score(img, w, h, stride, std::min(ini_th, min_th), raw.data());
  int n = 0;
  for (int i = 0; i < nRows; ++i) {
    int iniY = minBY + i * hCell, maxY = iniY + hCell + 6;
    if (iniY >= maxBY - 3) continue;
    if (maxY > maxBY) maxY = maxBY;
    for (int j = 0; j < nCols; ++j) {
      int iniX = minBX + j * wCell, maxX = iniX + wCell + 6;
      if (iniX >= maxBX - 6) continue;
      if (maxX > maxBX) maxX = maxBX;
      // evaluated interior of this cv::FAST call: [iniX+3,maxX-3) x [iniY+3,maxY-3)
      int x0 = iniX + 3, x1 = maxX - 3, y0 = iniY + 3, y1 = maxY - 3;
      int start = n;
      for (int pass = 0; pass < 2 && n == start; ++pass) {
        const uint8_t* S = raw.data();
        const int th = pass == 0 ? ini_th : min_th;
        for (int y = y0; y < y1; ++y)
          for (int x = x0; x < x1; ++x) {
            int sc = S[(size_t)y * w + x];
            if (sc < th) continue;
            bool ok = true;
            for (int yy = y - 1; yy <= y + 1 && ok; ++yy)
              for (int xx = x - 1; xx <= x + 1; ++xx) {
                if ((xx == x && yy == y) || xx < x0 || xx >= x1 || yy < y0 || yy >= y1) continue;
                const int sn = S[(size_t)yy * w + xx];
                if (sn >= th && sn >= sc) { ok = false; break; }
              }
            if (!ok) continue;
            if (n < cap) {
              // cell-local coordinate + (j*wCell, i*hCell)  (:820-825) == level coordinate - minBorder
              xyr[3 * n] = (float)(x - iniX) + (float)(j * wCell);
              xyr[3 * n + 1] = (float)(y - iniY) + (float)(i * hCell);
              xyr[3 * n + 2] = (float)sc;
            }
            ++n;
          }
      }
    }
  }
  return n;
}

}  // extern "C"

namespace {

struct Node {
  int ULx, ULy, URx, BRy;
  std::vector<int> keys;  // candidate indices, inherited order
  bool no_more = false;
  int seq = -1;  // creation sequence: the declared stand-in for the node's heap address (H1)
  std::list<int>::iterator lit;
};

// ExtractorNode::DivideNode, ORBextractor.cc:481-537
void divide(const Node& n, const float* xyr, Node c[4]) {
  const int halfX = (int)ceilf((float)(n.URx - n.ULx) / 2), halfY = (int)ceilf((float)(n.BRy - n.ULy) / 2);
  const int mx = n.ULx + halfX, my = n.ULy + halfY;
  c[0].ULx = n.ULx; c[0].URx = mx;    c[0].ULy = n.ULy; c[0].BRy = my;
  c[1].ULx = mx;    c[1].URx = n.URx; c[1].ULy = n.ULy; c[1].BRy = my;
  c[2].ULx = n.ULx; c[2].URx = mx;    c[2].ULy = my;    c[2].BRy = n.BRy;
  c[3].ULx = mx;    c[3].URx = n.URx; c[3].ULy = my;    c[3].BRy = n.BRy;
  for (int k : n.keys) {
    float x = xyr[3 * k], y = xyr[3 * k + 1];
    if (x < mx) (y < my ? c[0] : c[2]).keys.push_back(k);
    else if (y < my) c[1].keys.push_back(k);
    else c[3].keys.push_back(k);
  }
  for (int q = 0; q < 4; ++q) c[q].no_more = c[q].keys.size() == 1;
}

}  // namespace

extern "C" {

// ORBextractor::DistributeOctTree, ORBextractor.cc:539-763.
int orc_octree(const float* xyr, int n, int min_x, int max_x, int min_y, int max_y, int N, float* out_xyr, int cap) {
  if (n <= 0) return 0;
  const int nIni = (int)roundf((float)(max_x - min_x) / (max_y - min_y));
  if (nIni <= 0) return -1;
  const float hX = (float)(max_x - min_x) / nIni;
  std::vector<Node> pool;
  pool.reserve((size_t)4 * n + 16);
  std::list<int> L;
  int seq = 0;
  for (int i = 0; i < nIni; ++i) {
    Node ni;
    ni.ULx = (int)(hX * (float)i);
    ni.URx = (int)(hX * (float)(i + 1));
    ni.ULy = 0;
    ni.BRy = max_y - min_y;
    ni.seq = seq++;
    pool.push_back(ni);
    L.push_back((int)pool.size() - 1);
  }
  for (int k = 0; k < n; ++k) {
    int r = (int)(xyr[3 * k] / hX);
    if (r < 0 || r >= nIni) return -1;
    pool[r].keys.push_back(k);
  }
  for (auto it = L.begin(); it != L.end();) {
    Node& nd = pool[*it];
    if (nd.keys.size() == 1) { nd.no_more = true; ++it; }
    else if (nd.keys.empty()) it = L.erase(it);
    else ++it;
  }
  auto push_children = [&](Node c[4], std::vector<int>& expandable, int* nToExpand) {
    for (int q = 0; q < 4; ++q) {
      if (c[q].keys.empty()) continue;
      c[q].seq = seq++;
      pool.push_back(c[q]);
      int id = (int)pool.size() - 1;
      L.push_front(id);
      if (pool[id].keys.size() > 1) {
        if (nToExpand) ++*nToExpand;
        expandable.push_back(id);
        pool[id].lit = L.begin();
      }
    }
  };
  bool finish = false;
  std::vector<int> last;
  while (!finish) {
    int prev = (int)L.size(), nToExpand = 0;
    last.clear();
    for (auto it = L.begin(); it != L.end();) {
      if (pool[*it].no_more) { ++it; continue; }
      Node c[4];
      divide(pool[*it], xyr, c);
      push_children(c, last, &nToExpand);
      it = L.erase(it);
    }
    if ((int)L.size() >= N || (int)L.size() == prev) finish = true;
    else if ((int)L.size() + nToExpand * 3 > N) {
      while (!finish) {
        prev = (int)L.size();
        std::vector<int> cand = last;
        last.clear();
        std::sort(cand.begin(), cand.end(), [&](int a, int b) {
          if (pool[a].keys.size() != pool[b].keys.size()) return pool[a].keys.size() < pool[b].keys.size();
          return pool[a].seq < pool[b].seq;
        });
        for (int j = (int)cand.size() - 1; j >= 0; --j) {
          Node c[4];
          divide(pool[cand[j]], xyr, c);
          push_children(c, last, nullptr);
          L.erase(pool[cand[j]].lit);
          if ((int)L.size() >= N) break;
        }
        if ((int)L.size() >= N || (int)L.size() == prev) finish = true;
      }
    }
  }
  int m = 0;
  for (int id : L) {  // :744-760 best response per node, first wins ties
    const Node& nd = pool[id];
    int best = nd.keys[0];
    for (size_t k = 1; k < nd.keys.size(); ++k)
      if (xyr[3 * nd.keys[k] + 2] > xyr[3 * best + 2]) best = nd.keys[k];
    if (m < cap) { out_xyr[3 * m] = xyr[3 * best]; out_xyr[3 * m + 1] = xyr[3 * best + 1]; out_xyr[3 * m + 2] = xyr[3 * best + 2]; }
    ++m;
  }
  return m;
}

// IC_Angle, ORBextractor.cc:77-104
float orc_ic_angle(const uint8_t* img, int stride, int x, int y) {
  static int umax[16] = {-1};
  if (umax[0] < 0) {
    orc_orb_params p = {1000, 1.2f, 8, 20, 7};
    int32_t um[16];
    orc_orb_tables(&p, nullptr, nullptr, nullptr, um);
    for (int i = 0; i < 16; ++i) umax[i] = um[i];
  }
  const uint8_t* c = img + (size_t)y * stride + x;
  int m01 = 0, m10 = 0;
  for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
  for (int v = 1; v <= kHalfPatch; ++v) {
    int vsum = 0, d = umax[v];
    for (int u = -d; u <= d; ++u) {
      int vp = c[u + v * stride], vm = c[u - v * stride];
      vsum += vp - vm;
      m10 += u * (vp + vm);
    }
    m01 += v * vsum;
  }
  return orc_fast_atan2((float)m01, (float)m10);
}

// computeOrbDescriptor, ORBextractor.cc:108-147.  H2: cos/sin in double, rounded to float.
void orc_brief(const uint8_t* blur, int stride, float x, float y, float angle_deg, uint8_t* desc) {
  const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
  float angle = angle_deg * factorPI;
  float a = (float)cos((double)angle), b = (float)sin((double)angle);
  const uint8_t* c = blur + (size_t)cv_round(y) * stride + cv_round(x);
  const int8_t* pat = kPattern;
  for (int i = 0; i < 32; ++i, pat += 32) {
    int val = 0;
    for (int t = 0; t < 8; ++t) {
      float x0 = pat[4 * t], y0 = pat[4 * t + 1], x1 = pat[4 * t + 2], y1 = pat[4 * t + 3];
      int t0 = c[cv_round(x0 * b + y0 * a) * stride + cv_round(x0 * a - y0 * b)];
      int t1 = c[cv_round(x1 * b + y1 * a) * stride + cv_round(x1 * a - y1 * b)];
      val |= (t0 < t1) << t;
    }
    desc[i] = (uint8_t)val;
  }
}

int orc_orb_extract(const orc_orb_params* p, const uint8_t* gray, int w, int h, int stride, psl_keypoint* kps,
                    uint8_t* desc, int cap, int* n_out) {
  *n_out = 0;
  if (w <= 0 || h <= 0) return 0;  // :1046-1047
  const int L = p->nlevels;
  std::vector<float> scale(L), inv(L);
  std::vector<int32_t> quota(L);
  orc_orb_tables(p, scale.data(), inv.data(), quota.data(), nullptr);
  static thread_local std::vector<std::vector<uint8_t>> pyr;
  pyr.resize(L);
  std::vector<int> lw(L), lh(L);
  for (int l = 0; l < L; ++l) {  // ComputePyramid :1107-1132 (border never read → not built)
    lw[l] = cv_round((float)w * inv[l]);
    lh[l] = cv_round((float)h * inv[l]);
    if (lw[l] < 2 * kEdge + 7 || lh[l] < 2 * kEdge + 7) return -1;
    pyr[l].resize((size_t)lw[l] * lh[l]);
    if (l == 0)
      for (int y = 0; y < h; ++y) std::memcpy(&pyr[0][(size_t)y * w], gray + (size_t)y * stride, w);
    else
      orc_resize_linear_u8(pyr[l - 1].data(), lw[l - 1], lh[l - 1], lw[l - 1], pyr[l].data(), lw[l], lh[l], lw[l]);
  }
  int n = 0;
  static thread_local std::vector<float> cand, sel;
  static thread_local std::vector<uint8_t> blur;
  for (int l = 0; l < L; ++l) {
    int ccap = (lw[l] * lh[l]) / 4 + 16;
    cand.resize((size_t)3 * ccap);
    int nc = orc_fast_cells(pyr[l].data(), lw[l], lh[l], lw[l], p->ini_th, p->min_th, cand.data(), ccap);
    if (nc <= 0) continue;
    sel.resize((size_t)3 * (nc + 4));
    int ns = orc_octree(cand.data(), nc, kEdge - 3, lw[l] - kEdge + 3, kEdge - 3, lh[l] - kEdge + 3, quota[l],
                        sel.data(), nc + 4);
    if (ns < 0) return -1;
    if (ns == 0) continue;
    blur.resize(pyr[l].size());
    orc_gauss_blur_u8(pyr[l].data(), lw[l], lh[l], lw[l], blur.data(), lw[l], 7);  // :1085-1086
    const float size = (float)(int)(kPatch * scale[l]);                              // :837
    for (int i = 0; i < ns; ++i) {
      if (n >= cap) return PSL_E_CAPACITY;
      float x = sel[3 * i] + (float)(kEdge - 3), y = sel[3 * i + 1] + (float)(kEdge - 3);  // :843-844
      psl_keypoint& k = kps[n];
      k.angle = orc_ic_angle(pyr[l].data(), lw[l], cv_round(x), cv_round(y));
      orc_brief(blur.data(), lw[l], x, y, k.angle, desc + (size_t)32 * n);
      k.x = l ? x * scale[l] : x;  // :1095-1101
      k.y = l ? y * scale[l] : y;
      k.size = size;
      k.response = sel[3 * i + 2];
      k.octave = l;
      k.class_id = -1;
      ++n;
    }
  }
  *n_out = n;
  return 0;
}

}  // extern "C"

// ---- throughput harness for the CPU baseline: one frame per task, `nthreads` workers --------
#include <atomic>
#include <thread>
extern "C" int orc_orb_extract_batch_mt(const orc_orb_params* p, const uint8_t* gray, int B, int w, int h, int stride,
                                        int64_t frame_stride, int nthreads, int32_t* n_out, uint32_t* desc_xor) {
  std::atomic<int> next(0), err(0);
  std::atomic<uint32_t> acc(0);
  auto work = [&]() {
    const int cap = p->nfeatures + 4 * p->nlevels + 16;
    std::vector<psl_keypoint> kps(cap);
    std::vector<uint8_t> desc((size_t)cap * 32);
    for (;;) {
      const int b = next.fetch_add(1);
      if (b >= B) break;
      int n = 0;
      if (orc_orb_extract(p, gray + (size_t)b * frame_stride, w, h, stride, kps.data(), desc.data(), cap, &n)) err = 1;
      n_out[b] = n;
      uint32_t x = 0;
      for (int i = 0; i < n * 8; ++i) x ^= reinterpret_cast<const uint32_t*>(desc.data())[i];
      acc.fetch_xor(x);
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nthreads; ++t) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  if (desc_xor) *desc_xor = acc.load();
  return err.load() ? -1 : 0;
}
