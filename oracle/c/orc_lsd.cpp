// TEST INFRASTRUCTURE — CPU oracle for the LSD line segment detector behind
// LINEextractor::operator() (add_src/LineExtractor.cpp:336-337 -> line_descriptor::LSDDetector ->
// cv::createLineSegmentDetector(LSD_REFINE_STD)->detect).  The algorithm lives in OpenCV imgproc
// (lsd.cpp), which is NOT in the reference tree; this is a restatement of the published algorithm
// (Grompone von Gioi et al., "LSD: a Line Segment Detector", IPOL 2012) with OpenCV 4.x's
// data path (8-bit image, 7x7 sigma-0.75 Gaussian in Q8.8, 0.8x INTER_LINEAR_EXACT resize, integer
// 2x2 gradient, 1024-bin ordering), validated segment-for-segment against cv2 4.13 (tests/golden).
//
// order_mode 0: seeds ordered like cv2 4.13 (std::sort of the raster-ordered points by bin, descending;
//               the within-bin order is whatever libstdc++'s introsort produces) — pins the port.
// order_mode 1: bins descending, raster order inside a bin (OpenCV 3.x's bucket lists, i.e. the
//               authors' build; deterministic and what the CUDA path implements).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "psl_oracle.h"

namespace {

const double kPi = 3.14159265358979323846;
const double M_3_2_PI = (3 * kPi) / 2, M_2__PI = 2 * kPi, NOTDEF = -1024.0, DEG_TO_RADS = kPi / 180;
const double ANG_TH = 22.5, QUANT = 2.0, SCALE = 0.8, SIGMA_SCALE = 0.6, DENSITY_TH = 0.7;
const int N_BINS = 1024;

struct RegionPoint { int x, y; double angle, modgrad; };
struct Rect { double x1, y1, x2, y2, width, x, y, theta, dx, dy, prec, p; };
struct NormPoint { int x, y, norm; };

inline int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}
inline double dist_sq(double x1, double y1, double x2, double y2) { return (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1); }
inline double dist(double x1, double y1, double x2, double y2) { return sqrt(dist_sq(x1, y1, x2, y2)); }
inline double angle_diff_signed(double a, double b) {
  double diff = a - b;
  while (diff <= -kPi) diff += M_2__PI;
  while (diff > kPi) diff -= M_2__PI;
  return diff;
}
inline double angle_diff(double a, double b) {
  double d = angle_diff_signed(a, b);
  return d < 0 ? -d : d;
}

struct Lsd {
  int W = 0, H = 0;  // scaled image size
  std::vector<uint8_t> img;
  std::vector<double> angles, modgrad;
  std::vector<uint8_t> used;
  std::vector<NormPoint> ordered;

  // GaussianBlur(7x7, sigma 0.75) on CV_8U: Q8 taps [0,4,56,136,56,4,0] (bit-exact path, derived against cv2)
  // then resize(fx=fy=0.8, INTER_LINEAR_EXACT): Q8 weights, Q8.8 rows, (v + 2^15) >> 16.
  void prologue(const uint8_t* src, int w, int h, int stride) {
    static const int k[7] = {0, 4, 56, 136, 56, 4, 0};
    std::vector<uint16_t> hp((size_t)w * h);
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        int acc = 0;
        for (int i = -3; i <= 3; ++i) acc += k[i + 3] * src[(size_t)y * stride + reflect101(x + i, w)];
        hp[(size_t)y * w + x] = (uint16_t)acc;
      }
    std::vector<uint8_t> g((size_t)w * h);
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        uint32_t acc = 0;
        for (int j = -3; j <= 3; ++j) acc += (uint32_t)k[j + 3] * hp[(size_t)reflect101(y + j, h) * w + x];
        g[(size_t)y * w + x] = (uint8_t)((acc + 32768u) >> 16);
      }
    W = (int)lrint(w * SCALE);
    H = (int)lrint(h * SCALE);
    const double scale = 1.0 / SCALE;
    auto coeffs = [&](int sn, int dn, std::vector<int>& ofs, std::vector<int>& c1) {
      ofs.assign(dn, 0);
      c1.assign(dn, -1);  // -1: single tap
      for (int d = 0; d < dn; ++d) {
        const double f = scale * (d + 0.5) - 0.5;
        const int i = (int)floor(f);
        if (i >= 0 && sn > 1) {
          if (i < sn - 1) { ofs[d] = i; c1[d] = (int)lrint((f - i) * 256.0); }
          else ofs[d] = sn - 1;
        }
      }
    };
    std::vector<int> xo, xc, yo, yc;
    coeffs(w, W, xo, xc);
    coeffs(h, H, yo, yc);
    std::vector<uint32_t> hr((size_t)h * W);
    for (int y = 0; y < h; ++y)
      for (int d = 0; d < W; ++d) {
        const uint8_t* r = &g[(size_t)y * w];
        hr[(size_t)y * W + d] = xc[d] >= 0 ? r[xo[d]] * (256 - xc[d]) + r[xo[d] + 1] * xc[d] : r[xo[d]] * 256;
      }
    img.resize((size_t)W * H);
    for (int d = 0; d < H; ++d)
      for (int x = 0; x < W; ++x) {
        const uint32_t a = hr[(size_t)yo[d] * W + x];
        const uint32_t v = yc[d] >= 0 ? a * (256 - yc[d]) + hr[(size_t)(yo[d] + 1) * W + x] * yc[d] : a * 256;
        img[(size_t)d * W + x] = (uint8_t)((v + 32768u) >> 16);
      }
  }

  void ll_angle(double threshold, int order_mode) {
    angles.assign((size_t)W * H, NOTDEF);
    modgrad.assign((size_t)W * H, 0.0);
    double max_grad = -1;
    for (int y = 0; y < H - 1; ++y)
      for (int x = 0; x < W - 1; ++x) {
        const int DA = img[(size_t)(y + 1) * W + x + 1] - img[(size_t)y * W + x];
        const int BC = img[(size_t)y * W + x + 1] - img[(size_t)(y + 1) * W + x];
        const int gx = DA + BC, gy = DA - BC;
        const double norm = std::sqrt((gx * gx + gy * gy) / 4.0);
        modgrad[(size_t)y * W + x] = norm;
        if (norm <= threshold) angles[(size_t)y * W + x] = NOTDEF;
        else {
          angles[(size_t)y * W + x] = orc_fast_atan2((float)gx, (float)-gy) * DEG_TO_RADS;
          if (norm > max_grad) max_grad = norm;
        }
      }
    const double bin_coef = (max_grad > 0) ? double(N_BINS - 1) / max_grad : 0;
    ordered.clear();
    ordered.reserve((size_t)W * H);
    for (int y = 0; y < H - 1; ++y)
      for (int x = 0; x < W - 1; ++x) ordered.push_back({x, y, int(modgrad[(size_t)y * W + x] * bin_coef)});
    if (order_mode == 0)
      std::sort(ordered.begin(), ordered.end(), [](const NormPoint& a, const NormPoint& b) { return a.norm > b.norm; });
    else
      std::stable_sort(ordered.begin(), ordered.end(), [](const NormPoint& a, const NormPoint& b) { return a.norm > b.norm; });
  }

  bool is_aligned(int x, int y, double theta, double prec) const {
    if (x < 0 || y < 0 || x >= W || y >= H) return false;
    const double a = angles[(size_t)y * W + x];
    if (a == NOTDEF) return false;
    double n_theta = theta - a;
    if (n_theta < 0) n_theta = -n_theta;
    if (n_theta > M_3_2_PI) {
      n_theta -= M_2__PI;
      if (n_theta < 0) n_theta = -n_theta;
    }
    return n_theta <= prec;
  }

  void region_grow(int sx, int sy, std::vector<RegionPoint>& reg, double& reg_angle, double prec) {
    reg.clear();
    reg_angle = angles[(size_t)sy * W + sx];
    reg.push_back({sx, sy, reg_angle, modgrad[(size_t)sy * W + sx]});
    float sumdx = float(std::cos(reg_angle)), sumdy = float(std::sin(reg_angle));
    used[(size_t)sy * W + sx] = 1;
    for (size_t i = 0; i < reg.size(); ++i) {
      const int px = reg[i].x, py = reg[i].y;
      const int xx_min = std::max(px - 1, 0), xx_max = std::min(px + 1, W - 1);
      const int yy_min = std::max(py - 1, 0), yy_max = std::min(py + 1, H - 1);
      for (int yy = yy_min; yy <= yy_max; ++yy)
        for (int xx = xx_min; xx <= xx_max; ++xx) {
          uint8_t& u = used[(size_t)yy * W + xx];
          if (u != 1 && is_aligned(xx, yy, reg_angle, prec)) {
            const double angle = angles[(size_t)yy * W + xx];
            u = 1;
            reg.push_back({xx, yy, angle, modgrad[(size_t)yy * W + xx]});
            sumdx += (float)cos((double)float(angle));
            sumdy += (float)sin((double)float(angle));
            reg_angle = orc_fast_atan2(sumdy, sumdx) * DEG_TO_RADS;
          }
        }
    }
  }

  double get_theta(const std::vector<RegionPoint>& reg, double x, double y, double reg_angle, double prec) const {
    double Ixx = 0, Iyy = 0, Ixy = 0;
    for (const RegionPoint& r : reg) {
      const double dx = double(r.x) - x, dy = double(r.y) - y, w = r.modgrad;
      Ixx += dy * dy * w;
      Iyy += dx * dx * w;
      Ixy -= dx * dy * w;
    }
    const double lambda = 0.5 * (Ixx + Iyy - sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
    double theta = (fabs(Ixx) > fabs(Iyy)) ? double(orc_fast_atan2(float(lambda - Ixx), float(Ixy)))
                                           : double(orc_fast_atan2(float(Ixy), float(lambda - Iyy)));
    theta *= DEG_TO_RADS;
    if (angle_diff(theta, reg_angle) > prec) theta += kPi;
    return theta;
  }

  void region2rect(const std::vector<RegionPoint>& reg, double reg_angle, double prec, double p, Rect& rec) const {
    double x = 0, y = 0, sum = 0;
    for (const RegionPoint& r : reg) {
      x += double(r.x) * r.modgrad;
      y += double(r.y) * r.modgrad;
      sum += r.modgrad;
    }
    x /= sum;
    y /= sum;
    const double theta = get_theta(reg, x, y, reg_angle, prec);
    const double dx = cos(theta), dy = sin(theta);
    double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
    for (const RegionPoint& r : reg) {
      const double rdx = double(r.x) - x, rdy = double(r.y) - y;
      const double l = rdx * dx + rdy * dy, w = -rdx * dy + rdy * dx;
      if (l > l_max) l_max = l; else if (l < l_min) l_min = l;
      if (w > w_max) w_max = w; else if (w < w_min) w_min = w;
    }
    rec.x1 = x + l_min * dx; rec.y1 = y + l_min * dy;
    rec.x2 = x + l_max * dx; rec.y2 = y + l_max * dy;
    rec.width = w_max - w_min;
    rec.x = x; rec.y = y; rec.theta = theta; rec.dx = dx; rec.dy = dy; rec.prec = prec; rec.p = p;
    if (rec.width < 1.0) rec.width = 1.0;
  }

  bool reduce_region_radius(std::vector<RegionPoint>& reg, double reg_angle, double prec, double p, Rect& rec,
                            double density, double density_th) {
    const double xc = double(reg[0].x), yc = double(reg[0].y);
    const double r1 = dist_sq(xc, yc, rec.x1, rec.y1), r2 = dist_sq(xc, yc, rec.x2, rec.y2);
    double radSq = r1 > r2 ? r1 : r2;
    while (density < density_th) {
      radSq *= 0.75 * 0.75;
      for (size_t i = 0; i < reg.size(); ++i) {
        if (dist_sq(xc, yc, double(reg[i].x), double(reg[i].y)) > radSq) {
          used[(size_t)reg[i].y * W + reg[i].x] = 0;
          std::swap(reg[i], reg[reg.size() - 1]);
          reg.pop_back();
          --i;
        }
      }
      if (reg.size() < 2) return false;
      region2rect(reg, reg_angle, prec, p, rec);
      density = double(reg.size()) / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
    }
    return true;
  }

  bool refine(std::vector<RegionPoint>& reg, double reg_angle, double prec, double p, Rect& rec, double density_th) {
    double density = double(reg.size()) / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
    if (density >= density_th) return true;
    const double xc = double(reg[0].x), yc = double(reg[0].y), ang_c = reg[0].angle;
    double sum = 0, s_sum = 0;
    int n = 0;
    for (const RegionPoint& r : reg) {
      used[(size_t)r.y * W + r.x] = 0;
      if (dist(xc, yc, r.x, r.y) < rec.width) {
        const double ang_d = angle_diff_signed(r.angle, ang_c);
        sum += ang_d;
        s_sum += ang_d * ang_d;
        ++n;
      }
    }
    const double mean_angle = sum / double(n);
    const double tau = 2.0 * sqrt((s_sum - 2.0 * mean_angle * sum) / double(n) + mean_angle * mean_angle);
    const int sx = reg[0].x, sy = reg[0].y;
    region_grow(sx, sy, reg, reg_angle, tau);
    if (reg.size() < 2) return false;
    region2rect(reg, reg_angle, prec, p, rec);
    density = double(reg.size()) / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
    if (density < density_th) return reduce_region_radius(reg, reg_angle, prec, p, rec, density, density_th);
    return true;
  }

  int detect(const uint8_t* src, int w, int h, int stride, int order_mode, float* lines, int cap) {
    const double prec = kPi * ANG_TH / 180, p = ANG_TH / 180, rho = QUANT / sin(prec);
    (void)SIGMA_SCALE;
    prologue(src, w, h, stride);
    ll_angle(rho, order_mode);
    const double LOG_NT = 5 * (log10(double(W)) + log10(double(H))) / 2 + log10(11.0);
    const size_t min_reg_size = size_t(-LOG_NT / log10(p));
    used.assign((size_t)W * H, 0);
    std::vector<RegionPoint> reg;
    int n = 0;
    for (const NormPoint& pt : ordered) {
      if (used[(size_t)pt.y * W + pt.x] == 0 && angles[(size_t)pt.y * W + pt.x] != NOTDEF) {
        double reg_angle;
        region_grow(pt.x, pt.y, reg, reg_angle, prec);
        if (reg.size() < min_reg_size) continue;
        Rect rec;
        region2rect(reg, reg_angle, prec, p, rec);
        if (!refine(reg, reg_angle, prec, p, rec, DENSITY_TH)) continue;
        rec.x1 += 0.5; rec.y1 += 0.5; rec.x2 += 0.5; rec.y2 += 0.5;
        rec.x1 /= SCALE; rec.y1 /= SCALE; rec.x2 /= SCALE; rec.y2 /= SCALE;
        if (n < cap) {
          lines[4 * n] = float(rec.x1); lines[4 * n + 1] = float(rec.y1);
          lines[4 * n + 2] = float(rec.x2); lines[4 * n + 3] = float(rec.y2);
        }
        ++n;
      }
    }
    return n;
  }
};

}  // namespace

// stage accessor: the blurred + 0.8x-resized 8-bit image LSD works on (for the prologue parity tests)
extern "C" int orc_lsd_scaled_image(const uint8_t* img, int w, int h, int stride, uint8_t* out, int* W, int* H) {
  Lsd L;
  L.prologue(img, w, h, stride);
  *W = L.W;
  *H = L.H;
  if (out) std::memcpy(out, L.img.data(), L.img.size());
  return 0;
}

extern "C" int orc_lsd_detect(const uint8_t* img, int w, int h, int stride, int order_mode, float* lines, int cap) {
  if (w <= 0 || h <= 0) return 0;
  static thread_local Lsd L;
  return L.detect(img, w, h, stride, order_mode, lines, cap);
}
