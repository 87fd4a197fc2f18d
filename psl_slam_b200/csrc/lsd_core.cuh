// Sequential core of the LSD line segment detector (cv::LineSegmentDetector with LSD_REFINE_STD, called
// through line_descriptor::LSDDetector by LINEextractor::operator(), add_src/LineExtractor.cpp:336-337;
// the algorithm itself is OpenCV imgproc's, after Grompone von Gioi et al., IPOL 2012).
//
// Region growing is order dependent: every accepted pixel updates the running region angle that the next
// neighbour is tested against, and regions claim pixels from later seeds.  There is no exact intra-frame
// parallel form, so one thread walks one frame's seed list; parallelism comes from the batch (thousands of
// frames in flight) while the per-pixel prologue (blur, 0.8x resize, gradient, bin ordering) is ordinary
// data-parallel CUDA (lsd_kernels.cu).  The same source compiles for the host (PSL_HOST_EMU) so the CPU
// test-suite can check it against the oracle without a GPU.
//
// Exactness: fp64 arithmetic in the reference's operation order (no contraction: -fmad=false /
// -ffp-contract=off), cv::fastAtan2 re-stated with explicit fp32 rounding, cosf/sinf of the running angle
// pinned to (float)cos((double)a) (identical to cv2 4.13 on every golden).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef PSL_HOST_EMU
#define PSL_LSD_HD inline
#else
#define PSL_LSD_HD __device__ __forceinline__
#endif

namespace psl {
namespace lsd {

constexpr double kPi = 3.14159265358979323846;
constexpr double kM32Pi = (3 * kPi) / 2, kM2Pi = 2 * kPi, kDegToRad = kPi / 180;
constexpr float kNotDefDeg = -1024.f;  // angle sentinel (NOTDEF)
constexpr double kAngTh = 22.5, kDensityTh = 0.7, kScale = 0.8;

// Per-frame view of the working set (all pointers to this frame's slices).
struct Frame {
  int W, H;               // scaled image size (0.8x)
  const float* deg;       // gradient angle in degrees (fastAtan2), kNotDefDeg where |grad| <= rho or on the border
  const int32_t* n2;      // gx^2 + gy^2 (modgrad = sqrt(n2 / 4.0))
  uint8_t* used;          // region membership map
  uint32_t* reg;          // region points, packed y << 16 | x
  const uint32_t* seeds;  // pixel indices (y*W + x), bins descending, raster order inside a bin
  int n_seeds;
  int min_reg_size;
  float* out;             // segments x1,y1,x2,y2
  int cap;
};

struct Rect { double x1, y1, x2, y2, width, x, y, theta, dx, dy; };

// cv::fastAtan2, every operation rounded to fp32
PSL_LSD_HD float fast_atan2(float y, float x) {
#ifdef PSL_HOST_EMU
#define PSL_FMUL(a, b) ((a) * (b))
#define PSL_FADD(a, b) ((a) + (b))
#define PSL_FSUB(a, b) ((a) - (b))
#define PSL_FDIV(a, b) ((a) / (b))
#else
#define PSL_FMUL(a, b) __fmul_rn((a), (b))
#define PSL_FADD(a, b) __fadd_rn((a), (b))
#define PSL_FSUB(a, b) __fsub_rn((a), (b))
#define PSL_FDIV(a, b) __fdiv_rn((a), (b))
#endif
  const float sc = (float)(180.0 / kPi);
  const float p1 = PSL_FMUL(0.9997878412794807f, sc), p3 = PSL_FMUL(-0.3258083974640975f, sc),
              p5 = PSL_FMUL(0.1555786518463281f, sc), p7 = PSL_FMUL(-0.04432655554792128f, sc);
  const float eps = (float)2.2204460492503131e-16;
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = PSL_FDIV(ay, PSL_FADD(ax, eps));
    c2 = PSL_FMUL(c, c);
    a = PSL_FMUL(PSL_FADD(PSL_FMUL(PSL_FADD(PSL_FMUL(PSL_FADD(PSL_FMUL(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = PSL_FDIV(ax, PSL_FADD(ay, eps));
    c2 = PSL_FMUL(c, c);
    a = PSL_FSUB(90.f, PSL_FMUL(PSL_FADD(PSL_FMUL(PSL_FADD(PSL_FMUL(PSL_FADD(PSL_FMUL(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0.f) a = PSL_FSUB(180.f, a);
  if (y < 0.f) a = PSL_FSUB(360.f, a);
  return a;
}

PSL_LSD_HD double angle_of(const Frame& f, int idx) { return (double)f.deg[idx] * kDegToRad; }
PSL_LSD_HD double modgrad_of(const Frame& f, int idx) { return sqrt((double)f.n2[idx] / 4.0); }
PSL_LSD_HD double dist_sq(double x1, double y1, double x2, double y2) {
  return (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1);
}
PSL_LSD_HD double angle_diff_signed(double a, double b) {
  double diff = a - b;
  while (diff <= -kPi) diff += kM2Pi;
  while (diff > kPi) diff -= kM2Pi;
  return diff;
}

PSL_LSD_HD bool is_aligned(const Frame& f, int idx, double theta, double prec) {
  const float d = f.deg[idx];
  if (d == kNotDefDeg) return false;
  double n_theta = theta - (double)d * kDegToRad;
  if (n_theta < 0) n_theta = -n_theta;
  if (n_theta > kM32Pi) {
    n_theta -= kM2Pi;
    if (n_theta < 0) n_theta = -n_theta;
  }
  return n_theta <= prec;
}

// region_grow: returns the region size; reg_angle is in/out
PSL_LSD_HD int region_grow(const Frame& f, int sx, int sy, double& reg_angle, double prec) {
  const int W = f.W, H = f.H;
  int n = 0;
  reg_angle = angle_of(f, sy * W + sx);
  f.reg[n++] = ((uint32_t)sy << 16) | (uint32_t)sx;
  float sumdx = (float)cos(reg_angle), sumdy = (float)sin(reg_angle);
  f.used[sy * W + sx] = 1;
  for (int i = 0; i < n; ++i) {
    const int px = (int)(f.reg[i] & 0xFFFFu), py = (int)(f.reg[i] >> 16);
    const int xx_min = px - 1 < 0 ? 0 : px - 1, xx_max = px + 1 > W - 1 ? W - 1 : px + 1;
    const int yy_min = py - 1 < 0 ? 0 : py - 1, yy_max = py + 1 > H - 1 ? H - 1 : py + 1;
    for (int yy = yy_min; yy <= yy_max; ++yy)
      for (int xx = xx_min; xx <= xx_max; ++xx) {
        const int idx = yy * W + xx;
        if (f.used[idx] != 1 && is_aligned(f, idx, reg_angle, prec)) {
          const double angle = angle_of(f, idx);
          f.used[idx] = 1;
          f.reg[n++] = ((uint32_t)yy << 16) | (uint32_t)xx;
          sumdx += (float)cos((double)(float)angle);
          sumdy += (float)sin((double)(float)angle);
          reg_angle = (double)fast_atan2(sumdy, sumdx) * kDegToRad;
        }
      }
  }
  return n;
}

PSL_LSD_HD double get_theta(const Frame& f, int n, double x, double y, double reg_angle, double prec) {
  double Ixx = 0, Iyy = 0, Ixy = 0;
  for (int i = 0; i < n; ++i) {
    const int px = (int)(f.reg[i] & 0xFFFFu), py = (int)(f.reg[i] >> 16);
    const double dx = (double)px - x, dy = (double)py - y, w = modgrad_of(f, py * f.W + px);
    Ixx += dy * dy * w;
    Iyy += dx * dx * w;
    Ixy -= dx * dy * w;
  }
  const double lambda = 0.5 * (Ixx + Iyy - sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
  double theta = (fabs(Ixx) > fabs(Iyy)) ? (double)fast_atan2((float)(lambda - Ixx), (float)Ixy)
                                         : (double)fast_atan2((float)Ixy, (float)(lambda - Iyy));
  theta *= kDegToRad;
  double d = angle_diff_signed(theta, reg_angle);
  if (d < 0) d = -d;
  if (d > prec) theta += kPi;
  return theta;
}

PSL_LSD_HD void region2rect(const Frame& f, int n, double reg_angle, double prec, Rect& rec) {
  double x = 0, y = 0, sum = 0;
  for (int i = 0; i < n; ++i) {
    const int px = (int)(f.reg[i] & 0xFFFFu), py = (int)(f.reg[i] >> 16);
    const double w = modgrad_of(f, py * f.W + px);
    x += (double)px * w;
    y += (double)py * w;
    sum += w;
  }
  x /= sum;
  y /= sum;
  const double theta = get_theta(f, n, x, y, reg_angle, prec);
  const double dx = cos(theta), dy = sin(theta);
  double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
  for (int i = 0; i < n; ++i) {
    const double rdx = (double)(int)(f.reg[i] & 0xFFFFu) - x, rdy = (double)(int)(f.reg[i] >> 16) - y;
    const double l = rdx * dx + rdy * dy, w = -rdx * dy + rdy * dx;
    if (l > l_max) l_max = l; else if (l < l_min) l_min = l;
    if (w > w_max) w_max = w; else if (w < w_min) w_min = w;
  }
  rec.x1 = x + l_min * dx; rec.y1 = y + l_min * dy;
  rec.x2 = x + l_max * dx; rec.y2 = y + l_max * dy;
  rec.width = w_max - w_min;
  rec.x = x; rec.y = y; rec.theta = theta; rec.dx = dx; rec.dy = dy;
  if (rec.width < 1.0) rec.width = 1.0;
}

PSL_LSD_HD double density_of(int n, const Rect& rec) {
  return (double)n / (sqrt(dist_sq(rec.x1, rec.y1, rec.x2, rec.y2)) * rec.width);
}

PSL_LSD_HD bool reduce_region_radius(const Frame& f, int& n, double reg_angle, double prec, Rect& rec, double density) {
  const double xc = (double)(int)(f.reg[0] & 0xFFFFu), yc = (double)(int)(f.reg[0] >> 16);
  const double r1 = dist_sq(xc, yc, rec.x1, rec.y1), r2 = dist_sq(xc, yc, rec.x2, rec.y2);
  double radSq = r1 > r2 ? r1 : r2;
  while (density < kDensityTh) {
    radSq *= 0.75 * 0.75;
    for (int i = 0; i < n; ++i) {
      const int px = (int)(f.reg[i] & 0xFFFFu), py = (int)(f.reg[i] >> 16);
      if (dist_sq(xc, yc, (double)px, (double)py) > radSq) {
        f.used[py * f.W + px] = 0;
        const uint32_t t = f.reg[i];
        f.reg[i] = f.reg[n - 1];
        f.reg[n - 1] = t;
        --n;
        --i;
      }
    }
    if (n < 2) return false;
    region2rect(f, n, reg_angle, prec, rec);
    density = density_of(n, rec);
  }
  return true;
}

PSL_LSD_HD bool refine(const Frame& f, int& n, double reg_angle, double prec, Rect& rec) {
  double density = density_of(n, rec);
  if (density >= kDensityTh) return true;
  const int sx = (int)(f.reg[0] & 0xFFFFu), sy = (int)(f.reg[0] >> 16);
  const double xc = (double)sx, yc = (double)sy, ang_c = angle_of(f, sy * f.W + sx);
  double sum = 0, s_sum = 0;
  int cnt = 0;
  for (int i = 0; i < n; ++i) {
    const int px = (int)(f.reg[i] & 0xFFFFu), py = (int)(f.reg[i] >> 16);
    f.used[py * f.W + px] = 0;
    if (sqrt(dist_sq(xc, yc, (double)px, (double)py)) < rec.width) {
      const double ang_d = angle_diff_signed(angle_of(f, py * f.W + px), ang_c);
      sum += ang_d;
      s_sum += ang_d * ang_d;
      ++cnt;
    }
  }
  const double mean_angle = sum / (double)cnt;
  const double tau = 2.0 * sqrt((s_sum - 2.0 * mean_angle * sum) / (double)cnt + mean_angle * mean_angle);
  n = region_grow(f, sx, sy, reg_angle, tau);
  if (n < 2) return false;
  region2rect(f, n, reg_angle, prec, rec);
  density = density_of(n, rec);
  if (density < kDensityTh) return reduce_region_radius(f, n, reg_angle, prec, rec, density);
  return true;
}

// The seed loop of LineSegmentDetectorImpl::flsd.  Returns the number of segments found (may exceed cap).
PSL_LSD_HD int detect(const Frame& f) {
  const double prec = kPi * kAngTh / 180;
  int nseg = 0;
  for (int s = 0; s < f.n_seeds; ++s) {
    const int idx = (int)f.seeds[s];
    if (f.used[idx] != 0 || f.deg[idx] == kNotDefDeg) continue;
    const int sy = idx / f.W, sx = idx - sy * f.W;
    double reg_angle;
    int n = region_grow(f, sx, sy, reg_angle, prec);
    if (n < f.min_reg_size) continue;
    Rect rec;
    region2rect(f, n, reg_angle, prec, rec);
    if (!refine(f, n, reg_angle, prec, rec)) continue;
    rec.x1 += 0.5; rec.y1 += 0.5; rec.x2 += 0.5; rec.y2 += 0.5;
    rec.x1 /= kScale; rec.y1 /= kScale; rec.x2 /= kScale; rec.y2 /= kScale;
    if (nseg < f.cap) {
      f.out[4 * nseg] = (float)rec.x1; f.out[4 * nseg + 1] = (float)rec.y1;
      f.out[4 * nseg + 2] = (float)rec.x2; f.out[4 * nseg + 3] = (float)rec.y2;
    }
    ++nseg;
  }
  return nseg;
}

}  // namespace lsd
}  // namespace psl
