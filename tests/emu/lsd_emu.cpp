// Host build of the device LSD core (lsd_core.cuh) plus a plain restatement of what the prologue kernels
// produce (gradient angle in degrees, squared norm, stable bin ordering), so that the CPU test-suite can
// check the device algorithm against the oracle without a GPU.
#define PSL_HOST_EMU 1
#include <algorithm>
#include <cmath>
#include <vector>
#include "../../psl_slam_b200/csrc/lsd_core.cuh"

extern "C" int emu_lsd(const uint8_t* scaled, int W, int H, float* out, int cap) {
  using namespace psl::lsd;
  std::vector<float> deg((size_t)W * H, kNotDefDeg);
  std::vector<int32_t> n2((size_t)W * H, 0);
  const double rho = 2.0 / sin(kPi * kAngTh / 180);
  int max_n2 = -1;
  for (int y = 0; y < H - 1; ++y)
    for (int x = 0; x < W - 1; ++x) {
      const int DA = scaled[(y + 1) * W + x + 1] - scaled[y * W + x], BC = scaled[y * W + x + 1] - scaled[(y + 1) * W + x];
      const int gx = DA + BC, gy = DA - BC, q = gx * gx + gy * gy;
      n2[y * W + x] = q;
      if (!(sqrt(q / 4.0) <= rho)) {
        deg[y * W + x] = fast_atan2((float)gx, (float)-gy);
        max_n2 = std::max(max_n2, q);
      }
    }
  const double max_grad = max_n2 >= 0 ? sqrt(max_n2 / 4.0) : -1;
  const double bin_coef = max_grad > 0 ? 1023.0 / max_grad : 0;
  // only pixels with a defined angle can seed a region; they keep their relative order
  std::vector<std::pair<int, uint32_t>> keyed;
  for (int y = 0; y < H - 1; ++y)
    for (int x = 0; x < W - 1; ++x)
      if (deg[y * W + x] != kNotDefDeg) keyed.push_back({(int)(sqrt(n2[y * W + x] / 4.0) * bin_coef), (uint32_t)(y * W + x)});
  std::stable_sort(keyed.begin(), keyed.end(), [](auto& a, auto& b) { return a.first > b.first; });
  std::vector<uint32_t> seeds(keyed.size());
  for (size_t i = 0; i < keyed.size(); ++i) seeds[i] = keyed[i].second;
  std::vector<uint8_t> used((size_t)W * H, 0);
  std::vector<uint32_t> reg((size_t)W * H);
  const double logNT = 5 * (log10((double)W) + log10((double)H)) / 2 + log10(11.0);
  Frame f{W, H, deg.data(), n2.data(), used.data(), reg.data(), seeds.data(), (int)seeds.size(),
          (int)(size_t)(-logNT / log10(kAngTh / 180)), out, cap};
  return detect(f);
}
