// Host build of the per-frame line post-processing (line_core.cuh) for the CPU test-suite.
#define PSL_HOST_EMU 1
#include <cstring>
#include <vector>
#include "../../psl_slam_b200/csrc/line_core.cuh"

extern "C" int emu_frame_lines(const float* raw4, int n_raw, int w, int h, int nfeatures, psl_keyline* kl, double* eq,
                               int kl_cap, int* overflow) {
  using namespace psl::line;
  const int cap = n_raw > 16 ? n_raw : 16;
  std::vector<Seg> raw(cap), t1(cap), t2(cap);
  std::memcpy(raw.data(), raw4, (size_t)n_raw * 16);
  std::vector<float> angles(cap), length(cap);
  std::vector<uint16_t> order(cap), tmp16(cap), nb((size_t)cap * kNbCap), nb_cnt(cap), check(cap), loc(cap);
  std::vector<int16_t> code(cap);
  std::vector<uint8_t> flag(cap, 0);
  MergeScratch S{cap, angles.data(), length.data(), order.data(), tmp16.data(), nb.data(), nb_cnt.data(), code.data(),
                 check.data(), loc.data(), flag.data(), 0, nullptr};
  int n = frame_lines(raw.data(), n_raw, t1.data(), t2.data(), w, h, nfeatures, S, kl, eq, kl_cap);
  *overflow = S.overflow;
  return n;
}
