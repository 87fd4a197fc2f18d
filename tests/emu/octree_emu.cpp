// Host emulation of the CTA-cooperative octree selection (orb_octree_core.cuh) so the CPU
// test-suite can check the device algorithm against the oracle without a GPU.
#define PSL_HOST_EMU 1
#include <vector>
#include "../../psl_slam_b200/csrc/orb_octree_core.cuh"

template <int NT, int CAP>
static int run(const uint32_t* cand, int n, int n_ini, float hx, int width, int height, int N, uint32_t* out, int cap) {
  auto* sh = new psl::octree::Shared<NT, CAP>();
  memset(sh, 0, sizeof(*sh));
  std::vector<uint32_t> k0(cand, cand + n), k1(n + 1);
  std::vector<uint16_t> kn0(n + 1), kn1(n + 1);
  int r = psl::octree::select<NT, CAP>(*sh, n, k0.data(), k1.data(), kn0.data(), kn1.data(), n_ini, hx, width, height, N,
                                       out, cap);
  if (sh->error == 2) r = -2;
  delete sh;
  return r;
}

extern "C" int emu_octree(const uint32_t* cand, int n, int n_ini, float hx, int width, int height, int N,
                          uint32_t* out, int cap, int nt) {
  if (n == 0) return 0;
  switch (nt) {
    case 32: return run<32, 1024>(cand, n, n_ini, hx, width, height, N, out, cap);
    case 128: return run<128, 256>(cand, n, n_ini, hx, width, height, N, out, cap);
    default: return run<256, 1024>(cand, n, n_ini, hx, width, height, N, out, cap);
  }
}
