// TEST INFRASTRUCTURE — C entry point over the reference's own ORB_SLAM2::ORBextractor, compiled unchanged from
// /root/reference/src/ORBextractor.cc against oracle/ref_shim (recipe: oracle/Makefile, output oracle/_ref/).
//
// Two of the oracle's declared choices are properties of the build, not of the source, and are selected here:
//   H1  DistributeOctTree sorts pair<int, ExtractorNode*> (ORBextractor.cc:684): node-count ties fall to the heap
//       address of the std::list node.  -DREF_PIN_ALLOC routes operator new to a bump arena that is rewound before
//       every extraction, so "larger address" == "created later" (the oracle's stated tie-break).  Without it the
//       addresses are glibc malloc's, as in the authors' binary.
//   H2  cos(float) / sin(float) (:112-113) resolve to libm cosf / sinf, which are not correctly rounded.
//       -DREF_PIN_TRIG gives the library private cosf / sinf = (float)cos((double)x).
// libpsl_ref_orb.so is built with both (the oracle's pinned reading), libpsl_ref_orb_native.so with neither;
// tests/test_oracle_ref.py requires oracle == pinned bit for bit and reports how far native moves.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/psl_frontend.h"
#include "ORBextractor.h"

#ifdef REF_PIN_TRIG
extern "C" {
__attribute__((visibility("hidden"))) float cosf(float x) noexcept { return (float)cos((double)x); }
__attribute__((visibility("hidden"))) float sinf(float x) noexcept { return (float)sin((double)x); }
// gcc folds the reference's cos(angle), sin(angle) pair into one sincosf call
__attribute__((visibility("hidden"))) void sincosf(float x, float* s, float* c) noexcept {
  *s = (float)sin((double)x);
  *c = (float)cos((double)x);
}
}
#endif

#ifdef REF_PIN_ALLOC
namespace {
constexpr size_t kArena = (size_t)1 << 30;
char* g_arena = nullptr;
size_t g_top = 0;
bool g_in_extract = false;
inline bool in_arena(void* p) { return g_arena && (char*)p >= g_arena && (char*)p < g_arena + kArena; }
}  // namespace
void* operator new(size_t n) {
  if (g_in_extract) {
    const size_t a = (g_top + 15) & ~(size_t)15;
    if (a + n <= kArena) {
      g_top = a + n;
      return g_arena + a;
    }
  }
  void* p = malloc(n ? n : 1);
  if (!p) throw std::bad_alloc();
  return p;
}
void* operator new[](size_t n) { return operator new(n); }
void operator delete(void* p) noexcept { if (!in_arena(p)) free(p); }
void operator delete[](void* p) noexcept { if (!in_arena(p)) free(p); }
void operator delete(void* p, size_t) noexcept { if (!in_arena(p)) free(p); }
void operator delete[](void* p, size_t) noexcept { if (!in_arena(p)) free(p); }
#endif

extern "C" {

// returns the number of keypoints (may exceed cap; only cap rows are written)
int ref_orb_extract(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, const uint8_t* gray, int w,
                    int h, int stride, psl_keypoint* kps, uint8_t* desc, int cap) {
  std::vector<cv::KeyPoint> keys;
  cv::Mat descriptors;
  int n = 0;
#ifdef REF_PIN_ALLOC
  if (!g_arena) g_arena = (char*)malloc(kArena);
  g_top = 0;
  g_in_extract = true;
#endif
  {
    ORB_SLAM2::ORBextractor ex(nfeatures, scale_factor, nlevels, ini_th, min_th);
    cv::Mat image(h, w, CV_8UC1, (void*)gray, (size_t)stride);
    ex(image, cv::Mat(), keys, descriptors);
    n = (int)keys.size();
    for (int i = 0; i < n && i < cap; ++i) {
      const cv::KeyPoint& k = keys[i];
      kps[i] = psl_keypoint{k.pt.x, k.pt.y, k.size, k.angle, k.response, k.octave, k.class_id};
      std::memcpy(desc + (size_t)i * 32, descriptors.ptr(i), 32);
    }
    // everything allocated during the call dies here, before the arena is rewound by the next call
    std::vector<cv::KeyPoint>().swap(keys);
    descriptors.release();
  }
#ifdef REF_PIN_ALLOC
  g_in_extract = false;
#endif
  return n;
}

// getters of the reference object (ORBextractor.h:63-83), for A0
int ref_orb_tables(int nfeatures, float scale_factor, int nlevels, float* scale, float* inv_scale, float* sigma2,
                   float* inv_sigma2) {
  ORB_SLAM2::ORBextractor ex(nfeatures, scale_factor, nlevels, 20, 7);
  const std::vector<float> a = ex.GetScaleFactors(), b = ex.GetInverseScaleFactors(), c = ex.GetScaleSigmaSquares(),
                           d = ex.GetInverseScaleSigmaSquares();
  for (int i = 0; i < nlevels; ++i) {
    scale[i] = a[i]; inv_scale[i] = b[i]; sigma2[i] = c[i]; inv_sigma2[i] = d[i];
  }
  return ex.GetLevels();
}

}  // extern "C"
