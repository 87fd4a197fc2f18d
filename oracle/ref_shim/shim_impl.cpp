// TEST INFRASTRUCTURE — the OpenCV primitives behind oracle/ref_shim/opencv2/*, each a thin adapter over the
// cv2-4.13-verified arithmetic of oracle/c/orc_orb.cpp.  cv::FAST here is a genuine
// per-call detector on the sub-image it is handed (score map of that sub-image + 3x3 non-maximum suppression,
// App. A3), i.e. the reference's 815 calls per frame run as written, not the oracle's one-map shortcut.
#include <algorithm>
#include <cstdlib>
#include <new>

#include "../c/psl_oracle.h"
#include "opencv2/core/core.hpp"

extern "C" {
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride);
void orc_gauss_blur_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride, int ksize);
float orc_fast_atan2(float y, float x);
void orc_fast_score(const uint8_t* img, int w, int h, int stride, int th, uint8_t* score);
}

namespace cv {

float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression) {
  const Mat img = image.getMat();
  keypoints.clear();
  const int w = img.cols, h = img.rows;
  if (w < 7 || h < 7) return;
  std::vector<uint8_t> score((size_t)w * h);
  orc_fast_score(img.data, w, h, (int)img.step, threshold, score.data());   // 0 where not a corner at `threshold`
  for (int y = 3; y < h - 3; ++y)
    for (int x = 3; x < w - 3; ++x) {
      const int s = score[(size_t)y * w + x];
      if (!s) continue;
      bool keep = true;
      if (nonmaxSuppression)
        for (int dy = -1; dy <= 1 && keep; ++dy)
          for (int dx = -1; dx <= 1; ++dx)
            if ((dx || dy) && score[(size_t)(y + dy) * w + (x + dx)] >= s) { keep = false; break; }
      if (keep) keypoints.push_back(KeyPoint((float)x, (float)y, 7.f, -1.f, (float)s));
    }
}

void resize(InputArray src_, OutputArray dst_, Size dsize, double, double, int) {
  const Mat src = src_.getMat();
  dst_.create(dsize.height, dsize.width, CV_8UC1);   // keeps a correctly sized ROI (ORBextractor.cc:1115-1120)
  Mat dst = dst_.getMat();
  orc_resize_linear_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}

static int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
  return i;
}

// BORDER_REFLECT_101 (+ BORDER_ISOLATED): the source may be a ROI of the destination (ComputePyramid's in-place
// use, ORBextractor.cc:1122), so it is copied first.  Without BORDER_ISOLATED OpenCV would read the pixels around a
// ROI; the reference passes a whole image in that case (:1127), where both agree.
void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int) {
  const Mat src = src_.getMat().clone();
  dst_.create(src.rows + top + bottom, src.cols + left + right, CV_8UC1);
  Mat dst = dst_.getMat();
  for (int y = 0; y < dst.rows; ++y) {
    const uchar* s = src.ptr(reflect101(y - top, src.rows));
    uchar* d = dst.ptr(y);
    for (int x = 0; x < dst.cols; ++x) d[x] = s[reflect101(x - left, src.cols)];
  }
}

void GaussianBlur(InputArray src_, OutputArray dst_, Size ksize, double sigmaX, double sigmaY, int) {
  // only the two kernels whose Q8 taps were verified against cv2 exist (7x7 sigma 2, 5x5 sigma 1)
  if (!((ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2) ||
        (ksize.width == 5 && ksize.height == 5 && sigmaX == 1 && sigmaY == 1)))
    abort();
  const Mat src = src_.getMat().clone();
  dst_.create(src.rows, src.cols, CV_8UC1);
  Mat dst = dst_.getMat();
  orc_gauss_blur_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, (int)dst.step, ksize.width);
}

void KeyPointsFilter::retainBest(std::vector<KeyPoint>&, int) { abort(); }   // dead code path in the reference

}  // namespace cv
