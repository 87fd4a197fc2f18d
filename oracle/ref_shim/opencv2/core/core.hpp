// TEST INFRASTRUCTURE — a minimal stand-in for the OpenCV headers that /root/reference/src/ORBextractor.cc
// includes, written from OpenCV's documented API so that the reference source compiles UNCHANGED into
// oracle/_ref (recipe: oracle/Makefile).  Only what that file uses exists: 8-bit single-channel cv::Mat with
// shared storage and ROIs, KeyPoint / Point_ / Size / Rect, Input/OutputArray over Mat, and the primitives
// FAST / resize / copyMakeBorder / GaussianBlur / fastAtan2 whose arithmetic is the cv2-4.13-verified
// restatement of oracle/c/orc_orb.cpp.  Never part of the product.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

typedef unsigned char uchar;
#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0

inline int cvRound(double v) { return (int)lrint(v); }   // round half to even (SSE2 cvtsd2si)
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {

template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <typename U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
};
// cv::Point2i(float, float) goes through saturate_cast<int>(float) = cvRound in OpenCV only for the converting
// constructor from Point_<float>; the (T, T) constructor takes ints, so float arguments convert by truncation
// exactly as in C++ (that is what ORBextractor.cc:555-556 relies on).
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
template <typename T> inline Point_<T>& operator*=(Point_<T>& p, float s) {   // Point_<float> *= float
  p.x = (T)(p.x * s);
  p.y = (T)(p.y * s);
  return p;
}

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
  Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};

struct KeyPoint {
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

class Mat {
 public:
  int rows, cols;
  size_t step;
  uchar* data;
  Mat() : rows(0), cols(0), step(0), data(nullptr) {}
  Mat(int r, int c, int type) : Mat() { create(r, c, type); }
  Mat(Size sz, int type) : Mat() { create(sz.height, sz.width, type); }
  Mat(int r, int c, int, void* ext, size_t step_) : rows(r), cols(c), step(step_), data((uchar*)ext) {}   // user data, not owned
  void create(int r, int c, int) {
    if (data && r == rows && c == cols) return;   // cv::Mat::create keeps a buffer of the right size (ROIs included)
    buf_.reset(new std::vector<uchar>((size_t)r * c));
    rows = r; cols = c; step = (size_t)c; data = buf_->data();
  }
  void create(Size sz, int type) { create(sz.height, sz.width, type); }
  void release() { buf_.reset(); rows = cols = 0; step = 0; data = nullptr; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return CV_8UC1; }
  size_t step1() const { return step; }
  template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  uchar* ptr(int y = 0) { return data + (size_t)y * step; }
  const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
  Mat operator()(const Rect& r) const {
    assert(r.x >= 0 && r.y >= 0 && r.x + r.width <= cols && r.y + r.height <= rows);
    Mat m(*this);
    m.data = data + (size_t)r.y * step + r.x;
    m.rows = r.height; m.cols = r.width;
    return m;
  }
  Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
  Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
  Mat clone() const {
    Mat m(rows, cols, CV_8UC1);
    for (int y = 0; y < rows; ++y) std::memcpy(m.ptr(y), ptr(y), (size_t)cols);
    return m;
  }
  // computeDescriptors assigns `descriptors = Mat::zeros(n, 32, CV_8UC1)` to a ROI header of the output matrix
  // (ORBextractor.cc:1037, 1089).  In OpenCV that is Mat::operator=(const MatExpr&): create() — a no-op for a
  // header that already has this size and type — followed by a fill, i.e. the zeros are written INTO the ROI and the
  // header stays bound to the caller's buffer.  MatExpr models exactly that.
  struct MatExpr {
    int r, c;
    operator Mat() const { Mat m; m = *this; return m; }
  };
  static MatExpr zeros(int r, int c, int) { return MatExpr{r, c}; }
  Mat& operator=(const MatExpr& e) {
    create(e.r, e.c, CV_8UC1);
    for (int y = 0; y < rows; ++y) std::memset(ptr(y), 0, (size_t)cols);
    return *this;
  }

 private:
  std::shared_ptr<std::vector<uchar>> buf_;
};

// _InputArray / _OutputArray over Mat only
class _InputArray {
 public:
  _InputArray() : m_(nullptr) {}
  _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
  Mat getMat() const { return m_ ? *m_ : Mat(); }
  bool empty() const { return !m_ || m_->empty(); }
 protected:
  Mat* m_;
};
class _OutputArray : public _InputArray {
 public:
  _OutputArray(Mat& m) : _InputArray(m) {}
  void create(int r, int c, int type) const { m_->create(r, c, type); }
  void release() const { m_->release(); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
enum { INTER_LINEAR = 1 };

float fastAtan2(float y, float x);
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType);
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0,
                  int borderType = BORDER_REFLECT_101);

struct KeyPointsFilter {   // only named by the dead ComputeKeyPointsOld (ORBextractor.cc:855-1032)
  static void retainBest(std::vector<KeyPoint>& keypoints, int npoints);
};

}  // namespace cv
