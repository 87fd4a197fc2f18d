// psl_orbslam_shim.hpp — the two extractor classes of PSL-SLAM re-created on top of the C-ABI (psl_frontend.h), with the
// reference's own names and signatures, so that Frame / Tracking compile unchanged:
//
//   ORB_SLAM2::ORBextractor   include/ORBextractor.h:45-112, src/ORBextractor.cc:410-470, 1043-1105
//   ORB_SLAM2::LINEextractor  add_inc/LineExtractor.h:159-253, add_src/LineExtractor.cpp:6-25, 325-366
//
// It replaces src/ORBextractor.cc and the extraction half of add_src/LineExtractor.cpp in PSL-SLAM's source list
// (CMakeLists.txt:55-87); compile it inside PSL-SLAM, where <opencv2/...>, opencv_contrib's line_descriptor and Eigen
// exist.  The only things it needs from those headers are cv::Mat / cv::KeyPoint / cv::InputArray / cv::OutputArray,
// cv::line_descriptor::KeyLine and Eigen::Vector3d.  This repository compiles it against small mock headers of
// exactly those types and runs it on the GPU (tests/shim, tests/test_shim.py).  The matcher classes take Frame /
// KeyFrame / MapPoint objects and therefore live with the reference's own headers; INTEGRATION.md §3-§5 shows them.
//
// Threading: like the reference's extractors (one per thread, Frame.cc:92-93) an object is not re-entrant.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "psl_frontend.h"

#ifndef PSL_SHIM_MAX_WIDTH
#define PSL_SHIM_MAX_WIDTH 1920   // largest frame the context allocates for
#define PSL_SHIM_MAX_HEIGHT 1080
#endif

namespace ORB_SLAM2 {

namespace psl_shim {
inline void check(psl_ctx* ctx, int rc, const char* what) {
  if (rc != PSL_OK) throw std::runtime_error(std::string(what) + ": " + (ctx ? psl_last_error(ctx) : "no context"));
}
}  // namespace psl_shim

class ORBextractor {
 public:
  enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

  ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST) {
    psl_config cfg;
    psl_default_config(&cfg);
    cfg.orb_nfeatures = nfeatures;
    cfg.orb_scale_factor = scaleFactor;
    cfg.orb_nlevels = nlevels;
    cfg.orb_ini_th_fast = iniThFAST;
    cfg.orb_min_th_fast = minThFAST;
    cfg.max_width = PSL_SHIM_MAX_WIDTH;
    cfg.max_height = PSL_SHIM_MAX_HEIGHT;
    cfg.max_batch = 1;   // online tracking: one frame per call
    if (psl_create(&cfg, &ctx_) != PSL_OK) throw std::runtime_error("psl_create failed: no B200 or bad settings");
    mvScaleFactor.resize(nlevels);
    mvInvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels);
    mvInvLevelSigma2.resize(nlevels);
    int32_t L = 0;
    psl_shim::check(ctx_, psl_orb_tables(ctx_, &L, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                                         mvInvLevelSigma2.data(), nullptr), "psl_orb_tables");
    scaleFactor_ = scaleFactor;
    cap_ = nfeatures + 4 * nlevels + 64;   // DistributeOctTree returns at most N + 3 keys per level
    kps_.resize(cap_);
    desc_.resize((size_t)cap_ * 32);
  }
  ~ORBextractor() { psl_destroy(ctx_); }
  ORBextractor(const ORBextractor&) = delete;
  ORBextractor& operator=(const ORBextractor&) = delete;

  // Same contract as ORBextractor.cc:1043-1105: the mask is ignored, an empty image returns silently, the image must be
  // CV_8UC1 (asserted by the reference at :1050), descriptors = N x 32 CV_8U (released when N = 0).
  void operator()(cv::InputArray image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint>& keypoints,
                  cv::OutputArray descriptors) {
    if (image.empty()) return;
    const cv::Mat im = image.getMat();
    int32_t n = 0;
    psl_shim::check(ctx_, psl_orb_extract(ctx_, im.data, im.cols, im.rows, (int32_t)im.step, kps_.data(), desc_.data(), cap_,
                                          &n), "psl_orb_extract");
    static_assert(sizeof(cv::KeyPoint) == sizeof(psl_keypoint), "cv::KeyPoint and psl_keypoint share one 28-byte layout");
    keypoints.resize((size_t)n);
    if (n) std::memcpy(static_cast<void*>(keypoints.data()), kps_.data(), (size_t)n * sizeof(psl_keypoint));
    if (n == 0) {
      descriptors.release();
      return;
    }
    descriptors.create(n, 32, CV_8U);
    cv::Mat out = descriptors.getMat();
    for (int i = 0; i < n; ++i) std::memcpy(out.ptr(i), &desc_[(size_t)i * 32], 32);
  }

  int GetLevels() { return (int)mvScaleFactor.size(); }
  float GetScaleFactor() { return scaleFactor_; }
  std::vector<float> GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> GetInverseScaleFactors() { return mvInvScaleFactor; }
  std::vector<float> GetScaleSigmaSquares() { return mvLevelSigma2; }
  std::vector<float> GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

  // Read by the stereo constructor only (Frame.cc:1172, 1262), which this fork does not build; left empty.
  std::vector<cv::Mat> mvImagePyramid;

 protected:
  psl_ctx* ctx_ = nullptr;
  int cap_ = 0;
  float scaleFactor_ = 1.f;
  std::vector<psl_keypoint> kps_;
  std::vector<uint8_t> desc_;
  std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
};

class LINEextractor {
 public:
  LINEextractor() {}
  LINEextractor(int _numOctaves, float _scale, unsigned int _nLSDFeature, double _min_line_length)
      : numOctaves(_numOctaves), scale(_scale), nLSDFeature(_nLSDFeature), min_line_length(_min_line_length) {
    psl_config cfg;
    psl_default_config(&cfg);
    cfg.line_nlevels = _numOctaves;
    cfg.line_scale_factor = _scale;
    cfg.line_nfeatures = (int32_t)_nLSDFeature;
    cfg.line_min_length = (float)_min_line_length;
    cfg.max_width = PSL_SHIM_MAX_WIDTH;
    cfg.max_height = PSL_SHIM_MAX_HEIGHT;
    cfg.max_batch = 1;
    cfg.line_chunk_frames = 1;
    if (psl_create(&cfg, &ctx_) != PSL_OK) throw std::runtime_error("psl_create failed: no B200 or bad settings");
    // LineExtractor.cpp:9-24: float tables built by repeated multiplication
    mvScaleFactor.assign((size_t)numOctaves, 1.f);
    mvLevelSigma2.assign((size_t)numOctaves, 1.f);
    for (int i = 1; i < numOctaves; ++i) {
      mvScaleFactor[i] = mvScaleFactor[i - 1] * scale;
      mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
    }
    mvInvScaleFactor.resize((size_t)numOctaves);
    mvInvLevelSigma2.resize((size_t)numOctaves);
    for (int i = 0; i < numOctaves; ++i) {
      mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i];
      mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i];
    }
    kl_.resize(nLSDFeature);
    desc_.resize((size_t)nLSDFeature * 32);
    eq_.resize((size_t)nLSDFeature * 3);
  }
  ~LINEextractor() { if (ctx_) psl_destroy(ctx_); }
  LINEextractor(const LINEextractor&) = delete;
  LINEextractor& operator=(const LINEextractor&) = delete;

  // LineExtractor.cpp:325-366: LSD -> long-line merge -> strongest nLSDFeature -> LBD -> 2-D line equations.
  // The reference's only caller passes an empty mask (Frame.cc:493); a mask is not supported.
  void operator()(cv::InputArray image, cv::InputArray /*mask*/, std::vector<cv::line_descriptor::KeyLine>& keylines,
                  cv::OutputArray descriptors, std::vector<Eigen::Vector3d>& lineVec2d) {
    if (image.empty()) return;
    const cv::Mat im = image.getMat();
    int32_t n = 0;
    psl_shim::check(ctx_, psl_line_extract(ctx_, im.data, im.cols, im.rows, (int32_t)im.step, kl_.data(), desc_.data(),
                                           eq_.data(), nullptr, (int32_t)nLSDFeature, &n), "psl_line_extract");
    static_assert(sizeof(cv::line_descriptor::KeyLine) == sizeof(psl_keyline), "KeyLine and psl_keyline share one 68-byte layout");
    keylines.resize((size_t)n);
    if (n) std::memcpy(static_cast<void*>(keylines.data()), kl_.data(), (size_t)n * sizeof(psl_keyline));
    lineVec2d.clear();
    for (int i = 0; i < n; ++i) lineVec2d.push_back(Eigen::Vector3d(eq_[3 * i], eq_[3 * i + 1], eq_[3 * i + 2]));
    if (n == 0) {
      descriptors.release();
      return;
    }
    descriptors.create(n, 32, CV_8U);
    cv::Mat out = descriptors.getMat();
    for (int i = 0; i < n; ++i) std::memcpy(out.ptr(i), &desc_[(size_t)i * 32], 32);
  }

  int GetLevels() { return numOctaves; }
  float GetScaleFactor() { return scale; }
  std::vector<float> GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> GetInverseScaleFactors() { return mvInvScaleFactor; }
  std::vector<float> GetScaleSigmaSquares() { return mvLevelSigma2; }
  std::vector<float> GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

 protected:
  psl_ctx* ctx_ = nullptr;
  int numOctaves = 1;
  float scale = 1.2f;
  unsigned int nLSDFeature = 200;
  double min_line_length = 0;
  std::vector<psl_keyline> kl_;
  std::vector<uint8_t> desc_;
  std::vector<double> eq_;
  std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
};

}  // namespace ORB_SLAM2
