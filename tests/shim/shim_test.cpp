// TEST INFRASTRUCTURE — drives include/psl_orbslam_shim.hpp exactly as Frame::ExtractORB / Frame::ExtractLSD do
// (src/Frame.cc:311-317, 489-494) and hands the results to the Python test through a C entry point.
#include <opencv2/core/core.hpp>
#include <opencv2/line_descriptor.hpp>
#include <Eigen/Core>

#include "psl_orbslam_shim.hpp"

extern "C" int shim_extract(const uint8_t* gray, int w, int h, int nfeatures, float scale, int nlevels, int ini, int mn,
                            psl_keypoint* kps, uint8_t* desc, int cap, psl_keyline* kl, uint8_t* ldesc, double* eq,
                            int line_cap, int* n_lines) {
  try {
    ORB_SLAM2::ORBextractor* orb = new ORB_SLAM2::ORBextractor(nfeatures, scale, nlevels, ini, mn);
    ORB_SLAM2::LINEextractor* lsd = new ORB_SLAM2::LINEextractor(1, 1.2f, (unsigned)line_cap, 0.0);
    const cv::Mat im(h, w, CV_8UC1, (void*)gray, (size_t)w);
    std::vector<cv::KeyPoint> mvKeys;
    cv::Mat mDescriptors;
    (*orb)(im, cv::Mat(), mvKeys, mDescriptors);                       // Frame.cc:314
    std::vector<cv::line_descriptor::KeyLine> mvKeylinesUn;
    cv::Mat mLdesc, mask;
    std::vector<Eigen::Vector3d> mvKeyLineFunctions;
    (*lsd)(im, mask, mvKeylinesUn, mLdesc, mvKeyLineFunctions);        // Frame.cc:494
    const int n = (int)mvKeys.size();
    if (n > cap || (int)mvKeylinesUn.size() > line_cap) return -2;
    if (n && (mDescriptors.rows != n || mDescriptors.cols != 32)) return -3;
    for (int i = 0; i < n; ++i) {
      std::memcpy(&kps[i], &mvKeys[i], sizeof(psl_keypoint));
      std::memcpy(desc + (size_t)i * 32, mDescriptors.ptr(i), 32);
    }
    *n_lines = (int)mvKeylinesUn.size();
    for (int i = 0; i < *n_lines; ++i) {
      std::memcpy(&kl[i], &mvKeylinesUn[i], sizeof(psl_keyline));
      std::memcpy(ldesc + (size_t)i * 32, mLdesc.ptr(i), 32);
      for (int k = 0; k < 3; ++k) eq[3 * i + k] = mvKeyLineFunctions[i](k);
    }
    if (orb->GetLevels() != nlevels || orb->GetScaleFactors().size() != (size_t)nlevels) return -4;
    // an empty image returns silently and leaves the outputs alone (ORBextractor.cc:1046-1047)
    std::vector<cv::KeyPoint> keep(3);
    cv::Mat none;
    (*orb)(cv::Mat(), cv::Mat(), keep, none);
    if (keep.size() != 3) return -5;
    delete orb;
    delete lsd;
    return n;
  } catch (const std::exception&) {
    return -1;
  }
}
