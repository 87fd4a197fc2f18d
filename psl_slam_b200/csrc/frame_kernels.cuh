#pragma once
#include "psl_common.cuh"

namespace psl {

struct QueryBuildParams {
  float th;  // SearchByProjection window factor (15 / 7 in Tracking.cc:1189-1193)
  int32_t mono;
  float min_x, min_y, max_x, max_y;
  float scale[kMaxLevels];  // mvScaleFactors
};

void launch_stereo(const psl_keypoint* kps, const int32_t* n, int cap, const uint16_t* depth, int stride_px,
                   int64_t frame_stride_px, float depth_factor, float bf, float* u_right, float* z, int B,
                   cudaStream_t st);
void launch_query_build(const psl_keypoint* kps, const float* z, const int32_t* n, int cap, const float* Tcw, int first,
                        const psl_camera& cam, const QueryBuildParams& prm, psl_proj_query* q, int32_t* nq, int B,
                        cudaStream_t st);

// correspondences of PoseOptimization after SearchByProjection(Current, Last) + the 4x4 prior poses (Tracking.cc:1193-1214)
void launch_pose_points(const psl_keypoint* kps, const float* u_right, const float* z, const int32_t* assign,
                        const int32_t* n, int cap, const float* Tcw, const psl_camera& cam, const float* inv_sigma2,
                        psl_pose_point* pts, float* T44, int B, cudaStream_t st);

// Frame::UndistortKeyPoints (Frame.cc:1062-1092): kps_un[b][i] = kps[b][i] with cv::undistortPoints applied to pt
void launch_undistort(const psl_keypoint* kps, const int32_t* n, int cap, const psl_distortion& cam, psl_keypoint* kps_un,
                      int B, cudaStream_t st);

// K0: cvtColor(... -> GRAY) in OpenCV 4.x Q15 arithmetic and depth.convertTo(CV_32F, factor) (Tracking.cc:219-235)
void launch_color_to_gray(const uint8_t* color, int channels, int rgb_order, int color_stride, int64_t color_fs,
                          uint8_t* gray, int gray_stride, int64_t gray_fs, int B, int w, int h, cudaStream_t st);
void launch_depth_to_float(const uint16_t* in, int stride_px, int64_t fs_px, float factor, float* out, int B, int w,
                           int h, cudaStream_t st);

}  // namespace psl
