/*
 * psl_oracle.h — TEST INFRASTRUCTURE.  CPU restatement of the reference's
 * front-end algorithms; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product
 * (libpsl_frontend.so) never links or calls anything declared here.
 *
 * Parity status: the reference cannot be compiled here (needs OpenCV 3.x +
 * contrib headers) and holds no tests or golden vectors for this path
 * (SURVEY.md §4, §8c).  The oracle is pinned against goldens produced by
 * oracle/pyref/ (the reference's control flow over the real cv2 4.13
 * primitives) — "pinned to cv2-primitive goldens, unpinned by the reference".
 */
#ifndef PSL_ORACLE_H
#define PSL_ORACLE_H
#include <stdint.h>

#include "../../include/psl_frontend.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_orb_params {
  int32_t nfeatures;
  float scale_factor;
  int32_t nlevels, ini_th, min_th;
} orc_orb_params;

/* primitives */
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride);
void orc_gauss_blur_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride, int ksize);
float orc_fast_atan2(float y, float x);
/* FAST-9/16 score of every pixel of a w×h image (0 where not evaluable / not a corner at `th`) */
void orc_fast_score(const uint8_t* img, int w, int h, int stride, int th, uint8_t* score);

/* ctor tables, ORBextractor.cc:410-470; arrays nlevels long */
void orc_orb_tables(const orc_orb_params* p, float* scale, float* inv_scale, int32_t* quota, int32_t* umax16);
void orc_orb_level_size(const orc_orb_params* p, int w, int h, int level, int* lw, int* lh);

/* stage functions; coordinates of candidates/selected are relative to minBorder (16) as in the reference */
int orc_fast_cells(const uint8_t* img, int w, int h, int stride, int ini_th, int min_th, float* xyr, int cap);
int orc_octree(const float* xyr, int n, int min_x, int max_x, int min_y, int max_y, int N, float* out_xyr, int cap);
float orc_ic_angle(const uint8_t* img, int stride, int x, int y);
void orc_brief(const uint8_t* blur, int stride, float x, float y, float angle_deg, uint8_t* desc32);

/* ORBextractor::operator(), ORBextractor.cc:1043-1105.  Returns 0 or <0. */
int orc_orb_extract(const orc_orb_params* p, const uint8_t* gray, int w, int h, int stride, psl_keypoint* kps,
                    uint8_t* desc, int cap, int* n);

/* ---- matchers (orc_match.cpp): ORBmatcher.cc + Frame.cc grid helpers ---- */
void orc_descriptor_distance(const uint8_t* a, const uint8_t* b, int n, int32_t* dist);
void orc_hamming_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx, int32_t* dist);
int orc_features_in_area(const psl_frame_view* f, float x, float y, float r, int minLevel, int maxLevel, int32_t* out,
                         int cap);
int orc_match_projection(const psl_frame_view* f, const psl_proj_query* qs, const uint8_t* qdesc, int nq,
                         const uint8_t* claimed_in, const psl_match_params* p, int32_t* assign, int32_t* nmatches);
int orc_match_bow(const uint8_t* kf_desc, const float* kf_angle, const uint8_t* kf_valid, int nkf,
                  const psl_feature_vector* kfv, const uint8_t* f_desc, const float* f_angle, int nf,
                  const psl_feature_vector* ffv, float nn_ratio, int th_low, int check_orientation, int32_t* match_f,
                  int32_t* nmatches);

int orc_match_triangulation(const psl_keyframe_view* kf1, const psl_feature_vector* fv1, const psl_keyframe_view* kf2,
                            const psl_feature_vector* fv2, const float* F12, float ex, float ey,
                            const float* scale_factors2, const float* level_sigma2_2, int only_stereo,
                            int check_orientation, int th_low, int32_t* matches12, int32_t* nmatches);

int orc_match_fuse(const psl_frame_view* kf, const psl_fuse_query* qs, const uint8_t* qdesc, int nq,
                   const float* inv_level_sigma2, int th_low, int32_t* best_idx, int32_t* best_dist);

/* loop closing / initialisation: ORBmatcher.cc:522-655, 1102-1326, 405-520 */
int orc_match_bow_kf(const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int n1,
                     const psl_feature_vector* fv1, const uint8_t* desc2, const float* angle2, const uint8_t* valid2,
                     int n2, const psl_feature_vector* fv2, float nn_ratio, int th_low, int check_orientation,
                     int32_t* matches12, int32_t* nmatches);
int orc_match_sim3(const psl_frame_view* kf1, const psl_frame_view* kf2, const psl_fuse_query* q12,
                   const uint8_t* mp_desc1, const psl_fuse_query* q21, const uint8_t* mp_desc2, int th_high,
                   int32_t* matches12, int32_t* nfound);
int orc_match_initialization(const psl_keypoint* kps1_un, const uint8_t* desc1, int n1, float* prev_matched,
                             const psl_frame_view* f2, int window_size, float nn_ratio, int th_low,
                             int check_orientation, int32_t* matches12, int32_t* nmatches);

/* ---- line junctions (orc_junction.cpp): PartiallyRecoverConnectivity.cpp:14-133, Frame.cc:380-472 ---- */
int orc_line_junctions(const psl_keyline* kl_un, const double* lines3d, int n, int img_w, int img_h, float radius,
                       float fan_thr, float* fans, psl_line_junction* junctions, int cap, int32_t* n_fans,
                       int32_t* n_junctions);

/* ---- pose-only optimisation (orc_pose.cpp): Optimizer.cc:239-1023 over the vendored g2o, point and LIL edges; UNPINNED ---- */
int orc_pose_optimization_lil(const float* Tcw_in, const psl_pose_point* pts, int n, const psl_pose_lil* lils, int n_lil,
                              float fx, float fy, float cx, float cy, float bf, float* Tcw_out, uint8_t* outlier,
                              uint8_t* lil_outlier);
int orc_pose_optimization(const float* Tcw_in, const psl_pose_point* pts, int n, float fx, float fy, float cx, float cy,
                          float bf, float* Tcw_out, uint8_t* outlier);

/* ---- Frame bookkeeping (orc_frame.cpp): Frame.cc:1342-1381, ORBmatcher.cc:1339-1393 ---- */
/* Frame::UndistortKeyPoints / ComputeImageBounds (Frame.cc:1062-1092, 1135-1163) over cv::undistortPoints */
void orc_undistort_keypoints(const psl_keypoint* kps, int n, const psl_distortion* cam, psl_keypoint* kps_un);
void orc_image_bounds(int cols, int rows, const psl_distortion* cam, float* bounds);
void orc_stereo_from_rgbd(const psl_keypoint* kps, int n, const uint16_t* depth, int w, int h, int stride_px,
                          float depth_factor, float bf, float* u_right, float* z);
void orc_queries_from_last_frame(const psl_keypoint* kps_last, const float* z_last, const uint8_t* valid_in,
                                 const uint8_t* claims_in, int n, const float* Tcw_last, const float* Tcw_cur,
                                 const float* cam, const float* scale_factors, float th, int mono, float min_x,
                                 float min_y, float max_x, float max_y, psl_proj_query* q);

/* ---- lines (orc_lsd.cpp ...): cv::LineSegmentDetector(LSD_REFINE_STD) behind LineExtractor.cpp:336-337 ---- */
int orc_lsd_detect(const uint8_t* img, int w, int h, int stride, int order_mode, float* lines, int cap);
int orc_lsd_scaled_image(const uint8_t* img, int w, int h, int stride, uint8_t* out, int* W, int* H);
int orc_line_iterator_count(int w, int h, float x1, float y1, float x2, float y2);
int orc_merge_lines_lsd(const float* lines, int n, float* out, int cap);
void orc_clamp_segments(float* lines, int n, int w, int h);
int orc_make_keylines(const float* lines, int n, int w, int h, int nfeatures, psl_keyline* kl);
void orc_lbd_gradients(const uint8_t* img, int w, int h, int stride, int16_t* dx, int16_t* dy);
void orc_lbd_one(const int16_t* dx, const int16_t* dy, int w, int h, const psl_keyline* kl, float* des72);
void orc_lbd_binarise(const float* des72, uint8_t* out32);
int orc_line_extract(const uint8_t* gray, int w, int h, int stride, int nfeatures, psl_keyline* kl, uint8_t* ldesc,
                     double* lineeq, float* lbd72, int cap, int* n_out);

/* ---- line matchers (orc_linematch.cpp): LSDmatcher.cpp, Frame.cc line grid, InsectlineMatch.cpp, Map.cc:204-272 ---- */
int orc_line_grid_cells(const psl_line_frame_view* f, int line, int32_t* cells, int cap);
int orc_lines_in_area(const psl_line_frame_view* f, float x1, float y1, float x2, float y2, float r, float TH,
                      int32_t* out, int cap);
int orc_line_match_nnr(const uint8_t* desc1, int n1, const uint8_t* desc2, int n2, float nnr, int32_t* matches12);
int orc_line_search_geom(const psl_keyline* kl_last, const uint8_t* desc_last, const uint8_t* has_ml, int n_last,
                         const psl_keyline* kl_cur, const uint8_t* desc_cur, int n_cur, const float* bounds,
                         float desc_th, int32_t* assign_cur);
void orc_line_frame_bf_match(const uint8_t* desc1, int n1, const uint8_t* desc2, int n2, float nn_ratio, float th,
                             int32_t* matches);
int orc_line_search_double(const uint8_t* desc1, int n1, const uint8_t* desc2, int n2, float nn_ratio, float th,
                           int32_t* matches12);
int orc_line_search_triangulation(const uint8_t* desc1, const uint8_t* has_ml1, int n1, const uint8_t* desc2,
                                  const uint8_t* has_ml2, int n2, float nn_ratio, float th, int is_double,
                                  int32_t* matches12);
int orc_line_search_triangulation_new(const psl_keyline* kl1, const uint8_t* desc1, const double* func1,
                                      const uint8_t* has_ml1, int n1, const psl_keyline* kl2, const uint8_t* desc2,
                                      const double* func2, const uint8_t* has_ml2, int n2, const float* F21,
                                      const float* F12, float nn_ratio, float th, int is_double, int32_t* pairs);
int orc_line_fuse(const psl_keyline* kl, int n_lines, const uint8_t* kf_desc, const psl_line_fuse_query* qs,
                  const uint8_t* qdesc, int nq, float th_cos, int th_low, int32_t* best_idx, int32_t* best_dist);
int orc_line_match_projection(const psl_line_frame_view* f, const psl_line_query* qs, const uint8_t* qdesc, int nq,
                              const uint8_t* claimed_in, int mode, float nn_ratio, int32_t* assign);
/* Frame::isLineGood (orc_line3d.cpp): Frame.cc:662-750, LineExtractor.cpp:27-323 */
void orc_lines_3d(const psl_keyline* kl, int n, const float* depth, int w, int h, int stride, const float* cam,
                  uint32_t seed, double* lines3d, float* line_eq);
int orc_plane_hypotheses(const psl_keyline* kl_un, const float* line_eq, const double* lines3d, int n_lines,
                         const psl_line_junction* js, int nj, double* le_l, float* planes, double* normals,
                         int32_t* junction_of, int cap);
int orc_plane_assoc(const float* planes_cam, const double* pts, int n_ljl, const float* Tcw, const float* map_planes,
                    const uint8_t* map_bad, int n_map, float d_th, float a_th, int mode, int32_t* assign);

/* CPU-baseline harness: B frames, one frame per task on `nthreads` std::threads. */
int orc_orb_extract_batch_mt(const orc_orb_params* p, const uint8_t* gray, int B, int w, int h, int stride,
                             int64_t frame_stride, int nthreads, int32_t* n_out, uint32_t* desc_xor);

int orc_track_batch_mt(const orc_orb_params* p, const uint8_t* gray, const uint16_t* depth, int B, int w, int h,
                       const float* Tcw, const float* cam, float th, float nn_ratio, int check_ori, int nthreads,
                       int32_t* n_out, int32_t* nmatches_out);

int orc_frontend_batch_mt(const orc_orb_params* p, const uint8_t* gray, const uint16_t* depth, int B, int w, int h,
                          const float* Tcw, const float* cam, float th, float nn_ratio, int check_ori,
                          int line_nfeatures, float line_desc_th, int nthreads, int32_t* n_out, int32_t* nmatches_out,
                          int32_t* nl_out, int32_t* line_nmatches_out);

#ifdef __cplusplus
}
#endif
#endif
