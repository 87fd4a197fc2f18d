/*
 * psl_frontend.h — C-ABI of the B200-native PSL-SLAM feature front end.
 *
 * Plain pointers and sizes only.  Every entry point names the reference
 * interface it replaces (file:line relative to the PSL-SLAM tree).  The
 * library (libpsl_frontend.so) is CUDA-only: there is no CPU fallback; every
 * call fails with PSL_E_CUDA when no sm_100 device is usable.
 *
 * Conventions
 *   - all functions return 0 (PSL_OK) or a negative PSL_E_* code and never throw;
 *     psl_last_error(ctx) returns a human-readable reason for the last failure.
 *   - the caller owns every buffer and passes capacities; the callee writes counts.
 *   - a psl_ctx is bound to one GPU and one stream and is NOT re-entrant — the same
 *     rule as the reference extractors (stateful mvImagePyramid member,
 *     include/ORBextractor.h:85; one extractor per thread, src/Frame.cc:92-93).
 *   - "_dev" variants take device pointers (inputs already resident in HBM) and
 *     enqueue on the ctx stream; the non-_dev variants take HOST pointers and
 *     include the H2D/D2H copies.
 */
#ifndef PSL_FRONTEND_H
#define PSL_FRONTEND_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSL_OK 0
#define PSL_E_INVALID (-1)   /* bad argument (null, size, stride, aspect) */
#define PSL_E_CUDA (-2)      /* CUDA runtime failure / no device */
#define PSL_E_CAPACITY (-3)  /* caller buffer or ctx limit too small */
#define PSL_E_INTERNAL (-4)

typedef struct psl_ctx psl_ctx;

/* Same field order and size (28 B) as cv::KeyPoint, the element type of
 * ORBextractor::operator()'s `keypoints` (include/ORBextractor.h:59-61). */
typedef struct psl_keypoint {
  float x, y;     /* pt, level-0 pixel coordinates (ORBextractor.cc:1095-1101) */
  float size;     /* 31*scale[octave] truncated (ORBextractor.cc:837) */
  float angle;    /* degrees, IC_Angle (ORBextractor.cc:77-104) */
  float response; /* FAST score */
  int32_t octave;
  int32_t class_id; /* always -1 */
} psl_keypoint;

/* Same field order and size (68 B) as cv::line_descriptor::KeyLine
 * (Thirdparty/line_descriptor/include/line_descriptor/descriptor_custom.hpp:107-146), the element type of
 * LINEextractor::operator()'s `keylines` (add_inc/LineExtractor.h:167). */
typedef struct psl_keyline {
  float angle;      /* atan2(dy, dx) of the segment */
  int32_t class_id; /* index of the line after the top-N filter (LineExtractor.cpp:346-347) */
  int32_t octave;   /* always 0 (uselongline.cpp:419) */
  float pt_x, pt_y; /* midpoint */
  float response;   /* length / max(cols, rows) */
  float size;       /* (ex-sx)*(ey-sy) */
  float start_x, start_y, end_x, end_y;             /* extremes in the original image */
  float s_oct_x, s_oct_y, e_oct_x, e_oct_y;         /* extremes in the octave image (same: octave 0) */
  float line_length;
  int32_t num_pixels; /* cv::LineIterator(img, start, end).count */
} psl_keyline;

/* Mirrors the YAML keys read in src/Tracking.cc:113-127
 * (ORBextractor.* / LINEextractor.*; Examples/RGB-D/TUM1.yaml:42-63). */
typedef struct psl_config {
  int32_t device;          /* CUDA ordinal */
  int32_t max_width;       /* largest frame the ctx must accept */
  int32_t max_height;
  int32_t max_batch;       /* frames in flight per *_batch call */
  int32_t orb_nfeatures;   /* ORBextractor.nFeatures   (1000) */
  float orb_scale_factor;  /* ORBextractor.scaleFactor (1.2)  */
  int32_t orb_nlevels;     /* ORBextractor.nLevels     (8)    */
  int32_t orb_ini_th_fast; /* ORBextractor.iniThFAST   (20)   */
  int32_t orb_min_th_fast; /* ORBextractor.minThFAST   (7)    */
  int32_t orb_max_candidates; /* FAST candidate pool per frame (the reference only reserves nfeatures*10 as a hint,
                                 ORBextractor.cc:779).  > 0: hard bound, overflow = PSL_E_CAPACITY.  0 / < 0: starts at
                                 max(16384, 32 nfeatures) / at |value| and doubles when a frame overflows it: the
                                 host-pointer entry points then run the call again themselves, the device-pointer
                                 ones return PSL_E_CAPACITY once and succeed when called again */
  int32_t chunk_frames;    /* frames per launch of the ORB stages; 0 = auto (max_batch clamped to [1, 512]; ~2.4 MB of HBM per frame at 640x480) */
  int32_t line_nfeatures;  /* LINEextractor.nFeatures  (200)  */
  float line_scale_factor; /* LINEextractor.scaleFactor (1.2; truncated to int 1 by the reference) */
  int32_t line_nlevels;    /* LINEextractor.nLevels    (1)    */
  float line_min_length;   /* LINEextractor.min_line_length (0; read by Tracking.cc:126 but never used by
                              LINEextractor::operator(), add_src/LineExtractor.cpp:325-366 -> ignored here too) */
  int32_t line_chunk_frames; /* frames per launch of the line stages; 0 = auto (max_batch clamped to [64, 4096]).
                                The LSD core runs one frame per warp, so the batch is its parallel axis;
                                the line buffers take about 7.5 MB of HBM per frame of the chunk at 640x480 */
  int32_t line_max_raw;    /* raw LSD segments kept per frame before the merge: > 0 hard bound, 0 / < 0 starts at
                              4096 / |value| and grows like orb_max_candidates (at most 65535) */
} psl_config;

void psl_default_config(psl_config* cfg);

/* ORBextractor::ORBextractor (src/ORBextractor.cc:410-470) + LINEextractor ctor
 * (add_src/LineExtractor.cpp:6-25): tables, device buffers, stream. */
int psl_create(const psl_config* cfg, psl_ctx** out);
void psl_destroy(psl_ctx* ctx);
const char* psl_last_error(const psl_ctx* ctx);
/* cudaStream_t of the ctx, as void* (for callers that order their own work after ours). */
void* psl_stream(psl_ctx* ctx);
int psl_sync(psl_ctx* ctx);

/* ORBextractor::GetLevels/GetScaleFactors/GetInverseScaleFactors/GetScaleSigmaSquares/
 * GetInverseScaleSigmaSquares (include/ORBextractor.h:63-83); each array nlevels long (may be NULL). */
int psl_orb_tables(const psl_ctx* ctx, int32_t* nlevels, float* scale, float* inv_scale, float* sigma2,
                   float* inv_sigma2, int32_t* features_per_level);

/* ORBextractor::operator()(image, mask, keypoints, descriptors)
 * (src/ORBextractor.cc:1043-1105; called from Frame::ExtractORB, src/Frame.cc:311-317).
 * gray: HOST CV_8UC1, `stride` bytes per row.  mask is ignored by the reference and absent here.
 * kps[cap], desc[cap*32]; *n = number written (≤ nfeatures + 3*nlevels).
 * Empty image (w==0||h==0) → *n = 0, PSL_OK (the reference returns silently, :1046-1047). */
int psl_orb_extract(psl_ctx* ctx, const uint8_t* gray, int32_t w, int32_t h, int32_t stride, psl_keypoint* kps,
                    uint8_t* desc, int32_t cap, int32_t* n);

/* Batched form of the same call: B frames of identical size, frame b at gray + b*frame_stride
 * (HOST).  Outputs are [B][cap] row blocks; n[b] = count of frame b. */
int psl_orb_extract_batch(psl_ctx* ctx, const uint8_t* gray, int32_t B, int32_t w, int32_t h, int32_t stride,
                          int64_t frame_stride, psl_keypoint* kps, uint8_t* desc, int32_t cap, int32_t* n);

/* Same, all pointers DEVICE memory; asynchronous on psl_stream(ctx). */
int psl_orb_extract_batch_dev(psl_ctx* ctx, const uint8_t* d_gray, int32_t B, int32_t w, int32_t h, int32_t stride,
                              int64_t frame_stride, psl_keypoint* d_kps, uint8_t* d_desc, int32_t cap,
                              int32_t* d_n);

/* Camera matrix and distortion of Tracking's settings (mK, mDistCoef: src/Tracking.cc:52-77): k1, k2, p1, p2, k3. */
typedef struct psl_distortion {
  float fx, fy, cx, cy;
  float k1, k2, p1, p2, k3;
} psl_distortion;

/* Frame::UndistortKeyPoints (src/Frame.cc:1062-1092; SURVEY "next" row N3): mvKeysUn = mvKeys with pt replaced by
 * cv::undistortPoints(pt, mK, mDistCoef, Mat(), mK) — OpenCV's five fixed-point iterations in double on the
 * normalised coordinates, re-projected with mK and rounded to float; with k1 == 0 the keypoints are copied (:1064-1068).
 * HOST pointers; kps_un may alias kps. */
int psl_undistort_keypoints(psl_ctx* ctx, const psl_keypoint* kps, int32_t n, const psl_distortion* cam,
                            psl_keypoint* kps_un);
/* Batched, DEVICE pointers, asynchronous: frame b owns rows [b*cap, b*cap + d_n[b]) of both arrays. */
int psl_undistort_keypoints_dev(psl_ctx* ctx, const psl_keypoint* d_kps, const int32_t* d_n, int32_t cap, int32_t B,
                                const psl_distortion* cam, psl_keypoint* d_kps_un);
/* Frame::ComputeImageBounds (src/Frame.cc:1135-1163): the undistorted image corners give
 * bounds[4] = mnMinX, mnMinY, mnMaxX, mnMaxY (the order psl_frame_view uses); k1 == 0: 0, 0, cols, rows. */
int psl_image_bounds(psl_ctx* ctx, int32_t cols, int32_t rows, const psl_distortion* cam, float* bounds);

/* Tracking::GrabImageRGBD input conversion (src/Tracking.cc:219-235; SURVEY "next" row N3, kernel K0):
 *   cvtColor(RGB|BGR|RGBA|BGRA -> GRAY)  Y = (R*9798 + G*19235 + B*3735 + 16384) >> 15   (OpenCV 4.x Q15 arithmetic)
 *   imDepth.convertTo(CV_32F, mDepthMapFactor)
 * Batched, DEVICE pointers, asynchronous.  color: B frames of h rows, `channels` (3 or 4) interleaved bytes per
 * pixel, `color_stride` bytes per row, frames `color_frame_stride` bytes apart; rgb_order != 0 for mbRGB (R first).
 * gray: [B][h][gray_stride] u8.  depth_in u16 (strides in pixels) -> depth_out float [B][h][w]; either half of the
 * call may be skipped by passing NULL. */
int psl_convert_rgbd_dev(psl_ctx* ctx, const uint8_t* d_color, int32_t channels, int32_t rgb_order, int32_t color_stride,
                         int64_t color_frame_stride, uint8_t* d_gray, int32_t gray_stride, int64_t gray_frame_stride,
                         const uint16_t* d_depth_in, int32_t depth_stride_px, int64_t depth_frame_stride_px,
                         float depth_factor, float* d_depth_out, int32_t B, int32_t w, int32_t h);
/* HOST pointers, tightly packed frames ([B][h][w][channels] -> [B][h][w]). */
int psl_convert_rgbd(psl_ctx* ctx, const uint8_t* color, int32_t channels, int32_t rgb_order, uint8_t* gray,
                     const uint16_t* depth_in, float depth_factor, float* depth_out, int32_t B, int32_t w, int32_t h);

/* LINEextractor::operator()(image, mask, keylines, descriptors, lineVec2d)
 * (add_src/LineExtractor.cpp:325-366; called from Frame::ExtractLSD, src/Frame.cc:494):
 *   LSDDetector::detect(img, kl, scale -> int 1, numOctaves 1)   :336-337  (cv::LineSegmentDetector, REFINE_STD)
 *   optimizeAndMergeLines_lsd                                      :338      (add_src/uselongline.cpp:449-485)
 *   sort by response, keep line_nfeatures, class_id = i            :342-348
 *   BinaryDescriptor::compute (LBD, 32 bytes per line)             :349-350
 *   2-D line equations sp x ep / |(l0, l1)|                        :352-363
 * gray: HOST CV_8UC1.  mask must be empty in the reference's only call (Frame.cc:493) and is absent here.
 * kl[cap], ldesc[cap*32], lineeq[cap*3] (fp64, lineVec2d), lbd72[cap*72] (optional, may be NULL: the float
 * descriptor before binarisation).  *n = number of lines.  Empty image -> *n = 0, PSL_OK (:327-328). */
int psl_line_extract(psl_ctx* ctx, const uint8_t* gray, int32_t w, int32_t h, int32_t stride, psl_keyline* kl,
                     uint8_t* ldesc, double* lineeq, float* lbd72, int32_t cap, int32_t* n);

/* Batched form: B frames of identical size, frame b at gray + b*frame_stride (HOST); outputs are [B][cap] blocks. */
int psl_line_extract_batch(psl_ctx* ctx, const uint8_t* gray, int32_t B, int32_t w, int32_t h, int32_t stride,
                           int64_t frame_stride, psl_keyline* kl, uint8_t* ldesc, double* lineeq, float* lbd72,
                           int32_t cap, int32_t* n);

/* Same, all pointers DEVICE memory; asynchronous on psl_stream(ctx). */
int psl_line_extract_batch_dev(psl_ctx* ctx, const uint8_t* d_gray, int32_t B, int32_t w, int32_t h, int32_t stride,
                               int64_t frame_stride, psl_keyline* d_kl, uint8_t* d_ldesc, double* d_lineeq,
                               float* d_lbd72, int32_t cap, int32_t* d_n);

/* ------------------------------------------------------------------------------------------------
 * Matching.  Only plain arrays cross the boundary: the caller (Frame / Tracking) keeps the MapPoint
 * objects and passes what the matchers read from them.
 * ------------------------------------------------------------------------------------------------ */

#define PSL_GRID_COLS 64 /* FRAME_GRID_COLS, include/Frame.h:46 */
#define PSL_GRID_ROWS 48 /* FRAME_GRID_ROWS, include/Frame.h:45 */

/* What the matchers read from the searched Frame (include/Frame.h:150-290): undistorted keypoints
 * (mvKeysUn), right coordinates (mvuRight; <=0 = none; may be NULL), descriptors (mDescriptors) and the
 * image bounds that define the 64x48 feature grid (mnMinX.., mfGridElementWidthInv..; Frame.cc:160-166).
 * The grid itself (Frame::AssignFeaturesToGrid, Frame.cc:269-284) is rebuilt by the library. */
typedef struct psl_frame_view {
  int32_t n;
  const psl_keypoint* kps_un;
  const float* u_right;
  const uint8_t* desc; /* n x 32 */
  float min_x, min_y, max_x, max_y;
  float grid_w_inv, grid_h_inv;
} psl_frame_view;

#define PSL_Q_VALID 1u  /* query takes part (MapPoint exists, not outlier, in frustum ...) */
#define PSL_Q_CLAIMS 2u /* its MapPoint has Observations()>0: a kp assigned to it is skipped by later queries
                           (ORBmatcher.cc:87-89, 1403-1405) */

/* One projected point to be matched: what SearchByProjection computes per MapPoint before calling
 * Frame::GetFeaturesInArea (ORBmatcher.cc:62-71 and :1365-1393). */
typedef struct psl_proj_query {
  float u, v;        /* projection in the searched frame */
  float radius;      /* window half-size, already scaled (th*scale[octave], r*scale[level]) */
  int32_t min_level; /* GetFeaturesInArea(minLevel,maxLevel): -1/-1 = no gating (Frame.cc:1006) */
  int32_t max_level;
  float u_right;     /* predicted right coordinate (u - bf/z, mTrackProjXR) */
  float angle;       /* keypoint angle used by the rotation histogram */
  uint32_t flags;    /* PSL_Q_* */
} psl_proj_query;

typedef struct psl_match_params {
  int32_t mode;              /* 0: best only, rotation histogram (ORBmatcher.cc:1328-1470)
                                1: best + second with same-level ratio test (ORBmatcher.cc:45-129) */
  int32_t th_dist;           /* TH_HIGH = 100 (ORBmatcher.cc:37) */
  float nn_ratio;            /* mfNNratio */
  int32_t check_orientation; /* mbCheckOrientation */
} psl_match_params;

/* ORBmatcher::DescriptorDistance (ORBmatcher.cc:1647-1663) for n pairs a[i], b[i] of 32-byte descriptors. */
int psl_descriptor_distance(psl_ctx* ctx, const uint8_t* a, const uint8_t* b, int32_t n, int32_t* dist);

/* cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, 2) as used by LSDmatcher::matchNNR / FrameBFMatch
 * (add_src/LSDmatcher.cpp:354-376, 492-516): idx/dist are nq x 2, sorted by distance, earliest train
 * index wins ties; missing neighbours (nt < 2) are -1. */
int psl_hamming_knn2(psl_ctx* ctx, const uint8_t* q, int32_t nq, const uint8_t* t, int32_t nt, int32_t* idx,
                     int32_t* dist);

/* ORBmatcher::SearchByProjection, both the Frame<-LastFrame form (ORBmatcher.cc:1328-1470, mode 0) and the
 * Frame<-local-map-points form (:45-129, mode 1).  claimed_in[n] (may be NULL) marks keypoints whose current
 * MapPoint has Observations()>0 before the call.  assign[n] = index of the query whose point ends up in
 * mvpMapPoints[i], or -1; *nmatches = the function's return value.
 * The relocalization overload SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (:1472-1599,
 * Tracking.cc:2142,2156) is mode 0 with other settings: frame->u_right = NULL (it has no right-coordinate gate),
 * every valid query flagged PSL_Q_CLAIMS and claimed_in[i] = (CurrentFrame.mvpMapPoints[i] != NULL) (:1545-1546 skips
 * any keypoint that holds a MapPoint, old or just assigned), levels pred-1 .. pred+1, angle = pKF->mvKeysUn[i].angle,
 * th_dist = ORBdist (tests: reloc_* goldens, made by a restatement of that overload alone). */
int psl_match_projection(psl_ctx* ctx, const psl_frame_view* frame, const psl_proj_query* queries,
                         const uint8_t* query_desc, int32_t nq, const uint8_t* claimed_in,
                         const psl_match_params* params, int32_t* assign, int32_t* nmatches);

/* ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) (ORBmatcher.cc:159-288).  The two FeatureVectors
 * (DBoW2 std::map<NodeId, vector<uint>>) are passed as CSR: node ids ascending, offs[n_nodes+1], indices.
 * kf_valid[i] != 0 iff the keyframe keypoint has a good MapPoint.  match_f[nf] = keyframe keypoint index
 * matched to frame keypoint i, or -1. */
typedef struct psl_feature_vector {
  int32_t n_nodes;
  const uint32_t* node_id;
  const int32_t* offs;
  const uint32_t* idx;
} psl_feature_vector;
int psl_match_bow(psl_ctx* ctx, const uint8_t* kf_desc, const float* kf_angle, const uint8_t* kf_valid, int32_t nkf,
                  const psl_feature_vector* kf_fv, const uint8_t* f_desc, const float* f_angle, int32_t nf,
                  const psl_feature_vector* f_fv, float nn_ratio, int32_t th_low, int32_t check_orientation,
                  int32_t* match_f, int32_t* nmatches);

/* What the KeyFrame-to-KeyFrame matchers read from a KeyFrame (include/KeyFrame.h): mvKeysUn, mvuRight (< 0 = no
 * stereo observation), mDescriptors and, per keypoint, whether GetMapPoint(i) is non-null. */
typedef struct psl_keyframe_view {
  int32_t n;
  const psl_keypoint* kps_un;
  const float* u_right;
  const uint8_t* desc;
  const uint8_t* has_mappoint;
} psl_keyframe_view;

/* ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo) (ORBmatcher.cc:657-823; "next" row
 * N1, called by LocalMapping::CreateNewMapPoints, LocalMapping.cc:336): within equal vocabulary nodes, for every
 * KF1 keypoint without a MapPoint the KF2 keypoint without a MapPoint of smallest Hamming distance <= th_low that
 * is far enough from the epipole (mono-mono only, :739-745) and within 3.84 sigma^2 of the epipolar line
 * (CheckDistEpipolarLine, :140-157); the last candidate wins distance ties (`dist > bestDist` skips); rotation
 * histogram as elsewhere.  F12: 3x3 row-major; (ex, ey): epipole in image 2 (:663-670, computed by the caller from
 * the poses); scale_factors2 / level_sigma2_2: pKF2->mvScaleFactors / mvLevelSigma2.
 * matches12[kf1.n] = KF2 index or -1 (vMatchedPairs = the non-negative entries in index order); *nmatches = return. */
int psl_match_triangulation(psl_ctx* ctx, const psl_keyframe_view* kf1, const psl_feature_vector* fv1,
                            const psl_keyframe_view* kf2, const psl_feature_vector* fv2, const float* F12, float ex,
                            float ey, const float* scale_factors2, const float* level_sigma2_2, int32_t nlevels,
                            int32_t only_stereo, int32_t check_orientation, int32_t th_low, int32_t* matches12,
                            int32_t* nmatches);

/* One MapPoint projected into a KeyFrame by ORBmatcher::Fuse before its window search (ORBmatcher.cc:846-892):
 * the caller keeps the MapPoint tests (isBad, IsInKeyFrame, depth, IsInImage, distance invariance, viewing angle)
 * and PredictScale; flags = PSL_Q_VALID when all passed. */
typedef struct psl_fuse_query {
  float u, v;          /* projection */
  float u_right;       /* u - bf * invz */
  float radius;        /* th * mvScaleFactors[nPredictedLevel] */
  int32_t pred_level;  /* nPredictedLevel */
  uint32_t flags;
} psl_fuse_query;

/* The matching part of ORBmatcher::Fuse(pKF, vpMapPoints, th) (ORBmatcher.cc:825-981; "next" row N1, called by
 * LocalMapping::SearchInNeighbors, LocalMapping.cc:796,821): per MapPoint the keypoint of KeyFrame::GetFeaturesInArea
 * (KeyFrame.cc:685-724, no level gate) with the smallest Hamming distance among those at level pred-1..pred whose
 * reprojection error passes the chi-square gate (7.8 stereo / 5.99 mono, :906-931); earliest candidate wins ties.
 * best_idx[nq] = keypoint index if that distance <= th_low (TH_LOW = 50), else -1; best_dist[nq] (may be NULL).
 * The replace-or-add bookkeeping on the MapPoint objects (:953-975) stays with the caller; every query is
 * independent.  inv_level_sigma2: pKF->mvInvLevelSigma2 (nlevels entries); NULL = no chi-square gate (the Sim3 form, below). */
int psl_match_fuse(psl_ctx* ctx, const psl_frame_view* kf, const psl_fuse_query* queries, const uint8_t* query_desc,
                   int32_t nq, const float* inv_level_sigma2, int32_t nlevels, int32_t th_low, int32_t* best_idx,
                   int32_t* best_dist);

/* The Sim3 form, ORBmatcher::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (ORBmatcher.cc:983-1100; "next" row N1, caller
 * LoopClosing::SearchAndFuse, LoopClosing.cc:599), has the same window search without the chi-square gate (:1046-1075):
 * call psl_match_fuse with inv_level_sigma2 = NULL (u_right of the queries and of kf is then not read).  Tests: the
 * qfuse / fuse_idx arrays of the loop_pair* goldens.
 *
 * ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th) (ORBmatcher.cc:290-403, caller
 * LoopClosing::ComputeSim3, LoopClosing.cc:375) is psl_match_projection mode 0 with: frame->u_right = NULL,
 * check_orientation = 0, th_dist = TH_LOW (50), every valid query flagged PSL_Q_CLAIMS, min_level = pred-1,
 * max_level = pred (the loop's level test :376-379 selects the same keypoints in the same order as the level gate of
 * Frame::GetFeaturesInArea), claimed_in[i] = (vpMatched[i] != NULL) (:370-371, which also skips the keypoints this call
 * has just filled).  assign[i] = the query whose point this call wrote into vpMatched[i].  Tests: qproj / proj_assign. */

/* ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12) (ORBmatcher.cc:522-655; "next" row N1, caller
 * LoopClosing::ComputeSim3, LoopClosing.cc:265).  As psl_match_bow, but both sides are keyframes: valid1 / valid2[i] != 0
 * iff the keypoint holds a MapPoint that is not bad (:561-566, :580-586), the acceptance is `bestDist1 < th_low`
 * (strict, :592) and the result is indexed by KF1: matches12[n1] = KF2 keypoint index or -1 (vpMatches12[idx1] =
 * vpMapPoints2[matches12[idx1]]); *nmatches = the return value. */
int psl_match_bow_kf(psl_ctx* ctx, const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int32_t n1,
                     const psl_feature_vector* fv1, const uint8_t* desc2, const float* angle2, const uint8_t* valid2,
                     int32_t n2, const psl_feature_vector* fv2, float nn_ratio, int32_t th_low,
                     int32_t check_orientation, int32_t* matches12, int32_t* nmatches);

/* ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th) (ORBmatcher.cc:1102-1326; "next" row N1, caller
 * LoopClosing::ComputeSim3, LoopClosing.cc:323).  The caller keeps the MapPoint side (:1140-1186, :1220-1262: the point
 * exists, is not bad, is not already matched, positive depth, IsInImage, distance invariance, PredictScale) and passes
 * one psl_fuse_query per keypoint: q12[i1] = the MapPoint of KF1 keypoint i1 projected into KF2 with radius
 * th * pKF2->mvScaleFactors[pred] (n = kf1->n queries, mp_desc1 = their descriptors), q21 likewise into KF1
 * (kf2->n queries); u_right is not read.  Both window searches (smallest Hamming distance at level pred-1..pred, first
 * candidate on ties, accepted if <= th_high = TH_HIGH) and the agreement test (:1306-1321) run on the device.
 * matches12[kf1->n] = KF2 keypoint index for the pairs found in both directions, else -1; *nfound = the return value. */
int psl_match_sim3(psl_ctx* ctx, const psl_frame_view* kf1, const psl_frame_view* kf2, const psl_fuse_query* q12,
                   const uint8_t* mp_desc1, const psl_fuse_query* q21, const uint8_t* mp_desc2, int32_t th_high,
                   int32_t* matches12, int32_t* nfound);

/* ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (ORBmatcher.cc:405-520; "next"
 * row N1, caller Tracking::MonocularInitialization, Tracking.cc:696).  kps1_un / desc1: F1.mvKeysUn / mDescriptors
 * (only level-0 keypoints search, :421-423); prev_matched [n1][2] = vbPrevMatched, read as the window centres and
 * updated in place for the matched keypoints (:514-517); f2 = the searched frame (window of half-size window_size over
 * its level-0 keypoints, :425).  The greedy loop is order dependent (vMatchedDistance / vnMatches21 un-match earlier
 * pairs, :441-442, :461-470): the candidates and their distances are found in parallel, the loop itself is replayed in
 * order by one warp.  matches12[n1] = F2 index or -1; *nmatches = the return value.  A window holding more than 256
 * level-0 keypoints of F2 is refused with PSL_E_CAPACITY (nothing is written to the outputs). */
int psl_match_initialization(psl_ctx* ctx, const psl_keypoint* kps1_un, const uint8_t* desc1, int32_t n1,
                             float* prev_matched, const psl_frame_view* f2, int32_t window_size, float nn_ratio,
                             int32_t th_low, int32_t check_orientation, int32_t* matches12, int32_t* nmatches);

/* ------------------------------------------------------------------------------------------------
 * Line matching (add_src/LSDmatcher.cpp, add_src/InsectlineMatch.cpp).  As for points, MapLine / InsectLine
 * objects stay with the caller; plain arrays cross the boundary.  All pointers HOST.
 * ------------------------------------------------------------------------------------------------ */

/* LSDmatcher::match -> matchNNR (LSDmatcher.cpp:354-413): BFMatcher(NORM_HAMMING).knnMatch(desc1, desc2, 2), keep
 * best if d0 < d1 * nnr.  matches12[n1] = index into desc2 or -1; *nmatches = return value.  With n2 < 2 the
 * reference indexes out of bounds (:369); here every query is unmatched. */
int psl_line_match_nnr(psl_ctx* ctx, const uint8_t* desc1, int32_t n1, const uint8_t* desc2, int32_t n2, float nnr,
                       int32_t* matches12, int32_t* nmatches);

/* LSDmatcher::SearchByGeomNApearance(CurrentFrame, LastFrame, desc_th) (LSDmatcher.cpp:36-110, called at
 * Tracking.cc:1183): matchNNR(last, cur, desc_th), then per last line with a MapLine (has_mapline_last[i] != 0):
 * reject if the matched current line has startPointX == 0, if |cos| of the 2-D directions < cos 20 deg, or if both
 * endpoints moved by more than 10 % of the image bounds.  assign_cur[n_cur] = last-frame line whose MapLine ends up
 * in CurrentFrame.mvpMapLines[i] (the last writer wins, as in the reference) or -1; *nmatches = return value.
 * bounds = mnMinX, mnMinY, mnMaxX, mnMaxY of the current frame. */
int psl_line_search_geom(psl_ctx* ctx, const psl_keyline* kl_last, const uint8_t* desc_last,
                         const uint8_t* has_mapline_last, int32_t n_last, const psl_keyline* kl_cur,
                         const uint8_t* desc_cur, int32_t n_cur, const float* bounds, float desc_th,
                         int32_t* assign_cur, int32_t* nmatches);

/* LSDmatcher::FrameBFMatch (LSDmatcher.cpp:492-516) with lineDescriptorMAD (:660-685): knn2, MAD-adaptive gap
 * threshold, d0 < th and d0 < nn_ratio * d1.  matches[n1] = index into desc2 or -1.  n1 == 0 or n2 < 2: all -1. */
int psl_line_frame_bf_match(psl_ctx* ctx, const uint8_t* desc1, int32_t n1, const uint8_t* desc2, int32_t n2,
                            float nn_ratio, float th, int32_t* matches);

/* LSDmatcher::SearchDouble(InitialFrame, CurrentFrame, LineMatches) (LSDmatcher.cpp:462-490): FrameBFMatch in both
 * directions (TH_LOW = 50) and mutual consistency.  matches12[n1]; *nmatches = return value.  The KeyFrame form
 * (:415-460) is the same call with the roles swapped plus the caller's MapLine lookup. */
int psl_line_search_double(psl_ctx* ctx, const uint8_t* desc1, int32_t n1, const uint8_t* desc2, int32_t n2,
                           float nn_ratio, float th, int32_t* matches12, int32_t* nmatches);

/* LSDmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs) (LSDmatcher.cpp:705-741, TH_LOW, mutual; called by
 * LocalMapping::CreateNewMapLines2, LocalMapping.cc:580) and its vector form (:743-779: TH_HIGH, mutual check only
 * when is_double): FrameBFMatch in both directions, then drop pairs where either line already has a MapLine
 * (has_mapline1 / has_mapline2).  matches12[n1] = line of KF2 or -1; *nmatches = return value. */
int psl_line_search_triangulation(psl_ctx* ctx, const uint8_t* desc1, const uint8_t* has_mapline1, int32_t n1,
                                  const uint8_t* desc2, const uint8_t* has_mapline2, int32_t n2, float nn_ratio, float th,
                                  int32_t is_double, int32_t* matches12, int32_t* nmatches);

/* LSDmatcher::SearchForTriangulationNew(pKF1, pKF2, vMatchedPairs, isDouble) over FrameBFMatchNew and mutualOverlap
 * (add_src/LSDmatcher.cpp:518-658, 783-824; "next" row N1 — the epipolar-overlap variant of the search above; nothing in
 * the reference calls it).  Per line of either keyframe: the best of cv::BFMatcher::knnMatch(k = 2) is kept if the
 * projections of its end points onto the matched line, taken along their epipolar lines F * p, overlap that line's segment
 * by more than 0.8 (mutualOverlap), its distance is < th and < nn_ratio * the second distance; then the mutual check
 * (is_double) and the MapLine gates as in psl_line_search_triangulation.  kl*: mvKeyLines; func*: mvKeyLineFunctions
 * (n*3 doubles); F21 = ComputeF12(pKF2, pKF1) maps points of KF1 to epipolar lines in KF2, F12 the reverse (3x3 row-major
 * float, :804-805, computed by the caller from the poses).  matched_pairs[n1] = KF2 line or -1; *nmatches = return. */
int psl_line_search_triangulation_new(psl_ctx* ctx, const psl_keyline* kl1, const uint8_t* desc1, const double* func1,
                                      const uint8_t* has_mapline1, int32_t n1, const psl_keyline* kl2,
                                      const uint8_t* desc2, const double* func2, const uint8_t* has_mapline2, int32_t n2,
                                      const float* F21, const float* F12, float nn_ratio, float th, int32_t is_double,
                                      int32_t* matched_pairs, int32_t* nmatches);

/* One MapLine projected into a KeyFrame by LSDmatcher::Fuse before its window search (LSDmatcher.cpp:862-918): the
 * caller keeps the MapLine tests (isBad, IsInKeyFrame, IsInImage of both endpoints, distance invariance, viewing
 * angle), the `return false` of the whole call on an endpoint behind the camera (:883-884), and PredictScale;
 * flags = PSL_Q_VALID when all passed. */
typedef struct psl_line_fuse_query {
  float u1, v1, u2, v2;  /* projected endpoints */
  float radius;          /* th * mvScaleFactorsLine[nPredictedLevel] */
  int32_t pred_level;    /* nPredictedLevel */
  uint32_t flags;
} psl_line_fuse_query;

/* The matching part of LSDmatcher::Fuse(pKF, vpMapLines, th) (LSDmatcher.cpp:847-984; "next" row N1, called by
 * LocalMapping::SearchInNeighbors, LocalMapping.cc:846,872): per MapLine, KeyFrame::GetLinesInArea (KeyFrame.cc:857-891:
 * every KeyLine whose midpoint lies within `radius` of the projected midpoint and whose 2-D direction has
 * |cos| >= th_cos, default 0.998, in index order), the octave gate pred-1..pred, and the smallest Hamming distance;
 * the earliest line wins ties.  kf_desc: the rows the reference compares against, which are those of
 * pKF->mDescriptors (the ORB descriptor matrix) at the LINE index (:938), not of mLineDescriptors; it must have at
 * least n_lines rows (the reference would throw beyond them).  best_idx[nq] = line index if that distance <= th_low
 * (TH_LOW = 50), else -1; best_dist[nq] (may be NULL) = the smallest distance, 256 when no line qualified.  The
 * replace-or-add bookkeeping on the MapLine objects (:955-978) stays with the caller; every query is independent. */
int psl_line_fuse(psl_ctx* ctx, const psl_keyline* kl, int32_t n_lines, const uint8_t* kf_desc, int32_t n_desc,
                  const psl_line_fuse_query* queries, const uint8_t* query_desc, int32_t nq, float th_cos, int32_t th_low,
                  int32_t* best_idx, int32_t* best_dist);

/* What LSDmatcher::SearchByProjection reads from the searched Frame (include/Frame.h): mvKeylinesUn, mLdesc,
 * mvKeyLineFunctions, mvLines3D (first/second endpoints, n x 6 doubles; only the map-line form needs it) and the
 * bounds of the 64x48 line grid (Frame::AssignFeaturesToGridForLine, Frame.cc:286-309; rebuilt by the library). */
typedef struct psl_line_frame_view {
  int32_t n;
  const psl_keyline* kl_un;
  const uint8_t* ldesc;
  const double* lineeq;
  const double* lines3d;
  float min_x, min_y, max_x, max_y;
  float grid_w_inv, grid_h_inv;
} psl_line_frame_view;

/* One projected MapLine: what Frame::isInFrustum(MapLine*, ...) leaves in mTrackProjX1.. (Frame.cc:828-898). */
typedef struct psl_line_query {
  float x1, y1, x2, y2;     /* mTrackProjX1, Y1, X2, Y2 */
  float radius;             /* th (LSDmatcher.cpp:152) or RadiusByViewingCos(viewCos)*th (:282-285) */
  float sx, sy, ex, ey;     /* mode 0: LastFrame.mvKeylinesUn[i] s/ePointInOctave (orientation gate :175-185) */
  float length;             /* mode 0: LastFrame.mvKeylinesUn[i].lineLength (:189-193) */
  double normal[3];         /* mode 1: pML->GetNormal() (:294, :312-322) */
  uint32_t flags;           /* PSL_Q_VALID | PSL_Q_CLAIMS */
  uint32_t pad_;
} psl_line_query;

/* LSDmatcher::SearchByProjection, mode 0 = (CurrentFrame, LastFrame, th) (LSDmatcher.cpp:112-215; window test with
 * TH 0.96, 2-D direction gate cos 10 deg, length ratio >= 0.75, best <= 95), mode 1 = (F, vpMapLines, eval_orient, th)
 * (:260-352; TH 0.998, 3-D direction gate cos 15 deg, best/second with same-level ratio test nn_ratio, best <= 95).
 * Candidates come from Frame::GetFeaturesInAreaForLine (Frame.cc:752-826) in its enumeration order.
 * claimed_in[n] (may be NULL) marks lines whose MapLine has Observations() > 0 before the call.
 * assign[n] = query whose MapLine ends up in mvpMapLines[i], or -1; *nmatches = return value. */
int psl_line_match_projection(psl_ctx* ctx, const psl_line_frame_view* frame, const psl_line_query* queries, const uint8_t* query_desc,
                              int32_t nq, const uint8_t* claimed_in, int32_t mode, float nn_ratio, int32_t* assign,
                              int32_t* nmatches);

/* InsectLineMatch::SearchMapInsectline (add_src/InsectlineMatch.cpp:9-60; mode 0) and its live twin
 * Map::AssociatePlanesByBoundary (src/Map.cc:204-272; mode 1: map planes flipped to d >= 0, the distance
 * threshold carried across structural lines, one match counted per assignment).
 * planes_cam[n_ljl*4]: Frame::mvPlanes; Tcw[16] row-major 4x4 float (Frame::ComputeWorldPlane, Frame.cc:918-925);
 * pts[n_ljl*15]: the four 3-D endpoints of the two lines and the junction (mvLines3D[..].first/second, CrossPoint_3D),
 * doubles; map_planes[n_map*4] float (InsectLine::GetWorldPos_plane); map_bad[n_map] (isBad, mode 0 only, may be
 * NULL).  assign[n_ljl] = map index in mvpMapInsecs[i] or -1. */
int psl_plane_assoc(psl_ctx* ctx, const float* planes_cam, const double* pts, int32_t n_ljl, const float* Tcw,
                    const float* map_planes, const uint8_t* map_bad, int32_t n_map, float d_th, float a_th,
                    int32_t mode, int32_t* assign, int32_t* nmatches);

/* ------------------------------------------------------------------------------------------------
 * Batched point front end of one tracking step (BASELINE config 2/4), everything resident in HBM.
 * For a batch of B consecutive RGB-D frames it does what the reference does per frame:
 *   Frame::Frame(gray, depth, ...)      ExtractORB (src/Frame.cc:179 -> ORBextractor::operator())
 *                                        ComputeStereoFromRGBD (src/Frame.cc:192, :1342-1363)
 *   Tracking::TrackWithMotionModel      ORBmatcher::SearchByProjection(Cur, Last, th, false)
 *                                        (src/Tracking.cc:1193 -> src/ORBmatcher.cc:1328-1470)
 * with frame b-1 of the batch playing LastFrame for frame b (every keypoint of the last frame that has depth
 * carries a MapPoint, as after Tracking::UpdateLastFrame), and Tcw[b] the pose prior of frame b.
 * Distortion must be zero (mvKeysUn == mvKeys, bounds = image; Frame.cc:155-158).
 * ------------------------------------------------------------------------------------------------ */
typedef struct psl_camera {
  float fx, fy, cx, cy; /* Camera.fx .. (Examples/RGB-D/ICL.yaml:8-11) */
  float bf;             /* Camera.bf */
  float depth_factor;   /* 1/DepthMapFactor as float (Tracking.cc:142-145) */
} psl_camera;

typedef struct psl_track_params {
  float th;                  /* window factor: 15 (Tracking.cc:1189) */
  float nn_ratio;            /* ORBmatcher(0.9, true) at Tracking.cc:1166 */
  int32_t check_orientation;
  int32_t th_dist;           /* TH_HIGH = 100 */
} psl_track_params;

/* DEVICE pointers, asynchronous on psl_stream(ctx).  gray: [B] frames of h rows, `gray_stride` bytes per row,
 * frames `gray_frame_stride` bytes apart; depth: u16, strides in PIXELS; Tcw: [B][12] row-major 3x4 float.
 * Outputs are [B][cap] blocks: kps, desc (x32), u_right, z (mvDepth), assign (index into frame b-1's
 * keypoints whose point landed on keypoint i of frame b, or -1), and n[B], nmatches[B] (nmatches[0] = 0). */
int psl_track_orb_batch_dev(psl_ctx* ctx, const uint8_t* d_gray, int32_t gray_stride, int64_t gray_frame_stride,
                            const uint16_t* d_depth, int32_t depth_stride_px, int64_t depth_frame_stride_px,
                            int32_t B, int32_t w, int32_t h, const float* d_Tcw, const psl_camera* cam,
                            const psl_track_params* prm, psl_keypoint* d_kps, uint8_t* d_desc, int32_t* d_n,
                            float* d_u_right, float* d_z, int32_t* d_assign, int32_t* d_nmatches, int32_t cap);

/* Same with HOST pointers (tightly packed frames: gray [B][h][w] u8, depth [B][h][w] u16, Tcw [B][12]);
 * H2D and D2H copies are part of the call. */
int psl_track_orb_batch(psl_ctx* ctx, const uint8_t* gray, const uint16_t* depth, int32_t B, int32_t w, int32_t h,
                        const float* Tcw, const psl_camera* cam, const psl_track_params* prm, psl_keypoint* kps,
                        uint8_t* desc, int32_t* n, float* u_right, float* z, int32_t* assign, int32_t* nmatches,
                        int32_t cap);

/* ------------------------------------------------------------------------------------------------
 * Batched combined front end of one tracking step (BASELINE config 4): per frame b of a batch of consecutive
 * RGB-D frames, everything the reference does between Tracking::GrabImageRGBD and Optimizer::PoseOptimization
 * on the feature side:
 *   Frame::Frame            ExtractORB + ExtractLSD (LINEextractor::operator()) + ComputeStereoFromRGBD
 *                           (src/Frame.cc:179-192)
 *   TrackWithMotionModel    LSDmatcher::SearchByGeomNApearance(Cur, Last, 0.95)   (src/Tracking.cc:1182-1183)
 *                           ORBmatcher::SearchByProjection(Cur, Last, th)         (src/Tracking.cc:1193)
 * with frame b-1 playing LastFrame (all of its lines / depth-valid points carrying map features).
 * Outputs are [B][cap] (points) and [B][line_cap] (lines) blocks; line_assign[b][i] = index of the line of
 * frame b-1 whose MapLine lands on line i of frame b, or -1; line_nmatches[0] = 0.
 * ------------------------------------------------------------------------------------------------ */
typedef struct psl_frontend_out {
  psl_keypoint* kps; uint8_t* desc; int32_t* n; float* u_right; float* z; int32_t* assign; int32_t* nmatches;
  int32_t cap;
  int32_t line_cap;
  psl_keyline* kl; uint8_t* ldesc; double* lineeq; int32_t* nl; int32_t* line_assign; int32_t* line_nmatches;
} psl_frontend_out;

/* DEVICE pointers (inputs as psl_track_orb_batch_dev, every psl_frontend_out pointer in HBM); asynchronous. */
int psl_track_frontend_batch_dev(psl_ctx* ctx, const uint8_t* d_gray, int32_t gray_stride, int64_t gray_frame_stride,
                                 const uint16_t* d_depth, int32_t depth_stride_px, int64_t depth_frame_stride_px,
                                 int32_t B, int32_t w, int32_t h, const float* d_Tcw, const psl_camera* cam,
                                 const psl_track_params* prm, float line_desc_th, const psl_frontend_out* out);

/* HOST pointers (tightly packed frames); H2D and D2H copies are part of the call. */
int psl_track_frontend_batch(psl_ctx* ctx, const uint8_t* gray, const uint16_t* depth, int32_t B, int32_t w, int32_t h,
                             const float* Tcw, const psl_camera* cam, const psl_track_params* prm, float line_desc_th,
                             const psl_frontend_out* out);

/* Tracking::GrabImageRGBD for a batch (src/Tracking.cc:214-243): cvtColor RGB / BGR (A) -> GRAY on the device (the
 * arithmetic of psl_convert_rgbd), then everything of psl_track_frontend_batch.  channels = 3 or 4; rgb_order != 0 for
 * RGB(A) input (mbRGB, Tracking.cc:222-232), 0 for BGR(A).
 *
 * _dev: DEVICE pointers, asynchronous.
 * Host form, one call: psl_track_rgbd_batch (tightly packed frames, width a multiple of 16).
 * Host form, pipelined: psl_track_rgbd_batch_begin only enqueues the uploads of a batch into one of two staging sets and
 * returns; psl_track_rgbd_batch_end runs the OLDEST begun batch and returns with its results in `out`.  Calling
 * begin(k+1) before end(k) keeps the PCIe link busy while batch k is computed (at most two batches in flight; the host
 * input buffers of a batch must stay valid and unchanged until its _end returns; pinned memory makes the copies
 * asynchronous).  PSL_E_INVALID when a third batch is begun or none is in flight. */
int psl_track_rgbd_batch_dev(psl_ctx* ctx, const uint8_t* d_color, int32_t channels, int32_t rgb_order, int32_t color_stride,
                             int64_t color_frame_stride, const uint16_t* d_depth, int32_t depth_stride_px,
                             int64_t depth_frame_stride_px, int32_t B, int32_t w, int32_t h, const float* d_Tcw,
                             const psl_camera* cam, const psl_track_params* prm, float line_desc_th,
                             const psl_frontend_out* out);
int psl_track_rgbd_batch(psl_ctx* ctx, const uint8_t* color, int32_t channels, int32_t rgb_order, const uint16_t* depth,
                         int32_t B, int32_t w, int32_t h, const float* Tcw, const psl_camera* cam,
                         const psl_track_params* prm, float line_desc_th, const psl_frontend_out* out);
int psl_track_rgbd_batch_begin(psl_ctx* ctx, const uint8_t* color, int32_t channels, int32_t rgb_order,
                               const uint16_t* depth, int32_t B, int32_t w, int32_t h, const float* Tcw);
int psl_track_rgbd_batch_end(psl_ctx* ctx, const psl_camera* cam, const psl_track_params* prm, float line_desc_th,
                             const psl_frontend_out* out);

/* The second half of Tracking::TrackWithMotionModel (src/Tracking.cc:1193-1240) on the arrays psl_track_*_batch_dev left in
 * HBM: for every frame b the keypoints that SearchByProjection matched (assign[b][i] = j >= 0) observe the MapPoint that
 * frame b-1 created at its keypoint j — Last.UnprojectStereo(j), Frame.cc:1367-1381, from that keypoint, its depth z and
 * the pose of frame b-1 — and Optimizer::PoseOptimization runs from the prior pose d_Tcw[b] (3x4 rows, as the batched
 * front end takes them).  d_Tcw_out [B][16] 4x4 row-major, d_outlier [B][cap] = mvbOutlier, d_n_inliers [B] = the return
 * value of PoseOptimization (0 and the prior pose for frame 0 and for frames with fewer than 3 matches).  Asynchronous. */
int psl_track_pose_batch_dev(psl_ctx* ctx, const psl_keypoint* d_kps, const float* d_u_right, const float* d_z,
                             const int32_t* d_assign, const int32_t* d_n, int32_t cap, int32_t B, const float* d_Tcw,
                             const psl_camera* cam, float* d_Tcw_out, uint8_t* d_outlier, int32_t* d_n_inliers);

/* Per-stage device timing (CUDA events on the ctx stream between the kernels of each stage).
 * Stages: 0 pyramid resize, 1 FAST cells, 2 octree selection, 3 Gaussian blur, 4 orientation+rBRIEF,
 * 5 single-pair matcher calls, 6 stereo + projection queries, 7 feature grid, 8 candidate lists,
 * 9 ordered resolve, 10 LSD prologue (blur, 0.8x resize, gradient, seed keys), 11 LSD seed ordering,
 * 12 LSD region growing, 13 line merge + KeyLines, 14 LBD, 15 line matching.  psl_profile_read synchronises, writes the accumulated
 * milliseconds and kernel-launch counts per stage since the last read (arrays of PSL_N_STAGES) and resets. */
#define PSL_N_STAGES 16
int psl_profile_enable(psl_ctx* ctx, int32_t on);
int psl_profile_read(psl_ctx* ctx, float* ms, int64_t* launches);
/* Kernel launches issued by this ctx since creation (all stages). */
int64_t psl_launch_count(const psl_ctx* ctx);

/* Diagnostic: copy an intermediate of the LAST ORB call (chunk-local frame index) to the host, so the
 * stage-level parity tests can compare with ComputePyramid (ORBextractor.cc:1107-1132), the FAST cell
 * loop (:789-829) and DistributeOctTree (:539-763) separately.
 *   what = 0: pyramid level image, tightly packed w*h bytes (level >= 1)
 *          1: blurred level image, tightly packed
 *          2: FAST candidates of the level in reference order, packed u32 (x-16)<<20 | (y-16)<<8 | score
 *          3: octree-selected keys of the level in list order, same packing
 *          4: (last LINE call) the 0.8x image LSD works on, tightly packed
 *          5: (last LINE call) raw LSD segments of the frame, float x1,y1,x2,y2 each
 * *n = number of bytes (0,1,4) or entries (2,3,5) written. */
int psl_debug_fetch(psl_ctx* ctx, int32_t what, int32_t frame, int32_t level, void* out, int64_t cap_bytes,
                    int64_t* n);

/* Frame::isLineGood (src/Frame.cc:662-750; SURVEY "next" row N2): the 3-D line of every KeyLine from the depth image.
 * Per line: min((int)length, 20) + 1 samples along it, nearest-pixel depth (> 0.01 m) back-projected with Frame's
 * fx, fy, cx, cy; at least 5 points; each point's covariance (LINEextractor::compPt3dCov, depth noise model of
 * LineExtractor.cpp:27-38) whitened through cv::SVD; a RANSAC of at most 10 draws under the Mahalanobis point-to-line
 * distance (< 3) with the 10-cell support test (extract3dline_mahdist, verify3dLine), the SVD refit loop
 * (computeLine3d_svd) and the two extreme inliers as end points; kept if they are more than 2 cm apart.
 * depth: CV_32F metres (the output of psl_convert_rgbd).  lines3d [n*6] = mvLines3D first / second (zeros when there is
 * no line), line_eq [n*3] = mvLineEq (unit direction as float; (-1,-1,-1) when there is no line).
 * rand() of random_unique (LineExtractor.h:25-37) is pinned to the ANSI C example generator, re-seeded per line with
 * seed * 1000003 + line index + 1 (the reference draws from the process-wide stream).  HOST pointers. */
int psl_lines_3d(psl_ctx* ctx, const psl_keyline* kl_un, int32_t n, const float* depth, int32_t w, int32_t h, float fx,
                 float fy, float cx, float cy, uint32_t seed, double* lines3d, float* line_eq);
/* Batched, DEVICE pointers, asynchronous: frame b owns rows [b*cap, b*cap + d_n[b]) of d_kl / d_lines3d / d_line_eq and
 * the depth image at d_depth + b * depth_frame_stride_px (row stride depth_stride_px floats). */
int psl_lines_3d_dev(psl_ctx* ctx, const psl_keyline* d_kl, const int32_t* d_n, int32_t cap, int32_t B, const float* d_depth,
                     int32_t w, int32_t h, int32_t depth_stride_px, int64_t depth_frame_stride_px, float fx, float fy,
                     float cx, float cy, uint32_t seed, double* d_lines3d, float* d_line_eq);

/* One entry of Frame::intersection_lines_plane (the junctions CPartiallyRecoverConnectivity / convertFansToKeyLines
 * found, Frame.cc:504-511): the two line indices, the 2-D junction and its 3-D position in the camera frame. */
typedef struct psl_line_junction {
  int32_t l1, l2;
  float cross2d_x, cross2d_y;
  double cross3d[3];
} psl_line_junction;

/* The plane hypotheses Frame::ExtractLSD builds from coplanar intersecting line pairs (src/Frame.cc:512-645 with
 * Frame::OldPlane :474-487; SURVEY "next" row N2, the producer of psl_plane_assoc's inputs).  Per junction, in order:
 * le_l[i*6..] = the two normalised 2-D line equations sp x ep / sqrt(l0^2 + l1^2) (mvle_l, pushed for every junction);
 * skipped if either line has mvLineEq == (0,0,0) or mvLines3D == 0 (Eigen isZero); normal = mvLineEq[l1] x mvLineEq[l2]
 * normalised in float; the signed distances of the four 3-D endpoints and the junction must span <= 0.05; plane =
 * (normal, -mean distance), flipped to d >= 0; dropped if OldPlane (|d - d'| <= 0.2 and |n.n'| >= 0.9397 against a plane
 * already kept in this call).  kl_un: mvKeylinesUn; line_eq [n_lines*3] float: mvLineEq; lines3d [n_lines*6] double:
 * mvLines3D first/second.  Outputs for the kept hypotheses, in order: planes [cap*4] (mvPlanes), normals [cap*3]
 * (mvPlaneNormal), junction_of [cap] = index of the junction (gives mvPlaneLineNo, CrossPoint_3D, CrossPoint_2D);
 * *n_planes = their number.  More than cap hypotheses: PSL_E_CAPACITY. */
int psl_plane_hypotheses(psl_ctx* ctx, const psl_keyline* kl_un, const float* line_eq, const double* lines3d,
                         int32_t n_lines, const psl_line_junction* junctions, int32_t n_junctions, double* le_l,
                         float* planes, double* normals, int32_t* junction_of, int32_t cap, int32_t* n_planes);

/* The junction detection of Frame::ExtractLSD (src/Frame.cc:504-507; SURVEY "next" row N2, the step between psl_lines_3d
 * and psl_plane_hypotheses): CPartiallyRecoverConnectivity(mLines, expandWidth, fans, im, fanThr)
 * (add_src/PartiallyRecoverConnectivity.cpp:14-133; radius = Frame::expandWidth = 20, fan_thr = Frame::fanThr = pi/4,
 * include/Frame.h:217-218) followed by Frame::convertFansToKeyLines / Frame_shortestDistance (src/Frame.cc:380-472).
 * kl_un: mvKeylinesUn (start / end points = Frame::keyLinesToMat); img_w / img_h: the image the lines come from.
 * fans [cap*4] = the rows (x, y, i, j) of `fans` after the duplicate removal, in the reference's order (this is also
 * LIL_gather: point + the two KeyLines i, j); *n_fans = their number (> cap: PSL_E_CAPACITY).
 * junctions [cap] (may be NULL; needs lines3d = mvLines3D, n*6 doubles) = Frame::intersection_lines_plane, the fans
 * whose two 3-D lines have a cross point (closest points of the two carrier lines, mid point; accepted if the mid-point
 * test of :417-421 holds and |cross| > DBL_EPSILON), ready for psl_plane_hypotheses; *n_junctions = their number.
 * Pinned: the cv::MatExpr of ptsDropInRotatedRect = one cv::addWeighted in double (cv2 4.13); the 2x2 system of
 * Frame_shortestDistance follows Eigen's ColPivHouseholderQR step by step; the function's missing `return` (undefined
 * behaviour when the mid-point test fails) reads "no cross point".  HOST pointers. */
int psl_line_junctions(psl_ctx* ctx, const psl_keyline* kl_un, const double* lines3d, int32_t n, int32_t img_w,
                       int32_t img_h, float radius, float fan_thr, float* fans, psl_line_junction* junctions,
                       int32_t cap, int32_t* n_fans, int32_t* n_junctions);
/* Batched, DEVICE pointers, asynchronous: frame b owns rows [b*line_cap, b*line_cap + d_n[b]) of d_kl / d_lines3d and
 * rows [b*cap, ...) of d_fans / d_junctions; d_n_fans[b] / d_n_junctions[b] = the counts (a count above cap means the
 * rows beyond cap were dropped).  d_lines3d and d_junctions are both NULL or both given. */
int psl_line_junctions_dev(psl_ctx* ctx, const psl_keyline* d_kl, const int32_t* d_n, int32_t line_cap, int32_t B,
                           const double* d_lines3d, int32_t img_w, int32_t img_h, float radius, float fan_thr,
                           float* d_fans, psl_line_junction* d_junctions, int32_t cap, int32_t* d_n_fans,
                           int32_t* d_n_junctions);

/* Optimizer::PoseOptimization (src/Optimizer.cc:239-1023; SURVEY "next" row N4) for a frame without InsectLine
 * observations: the pose-only Levenberg optimisation over the keypoints that hold a MapPoint — monocular / stereo
 * reprojection edges (EdgeSE3ProjectXYZOnlyPose / EdgeStereoSE3ProjectXYZOnlyPose of the vendored g2o) with Huber kernels
 * (delta^2 = 5.991 / 7.815), four rounds of at most ten Levenberg iterations from the pose prior, the chi-square
 * classification between rounds (outliers leave the system, :782-838), the kernels dropped after the third round.
 * One record per keypoint: flags & 1 = mvpMapPoints[i] != NULL; u_right < 0 = monocular observation.  Tcw: 4x4
 * row-major float (pFrame->mTcw in, the pose SetPose receives out); outlier[n] = mvbOutlier; *n_inliers = the return
 * value (nInitialCorrespondences - nBad; 0 and no change with fewer than 3 correspondences).  fp64 on the device; the
 * sums over the edges run in another order than g2o's, so the contract is a tolerance (pose to 1e-6, identical flags away
 * from the thresholds), and the oracle is a restatement that g2o cannot pin here.  This form carries no
 * structural-line (LIL) edges, i.e. it is PoseOptimization for a frame with N_LJL = 0; psl_pose_optimization_lil below
 * is the complete call. */
typedef struct psl_pose_point {
  float u, v, u_right;   /* mvKeysUn[i].pt, mvuRight[i] */
  float inv_sigma2;      /* mvInvLevelSigma2[mvKeysUn[i].octave] */
  float xw, yw, zw;      /* pMP->GetWorldPos() */
  uint32_t flags;
} psl_pose_point;
int psl_pose_optimization(psl_ctx* ctx, const float* Tcw_in, const psl_pose_point* pts, int32_t n, float fx, float fy,
                          float cx, float cy, float bf, float* Tcw_out, uint8_t* outlier, int32_t* n_inliers);
/* Batched, DEVICE pointers, asynchronous: frame b owns rows [b*cap, b*cap + d_n[b]) of d_pts / d_outlier and the 16
 * floats at d_Tcw_in / d_Tcw_out + 16*b; one CTA per frame. */
int psl_pose_optimization_dev(psl_ctx* ctx, const float* d_Tcw_in, const psl_pose_point* d_pts, const int32_t* d_n,
                              int32_t cap, int32_t B, float fx, float fy, float cx, float cy, float bf, float* d_Tcw_out,
                              uint8_t* d_outlier, int32_t* d_n_inliers);

/* The complete Optimizer::PoseOptimization: the point edges above plus one EdgeLILSE3ProjectXYZ per structural line of the
 * frame that holds a map InsectLine (src/Optimizer.cc:619-693; edge: add_inc/EdgeLIL.h:210-374).  The map InsectLine is a
 * FIXED vertex (two 3-D segments and their cross point, :637-649), so the edge only adds to the pose block: a 6-dim
 * error — the distances of the four projected end points to the two observed 2-D line equations and the reprojection
 * error of the cross point — with identity information and a Huber kernel of delta^2 = 11.07.  The reference's
 * Jacobian reads both end points of the second segment from the same three numbers (EdgeLIL.h:276-279); reproduced.
 * Between rounds the LIL edges are classified like the points (chi2 > 11.07 -> mvbOutlier_Insec, level 1, :976-1007);
 * lil_outlier[n_lil] receives mvbOutlier_Insec.  *n_inliers = nInitialCorrespondences - nBad, where the initial count
 * includes the LIL edges and nBad only the point outliers (nLineBad is never incremented, :712, :1021). */
typedef struct psl_pose_lil {
  double line1[6]; /* pLIL->line1: start, end in world coordinates */
  double line2[6]; /* pLIL->line2 */
  double cross[3]; /* pLIL->crosspoint */
  double obs1[3];  /* pFrame->mvle_l[i].first: the observed 2-D line equation of the first segment */
  double obs2[3];  /* pFrame->mvle_l[i].second */
  double ins[2];   /* pFrame->CrossPoint_2D[i] */
  uint32_t flags;  /* & 1: mvpMapInsecs[i] != NULL and not bad */
  uint32_t pad_;
} psl_pose_lil;    /* 192 B */
int psl_pose_optimization_lil(psl_ctx* ctx, const float* Tcw_in, const psl_pose_point* pts, int32_t n,
                              const psl_pose_lil* lils, int32_t n_lil, float fx, float fy, float cx, float cy, float bf,
                              float* Tcw_out, uint8_t* outlier, uint8_t* lil_outlier, int32_t* n_inliers);
/* Batched, DEVICE pointers, asynchronous: frame b additionally owns rows [b*lil_cap, b*lil_cap + d_n_lil[b]) of d_lils /
 * d_lil_outlier.  d_lils may be NULL with lil_cap = 0 (no structural lines). */
int psl_pose_optimization_lil_dev(psl_ctx* ctx, const float* d_Tcw_in, const psl_pose_point* d_pts, const int32_t* d_n,
                                  int32_t cap, const psl_pose_lil* d_lils, const int32_t* d_n_lil, int32_t lil_cap, int32_t B,
                                  float fx, float fy, float cx, float cy, float bf, float* d_Tcw_out, uint8_t* d_outlier,
                                  uint8_t* d_lil_outlier, int32_t* d_n_inliers);

#ifdef __cplusplus
}
#endif
#endif /* PSL_FRONTEND_H */
