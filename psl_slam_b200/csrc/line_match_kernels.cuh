// Launchers of the line matchers (all asynchronous on `st`), batch-first: B independent pairs, [B][cap] blocks.
#pragma once
#include "psl_common.cuh"

namespace psl {

constexpr int kLineCells = 80;     // grid cells one line can cross (Bresenham over a 64x48 grid: <= 66 in-grid steps)
constexpr int kMaxLinesPerFrame = 4096;  // line index field of the projection keys

// the lines of a batch of frames: frame b owns rows [b*cap, b*cap + n[b])
struct LineSet {
  const psl_keyline* kl;  // may be null for descriptor-only matchers
  const uint8_t* desc;
  const int32_t* n;       // device array [B]
  int32_t cap;
};

void launch_line_knn2(const LineSet& Q, const LineSet& T, uint2* knn, int B, cudaStream_t st);
void launch_line_nnr(const LineSet& Q, const uint2* knn, float nnr, int32_t* m12, int32_t* nmatches, int B,
                     cudaStream_t st);
void launch_line_geom(const LineSet& Last, const uint8_t* has_ml, const LineSet& Cur, const uint2* knn, float desc_th,
                      float bounds_w, float bounds_h, int32_t* assign_cur, int32_t* nmatches, int B, cudaStream_t st);
void launch_line_bfmatch(const LineSet& Q, const LineSet& T, const uint2* knn, float nn_ratio, float th,
                         int32_t* matches, int B, cudaStream_t st);
void launch_line_mutual(const LineSet& A, const int32_t* m21, int cap2, int32_t* m12, int32_t* nmatches, int B,
                        cudaStream_t st);
// FrameBFMatchNew: knn = launch_line_knn2(Q, T); Q.kl / T.kl and funcT (line equations of T, [B][cap][3]) are read
void launch_line_bfmatch_new(const LineSet& Q, const LineSet& T, const double* funcT, const uint2* knn, const float* F,
                             float nn_ratio, float th, int32_t* matches, int B, cudaStream_t st);
void launch_line_triang(const LineSet& A, const int32_t* m21, int cap2, const uint8_t* ml1, const uint8_t* ml2,
                        int is_double, int32_t* m12, int32_t* nmatches, int B, cudaStream_t st);
// window search of LSDmatcher::Fuse: one warp per query over the n_lines KeyLines of one KeyFrame
void launch_line_fuse(const psl_keyline* kl, int n_lines, const uint8_t* kf_desc, const psl_line_fuse_query* queries,
                      const uint8_t* qdesc, int nq, float th_cos, int th_low, int32_t* best_idx, int32_t* best_dist,
                      cudaStream_t st);
// 3 launches: line grid cells, static keys, ordered resolve
void launch_line_projection(const LineSet& F, const double* lineeq, const double* lines3d, const psl_line_query* queries,
                            const uint8_t* qdesc, const int32_t* nq, int qcap, int max_nq, float min_x, float min_y,
                            float w_inv, float h_inv, int mode, float nn_ratio, const uint8_t* claimed_in,
                            uint16_t* cells, uint8_t* ncell, unsigned long long* keys, uint8_t* claimed, int32_t* assign,
                            int32_t* nmatches, int B, cudaStream_t st);
// Frame::isLineGood (Frame.cc:662-750): mvLines3D [B][cap][6] f64 and mvLineEq [B][cap][3] f32 of frame b's n_lines[b] KeyLines
void launch_lines3d(const psl_keyline* kl, const int32_t* n_lines, int cap, const float* depth, int w, int h, int stride,
                    int64_t frame_stride, float fx, float fy, float cx, float cy, uint32_t seed, double* lines3d,
                    float* line_eq, int B, cudaStream_t st);
// Frame::ExtractLSD plane hypotheses (Frame.cc:512-645): one warp; n_planes counts every kept hypothesis (> cap = overflow)
void launch_plane_hypotheses(const psl_keyline* kl_un, const float* line_eq, const double* lines3d,
                             const psl_line_junction* js, int nj, double* le_l, float* planes, double* normals,
                             int32_t* junction_of, int cap, int32_t* n_planes, float* kept_scratch, cudaStream_t st);
void launch_plane_assoc(const float* planes_cam, const double* pts, int n_ljl, const float* Tcw, const float* map_planes,
                        const uint8_t* map_bad, int n_map, float d_th, float a_th, int mode, int32_t* assign,
                        int32_t* nmatches, cudaStream_t st);

}  // namespace psl
