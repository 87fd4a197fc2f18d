// Launchers of the Hamming matchers (all asynchronous on `st`).  Batch-first: every launcher
// works on B independent (frame, query-set) pairs laid out as [B][cap] row blocks.
#pragma once
#include "psl_common.cuh"

namespace psl {

constexpr int kGridCells = PSL_GRID_COLS * PSL_GRID_ROWS;  // 3072
constexpr int kCandCap = 256;                              // candidates kept per projected query

// The searched frames of a batch: frame b owns rows [b*cap, b*cap + n[b]).
struct MatchFrames {
  const psl_keypoint* kps;  // undistorted keypoints (mvKeysUn)
  const float* u_right;     // may be null
  const uint8_t* desc;
  const int32_t* n;         // device array [B]
  int32_t cap;
  float min_x, min_y, grid_w_inv, grid_h_inv;
};

struct MatchQueries {
  const psl_proj_query* q;
  const uint8_t* desc;
  const int32_t* nq;  // device array [B]
  int32_t cap;
};

// per-frame CSR of the 64x48 feature grid: start[B][3073], items[B][cap]
void launch_grid_build(const MatchFrames& f, int32_t* cell_start, uint16_t* cell_items, int B, cudaStream_t st);

// ordered candidate lists with distances: cand[B][qcap][kCandCap] = idx<<16 | dist, count[B][qcap];
// best[B][qcap] = the two lexicographically smallest (dist<<16 | position) of each list (0xFFFFFFFF = none)
void launch_proj_candidates(const MatchFrames& f, const MatchQueries& q, const int32_t* cell_start,
                            const uint16_t* cell_items, uint32_t* cand, int32_t* cand_count, uint2* cand_best,
                            uint32_t* status, int B, cudaStream_t st);

// ordered greedy resolve + rotation-histogram filter; claimed_in may be null
void launch_proj_resolve(const MatchFrames& f, const MatchQueries& q, const uint32_t* cand, const int32_t* cand_count,
                         const uint2* cand_best, const uint8_t* claimed_in, psl_match_params prm,
                         uint32_t* accepted_scratch, int32_t* assign, int32_t* nmatches, int B, cudaStream_t st);

void launch_descriptor_distance(const uint8_t* a, const uint8_t* b, int n, int32_t* dist, cudaStream_t st);
void launch_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx, int32_t* dist, cudaStream_t st);

// SearchByBoW: `pairs` = (kf node slot, frame node slot) of equal node ids.  The KeyFrame-KeyFrame form
// (ORBmatcher.cc:522-655) passes f_valid (MapPoint test on the searched side), strict = 1 (`< TH_LOW`) and m12[nkf]
// (the result indexed by the first keyframe); the KeyFrame-Frame form passes nullptr, 0, nullptr.
void launch_bow(const uint8_t* kf_desc, const float* kf_angle, const uint8_t* kf_valid, const int32_t* kf_offs,
                const uint32_t* kf_idx, const uint8_t* f_desc, const float* f_angle, const uint8_t* f_valid,
                const int32_t* f_offs, const uint32_t* f_idx, const int2* pairs, int npairs, float nn_ratio, int th_low,
                int strict, int check_ori, int nf, int32_t* match_f, int32_t* hist /*[32]*/, uint32_t* accepted,
                int32_t* n_accepted, int32_t* nmatches, int32_t* m12, int nkf, cudaStream_t st);

// SearchBySim3: agreement of the two directions (m1[n1] -> KF2 index, m2[n2] -> KF1 index)
void launch_sim3_agree(const int32_t* m1, int n1, const int32_t* m2, int32_t* out, int32_t* nfound, cudaStream_t st);

// SearchForInitialization: ordered greedy loop over the candidate lists of launch_proj_candidates (B = 1).
// mdist / m21: int32[n2] scratch; accepted: uint32[n1] scratch; prev_matched [n1][2] updated in place.
void launch_init_resolve(const uint32_t* cand, const int32_t* cand_count, const psl_keypoint* kps1, int n1,
                         const psl_keypoint* kps2, int n2, float nn_ratio, int th_low, int check_ori, int32_t* mdist,
                         int32_t* m21, uint32_t* accepted, int32_t* m12, float* prev_matched, int32_t* nmatches,
                         cudaStream_t st);

// the window search of ORBmatcher::Fuse for one keyframe (f: B = 1 views; grid from launch_grid_build)
void launch_fuse(const MatchFrames& f, const psl_fuse_query* qs, const uint8_t* qdesc, int nq, const int32_t* cell_start,
                 const uint16_t* cell_items, const float* inv_sigma2 /* nullptr: no chi-square gate */, int th_low,
                 int32_t* best_idx, int32_t* best_dist, cudaStream_t st);

// SearchForTriangulation: candidate scan per vocabulary-node pair + rotation filter.  hist: int32[33] (30 bins used, [32] = nmatches)
void launch_triangulation(const psl_keypoint* kps1, const float* ur1, const uint8_t* desc1, const uint8_t* mp1,
                          const int32_t* offs1, const uint32_t* idx1, int n1, const psl_keypoint* kps2, const float* ur2,
                          const uint8_t* desc2, const uint8_t* mp2, const int32_t* offs2, const uint32_t* idx2,
                          const int2* pairs, int npairs, const float* F12, float ex, float ey, const float* scale2,
                          const float* sigma2, int only_stereo, int th_low, int check_ori, int32_t* m12, int32_t* hist,
                          int32_t* nmatches, cudaStream_t st);

}  // namespace psl
