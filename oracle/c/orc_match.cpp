// TEST INFRASTRUCTURE — CPU oracle for the point matchers (see psl_oracle.h).
// Sequential restatement of /root/reference/src/ORBmatcher.cc and the Frame helpers it calls
// (src/Frame.cc:269-284, 985-1050) on plain arrays.  Nothing here is used by the product.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "psl_oracle.h"

namespace {

const int HISTO_LENGTH = 30;  // ORBmatcher.cc:38

// ORBmatcher::DescriptorDistance, ORBmatcher.cc:1647-1663 (SWAR popcount over 8 words)
int desc_dist(const uint8_t* a, const uint8_t* b) {
  int dist = 0;
  for (int i = 0; i < 8; ++i) {
    uint32_t x, y;
    std::memcpy(&x, a + 4 * i, 4);
    std::memcpy(&y, b + 4 * i, 4);
    uint32_t v = x ^ y;
    v = v - ((v >> 1) & 0x55555555u);
    v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
    dist += (int)((((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24);
  }
  return dist;
}

// Frame::AssignFeaturesToGrid + PosInGrid, Frame.cc:269-284,1040-1050
struct Grid {
  std::vector<int> cell[PSL_GRID_COLS][PSL_GRID_ROWS];
  explicit Grid(const psl_frame_view& f) {
    for (int i = 0; i < f.n; ++i) {
      const int px = (int)roundf((f.kps_un[i].x - f.min_x) * f.grid_w_inv);
      const int py = (int)roundf((f.kps_un[i].y - f.min_y) * f.grid_h_inv);
      if (px < 0 || px >= PSL_GRID_COLS || py < 0 || py >= PSL_GRID_ROWS) continue;
      cell[px][py].push_back(i);
    }
  }
};

// Frame::GetFeaturesInArea, Frame.cc:985-1038
void features_in_area(const psl_frame_view& f, const Grid& g, float x, float y, float r, int minLevel, int maxLevel,
                      std::vector<int>& out) {
  out.clear();
  const int nMinCellX = std::max(0, (int)floorf((x - f.min_x - r) * f.grid_w_inv));
  if (nMinCellX >= PSL_GRID_COLS) return;
  const int nMaxCellX = std::min(PSL_GRID_COLS - 1, (int)ceilf((x - f.min_x + r) * f.grid_w_inv));
  if (nMaxCellX < 0) return;
  const int nMinCellY = std::max(0, (int)floorf((y - f.min_y - r) * f.grid_h_inv));
  if (nMinCellY >= PSL_GRID_ROWS) return;
  const int nMaxCellY = std::min(PSL_GRID_ROWS - 1, (int)ceilf((y - f.min_y + r) * f.grid_h_inv));
  if (nMaxCellY < 0) return;
  const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
  for (int ix = nMinCellX; ix <= nMaxCellX; ++ix)
    for (int iy = nMinCellY; iy <= nMaxCellY; ++iy)
      for (int i : g.cell[ix][iy]) {
        const psl_keypoint& kp = f.kps_un[i];
        if (bCheckLevels) {
          if (kp.octave < minLevel) continue;
          if (maxLevel >= 0 && kp.octave > maxLevel) continue;
        }
        const float dx = kp.x - x, dy = kp.y - y;
        if (fabsf(dx) < r && fabsf(dy) < r) out.push_back(i);
      }
}

// ORBmatcher::ComputeThreeMaxima, ORBmatcher.cc:1601-1642
void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  for (int i = 0; i < L; ++i) {
    const int s = (int)histo[i].size();
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
    else if (s > max3) { max3 = s; ind3 = i; }
  }
  if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
  else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

int rot_bin(float a1, float a2) {
  const float factor = 1.0f / HISTO_LENGTH;  // the reference's quirk: only bins 0..12 are ever hit
  float rot = a1 - a2;
  if (rot < 0.0) rot += 360.0f;
  int bin = (int)roundf(rot * factor);
  if (bin == HISTO_LENGTH) bin = 0;
  return bin;
}

}  // namespace

extern "C" {

void orc_descriptor_distance(const uint8_t* a, const uint8_t* b, int n, int32_t* dist) {
  for (int i = 0; i < n; ++i) dist[i] = desc_dist(a + 32 * (size_t)i, b + 32 * (size_t)i);
}

// cv::BFMatcher(NORM_HAMMING).knnMatch(k=2): ascending distance, earliest index on ties
void orc_hamming_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx, int32_t* dist) {
  for (int i = 0; i < nq; ++i) {
    int b0 = -1, b1 = -1, d0 = 1 << 30, d1 = 1 << 30;
    for (int j = 0; j < nt; ++j) {
      const int d = desc_dist(q + 32 * (size_t)i, t + 32 * (size_t)j);
      if (d < d0) { d1 = d0; b1 = b0; d0 = d; b0 = j; }
      else if (d < d1) { d1 = d; b1 = j; }
    }
    idx[2 * i] = b0; idx[2 * i + 1] = b1;
    dist[2 * i] = b0 < 0 ? -1 : d0;
    dist[2 * i + 1] = b1 < 0 ? -1 : d1;
  }
}

// Candidate list of one query in the reference's enumeration order (for stage tests)
int orc_features_in_area(const psl_frame_view* f, float x, float y, float r, int minLevel, int maxLevel, int32_t* out,
                         int cap) {
  Grid g(*f);
  std::vector<int> v;
  features_in_area(*f, g, x, y, r, minLevel, maxLevel, v);
  for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = v[i];
  return (int)v.size();
}

// ORBmatcher::SearchByProjection — mode 0: (Frame&, const Frame& LastFrame, th, bMono) ORBmatcher.cc:1328-1470
//                                  mode 1: (Frame&, vector<MapPoint*>&, th)            ORBmatcher.cc:45-129
int orc_match_projection(const psl_frame_view* f, const psl_proj_query* qs, const uint8_t* qdesc, int nq,
                         const uint8_t* claimed_in, const psl_match_params* p, int32_t* assign, int32_t* nmatches) {
  Grid g(*f);
  std::vector<uint8_t> claimed(f->n, 0);
  if (claimed_in) std::memcpy(claimed.data(), claimed_in, f->n);
  for (int i = 0; i < f->n; ++i) assign[i] = -1;
  std::vector<int> rotHist[HISTO_LENGTH];
  std::vector<int> cand;
  int nm = 0;
  for (int q = 0; q < nq; ++q) {
    const psl_proj_query& Q = qs[q];
    if (!(Q.flags & PSL_Q_VALID)) continue;
    features_in_area(*f, g, Q.u, Q.v, Q.radius, Q.min_level, Q.max_level, cand);
    if (cand.empty()) continue;
    const uint8_t* dq = qdesc + 32 * (size_t)q;
    int bestDist = 256, bestIdx = -1, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1;
    for (int i2 : cand) {
      if (claimed[i2]) continue;  // mvpMapPoints[i2] && Observations()>0
      if (f->u_right && f->u_right[i2] > 0) {
        const float er = fabsf(Q.u_right - f->u_right[i2]);
        if (er > Q.radius) continue;
      }
      const int dist = desc_dist(dq, f->desc + 32 * (size_t)i2);
      if (dist < bestDist) {
        bestDist2 = bestDist; bestLevel2 = bestLevel;
        bestDist = dist; bestLevel = f->kps_un[i2].octave; bestIdx = i2;
      } else if (p->mode == 1 && dist < bestDist2) {
        bestLevel2 = f->kps_un[i2].octave; bestDist2 = dist;
      }
    }
    if (bestDist <= p->th_dist) {
      if (p->mode == 1 && bestLevel == bestLevel2 && bestDist > p->nn_ratio * bestDist2) continue;
      assign[bestIdx] = q;
      claimed[bestIdx] = (Q.flags & PSL_Q_CLAIMS) ? 1 : 0;
      ++nm;
      if (p->mode == 0 && p->check_orientation) rotHist[rot_bin(Q.angle, f->kps_un[bestIdx].angle)].push_back(bestIdx);
    }
  }
  if (p->mode == 0 && p->check_orientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; ++i)
      if (i != ind1 && i != ind2 && i != ind3)
        for (int idx : rotHist[i]) { assign[idx] = -1; --nm; }
  }
  *nmatches = nm;
  return 0;
}

// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&), ORBmatcher.cc:159-288
int orc_match_bow(const uint8_t* kf_desc, const float* kf_angle, const uint8_t* kf_valid, int nkf,
                  const psl_feature_vector* kfv, const uint8_t* f_desc, const float* f_angle, int nf,
                  const psl_feature_vector* ffv, float nn_ratio, int th_low, int check_orientation, int32_t* match_f,
                  int32_t* nmatches) {
  (void)nkf;
  for (int i = 0; i < nf; ++i) match_f[i] = -1;
  std::vector<int> rotHist[HISTO_LENGTH];
  int nm = 0, a = 0, b = 0;
  while (a < kfv->n_nodes && b < ffv->n_nodes) {
    if (kfv->node_id[a] == ffv->node_id[b]) {
      for (int ik = kfv->offs[a]; ik < kfv->offs[a + 1]; ++ik) {
        const int realIdxKF = (int)kfv->idx[ik];
        if (!kf_valid[realIdxKF]) continue;
        const uint8_t* dKF = kf_desc + 32 * (size_t)realIdxKF;
        int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
        for (int jf = ffv->offs[b]; jf < ffv->offs[b + 1]; ++jf) {
          const int realIdxF = (int)ffv->idx[jf];
          if (match_f[realIdxF] >= 0) continue;
          const int dist = desc_dist(dKF, f_desc + 32 * (size_t)realIdxF);
          if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF; }
          else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist1 <= th_low && (float)bestDist1 < nn_ratio * (float)bestDist2) {
          match_f[bestIdxF] = realIdxKF;
          if (check_orientation) rotHist[rot_bin(kf_angle[realIdxKF], f_angle[bestIdxF])].push_back(bestIdxF);
          ++nm;
        }
      }
      ++a; ++b;
    } else if (kfv->node_id[a] < ffv->node_id[b]) {
      while (a < kfv->n_nodes && kfv->node_id[a] < ffv->node_id[b]) ++a;  // lower_bound
    } else {
      while (b < ffv->n_nodes && ffv->node_id[b] < kfv->node_id[a]) ++b;
    }
  }
  if (check_orientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; ++i) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx : rotHist[i]) { match_f[idx] = -1; --nm; }
    }
  }
  *nmatches = nm;
  return 0;
}


// ORBmatcher::SearchForTriangulation, ORBmatcher.cc:657-823 (+ CheckDistEpipolarLine :140-157).  vbMatched2 is
// declared but never set by the reference, so KF2 keypoints can be matched more than once; kept.
int orc_match_triangulation(const psl_keyframe_view* kf1, const psl_feature_vector* fv1, const psl_keyframe_view* kf2,
                            const psl_feature_vector* fv2, const float* F12, float ex, float ey,
                            const float* scale_factors2, const float* level_sigma2_2, int only_stereo,
                            int check_orientation, int th_low, int32_t* matches12, int32_t* nmatches) {
  for (int i = 0; i < kf1->n; ++i) matches12[i] = -1;
  std::vector<int> rotHist[HISTO_LENGTH];
  int nm = 0, a = 0, b = 0;
  while (a < fv1->n_nodes && b < fv2->n_nodes) {
    if (fv1->node_id[a] == fv2->node_id[b]) {
      for (int i1 = fv1->offs[a]; i1 < fv1->offs[a + 1]; ++i1) {
        const int idx1 = (int)fv1->idx[i1];
        if (kf1->has_mappoint[idx1]) continue;
        const bool bStereo1 = kf1->u_right[idx1] >= 0;
        if (only_stereo && !bStereo1) continue;
        const psl_keypoint& kp1 = kf1->kps_un[idx1];
        const uint8_t* d1 = kf1->desc + 32 * (size_t)idx1;
        int bestDist = th_low, bestIdx2 = -1;
        for (int i2 = fv2->offs[b]; i2 < fv2->offs[b + 1]; ++i2) {
          const int idx2 = (int)fv2->idx[i2];
          if (kf2->has_mappoint[idx2]) continue;
          const bool bStereo2 = kf2->u_right[idx2] >= 0;
          if (only_stereo && !bStereo2) continue;
          const int dist = desc_dist(d1, kf2->desc + 32 * (size_t)idx2);
          if (dist > th_low || dist > bestDist) continue;
          const psl_keypoint& kp2 = kf2->kps_un[idx2];
          if (!bStereo1 && !bStereo2) {
            const float distex = ex - kp2.x, distey = ey - kp2.y;
            if (distex * distex + distey * distey < 100 * scale_factors2[kp2.octave]) continue;
          }
          // CheckDistEpipolarLine
          const float ea = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
          const float eb = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
          const float ec = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
          const float num = ea * kp2.x + eb * kp2.y + ec;
          const float den = ea * ea + eb * eb;
          if (den == 0) continue;
          const float dsqr = num * num / den;
          if (dsqr < 3.84 * level_sigma2_2[kp2.octave]) {
            bestIdx2 = idx2;
            bestDist = dist;
          }
        }
        if (bestIdx2 >= 0) {
          matches12[idx1] = bestIdx2;
          ++nm;
          if (check_orientation) rotHist[rot_bin(kp1.angle, kf2->kps_un[bestIdx2].angle)].push_back(idx1);
        }
      }
      ++a; ++b;
    } else if (fv1->node_id[a] < fv2->node_id[b]) {
      while (a < fv1->n_nodes && fv1->node_id[a] < fv2->node_id[b]) ++a;
    } else {
      while (b < fv2->n_nodes && fv2->node_id[b] < fv1->node_id[a]) ++b;
    }
  }
  if (check_orientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; ++i) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx : rotHist[i]) { matches12[idx] = -1; --nm; }
    }
  }
  *nmatches = nm;
  return 0;
}


// the window search of ORBmatcher::Fuse, ORBmatcher.cc:893-950 (KeyFrame::GetFeaturesInArea = the Frame version
// without a level gate, KeyFrame.cc:685-724)
int orc_match_fuse(const psl_frame_view* kf, const psl_fuse_query* qs, const uint8_t* qdesc, int nq,
                   const float* inv_level_sigma2, int th_low, int32_t* best_idx, int32_t* best_dist) {
  Grid g(*kf);
  std::vector<int> cand;
  for (int q = 0; q < nq; ++q) {
    best_idx[q] = -1;
    if (best_dist) best_dist[q] = 256;
    const psl_fuse_query& Q = qs[q];
    if (!(Q.flags & PSL_Q_VALID)) continue;
    features_in_area(*kf, g, Q.u, Q.v, Q.radius, -1, -1, cand);
    int bestDist = 256, bestIdx = -1;
    for (int idx : cand) {
      const psl_keypoint& kp = kf->kps_un[idx];
      const int kpLevel = kp.octave;
      if (kpLevel < Q.pred_level - 1 || kpLevel > Q.pred_level) continue;
      const float ex = Q.u - kp.x, ey = Q.v - kp.y;
      if (!inv_level_sigma2) {
        // the Sim3 form (ORBmatcher.cc:1046-1075) has no reprojection gate
      } else if (kf->u_right && kf->u_right[idx] >= 0) {
        const float er = Q.u_right - kf->u_right[idx];
        const float e2 = ex * ex + ey * ey + er * er;
        if (e2 * inv_level_sigma2[kpLevel] > 7.8) continue;
      } else {
        const float e2 = ex * ex + ey * ey;
        if (e2 * inv_level_sigma2[kpLevel] > 5.99) continue;
      }
      const int dist = desc_dist(qdesc + 32 * (size_t)q, kf->desc + 32 * (size_t)idx);
      if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
    }
    if (best_dist) best_dist[q] = bestDist;
    if (bestDist <= th_low) best_idx[q] = bestIdx;
  }
  return 0;
}

// ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&), ORBmatcher.cc:522-655
int orc_match_bow_kf(const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int n1,
                     const psl_feature_vector* fv1, const uint8_t* desc2, const float* angle2, const uint8_t* valid2,
                     int n2, const psl_feature_vector* fv2, float nn_ratio, int th_low, int check_orientation,
                     int32_t* matches12, int32_t* nmatches) {
  for (int i = 0; i < n1; ++i) matches12[i] = -1;
  std::vector<char> vbMatched2((size_t)n2, 0);
  std::vector<int> rotHist[HISTO_LENGTH];
  int nm = 0, a = 0, b = 0;
  while (a < fv1->n_nodes && b < fv2->n_nodes) {
    if (fv1->node_id[a] == fv2->node_id[b]) {
      for (int i1 = fv1->offs[a]; i1 < fv1->offs[a + 1]; ++i1) {
        const int idx1 = (int)fv1->idx[i1];
        if (!valid1[idx1]) continue;  // :561-566
        const uint8_t* d1 = desc1 + 32 * (size_t)idx1;
        int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
        for (int i2 = fv2->offs[b]; i2 < fv2->offs[b + 1]; ++i2) {
          const int idx2 = (int)fv2->idx[i2];
          if (vbMatched2[idx2] || !valid2[idx2]) continue;  // :580-586
          const int dist = desc_dist(d1, desc2 + 32 * (size_t)idx2);
          if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
          else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist1 < th_low) {  // :592, strict
          if ((float)bestDist1 < nn_ratio * (float)bestDist2) {
            matches12[idx1] = bestIdx2;
            vbMatched2[bestIdx2] = 1;
            if (check_orientation) rotHist[rot_bin(angle1[idx1], angle2[bestIdx2])].push_back(idx1);
            ++nm;
          }
        }
      }
      ++a; ++b;
    } else if (fv1->node_id[a] < fv2->node_id[b]) {
      while (a < fv1->n_nodes && fv1->node_id[a] < fv2->node_id[b]) ++a;
    } else {
      while (b < fv2->n_nodes && fv2->node_id[b] < fv1->node_id[a]) ++b;
    }
  }
  if (check_orientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; ++i) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx : rotHist[i]) { matches12[idx] = -1; --nm; }
    }
  }
  *nmatches = nm;
  return 0;
}

namespace {
// one direction of SearchBySim3 after the projection, ORBmatcher.cc:1188-1216 / :1264-1296
void sim3_direction(const psl_frame_view* kf, const psl_fuse_query* qs, const uint8_t* qdesc, int nq, int th_high,
                    std::vector<int>& vnMatch) {
  Grid g(*kf);
  std::vector<int> cand;
  vnMatch.assign((size_t)nq, -1);
  for (int q = 0; q < nq; ++q) {
    const psl_fuse_query& Q = qs[q];
    if (!(Q.flags & PSL_Q_VALID)) continue;
    features_in_area(*kf, g, Q.u, Q.v, Q.radius, -1, -1, cand);
    int bestDist = INT32_MAX, bestIdx = -1;
    for (int idx : cand) {
      const psl_keypoint& kp = kf->kps_un[idx];
      if (kp.octave < Q.pred_level - 1 || kp.octave > Q.pred_level) continue;
      const int dist = desc_dist(qdesc + 32 * (size_t)q, kf->desc + 32 * (size_t)idx);
      if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
    }
    if (bestDist <= th_high) vnMatch[q] = bestIdx;
  }
}
}  // namespace

// ORBmatcher::SearchBySim3, ORBmatcher.cc:1102-1326 (the MapPoint tests and the projections are the caller's)
int orc_match_sim3(const psl_frame_view* kf1, const psl_frame_view* kf2, const psl_fuse_query* q12,
                   const uint8_t* mp_desc1, const psl_fuse_query* q21, const uint8_t* mp_desc2, int th_high,
                   int32_t* matches12, int32_t* nfound) {
  std::vector<int> vnMatch1, vnMatch2;
  sim3_direction(kf2, q12, mp_desc1, kf1->n, th_high, vnMatch1);
  sim3_direction(kf1, q21, mp_desc2, kf2->n, th_high, vnMatch2);
  int nFound = 0;
  for (int i1 = 0; i1 < kf1->n; ++i1) {  // :1306-1321
    matches12[i1] = -1;
    const int idx2 = vnMatch1[i1];
    if (idx2 >= 0) {
      const int idx1 = vnMatch2[idx2];
      if (idx1 == i1) { matches12[i1] = idx2; ++nFound; }
    }
  }
  *nfound = nFound;
  return 0;
}

// ORBmatcher::SearchForInitialization, ORBmatcher.cc:405-520
int orc_match_initialization(const psl_keypoint* kps1_un, const uint8_t* desc1, int n1, float* prev_matched,
                             const psl_frame_view* f2, int window_size, float nn_ratio, int th_low,
                             int check_orientation, int32_t* matches12, int32_t* nmatches) {
  int nm = 0;
  for (int i = 0; i < n1; ++i) matches12[i] = -1;
  std::vector<int> rotHist[HISTO_LENGTH];
  std::vector<int> vMatchedDistance((size_t)f2->n, INT32_MAX), vnMatches21((size_t)f2->n, -1);
  Grid g(*f2);
  std::vector<int> vIndices2;
  for (int i1 = 0; i1 < n1; ++i1) {
    const int level1 = kps1_un[i1].octave;
    if (level1 > 0) continue;
    features_in_area(*f2, g, prev_matched[2 * i1], prev_matched[2 * i1 + 1], (float)window_size, level1, level1, vIndices2);
    if (vIndices2.empty()) continue;
    const uint8_t* d1 = desc1 + 32 * (size_t)i1;
    int bestDist = INT32_MAX, bestDist2 = INT32_MAX, bestIdx2 = -1;
    for (int i2 : vIndices2) {
      const int dist = desc_dist(d1, f2->desc + 32 * (size_t)i2);
      if (vMatchedDistance[i2] <= dist) continue;
      if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
      else if (dist < bestDist2) bestDist2 = dist;
    }
    if (bestDist <= th_low) {
      if (bestDist < (float)bestDist2 * nn_ratio) {
        if (vnMatches21[bestIdx2] >= 0) { matches12[vnMatches21[bestIdx2]] = -1; --nm; }
        matches12[i1] = bestIdx2;
        vnMatches21[bestIdx2] = i1;
        vMatchedDistance[bestIdx2] = bestDist;
        ++nm;
        if (check_orientation) rotHist[rot_bin(kps1_un[i1].angle, f2->kps_un[bestIdx2].angle)].push_back(i1);
      }
    }
  }
  if (check_orientation) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; ++i) {
      if (i == ind1 || i == ind2 || i == ind3) continue;
      for (int idx1 : rotHist[i])
        if (matches12[idx1] >= 0) { matches12[idx1] = -1; --nm; }
    }
  }
  for (int i1 = 0; i1 < n1; ++i1)  // :514-517
    if (matches12[i1] >= 0) {
      prev_matched[2 * i1] = f2->kps_un[matches12[i1]].x;
      prev_matched[2 * i1 + 1] = f2->kps_un[matches12[i1]].y;
    }
  *nmatches = nm;
  return 0;
}

}  // extern "C"
