// C-ABI of the loop-closing / initialisation matchers ("next" row N1, second batch; include/psl_frontend.h):
// SearchByBoW(KF, KF), SearchBySim3, SearchForInitialization.  Host-pointer, single-pair entry points like the
// rest of the matcher API: the plain arrays are staged in HBM, the kernels of match_kernels.cu run with B = 1 and
// the result is copied back.  (The Sim3 forms of Fuse and SearchByProjection are settings of psl_match_fuse /
// psl_match_projection, see the header.)
#include <algorithm>
#include <cstring>
#include <vector>

#include "match_kernels.cuh"
#include "psl_ctx.cuh"

using namespace psl;

#define PSL_UP(buf, src, nbytes)                                                                    \
  do {                                                                                              \
    int rc__ = ensure(ctx, buf, (nbytes));                                                          \
    if (rc__) return rc__;                                                                          \
    if ((nbytes) > 0) PSL_CK(cudaMemcpyAsync((buf).p, (src), (nbytes), cudaMemcpyHostToDevice, ctx->stream)); \
  } while (0)

namespace {
bool view_ok(const psl_frame_view* v) {
  return v && v->n >= 0 && v->n <= 65535 && (v->n == 0 || (v->kps_un && v->desc));
}
}  // namespace

extern "C" {

int psl_match_bow_kf(psl_ctx* ctx, const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int32_t n1,
                     const psl_feature_vector* fv1, const uint8_t* desc2, const float* angle2, const uint8_t* valid2,
                     int32_t n2, const psl_feature_vector* fv2, float nn_ratio, int32_t th_low,
                     int32_t check_orientation, int32_t* matches12, int32_t* nmatches) {
  if (!ctx) return PSL_E_INVALID;
  if (!fv1 || !fv2 || !nmatches || n1 < 0 || n2 < 0 || (n1 > 0 && (!desc1 || !angle1 || !valid1 || !matches12)) ||
      (n2 > 0 && (!desc2 || !angle2 || !valid2)))
    return fail(ctx, PSL_E_INVALID, "bad argument");
  *nmatches = 0;
  for (int i = 0; i < n1; ++i) matches12[i] = -1;
  if (n1 == 0 || n2 == 0) return PSL_OK;
  // merge walk over the two sorted node lists (std::map iteration + lower_bound, ORBmatcher.cc:550-620)
  std::vector<int2> pairs;
  for (int a = 0, b = 0; a < fv1->n_nodes && b < fv2->n_nodes;) {
    if (fv1->node_id[a] == fv2->node_id[b]) pairs.push_back(make_int2(a++, b++));
    else if (fv1->node_id[a] < fv2->node_id[b]) ++a;
    else ++b;
  }
  const size_t ni1 = fv1->n_nodes ? (size_t)fv1->offs[fv1->n_nodes] : 0, ni2 = fv2->n_nodes ? (size_t)fv2->offs[fv2->n_nodes] : 0;
  for (size_t i = 0; i < ni1; ++i) if ((int)fv1->idx[i] >= n1) return fail(ctx, PSL_E_INVALID, "KF1 feature index out of range");
  for (size_t i = 0; i < ni2; ++i) if ((int)fv2->idx[i] >= n2) return fail(ctx, PSL_E_INVALID, "KF2 feature index out of range");
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  DevBuf* M = ctx->m_misc;
  PSL_UP(M[0], desc1, (size_t)n1 * 32);
  PSL_UP(M[1], angle1, (size_t)n1 * 4);
  PSL_UP(M[2], valid1, (size_t)n1);
  PSL_UP(M[3], fv1->offs, (size_t)(fv1->n_nodes + 1) * 4);
  PSL_UP(M[4], fv1->idx, ni1 * 4);
  PSL_UP(M[5], desc2, (size_t)n2 * 32);
  PSL_UP(M[6], angle2, (size_t)n2 * 4);
  PSL_UP(M[7], fv2->offs, (size_t)(fv2->n_nodes + 1) * 4);
  PSL_UP(M[8], fv2->idx, ni2 * 4);
  PSL_UP(M[9], pairs.data(), pairs.size() * sizeof(int2));
  PSL_UP(ctx->m_claimed, valid2, (size_t)n2);
  int rc;
  if ((rc = ensure(ctx, M[10], (size_t)n2 * 4 + 34 * 4))) return rc;  // match over KF2 | hist[32] | n_accepted | nmatches
  if ((rc = ensure(ctx, M[11], (size_t)n1 * 4))) return rc;           // accepted list
  if ((rc = ensure(ctx, ctx->m_assign, (size_t)n1 * 4))) return rc;   // matches12
  int32_t* d_match = M[10].as<int32_t>();
  int32_t* d_hist = d_match + n2;
  launch_bow(M[0].as<uint8_t>(), M[1].as<float>(), M[2].as<uint8_t>(), M[3].as<int32_t>(), M[4].as<uint32_t>(),
             M[5].as<uint8_t>(), M[6].as<float>(), ctx->m_claimed.as<uint8_t>(), M[7].as<int32_t>(), M[8].as<uint32_t>(),
             M[9].as<int2>(), (int)pairs.size(), nn_ratio, th_low, 1, check_orientation, n2, d_match, d_hist,
             M[11].as<uint32_t>(), d_hist + 32, d_hist + 33, ctx->m_assign.as<int32_t>(), n1, ctx->stream);
  prof_span(ctx, 5, prof_mark(ctx), 3);
  PSL_CK(cudaGetLastError());
  PSL_CK(cudaMemcpyAsync(matches12, ctx->m_assign.p, (size_t)n1 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  PSL_CK(cudaMemcpyAsync(nmatches, d_hist + 33, 4, cudaMemcpyDeviceToHost, ctx->stream));
  return check_status(ctx);
}

int psl_match_sim3(psl_ctx* ctx, const psl_frame_view* kf1, const psl_frame_view* kf2, const psl_fuse_query* q12,
                   const uint8_t* mp_desc1, const psl_fuse_query* q21, const uint8_t* mp_desc2, int32_t th_high,
                   int32_t* matches12, int32_t* nfound) {
  if (!ctx) return PSL_E_INVALID;
  if (!view_ok(kf1) || !view_ok(kf2) || !nfound || (kf1->n > 0 && (!q12 || !mp_desc1 || !matches12)) ||
      (kf2->n > 0 && (!q21 || !mp_desc2)))
    return fail(ctx, PSL_E_INVALID, "bad argument");
  *nfound = 0;
  const int n1 = kf1->n, n2 = kf2->n;
  for (int i = 0; i < n1; ++i) matches12[i] = -1;
  if (n1 == 0 || n2 == 0) return PSL_OK;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  DevBuf* M = ctx->m_misc;
  // keyframe 1: m_kps / m_desc / grid in m_cell_*; keyframe 2: M[0] / M[1] / grid in M[2], M[3]
  PSL_UP(ctx->m_kps, kf1->kps_un, (size_t)n1 * sizeof(psl_keypoint));
  PSL_UP(ctx->m_desc, kf1->desc, (size_t)n1 * 32);
  PSL_UP(M[0], kf2->kps_un, (size_t)n2 * sizeof(psl_keypoint));
  PSL_UP(M[1], kf2->desc, (size_t)n2 * 32);
  PSL_UP(M[4], q12, (size_t)n1 * sizeof(psl_fuse_query));
  PSL_UP(M[5], mp_desc1, (size_t)n1 * 32);
  PSL_UP(M[6], q21, (size_t)n2 * sizeof(psl_fuse_query));
  PSL_UP(M[7], mp_desc2, (size_t)n2 * 32);
  const int32_t nn[2] = {n1, n2};
  PSL_UP(ctx->m_n, nn, sizeof(nn));
  int rc;
  if ((rc = ensure(ctx, ctx->m_cell_start, (size_t)(kGridCells + 1) * 4))) return rc;
  if ((rc = ensure(ctx, ctx->m_cell_items, (size_t)n1 * 2))) return rc;
  if ((rc = ensure(ctx, M[2], (size_t)(kGridCells + 1) * 4))) return rc;
  if ((rc = ensure(ctx, M[3], (size_t)n2 * 2))) return rc;
  if ((rc = ensure(ctx, M[8], (size_t)n1 * 4))) return rc;           // vnMatch1
  if ((rc = ensure(ctx, M[9], (size_t)n2 * 4))) return rc;           // vnMatch2
  if ((rc = ensure(ctx, ctx->m_assign, (size_t)n1 * 4))) return rc;  // matches12
  if ((rc = ensure(ctx, ctx->m_nm, 4))) return rc;
  MatchFrames F1{ctx->m_kps.as<psl_keypoint>(), nullptr, ctx->m_desc.as<uint8_t>(), ctx->m_n.as<int32_t>(), n1,
                 kf1->min_x, kf1->min_y, kf1->grid_w_inv, kf1->grid_h_inv};
  MatchFrames F2{M[0].as<psl_keypoint>(), nullptr, M[1].as<uint8_t>(), ctx->m_n.as<int32_t>() + 1, n2,
                 kf2->min_x, kf2->min_y, kf2->grid_w_inv, kf2->grid_h_inv};
  size_t e = prof_mark(ctx);
  launch_grid_build(F1, ctx->m_cell_start.as<int32_t>(), ctx->m_cell_items.as<uint16_t>(), 1, ctx->stream);
  launch_grid_build(F2, M[2].as<int32_t>(), M[3].as<uint16_t>(), 1, ctx->stream);
  // KF1's points searched in KF2 (:1140-1216), KF2's points searched in KF1 (:1220-1296)
  launch_fuse(F2, M[4].as<psl_fuse_query>(), M[5].as<uint8_t>(), n1, M[2].as<int32_t>(), M[3].as<uint16_t>(), nullptr,
              th_high, M[8].as<int32_t>(), nullptr, ctx->stream);
  launch_fuse(F1, M[6].as<psl_fuse_query>(), M[7].as<uint8_t>(), n2, ctx->m_cell_start.as<int32_t>(),
              ctx->m_cell_items.as<uint16_t>(), nullptr, th_high, M[9].as<int32_t>(), nullptr, ctx->stream);
  launch_sim3_agree(M[8].as<int32_t>(), n1, M[9].as<int32_t>(), ctx->m_assign.as<int32_t>(), ctx->m_nm.as<int32_t>(),
                    ctx->stream);
  prof_span(ctx, 5, e, 5);
  PSL_CK(cudaGetLastError());
  PSL_CK(cudaMemcpyAsync(matches12, ctx->m_assign.p, (size_t)n1 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  PSL_CK(cudaMemcpyAsync(nfound, ctx->m_nm.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  return check_status(ctx);
}

int psl_match_initialization(psl_ctx* ctx, const psl_keypoint* kps1_un, const uint8_t* desc1, int32_t n1,
                             float* prev_matched, const psl_frame_view* f2, int32_t window_size, float nn_ratio,
                             int32_t th_low, int32_t check_orientation, int32_t* matches12, int32_t* nmatches) {
  if (!ctx) return PSL_E_INVALID;
  if (!view_ok(f2) || !nmatches || n1 < 0 || n1 > 65535 || window_size < 0 ||
      (n1 > 0 && (!kps1_un || !desc1 || !prev_matched || !matches12)))
    return fail(ctx, PSL_E_INVALID, "bad argument");
  *nmatches = 0;
  for (int i = 0; i < n1; ++i) matches12[i] = -1;
  const int n2 = f2->n;
  if (n1 == 0 || n2 == 0) return PSL_OK;
  // one window per F1 keypoint of level 0 (:418-425): GetFeaturesInArea(prev.x, prev.y, windowSize, level1, level1)
  std::vector<psl_proj_query> q((size_t)n1);
  for (int i = 0; i < n1; ++i) {
    const int level1 = kps1_un[i].octave;
    q[i] = psl_proj_query{prev_matched[2 * i], prev_matched[2 * i + 1], (float)window_size, level1, level1, 0.f,
                          kps1_un[i].angle, level1 > 0 ? 0u : PSL_Q_VALID};
  }
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  DevBuf* M = ctx->m_misc;
  PSL_UP(ctx->m_kps, f2->kps_un, (size_t)n2 * sizeof(psl_keypoint));
  PSL_UP(ctx->m_desc, f2->desc, (size_t)n2 * 32);
  PSL_UP(ctx->m_q, q.data(), (size_t)n1 * sizeof(psl_proj_query));
  PSL_UP(ctx->m_qdesc, desc1, (size_t)n1 * 32);
  PSL_UP(M[0], kps1_un, (size_t)n1 * sizeof(psl_keypoint));
  PSL_UP(M[1], prev_matched, (size_t)n1 * 8);
  const int32_t nn[2] = {n2, n1};
  PSL_UP(ctx->m_n, nn, sizeof(nn));
  int rc;
  if ((rc = ensure(ctx, ctx->m_cell_start, (size_t)(kGridCells + 1) * 4))) return rc;
  if ((rc = ensure(ctx, ctx->m_cell_items, (size_t)n2 * 2))) return rc;
  if ((rc = ensure(ctx, ctx->m_cand, (size_t)n1 * kCandCap * 4))) return rc;
  if ((rc = ensure(ctx, ctx->m_cand_count, (size_t)n1 * 4))) return rc;
  if ((rc = ensure(ctx, ctx->m_best, (size_t)n1 * 8))) return rc;
  if ((rc = ensure(ctx, ctx->m_accepted, (size_t)n1 * 4))) return rc;
  if ((rc = ensure(ctx, ctx->m_assign, (size_t)n1 * 4))) return rc;
  if ((rc = ensure(ctx, ctx->m_nm, 4))) return rc;
  if ((rc = ensure(ctx, M[2], (size_t)n2 * 4))) return rc;  // vMatchedDistance
  if ((rc = ensure(ctx, M[3], (size_t)n2 * 4))) return rc;  // vnMatches21
  MatchFrames F{ctx->m_kps.as<psl_keypoint>(), nullptr, ctx->m_desc.as<uint8_t>(), ctx->m_n.as<int32_t>(), n2,
                f2->min_x, f2->min_y, f2->grid_w_inv, f2->grid_h_inv};
  MatchQueries Q{ctx->m_q.as<psl_proj_query>(), ctx->m_qdesc.as<uint8_t>(), ctx->m_n.as<int32_t>() + 1, n1};
  size_t e = prof_mark(ctx);
  launch_grid_build(F, ctx->m_cell_start.as<int32_t>(), ctx->m_cell_items.as<uint16_t>(), 1, ctx->stream);
  launch_proj_candidates(F, Q, ctx->m_cell_start.as<int32_t>(), ctx->m_cell_items.as<uint16_t>(),
                         ctx->m_cand.as<uint32_t>(), ctx->m_cand_count.as<int32_t>(), ctx->m_best.as<uint2>(),
                         ctx->d_status, 1, ctx->stream);
  launch_init_resolve(ctx->m_cand.as<uint32_t>(), ctx->m_cand_count.as<int32_t>(), M[0].as<psl_keypoint>(), n1,
                      ctx->m_kps.as<psl_keypoint>(), n2, nn_ratio, th_low, check_orientation, M[2].as<int32_t>(),
                      M[3].as<int32_t>(), ctx->m_accepted.as<uint32_t>(), ctx->m_assign.as<int32_t>(), M[1].as<float>(),
                      ctx->m_nm.as<int32_t>(), ctx->stream);
  prof_span(ctx, 5, e, 3);
  PSL_CK(cudaGetLastError());
  // results land in caller memory only if the call succeeds (a window with more than 256 candidates is refused)
  std::vector<float> pm((size_t)n1 * 2);
  std::vector<int32_t> m12((size_t)n1);
  int32_t nm = 0;
  PSL_CK(cudaMemcpyAsync(m12.data(), ctx->m_assign.p, (size_t)n1 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  PSL_CK(cudaMemcpyAsync(pm.data(), M[1].p, (size_t)n1 * 8, cudaMemcpyDeviceToHost, ctx->stream));
  PSL_CK(cudaMemcpyAsync(&nm, ctx->m_nm.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  rc = check_status(ctx);
  if (rc == PSL_E_CAPACITY) return fail(ctx, rc, "more than 256 candidates in one search window");
  if (rc) return rc;
  std::memcpy(matches12, m12.data(), (size_t)n1 * 4);
  std::memcpy(prev_matched, pm.data(), (size_t)n1 * 8);
  *nmatches = nm;
  return PSL_OK;
}

}  // extern "C"
