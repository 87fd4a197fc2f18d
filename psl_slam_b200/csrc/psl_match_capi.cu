// C-ABI of the matchers (include/psl_frontend.h): host-pointer, single-pair entry points that
// stage the plain arrays in HBM, run the batch kernels with B = 1 and copy the result back.
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

extern "C" {

int psl_descriptor_distance(psl_ctx* ctx, const uint8_t* a, const uint8_t* b, int32_t n, int32_t* dist) {
  if (!ctx) return PSL_E_INVALID;
  if (n < 0 || (n > 0 && (!a || !b || !dist))) return fail(ctx, PSL_E_INVALID, "bad argument");
  if (n == 0) return PSL_OK;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  PSL_UP(ctx->m_desc, a, (size_t)n * 32);
  PSL_UP(ctx->m_qdesc, b, (size_t)n * 32);
  int rc = ensure(ctx, ctx->m_assign, (size_t)n * 4);
  if (rc) return rc;
  launch_descriptor_distance(ctx->m_desc.as<uint8_t>(), ctx->m_qdesc.as<uint8_t>(), n, ctx->m_assign.as<int32_t>(),
                             ctx->stream);
  prof_span(ctx, 5, prof_mark(ctx), 1);
  PSL_CK(cudaMemcpyAsync(dist, ctx->m_assign.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  return check_status(ctx);
}

int psl_hamming_knn2(psl_ctx* ctx, const uint8_t* q, int32_t nq, const uint8_t* t, int32_t nt, int32_t* idx,
                     int32_t* dist) {
  if (!ctx) return PSL_E_INVALID;
  if (nq < 0 || nt < 0 || nt > 65535 || (nq > 0 && (!q || !idx || !dist)) || (nt > 0 && !t))
    return fail(ctx, PSL_E_INVALID, "bad argument (nt <= 65535)");
  if (nq == 0) return PSL_OK;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  PSL_UP(ctx->m_qdesc, q, (size_t)nq * 32);
  PSL_UP(ctx->m_desc, t, (size_t)nt * 32);
  int rc = ensure(ctx, ctx->m_assign, (size_t)nq * 8);
  if (rc) return rc;
  if ((rc = ensure(ctx, ctx->m_cand_count, (size_t)nq * 8))) return rc;
  launch_knn2(ctx->m_qdesc.as<uint8_t>(), nq, ctx->m_desc.as<uint8_t>(), nt, ctx->m_assign.as<int32_t>(),
              ctx->m_cand_count.as<int32_t>(), ctx->stream);
  prof_span(ctx, 5, prof_mark(ctx), 1);
  PSL_CK(cudaMemcpyAsync(idx, ctx->m_assign.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, ctx->stream));
  PSL_CK(cudaMemcpyAsync(dist, ctx->m_cand_count.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, ctx->stream));
  return check_status(ctx);
}

int psl_match_projection(psl_ctx* ctx, const psl_frame_view* fv, const psl_proj_query* queries,
                         const uint8_t* query_desc, int32_t nq, const uint8_t* claimed_in,
                         const psl_match_params* prm, int32_t* assign, int32_t* nmatches) {
  if (!ctx) return PSL_E_INVALID;
  // the resolve stage packs an accepted match as query << 16 | keypoint: both sides are bounded by 65535
  if (!fv || !prm || !nmatches || nq < 0 || nq > 65535 || fv->n < 0 || fv->n > 65535 || (fv->n > 0 && (!fv->kps_un || !fv->desc || !assign)) ||
      (nq > 0 && (!queries || !query_desc)) || (prm->mode != 0 && prm->mode != 1))
    return fail(ctx, PSL_E_INVALID, "bad argument");
  *nmatches = 0;
  const int n = fv->n;
  if (n == 0) return PSL_OK;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  const int ncap = std::max(n, 1), qcap = std::max(nq, 1);
  PSL_UP(ctx->m_kps, fv->kps_un, (size_t)n * sizeof(psl_keypoint));
  if (fv->u_right) PSL_UP(ctx->m_ur, fv->u_right, (size_t)n * 4);
  PSL_UP(ctx->m_desc, fv->desc, (size_t)n * 32);
  PSL_UP(ctx->m_q, queries, (size_t)nq * sizeof(psl_proj_query));
  PSL_UP(ctx->m_qdesc, query_desc, (size_t)nq * 32);
  if (claimed_in) PSL_UP(ctx->m_claimed, claimed_in, (size_t)n);
  const int32_t nn[2] = {n, nq};
  PSL_UP(ctx->m_n, nn, sizeof(nn));
  int rc;
  if ((rc = ensure(ctx, ctx->m_cell_start, (size_t)(kGridCells + 1) * 4))) return rc;
  if ((rc = ensure(ctx, ctx->m_cell_items, (size_t)ncap * 2))) return rc;
  if ((rc = ensure(ctx, ctx->m_cand, (size_t)qcap * kCandCap * 4))) return rc;
  if ((rc = ensure(ctx, ctx->m_cand_count, (size_t)qcap * 4))) return rc;
  if ((rc = ensure(ctx, ctx->m_best, (size_t)qcap * 8))) return rc;
  if ((rc = ensure(ctx, ctx->m_accepted, (size_t)qcap * 4))) return rc;
  if ((rc = ensure(ctx, ctx->m_assign, (size_t)ncap * 4))) return rc;
  if ((rc = ensure(ctx, ctx->m_nm, 4))) return rc;
  MatchFrames F{ctx->m_kps.as<psl_keypoint>(), fv->u_right ? ctx->m_ur.as<float>() : nullptr, ctx->m_desc.as<uint8_t>(),
                ctx->m_n.as<int32_t>(), ncap, fv->min_x, fv->min_y, fv->grid_w_inv, fv->grid_h_inv};
  MatchQueries Q{ctx->m_q.as<psl_proj_query>(), ctx->m_qdesc.as<uint8_t>(), ctx->m_n.as<int32_t>() + 1, qcap};
  size_t e = prof_mark(ctx);
  launch_grid_build(F, ctx->m_cell_start.as<int32_t>(), ctx->m_cell_items.as<uint16_t>(), 1, ctx->stream);
  launch_proj_candidates(F, Q, ctx->m_cell_start.as<int32_t>(), ctx->m_cell_items.as<uint16_t>(),
                         ctx->m_cand.as<uint32_t>(), ctx->m_cand_count.as<int32_t>(), ctx->m_best.as<uint2>(),
                         ctx->d_status, 1, ctx->stream);
  launch_proj_resolve(F, Q, ctx->m_cand.as<uint32_t>(), ctx->m_cand_count.as<int32_t>(), ctx->m_best.as<uint2>(),
                      claimed_in ? ctx->m_claimed.as<uint8_t>() : nullptr, *prm, ctx->m_accepted.as<uint32_t>(),
                      ctx->m_assign.as<int32_t>(), ctx->m_nm.as<int32_t>(), 1, ctx->stream);
  prof_span(ctx, 5, e, 3);
  PSL_CK(cudaGetLastError());
  PSL_CK(cudaMemcpyAsync(assign, ctx->m_assign.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  PSL_CK(cudaMemcpyAsync(nmatches, ctx->m_nm.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  rc = check_status(ctx);
  if (rc == PSL_E_CAPACITY) return fail(ctx, rc, "more than 256 candidates in one search window");
  return rc;
}

int psl_match_bow(psl_ctx* ctx, const uint8_t* kf_desc, const float* kf_angle, const uint8_t* kf_valid, int32_t nkf,
                  const psl_feature_vector* kfv, const uint8_t* f_desc, const float* f_angle, int32_t nf,
                  const psl_feature_vector* ffv, float nn_ratio, int32_t th_low, int32_t check_orientation,
                  int32_t* match_f, int32_t* nmatches) {
  if (!ctx) return PSL_E_INVALID;
  if (!kfv || !ffv || !nmatches || nkf < 0 || nf < 0 || (nf > 0 && (!f_desc || !f_angle || !match_f)) ||
      (nkf > 0 && (!kf_desc || !kf_angle || !kf_valid)))
    return fail(ctx, PSL_E_INVALID, "bad argument");
  *nmatches = 0;
  for (int i = 0; i < nf; ++i) match_f[i] = -1;
  if (nf == 0 || nkf == 0) return PSL_OK;
  // merge walk over the two sorted node lists (std::map iteration + lower_bound, ORBmatcher.cc:180-262)
  std::vector<int2> pairs;
  for (int a = 0, b = 0; a < kfv->n_nodes && b < ffv->n_nodes;) {
    if (kfv->node_id[a] == ffv->node_id[b]) pairs.push_back(make_int2(a++, b++));
    else if (kfv->node_id[a] < ffv->node_id[b]) ++a;
    else ++b;
  }
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  const size_t nki = kfv->n_nodes ? (size_t)kfv->offs[kfv->n_nodes] : 0, nfi = ffv->n_nodes ? (size_t)ffv->offs[ffv->n_nodes] : 0;
  for (size_t i = 0; i < nki; ++i) if ((int)kfv->idx[i] >= nkf) return fail(ctx, PSL_E_INVALID, "kf feature index out of range");
  for (size_t i = 0; i < nfi; ++i) if ((int)ffv->idx[i] >= nf) return fail(ctx, PSL_E_INVALID, "frame feature index out of range");
  DevBuf* M = ctx->m_misc;
  PSL_UP(M[0], kf_desc, (size_t)nkf * 32);
  PSL_UP(M[1], kf_angle, (size_t)nkf * 4);
  PSL_UP(M[2], kf_valid, (size_t)nkf);
  PSL_UP(M[3], kfv->offs, (size_t)(kfv->n_nodes + 1) * 4);
  PSL_UP(M[4], kfv->idx, nki * 4);
  PSL_UP(M[5], f_desc, (size_t)nf * 32);
  PSL_UP(M[6], f_angle, (size_t)nf * 4);
  PSL_UP(M[7], ffv->offs, (size_t)(ffv->n_nodes + 1) * 4);
  PSL_UP(M[8], ffv->idx, nfi * 4);
  PSL_UP(M[9], pairs.data(), pairs.size() * sizeof(int2));
  int rc;
  if ((rc = ensure(ctx, M[10], (size_t)nf * 4 + 34 * 4))) return rc;      // match_f | hist[32] | n_accepted | nmatches
  if ((rc = ensure(ctx, M[11], (size_t)std::max(nkf, 1) * 4))) return rc;  // accepted list
  int32_t* d_match = M[10].as<int32_t>();
  int32_t* d_hist = d_match + nf;
  launch_bow(M[0].as<uint8_t>(), M[1].as<float>(), M[2].as<uint8_t>(), M[3].as<int32_t>(), M[4].as<uint32_t>(),
             M[5].as<uint8_t>(), M[6].as<float>(), nullptr, M[7].as<int32_t>(), M[8].as<uint32_t>(), M[9].as<int2>(),
             (int)pairs.size(), nn_ratio, th_low, 0, check_orientation, nf, d_match, d_hist, M[11].as<uint32_t>(),
             d_hist + 32, d_hist + 33, nullptr, nkf, ctx->stream);
  prof_span(ctx, 5, prof_mark(ctx), 2);
  PSL_CK(cudaGetLastError());
  PSL_CK(cudaMemcpyAsync(match_f, d_match, (size_t)nf * 4, cudaMemcpyDeviceToHost, ctx->stream));
  PSL_CK(cudaMemcpyAsync(nmatches, d_hist + 33, 4, cudaMemcpyDeviceToHost, ctx->stream));
  return check_status(ctx);
}


int psl_match_triangulation(psl_ctx* ctx, const psl_keyframe_view* kf1, const psl_feature_vector* fv1,
                            const psl_keyframe_view* kf2, const psl_feature_vector* fv2, const float* F12, float ex,
                            float ey, const float* scale_factors2, const float* level_sigma2_2, int32_t nlevels,
                            int32_t only_stereo, int32_t check_orientation, int32_t th_low, int32_t* matches12,
                            int32_t* nmatches) {
  if (!ctx) return PSL_E_INVALID;
  if (!kf1 || !kf2 || !fv1 || !fv2 || !F12 || !scale_factors2 || !level_sigma2_2 || !nmatches || nlevels < 1 ||
      nlevels > kMaxLevels || kf1->n < 0 || kf2->n < 0 || kf2->n > 65535 ||
      (kf1->n > 0 && (!kf1->kps_un || !kf1->u_right || !kf1->desc || !kf1->has_mappoint || !matches12)) ||
      (kf2->n > 0 && (!kf2->kps_un || !kf2->u_right || !kf2->desc || !kf2->has_mappoint)))
    return fail(ctx, PSL_E_INVALID, "bad argument");
  *nmatches = 0;
  const int n1 = kf1->n, n2 = kf2->n;
  for (int i = 0; i < n1; ++i) matches12[i] = -1;
  if (n1 == 0 || n2 == 0) return PSL_OK;
  for (int i = 0; i < n2; ++i)
    if (kf2->kps_un[i].octave < 0 || kf2->kps_un[i].octave >= nlevels) return fail(ctx, PSL_E_INVALID, "octave out of range");
  std::vector<int2> pairs;
  for (int a = 0, b = 0; a < fv1->n_nodes && b < fv2->n_nodes;) {
    if (fv1->node_id[a] == fv2->node_id[b]) pairs.push_back(make_int2(a++, b++));
    else if (fv1->node_id[a] < fv2->node_id[b]) ++a;
    else ++b;
  }
  const size_t ni1 = fv1->n_nodes ? (size_t)fv1->offs[fv1->n_nodes] : 0, ni2 = fv2->n_nodes ? (size_t)fv2->offs[fv2->n_nodes] : 0;
  for (size_t i = 0; i < ni1; ++i) if ((int)fv1->idx[i] >= n1) return fail(ctx, PSL_E_INVALID, "feature index out of range");
  for (size_t i = 0; i < ni2; ++i) if ((int)fv2->idx[i] >= n2) return fail(ctx, PSL_E_INVALID, "feature index out of range");
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  DevBuf* M = ctx->m_misc;
  PSL_UP(M[0], kf1->kps_un, (size_t)n1 * sizeof(psl_keypoint));
  PSL_UP(M[1], kf1->u_right, (size_t)n1 * 4);
  PSL_UP(ctx->m_qdesc, kf1->desc, (size_t)n1 * 32);
  PSL_UP(M[2], kf1->has_mappoint, (size_t)n1);
  PSL_UP(M[3], fv1->offs, (size_t)(fv1->n_nodes + 1) * 4);
  PSL_UP(M[4], fv1->idx, ni1 * 4);
  PSL_UP(M[5], kf2->kps_un, (size_t)n2 * sizeof(psl_keypoint));
  PSL_UP(M[6], kf2->u_right, (size_t)n2 * 4);
  PSL_UP(ctx->m_desc, kf2->desc, (size_t)n2 * 32);
  PSL_UP(M[7], kf2->has_mappoint, (size_t)n2);
  PSL_UP(M[8], fv2->offs, (size_t)(fv2->n_nodes + 1) * 4);
  PSL_UP(M[9], fv2->idx, ni2 * 4);
  PSL_UP(M[10], pairs.data(), pairs.size() * sizeof(int2));
  float tab[9 + 2 * kMaxLevels] = {0};
  std::memcpy(tab, F12, 36);
  std::memcpy(tab + 9, scale_factors2, (size_t)nlevels * 4);
  std::memcpy(tab + 9 + kMaxLevels, level_sigma2_2, (size_t)nlevels * 4);
  PSL_UP(M[11], tab, sizeof(tab));
  int rc;
  if ((rc = ensure(ctx, ctx->m_assign, (size_t)n1 * 4))) return rc;
  if ((rc = ensure(ctx, ctx->m_nm, 33 * 4))) return rc;
  const float* d_tab = M[11].as<float>();
  int32_t* d_hist = ctx->m_nm.as<int32_t>();
  launch_triangulation(M[0].as<psl_keypoint>(), M[1].as<float>(), ctx->m_qdesc.as<uint8_t>(), M[2].as<uint8_t>(),
                       M[3].as<int32_t>(), M[4].as<uint32_t>(), n1, M[5].as<psl_keypoint>(), M[6].as<float>(),
                       ctx->m_desc.as<uint8_t>(), M[7].as<uint8_t>(), M[8].as<int32_t>(), M[9].as<uint32_t>(),
                       M[10].as<int2>(), (int)pairs.size(), d_tab, ex, ey, d_tab + 9, d_tab + 9 + kMaxLevels, only_stereo,
                       th_low, check_orientation, ctx->m_assign.as<int32_t>(), d_hist, d_hist + 32, ctx->stream);
  prof_span(ctx, 5, prof_mark(ctx), 2);
  PSL_CK(cudaGetLastError());
  PSL_CK(cudaMemcpyAsync(matches12, ctx->m_assign.p, (size_t)n1 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  PSL_CK(cudaMemcpyAsync(nmatches, d_hist + 32, 4, cudaMemcpyDeviceToHost, ctx->stream));
  return check_status(ctx);
}


int psl_match_fuse(psl_ctx* ctx, const psl_frame_view* kf, const psl_fuse_query* queries, const uint8_t* query_desc,
                   int32_t nq, const float* inv_level_sigma2, int32_t nlevels, int32_t th_low, int32_t* best_idx,
                   int32_t* best_dist) {
  if (!ctx) return PSL_E_INVALID;
  if (!kf || nq < 0 || kf->n < 0 || kf->n > 65535 || nlevels < 1 || nlevels > kMaxLevels ||
      (kf->n > 0 && (!kf->kps_un || !kf->desc)) || (nq > 0 && (!queries || !query_desc || !best_idx)))
    return fail(ctx, PSL_E_INVALID, "bad argument");
  if (nq == 0) return PSL_OK;
  const int n = kf->n;
  if (n == 0) {
    for (int i = 0; i < nq; ++i) { best_idx[i] = -1; if (best_dist) best_dist[i] = 256; }
    return PSL_OK;
  }
  for (int i = 0; i < n; ++i)
    if (kf->kps_un[i].octave < 0 || kf->kps_un[i].octave >= nlevels) return fail(ctx, PSL_E_INVALID, "octave out of range");
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  DevBuf* M = ctx->m_misc;
  PSL_UP(ctx->m_kps, kf->kps_un, (size_t)n * sizeof(psl_keypoint));
  if (kf->u_right) PSL_UP(ctx->m_ur, kf->u_right, (size_t)n * 4);
  PSL_UP(ctx->m_desc, kf->desc, (size_t)n * 32);
  PSL_UP(M[0], queries, (size_t)nq * sizeof(psl_fuse_query));
  PSL_UP(ctx->m_qdesc, query_desc, (size_t)nq * 32);
  float tab[kMaxLevels] = {0};
  if (inv_level_sigma2) std::memcpy(tab, inv_level_sigma2, (size_t)nlevels * 4);  // NULL: the Sim3 form, no gate
  PSL_UP(M[1], tab, sizeof(tab));
  const int32_t nn[1] = {n};
  PSL_UP(ctx->m_n, nn, sizeof(nn));
  int rc;
  if ((rc = ensure(ctx, ctx->m_cell_start, (size_t)(kGridCells + 1) * 4))) return rc;
  if ((rc = ensure(ctx, ctx->m_cell_items, (size_t)n * 2))) return rc;
  if ((rc = ensure(ctx, ctx->m_assign, (size_t)nq * 4))) return rc;
  if ((rc = ensure(ctx, ctx->m_nm, (size_t)nq * 4))) return rc;
  MatchFrames F{ctx->m_kps.as<psl_keypoint>(), kf->u_right ? ctx->m_ur.as<float>() : nullptr, ctx->m_desc.as<uint8_t>(),
                ctx->m_n.as<int32_t>(), n, kf->min_x, kf->min_y, kf->grid_w_inv, kf->grid_h_inv};
  launch_grid_build(F, ctx->m_cell_start.as<int32_t>(), ctx->m_cell_items.as<uint16_t>(), 1, ctx->stream);
  launch_fuse(F, M[0].as<psl_fuse_query>(), ctx->m_qdesc.as<uint8_t>(), nq, ctx->m_cell_start.as<int32_t>(),
              ctx->m_cell_items.as<uint16_t>(), inv_level_sigma2 ? M[1].as<float>() : nullptr, th_low, ctx->m_assign.as<int32_t>(),
              ctx->m_nm.as<int32_t>(), ctx->stream);
  prof_span(ctx, 5, prof_mark(ctx), 2);
  PSL_CK(cudaGetLastError());
  PSL_CK(cudaMemcpyAsync(best_idx, ctx->m_assign.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (best_dist) PSL_CK(cudaMemcpyAsync(best_dist, ctx->m_nm.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream));
  return check_status(ctx);
}

}  // extern "C"
