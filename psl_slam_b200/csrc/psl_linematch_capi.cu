// C-ABI of the line matchers (include/psl_frontend.h): host-pointer, single-pair entry points that stage the
// plain arrays in HBM, run the batch kernels with B = 1 and copy the result back.
#include <algorithm>
#include <cstring>

#include "line_match_kernels.cuh"
#include "psl_ctx.cuh"

using namespace psl;

#define PSL_UP(buf, src, nbytes)                                                                    \
  do {                                                                                              \
    int rc__ = ensure(ctx, buf, (nbytes));                                                          \
    if (rc__) return rc__;                                                                          \
    if ((nbytes) > 0) PSL_CK(cudaMemcpyAsync((buf).p, (src), (nbytes), cudaMemcpyHostToDevice, ctx->stream)); \
  } while (0)
#define PSL_ENS(buf, nbytes)                    \
  do {                                          \
    int rc__ = ensure(ctx, buf, (nbytes));      \
    if (rc__) return rc__;                      \
  } while (0)

namespace {
// uploads desc1 -> m_qdesc, desc2 -> m_desc, (n1, n2) -> m_n and returns the two LineSets (cap = n, B = 1)
int stage_desc_pair(psl_ctx* ctx, const uint8_t* d1, int n1, const uint8_t* d2, int n2, LineSet& A, LineSet& Bs) {
  PSL_UP(ctx->m_qdesc, d1, (size_t)n1 * 32);
  PSL_UP(ctx->m_desc, d2, (size_t)n2 * 32);
  const int32_t nn[2] = {n1, n2};
  PSL_UP(ctx->m_n, nn, sizeof(nn));
  A = LineSet{nullptr, ctx->m_qdesc.as<uint8_t>(), ctx->m_n.as<int32_t>(), std::max(n1, 1)};
  Bs = LineSet{nullptr, ctx->m_desc.as<uint8_t>(), ctx->m_n.as<int32_t>() + 1, std::max(n2, 1)};
  return PSL_OK;
}
bool bad_desc_args(const uint8_t* d1, int n1, const uint8_t* d2, int n2) {
  return n1 < 0 || n2 < 0 || n1 > 65535 || n2 > 65535 || (n1 > 0 && !d1) || (n2 > 0 && !d2);
}
}  // namespace

extern "C" {

int psl_line_match_nnr(psl_ctx* ctx, const uint8_t* desc1, int32_t n1, const uint8_t* desc2, int32_t n2, float nnr,
                       int32_t* matches12, int32_t* nmatches) {
  if (!ctx) return PSL_E_INVALID;
  if (bad_desc_args(desc1, n1, desc2, n2) || !nmatches || (n1 > 0 && !matches12)) return fail(ctx, PSL_E_INVALID, "bad argument");
  *nmatches = 0;
  if (n1 == 0) return PSL_OK;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  LineSet A, T;
  int rc = stage_desc_pair(ctx, desc1, n1, desc2, n2, A, T);
  if (rc) return rc;
  PSL_ENS(ctx->m_best, (size_t)n1 * 8);
  PSL_ENS(ctx->m_assign, (size_t)n1 * 4);
  PSL_ENS(ctx->m_nm, 4);
  size_t e = prof_mark(ctx);
  launch_line_knn2(A, T, ctx->m_best.as<uint2>(), 1, ctx->stream);
  launch_line_nnr(A, ctx->m_best.as<uint2>(), nnr, ctx->m_assign.as<int32_t>(), ctx->m_nm.as<int32_t>(), 1, ctx->stream);
  prof_span(ctx, 15, e, 2);
  PSL_CK(cudaMemcpyAsync(matches12, ctx->m_assign.p, (size_t)n1 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  PSL_CK(cudaMemcpyAsync(nmatches, ctx->m_nm.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  return check_status(ctx);
}

int psl_line_search_geom(psl_ctx* ctx, const psl_keyline* kl_last, const uint8_t* desc_last,
                         const uint8_t* has_mapline_last, int32_t n_last, const psl_keyline* kl_cur,
                         const uint8_t* desc_cur, int32_t n_cur, const float* bounds, float desc_th,
                         int32_t* assign_cur, int32_t* nmatches) {
  if (!ctx) return PSL_E_INVALID;
  if (bad_desc_args(desc_last, n_last, desc_cur, n_cur) || !nmatches || !bounds || (n_last > 0 && (!kl_last || !has_mapline_last)) ||
      (n_cur > 0 && (!kl_cur || !assign_cur)))
    return fail(ctx, PSL_E_INVALID, "bad argument");
  *nmatches = 0;
  if (n_cur == 0) return PSL_OK;  // CurrentFrame.mLdesc.empty(), LSDmatcher.cpp:40-43
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  LineSet L, Cn;
  int rc = stage_desc_pair(ctx, desc_last, n_last, desc_cur, n_cur, L, Cn);
  if (rc) return rc;
  PSL_UP(ctx->m_misc[0], kl_last, (size_t)n_last * sizeof(psl_keyline));
  PSL_UP(ctx->m_misc[1], kl_cur, (size_t)n_cur * sizeof(psl_keyline));
  PSL_UP(ctx->m_claimed, has_mapline_last, (size_t)n_last);
  L.kl = ctx->m_misc[0].as<psl_keyline>();
  Cn.kl = ctx->m_misc[1].as<psl_keyline>();
  PSL_ENS(ctx->m_best, (size_t)std::max(n_last, 1) * 8);
  PSL_ENS(ctx->m_assign, (size_t)n_cur * 4);
  PSL_ENS(ctx->m_nm, 4);
  size_t e = prof_mark(ctx);
  launch_line_knn2(L, Cn, ctx->m_best.as<uint2>(), 1, ctx->stream);
  launch_line_geom(L, ctx->m_claimed.as<uint8_t>(), Cn, ctx->m_best.as<uint2>(), desc_th, bounds[2] - bounds[0],
                   bounds[3] - bounds[1], ctx->m_assign.as<int32_t>(), ctx->m_nm.as<int32_t>(), 1, ctx->stream);
  prof_span(ctx, 15, e, 2);
  PSL_CK(cudaMemcpyAsync(assign_cur, ctx->m_assign.p, (size_t)n_cur * 4, cudaMemcpyDeviceToHost, ctx->stream));
  PSL_CK(cudaMemcpyAsync(nmatches, ctx->m_nm.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  return check_status(ctx);
}

int psl_line_frame_bf_match(psl_ctx* ctx, const uint8_t* desc1, int32_t n1, const uint8_t* desc2, int32_t n2,
                            float nn_ratio, float th, int32_t* matches) {
  if (!ctx) return PSL_E_INVALID;
  if (bad_desc_args(desc1, n1, desc2, n2) || (n1 > 0 && !matches)) return fail(ctx, PSL_E_INVALID, "bad argument");
  if (n1 == 0) return PSL_OK;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  LineSet A, T;
  int rc = stage_desc_pair(ctx, desc1, n1, desc2, n2, A, T);
  if (rc) return rc;
  PSL_ENS(ctx->m_best, (size_t)n1 * 8);
  PSL_ENS(ctx->m_assign, (size_t)n1 * 4);
  size_t e = prof_mark(ctx);
  launch_line_knn2(A, T, ctx->m_best.as<uint2>(), 1, ctx->stream);
  launch_line_bfmatch(A, T, ctx->m_best.as<uint2>(), nn_ratio, th, ctx->m_assign.as<int32_t>(), 1, ctx->stream);
  prof_span(ctx, 15, e, 2);
  PSL_CK(cudaMemcpyAsync(matches, ctx->m_assign.p, (size_t)n1 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  return check_status(ctx);
}

int psl_line_search_double(psl_ctx* ctx, const uint8_t* desc1, int32_t n1, const uint8_t* desc2, int32_t n2,
                           float nn_ratio, float th, int32_t* matches12, int32_t* nmatches) {
  if (!ctx) return PSL_E_INVALID;
  if (bad_desc_args(desc1, n1, desc2, n2) || !nmatches || (n1 > 0 && !matches12)) return fail(ctx, PSL_E_INVALID, "bad argument");
  *nmatches = 0;
  if (n1 == 0) return PSL_OK;
  if (n2 == 0) {  // LSDmatcher.cpp:469-470
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    return PSL_OK;
  }
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  LineSet A, T;
  int rc = stage_desc_pair(ctx, desc1, n1, desc2, n2, A, T);
  if (rc) return rc;
  PSL_ENS(ctx->m_best, (size_t)n1 * 8);
  PSL_ENS(ctx->m_cand_count, (size_t)n2 * 8);
  PSL_ENS(ctx->m_assign, (size_t)n1 * 4);
  PSL_ENS(ctx->m_accepted, (size_t)n2 * 4);
  PSL_ENS(ctx->m_nm, 4);
  cudaStream_t st = ctx->stream;
  size_t e = prof_mark(ctx);
  launch_line_knn2(A, T, ctx->m_best.as<uint2>(), 1, st);
  launch_line_bfmatch(A, T, ctx->m_best.as<uint2>(), nn_ratio, th, ctx->m_assign.as<int32_t>(), 1, st);
  launch_line_knn2(T, A, ctx->m_cand_count.as<uint2>(), 1, st);
  launch_line_bfmatch(T, A, ctx->m_cand_count.as<uint2>(), nn_ratio, th, ctx->m_accepted.as<int32_t>(), 1, st);
  launch_line_mutual(A, ctx->m_accepted.as<int32_t>(), T.cap, ctx->m_assign.as<int32_t>(), ctx->m_nm.as<int32_t>(), 1, st);
  prof_span(ctx, 15, e, 5);
  PSL_CK(cudaMemcpyAsync(matches12, ctx->m_assign.p, (size_t)n1 * 4, cudaMemcpyDeviceToHost, st));
  PSL_CK(cudaMemcpyAsync(nmatches, ctx->m_nm.p, 4, cudaMemcpyDeviceToHost, st));
  return check_status(ctx);
}

int psl_line_search_triangulation(psl_ctx* ctx, const uint8_t* desc1, const uint8_t* has_mapline1, int32_t n1,
                                  const uint8_t* desc2, const uint8_t* has_mapline2, int32_t n2, float nn_ratio, float th,
                                  int32_t is_double, int32_t* matches12, int32_t* nmatches) {
  if (!ctx) return PSL_E_INVALID;
  if (bad_desc_args(desc1, n1, desc2, n2) || !nmatches || (n1 > 0 && (!matches12 || !has_mapline1)) || (n2 > 0 && !has_mapline2))
    return fail(ctx, PSL_E_INVALID, "bad argument");
  *nmatches = 0;
  if (n1 == 0) return PSL_OK;
  if (n2 == 0) {  // LSDmatcher.cpp:714-715
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    return PSL_OK;
  }
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  LineSet A, T;
  int rc = stage_desc_pair(ctx, desc1, n1, desc2, n2, A, T);
  if (rc) return rc;
  PSL_UP(ctx->m_misc[0], has_mapline1, (size_t)n1);
  PSL_UP(ctx->m_misc[1], has_mapline2, (size_t)n2);
  PSL_ENS(ctx->m_best, (size_t)n1 * 8);
  PSL_ENS(ctx->m_cand_count, (size_t)n2 * 8);
  PSL_ENS(ctx->m_assign, (size_t)n1 * 4);
  PSL_ENS(ctx->m_accepted, (size_t)n2 * 4);
  PSL_ENS(ctx->m_nm, 4);
  cudaStream_t st = ctx->stream;
  size_t e = prof_mark(ctx);
  launch_line_knn2(A, T, ctx->m_best.as<uint2>(), 1, st);
  launch_line_bfmatch(A, T, ctx->m_best.as<uint2>(), nn_ratio, th, ctx->m_assign.as<int32_t>(), 1, st);
  launch_line_knn2(T, A, ctx->m_cand_count.as<uint2>(), 1, st);
  launch_line_bfmatch(T, A, ctx->m_cand_count.as<uint2>(), nn_ratio, th, ctx->m_accepted.as<int32_t>(), 1, st);
  launch_line_triang(A, ctx->m_accepted.as<int32_t>(), T.cap, ctx->m_misc[0].as<uint8_t>(), ctx->m_misc[1].as<uint8_t>(),
                     is_double, ctx->m_assign.as<int32_t>(), ctx->m_nm.as<int32_t>(), 1, st);
  prof_span(ctx, 15, e, 5);
  PSL_CK(cudaMemcpyAsync(matches12, ctx->m_assign.p, (size_t)n1 * 4, cudaMemcpyDeviceToHost, st));
  PSL_CK(cudaMemcpyAsync(nmatches, ctx->m_nm.p, 4, cudaMemcpyDeviceToHost, st));
  return check_status(ctx);
}

int psl_line_search_triangulation_new(psl_ctx* ctx, const psl_keyline* kl1, const uint8_t* desc1, const double* func1,
                                      const uint8_t* has_mapline1, int32_t n1, const psl_keyline* kl2,
                                      const uint8_t* desc2, const double* func2, const uint8_t* has_mapline2, int32_t n2,
                                      const float* F21, const float* F12, float nn_ratio, float th, int32_t is_double,
                                      int32_t* matched_pairs, int32_t* nmatches) {
  if (!ctx) return PSL_E_INVALID;
  if (bad_desc_args(desc1, n1, desc2, n2) || !nmatches || !F21 || !F12 ||
      (n1 > 0 && (!matched_pairs || !has_mapline1 || !kl1 || !func1)) || (n2 > 0 && (!has_mapline2 || !kl2 || !func2)))
    return fail(ctx, PSL_E_INVALID, "bad argument");
  *nmatches = 0;
  if (n1 == 0) return PSL_OK;
  for (int i = 0; i < n1; ++i) matched_pairs[i] = -1;
  if (n2 == 0) return PSL_OK;  // LSDmatcher.cpp:792-793
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  LineSet A, T;
  int rc = stage_desc_pair(ctx, desc1, n1, desc2, n2, A, T);
  if (rc) return rc;
  DevBuf* M = ctx->m_misc;
  PSL_UP(M[0], has_mapline1, (size_t)n1);
  PSL_UP(M[1], has_mapline2, (size_t)n2);
  PSL_UP(M[2], kl1, (size_t)n1 * sizeof(psl_keyline));
  PSL_UP(M[3], kl2, (size_t)n2 * sizeof(psl_keyline));
  PSL_UP(M[4], func1, (size_t)n1 * 24);
  PSL_UP(M[5], func2, (size_t)n2 * 24);
  float FF[18];
  std::memcpy(FF, F21, 36);
  std::memcpy(FF + 9, F12, 36);
  PSL_UP(M[6], FF, sizeof(FF));
  A.kl = M[2].as<psl_keyline>();
  T.kl = M[3].as<psl_keyline>();
  PSL_ENS(ctx->m_best, (size_t)n1 * 8);
  PSL_ENS(ctx->m_cand_count, (size_t)n2 * 8);
  PSL_ENS(ctx->m_assign, (size_t)n1 * 4);
  PSL_ENS(ctx->m_accepted, (size_t)n2 * 4);
  PSL_ENS(ctx->m_nm, 4);
  cudaStream_t st = ctx->stream;
  size_t e = prof_mark(ctx);
  launch_line_knn2(A, T, ctx->m_best.as<uint2>(), 1, st);
  launch_line_bfmatch_new(A, T, M[5].as<double>(), ctx->m_best.as<uint2>(), M[6].as<float>(), nn_ratio, th,
                          ctx->m_assign.as<int32_t>(), 1, st);   // thread12: F21 (:805)
  launch_line_knn2(T, A, ctx->m_cand_count.as<uint2>(), 1, st);
  launch_line_bfmatch_new(T, A, M[4].as<double>(), ctx->m_cand_count.as<uint2>(), M[6].as<float>() + 9, nn_ratio, th,
                          ctx->m_accepted.as<int32_t>(), 1, st);  // thread21: F12 (:806)
  launch_line_triang(A, ctx->m_accepted.as<int32_t>(), T.cap, M[0].as<uint8_t>(), M[1].as<uint8_t>(), is_double,
                     ctx->m_assign.as<int32_t>(), ctx->m_nm.as<int32_t>(), 1, st);
  prof_span(ctx, 15, e, 5);
  PSL_CK(cudaGetLastError());
  PSL_CK(cudaMemcpyAsync(matched_pairs, ctx->m_assign.p, (size_t)n1 * 4, cudaMemcpyDeviceToHost, st));
  PSL_CK(cudaMemcpyAsync(nmatches, ctx->m_nm.p, 4, cudaMemcpyDeviceToHost, st));
  return check_status(ctx);
}

int psl_line_fuse(psl_ctx* ctx, const psl_keyline* kl, int32_t n_lines, const uint8_t* kf_desc, int32_t n_desc,
                  const psl_line_fuse_query* queries, const uint8_t* query_desc, int32_t nq, float th_cos, int32_t th_low,
                  int32_t* best_idx, int32_t* best_dist) {
  if (!ctx) return PSL_E_INVALID;
  if (nq < 0 || n_lines < 0 || n_lines > 65534 || n_desc < n_lines || (n_lines > 0 && (!kl || !kf_desc)) ||
      (nq > 0 && (!queries || !query_desc || !best_idx)))
    return fail(ctx, PSL_E_INVALID, "bad argument (kf_desc needs a row per line: LSDmatcher.cpp:938)");
  if (nq == 0) return PSL_OK;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  PSL_UP(ctx->m_misc[0], kl, (size_t)n_lines * sizeof(psl_keyline));
  PSL_UP(ctx->m_desc, kf_desc, (size_t)n_lines * 32);
  PSL_UP(ctx->m_misc[3], queries, (size_t)nq * sizeof(psl_line_fuse_query));
  PSL_UP(ctx->m_qdesc, query_desc, (size_t)nq * 32);
  PSL_ENS(ctx->m_assign, (size_t)nq * 4);
  PSL_ENS(ctx->m_accepted, (size_t)nq * 4);
  cudaStream_t st = ctx->stream;
  size_t e = prof_mark(ctx);
  launch_line_fuse(ctx->m_misc[0].as<psl_keyline>(), n_lines, ctx->m_desc.as<uint8_t>(),
                   ctx->m_misc[3].as<psl_line_fuse_query>(), ctx->m_qdesc.as<uint8_t>(), nq, th_cos, th_low,
                   ctx->m_assign.as<int32_t>(), ctx->m_accepted.as<int32_t>(), st);
  prof_span(ctx, 15, e, 1);
  PSL_CK(cudaGetLastError());
  PSL_CK(cudaMemcpyAsync(best_idx, ctx->m_assign.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
  if (best_dist) PSL_CK(cudaMemcpyAsync(best_dist, ctx->m_accepted.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
  return check_status(ctx);
}

int psl_line_match_projection(psl_ctx* ctx, const psl_line_frame_view* fv, const psl_line_query* queries,
                              const uint8_t* query_desc, int32_t nq, const uint8_t* claimed_in, int32_t mode,
                              float nn_ratio, int32_t* assign, int32_t* nmatches) {
  if (!ctx) return PSL_E_INVALID;
  if (!fv || !nmatches || nq < 0 || fv->n < 0 || fv->n > kMaxLinesPerFrame || (mode != 0 && mode != 1) ||
      (fv->n > 0 && (!fv->kl_un || !fv->ldesc || !fv->lineeq || !assign || (mode == 1 && !fv->lines3d))) ||
      (nq > 0 && (!queries || !query_desc)))
    return fail(ctx, PSL_E_INVALID, "bad argument (at most 4096 lines per frame)");
  *nmatches = 0;
  const int n = fv->n;
  if (n == 0) return PSL_OK;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  const int qcap = std::max(nq, 1);
  PSL_UP(ctx->m_misc[0], fv->kl_un, (size_t)n * sizeof(psl_keyline));
  PSL_UP(ctx->m_desc, fv->ldesc, (size_t)n * 32);
  PSL_UP(ctx->m_misc[1], fv->lineeq, (size_t)n * 24);
  if (mode == 1) PSL_UP(ctx->m_misc[2], fv->lines3d, (size_t)n * 48);
  PSL_UP(ctx->m_misc[3], queries, (size_t)nq * sizeof(psl_line_query));
  PSL_UP(ctx->m_qdesc, query_desc, (size_t)nq * 32);
  if (claimed_in) PSL_UP(ctx->m_claimed, claimed_in, (size_t)n);
  const int32_t nn[2] = {n, nq};
  PSL_UP(ctx->m_n, nn, sizeof(nn));
  PSL_ENS(ctx->m_misc[4], (size_t)n * kLineCells * 2);
  PSL_ENS(ctx->m_misc[5], (size_t)n);
  PSL_ENS(ctx->m_misc[6], (size_t)qcap * n * 8);
  PSL_ENS(ctx->m_misc[7], (size_t)n);
  PSL_ENS(ctx->m_assign, (size_t)n * 4);
  PSL_ENS(ctx->m_nm, 4);
  LineSet F{ctx->m_misc[0].as<psl_keyline>(), ctx->m_desc.as<uint8_t>(), ctx->m_n.as<int32_t>(), n};
  size_t e = prof_mark(ctx);
  launch_line_projection(F, ctx->m_misc[1].as<double>(), mode == 1 ? ctx->m_misc[2].as<double>() : nullptr,
                         ctx->m_misc[3].as<psl_line_query>(), ctx->m_qdesc.as<uint8_t>(), ctx->m_n.as<int32_t>() + 1, qcap,
                         nq, fv->min_x, fv->min_y, fv->grid_w_inv, fv->grid_h_inv, mode, nn_ratio,
                         claimed_in ? ctx->m_claimed.as<uint8_t>() : nullptr, ctx->m_misc[4].as<uint16_t>(),
                         ctx->m_misc[5].as<uint8_t>(), ctx->m_misc[6].as<unsigned long long>(), ctx->m_misc[7].as<uint8_t>(),
                         ctx->m_assign.as<int32_t>(), ctx->m_nm.as<int32_t>(), 1, ctx->stream);
  prof_span(ctx, 15, e, 3);
  PSL_CK(cudaMemcpyAsync(assign, ctx->m_assign.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  PSL_CK(cudaMemcpyAsync(nmatches, ctx->m_nm.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  return check_status(ctx);
}

int psl_plane_assoc(psl_ctx* ctx, const float* planes_cam, const double* pts, int32_t n_ljl, const float* Tcw,
                    const float* map_planes, const uint8_t* map_bad, int32_t n_map, float d_th, float a_th,
                    int32_t mode, int32_t* assign, int32_t* nmatches) {
  if (!ctx) return PSL_E_INVALID;
  if (!nmatches || n_ljl < 0 || n_map < 0 || (mode != 0 && mode != 1) || !Tcw ||
      (n_ljl > 0 && (!planes_cam || !pts || !assign)) || (n_map > 0 && !map_planes))
    return fail(ctx, PSL_E_INVALID, "bad argument");
  *nmatches = 0;
  if (n_ljl == 0) return PSL_OK;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  PSL_UP(ctx->m_misc[0], planes_cam, (size_t)n_ljl * 16);
  PSL_UP(ctx->m_misc[1], pts, (size_t)n_ljl * 120);
  PSL_UP(ctx->m_misc[2], Tcw, 64);
  PSL_UP(ctx->m_misc[3], map_planes, (size_t)n_map * 16);
  if (map_bad) PSL_UP(ctx->m_claimed, map_bad, (size_t)n_map);
  PSL_ENS(ctx->m_assign, (size_t)n_ljl * 4);
  PSL_ENS(ctx->m_nm, 4);
  size_t e = prof_mark(ctx);
  launch_plane_assoc(ctx->m_misc[0].as<float>(), ctx->m_misc[1].as<double>(), n_ljl, ctx->m_misc[2].as<float>(),
                     ctx->m_misc[3].as<float>(), map_bad ? ctx->m_claimed.as<uint8_t>() : nullptr, n_map, d_th, a_th, mode,
                     ctx->m_assign.as<int32_t>(), ctx->m_nm.as<int32_t>(), ctx->stream);
  prof_span(ctx, 15, e, 1);
  PSL_CK(cudaMemcpyAsync(assign, ctx->m_assign.p, (size_t)n_ljl * 4, cudaMemcpyDeviceToHost, ctx->stream));
  PSL_CK(cudaMemcpyAsync(nmatches, ctx->m_nm.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  return check_status(ctx);
}

int psl_plane_hypotheses(psl_ctx* ctx, const psl_keyline* kl_un, const float* line_eq, const double* lines3d,
                         int32_t n_lines, const psl_line_junction* junctions, int32_t n_junctions, double* le_l,
                         float* planes, double* normals, int32_t* junction_of, int32_t cap, int32_t* n_planes) {
  if (!ctx) return PSL_E_INVALID;
  if (!n_planes || n_lines < 0 || n_junctions < 0 || cap < 0 ||
      (n_junctions > 0 && (!kl_un || !line_eq || !lines3d || !junctions || !le_l || n_lines < 1)) ||
      (cap > 0 && (!planes || !normals || !junction_of)))
    return fail(ctx, PSL_E_INVALID, "bad argument");
  *n_planes = 0;
  if (n_junctions == 0) return PSL_OK;
  for (int i = 0; i < n_junctions; ++i)
    if (junctions[i].l1 < 0 || junctions[i].l1 >= n_lines || junctions[i].l2 < 0 || junctions[i].l2 >= n_lines)
      return fail(ctx, PSL_E_INVALID, "junction refers to a line outside [0, n_lines)");
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  const int ocap = std::max(cap, 1);
  PSL_UP(ctx->m_misc[0], kl_un, (size_t)n_lines * sizeof(psl_keyline));
  PSL_UP(ctx->m_misc[1], line_eq, (size_t)n_lines * 12);
  PSL_UP(ctx->m_misc[2], lines3d, (size_t)n_lines * 48);
  PSL_UP(ctx->m_misc[3], junctions, (size_t)n_junctions * sizeof(psl_line_junction));
  PSL_ENS(ctx->m_misc[4], (size_t)n_junctions * 48);   // le_l
  PSL_ENS(ctx->m_misc[5], (size_t)ocap * 16);          // planes
  PSL_ENS(ctx->m_misc[6], (size_t)ocap * 24);          // normals
  PSL_ENS(ctx->m_assign, (size_t)ocap * 4);            // junction_of
  PSL_ENS(ctx->m_misc[7], (size_t)n_junctions * 16);   // every kept plane (OldPlane reads them)
  PSL_ENS(ctx->m_nm, 4);
  cudaStream_t st = ctx->stream;
  size_t e = prof_mark(ctx);
  launch_plane_hypotheses(ctx->m_misc[0].as<psl_keyline>(), ctx->m_misc[1].as<float>(), ctx->m_misc[2].as<double>(),
                          ctx->m_misc[3].as<psl_line_junction>(), n_junctions, ctx->m_misc[4].as<double>(),
                          ctx->m_misc[5].as<float>(), ctx->m_misc[6].as<double>(), ctx->m_assign.as<int32_t>(), cap,
                          ctx->m_nm.as<int32_t>(), ctx->m_misc[7].as<float>(), st);
  prof_span(ctx, 15, e, 1);
  PSL_CK(cudaGetLastError());
  PSL_CK(cudaMemcpyAsync(le_l, ctx->m_misc[4].p, (size_t)n_junctions * 48, cudaMemcpyDeviceToHost, st));
  PSL_CK(cudaMemcpyAsync(n_planes, ctx->m_nm.p, 4, cudaMemcpyDeviceToHost, st));
  int rc = check_status(ctx);  // synchronises: *n_planes is valid now
  if (rc) return rc;
  const int np = std::min(*n_planes, cap);
  if (np > 0) {
    PSL_CK(cudaMemcpyAsync(planes, ctx->m_misc[5].p, (size_t)np * 16, cudaMemcpyDeviceToHost, st));
    PSL_CK(cudaMemcpyAsync(normals, ctx->m_misc[6].p, (size_t)np * 24, cudaMemcpyDeviceToHost, st));
    PSL_CK(cudaMemcpyAsync(junction_of, ctx->m_assign.p, (size_t)np * 4, cudaMemcpyDeviceToHost, st));
    PSL_CK(cudaStreamSynchronize(st));
  }
  if (*n_planes > cap) return fail(ctx, PSL_E_CAPACITY, "more plane hypotheses than `cap`");
  return PSL_OK;
}

int psl_lines_3d_dev(psl_ctx* ctx, const psl_keyline* d_kl, const int32_t* d_n, int32_t cap, int32_t B, const float* d_depth,
                     int32_t w, int32_t h, int32_t depth_stride_px, int64_t depth_frame_stride_px, float fx, float fy,
                     float cx, float cy, uint32_t seed, double* d_lines3d, float* d_line_eq) {
  if (!ctx) return PSL_E_INVALID;
  if (B < 0 || cap < 1 || w <= 0 || h <= 0 || depth_stride_px < w || !(fx != 0.f) || !(fy != 0.f) ||
      (B > 0 && (!d_kl || !d_n || !d_depth || !d_lines3d || !d_line_eq)))
    return fail(ctx, PSL_E_INVALID, "bad argument");
  if (B == 0) return PSL_OK;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  size_t e = prof_mark(ctx);
  launch_lines3d(d_kl, d_n, cap, d_depth, w, h, depth_stride_px, depth_frame_stride_px, fx, fy, cx, cy, seed, d_lines3d,
                 d_line_eq, B, ctx->stream);
  prof_span(ctx, 15, e, 1);
  PSL_CK(cudaGetLastError());
  return PSL_OK;
}

int psl_lines_3d(psl_ctx* ctx, const psl_keyline* kl_un, int32_t n, const float* depth, int32_t w, int32_t h, float fx,
                 float fy, float cx, float cy, uint32_t seed, double* lines3d, float* line_eq) {
  if (!ctx) return PSL_E_INVALID;
  if (n < 0 || w <= 0 || h <= 0 || !depth || (n > 0 && (!kl_un || !lines3d || !line_eq)))
    return fail(ctx, PSL_E_INVALID, "bad argument");
  if (n == 0) return PSL_OK;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  PSL_UP(ctx->m_misc[0], kl_un, (size_t)n * sizeof(psl_keyline));
  PSL_UP(ctx->m_misc[1], depth, (size_t)w * h * 4);
  PSL_UP(ctx->m_n, &n, 4);
  PSL_ENS(ctx->m_misc[2], (size_t)n * 48);
  PSL_ENS(ctx->m_misc[3], (size_t)n * 12);
  int rc = psl_lines_3d_dev(ctx, ctx->m_misc[0].as<psl_keyline>(), ctx->m_n.as<int32_t>(), n, 1, ctx->m_misc[1].as<float>(), w,
                            h, w, (int64_t)w * h, fx, fy, cx, cy, seed, ctx->m_misc[2].as<double>(),
                            ctx->m_misc[3].as<float>());
  if (rc) return rc;
  PSL_CK(cudaMemcpyAsync(lines3d, ctx->m_misc[2].p, (size_t)n * 48, cudaMemcpyDeviceToHost, ctx->stream));
  PSL_CK(cudaMemcpyAsync(line_eq, ctx->m_misc[3].p, (size_t)n * 12, cudaMemcpyDeviceToHost, ctx->stream));
  return check_status(ctx);
}

}  // extern "C"
