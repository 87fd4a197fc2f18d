// C-ABI of the batched point front end (psl_track_orb_batch[_dev]): ORB extraction, RGB-D stereo
// bookkeeping and SearchByProjection against the previous frame, chunk by chunk, all in HBM.
#include <algorithm>
#include <cstring>

#include "frame_kernels.cuh"
#include "line_match_kernels.cuh"
#include "match_kernels.cuh"
#include "psl_ctx.cuh"

using namespace psl;

namespace psl {
int track_after_extract(psl_ctx* ctx, const uint16_t* d_depth, int32_t depth_stride_px, int64_t depth_frame_stride_px,
                        int32_t B, int32_t w, int32_t h, const float* d_Tcw, const psl_camera* cam,
                        const psl_track_params* prm, psl_keypoint* d_kps, uint8_t* d_desc, int32_t* d_n, float* d_u_right,
                        float* d_z, int32_t* d_assign, int32_t* d_nmatches, int32_t cap);
}

extern "C" {

int psl_track_orb_batch_dev(psl_ctx* ctx, const uint8_t* d_gray, int32_t gray_stride, int64_t gray_frame_stride,
                            const uint16_t* d_depth, int32_t depth_stride_px, int64_t depth_frame_stride_px,
                            int32_t B, int32_t w, int32_t h, const float* d_Tcw, const psl_camera* cam,
                            const psl_track_params* prm, psl_keypoint* d_kps, uint8_t* d_desc, int32_t* d_n,
                            float* d_u_right, float* d_z, int32_t* d_assign, int32_t* d_nmatches, int32_t cap) {
  if (!ctx) return PSL_E_INVALID;
  if (!d_gray || !d_depth || !d_Tcw || !cam || !prm || !d_kps || !d_desc || !d_n || !d_u_right || !d_z || !d_assign ||
      !d_nmatches || B < 0 || w <= 0 || h <= 0 || cap < 1 || cap > 65535 || depth_stride_px < w)
    return fail(ctx, PSL_E_INVALID, "bad argument");
  if (B == 0) return PSL_OK;
  int rc = psl_orb_extract_batch_dev(ctx, d_gray, B, w, h, gray_stride, gray_frame_stride, d_kps, d_desc, cap, d_n);
  if (rc) return rc;
  return psl::track_after_extract(ctx, d_depth, depth_stride_px, depth_frame_stride_px, B, w, h, d_Tcw, cam, prm, d_kps,
                                  d_desc, d_n, d_u_right, d_z, d_assign, d_nmatches, cap);
}

}  // extern "C"

namespace psl {
// Everything of the batched point front end that follows the extraction: ComputeStereoFromRGBD and
// SearchByProjection against the previous frame (arguments as psl_track_orb_batch_dev, already validated).
int track_after_extract(psl_ctx* ctx, const uint16_t* d_depth, int32_t depth_stride_px, int64_t depth_frame_stride_px,
                        int32_t B, int32_t w, int32_t h, const float* d_Tcw, const psl_camera* cam,
                        const psl_track_params* prm, psl_keypoint* d_kps, uint8_t* d_desc, int32_t* d_n, float* d_u_right,
                        float* d_z, int32_t* d_assign, int32_t* d_nmatches, int32_t cap) {
  int rc;
  cudaStream_t st = ctx->stream;
  size_t e = prof_mark(ctx);
  launch_stereo(d_kps, d_n, cap, d_depth, depth_stride_px, depth_frame_stride_px, cam->depth_factor, cam->bf, d_u_right,
                d_z, B, st);
  prof_span(ctx, 6, e, 1);

  // The matching kernels are latency-bound per frame (one warp walks a frame's queries in order), so they
  // run over many more frames per launch than the L2-sized extraction chunks.
  const int C = std::min(B, 4096);   // frames per matching launch (the ordered resolve is one warp per frame: it wants them all)
  if ((rc = ensure(ctx, ctx->m_q, (size_t)C * cap * sizeof(psl_proj_query)))) return rc;
  if ((rc = ensure(ctx, ctx->m_n, (size_t)C * 4))) return rc;
  if ((rc = ensure(ctx, ctx->m_cell_start, (size_t)C * (kGridCells + 1) * 4))) return rc;
  if ((rc = ensure(ctx, ctx->m_cell_items, (size_t)C * cap * 2))) return rc;
  if ((rc = ensure(ctx, ctx->m_cand, (size_t)C * cap * kCandCap * 4))) return rc;
  if ((rc = ensure(ctx, ctx->m_cand_count, (size_t)C * cap * 4))) return rc;
  if ((rc = ensure(ctx, ctx->m_best, (size_t)C * cap * 8))) return rc;
  if ((rc = ensure(ctx, ctx->m_accepted, (size_t)C * cap * 4))) return rc;

  QueryBuildParams qp{};
  qp.th = prm->th;
  qp.mono = 0;
  qp.min_x = 0.f;  // zero distortion: mnMinX = 0, mnMaxX = cols (Frame.cc:155-158)
  qp.min_y = 0.f;
  qp.max_x = (float)w;
  qp.max_y = (float)h;
  for (int l = 0; l < ctx->cfg.orb_nlevels; ++l) qp.scale[l] = ctx->scale[l];
  const float gw = (float)PSL_GRID_COLS / (qp.max_x - qp.min_x), gh = (float)PSL_GRID_ROWS / (qp.max_y - qp.min_y);
  psl_match_params mp{0, prm->th_dist, prm->nn_ratio, prm->check_orientation};

  for (int c0 = 0; c0 < B; c0 += C) {
    const int nb = std::min(C, B - c0);
    const size_t off = (size_t)c0 * cap;
    e = prof_mark(ctx);
    launch_query_build(d_kps, d_z, d_n, cap, d_Tcw, c0, *cam, qp, ctx->m_q.as<psl_proj_query>(), ctx->m_n.as<int32_t>(),
                       nb, st);
    prof_span(ctx, 6, e, 1);
    MatchFrames F{d_kps + off, d_u_right + off, d_desc + off * 32, d_n + c0, cap, qp.min_x, qp.min_y, gw, gh};
    // descriptors of the query side are those of frame b-1 in the same [B][cap] array
    const uint8_t* qdesc = d_desc + ((ptrdiff_t)c0 - 1) * (ptrdiff_t)cap * 32;
    MatchQueries Q{ctx->m_q.as<psl_proj_query>(), qdesc, ctx->m_n.as<int32_t>(), cap};
    e = prof_mark(ctx);
    launch_grid_build(F, ctx->m_cell_start.as<int32_t>(), ctx->m_cell_items.as<uint16_t>(), nb, st);
    prof_span(ctx, 7, e, 1);
    e = prof_mark(ctx);
    launch_proj_candidates(F, Q, ctx->m_cell_start.as<int32_t>(), ctx->m_cell_items.as<uint16_t>(),
                           ctx->m_cand.as<uint32_t>(), ctx->m_cand_count.as<int32_t>(), ctx->m_best.as<uint2>(),
                         ctx->d_status, nb, st);
    prof_span(ctx, 8, e, 1);
    e = prof_mark(ctx);
    launch_proj_resolve(F, Q, ctx->m_cand.as<uint32_t>(), ctx->m_cand_count.as<int32_t>(), ctx->m_best.as<uint2>(),
                        nullptr, mp,
                        ctx->m_accepted.as<uint32_t>(), d_assign + off, d_nmatches + c0, nb, st);
    prof_span(ctx, 9, e, 1);
  }
  PSL_CK(cudaGetLastError());
  return PSL_OK;
}

}  // namespace psl

extern "C" {

int psl_track_orb_batch(psl_ctx* ctx, const uint8_t* gray, const uint16_t* depth, int32_t B, int32_t w, int32_t h,
                        const float* Tcw, const psl_camera* cam, const psl_track_params* prm, psl_keypoint* kps,
                        uint8_t* desc, int32_t* n, float* u_right, float* z, int32_t* assign, int32_t* nmatches,
                        int32_t cap) {
  if (!ctx) return PSL_E_INVALID;
  if (!gray || !depth || !Tcw || !kps || !desc || !n || !u_right || !z || !assign || !nmatches || B < 0 || w <= 0 ||
      h <= 0 || cap < 1)
    return fail(ctx, PSL_E_INVALID, "bad argument");
  if (B == 0) return PSL_OK;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  const size_t px = (size_t)w * h, nk = (size_t)B * cap;
  DevBuf* M = ctx->m_misc;
  int rc;
  if ((rc = ensure(ctx, M[0], px * B))) return rc;
  if ((rc = ensure(ctx, M[1], px * B * 2))) return rc;
  if ((rc = ensure(ctx, M[2], (size_t)B * 12 * 4))) return rc;
  if ((rc = ensure(ctx, M[3], nk * sizeof(psl_keypoint)))) return rc;
  if ((rc = ensure(ctx, M[4], nk * 32))) return rc;
  if ((rc = ensure(ctx, M[5], (size_t)B * 8))) return rc;  // n | nmatches
  if ((rc = ensure(ctx, M[6], nk * 4))) return rc;
  if ((rc = ensure(ctx, M[7], nk * 4))) return rc;
  if ((rc = ensure(ctx, M[8], nk * 4))) return rc;
  cudaStream_t st = ctx->stream;
  PSL_CK(cudaMemcpyAsync(M[0].p, gray, px * B, cudaMemcpyHostToDevice, st));
  PSL_CK(cudaMemcpyAsync(M[1].p, depth, px * B * 2, cudaMemcpyHostToDevice, st));
  PSL_CK(cudaMemcpyAsync(M[2].p, Tcw, (size_t)B * 48, cudaMemcpyHostToDevice, st));
  int32_t* d_n = M[5].as<int32_t>();
  rc = psl_track_orb_batch_dev(ctx, M[0].as<uint8_t>(), w, (int64_t)px, M[1].as<uint16_t>(), w, (int64_t)px, B, w, h,
                               M[2].as<float>(), cam, prm, M[3].as<psl_keypoint>(), M[4].as<uint8_t>(), d_n,
                               M[6].as<float>(), M[7].as<float>(), M[8].as<int32_t>(), d_n + B, cap);
  if (rc) return rc;
  PSL_CK(cudaMemcpyAsync(kps, M[3].p, nk * sizeof(psl_keypoint), cudaMemcpyDeviceToHost, st));
  PSL_CK(cudaMemcpyAsync(desc, M[4].p, nk * 32, cudaMemcpyDeviceToHost, st));
  PSL_CK(cudaMemcpyAsync(n, d_n, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  PSL_CK(cudaMemcpyAsync(nmatches, d_n + B, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  PSL_CK(cudaMemcpyAsync(u_right, M[6].p, nk * 4, cudaMemcpyDeviceToHost, st));
  PSL_CK(cudaMemcpyAsync(z, M[7].p, nk * 4, cudaMemcpyDeviceToHost, st));
  PSL_CK(cudaMemcpyAsync(assign, M[8].p, nk * 4, cudaMemcpyDeviceToHost, st));
  return check_status(ctx);
}


}  // extern "C"

namespace {
struct Copy { void* dst; const void* src; size_t bytes; };
// Host inputs of the host-pointer entry point (null members = data already in HBM)
struct HostFeed { const uint8_t* gray; const uint16_t* depth; const float* Tcw; };

// The combined front end.  With `feed` the inputs come from host memory: the gray frames go up slice by slice
// (one extraction chunk each) on a copy stream, the ORB extraction of a slice starts as soon as the slice has
// landed, the line path (which needs all gray frames and is the longer of the two) starts after the last slice,
// depth and poses follow behind the gray frames, and `early_out` (device->host copies of the point results)
// travel back while the line path is still running.
int frontend_core_impl(psl_ctx* ctx, uint8_t* d_gray, int32_t gray_stride, int64_t gray_frame_stride,
                       uint16_t* d_depth, int32_t depth_stride_px, int64_t depth_frame_stride_px, int32_t B,
                       int32_t w, int32_t h, float* d_Tcw, const psl_camera* cam, const psl_track_params* prm,
                       float line_desc_th, const psl_frontend_out* o, const HostFeed* feed, const Copy* early_out,
                       int n_out) {
  if (!ctx) return PSL_E_INVALID;
  if (!o || !o->kl || !o->ldesc || !o->lineeq || !o->nl || !o->line_assign || !o->line_nmatches || o->line_cap < 1 ||
      o->line_cap > kMaxLinesPerFrame)
    return fail(ctx, PSL_E_INVALID, "bad line output block");
  if (!d_gray || !d_depth || !d_Tcw || !cam || !prm || !o->kps || !o->desc || !o->n || !o->u_right || !o->z ||
      !o->assign || !o->nmatches || w <= 0 || h <= 0 || o->cap < 1 || o->cap > 65535 || depth_stride_px < w)
    return fail(ctx, PSL_E_INVALID, "bad argument");
  if (B <= 0) return B == 0 ? PSL_OK : fail(ctx, PSL_E_INVALID, "bad argument");
  // The line path (latency-bound: one warp walks one frame) and the point path (bandwidth / ALU-bound) are
  // independent until the matchers, so they run on two streams and share the SMs; with per-stage profiling on
  // they are serialised so that the stage timings stay meaningful.
  const int lc = o->line_cap;
  cudaStream_t main_st = ctx->stream;
  const bool overlap = !ctx->prof;
  int rc;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  const int slice = ctx->chunk;
  const int nslices = (B + slice - 1) / slice;
  const bool sliced = feed && overlap;
  if (feed && !sliced) {  // profiling: everything in order on one stream
    PSL_CK(cudaMemcpyAsync(d_gray, feed->gray, (size_t)gray_frame_stride * B, cudaMemcpyHostToDevice, main_st));
    PSL_CK(cudaMemcpyAsync(d_depth, feed->depth, (size_t)depth_frame_stride_px * 2 * B, cudaMemcpyHostToDevice, main_st));
    PSL_CK(cudaMemcpyAsync(d_Tcw, feed->Tcw, (size_t)B * 48, cudaMemcpyHostToDevice, main_st));
  }
  if (overlap) PSL_CK(cudaEventRecord(ctx->ev_fork, main_st));
  if (sliced) {
    cudaStream_t cs = ctx->stream_copy;
    while ((int)ctx->ev_slice.size() < nslices + 1) {
      cudaEvent_t ev;
      PSL_CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      ctx->ev_slice.push_back(ev);
    }
    PSL_CK(cudaStreamWaitEvent(cs, ctx->ev_fork, 0));  // earlier work on the main stream may still read the staging
    for (int s0 = 0; s0 < nslices; ++s0) {
      const int f0 = s0 * slice, nb = std::min(slice, B - f0);
      PSL_CK(cudaMemcpyAsync(d_gray + (size_t)f0 * gray_frame_stride, feed->gray + (size_t)f0 * gray_frame_stride,
                             (size_t)gray_frame_stride * nb, cudaMemcpyHostToDevice, cs));
      PSL_CK(cudaEventRecord(ctx->ev_slice[s0], cs));
    }
    PSL_CK(cudaMemcpyAsync(d_depth, feed->depth, (size_t)depth_frame_stride_px * 2 * B, cudaMemcpyHostToDevice, cs));
    PSL_CK(cudaMemcpyAsync(d_Tcw, feed->Tcw, (size_t)B * 48, cudaMemcpyHostToDevice, cs));
    PSL_CK(cudaEventRecord(ctx->ev_slice[nslices], cs));
  }
  if (overlap) {
    PSL_CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
    if (sliced) PSL_CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_slice[nslices - 1], 0));
    ctx->stream = ctx->stream2;
  }
  rc = psl_line_extract_batch_dev(ctx, d_gray, B, w, h, gray_stride, gray_frame_stride, o->kl, o->ldesc, o->lineeq, nullptr,
                                  lc, o->nl);
  ctx->stream = main_st;
  if (overlap && !rc) {
    PSL_CK(cudaEventRecord(ctx->ev_join, ctx->stream2));
  }
  if (rc) return rc;
  if (sliced) {
    for (int s0 = 0; s0 < nslices; ++s0) {
      const int f0 = s0 * slice, nb = std::min(slice, B - f0);
      PSL_CK(cudaStreamWaitEvent(main_st, ctx->ev_slice[s0], 0));
      rc = psl_orb_extract_batch_dev(ctx, d_gray + (size_t)f0 * gray_frame_stride, nb, w, h, gray_stride, gray_frame_stride,
                                     o->kps + (size_t)f0 * o->cap, o->desc + (size_t)f0 * o->cap * 32, o->cap, o->n + f0);
      if (rc) return rc;
    }
    PSL_CK(cudaStreamWaitEvent(main_st, ctx->ev_slice[nslices], 0));
  } else {
    rc = psl_orb_extract_batch_dev(ctx, d_gray, B, w, h, gray_stride, gray_frame_stride, o->kps, o->desc, o->cap, o->n);
    if (rc) return rc;
  }
  rc = track_after_extract(ctx, d_depth, depth_stride_px, depth_frame_stride_px, B, w, h, d_Tcw, cam, prm, o->kps, o->desc,
                           o->n, o->u_right, o->z, o->assign, o->nmatches, o->cap);
  if (rc) return rc;
  for (int i = 0; i < n_out; ++i)
    PSL_CK(cudaMemcpyAsync(early_out[i].dst, early_out[i].src, early_out[i].bytes, cudaMemcpyDeviceToHost, main_st));
  if (overlap) PSL_CK(cudaStreamWaitEvent(main_st, ctx->ev_join, 0));
  // SearchByGeomNApearance(frame b, frame b-1): the Last set is the same [B][line_cap] block shifted by one frame
  cudaStream_t st = ctx->stream;
  if ((rc = ensure(ctx, ctx->m_misc[9], (size_t)B * lc * 8))) return rc;
  if ((rc = ensure(ctx, ctx->m_misc[10], (size_t)B * 4))) return rc;
  int32_t* n_last = ctx->m_misc[10].as<int32_t>();
  PSL_CK(cudaMemsetAsync(n_last, 0, 4, st));
  if (B > 1) PSL_CK(cudaMemcpyAsync(n_last + 1, o->nl, (size_t)(B - 1) * 4, cudaMemcpyDeviceToDevice, st));
  LineSet Cur{o->kl, o->ldesc, o->nl, lc};
  LineSet Last{o->kl - lc, o->ldesc - (ptrdiff_t)lc * 32, n_last, lc};
  size_t e = prof_mark(ctx);
  launch_line_knn2(Last, Cur, ctx->m_misc[9].as<uint2>(), B, st);
  launch_line_geom(Last, nullptr, Cur, ctx->m_misc[9].as<uint2>(), line_desc_th, (float)w, (float)h, o->line_assign,
                   o->line_nmatches, B, st);
  prof_span(ctx, 15, e, 2);
  PSL_CK(cudaGetLastError());
  return PSL_OK;
}

// An error after the fork leaves work on the line and copy streams that the main stream never joined; the next call
// would reuse the staging those streams still read.  Drain them before reporting the error.
int frontend_core(psl_ctx* ctx, uint8_t* d_gray, int32_t gray_stride, int64_t gray_frame_stride, uint16_t* d_depth,
                  int32_t depth_stride_px, int64_t depth_frame_stride_px, int32_t B, int32_t w, int32_t h, float* d_Tcw,
                  const psl_camera* cam, const psl_track_params* prm, float line_desc_th, const psl_frontend_out* o,
                  const HostFeed* feed, const Copy* early_out, int n_out) {
  cudaStream_t main_st = ctx ? ctx->stream : nullptr;
  const int rc = frontend_core_impl(ctx, d_gray, gray_stride, gray_frame_stride, d_depth, depth_stride_px,
                                    depth_frame_stride_px, B, w, h, d_Tcw, cam, prm, line_desc_th, o, feed, early_out, n_out);
  if (rc && ctx) {
    ctx->stream = main_st;
    if (ctx->stream2) cudaStreamSynchronize(ctx->stream2);
    if (ctx->stream_copy) cudaStreamSynchronize(ctx->stream_copy);
    cudaStreamSynchronize(main_st);
  }
  return rc;
}
}  // namespace

extern "C" {

int psl_track_frontend_batch_dev(psl_ctx* ctx, const uint8_t* d_gray, int32_t gray_stride, int64_t gray_frame_stride,
                                 const uint16_t* d_depth, int32_t depth_stride_px, int64_t depth_frame_stride_px,
                                 int32_t B, int32_t w, int32_t h, const float* d_Tcw, const psl_camera* cam,
                                 const psl_track_params* prm, float line_desc_th, const psl_frontend_out* o) {
  return frontend_core(ctx, const_cast<uint8_t*>(d_gray), gray_stride, gray_frame_stride, const_cast<uint16_t*>(d_depth),
                       depth_stride_px, depth_frame_stride_px, B, w, h, const_cast<float*>(d_Tcw), cam, prm, line_desc_th, o,
                       nullptr, nullptr, 0);
}

}  // extern "C"

namespace {
// The combined front end with HOST outputs: results are staged in HBM and copied back; the point results travel while
// the line path is still running.  Inputs are either in HBM already (feed == nullptr) or come from `feed` (then
// d_gray / d_depth / d_Tcw are the staging buffers they are uploaded to).  Synchronises before returning.
int frontend_to_host(psl_ctx* ctx, uint8_t* d_gray, uint16_t* d_depth, float* d_Tcw, const HostFeed* feed, int32_t B,
                     int32_t w, int32_t h, const psl_camera* cam, const psl_track_params* prm, float line_desc_th,
                     const psl_frontend_out* o) {
  const size_t px = (size_t)w * h, nk = (size_t)B * o->cap, nl = (size_t)B * o->line_cap;
  DevBuf* M = ctx->m_misc;
  int rc;
  if ((rc = ensure(ctx, M[3], nk * sizeof(psl_keypoint)))) return rc;
  if ((rc = ensure(ctx, M[4], nk * 32))) return rc;
  if ((rc = ensure(ctx, M[5], (size_t)B * 16))) return rc;  // n | nmatches | nl | line_nmatches
  if ((rc = ensure(ctx, M[6], nk * 4))) return rc;
  if ((rc = ensure(ctx, M[7], nk * 4))) return rc;
  if ((rc = ensure(ctx, M[8], nk * 4))) return rc;
  if ((rc = ensure(ctx, ctx->l_kl, nl * sizeof(psl_keyline)))) return rc;
  if ((rc = ensure(ctx, ctx->l_desc, nl * 32))) return rc;
  if ((rc = ensure(ctx, ctx->l_eq, nl * 24))) return rc;
  if ((rc = ensure(ctx, ctx->l_n, nl * 4))) return rc;  // line_assign
  cudaStream_t st = ctx->stream;
  int32_t* d_n = M[5].as<int32_t>();
  psl_frontend_out d = *o;
  d.kps = M[3].as<psl_keypoint>(); d.desc = M[4].as<uint8_t>(); d.n = d_n; d.nmatches = d_n + B;
  d.u_right = M[6].as<float>(); d.z = M[7].as<float>(); d.assign = M[8].as<int32_t>();
  d.kl = ctx->l_kl.as<psl_keyline>(); d.ldesc = ctx->l_desc.as<uint8_t>(); d.lineeq = ctx->l_eq.as<double>();
  d.nl = d_n + 2 * B; d.line_nmatches = d_n + 3 * B; d.line_assign = ctx->l_n.as<int32_t>();
  const Copy early_out[7] = {{o->kps, d.kps, nk * sizeof(psl_keypoint)}, {o->desc, d.desc, nk * 32},
                             {o->n, d.n, (size_t)B * 4},  {o->nmatches, d.nmatches, (size_t)B * 4},
                             {o->u_right, d.u_right, nk * 4}, {o->z, d.z, nk * 4}, {o->assign, d.assign, nk * 4}};
  rc = frontend_core(ctx, d_gray, w, (int64_t)px, d_depth, w, (int64_t)px, B, w, h, d_Tcw, cam, prm, line_desc_th, &d, feed,
                     early_out, 7);
  if (rc) return rc;
  PSL_CK(cudaMemcpyAsync(o->kl, d.kl, nl * sizeof(psl_keyline), cudaMemcpyDeviceToHost, st));
  PSL_CK(cudaMemcpyAsync(o->ldesc, d.ldesc, nl * 32, cudaMemcpyDeviceToHost, st));
  PSL_CK(cudaMemcpyAsync(o->lineeq, d.lineeq, nl * 24, cudaMemcpyDeviceToHost, st));
  PSL_CK(cudaMemcpyAsync(o->nl, d.nl, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  PSL_CK(cudaMemcpyAsync(o->line_nmatches, d.line_nmatches, (size_t)B * 4, cudaMemcpyDeviceToHost, st));
  PSL_CK(cudaMemcpyAsync(o->line_assign, d.line_assign, nl * 4, cudaMemcpyDeviceToHost, st));
  return check_status(ctx);
}

bool bad_host_out(const psl_frontend_out* o) {
  return !o || !o->kps || !o->desc || !o->n || !o->u_right || !o->z || !o->assign || !o->nmatches || !o->kl || !o->ldesc ||
         !o->lineeq || !o->nl || !o->line_assign || !o->line_nmatches || o->cap < 1 || o->line_cap < 1;
}
}  // namespace

extern "C" {

int psl_track_frontend_batch(psl_ctx* ctx, const uint8_t* gray, const uint16_t* depth, int32_t B, int32_t w, int32_t h,
                             const float* Tcw, const psl_camera* cam, const psl_track_params* prm, float line_desc_th,
                             const psl_frontend_out* o) {
  if (!ctx) return PSL_E_INVALID;
  if (!gray || !depth || !Tcw || bad_host_out(o) || B < 0 || w <= 0 || h <= 0) return fail(ctx, PSL_E_INVALID, "bad argument");
  if (B == 0) return PSL_OK;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  const size_t px = (size_t)w * h;
  DevBuf* M = ctx->m_misc;
  int rc;
  if ((rc = ensure(ctx, M[0], px * B))) return rc;
  if ((rc = ensure(ctx, M[1], px * B * 2))) return rc;
  if ((rc = ensure(ctx, M[2], (size_t)B * 12 * 4))) return rc;
  const HostFeed feed{gray, depth, Tcw};
  return frontend_to_host(ctx, M[0].as<uint8_t>(), M[1].as<uint16_t>(), M[2].as<float>(), &feed, B, w, h, cam, prm,
                          line_desc_th, o);
}

// ---- Tracking::GrabImageRGBD for a batch (src/Tracking.cc:214-243): colour -> gray on the device, then the combined
// front end.  The host-pointer form is split in two so that a caller can keep the PCIe link busy: _begin only enqueues the
// uploads of a batch into one of two staging sets, _end runs the oldest begun batch and returns its results; begin(k+1)
// before end(k) overlaps the upload of batch k+1 with the kernels of batch k.
int psl_track_rgbd_batch_dev(psl_ctx* ctx, const uint8_t* d_color, int32_t channels, int32_t rgb_order, int32_t color_stride,
                             int64_t color_frame_stride, const uint16_t* d_depth, int32_t depth_stride_px,
                             int64_t depth_frame_stride_px, int32_t B, int32_t w, int32_t h, const float* d_Tcw,
                             const psl_camera* cam, const psl_track_params* prm, float line_desc_th,
                             const psl_frontend_out* out) {
  if (!ctx) return PSL_E_INVALID;
  if (!d_color || (channels != 3 && channels != 4) || B < 0 || w <= 0 || h <= 0 || color_stride < w * channels)
    return fail(ctx, PSL_E_INVALID, "bad argument (3 or 4 channels)");
  if (B == 0) return PSL_OK;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  int rc;
  const int gp = (w + 15) & ~15;
  if ((rc = ensure(ctx, ctx->feed_gray, (size_t)gp * h * B))) return rc;
  size_t e = prof_mark(ctx);
  launch_color_to_gray(d_color, channels, rgb_order, color_stride, color_frame_stride, ctx->feed_gray.as<uint8_t>(), gp,
                       (int64_t)gp * h, B, w, h, ctx->stream);
  prof_span(ctx, 6, e, 1);
  return frontend_core(ctx, ctx->feed_gray.as<uint8_t>(), gp, (int64_t)gp * h, const_cast<uint16_t*>(d_depth), depth_stride_px,
                       depth_frame_stride_px, B, w, h, const_cast<float*>(d_Tcw), cam, prm, line_desc_th, out, nullptr,
                       nullptr, 0);
}

int psl_track_rgbd_batch_begin(psl_ctx* ctx, const uint8_t* color, int32_t channels, int32_t rgb_order,
                               const uint16_t* depth, int32_t B, int32_t w, int32_t h, const float* Tcw) {
  if (!ctx) return PSL_E_INVALID;
  if (!color || !depth || !Tcw || (channels != 3 && channels != 4) || B < 1 || w <= 0 || h <= 0)
    return fail(ctx, PSL_E_INVALID, "bad argument (3 or 4 channels)");
  psl_ctx::FeedSlot& s = ctx->feed[ctx->feed_head];
  if (s.pending) return fail(ctx, PSL_E_INVALID, "two batches are already in flight: call psl_track_rgbd_batch_end first");
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  const size_t px = (size_t)w * h;
  int rc;
  if ((rc = ensure(ctx, s.color, px * channels * B))) return rc;
  if ((rc = ensure(ctx, s.depth, px * 2 * B))) return rc;
  if ((rc = ensure(ctx, s.Tcw, (size_t)B * 48))) return rc;
  if (!s.up_done) {
    PSL_CK(cudaEventCreateWithFlags(&s.up_done, cudaEventDisableTiming));
    PSL_CK(cudaEventCreateWithFlags(&s.free_ev, cudaEventDisableTiming));
  } else {
    PSL_CK(cudaStreamWaitEvent(ctx->stream_copy, s.free_ev, 0));   // the batch that used this set has been computed
  }
  cudaStream_t cs = ctx->stream_copy;
  PSL_CK(cudaMemcpyAsync(s.color.p, color, px * channels * B, cudaMemcpyHostToDevice, cs));
  PSL_CK(cudaMemcpyAsync(s.depth.p, depth, px * 2 * B, cudaMemcpyHostToDevice, cs));
  PSL_CK(cudaMemcpyAsync(s.Tcw.p, Tcw, (size_t)B * 48, cudaMemcpyHostToDevice, cs));
  PSL_CK(cudaEventRecord(s.up_done, cs));
  s.B = B; s.w = w; s.h = h; s.channels = channels; s.rgb_order = rgb_order;
  s.pending = true;
  ctx->feed_head ^= 1;
  return PSL_OK;
}

int psl_track_rgbd_batch_end(psl_ctx* ctx, const psl_camera* cam, const psl_track_params* prm, float line_desc_th,
                             const psl_frontend_out* o) {
  if (!ctx) return PSL_E_INVALID;
  psl_ctx::FeedSlot& s = ctx->feed[ctx->feed_tail];
  if (!s.pending) return fail(ctx, PSL_E_INVALID, "no batch in flight: call psl_track_rgbd_batch_begin first");
  if (!cam || !prm || bad_host_out(o)) return fail(ctx, PSL_E_INVALID, "bad argument");
  // frontend_to_host works on tightly packed device frames; the gray staging is written that way when w is a multiple of 16
  if (s.w & 15) return fail(ctx, PSL_E_INVALID, "psl_track_rgbd_batch: width must be a multiple of 16");
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  s.pending = false;
  ctx->feed_tail ^= 1;
  int rc;
  const int gp = (s.w + 15) & ~15;
  if ((rc = ensure(ctx, ctx->feed_gray, (size_t)gp * s.h * s.B))) return rc;
  PSL_CK(cudaStreamWaitEvent(ctx->stream, s.up_done, 0));
  size_t e = prof_mark(ctx);
  launch_color_to_gray(s.color.as<uint8_t>(), s.channels, s.rgb_order, s.w * s.channels, (int64_t)s.w * s.h * s.channels,
                       ctx->feed_gray.as<uint8_t>(), gp, (int64_t)gp * s.h, s.B, s.w, s.h, ctx->stream);
  prof_span(ctx, 6, e, 1);
  rc = frontend_to_host(ctx, ctx->feed_gray.as<uint8_t>(), s.depth.as<uint16_t>(), s.Tcw.as<float>(), nullptr, s.B, s.w, s.h,
                        cam, prm, line_desc_th, o);
  // the set may be overwritten once everything enqueued so far has run (frontend_to_host synchronised on success)
  cudaEventRecord(s.free_ev, ctx->stream);
  return rc;
}

int psl_track_rgbd_batch(psl_ctx* ctx, const uint8_t* color, int32_t channels, int32_t rgb_order, const uint16_t* depth,
                         int32_t B, int32_t w, int32_t h, const float* Tcw, const psl_camera* cam,
                         const psl_track_params* prm, float line_desc_th, const psl_frontend_out* out) {
  if (!ctx) return PSL_E_INVALID;
  if (B == 0) return PSL_OK;
  if (ctx->feed[0].pending || ctx->feed[1].pending)
    return fail(ctx, PSL_E_INVALID, "a begun batch is in flight: finish it with psl_track_rgbd_batch_end");
  const int rc = psl_track_rgbd_batch_begin(ctx, color, channels, rgb_order, depth, B, w, h, Tcw);
  if (rc) return rc;
  return psl_track_rgbd_batch_end(ctx, cam, prm, line_desc_th, out);
}


// TrackWithMotionModel after the matcher (src/Tracking.cc:1193-1240): the matched keypoints of every frame, with the
// MapPoints their partners in the previous frame created (Frame::UnprojectStereo), go through
// Optimizer::PoseOptimization from the given prior pose; mvbOutlier and the inlier count come back.
int psl_track_pose_batch_dev(psl_ctx* ctx, const psl_keypoint* d_kps, const float* d_u_right, const float* d_z,
                             const int32_t* d_assign, const int32_t* d_n, int32_t cap, int32_t B, const float* d_Tcw,
                             const psl_camera* cam, float* d_Tcw_out, uint8_t* d_outlier, int32_t* d_n_inliers) {
  if (!ctx) return PSL_E_INVALID;
  if (B < 0 || cap < 1 || !cam || (B > 0 && (!d_kps || !d_u_right || !d_z || !d_assign || !d_n || !d_Tcw || !d_Tcw_out ||
      !d_outlier || !d_n_inliers)))
    return fail(ctx, PSL_E_INVALID, "bad argument");
  if (B == 0) return PSL_OK;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  int rc;
  DevBuf* M = ctx->m_misc;
  if ((rc = ensure(ctx, M[14], (size_t)B * cap * sizeof(psl_pose_point)))) return rc;
  if ((rc = ensure(ctx, M[15], (size_t)B * 64 + kMaxLevels * 4))) return rc;
  float* d_T44 = M[15].as<float>();
  float* d_is2 = d_T44 + (size_t)B * 16;
  PSL_CK(cudaMemcpyAsync(d_is2, ctx->inv_sigma2.data(), (size_t)ctx->cfg.orb_nlevels * 4, cudaMemcpyHostToDevice, ctx->stream));
  size_t e = prof_mark(ctx);
  launch_pose_points(d_kps, d_u_right, d_z, d_assign, d_n, cap, d_Tcw, *cam, d_is2, M[14].as<psl_pose_point>(), d_T44, B,
                     ctx->stream);
  prof_span(ctx, 6, e, 1);
  return psl_pose_optimization_dev(ctx, d_T44, M[14].as<psl_pose_point>(), d_n, cap, B, cam->fx, cam->fy, cam->cx, cam->cy,
                                   cam->bf, d_Tcw_out, d_outlier, d_n_inliers);
}

int psl_convert_rgbd_dev(psl_ctx* ctx, const uint8_t* d_color, int32_t channels, int32_t rgb_order, int32_t color_stride,
                         int64_t color_frame_stride, uint8_t* d_gray, int32_t gray_stride, int64_t gray_frame_stride,
                         const uint16_t* d_depth_in, int32_t depth_stride_px, int64_t depth_frame_stride_px,
                         float depth_factor, float* d_depth_out, int32_t B, int32_t w, int32_t h) {
  if (!ctx) return PSL_E_INVALID;
  const bool do_color = d_color || d_gray, do_depth = d_depth_in || d_depth_out;
  if (B < 0 || w <= 0 || h <= 0 || (do_color && (!d_color || !d_gray || (channels != 3 && channels != 4) ||
      color_stride < w * channels || gray_stride < w)) || (do_depth && (!d_depth_in || !d_depth_out || depth_stride_px < w)))
    return fail(ctx, PSL_E_INVALID, "bad argument (3 or 4 channels)");
  if (B == 0) return PSL_OK;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  size_t e = prof_mark(ctx);
  if (do_color)
    launch_color_to_gray(d_color, channels, rgb_order, color_stride, color_frame_stride, d_gray, gray_stride,
                         gray_frame_stride, B, w, h, ctx->stream);
  if (do_depth)
    launch_depth_to_float(d_depth_in, depth_stride_px, depth_frame_stride_px, depth_factor, d_depth_out, B, w, h,
                          ctx->stream);
  prof_span(ctx, 6, e, (int)do_color + (int)do_depth);
  PSL_CK(cudaGetLastError());
  return PSL_OK;
}

int psl_convert_rgbd(psl_ctx* ctx, const uint8_t* color, int32_t channels, int32_t rgb_order, uint8_t* gray,
                     const uint16_t* depth_in, float depth_factor, float* depth_out, int32_t B, int32_t w, int32_t h) {
  if (!ctx) return PSL_E_INVALID;
  if (B < 0 || w <= 0 || h <= 0 || (channels != 3 && channels != 4)) return fail(ctx, PSL_E_INVALID, "bad argument");
  if (B == 0) return PSL_OK;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  const size_t px = (size_t)w * h * B;
  DevBuf* M = ctx->m_misc;
  cudaStream_t st = ctx->stream;
  int rc;
  const bool do_color = color && gray, do_depth = depth_in && depth_out;
  if ((color || gray) && !do_color) return fail(ctx, PSL_E_INVALID, "color and gray must be given together");
  if ((depth_in || depth_out) && !do_depth) return fail(ctx, PSL_E_INVALID, "depth_in and depth_out must be given together");
  if (do_color) {
    if ((rc = ensure(ctx, M[0], px * channels))) return rc;
    if ((rc = ensure(ctx, M[1], px))) return rc;
    PSL_CK(cudaMemcpyAsync(M[0].p, color, px * channels, cudaMemcpyHostToDevice, st));
  }
  if (do_depth) {
    if ((rc = ensure(ctx, M[2], px * 2))) return rc;
    if ((rc = ensure(ctx, M[3], px * 4))) return rc;
    PSL_CK(cudaMemcpyAsync(M[2].p, depth_in, px * 2, cudaMemcpyHostToDevice, st));
  }
  rc = psl_convert_rgbd_dev(ctx, do_color ? M[0].as<uint8_t>() : nullptr, channels, rgb_order, w * channels,
                            (int64_t)w * h * channels, do_color ? M[1].as<uint8_t>() : nullptr, w, (int64_t)w * h,
                            do_depth ? M[2].as<uint16_t>() : nullptr, w, (int64_t)w * h, depth_factor,
                            do_depth ? M[3].as<float>() : nullptr, B, w, h);
  if (rc) return rc;
  if (do_color) PSL_CK(cudaMemcpyAsync(gray, M[1].p, px, cudaMemcpyDeviceToHost, st));
  if (do_depth) PSL_CK(cudaMemcpyAsync(depth_out, M[3].p, px * 4, cudaMemcpyDeviceToHost, st));
  return check_status(ctx);
}

int psl_undistort_keypoints_dev(psl_ctx* ctx, const psl_keypoint* d_kps, const int32_t* d_n, int32_t cap, int32_t B,
                                const psl_distortion* cam, psl_keypoint* d_kps_un) {
  if (!ctx) return PSL_E_INVALID;
  if (B < 0 || cap < 1 || !cam || (B > 0 && (!d_kps || !d_n || !d_kps_un))) return fail(ctx, PSL_E_INVALID, "bad argument");
  if (B == 0) return PSL_OK;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  size_t e = prof_mark(ctx);
  launch_undistort(d_kps, d_n, cap, *cam, d_kps_un, B, ctx->stream);
  prof_span(ctx, 6, e, 1);
  PSL_CK(cudaGetLastError());
  return PSL_OK;
}

int psl_undistort_keypoints(psl_ctx* ctx, const psl_keypoint* kps, int32_t n, const psl_distortion* cam,
                            psl_keypoint* kps_un) {
  if (!ctx) return PSL_E_INVALID;
  if (n < 0 || !cam || (n > 0 && (!kps || !kps_un))) return fail(ctx, PSL_E_INVALID, "bad argument");
  if (n == 0) return PSL_OK;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  DevBuf* M = ctx->m_misc;
  int rc;
  if ((rc = ensure(ctx, M[0], (size_t)n * sizeof(psl_keypoint)))) return rc;
  if ((rc = ensure(ctx, M[1], (size_t)n * sizeof(psl_keypoint)))) return rc;
  if ((rc = ensure(ctx, M[2], 4))) return rc;
  cudaStream_t st = ctx->stream;
  PSL_CK(cudaMemcpyAsync(M[0].p, kps, (size_t)n * sizeof(psl_keypoint), cudaMemcpyHostToDevice, st));
  PSL_CK(cudaMemcpyAsync(M[2].p, &n, 4, cudaMemcpyHostToDevice, st));
  rc = psl_undistort_keypoints_dev(ctx, M[0].as<psl_keypoint>(), M[2].as<int32_t>(), n, 1, cam, M[1].as<psl_keypoint>());
  if (rc) return rc;
  PSL_CK(cudaMemcpyAsync(kps_un, M[1].p, (size_t)n * sizeof(psl_keypoint), cudaMemcpyDeviceToHost, st));
  return check_status(ctx);
}

int psl_image_bounds(psl_ctx* ctx, int32_t cols, int32_t rows, const psl_distortion* cam, float* bounds) {
  if (!ctx) return PSL_E_INVALID;
  if (!cam || !bounds || cols <= 0 || rows <= 0) return fail(ctx, PSL_E_INVALID, "bad argument");
  if (cam->k1 == 0.f) {  // Frame.cc:1156-1162
    bounds[0] = 0.f; bounds[1] = 0.f; bounds[2] = (float)cols; bounds[3] = (float)rows;
    return PSL_OK;
  }
  psl_keypoint c[4] = {};  // :1139-1143
  c[1].x = (float)cols;
  c[2].y = (float)rows;
  c[3].x = (float)cols; c[3].y = (float)rows;
  const int rc = psl_undistort_keypoints(ctx, c, 4, cam, c);
  if (rc) return rc;
  bounds[0] = std::min(c[0].x, c[2].x);  // mnMinX :1150
  bounds[2] = std::max(c[1].x, c[3].x);  // mnMaxX
  bounds[1] = std::min(c[0].y, c[1].y);  // mnMinY
  bounds[3] = std::max(c[2].y, c[3].y);  // mnMaxY
  return PSL_OK;
}

}  // extern "C"
