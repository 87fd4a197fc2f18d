// C-ABI of the line extractor (LINEextractor::operator(), add_src/LineExtractor.cpp:325-366): per-size
// geometry and buffers of the line path, the chunk loop and the host-pointer staging.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "psl_ctx.cuh"

namespace psl {

void free_line_geometry(psl_ctx* c) {
  for (void* p : c->line_allocs) cudaFree(p);
  c->line_allocs.clear();
  c->lgeo_w = c->lgeo_h = 0;
}

template <class T>
static bool lalloc(psl_ctx* c, T*& p, size_t count, bool zero = false) {
  void* q = nullptr;
  if (cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)) != cudaSuccess) return false;
  c->line_allocs.push_back(q);
  if (zero && cudaMemsetAsync(q, 0, std::max<size_t>(count, 1) * sizeof(T), c->stream) != cudaSuccess) return false;
  p = reinterpret_cast<T*>(q);
  return true;
}

// one axis of cv::resize(fx = 0.8, INTER_LINEAR_EXACT) on CV_8U: source index + Q8 weight of the right tap (-1: single tap)
static void exact_axis(int sn, int dn, std::vector<short2>& t) {
  t.resize(dn);
  const double scale = 1.0 / 0.8;
  for (int d = 0; d < dn; ++d) {
    const double f = scale * (d + 0.5) - 0.5;
    const int i = (int)floor(f);
    short ofs = 0, c1 = -1;
    if (i >= 0 && sn > 1) {
      if (i < sn - 1) { ofs = (short)i; c1 = (short)lrint((f - i) * 256.0); }
      else ofs = (short)(sn - 1);
    }
    t[d] = make_short2(ofs, c1);
  }
}

static int set_line_geometry(psl_ctx* ctx, int w, int h) {
  if (ctx->lgeo_w == w && ctx->lgeo_h == h) return PSL_OK;
  if (w > ctx->cfg.max_width || h > ctx->cfg.max_height)
    return fail(ctx, PSL_E_CAPACITY, "frame larger than psl_config.max_width/max_height");
  if (w < 8 || h < 8 || w > 16384 || h > 16384) return fail(ctx, PSL_E_INVALID, "frame size not supported by the line path");
  PSL_CK(cudaStreamSynchronize(ctx->stream));
  free_line_geometry(ctx);
  LineBuffers& L = ctx->lb;
  std::memset(&L, 0, sizeof(L));
  L.w = w;
  L.h = h;
  L.pitch = (w + 127) & ~127;
  L.Ws = (int)lrint(w * 0.8);
  L.Hs = (int)lrint(h * 0.8);
  const double logNT = 5 * (log10((double)L.Ws) + log10((double)L.Hs)) / 2 + log10(11.0);
  L.min_reg_size = (int)(size_t)(-logNT / log10(22.5 / 180));
  L.raw_cap = ctx->raw_cap;
  if (L.raw_cap > 65535) return fail(ctx, PSL_E_INVALID, "line_max_raw must be <= 65535");
  const size_t C = ctx->line_chunk, npx = (size_t)L.Ws * L.Hs, R = L.raw_cap;
  std::vector<short2> xt, yt;
  exact_axis(w, L.Ws, xt);
  exact_axis(h, L.Hs, yt);
  short2 *dx = nullptr, *dy = nullptr;
  uint8_t* lut = nullptr;
  uint8_t* seed_lut = nullptr;
  // the x taps once more, per group of 4 outputs, for the word-based resize kernel
  const int n4 = (L.Ws + 3) >> 2;
  std::vector<short4> xt4(xt.size());
  for (size_t i = 0; i < xt.size(); ++i)
    xt4[i] = xt[i].y >= 0 ? make_short4(xt[i].x, (short)(xt[i].x + 1), (short)(256 - xt[i].y), xt[i].y)
                          : make_short4(xt[i].x, xt[i].x, 256, 0);
  std::vector<uint4> xw4(n4);
  std::vector<uint32_t> xo4(n4);
  const bool grouped = resize_group_tables(xt4.data(), L.Ws, xw4.data(), xo4.data());
  uint4* dxw = nullptr;
  uint32_t* dxo = nullptr;
  // the records of the row above a frame must read "not available" (lsd_kernels.cu, load_nbr): for every frame but the
  // first that row is the previous frame's last one (NOTDEF); the first frame gets Ws + 1 such records in front
  const size_t pix_pad = (2 * ((size_t)L.Ws + 1) + 7) & ~(size_t)7;   // whole 128-byte lines, so the records stay line-aligned;
  // one row + 1 of unavailable records before the first frame, twice that (and as much after the last frame) so that
  // region growing's look-ahead prefetch stays inside the allocation
  bool ok = lalloc(ctx, dx, xt.size()) && lalloc(ctx, dy, yt.size()) && lalloc(ctx, dxw, (size_t)n4) && lalloc(ctx, dxo, (size_t)n4) && lalloc(ctx, lut, lsd_lut_bytes()) && lalloc(ctx, seed_lut, lsd_seed_lut_bytes()) && lalloc(ctx, L.blur, C * L.pitch * h) &&
            lalloc(ctx, L.scaled, C * npx) && lalloc(ctx, L.pix, C * npx + 2 * pix_pad) &&
            lalloc(ctx, L.reg, C * npx) && lalloc(ctx, L.max_n2, C) &&
            lalloc(ctx, L.row_cnt, C * L.Hs) && lalloc(ctx, L.n_def, C) && lalloc(ctx, L.key_in, C * npx) &&
            lalloc(ctx, L.val_in, C * npx) && lalloc(ctx, L.val_out, C * npx) &&
            lalloc(ctx, L.raw, C * R * 4) && lalloc(ctx, L.n_raw, C) && lalloc(ctx, L.t1, C * R) &&
            lalloc(ctx, L.t2, C * R) && lalloc(ctx, L.m_angles, C * R) && lalloc(ctx, L.m_length, C * R) && lalloc(ctx, L.m_sangles, C * R) &&
            lalloc(ctx, L.m_order, C * R) && lalloc(ctx, L.m_tmp16, C * R) &&
            lalloc(ctx, L.m_nb, C * R * line::kNbCap) && lalloc(ctx, L.m_fw, C * R * line::kNbCap) && lalloc(ctx, L.m_scan, C * R) && lalloc(ctx, L.m_cnt, C * 2) && lalloc(ctx, L.m_nb_cnt, C * R) &&
            lalloc(ctx, L.m_code, C * R) && lalloc(ctx, L.m_check, C * R) && lalloc(ctx, L.m_loc, C * R) &&
            lalloc(ctx, L.m_flag, C * R, true) && lalloc(ctx, L.gxy, C * (size_t)w * h);
  if (!ok) {
    free_line_geometry(ctx);
    return fail(ctx, PSL_E_CUDA, "line buffers: out of device memory (lower psl_config.line_chunk_frames)");
  }
  PSL_CK(cudaMemsetAsync(L.pix, 0xFF, pix_pad * sizeof(float4), ctx->stream));
  L.pix += pix_pad;
  L.lut = reinterpret_cast<const float4*>(lut);
  L.seed_lut = reinterpret_cast<const float2*>(seed_lut);
  launch_lsd_lut(reinterpret_cast<float4*>(lut), reinterpret_cast<float2*>(seed_lut), ctx->stream);
  L.xtab = dx;
  L.ytab = dy;
  L.xw4 = grouped ? dxw : nullptr;
  L.xo4 = grouped ? dxo : nullptr;
  PSL_CK(cudaMemcpyAsync(dxw, xw4.data(), xw4.size() * sizeof(uint4), cudaMemcpyHostToDevice, ctx->stream));
  PSL_CK(cudaMemcpyAsync(dxo, xo4.data(), xo4.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  PSL_CK(cudaMemcpyAsync(dx, xt.data(), xt.size() * sizeof(short2), cudaMemcpyHostToDevice, ctx->stream));
  PSL_CK(cudaMemcpyAsync(dy, yt.data(), yt.size() * sizeof(short2), cudaMemcpyHostToDevice, ctx->stream));
  upload_lbd_tables();
  PSL_CK(cudaStreamSynchronize(ctx->stream));  // xt / yt are host temporaries
  PSL_CK(cudaGetLastError());
  ctx->lgeo_w = w;
  ctx->lgeo_h = h;
  return PSL_OK;
}

static int run_line_chunk(psl_ctx* ctx, ImgBatch in, int nb, psl_keyline* d_kl, uint8_t* d_ldesc, double* d_lineeq,
                          float* d_lbd72, int cap, int32_t* d_n) {
  const LineBuffers& L = ctx->lb;
  cudaStream_t st = ctx->stream;
  const int nfeat = ctx->cfg.line_nfeatures;
  size_t e = prof_mark(ctx);
  launch_lsd_prologue(L, in, nb, st);
  prof_span(ctx, 10, e, 5 + psl::kBlurLaunches);
  e = prof_mark(ctx);
  launch_lsd_order(L, nb, st);
  prof_span(ctx, 11, e, 1);
  e = prof_mark(ctx);
  launch_lsd_core(L, nb, ctx->d_status, st);
  prof_span(ctx, 12, e, 1);
  e = prof_mark(ctx);
  launch_line_post(L, nb, nfeat, d_kl, d_lineeq, cap, d_n, ctx->d_status, st);
  prof_span(ctx, 13, e, psl::kLinePostLaunches);
  e = prof_mark(ctx);
  launch_lbd(L, in, nb, nfeat, d_kl, d_n, cap, d_ldesc, d_lbd72, st);
  prof_span(ctx, 14, e, 2 + psl::kBlurLaunches);
  PSL_CK(cudaGetLastError());
  return PSL_OK;
}

}  // namespace psl

using namespace psl;

extern "C" {

int psl_line_extract_batch_dev(psl_ctx* ctx, const uint8_t* d_gray, int32_t B, int32_t w, int32_t h, int32_t stride,
                               int64_t frame_stride, psl_keyline* d_kl, uint8_t* d_ldesc, double* d_lineeq,
                               float* d_lbd72, int32_t cap, int32_t* d_n) {
  if (!ctx) return PSL_E_INVALID;
  if (B < 0 || w < 0 || h < 0 || cap < 1 || !d_n) return fail(ctx, PSL_E_INVALID, "bad argument");
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  if (B == 0) return PSL_OK;
  if (w == 0 || h == 0) {  // empty image: silent return, LineExtractor.cpp:327-328
    PSL_CK(cudaMemsetAsync(d_n, 0, (size_t)B * sizeof(int32_t), ctx->stream));
    return PSL_OK;
  }
  if (!d_gray || !d_kl || !d_ldesc || !d_lineeq || stride < w || (B > 1 && frame_stride < (int64_t)stride * h))
    return fail(ctx, PSL_E_INVALID, "bad image pointer / stride");
  if (ctx->cfg.line_nfeatures < 1) return fail(ctx, PSL_E_INVALID, "line_nfeatures < 1");
  int rc = set_line_geometry(ctx, w, h);
  if (rc) return rc;
  for (int c0 = 0; c0 < B; c0 += ctx->line_chunk) {
    const int nb = std::min(ctx->line_chunk, B - c0);
    ImgBatch in{d_gray + (size_t)c0 * frame_stride, stride, frame_stride, w, h};
    rc = run_line_chunk(ctx, in, nb, d_kl + (size_t)c0 * cap, d_ldesc + (size_t)c0 * cap * 32,
                        d_lineeq + (size_t)c0 * cap * 3, d_lbd72 ? d_lbd72 + (size_t)c0 * cap * 72 : nullptr, cap, d_n + c0);
    if (rc) return rc;
  }
  return PSL_OK;
}

int psl_line_extract_batch(psl_ctx* ctx, const uint8_t* gray, int32_t B, int32_t w, int32_t h, int32_t stride,
                           int64_t frame_stride, psl_keyline* kl, uint8_t* ldesc, double* lineeq, float* lbd72,
                           int32_t cap, int32_t* n) {
  if (!ctx) return PSL_E_INVALID;
  if (B < 0 || w < 0 || h < 0 || cap < 1 || !n) return fail(ctx, PSL_E_INVALID, "bad argument");
  if (B == 0) return PSL_OK;
  if (w == 0 || h == 0) {
    std::memset(n, 0, (size_t)B * sizeof(int32_t));
    return PSL_OK;
  }
  if (!gray || !kl || !ldesc || !lineeq || stride < w || (B > 1 && frame_stride < (int64_t)stride * h))
    return fail(ctx, PSL_E_INVALID, "bad image pointer / stride");
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  const int pitch = (w + 127) & ~127;
  const size_t fs = (size_t)pitch * h, N = (size_t)cap * B;
  int rc;
  if ((rc = ensure(ctx, ctx->l_in, fs * B))) return rc;
  if ((rc = ensure(ctx, ctx->l_kl, sizeof(psl_keyline) * N))) return rc;
  if ((rc = ensure(ctx, ctx->l_desc, 32 * N))) return rc;
  if ((rc = ensure(ctx, ctx->l_eq, 24 * N))) return rc;
  if (lbd72 && (rc = ensure(ctx, ctx->l_lbd, 288 * N))) return rc;
  if ((rc = ensure(ctx, ctx->l_n, sizeof(int32_t) * (size_t)B))) return rc;
  for (int b = 0; b < B; ++b)
    PSL_CK(cudaMemcpy2DAsync(ctx->l_in.as<uint8_t>() + b * fs, pitch, gray + (size_t)b * frame_stride, stride, w, h,
                             cudaMemcpyHostToDevice, ctx->stream));
  for (;;) {   // (again with more room for raw segments when a frame overflowed an auto-sized bound)
    rc = psl_line_extract_batch_dev(ctx, ctx->l_in.as<uint8_t>(), B, w, h, pitch, (int64_t)fs, ctx->l_kl.as<psl_keyline>(),
                                    ctx->l_desc.as<uint8_t>(), ctx->l_eq.as<double>(), lbd72 ? ctx->l_lbd.as<float>() : nullptr,
                                    cap, ctx->l_n.as<int32_t>());
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    PSL_CK(cudaMemcpyAsync(n, ctx->l_n.p, sizeof(int32_t) * (size_t)B, cudaMemcpyDeviceToHost, st));
    PSL_CK(cudaMemcpyAsync(kl, ctx->l_kl.p, sizeof(psl_keyline) * N, cudaMemcpyDeviceToHost, st));
    PSL_CK(cudaMemcpyAsync(ldesc, ctx->l_desc.p, 32 * N, cudaMemcpyDeviceToHost, st));
    PSL_CK(cudaMemcpyAsync(lineeq, ctx->l_eq.p, 24 * N, cudaMemcpyDeviceToHost, st));
    if (lbd72) PSL_CK(cudaMemcpyAsync(lbd72, ctx->l_lbd.p, 288 * N, cudaMemcpyDeviceToHost, st));
    rc = check_status(ctx);
    if (rc != PSL_E_CAPACITY || !ctx->grew) return rc;
  }
}

int psl_line_extract(psl_ctx* ctx, const uint8_t* gray, int32_t w, int32_t h, int32_t stride, psl_keyline* kl,
                     uint8_t* ldesc, double* lineeq, float* lbd72, int32_t cap, int32_t* n) {
  return psl_line_extract_batch(ctx, gray, 1, w, h, stride, (int64_t)stride * h, kl, ldesc, lineeq, lbd72, cap, n);
}

}  // extern "C"

// line selectors of psl_debug_fetch (called from psl_capi.cu)
int psl_line_debug_fetch(psl_ctx* ctx, int32_t what, int32_t frame, void* out, int64_t cap_bytes, int64_t* n) {
  const LineBuffers& L = ctx->lb;
  if (!ctx->lgeo_w || frame < 0 || frame >= ctx->line_chunk) return fail(ctx, PSL_E_INVALID, "debug_fetch: no line geometry / bad frame");
  PSL_CK(cudaStreamSynchronize(ctx->stream));
  if (what == 4) {
    const int64_t npx = (int64_t)L.Ws * L.Hs;
    if (npx > cap_bytes) return fail(ctx, PSL_E_CAPACITY, "debug_fetch: buffer too small");
    PSL_CK(cudaMemcpy(out, L.scaled + (size_t)frame * npx, npx, cudaMemcpyDeviceToHost));
    *n = npx;
    return PSL_OK;
  }
  int32_t cnt = 0;
  PSL_CK(cudaMemcpy(&cnt, L.n_raw + frame, sizeof(int32_t), cudaMemcpyDeviceToHost));
  if ((int64_t)cnt * 16 > cap_bytes) return fail(ctx, PSL_E_CAPACITY, "debug_fetch: buffer too small");
  PSL_CK(cudaMemcpy(out, L.raw + (size_t)frame * L.raw_cap * 4, (size_t)cnt * 16, cudaMemcpyDeviceToHost));
  *n = cnt;
  return PSL_OK;
}
