// C-ABI of the front end (include/psl_frontend.h): context, ORB tables and the ORB extractor
// entry points.  Host code is C++ as in the reference; it owns no algorithmic work beyond
// the constructor tables (ORBextractor.cc:410-470) and the per-size level geometry.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "psl_ctx.cuh"

namespace psl {

int fail(psl_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}
int cuda_fail(psl_ctx* c, cudaError_t e, const char* what) {
  return fail(c, PSL_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
int ensure_bytes(psl_ctx* ctx, void** p, size_t* have, size_t need) {
  if (*have >= need) return PSL_OK;
  if (*p) PSL_CK(cudaFree(*p));
  *p = nullptr;
  *have = 0;
  PSL_CK(cudaMalloc(p, need));
  *have = need;
  return PSL_OK;
}
static void free_geometry(psl_ctx* c);

int check_status(psl_ctx* ctx) {
  PSL_CK(cudaMemcpyAsync(ctx->h_status, ctx->d_status, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  PSL_CK(cudaStreamSynchronize(ctx->stream));
  const uint32_t s = ctx->h_status[0];
  ctx->last_flags = s;
  ctx->grew = false;
  if (!s) return PSL_OK;
  // word 1: 1 + the largest frame index (inside its launch chunk) that raised a flag, so that the caller of a batched
  // entry point knows where to look; the other frames of the batch are complete
  const std::string where = ctx->h_status[1] ? " (frame " + std::to_string(ctx->h_status[1] - 1) + " of its launch chunk)" : "";
  PSL_CK(cudaMemsetAsync(ctx->d_status, 0, 2 * sizeof(uint32_t), ctx->stream));
  if (s & kStatBadRoot) return fail(ctx, PSL_E_INVALID, "octree: candidate outside the root nodes (aspect ratio)");
  if (s & kStatNodeOverflow) return fail(ctx, PSL_E_INTERNAL, "octree: node table bound violated");
  // a capacity the context sizes itself is doubled right here: the call that hit it fails once (the host-pointer
  // extractors run it again themselves), the next one has the room
  ctx->grew = grow_capacity(ctx);
  const std::string again = ctx->grew ? " -- capacity doubled, call again" : "";
  if (s & kStatCandOverflow)
    return fail(ctx, PSL_E_CAPACITY, std::string("FAST candidate pool overflow: raise psl_config.orb_max_candidates") + where + again);
  if (s & kStatOutOverflow) return fail(ctx, PSL_E_CAPACITY, std::string("output capacity `cap` too small") + where);
  if (s & kStatLineRaw)
    return fail(ctx, PSL_E_CAPACITY, std::string("LSD raw segment overflow: raise psl_config.line_max_raw") + where + again);
  if (s & kStatLineNeighbours) return fail(ctx, PSL_E_CAPACITY, std::string("line merge: neighbour list bound violated") + where);
  return fail(ctx, PSL_E_INTERNAL, "unknown device status");
}

size_t prof_mark(psl_ctx* c) {
  if (!c->prof) return 0;
  if (c->ev_used == c->ev_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    c->ev_pool.push_back(e);
  }
  cudaEventRecord(c->ev_pool[c->ev_used], c->stream);
  return c->ev_used++;
}
void prof_span(psl_ctx* c, int stage, size_t e0, int nlaunch) {
  c->launches += nlaunch;
  c->stage_launches[stage] += nlaunch;
  if (!c->prof) return;
  const size_t e1 = prof_mark(c);
  c->spans.push_back({stage, e0, e1});
}

static inline int cv_round(float v) { return (int)lrintf(v); }  // cvRound: half-to-even

// One axis of cv::resize(INTER_LINEAR) for CV_8U: tap indices and Q11 weights.
static void resize_axis(int sn, int dn, bool zero_f, std::vector<short4>& out) {
  out.resize(dn);
  const double scale = (double)sn / dn;
  for (int d = 0; d < dn; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= s;
    if (zero_f) {
      if (s < 0) { s = 0; f = 0.f; }
      if (s >= sn - 1) { s = sn - 1; f = 0.f; }
    }
    const int s0 = std::min(std::max(s, 0), sn - 1), s1 = std::min(std::max(s + 1, 0), sn - 1);
    out[d] = make_short4((short)s0, (short)s1, (short)cv_round((1.f - f) * 2048.f), (short)cv_round(f * 2048.f));
  }
}

static void free_geometry(psl_ctx* c) {
  cudaFree(c->d_tables); c->d_tables = nullptr;
  cudaFree(c->d_levels); c->d_levels = nullptr;
  cudaFree(c->d_pool); c->d_pool = nullptr;
  cudaFree(c->d_pool_count); c->d_pool_count = nullptr;
  cudaFree(c->d_cell_tab); c->d_cell_tab = nullptr;
  cudaFree(c->d_fb_list); c->d_fb_list = nullptr;
  cudaFree(c->d_fast_tab); c->d_fast_tab = nullptr;
  cudaFree(c->d_key_scratch); c->d_key_scratch = nullptr;
  cudaFree(c->d_node_scratch); c->d_node_scratch = nullptr;
  cudaFree(c->d_sel); c->d_sel = nullptr;
  cudaFree(c->d_sel_count); c->d_sel_count = nullptr;
  c->geo_w = c->geo_h = 0;
}

// After a PSL_E_CAPACITY from check_status(): double the capacity that overflowed if the context owns it (auto mode)
// and drop the buffers sized by it, so that the same call can simply be made again.  False: nothing to grow.
bool grow_capacity(psl_ctx* ctx) {
  const uint32_t s = ctx->last_flags;
  bool grown = false;
  if ((s & kStatCandOverflow) && ctx->pool_auto && ctx->pool_cap < (1 << 22)) {
    ctx->pool_cap *= 2;
    cudaStreamSynchronize(ctx->stream);
    free_geometry(ctx);
    grown = true;
  }
  if ((s & kStatLineRaw) && ctx->raw_auto && ctx->raw_cap < 65535) {
    ctx->raw_cap = std::min(2 * ctx->raw_cap, 65535);
    cudaStreamSynchronize(ctx->stream);
    free_line_geometry(ctx);
    grown = true;
  }
  return grown;
}

// Level sizes, cell grids, octree roots and device buffers for frames of w x h.
static int set_geometry(psl_ctx* ctx, int w, int h) {
  if (ctx->geo_w == w && ctx->geo_h == h) return PSL_OK;
  if (w > ctx->cfg.max_width || h > ctx->cfg.max_height)
    return fail(ctx, PSL_E_CAPACITY, "frame larger than psl_config.max_width/max_height");
  PSL_CK(cudaStreamSynchronize(ctx->stream));
  free_geometry(ctx);
  OrbGeometry& g = ctx->geo;
  std::memset(&g, 0, sizeof(g));
  const int L = ctx->cfg.orb_nlevels;
  g.nlevels = L;
  size_t bytes = 0;
  int cells = 0, sel = 0;
  std::vector<size_t> lvl_off(L), blur_off(L);
  for (int l = 0; l < L; ++l) {
    const int lw = cv_round((float)w * ctx->inv_scale[l]), lh = cv_round((float)h * ctx->inv_scale[l]);  // :1111-1112
    if (lw < 2 * kEdge + 7 || lh < 2 * kEdge + 7 || lw > 4096 + 32 || lh > 4096 + 32)
      return fail(ctx, PSL_E_INVALID, "frame size not supported at this pyramid depth");
    const int pitch = (lw + 127) & ~127;
    const int64_t fs = (int64_t)pitch * lh;
    g.level[l] = ImgBatchMut{nullptr, pitch, fs, lw, lh};
    g.blur[l] = ImgBatchMut{nullptr, pitch, fs, lw, lh};
    if (l) { lvl_off[l] = bytes; bytes += (size_t)fs * ctx->chunk; }
    blur_off[l] = bytes;
    bytes += (size_t)fs * ctx->chunk;
    // cell grid :773-787
    CellGrid& cg = g.grid[l];
    cg.max_bx = lw - kEdge + 3;
    cg.max_by = lh - kEdge + 3;
    const float width = (float)(cg.max_bx - kMinBorder), height = (float)(cg.max_by - kMinBorder);
    cg.n_cols = (int)(width / (float)kCellW);
    cg.n_rows = (int)(height / (float)kCellW);
    if (cg.n_cols <= 0 || cg.n_rows <= 0) return fail(ctx, PSL_E_INVALID, "pyramid level smaller than one FAST cell");
    cg.w_cell = (int)ceilf(width / cg.n_cols);
    cg.h_cell = (int)ceilf(height / cg.n_rows);
    if (cg.w_cell + 6 > kMaxCellDim || cg.h_cell + 6 > kMaxCellDim)
      return fail(ctx, PSL_E_INVALID, "FAST cell larger than the kernel tile");
    cg.first_cell = cells;
    cells += cg.n_cols * cg.n_rows;
    // octree roots :543-545
    g.quota[l] = ctx->quota[l];
    g.n_ini[l] = (int)roundf(width / height);
    if (g.n_ini[l] < 1 || g.n_ini[l] > 4)
      return fail(ctx, PSL_E_INVALID, "aspect ratio needs 1..4 octree roots (round(W/H))");
    g.hx[l] = width / g.n_ini[l];
    g.sel_cap[l] = std::max(g.quota[l], 4 * g.n_ini[l]) + 8;
    g.sel_off[l] = sel;
    sel += g.sel_cap[l];
    g.scale[l] = ctx->scale[l];
    g.kp_size[l] = (float)(int)(kPatch * ctx->scale[l]);  // :837
  }
  g.total_cells = cells;
  g.total_sel = sel;
  PSL_CK(cudaMalloc(&ctx->d_levels, bytes));
  for (int l = 0; l < L; ++l) {
    if (l) g.level[l].ptr = ctx->d_levels + lvl_off[l];
    g.blur[l].ptr = ctx->d_levels + blur_off[l];
  }
  // resize tables: per-pixel (x, y) and, for the word kernel, the x table per group of 4 outputs
  std::vector<short4> all;
  std::vector<size_t> xo(L), yo(L), go(L);
  std::vector<uint8_t> grouped;  // per level: [n4] uint4 weights, then [n4] u32 offsets (16-byte aligned blocks)
  std::vector<char> grouped_ok(L, 0);
  std::vector<short4> t;
  for (int l = 1; l < L; ++l) {
    resize_axis(g.level[l - 1].w, g.level[l].w, true, t);
    xo[l] = all.size();
    all.insert(all.end(), t.begin(), t.end());
    const int n4 = (g.level[l].w + 3) >> 2;
    const size_t blk = ((size_t)n4 * 20 + 15) & ~(size_t)15;
    go[l] = grouped.size();
    grouped.resize(grouped.size() + blk);
    std::vector<uint4> xw(n4);
    std::vector<uint32_t> xoff(n4);
    if (resize_group_tables(t.data(), g.level[l].w, xw.data(), xoff.data())) {
      grouped_ok[l] = 1;
      std::memcpy(grouped.data() + go[l], xw.data(), (size_t)n4 * 16);
      std::memcpy(grouped.data() + go[l] + (size_t)n4 * 16, xoff.data(), (size_t)n4 * 4);
    }
    resize_axis(g.level[l - 1].h, g.level[l].h, false, t);
    yo[l] = all.size();
    all.insert(all.end(), t.begin(), t.end());
  }
  const size_t all_bytes = (all.size() * sizeof(short4) + 15) & ~(size_t)15;
  PSL_CK(cudaMalloc(&ctx->d_tables, std::max<size_t>(all_bytes + grouped.size(), 16)));
  PSL_CK(cudaMemcpyAsync(ctx->d_tables, all.data(), all.size() * sizeof(short4), cudaMemcpyHostToDevice, ctx->stream));
  if (!grouped.empty())
    PSL_CK(cudaMemcpyAsync((uint8_t*)ctx->d_tables + all_bytes, grouped.data(), grouped.size(), cudaMemcpyHostToDevice,
                           ctx->stream));
  ctx->rtab.assign(L, ResizeTables{nullptr, nullptr, nullptr, nullptr});
  for (int l = 1; l < L; ++l) {
    const uint8_t* gb = (const uint8_t*)ctx->d_tables + all_bytes + go[l];
    const int n4 = (g.level[l].w + 3) >> 2;
    ctx->rtab[l] = ResizeTables{(const short4*)ctx->d_tables + xo[l], (const short4*)ctx->d_tables + yo[l],
                                grouped_ok[l] ? (const uint4*)gb : nullptr,
                                grouped_ok[l] ? (const uint32_t*)(gb + (size_t)n4 * 16) : nullptr};
  }
  const size_t C = ctx->chunk, P = ctx->pool_cap;
  PSL_CK(cudaMalloc(&ctx->d_pool, C * P * sizeof(uint32_t)));
  PSL_CK(cudaMalloc(&ctx->d_pool_count, C * sizeof(uint32_t)));
  PSL_CK(cudaMalloc(&ctx->d_cell_tab, C * cells * sizeof(uint2)));
  PSL_CK(cudaMalloc(&ctx->d_fb_list, (C * cells + 1) * sizeof(uint32_t)));
  PSL_CK(cudaMalloc(&ctx->d_key_scratch, C * 2 * P * sizeof(uint32_t)));
  PSL_CK(cudaMalloc(&ctx->d_node_scratch, C * 2 * P * sizeof(uint16_t)));
  PSL_CK(cudaMalloc(&ctx->d_sel, C * sel * sizeof(uint32_t)));
  PSL_CK(cudaMalloc(&ctx->d_sel_count, C * L * sizeof(int32_t)));
  // dense FAST path: tile / cell lookup tables and the TMA descriptors of our own levels
  g.n_tiles = fast_tile_count(g);
  std::vector<uint32_t> ftab((size_t)g.n_tiles + cells);
  fast_build_tab(g, ftab.data());
  PSL_CK(cudaMalloc(&ctx->d_fast_tab, ftab.size() * sizeof(uint32_t)));
  PSL_CK(cudaMemcpyAsync(ctx->d_fast_tab, ftab.data(), ftab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice,
                         ctx->stream));
  g.fast_tab = ctx->d_fast_tab;
  ctx->fast_maps.valid = 0;
  for (int l = 1; l < L; ++l)
    fast_encode_map(ctx->fast_maps, l, g.level[l].ptr, g.level[l].w, g.level[l].h, g.level[l].pitch,
                    g.level[l].frame_stride, ctx->chunk, g.grid[l].h_cell + 6);
  PSL_CK(cudaMemcpyAsync(ctx->d_geo, &g, sizeof(g), cudaMemcpyHostToDevice, ctx->stream));
  PSL_CK(cudaStreamSynchronize(ctx->stream));  // `all`, `grouped`, `ftab` and `g` are host temporaries
  ctx->geo_w = w;
  ctx->geo_h = h;
  return PSL_OK;
}

// The whole extractor for `nb` <= chunk frames resident in HBM.
static int run_chunk(psl_ctx* ctx, ImgBatch in0, int nb, psl_keypoint* d_kps, uint8_t* d_desc, int cap, int32_t* d_n) {
  const OrbGeometry& g = ctx->geo;
  cudaStream_t st = ctx->stream;
  PSL_CK(cudaMemsetAsync(ctx->d_pool_count, 0, nb * sizeof(uint32_t), st));
  size_t e = prof_mark(ctx);
  for (int l = 1; l < g.nlevels; ++l) {  // ComputePyramid :1107-1132
    ImgBatch src = l == 1 ? in0
                          : ImgBatch{g.level[l - 1].ptr, g.level[l - 1].pitch, g.level[l - 1].frame_stride,
                                     g.level[l - 1].w, g.level[l - 1].h};
    launch_resize(src, g.level[l], ctx->rtab[l], nb, st);
  }
  prof_span(ctx, 0, e, g.nlevels - 1);
  e = prof_mark(ctx);
  launch_fast_cells(ctx->d_geo, g, in0, ctx->cfg.orb_ini_th_fast, ctx->cfg.orb_min_th_fast, ctx->d_pool,
                    ctx->pool_cap, ctx->d_pool_count, ctx->d_cell_tab, ctx->d_fb_list,
                    ctx->d_fb_list + (size_t)ctx->chunk * g.total_cells, ctx->fast_maps, ctx->d_status, nb, st);
  prof_span(ctx, 1, e, kFastLaunches);
  e = prof_mark(ctx);
  launch_octree(ctx->d_geo, g, ctx->d_pool, ctx->pool_cap, ctx->d_cell_tab, ctx->d_key_scratch, ctx->d_node_scratch,
                ctx->d_sel, ctx->d_sel_count, ctx->d_status, nb, st);
  prof_span(ctx, 2, e, 1);
  e = prof_mark(ctx);
  launch_gauss7(g, in0, nb, st);
  prof_span(ctx, 3, e, kBlurLaunches * g.nlevels);
  e = prof_mark(ctx);
  launch_describe(ctx->d_geo, g, in0, ctx->d_sel, ctx->d_sel_count, d_kps, d_desc, cap, d_n, ctx->d_status, nb, st);
  prof_span(ctx, 4, e, 1);
  PSL_CK(cudaGetLastError());
  return PSL_OK;
}

}  // namespace psl

using namespace psl;

extern "C" {

void psl_default_config(psl_config* cfg) {
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->device = 0;
  cfg->max_width = 640;
  cfg->max_height = 480;
  cfg->max_batch = 64;
  cfg->orb_nfeatures = 1000;  // Examples/RGB-D/TUM1.yaml:42-55
  cfg->orb_scale_factor = 1.2f;
  cfg->orb_nlevels = 8;
  cfg->orb_ini_th_fast = 20;
  cfg->orb_min_th_fast = 7;
  cfg->orb_max_candidates = 0;
  cfg->chunk_frames = 0;
  cfg->line_nfeatures = 200;  // TUM1.yaml:60-63
  cfg->line_scale_factor = 1.2f;
  cfg->line_nlevels = 1;
  cfg->line_min_length = 0.f;
  cfg->line_chunk_frames = 0;
  cfg->line_max_raw = 0;
}

int psl_create(const psl_config* cfg, psl_ctx** out) {
  if (!cfg || !out) return PSL_E_INVALID;
  *out = nullptr;
  if (cfg->orb_nlevels < 1 || cfg->orb_nlevels > kMaxLevels || cfg->orb_nfeatures < 1 ||
      !(cfg->orb_scale_factor > 1.0f) || cfg->orb_ini_th_fast < 1 || cfg->orb_min_th_fast < 1 ||
      cfg->orb_ini_th_fast > 254 || cfg->orb_min_th_fast > 254 || cfg->max_width < 1 || cfg->max_height < 1)
    return PSL_E_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || cfg->device < 0 || cfg->device >= ndev) return PSL_E_CUDA;
  if (cudaSetDevice(cfg->device) != cudaSuccess) return PSL_E_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess || prop.major < 10) return PSL_E_CUDA;
  psl_ctx* ctx = new psl_ctx();
  ctx->cfg = *cfg;
  const int L = cfg->orb_nlevels;
  // ORBextractor ctor :410-446
  const double sf = (double)cfg->orb_scale_factor;
  ctx->scale.assign(L, 1.f);
  for (int i = 1; i < L; ++i) ctx->scale[i] = (float)(ctx->scale[i - 1] * sf);
  ctx->inv_scale.resize(L);
  ctx->sigma2.resize(L);
  ctx->inv_sigma2.resize(L);
  for (int i = 0; i < L; ++i) {
    ctx->sigma2[i] = ctx->scale[i] * ctx->scale[i];
    ctx->inv_scale[i] = 1.0f / ctx->scale[i];
    ctx->inv_sigma2[i] = 1.0f / ctx->sigma2[i];
  }
  ctx->quota.resize(L);
  {
    const float factor = (float)(1.0f / sf);
    float nd = cfg->orb_nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)L));
    int sum = 0;
    for (int l = 0; l < L - 1; ++l) {
      ctx->quota[l] = cv_round(nd);
      sum += ctx->quota[l];
      nd *= factor;
    }
    ctx->quota[L - 1] = std::max(cfg->orb_nfeatures - sum, 0);
  }
  // frames per extraction chunk: every per-chunk buffer is sized by it, so a context made for single frames stays small
  ctx->chunk = cfg->chunk_frames > 0 ? cfg->chunk_frames : std::min(512, std::max(cfg->max_batch, 1));
  ctx->line_chunk = cfg->line_chunk_frames > 0 ? cfg->line_chunk_frames : std::min(std::max(cfg->max_batch, 64), 4096);
  // capacities that depend on the image content: a positive value is a hard bound (overflow = PSL_E_CAPACITY), 0 or a
  // negative value starts at the default / at |value| and grows when a frame overflows it (grow_capacity)
  ctx->pool_cap = cfg->orb_max_candidates > 0   ? cfg->orb_max_candidates
                  : cfg->orb_max_candidates < 0 ? -cfg->orb_max_candidates
                                                : std::max(16384, 32 * cfg->orb_nfeatures);
  ctx->pool_auto = cfg->orb_max_candidates <= 0;
  ctx->raw_cap = cfg->line_max_raw > 0 ? cfg->line_max_raw : (cfg->line_max_raw < 0 ? -cfg->line_max_raw : 4096);
  ctx->raw_auto = cfg->line_max_raw <= 0;
  bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&ctx->stream_copy, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) == cudaSuccess &&
            cudaMalloc(&ctx->d_geo, sizeof(OrbGeometry)) == cudaSuccess &&
            cudaMalloc(&ctx->d_status, 2 * sizeof(uint32_t)) == cudaSuccess &&
            cudaMemset(ctx->d_status, 0, 2 * sizeof(uint32_t)) == cudaSuccess &&
            cudaMallocHost(&ctx->h_status, 2 * sizeof(uint32_t)) == cudaSuccess;
  if (!ok) {
    psl_destroy(ctx);
    return PSL_E_CUDA;
  }
  *out = ctx;
  return PSL_OK;
}

void psl_destroy(psl_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->cfg.device);
  if (ctx->stream2) cudaStreamSynchronize(ctx->stream2);
  if (ctx->stream_copy) cudaStreamSynchronize(ctx->stream_copy);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  free_geometry(ctx);
  free_line_geometry(ctx);
  cudaFree(ctx->d_geo);
  cudaFree(ctx->d_status);
  cudaFreeHost(ctx->h_status);
  cudaFree(ctx->d_in);
  cudaFree(ctx->d_kps);
  cudaFree(ctx->d_desc);
  cudaFree(ctx->d_n);
  for (DevBuf* b : {&ctx->m_kps, &ctx->m_ur, &ctx->m_desc, &ctx->m_q, &ctx->m_qdesc, &ctx->m_claimed, &ctx->m_n,
                    &ctx->m_cell_start, &ctx->m_cell_items, &ctx->m_cand, &ctx->m_cand_count, &ctx->m_best, &ctx->m_accepted,
                    &ctx->m_assign, &ctx->m_nm})
    cudaFree(b->p);
  for (DevBuf& b : ctx->m_misc) cudaFree(b.p);
  for (psl_ctx::FeedSlot& f : ctx->feed) {
    cudaFree(f.color.p); cudaFree(f.depth.p); cudaFree(f.Tcw.p);
    if (f.up_done) cudaEventDestroy(f.up_done);
    if (f.free_ev) cudaEventDestroy(f.free_ev);
  }
  cudaFree(ctx->feed_gray.p);
  for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
  if (ctx->stream_copy) cudaStreamDestroy(ctx->stream_copy);
  for (cudaEvent_t ev : ctx->ev_slice) cudaEventDestroy(ev);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* psl_last_error(const psl_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void* psl_stream(psl_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int psl_sync(psl_ctx* ctx) {
  if (!ctx) return PSL_E_INVALID;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  return check_status(ctx);
}

int psl_orb_tables(const psl_ctx* ctx, int32_t* nlevels, float* scale, float* inv_scale, float* sigma2,
                   float* inv_sigma2, int32_t* features_per_level) {
  if (!ctx) return PSL_E_INVALID;
  const int L = ctx->cfg.orb_nlevels;
  if (nlevels) *nlevels = L;
  for (int i = 0; i < L; ++i) {
    if (scale) scale[i] = ctx->scale[i];
    if (inv_scale) inv_scale[i] = ctx->inv_scale[i];
    if (sigma2) sigma2[i] = ctx->sigma2[i];
    if (inv_sigma2) inv_sigma2[i] = ctx->inv_sigma2[i];
    if (features_per_level) features_per_level[i] = ctx->quota[i];
  }
  return PSL_OK;
}

int psl_orb_extract_batch_dev(psl_ctx* ctx, const uint8_t* d_gray, int32_t B, int32_t w, int32_t h, int32_t stride,
                              int64_t frame_stride, psl_keypoint* d_kps, uint8_t* d_desc, int32_t cap,
                              int32_t* d_n) {
  if (!ctx) return PSL_E_INVALID;
  if (B < 0 || w < 0 || h < 0 || cap < 1 || !d_n) return fail(ctx, PSL_E_INVALID, "bad argument");
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  if (B == 0) return PSL_OK;
  if (w == 0 || h == 0) {  // empty image: silent return, ORBextractor.cc:1046-1047
    PSL_CK(cudaMemsetAsync(d_n, 0, (size_t)B * sizeof(int32_t), ctx->stream));
    return PSL_OK;
  }
  if (!d_gray || !d_kps || !d_desc || stride < w || (B > 1 && frame_stride < (int64_t)stride * h))
    return fail(ctx, PSL_E_INVALID, "bad image pointer / stride");
  int rc = set_geometry(ctx, w, h);
  if (rc) return rc;
  for (int c0 = 0; c0 < B; c0 += ctx->chunk) {
    const int nb = std::min(ctx->chunk, B - c0);
    ImgBatch in0{d_gray + (size_t)c0 * frame_stride, stride, frame_stride, w, h};
    rc = run_chunk(ctx, in0, nb, d_kps + (size_t)c0 * cap, d_desc + (size_t)c0 * cap * 32, cap, d_n + c0);
    if (rc) return rc;
  }
  return PSL_OK;
}

int psl_orb_extract_batch(psl_ctx* ctx, const uint8_t* gray, int32_t B, int32_t w, int32_t h, int32_t stride,
                          int64_t frame_stride, psl_keypoint* kps, uint8_t* desc, int32_t cap, int32_t* n) {
  if (!ctx) return PSL_E_INVALID;
  if (B < 0 || w < 0 || h < 0 || cap < 1 || !n) return fail(ctx, PSL_E_INVALID, "bad argument");
  if (B == 0) return PSL_OK;
  if (w == 0 || h == 0) {
    std::memset(n, 0, (size_t)B * sizeof(int32_t));
    return PSL_OK;
  }
  if (!gray || !kps || !desc || stride < w || (B > 1 && frame_stride < (int64_t)stride * h))
    return fail(ctx, PSL_E_INVALID, "bad image pointer / stride");
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  const int pitch = (w + 127) & ~127;
  const size_t fs = (size_t)pitch * h;
  int rc;
  if ((rc = ensure_bytes(ctx, (void**)&ctx->d_in, &ctx->d_in_bytes, fs * B))) return rc;
  if ((rc = ensure_bytes(ctx, (void**)&ctx->d_kps, &ctx->d_kps_bytes, sizeof(psl_keypoint) * (size_t)cap * B))) return rc;
  if ((rc = ensure_bytes(ctx, (void**)&ctx->d_desc, &ctx->d_desc_bytes, (size_t)32 * cap * B))) return rc;
  if ((rc = ensure_bytes(ctx, (void**)&ctx->d_n, &ctx->d_n_bytes, sizeof(int32_t) * (size_t)B))) return rc;
  for (int b = 0; b < B; ++b)
    PSL_CK(cudaMemcpy2DAsync(ctx->d_in + b * fs, pitch, gray + (size_t)b * frame_stride, stride, w, h,
                             cudaMemcpyHostToDevice, ctx->stream));
  for (;;) {   // (again with a larger candidate pool when a frame overflowed an auto-sized one)
    rc = psl_orb_extract_batch_dev(ctx, ctx->d_in, B, w, h, pitch, (int64_t)fs, ctx->d_kps, ctx->d_desc, cap, ctx->d_n);
    if (rc) return rc;
    PSL_CK(cudaMemcpyAsync(n, ctx->d_n, sizeof(int32_t) * (size_t)B, cudaMemcpyDeviceToHost, ctx->stream));
    PSL_CK(cudaMemcpyAsync(kps, ctx->d_kps, sizeof(psl_keypoint) * (size_t)cap * B, cudaMemcpyDeviceToHost, ctx->stream));
    PSL_CK(cudaMemcpyAsync(desc, ctx->d_desc, (size_t)32 * cap * B, cudaMemcpyDeviceToHost, ctx->stream));
    rc = check_status(ctx);
    if (rc != PSL_E_CAPACITY || !ctx->grew) return rc;
  }
}

int psl_profile_enable(psl_ctx* ctx, int32_t on) {
  if (!ctx) return PSL_E_INVALID;
  ctx->prof = on != 0;
  return PSL_OK;
}

int psl_profile_read(psl_ctx* ctx, float* ms, int64_t* launches) {
  if (!ctx || !ms || !launches) return PSL_E_INVALID;
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  PSL_CK(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < PSL_N_STAGES; ++i) { ms[i] = 0.f; launches[i] = ctx->stage_launches[i]; ctx->stage_launches[i] = 0; }
  for (const auto& sp : ctx->spans) {
    float t = 0.f;
    PSL_CK(cudaEventElapsedTime(&t, ctx->ev_pool[sp.e0], ctx->ev_pool[sp.e1]));
    ms[sp.stage] += t;
  }
  ctx->spans.clear();
  ctx->ev_used = 0;
  return PSL_OK;
}

int64_t psl_launch_count(const psl_ctx* ctx) { return ctx ? ctx->launches : 0; }

int psl_debug_fetch(psl_ctx* ctx, int32_t what, int32_t frame, int32_t level, void* out, int64_t cap_bytes,
                    int64_t* n) {
  if (!ctx || !out || !n) return PSL_E_INVALID;
  if (what == 4 || what == 5) {
    PSL_CK(cudaSetDevice(ctx->cfg.device));
    *n = 0;
    return psl_line_debug_fetch(ctx, what, frame, out, cap_bytes, n);
  }
  const OrbGeometry& g = ctx->geo;
  if (!ctx->geo_w || frame < 0 || frame >= ctx->chunk || level < 0 || level >= g.nlevels)
    return fail(ctx, PSL_E_INVALID, "debug_fetch: no geometry / bad frame or level");
  PSL_CK(cudaSetDevice(ctx->cfg.device));
  PSL_CK(cudaStreamSynchronize(ctx->stream));
  *n = 0;
  if (what == 0 || what == 1) {
    const ImgBatchMut& im = what == 0 ? g.level[level] : g.blur[level];
    if (!im.ptr) return fail(ctx, PSL_E_INVALID, "debug_fetch: level 0 aliases the caller's image");
    if ((int64_t)im.w * im.h > cap_bytes) return fail(ctx, PSL_E_CAPACITY, "debug_fetch: buffer too small");
    PSL_CK(cudaMemcpy2D(out, im.w, im.ptr + (size_t)frame * im.frame_stride, im.pitch, im.w, im.h,
                        cudaMemcpyDeviceToHost));
    *n = (int64_t)im.w * im.h;
    return PSL_OK;
  }
  if (what == 2) {
    const CellGrid& cg = g.grid[level];
    const int nc = cg.n_cols * cg.n_rows;
    std::vector<uint2> tab(nc);
    PSL_CK(cudaMemcpy(tab.data(), ctx->d_cell_tab + (size_t)frame * g.total_cells + cg.first_cell, nc * sizeof(uint2),
                      cudaMemcpyDeviceToHost));
    std::vector<uint32_t> pool(ctx->pool_cap);
    PSL_CK(cudaMemcpy(pool.data(), ctx->d_pool + (size_t)frame * ctx->pool_cap, pool.size() * sizeof(uint32_t),
                      cudaMemcpyDeviceToHost));
    int64_t m = 0;
    uint32_t* o = (uint32_t*)out;
    for (int c = 0; c < nc; ++c)
      for (uint32_t k = 0; k < tab[c].y; ++k) {
        if ((m + 1) * 4 > cap_bytes || tab[c].x + k >= pool.size()) return fail(ctx, PSL_E_CAPACITY, "debug_fetch: buffer too small");
        o[m++] = pool[tab[c].x + k];
      }
    *n = m;
    return PSL_OK;
  }
  if (what == 3) {
    int32_t cnt = 0;
    PSL_CK(cudaMemcpy(&cnt, ctx->d_sel_count + (size_t)frame * g.nlevels + level, sizeof(int32_t), cudaMemcpyDeviceToHost));
    if ((int64_t)cnt * 4 > cap_bytes) return fail(ctx, PSL_E_CAPACITY, "debug_fetch: buffer too small");
    PSL_CK(cudaMemcpy(out, ctx->d_sel + (size_t)frame * g.total_sel + g.sel_off[level], cnt * sizeof(uint32_t),
                      cudaMemcpyDeviceToHost));
    *n = cnt;
    return PSL_OK;
  }
  return fail(ctx, PSL_E_INVALID, "debug_fetch: unknown selector");
}

int psl_orb_extract(psl_ctx* ctx, const uint8_t* gray, int32_t w, int32_t h, int32_t stride, psl_keypoint* kps,
                    uint8_t* desc, int32_t cap, int32_t* n) {
  return psl_orb_extract_batch(ctx, gray, 1, w, h, stride, (int64_t)stride * h, kps, desc, cap, n);
}

}  // extern "C"
