// TEST INFRASTRUCTURE — throughput harness of the CPU baseline for the point front end (cfg 2/4):
// per frame ORBextractor + ComputeStereoFromRGBD, per consecutive pair the SearchByProjection
// prologue + matcher, spread over `nthreads` std::threads (frames, then pairs, one per task).
#include <atomic>
#include <thread>
#include <vector>

#include "psl_oracle.h"

extern "C" int orc_track_batch_mt(const orc_orb_params* p, const uint8_t* gray, const uint16_t* depth, int B, int w,
                                  int h, const float* Tcw /*[B][12]*/, const float* cam /*fx fy cx cy bf depth_factor*/,
                                  float th, float nn_ratio, int check_ori, int nthreads, int32_t* n_out,
                                  int32_t* nmatches_out) {
  const int cap = p->nfeatures + 4 * p->nlevels + 16;
  std::vector<psl_keypoint> kps((size_t)B * cap);
  std::vector<uint8_t> desc((size_t)B * cap * 32);
  std::vector<float> ur((size_t)B * cap), z((size_t)B * cap);
  std::vector<float> scale(p->nlevels);
  orc_orb_tables(p, scale.data(), nullptr, nullptr, nullptr);
  std::atomic<int> next(0), err(0);
  auto run = [&](auto&& fn) {
    next = 0;
    std::vector<std::thread> th_;
    for (int t = 1; t < nthreads; ++t) th_.emplace_back(fn);
    fn();
    for (auto& t : th_) t.join();
  };
  run([&]() {
    for (;;) {
      const int b = next.fetch_add(1);
      if (b >= B) break;
      int n = 0;
      if (orc_orb_extract(p, gray + (size_t)b * w * h, w, h, w, &kps[(size_t)b * cap], &desc[(size_t)b * cap * 32], cap, &n))
        err = 1;
      n_out[b] = n;
      orc_stereo_from_rgbd(&kps[(size_t)b * cap], n, depth + (size_t)b * w * h, w, h, w, cam[5], cam[4],
                           &ur[(size_t)b * cap], &z[(size_t)b * cap]);
    }
  });
  nmatches_out[0] = 0;
  run([&]() {
    std::vector<psl_proj_query> q(cap);
    std::vector<int32_t> assign(cap);
    for (;;) {
      const int b = 1 + next.fetch_add(1);
      if (b >= B) break;
      const size_t o0 = (size_t)(b - 1) * cap, o1 = (size_t)b * cap;
      orc_queries_from_last_frame(&kps[o0], &z[o0], nullptr, nullptr, n_out[b - 1], Tcw + (size_t)(b - 1) * 12,
                                  Tcw + (size_t)b * 12, cam, scale.data(), th, 0, 0.f, 0.f, (float)w, (float)h, q.data());
      psl_frame_view f{n_out[b], &kps[o1], &ur[o1], &desc[o1 * 32], 0.f, 0.f, (float)w, (float)h,
                       (float)PSL_GRID_COLS / (float)w, (float)PSL_GRID_ROWS / (float)h};
      psl_match_params mp{0, 100, nn_ratio, check_ori};
      int32_t nm = 0;
      orc_match_projection(&f, q.data(), &desc[o0 * 32], n_out[b - 1], nullptr, &mp, assign.data(), &nm);
      nmatches_out[b] = nm;
    }
  });
  return err.load() ? -1 : 0;
}

// Combined front end (cfg 4): the point chain above plus, per frame, LINEextractor::operator() and, per
// consecutive pair, LSDmatcher::SearchByGeomNApearance(Cur, Last, desc_th) (Tracking.cc:1182-1183).
extern "C" int orc_frontend_batch_mt(const orc_orb_params* p, const uint8_t* gray, const uint16_t* depth, int B, int w,
                                     int h, const float* Tcw, const float* cam, float th, float nn_ratio, int check_ori,
                                     int line_nfeatures, float line_desc_th, int nthreads, int32_t* n_out,
                                     int32_t* nmatches_out, int32_t* nl_out, int32_t* line_nmatches_out) {
  int rc = orc_track_batch_mt(p, gray, depth, B, w, h, Tcw, cam, th, nn_ratio, check_ori, nthreads, n_out, nmatches_out);
  if (rc) return rc;
  const int cap = line_nfeatures;
  std::vector<psl_keyline> kl((size_t)B * cap);
  std::vector<uint8_t> ld((size_t)B * cap * 32);
  std::vector<double> eq((size_t)B * cap * 3);
  std::atomic<int> next(0), err(0);
  auto run = [&](auto&& fn) {
    next = 0;
    std::vector<std::thread> th_;
    for (int t = 1; t < nthreads; ++t) th_.emplace_back(fn);
    fn();
    for (auto& t : th_) t.join();
  };
  run([&]() {
    for (;;) {
      const int b = next.fetch_add(1);
      if (b >= B) break;
      int n = 0;
      if (orc_line_extract(gray + (size_t)b * w * h, w, h, w, line_nfeatures, &kl[(size_t)b * cap], &ld[(size_t)b * cap * 32],
                           &eq[(size_t)b * cap * 3], nullptr, cap, &n))
        err = 1;
      nl_out[b] = n;
    }
  });
  line_nmatches_out[0] = 0;
  const float bounds[4] = {0.f, 0.f, (float)w, (float)h};
  run([&]() {
    std::vector<int32_t> assign(cap);
    std::vector<uint8_t> has(cap, 1);
    for (;;) {
      const int b = 1 + next.fetch_add(1);
      if (b >= B) break;
      const size_t o0 = (size_t)(b - 1) * cap, o1 = (size_t)b * cap;
      line_nmatches_out[b] = orc_line_search_geom(&kl[o0], &ld[o0 * 32], has.data(), nl_out[b - 1], &kl[o1], &ld[o1 * 32],
                                                  nl_out[b], bounds, line_desc_th, assign.data());
    }
  });
  return err.load() ? -1 : 0;
}

// cvtColor RGB / BGR (A) -> GRAY as cv2 4.13 computes it: Y = (R 9798 + G 19235 + B 3735 + 2^14) >> 15;
// what Tracking::GrabImageRGBD does before the Frame is built (src/Tracking.cc:219-232).
extern "C" void orc_color_to_gray(const uint8_t* color, int channels, int rgb_order, uint8_t* gray, int64_t n_px) {
  const int ri = rgb_order ? 0 : 2, bi = rgb_order ? 2 : 0;
  for (int64_t i = 0; i < n_px; ++i) {
    const uint8_t* c = color + i * channels;
    gray[i] = (uint8_t)((c[ri] * 9798 + c[1] * 19235 + c[bi] * 3735 + 16384) >> 15);
  }
}

// GrabImageRGBD for a batch: the conversion above, then orc_frontend_batch_mt (the CPU arm of bench.py with RGB-D input)
extern "C" int orc_rgbd_frontend_batch_mt(const orc_orb_params* p, const uint8_t* color, int channels, int rgb_order,
                                          const uint16_t* depth, int B, int w, int h, const float* Tcw, const float* cam,
                                          float th, float nn_ratio, int check_ori, int line_nfeatures, float line_desc_th,
                                          int nthreads, int32_t* n_out, int32_t* nmatches_out, int32_t* nl_out,
                                          int32_t* line_nmatches_out) {
  const size_t px = (size_t)w * h;
  std::vector<uint8_t> gray(px * B);
  std::atomic<int> next(0);
  auto fn = [&]() {
    for (;;) {
      const int b = next.fetch_add(1);
      if (b >= B) break;
      orc_color_to_gray(color + (size_t)b * px * channels, channels, rgb_order, &gray[(size_t)b * px], (int64_t)px);
    }
  };
  std::vector<std::thread> th_;
  for (int t = 1; t < nthreads; ++t) th_.emplace_back(fn);
  fn();
  for (auto& t : th_) t.join();
  return orc_frontend_batch_mt(p, gray.data(), depth, B, w, h, Tcw, cam, th, nn_ratio, check_ori, line_nfeatures,
                               line_desc_th, nthreads, n_out, nmatches_out, nl_out, line_nmatches_out);
}
