// Host-side context shared by the C-ABI translation units.
#pragma once
#include <string>
#include <vector>

#include "line_kernels.cuh"
#include "orb_kernels.cuh"

// grow-only device buffer
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct psl_ctx {
  psl_config cfg{};
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;          // the line path of the combined front end runs beside the point path
  cudaStream_t stream_copy = nullptr;      // host->device uploads of the host-pointer combined entry point
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::vector<cudaEvent_t> ev_slice;       // one per uploaded slice of gray frames (+ one for depth and poses)
  std::string err;
  int chunk = 0;
  int pool_cap = 0;
  bool pool_auto = false;   // orb_max_candidates <= 0: the pool grows when a frame overflows it
  int raw_cap = 0;          // LSD raw segments per frame (LineBuffers::raw_cap of the next geometry)
  bool raw_auto = false;    // line_max_raw <= 0: grows likewise
  uint32_t last_flags = 0;  // device status word of the last check_status()
  bool grew = false;        // ... and whether it doubled a capacity (grow_capacity)

  // ORBextractor ctor tables (ORBextractor.cc:410-446)
  std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
  std::vector<int32_t> quota;

  // geometry of the current frame size
  int geo_w = 0, geo_h = 0;
  psl::OrbGeometry geo{};
  psl::OrbGeometry* d_geo = nullptr;
  std::vector<psl::ResizeTables> rtab;
  void* d_tables = nullptr;       // resize tables of all levels
  uint8_t* d_levels = nullptr;    // pyramid levels 1.. and blurred levels 0.. of one chunk
  uint32_t* d_pool = nullptr;     // [chunk][pool_cap] FAST candidates
  uint32_t* d_pool_count = nullptr;
  uint2* d_cell_tab = nullptr;    // [chunk][total_cells] (offset,count)
  uint32_t* d_fb_list = nullptr;  // [chunk * total_cells] cells to redo at minThFAST, + 1 counter at the end
  uint32_t* d_fast_tab = nullptr; // OrbGeometry::fast_tab
  psl::FastMaps fast_maps{};      // TMA descriptors of the pyramid levels
  uint32_t* d_key_scratch = nullptr;   // [chunk][2][pool_cap]
  uint16_t* d_node_scratch = nullptr;  // [chunk][2][pool_cap]
  uint32_t* d_sel = nullptr;      // [chunk][total_sel]
  int32_t* d_sel_count = nullptr; // [chunk][nlevels]
  uint32_t* d_status = nullptr;
  uint32_t* h_status = nullptr;   // pinned mirror

  // line front end (allocated on first use): geometry + per-chunk buffers for frames of lgeo_w x lgeo_h
  int line_chunk = 0;
  int lgeo_w = 0, lgeo_h = 0;
  psl::LineBuffers lb{};
  std::vector<void*> line_allocs;
  // staging of the host-pointer line entry points
  DevBuf l_in, l_kl, l_desc, l_eq, l_lbd, l_n;

  // per-stage profiling (psl_profile_*)
  bool prof = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  struct Span { int stage; size_t e0, e1; };
  std::vector<Span> spans;
  int64_t stage_launches[PSL_N_STAGES] = {0};
  int64_t launches = 0;

  // matcher scratch (single-pair host API and the batched pipeline)
  DevBuf m_kps, m_ur, m_desc, m_q, m_qdesc, m_claimed, m_n, m_cell_start, m_cell_items, m_cand, m_cand_count,
      m_best, m_accepted, m_assign, m_nm, m_misc[16];

  // psl_track_rgbd_batch_begin / _end: two staging sets of the raw inputs (colour, depth, poses) + the gray frames
  struct FeedSlot {
    DevBuf color, depth, Tcw;
    cudaEvent_t up_done = nullptr, free_ev = nullptr;
    int B = 0, w = 0, h = 0, channels = 0, rgb_order = 0;
    bool pending = false;
  } feed[2];
  int feed_head = 0, feed_tail = 0;
  DevBuf feed_gray;

  // staging for the host-pointer entry points (grown on demand)
  uint8_t* d_in = nullptr; size_t d_in_bytes = 0;
  psl_keypoint* d_kps = nullptr; size_t d_kps_bytes = 0;
  uint8_t* d_desc = nullptr; size_t d_desc_bytes = 0;
  int32_t* d_n = nullptr; size_t d_n_bytes = 0;
};

namespace psl {
int fail(psl_ctx* c, int code, const std::string& msg);
int cuda_fail(psl_ctx* c, cudaError_t e, const char* what);
int ensure_bytes(psl_ctx* c, void** p, size_t* have, size_t need);
inline int ensure(psl_ctx* c, DevBuf& b, size_t need) { return ensure_bytes(c, &b.p, &b.bytes, need ? need : 16); }
void free_line_geometry(psl_ctx* c);
bool grow_capacity(psl_ctx* ctx);
int check_status(psl_ctx* c);  // sync + translate the device status word
// RAII-free stage bracket: begin/end record events when profiling is on and count launches.
size_t prof_mark(psl_ctx* c);
void prof_span(psl_ctx* c, int stage, size_t e0, int nlaunch);
}  // namespace psl

int psl_line_debug_fetch(psl_ctx* ctx, int32_t what, int32_t frame, void* out, int64_t cap_bytes, int64_t* n);

#define PSL_CK(call)                                                     \
  do {                                                                   \
    cudaError_t e__ = (call);                                            \
    if (e__ != cudaSuccess) return psl::cuda_fail(ctx, e__, #call);      \
  } while (0)
