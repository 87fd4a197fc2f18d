// K7/K8/K9: Hamming matchers of ORBmatcher / LSDmatcher on CUDA cores (__popc; nothing here is a
// dense contraction).  The reference matchers are greedy and sequential: a keypoint already given to
// an earlier query is skipped by later ones (ORBmatcher.cc:87-89, 209-210, 1403-1405), and distance
// ties go to the first candidate in Frame::GetFeaturesInArea's enumeration order.  The work is split
// into an embarrassingly parallel stage (ordered candidate lists with distances, one warp per query)
// and an ordered resolve stage (one warp per frame walks the queries in order; the 32 lanes scan a
// query's candidates together and pick the lexicographic minimum of (distance, position)).
#include <climits>

#include "match_kernels.cuh"

namespace psl {

__device__ __forceinline__ int hamming256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
         __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__device__ __forceinline__ unsigned warp_min_u32(unsigned v) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}

// rotation-histogram bin, ORBmatcher.cc:1431-1438 (factor = 1.0f/HISTO_LENGTH quirk kept)
__device__ __forceinline__ int rot_bin(float a1, float a2) {
  float rot = __fsub_rn(a1, a2);
  if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
  int bin = (int)roundf(__fmul_rn(rot, 1.0f / 30));
  return bin == 30 ? 0 : bin;
}

// ComputeThreeMaxima, ORBmatcher.cc:1601-1642
__device__ void three_maxima(const int* h, int& i1, int& i2, int& i3) {
  int m1 = 0, m2 = 0, m3 = 0;
  i1 = i2 = i3 = -1;
  for (int i = 0; i < 30; ++i) {
    const int s = h[i];
    if (s > m1) { m3 = m2; m2 = m1; m1 = s; i3 = i2; i2 = i1; i1 = i; }
    else if (s > m2) { m3 = m2; m2 = s; i3 = i2; i2 = i; }
    else if (s > m3) { m3 = s; i3 = i; }
  }
  if ((float)m2 < __fmul_rn(0.1f, (float)m1)) { i2 = -1; i3 = -1; }
  else if ((float)m3 < __fmul_rn(0.1f, (float)m1)) { i3 = -1; }
}

// ---------------------------------------------------------------------------------------------
// Frame::AssignFeaturesToGrid / PosInGrid (Frame.cc:269-284, 1040-1050): CSR over 64x48 cells,
// items inside a cell in ascending keypoint index (= push_back order).  One CTA per frame.
// ---------------------------------------------------------------------------------------------
constexpr int kGridThreads = 256;

__global__ void __launch_bounds__(kGridThreads)
    grid_build_kernel(MatchFrames f, int32_t* __restrict__ cell_start, uint16_t* __restrict__ cell_items) {
  __shared__ int s_cnt[kGridCells];
  __shared__ int s_warp[kGridThreads / 32];
  __shared__ int s_carry;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = f.n[b];
  const psl_keypoint* kps = f.kps + (size_t)b * f.cap;
  int32_t* start = cell_start + (size_t)b * (kGridCells + 1);
  uint16_t* items = cell_items + (size_t)b * f.cap;
  for (int c = tid; c < kGridCells; c += kGridThreads) s_cnt[c] = 0;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  auto cell_of = [&](int i) {
    const int px = (int)roundf(__fmul_rn(__fsub_rn(kps[i].x, f.min_x), f.grid_w_inv));
    const int py = (int)roundf(__fmul_rn(__fsub_rn(kps[i].y, f.min_y), f.grid_h_inv));
    if (px < 0 || px >= PSL_GRID_COLS || py < 0 || py >= PSL_GRID_ROWS) return -1;
    return px * PSL_GRID_ROWS + py;
  };
  for (int i = tid; i < n; i += kGridThreads) {
    const int c = cell_of(i);
    if (c >= 0) atomicAdd(&s_cnt[c], 1);
  }
  __syncthreads();
  // exclusive scan of the 3072 counts, 256 at a time
  for (int base = 0; base < kGridCells; base += kGridThreads) {
    const int v = s_cnt[base + tid];
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, inc, d);
      if ((tid & 31) >= d) inc += o;
    }
    if ((tid & 31) == 31) s_warp[tid >> 5] = inc;
    __syncthreads();
    int wbase = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kGridThreads / 32; ++w) {
      if (w < (tid >> 5)) wbase += s_warp[w];
      tot += s_warp[w];
    }
    const int excl = s_carry + wbase + inc - v;
    start[base + tid] = excl;
    s_cnt[base + tid] = excl;  // becomes the fill cursor
    __syncthreads();
    if (tid == 0) s_carry += tot;
    __syncthreads();
  }
  if (tid == 0) start[kGridCells] = s_carry;
  for (int i = tid; i < n; i += kGridThreads) {
    const int c = cell_of(i);
    if (c >= 0) items[atomicAdd(&s_cnt[c], 1)] = (uint16_t)i;
  }
  __syncthreads();
  // restore push_back order inside every cell (cells hold a handful of items)
  for (int c = tid; c < kGridCells; c += kGridThreads) {
    const int s = start[c], e = s_cnt[c];
    for (int i = s + 1; i < e; ++i) {
      const uint16_t v = items[i];
      int j = i - 1;
      while (j >= s && items[j] > v) { items[j + 1] = items[j]; --j; }
      items[j + 1] = v;
    }
  }
}

void launch_grid_build(const MatchFrames& f, int32_t* cell_start, uint16_t* cell_items, int B, cudaStream_t st) {
  grid_build_kernel<<<B, kGridThreads, 0, st>>>(f, cell_start, cell_items);
}

// ---------------------------------------------------------------------------------------------
// Candidate lists: Frame::GetFeaturesInArea (Frame.cc:985-1038) + the static gates of
// SearchByProjection (right-coordinate test, ORBmatcher.cc:91-96 / 1407-1413) + Hamming distance.
// One warp per query; 32 grid cells are examined per step in the reference's (ix outer, iy inner)
// order, and a warp prefix sum keeps the output in enumeration order.
// ---------------------------------------------------------------------------------------------
constexpr int kCandWarps = 8;

// (8 CTAs per SM: the kernel waits on its gathers, 64 resident warps at 32 registers beat 32 at 60: 5.4 -> 4.3 ms per 4096 frames)
__global__ void __launch_bounds__(kCandWarps * 32, 8)
    proj_candidates_kernel(MatchFrames f, MatchQueries qs, const int32_t* __restrict__ cell_start,
                           const uint16_t* __restrict__ cell_items, uint32_t* __restrict__ cand,
                           int32_t* __restrict__ cand_count, uint2* __restrict__ cand_best,
                           uint32_t* __restrict__ status) {
  const int lane = threadIdx.x & 31, b = blockIdx.y;
  const int qi = blockIdx.x * kCandWarps + (threadIdx.x >> 5);
  if (qi >= qs.nq[b]) return;
  const psl_proj_query Q = qs.q[(size_t)b * qs.cap + qi];
  int32_t* out_count = cand_count + (size_t)b * qs.cap + qi;
  uint2* out_best = cand_best + (size_t)b * qs.cap + qi;
  if (!(Q.flags & PSL_Q_VALID)) {
    if (lane == 0) { *out_count = 0; *out_best = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu); }
    return;
  }
  const float x = Q.u, y = Q.v, r = Q.radius;
  const float dxm = __fsub_rn(x, f.min_x), dym = __fsub_rn(y, f.min_y);
  const int c0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(dxm, r), f.grid_w_inv)));
  const int c1 = min(PSL_GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(dxm, r), f.grid_w_inv)));
  const int r0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(dym, r), f.grid_h_inv)));
  const int r1 = min(PSL_GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(dym, r), f.grid_h_inv)));
  if (c0 >= PSL_GRID_COLS || c1 < 0 || r0 >= PSL_GRID_ROWS || r1 < 0) {
    if (lane == 0) { *out_count = 0; *out_best = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu); }
    return;
  }
  const bool check_levels = (Q.min_level > 0) || (Q.max_level >= 0);
  const psl_keypoint* kps = f.kps + (size_t)b * f.cap;
  const float* ur = f.u_right ? f.u_right + (size_t)b * f.cap : nullptr;
  const uint4* fdesc = reinterpret_cast<const uint4*>(f.desc + (size_t)b * f.cap * 32);
  const uint4* qd = reinterpret_cast<const uint4*>(qs.desc + ((size_t)b * qs.cap + qi) * 32);
  const uint4 q0 = __ldg(qd), q1 = __ldg(qd + 1);
  const int32_t* start = cell_start + (size_t)b * (kGridCells + 1);
  const uint16_t* items = cell_items + (size_t)b * f.cap;
  uint32_t* out = cand + ((size_t)b * qs.cap + qi) * kCandCap;

  auto passes = [&](int i) {
    const psl_keypoint kp = kps[i];
    if (check_levels) {
      if (kp.octave < Q.min_level) return false;
      if (Q.max_level >= 0 && kp.octave > Q.max_level) return false;
    }
    if (!(fabsf(__fsub_rn(kp.x, x)) < r && fabsf(__fsub_rn(kp.y, y)) < r)) return false;
    if (ur) {
      const float u2 = ur[i];
      if (u2 > 0.f && fabsf(__fsub_rn(Q.u_right, u2)) > r) return false;
    }
    return true;
  };

  // The CSR grid is cell-major with ix outer and iy inner, exactly the enumeration order of GetFeaturesInArea: the cells
  // (ix, r0 .. r1) of one window column are one contiguous run of items.  A window is therefore a handful of runs
  // (one per column); the lanes take the items of a run side by side, a ballot keeps the passing ones in order.
  const int ncols = c1 - c0 + 1;
  int total = 0;
  unsigned k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;  // two smallest (dist << 16 | position) seen by this lane
  const unsigned lt = (1u << lane) - 1u;
  for (int cb = 0; cb < ncols; cb += 32) {
    int rs = 0, re = 0;   // run of window column cb + lane
    if (cb + lane < ncols) {
      const int ix = c0 + cb + lane;
      rs = start[ix * PSL_GRID_ROWS + r0];
      re = start[ix * PSL_GRID_ROWS + r1 + 1];
    }
    const int nc = min(32, ncols - cb);
    for (int c = 0; c < nc; ++c) {
      const int s = __shfl_sync(0xffffffffu, rs, c), e = __shfl_sync(0xffffffffu, re, c);
      for (int k0 = s; k0 < e; k0 += 32) {
        const int k = k0 + lane;
        int i = 0;
        bool ok = false;
        if (k < e) {
          i = items[k];
          ok = passes(i);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        const int pos = total + __popc(bal & lt);
        if (ok && pos < kCandCap) {
          const uint4 d0 = __ldg(fdesc + 2 * i), d1 = __ldg(fdesc + 2 * i + 1);
          const unsigned dist = (unsigned)hamming256(q0, q1, d0, d1);
          out[pos] = ((uint32_t)i << 16) | dist;
          const unsigned key = (dist << 16) | (unsigned)pos;
          if (key < k1) { k2 = k1; k1 = key; }
          else if (key < k2) k2 = key;
        }
        total += __popc(bal);
      }
    }
  }
  const unsigned best = warp_min_u32(k1);
  const unsigned second = warp_min_u32(k1 == best ? k2 : k1);
  if (lane == 0) {
    *out_count = min(total, kCandCap);
    *out_best = make_uint2(best, second);
    if (total > kCandCap) { atomicOr(status, kStatCandOverflow); atomicMax(status + 1, (uint32_t)b + 1u); }
  }
}

void launch_proj_candidates(const MatchFrames& f, const MatchQueries& q, const int32_t* cell_start,
                            const uint16_t* cell_items, uint32_t* cand, int32_t* cand_count, uint2* cand_best,
                            uint32_t* status, int B, cudaStream_t st) {
  dim3 grid((q.cap + kCandWarps - 1) / kCandWarps, B);
  proj_candidates_kernel<<<grid, kCandWarps * 32, 0, st>>>(f, q, cell_start, cell_items, cand, cand_count, cand_best,
                                                           status);
}

// ---------------------------------------------------------------------------------------------
// Ordered resolve: the loops of SearchByProjection (ORBmatcher.cc:64-124 / 1353-1444) with the
// dynamic "already claimed" test, then the rotation-consistency filter (:1447-1467).
// One warp per frame; dynamic shared memory: claimed[cap] bytes.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
    proj_resolve_kernel(MatchFrames f, MatchQueries qs, const uint32_t* __restrict__ cand,
                        const int32_t* __restrict__ cand_count, const uint2* __restrict__ cand_best,
                        const uint8_t* __restrict__ claimed_in, psl_match_params prm,
                        uint32_t* __restrict__ accepted_scratch, int32_t* __restrict__ assign_out,
                        int32_t* __restrict__ nmatches) {
  extern __shared__ uint8_t s_claimed[];
  __shared__ int s_hist[32];
  const int lane = threadIdx.x, b = blockIdx.x;
  const int n = f.n[b], nq = qs.nq[b];
  int32_t* assign = assign_out + (size_t)b * f.cap;
  const psl_keypoint* kps = f.kps + (size_t)b * f.cap;
  uint32_t* accepted = accepted_scratch + (size_t)b * qs.cap;  // (query << 16 | keypoint) of accepted matches
  for (int i = lane; i < n; i += 32) {
    s_claimed[i] = claimed_in ? claimed_in[(size_t)b * f.cap + i] : 0;
    assign[i] = -1;
  }
  s_hist[lane] = 0;
  __syncwarp();
  int nacc = 0;
  for (int base = 0; base < nq; base += 32) {
    // 32 queries at a time: every lane preloads one query's unconstrained best/second and its flags
    const int ql = base + lane;
    uint2 bs = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
    unsigned bidx = 0, sidx = 0, claims = 0;
    if (ql < nq) {
      bs = cand_best[(size_t)b * qs.cap + ql];
      const uint32_t* cl = cand + ((size_t)b * qs.cap + ql) * kCandCap;
      if (bs.x != 0xFFFFFFFFu) bidx = cl[bs.x & 0xFFFFu] >> 16;
      if (bs.y != 0xFFFFFFFFu) sidx = cl[bs.y & 0xFFFFu] >> 16;
      claims = (qs.q[(size_t)b * qs.cap + ql].flags & PSL_Q_CLAIMS) ? 1u : 0u;
    }
    const int lim = min(32, nq - base);
    for (int j = 0; j < lim; ++j) {
      unsigned best = __shfl_sync(0xffffffffu, bs.x, j);
      if (best == 0xFFFFFFFFu) continue;  // invalid query or empty window
      unsigned second = __shfl_sync(0xffffffffu, bs.y, j);
      unsigned bi = __shfl_sync(0xffffffffu, bidx, j), si = __shfl_sync(0xffffffffu, sidx, j);
      const unsigned cf = __shfl_sync(0xffffffffu, claims, j);
      const int q = base + j;
      const uint32_t* cl = cand + ((size_t)b * qs.cap + q) * kCandCap;
      // the precomputed pair is still right unless one of the two keypoints was claimed meanwhile
      const bool stale = s_claimed[bi] || (prm.mode == 1 && second != 0xFFFFFFFFu && s_claimed[si]);
      if (stale) {
        const int cnt = cand_count[(size_t)b * qs.cap + q];
        unsigned k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
        for (int k = lane; k < cnt; k += 32) {
          const uint32_t e = cl[k];
          if (s_claimed[e >> 16]) continue;
          const unsigned key = ((e & 0xFFFFu) << 16) | (unsigned)k;
          if (key < k1) { k2 = k1; k1 = key; }
          else if (key < k2) k2 = key;
        }
        best = warp_min_u32(k1);
        if (best == 0xFFFFFFFFu) continue;
        second = warp_min_u32(k1 == best ? k2 : k1);
        bi = cl[best & 0xFFFFu] >> 16;
        if (second != 0xFFFFFFFFu) si = cl[second & 0xFFFFu] >> 16;
      }
      const int bestDist = (int)(best >> 16);
      if (bestDist > prm.th_dist) continue;
      if (prm.mode == 1 && second != 0xFFFFFFFFu) {
        // ratio test only when best and second live on the same pyramid level (:118-119); with no second
        // candidate bestLevel2 = -1 never equals a real level (:76-79)
        const int d2 = (int)(second >> 16);
        if (kps[bi].octave == kps[si].octave && (float)bestDist > __fmul_rn(prm.nn_ratio, (float)d2)) continue;
      }
      if (lane == 0) {
        assign[bi] = q;
        s_claimed[bi] = (uint8_t)cf;
        accepted[nacc] = ((uint32_t)q << 16) | bi;
      }
      ++nacc;
      __syncwarp();
    }
  }
  int nm = nacc;
  if (prm.mode == 0 && prm.check_orientation) {
    // rotation histogram of the accepted matches (:1431-1441), then keep the three best bins (:1447-1467)
    __syncwarp();
    for (int k = lane; k < nacc; k += 32) {
      const uint32_t e = accepted[k];
      const int bin = rot_bin(qs.q[(size_t)b * qs.cap + (e >> 16)].angle, kps[e & 0xFFFFu].angle);
      atomicAdd(&s_hist[bin], 1);
      accepted[k] = ((e & 0xFFFFu) << 8) | (uint32_t)bin;
    }
    __syncwarp();
    int i1, i2, i3;
    three_maxima(s_hist, i1, i2, i3);
    int dropped = 0;
    for (int k = lane; k < nacc; k += 32) {
      const uint32_t e = accepted[k];
      const int bin = (int)(e & 0xFFu);
      if (bin != i1 && bin != i2 && bin != i3) {
        assign[e >> 8] = -1;
        ++dropped;
      }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) dropped += __shfl_xor_sync(0xffffffffu, dropped, d);
    nm -= dropped;
  }
  if (lane == 0) nmatches[b] = nm;
}

void launch_proj_resolve(const MatchFrames& f, const MatchQueries& q, const uint32_t* cand, const int32_t* cand_count,
                         const uint2* cand_best, const uint8_t* claimed_in, psl_match_params prm,
                         uint32_t* accepted_scratch, int32_t* assign, int32_t* nmatches, int B, cudaStream_t st) {
  // claimed[cap] bytes of dynamic shared memory: up to 65535 keypoints, i.e. beyond the 48 KB a kernel gets by default
  static bool once = [] {
    cudaFuncSetAttribute(proj_resolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    return true;
  }();
  (void)once;
  proj_resolve_kernel<<<B, 32, (size_t)f.cap, st>>>(f, q, cand, cand_count, cand_best, claimed_in, prm,
                                                    accepted_scratch, assign, nmatches);
}

// ---------------------------------------------------------------------------------------------
// DescriptorDistance (ORBmatcher.cc:1647-1663) and BFMatcher knn2 (LSDmatcher.cpp:354-376)
// ---------------------------------------------------------------------------------------------
__global__ void descriptor_distance_kernel(const uint4* a, const uint4* b, int n, int32_t* dist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dist[i] = hamming256(a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1]);
}

void launch_descriptor_distance(const uint8_t* a, const uint8_t* b, int n, int32_t* dist, cudaStream_t st) {
  if (n <= 0) return;
  descriptor_distance_kernel<<<(n + 127) / 128, 128, 0, st>>>((const uint4*)a, (const uint4*)b, n, dist);
}

constexpr int kKnnWarps = 4;

__global__ void __launch_bounds__(kKnnWarps * 32)
    knn2_kernel(const uint4* __restrict__ q, int nq, const uint4* __restrict__ t, int nt, int32_t* __restrict__ idx,
                int32_t* __restrict__ dist) {
  const int lane = threadIdx.x & 31, qi = blockIdx.x * kKnnWarps + (threadIdx.x >> 5);
  if (qi >= nq) return;
  const uint4 q0 = __ldg(q + 2 * qi), q1 = __ldg(q + 2 * qi + 1);
  unsigned k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;  // (dist << 16) | train index: earliest index wins ties
  for (int j = lane; j < nt; j += 32) {
    const unsigned key = ((unsigned)hamming256(q0, q1, __ldg(t + 2 * j), __ldg(t + 2 * j + 1)) << 16) | (unsigned)j;
    if (key < k1) { k2 = k1; k1 = key; }
    else if (key < k2) k2 = key;
  }
  const unsigned best = warp_min_u32(k1);
  const unsigned second = warp_min_u32(k1 == best ? k2 : k1);
  if (lane == 0) {
    idx[2 * qi] = best == 0xFFFFFFFFu ? -1 : (int)(best & 0xFFFFu);
    dist[2 * qi] = best == 0xFFFFFFFFu ? -1 : (int)(best >> 16);
    idx[2 * qi + 1] = second == 0xFFFFFFFFu ? -1 : (int)(second & 0xFFFFu);
    dist[2 * qi + 1] = second == 0xFFFFFFFFu ? -1 : (int)(second >> 16);
  }
}

void launch_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx, int32_t* dist, cudaStream_t st) {
  if (nq <= 0) return;
  knn2_kernel<<<(nq + kKnnWarps - 1) / kKnnWarps, kKnnWarps * 32, 0, st>>>((const uint4*)q, nq, (const uint4*)t, nt,
                                                                           idx, dist);
}

// ---------------------------------------------------------------------------------------------
// SearchByBoW (ORBmatcher.cc:159-288): one warp per pair of equal vocabulary nodes.  A frame
// keypoint belongs to exactly one node, so the greedy "already matched" test never crosses warps.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    bow_kernel(const uint4* __restrict__ kf_desc, const float* __restrict__ kf_angle,
               const uint8_t* __restrict__ kf_valid, const int32_t* __restrict__ kf_offs,
               const uint32_t* __restrict__ kf_idx, const uint4* __restrict__ f_desc, const float* __restrict__ f_angle,
               const uint8_t* __restrict__ f_valid, const int32_t* __restrict__ f_offs,
               const uint32_t* __restrict__ f_idx, const int2* __restrict__ pairs, int npairs, float nn_ratio,
               int th_low, int strict, int check_ori, int32_t* match_f, int32_t* hist, uint32_t* accepted,
               int32_t* n_accepted) {
  const int lane = threadIdx.x & 31, g = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (g >= npairs) return;
  const int2 pr = pairs[g];
  const int ks = kf_offs[pr.x], ke = kf_offs[pr.x + 1], fs = f_offs[pr.y], fe = f_offs[pr.y + 1];
  for (int ik = ks; ik < ke; ++ik) {
    const int rk = (int)kf_idx[ik];
    if (!kf_valid[rk]) continue;
    const uint4 q0 = __ldg(kf_desc + 2 * rk), q1 = __ldg(kf_desc + 2 * rk + 1);
    unsigned k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
    for (int j = fs + lane; j < fe; j += 32) {
      const int rf = (int)f_idx[j];
      if (match_f[rf] >= 0) continue;  // :209-210 (KF-KF form: vbMatched2, :583)
      if (f_valid && !f_valid[rf]) continue;  // KF-KF form: the KF2 keypoint holds no good MapPoint (:583-587)
      const unsigned key = ((unsigned)hamming256(q0, q1, __ldg(f_desc + 2 * rf), __ldg(f_desc + 2 * rf + 1)) << 16) |
                           (unsigned)(j - fs);
      if (key < k1) { k2 = k1; k1 = key; }
      else if (key < k2) k2 = key;
    }
    const unsigned best = warp_min_u32(k1);
    if (best == 0xFFFFFFFFu) continue;
    const unsigned second = warp_min_u32(k1 == best ? k2 : k1);
    const int d1 = (int)(best >> 16), d2 = second == 0xFFFFFFFFu ? 256 : (int)(second >> 16);
    if ((strict ? d1 < th_low : d1 <= th_low) && (float)d1 < __fmul_rn(nn_ratio, (float)d2)) {  // :224 / :592
      const int rf = (int)f_idx[fs + (int)(best & 0xFFFFu)];
      if (lane == 0) {
        match_f[rf] = rk;
        if (check_ori) {
          const int bin = rot_bin(kf_angle[rk], f_angle[rf]);
          atomicAdd(&hist[bin], 1);
          accepted[atomicAdd(n_accepted, 1)] = ((uint32_t)rf << 8) | (uint32_t)bin;
        } else {
          atomicAdd(n_accepted, 1);
        }
      }
    }
    __syncwarp();
  }
}

__global__ void bow_finalize_kernel(int check_ori, int32_t* match_f, const int32_t* hist, const uint32_t* accepted,
                                    const int32_t* n_accepted, int32_t* nmatches) {
  __shared__ int s_drop;
  if (threadIdx.x == 0) s_drop = 0;
  __syncthreads();
  const int n = *n_accepted;
  if (check_ori) {
    int i1, i2, i3;
    three_maxima(hist, i1, i2, i3);
    int d = 0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
      const uint32_t e = accepted[k];
      const int bin = (int)(e & 0xFFu);
      if (bin != i1 && bin != i2 && bin != i3) { match_f[e >> 8] = -1; ++d; }
    }
    atomicAdd(&s_drop, d);
  }
  __syncthreads();
  if (threadIdx.x == 0) *nmatches = n - s_drop;
}

__global__ void bow_invert_kernel(const int32_t* __restrict__ match_f, int nf, int32_t* __restrict__ m12) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nf && match_f[i] >= 0) m12[match_f[i]] = i;  // a KF1 keypoint is matched at most once
}

void launch_bow(const uint8_t* kf_desc, const float* kf_angle, const uint8_t* kf_valid, const int32_t* kf_offs,
                const uint32_t* kf_idx, const uint8_t* f_desc, const float* f_angle, const uint8_t* f_valid,
                const int32_t* f_offs, const uint32_t* f_idx, const int2* pairs, int npairs, float nn_ratio, int th_low,
                int strict, int check_ori, int nf, int32_t* match_f, int32_t* hist, uint32_t* accepted,
                int32_t* n_accepted, int32_t* nmatches, int32_t* m12, int nkf, cudaStream_t st) {
  cudaMemsetAsync(match_f, 0xFF, (size_t)nf * sizeof(int32_t), st);
  cudaMemsetAsync(hist, 0, 32 * sizeof(int32_t), st);
  cudaMemsetAsync(n_accepted, 0, sizeof(int32_t), st);
  if (npairs > 0)
    bow_kernel<<<(npairs + 3) / 4, 128, 0, st>>>((const uint4*)kf_desc, kf_angle, kf_valid, kf_offs, kf_idx,
                                                 (const uint4*)f_desc, f_angle, f_valid, f_offs, f_idx, pairs, npairs,
                                                 nn_ratio, th_low, strict, check_ori, match_f, hist, accepted,
                                                 n_accepted);
  bow_finalize_kernel<<<1, 128, 0, st>>>(check_ori, match_f, hist, accepted, n_accepted, nmatches);
  if (m12) {  // KF-KF form: the result is indexed by the first keyframe (vpMatches12)
    cudaMemsetAsync(m12, 0xFF, (size_t)nkf * sizeof(int32_t), st);
    if (nf > 0) bow_invert_kernel<<<(nf + 127) / 128, 128, 0, st>>>(match_f, nf, m12);
  }
}

// ---------------------------------------------------------------------------------------------
// SearchForTriangulation (ORBmatcher.cc:657-823): one warp per pair of equal vocabulary nodes walks the
// KF1 keypoints of the node; the lanes scan the KF2 keypoints.  The reference never sets vbMatched2, so the
// KF1 keypoints are independent: best = smallest distance among the candidates that pass every gate, the
// LAST one on ties (`dist > bestDist` skips, an equal distance replaces).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    triang_kernel(const psl_keypoint* __restrict__ kps1, const float* __restrict__ ur1, const uint4* __restrict__ desc1,
                  const uint8_t* __restrict__ mp1, const int32_t* __restrict__ offs1, const uint32_t* __restrict__ idx1,
                  const psl_keypoint* __restrict__ kps2, const float* __restrict__ ur2, const uint4* __restrict__ desc2,
                  const uint8_t* __restrict__ mp2, const int32_t* __restrict__ offs2, const uint32_t* __restrict__ idx2,
                  const int2* __restrict__ pairs, int npairs, const float* __restrict__ F12, float ex, float ey,
                  const float* __restrict__ scale2, const float* __restrict__ sigma2, int only_stereo, int th_low,
                  int check_ori, int32_t* __restrict__ m12, int32_t* __restrict__ hist, int32_t* __restrict__ nmatches) {
  const int lane = threadIdx.x & 31, g = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (g >= npairs) return;
  const int2 pr = pairs[g];
  const int s1 = offs1[pr.x], e1 = offs1[pr.x + 1], s2 = offs2[pr.y], e2 = offs2[pr.y + 1];
  const float f00 = F12[0], f01 = F12[1], f02 = F12[2], f10 = F12[3], f11 = F12[4], f12 = F12[5], f20 = F12[6],
              f21 = F12[7], f22 = F12[8];
  int nm = 0;
  for (int i1 = s1; i1 < e1; ++i1) {
    const int a = (int)idx1[i1];
    if (mp1[a]) continue;
    const bool st1 = ur1[a] >= 0.f;
    if (only_stereo && !st1) continue;
    const psl_keypoint kp1 = kps1[a];
    const uint4 q0 = __ldg(desc1 + 2 * a), q1 = __ldg(desc1 + 2 * a + 1);
    // epipolar line l = x1' F12 (CheckDistEpipolarLine :142-145), no contraction
    const float ea = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, f00), __fmul_rn(kp1.y, f10)), f20);
    const float eb = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, f01), __fmul_rn(kp1.y, f11)), f21);
    const float ec = __fadd_rn(__fadd_rn(__fmul_rn(kp1.x, f02), __fmul_rn(kp1.y, f12)), f22);
    const float den = __fadd_rn(__fmul_rn(ea, ea), __fmul_rn(eb, eb));
    unsigned best = 0xFFFFFFFFu;
    for (int j = s2 + lane; j < e2; j += 32) {
      const int b = (int)idx2[j];
      if (mp2[b]) continue;
      const bool st2 = ur2[b] >= 0.f;
      if (only_stereo && !st2) continue;
      const int dist = hamming256(q0, q1, __ldg(desc2 + 2 * b), __ldg(desc2 + 2 * b + 1));
      if (dist > th_low) continue;
      const psl_keypoint kp2 = kps2[b];
      if (!st1 && !st2) {
        const float dx = __fsub_rn(ex, kp2.x), dy = __fsub_rn(ey, kp2.y);
        if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.f, scale2[kp2.octave])) continue;
      }
      if (den == 0.f) continue;
      const float num = __fadd_rn(__fadd_rn(__fmul_rn(ea, kp2.x), __fmul_rn(eb, kp2.y)), ec);
      const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
      if (!((double)dsqr < 3.84 * (double)sigma2[kp2.octave])) continue;
      const unsigned key = ((unsigned)dist << 16) | (unsigned)(0xFFFF - (j - s2));  // later position = smaller key
      best = min(best, key);
    }
    best = warp_min_u32(best);
    if (best == 0xFFFFFFFFu) continue;
    const int b = (int)idx2[s2 + (0xFFFF - (int)(best & 0xFFFFu))];
    if (lane == 0) {
      m12[a] = b;
      if (check_ori) atomicAdd(&hist[rot_bin(kp1.angle, kps2[b].angle)], 1);
    }
    ++nm;
  }
  if (lane == 0 && nm) atomicAdd(nmatches, nm);
}

// rotation-consistency filter (:792-808): drop the matches outside the three dominant bins
__global__ void triang_filter_kernel(const psl_keypoint* __restrict__ kps1, const psl_keypoint* __restrict__ kps2, int n1,
                                     const int32_t* __restrict__ hist, int32_t* __restrict__ m12,
                                     int32_t* __restrict__ nmatches) {
  int ind1 = -1, ind2 = -1, ind3 = -1;
  three_maxima(hist, ind1, ind2, ind3);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n1) return;
  const int b = m12[i];
  if (b < 0) return;
  const int bin = rot_bin(kps1[i].angle, kps2[b].angle);
  if (bin != ind1 && bin != ind2 && bin != ind3) {
    m12[i] = -1;
    atomicSub(nmatches, 1);
  }
}

void launch_triangulation(const psl_keypoint* kps1, const float* ur1, const uint8_t* desc1, const uint8_t* mp1,
                          const int32_t* offs1, const uint32_t* idx1, int n1, const psl_keypoint* kps2, const float* ur2,
                          const uint8_t* desc2, const uint8_t* mp2, const int32_t* offs2, const uint32_t* idx2,
                          const int2* pairs, int npairs, const float* F12, float ex, float ey, const float* scale2,
                          const float* sigma2, int only_stereo, int th_low, int check_ori, int32_t* m12, int32_t* hist,
                          int32_t* nmatches, cudaStream_t st) {
  cudaMemsetAsync(m12, 0xFF, (size_t)n1 * sizeof(int32_t), st);
  cudaMemsetAsync(hist, 0, 33 * sizeof(int32_t), st);  // hist[32] | nmatches
  if (npairs > 0)
    triang_kernel<<<(npairs + 3) / 4, 128, 0, st>>>(kps1, ur1, (const uint4*)desc1, mp1, offs1, idx1, kps2, ur2,
                                                    (const uint4*)desc2, mp2, offs2, idx2, pairs, npairs, F12, ex, ey,
                                                    scale2, sigma2, only_stereo, th_low, check_ori, m12, hist, nmatches);
  if (check_ori && n1 > 0) triang_filter_kernel<<<(n1 + 127) / 128, 128, 0, st>>>(kps1, kps2, n1, hist, m12, nmatches);
}

// ---------------------------------------------------------------------------------------------
// The window search of ORBmatcher::Fuse (ORBmatcher.cc:893-950): warp per projected MapPoint, lanes over the
// grid cells of KeyFrame::GetFeaturesInArea.  Queries are independent; best = smallest distance, earliest
// candidate in the reference's enumeration (cell column-major, in-cell order) on ties ->
// key = dist << 32 | cell rank << 16 | slot.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kCandWarps * 32)
    fuse_kernel(MatchFrames f, const psl_fuse_query* __restrict__ qs, const uint8_t* __restrict__ qdesc, int nq,
                const int32_t* __restrict__ cell_start, const uint16_t* __restrict__ cell_items,
                const float* __restrict__ inv_sigma2, int th_low, int32_t* __restrict__ best_idx,
                int32_t* __restrict__ best_dist) {
  const int lane = threadIdx.x & 31, qi = blockIdx.x * kCandWarps + (threadIdx.x >> 5);
  if (qi >= nq) return;
  const psl_fuse_query Q = qs[qi];
  unsigned long long best = ~0ull;
  if (Q.flags & PSL_Q_VALID) {
    const float x = Q.u, y = Q.v, r = Q.radius;
    const float dxm = __fsub_rn(x, f.min_x), dym = __fsub_rn(y, f.min_y);
    const int c0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(dxm, r), f.grid_w_inv)));
    const int c1 = min(PSL_GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(dxm, r), f.grid_w_inv)));
    const int r0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(dym, r), f.grid_h_inv)));
    const int r1 = min(PSL_GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(dym, r), f.grid_h_inv)));
    if (c0 < PSL_GRID_COLS && c1 >= 0 && r0 < PSL_GRID_ROWS && r1 >= 0) {
      const uint4* qd = reinterpret_cast<const uint4*>(qdesc + (size_t)qi * 32);
      const uint4 q0 = __ldg(qd), q1 = __ldg(qd + 1);
      const uint4* fdesc = reinterpret_cast<const uint4*>(f.desc);
      const int ncy = r1 - r0 + 1, ncells = (c1 - c0 + 1) * ncy;
      for (int c = lane; c < ncells; c += 32) {
        const int ix = c0 + c / ncy, iy = r0 + c % ncy;
        const int s = cell_start[ix * PSL_GRID_ROWS + iy], e = cell_start[ix * PSL_GRID_ROWS + iy + 1];
        for (int k = s; k < e; ++k) {
          const int i = cell_items[k];
          const psl_keypoint kp = f.kps[i];
          if (!(fabsf(__fsub_rn(kp.x, x)) < r && fabsf(__fsub_rn(kp.y, y)) < r)) continue;
          if (kp.octave < Q.pred_level - 1 || kp.octave > Q.pred_level) continue;
          const float ex = __fsub_rn(x, kp.x), ey = __fsub_rn(y, kp.y);
          const float kr = f.u_right ? f.u_right[i] : -1.f;
          if (!inv_sigma2) {
            // the Sim3 forms (ORBmatcher.cc:1046-1075, 1188-1216) have no reprojection gate
          } else if (kr >= 0.f) {
            const float er = __fsub_rn(Q.u_right, kr);
            const float e2 = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(er, er));
            if ((double)__fmul_rn(e2, inv_sigma2[kp.octave]) > 7.8) continue;
          } else {
            const float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
            if ((double)__fmul_rn(e2, inv_sigma2[kp.octave]) > 5.99) continue;
          }
          const unsigned dist = (unsigned)hamming256(q0, q1, __ldg(fdesc + 2 * i), __ldg(fdesc + 2 * i + 1));
          const unsigned long long key = ((unsigned long long)dist << 32) | ((unsigned long long)c << 16) | (unsigned)(k - s);
          best = key < best ? key : best;
        }
      }
#pragma unroll
      for (int d = 16; d; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, d);
        best = o < best ? o : best;
      }
      if (best != ~0ull) {
        // recover the keypoint index from (cell rank, slot)
        const int c = (int)((best >> 16) & 0xFFFFu), k = (int)(best & 0xFFFFu);
        const int ix = c0 + c / ncy, iy = r0 + c % ncy;
        const int i = cell_items[cell_start[ix * PSL_GRID_ROWS + iy] + k];
        const int dist = (int)(best >> 32);
        if (lane == 0) {
          best_idx[qi] = dist <= th_low ? i : -1;
          if (best_dist) best_dist[qi] = dist;
        }
        return;
      }
    }
  }
  if (lane == 0) {
    best_idx[qi] = -1;
    if (best_dist) best_dist[qi] = 256;
  }
}

void launch_fuse(const MatchFrames& f, const psl_fuse_query* qs, const uint8_t* qdesc, int nq, const int32_t* cell_start,
                 const uint16_t* cell_items, const float* inv_sigma2, int th_low, int32_t* best_idx, int32_t* best_dist,
                 cudaStream_t st) {
  if (nq <= 0) return;
  fuse_kernel<<<(nq + kCandWarps - 1) / kCandWarps, kCandWarps * 32, 0, st>>>(f, qs, qdesc, nq, cell_start, cell_items,
                                                                             inv_sigma2, th_low, best_idx, best_dist);
}

// ---------------------------------------------------------------------------------------------
// SearchBySim3 (ORBmatcher.cc:1102-1326): the two window searches are fuse_kernel launches without a gate; this is
// the agreement test (:1306-1321).
// ---------------------------------------------------------------------------------------------
__global__ void sim3_agree_kernel(const int32_t* __restrict__ m1, int n1, const int32_t* __restrict__ m2,
                                  int32_t* __restrict__ out, int32_t* __restrict__ nfound) {
  const int i1 = blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = false;
  if (i1 < n1) {
    const int idx2 = m1[i1];
    ok = idx2 >= 0 && m2[idx2] == i1;
    out[i1] = ok ? idx2 : -1;
  }
  const unsigned b = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(nfound, __popc(b));
}

void launch_sim3_agree(const int32_t* m1, int n1, const int32_t* m2, int32_t* out, int32_t* nfound, cudaStream_t st) {
  cudaMemsetAsync(nfound, 0, sizeof(int32_t), st);
  if (n1 > 0) sim3_agree_kernel<<<(n1 + 127) / 128, 128, 0, st>>>(m1, n1, m2, out, nfound);
}

// ---------------------------------------------------------------------------------------------
// SearchForInitialization (ORBmatcher.cc:405-520): the greedy loop over the F1 keypoints, replayed in order by one
// warp on the candidate lists of proj_candidates_kernel (window order, with distances).  A later keypoint may take
// an F2 keypoint from an earlier one if it is strictly closer (vMatchedDistance, :441-442; the earlier pair is
// dissolved, :461-465), so nothing about a list can be decided before its turn.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
    init_resolve_kernel(const uint32_t* __restrict__ cand, const int32_t* __restrict__ cand_count,
                        const psl_keypoint* __restrict__ kps1, int n1, const psl_keypoint* __restrict__ kps2, int n2,
                        float nn_ratio, int th_low, int check_ori, int32_t* mdist, int32_t* m21, uint32_t* accepted,
                        int32_t* m12, float* prev_matched, int32_t* nmatches) {
  __shared__ int s_hist[32];
  const int lane = threadIdx.x;
  for (int i = lane; i < n2; i += 32) { mdist[i] = INT_MAX; m21[i] = -1; }
  for (int i = lane; i < n1; i += 32) m12[i] = -1;
  s_hist[lane] = 0;
  __syncwarp();
  int nm = 0, nacc = 0;
  for (int base = 0; base < n1; base += 32) {
    const int my_cnt = base + lane < n1 ? cand_count[base + lane] : 0;
    unsigned todo = __ballot_sync(0xffffffffu, my_cnt > 0);
    while (todo) {
      const int j = __ffs(todo) - 1;
      todo &= todo - 1;
      const int i1 = base + j, cnt = __shfl_sync(0xffffffffu, my_cnt, j);
      const uint32_t* cl = cand + (size_t)i1 * kCandCap;
      unsigned k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
      for (int k = lane; k < cnt; k += 32) {
        const uint32_t e = cl[k];
        const int d = (int)(e & 0xFFFFu);
        if (mdist[e >> 16] <= d) continue;  // :441-442
        const unsigned key = ((unsigned)d << 16) | (unsigned)k;
        if (key < k1) { k2 = k1; k1 = key; }
        else if (key < k2) k2 = key;
      }
      const unsigned best = warp_min_u32(k1);
      if (best == 0xFFFFFFFFu) continue;
      const unsigned second = warp_min_u32(k1 == best ? k2 : k1);
      const int bestDist = (int)(best >> 16);
      const float d2 = second == 0xFFFFFFFFu ? (float)INT_MAX : (float)(int)(second >> 16);
      if (bestDist <= th_low && (float)bestDist < __fmul_rn(d2, nn_ratio)) {  // :456-458
        const int i2 = (int)(cl[best & 0xFFFFu] >> 16);
        const int old = __shfl_sync(0xffffffffu, lane == 0 ? m21[i2] : 0, 0);  // read by the lane that rewrites it
        if (lane == 0) {
          if (old >= 0) m12[old] = -1;
          m12[i1] = i2;
          m21[i2] = i1;
          mdist[i2] = bestDist;
          if (check_ori) {
            const int bin = rot_bin(kps1[i1].angle, kps2[i2].angle);
            s_hist[bin] += 1;  // a dissolved pair stays in its bin (rotHist is never shrunk, :461-465)
            accepted[nacc] = ((uint32_t)i1 << 8) | (uint32_t)bin;
          }
        }
        nm += old >= 0 ? 0 : 1;
        ++nacc;
        __syncwarp();
      }
    }
  }
  __syncwarp();
  if (check_ori) {
    int b1, b2, b3;
    three_maxima(s_hist, b1, b2, b3);
    int dropped = 0;
    for (int k = lane; k < nacc; k += 32) {
      const uint32_t e = accepted[k];
      const int bin = (int)(e & 0xFFu);
      if (bin != b1 && bin != b2 && bin != b3 && m12[e >> 8] >= 0) {  // :499-503
        m12[e >> 8] = -1;
        ++dropped;
      }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) dropped += __shfl_xor_sync(0xffffffffu, dropped, d);
    nm -= dropped;
    __syncwarp();
  }
  for (int i = lane; i < n1; i += 32) {  // :514-517
    const int i2 = m12[i];
    if (i2 >= 0) {
      prev_matched[2 * i] = kps2[i2].x;
      prev_matched[2 * i + 1] = kps2[i2].y;
    }
  }
  if (lane == 0) *nmatches = nm;
}

void launch_init_resolve(const uint32_t* cand, const int32_t* cand_count, const psl_keypoint* kps1, int n1,
                         const psl_keypoint* kps2, int n2, float nn_ratio, int th_low, int check_ori, int32_t* mdist,
                         int32_t* m21, uint32_t* accepted, int32_t* m12, float* prev_matched, int32_t* nmatches,
                         cudaStream_t st) {
  init_resolve_kernel<<<1, 32, 0, st>>>(cand, cand_count, kps1, n1, kps2, n2, nn_ratio, th_low, check_ori, mdist, m21,
                                        accepted, m12, prev_matched, nmatches);
}

}  // namespace psl
