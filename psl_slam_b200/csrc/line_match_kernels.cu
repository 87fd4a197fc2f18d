// Line matchers on the device (add_src/LSDmatcher.cpp, add_src/InsectlineMatch.cpp, src/Map.cc:204-272).
// Batch-first like the point matchers: every kernel works on B independent (query set, searched frame) pairs
// laid out as [B][cap] row blocks; the single-pair C-ABI calls run them with B = 1.  Work per pair is tiny
// (<= a few hundred 32-byte descriptors): warp-cooperative __popc Hamming with ballot/shuffle reductions.
#include "line_match_kernels.cuh"

namespace psl {

__device__ __forceinline__ int ham256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
         __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}
__device__ __forceinline__ unsigned wmin32(unsigned v) {
#pragma unroll
  for (int d = 16; d; d >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}
__device__ __forceinline__ unsigned long long wmin64(unsigned long long v) {
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, d);
    v = o < v ? o : v;
  }
  return v;
}

// ---------------------------------------------------------------------------------------------------
// cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, 2): warp per query, earliest train index wins ties.
// knn[B][cap] = (best key, second key), key = dist << 16 | train index, 0xFFFFFFFF = none.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
    line_knn2_kernel(LineSet Q, LineSet T, uint2* __restrict__ knn) {
  const int lane = threadIdx.x & 31, b = blockIdx.y, qi = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int nq = Q.n[b], nt = T.n[b];
  if (qi >= nq) return;
  const uint4* qd = reinterpret_cast<const uint4*>(Q.desc + ((size_t)b * Q.cap + qi) * 32);
  const uint4* td = reinterpret_cast<const uint4*>(T.desc + (size_t)b * T.cap * 32);
  const uint4 q0 = __ldg(qd), q1 = __ldg(qd + 1);
  unsigned k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
  for (int j = lane; j < nt; j += 32) {
    const unsigned key = ((unsigned)ham256(q0, q1, __ldg(td + 2 * j), __ldg(td + 2 * j + 1)) << 16) | (unsigned)j;
    if (key < k1) { k2 = k1; k1 = key; }
    else if (key < k2) k2 = key;
  }
  const unsigned best = wmin32(k1);
  const unsigned second = wmin32(k1 == best ? k2 : k1);
  if (lane == 0) knn[(size_t)b * Q.cap + qi] = make_uint2(best, second);
}

void launch_line_knn2(const LineSet& Q, const LineSet& T, uint2* knn, int B, cudaStream_t st) {
  dim3 grid((Q.cap + 3) / 4, B);
  line_knn2_kernel<<<grid, 128, 0, st>>>(Q, T, knn);
}

// LSDmatcher::matchNNR (LSDmatcher.cpp:354-376): d0 < d1 * nnr on the float distances
__global__ void line_nnr_kernel(LineSet Q, const uint2* __restrict__ knn, float nnr, int32_t* __restrict__ m12,
                                int32_t* __restrict__ nmatches) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Q.n[b]) return;
  const uint2 k = knn[(size_t)b * Q.cap + i];
  int m = -1;
  if (k.y != 0xFFFFFFFFu && (float)(k.x >> 16) < __fmul_rn((float)(k.y >> 16), nnr)) m = (int)(k.x & 0xFFFFu);
  m12[(size_t)b * Q.cap + i] = m;
  if (m >= 0) atomicAdd(nmatches + b, 1);
}

void launch_line_nnr(const LineSet& Q, const uint2* knn, float nnr, int32_t* m12, int32_t* nmatches, int B,
                     cudaStream_t st) {
  cudaMemsetAsync(nmatches, 0, (size_t)B * sizeof(int32_t), st);
  dim3 grid((Q.cap + 127) / 128, B);
  line_nnr_kernel<<<grid, 128, 0, st>>>(Q, knn, nnr, m12, nmatches);
}

// LSDmatcher::computeAngle2D (LSDmatcher.cpp:19-34) on (e - s) float differences held as doubles
__device__ __forceinline__ double angle2d(double ax, double ay, double bx, double by) {
  const double dot = ax * bx + ay * by;
  const double ma = sqrt(ax * ax + ay * ay), mb = sqrt(bx * bx + by * by);
  return fabs(dot / (ma * mb));
}

// LSDmatcher::SearchByGeomNApearance (LSDmatcher.cpp:36-110) after matchNNR: thread per last-frame line.
// Several last lines may pass for one current line; the reference overwrites in i1 order -> atomicMax.
__global__ void line_geom_kernel(LineSet Last, const uint8_t* __restrict__ has_ml, LineSet Cur,
                                 const uint2* __restrict__ knn, float desc_th, float bw, float bh,
                                 int32_t* __restrict__ assign_cur, int32_t* __restrict__ nmatches) {
  const int b = blockIdx.y, i1 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i1 >= Last.n[b] || Cur.n[b] == 0) return;
  if (has_ml && !has_ml[(size_t)b * Last.cap + i1]) return;
  const uint2 k = knn[(size_t)b * Last.cap + i1];
  if (k.y == 0xFFFFFFFFu || !((float)(k.x >> 16) < __fmul_rn((float)(k.y >> 16), desc_th))) return;
  const int i2 = (int)(k.x & 0xFFFFu);
  const psl_keyline& c = Cur.kl[(size_t)b * Cur.cap + i2];
  const psl_keyline& l = Last.kl[(size_t)b * Last.cap + i1];
  if (c.start_x == 0) return;
  const double ang = angle2d((double)__fsub_rn(c.e_oct_x, c.s_oct_x), (double)__fsub_rn(c.e_oct_y, c.s_oct_y),
                             (double)__fsub_rn(l.e_oct_x, l.s_oct_x), (double)__fsub_rn(l.e_oct_y, l.s_oct_y));
  const double cos_th = cos(20.0 / 180.0 * 3.14159265358979323846);
  if (ang < cos_th) return;
  const double dW = (double)bw * 0.1, dH = (double)bh * 0.1;
  const bool far_s = fabs((double)__fsub_rn(c.s_oct_x, l.s_oct_x)) > dW || fabs((double)__fsub_rn(c.s_oct_y, l.s_oct_y)) > dH;
  const bool far_e = fabs((double)__fsub_rn(c.e_oct_x, l.e_oct_x)) > dW || fabs((double)__fsub_rn(c.e_oct_y, l.e_oct_y)) > dH;
  if (far_s && far_e) return;
  atomicMax(assign_cur + (size_t)b * Cur.cap + i2, i1);
  atomicAdd(nmatches + b, 1);
}

void launch_line_geom(const LineSet& Last, const uint8_t* has_ml, const LineSet& Cur, const uint2* knn, float desc_th,
                      float bounds_w, float bounds_h, int32_t* assign_cur, int32_t* nmatches, int B, cudaStream_t st) {
  cudaMemsetAsync(assign_cur, 0xFF, (size_t)B * Cur.cap * sizeof(int32_t), st);
  cudaMemsetAsync(nmatches, 0, (size_t)B * sizeof(int32_t), st);
  dim3 grid((Last.cap + 127) / 128, B);
  line_geom_kernel<<<grid, 128, 0, st>>>(Last, has_ml, Cur, knn, desc_th, bounds_w, bounds_h, assign_cur, nmatches);
}

// LSDmatcher::FrameBFMatch (LSDmatcher.cpp:492-516) with lineDescriptorMAD (:660-685).  Hamming distances are
// integers in [0,256], so every median the reference takes from a sorted copy is a rank query on a 257-bin
// histogram (order-insensitive: only the value at position size/2 is used).  Block per pair.
__device__ int rank_value(const int* hist, int rank) {  // value at position `rank` of the sorted multiset
  int acc = 0;
  for (int v = 0; v <= 256; ++v) {
    acc += hist[v];
    if (acc > rank) return v;
  }
  return 256;
}

__global__ void __launch_bounds__(256)
    line_bfmatch_kernel(LineSet Q, LineSet T, const uint2* __restrict__ knn, float nn_ratio, float th,
                        int32_t* __restrict__ matches) {
  const int b = blockIdx.x, t = threadIdx.x;
  const int n = Q.n[b], nt = T.n[b];
  __shared__ int h1[257], h2[257];
  __shared__ int med12;
  const uint2* kk = knn + (size_t)b * Q.cap;
  int32_t* out = matches + (size_t)b * Q.cap;
  if (n <= 0) return;
  if (nt < 2) {
    for (int i = t; i < n; i += 256) out[i] = -1;
    return;
  }
  for (int v = t; v <= 256; v += 256) { h1[v] = 0; h2[v] = 0; }
  if (t == 0) { h1[256] = 0; h2[256] = 0; }
  __syncthreads();
  for (int i = t; i < n; i += 256) atomicAdd(&h1[(int)(kk[i].y >> 16) - (int)(kk[i].x >> 16)], 1);  // d1 - d0 >= 0
  __syncthreads();
  if (t == 0) med12 = rank_value(h1, n / 2);
  __syncthreads();
  for (int i = t; i < n; i += 256) atomicAdd(&h2[abs((int)(kk[i].y >> 16) - (int)(kk[i].x >> 16) - med12)], 1);
  __syncthreads();
  const double nn12_th = 1.4826 * (double)rank_value(h2, n / 2) * 0.5;
  for (int i = t; i < n; i += 256) {
    const float d0 = (float)(kk[i].x >> 16), d1 = (float)(kk[i].y >> 16);
    const double dist_12 = (double)__fsub_rn(d1, d0);
    out[i] = (dist_12 > nn12_th && d0 < th && d0 < __fmul_rn(nn_ratio, d1)) ? (int)(kk[i].x & 0xFFFFu) : -1;
  }
}

void launch_line_bfmatch(const LineSet& Q, const LineSet& T, const uint2* knn, float nn_ratio, float th,
                         int32_t* matches, int B, cudaStream_t st) {
  line_bfmatch_kernel<<<B, 256, 0, st>>>(Q, T, knn, nn_ratio, th, matches);
}

// LSDmatcher::FrameBFMatchNew + mutualOverlap (LSDmatcher.cpp:518-658): thread per query line.  The best knn match is
// kept if the projections of the query's end points onto the matched line (along their epipolar lines F * p) overlap
// that line's segment by more than 0.8 and the distance passes th and the ratio test.  cv::Mat float arithmetic as
// pinned in DESIGN.md: 3x3 * 3x1 with double accumulation and one rounding, Mat::cross in float, `Mat /= s` =
// convertTo(alpha = 1 / s) in float, cv::norm of a float difference with the squares summed in double.
struct V3f { float v[3]; };
__device__ __forceinline__ V3f matvec3(const float* F, float x, float y) {
  V3f r;
#pragma unroll
  for (int k = 0; k < 3; ++k) r.v[k] = (float)((double)F[3 * k] * x + (double)F[3 * k + 1] * y + (double)F[3 * k + 2] * 1.0);
  return r;
}
__device__ __forceinline__ V3f cross3(const V3f& a, const V3f& b) {
  return V3f{{a.v[1] * b.v[2] - a.v[2] * b.v[1], a.v[2] * b.v[0] - a.v[0] * b.v[2], a.v[0] * b.v[1] - a.v[1] * b.v[0]}};
}
__device__ __forceinline__ double norm_diff(const V3f& a, const V3f& b) {
  const float d0 = a.v[0] - b.v[0], d1 = a.v[1] - b.v[1], d2 = a.v[2] - b.v[2];
  return sqrt((double)d0 * d0 + (double)d1 * d1 + (double)d2 * d2);
}
__device__ float mutual_overlap(const V3f* p) {
  float max_dist = 0.0f;
  int outer1 = 0, outer2 = 3;
  for (int i = 0; i < 3; ++i)
    for (int j = i + 1; j < 4; ++j) {
      const float dist = (float)norm_diff(p[i], p[j]);   // float dist = norm(...)
      if (dist > max_dist) { max_dist = dist; outer1 = i; outer2 = j; }
    }
  if (max_dist < 1.0f) return 0.0f;
  int inner[2], c = 0;
  for (int k = 0; k < 4; ++k)
    if (k != outer1 && k != outer2) inner[c++] = k;
  return (float)(norm_diff(p[inner[0]], p[inner[1]]) / (double)max_dist);   // double norm / float -> float
}

__global__ void line_bfmatch_new_kernel(LineSet Q, LineSet T, const double* __restrict__ funcT, const uint2* __restrict__ knn,
                                        const float* __restrict__ F, float nn_ratio, float th, int32_t* __restrict__ matches) {
  const int b = blockIdx.y, q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q.n[b]) return;
  int32_t* out = matches + (size_t)b * Q.cap;
  out[q] = -1;
  if (T.n[b] < 2) return;  // one train row: knnMatch has one entry and the loop `j < size() - 1` is empty
  const uint2 kk = knn[(size_t)b * Q.cap + q];
  const int t = (int)(kk.x & 0xFFFFu);
  const psl_keyline k1 = Q.kl[(size_t)b * Q.cap + q], k2 = T.kl[(size_t)b * T.cap + t];
  const double* f2 = funcT + ((size_t)b * T.cap + t) * 3;
  const V3f e1 = matvec3(F, k1.start_x, k1.start_y), e2 = matvec3(F, k1.end_x, k1.end_y);
  const V3f l2{{(float)f2[0], (float)f2[1], (float)f2[2]}};
  V3f p1 = cross3(l2, e1), p2 = cross3(l2, e2);
  if (!((double)fabsf(p1.v[2]) > 1e-12 && (double)fabsf(p2.v[2]) > 1e-12)) return;
  const float s1 = (float)(1.0 / (double)p1.v[2]), s2 = (float)(1.0 / (double)p2.v[2]);
#pragma unroll
  for (int k = 0; k < 3; ++k) { p1.v[k] = p1.v[k] * s1; p2.v[k] = p2.v[k] * s2; }
  const V3f pts[4] = {p1, p2, V3f{{k2.start_x, k2.start_y, 1.0f}}, V3f{{k2.end_x, k2.end_y, 1.0f}}};
  const float score = mutual_overlap(pts);
  const float d0 = (float)(kk.x >> 16), d1 = (float)(kk.y >> 16);
  if (d0 < th && (double)score > 0.8 && d0 < nn_ratio * d1) out[q] = t;
}

void launch_line_bfmatch_new(const LineSet& Q, const LineSet& T, const double* funcT, const uint2* knn, const float* F,
                             float nn_ratio, float th, int32_t* matches, int B, cudaStream_t st) {
  dim3 grid((Q.cap + 127) / 128, B);
  line_bfmatch_new_kernel<<<grid, 128, 0, st>>>(Q, T, funcT, knn, F, nn_ratio, th, matches);
}

// mutual consistency of LSDmatcher::SearchDouble (LSDmatcher.cpp:474-487)
__global__ void line_mutual_kernel(LineSet A, const int32_t* __restrict__ m21, int cap2, int32_t* __restrict__ m12,
                                   int32_t* __restrict__ nmatches) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n[b]) return;
  const int j = m12[(size_t)b * A.cap + i];
  if (j < 0) return;
  if (m21[(size_t)b * cap2 + j] != i) m12[(size_t)b * A.cap + i] = -1;
  else atomicAdd(nmatches + b, 1);
}

// LSDmatcher::SearchForTriangulation (LSDmatcher.cpp:721-737, 759-775): mutual check (optional) + MapLine gates
__global__ void line_triang_kernel(LineSet A, const int32_t* __restrict__ m21, int cap2, const uint8_t* __restrict__ ml1,
                                   const uint8_t* __restrict__ ml2, int is_double, int32_t* __restrict__ m12,
                                   int32_t* __restrict__ nmatches) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n[b]) return;
  const int j = m12[(size_t)b * A.cap + i];
  if (j < 0) return;
  const bool drop = (is_double && m21[(size_t)b * cap2 + j] != i) || ml1[(size_t)b * A.cap + i] || ml2[(size_t)b * cap2 + j];
  if (drop) m12[(size_t)b * A.cap + i] = -1;
  else atomicAdd(nmatches + b, 1);
}

// LSDmatcher::Fuse window search (LSDmatcher.cpp:920-953) over KeyFrame::GetLinesInArea (KeyFrame.cc:857-891).
// The float / double mix of the reference is kept: the midpoint offsets are formed and squared in double and their
// sum is narrowed to float before the comparison with r * r; directions are normalised in float.  A degenerate
// (zero-length) direction gives NaN, which fails `CosSita < TH` and therefore passes, as in the reference.
__global__ void __launch_bounds__(128)
    line_fuse_kernel(const psl_keyline* __restrict__ kl, int n_lines, const uint8_t* __restrict__ kf_desc,
                     const psl_line_fuse_query* __restrict__ queries, const uint8_t* __restrict__ qdesc, int nq,
                     float th_cos, int th_low, int32_t* __restrict__ best_idx, int32_t* __restrict__ best_dist) {
  const int q = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (q >= nq) return;
  const psl_line_fuse_query Q = queries[q];
  unsigned best = (256u << 16) | 0xFFFFu;  // distance << 16 | line: the minimum is the earliest smallest distance
  if (Q.flags & PSL_Q_VALID) {
    float d1x = __fsub_rn(Q.u1, Q.u2), d1y = __fsub_rn(Q.v1, Q.v2);
    const float n1 = __fsqrt_rn(__fadd_rn(__fmul_rn(d1x, d1x), __fmul_rn(d1y, d1y)));
    d1x = __fdiv_rn(d1x, n1);
    d1y = __fdiv_rn(d1y, n1);
    const double mx = 0.5 * (double)__fadd_rn(Q.u1, Q.u2), my = 0.5 * (double)__fadd_rn(Q.v1, Q.v2);
    const float r2 = __fmul_rn(Q.radius, Q.radius);
    const uint4* qd = reinterpret_cast<const uint4*>(qdesc) + 2 * (size_t)q;
    const uint4 q0 = __ldg(qd), q1 = __ldg(qd + 1);
    for (int i = lane; i < n_lines; i += 32) {
      const psl_keyline k = kl[i];
      const double ex = mx - (double)k.pt_x, ey = my - (double)k.pt_y;
      const float distance = (float)__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey));
      if (distance > r2) continue;
      float d2x = __fsub_rn(k.start_x, k.end_x), d2y = __fsub_rn(k.start_y, k.end_y);
      const float n2 = __fsqrt_rn(__fadd_rn(__fmul_rn(d2x, d2x), __fmul_rn(d2y, d2y)));
      d2x = __fdiv_rn(d2x, n2);
      d2y = __fdiv_rn(d2y, n2);
      const float cs = fabsf(__fadd_rn(__fmul_rn(d1x, d2x), __fmul_rn(d1y, d2y)));
      if (cs < th_cos) continue;
      if (k.octave < Q.pred_level - 1 || k.octave > Q.pred_level) continue;
      const uint4* d = reinterpret_cast<const uint4*>(kf_desc) + 2 * (size_t)i;
      const unsigned key = ((unsigned)ham256(q0, q1, __ldg(d), __ldg(d + 1)) << 16) | (unsigned)i;
      best = min(best, key);
    }
  }
  best = __reduce_min_sync(0xffffffffu, best);
  if (lane == 0) {
    const int dist = (int)(best >> 16);
    best_idx[q] = (dist <= th_low && (best & 0xFFFFu) != 0xFFFFu) ? (int)(best & 0xFFFFu) : -1;
    if (best_dist) best_dist[q] = dist;
  }
}

void launch_line_fuse(const psl_keyline* kl, int n_lines, const uint8_t* kf_desc, const psl_line_fuse_query* queries,
                      const uint8_t* qdesc, int nq, float th_cos, int th_low, int32_t* best_idx, int32_t* best_dist,
                      cudaStream_t st) {
  if (nq <= 0) return;
  line_fuse_kernel<<<(nq + 3) / 4, 128, 0, st>>>(kl, n_lines, kf_desc, queries, qdesc, nq, th_cos, th_low, best_idx,
                                                 best_dist);
}

void launch_line_triang(const LineSet& A, const int32_t* m21, int cap2, const uint8_t* ml1, const uint8_t* ml2,
                        int is_double, int32_t* m12, int32_t* nmatches, int B, cudaStream_t st) {
  cudaMemsetAsync(nmatches, 0, (size_t)B * sizeof(int32_t), st);
  dim3 grid((A.cap + 127) / 128, B);
  line_triang_kernel<<<grid, 128, 0, st>>>(A, m21, cap2, ml1, ml2, is_double, m12, nmatches);
}

void launch_line_mutual(const LineSet& A, const int32_t* m21, int cap2, int32_t* m12, int32_t* nmatches, int B,
                        cudaStream_t st) {
  cudaMemsetAsync(nmatches, 0, (size_t)B * sizeof(int32_t), st);
  dim3 grid((A.cap + 127) / 128, B);
  line_mutual_kernel<<<grid, 128, 0, st>>>(A, m21, cap2, m12, nmatches);
}

// ---------------------------------------------------------------------------------------------------
// Line grid (Frame::AssignFeaturesToGridForLine, Frame.cc:286-309): the grid cells a line crosses, by the
// reference's Bresenham iterator (add_src/lineIterator.cpp:33-77).  cells[B][cap][kLineCells] (u16 cell id =
// ix * 48 + iy, in-grid cells only), ncell[B][cap].
// ---------------------------------------------------------------------------------------------------
__global__ void line_cells_kernel(LineSet F, float w_inv, float h_inv, uint16_t* __restrict__ cells,
                                  uint8_t* __restrict__ ncell) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F.n[b]) return;
  const psl_keyline& kl = F.kl[(size_t)b * F.cap + i];
  double x1 = (double)__fmul_rn(kl.start_x, w_inv), y1 = (double)__fmul_rn(kl.start_y, h_inv);
  double x2 = (double)__fmul_rn(kl.end_x, w_inv), y2 = (double)__fmul_rn(kl.end_y, h_inv);
  const bool steep = fabs(y2 - y1) > fabs(x2 - x1);
  double t;
  if (steep) { t = x1; x1 = y1; y1 = t; t = x2; x2 = y2; y2 = t; }
  if (x1 > x2) { t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }
  const double dx = x2 - x1, dy = fabs(y2 - y1);
  double error = dx / 2.0;
  const int ystep = (y1 < y2) ? 1 : -1;
  int x = (int)x1, y = (int)y1;
  const int maxX = (int)x2;
  uint16_t* out = cells + ((size_t)b * F.cap + i) * kLineCells;
  int n = 0;
  for (; x <= maxX; ++x) {
    const int px = steep ? y : x, py = steep ? x : y;
    if (px >= 0 && px < PSL_GRID_COLS && py >= 0 && py < PSL_GRID_ROWS && n < kLineCells)
      out[n++] = (uint16_t)(px * PSL_GRID_ROWS + py);
    error -= dy;
    if (error < 0) { y += ystep; error += dx; }
  }
  ncell[(size_t)b * F.cap + i] = (uint8_t)n;
}

// Static part of LSDmatcher::SearchByProjection for one (query, line): Frame::GetFeaturesInAreaForLine
// (Frame.cc:752-826) membership + its enumeration position, the per-mode gates and the Hamming distance.
// key = dist << 32 | (sample * 3072 + first cell) << 12 | line   (ascending key = ascending distance, ties by the
// reference's candidate order: sample, cell column-major, line index); ~0 = not a candidate.
__global__ void __launch_bounds__(128)
    line_proj_keys_kernel(LineSet F, const double* __restrict__ lineeq, const double* __restrict__ lines3d,
                          const psl_line_query* __restrict__ queries, const uint8_t* __restrict__ qdesc,
                          const int32_t* __restrict__ nq, int qcap, const uint16_t* __restrict__ cells,
                          const uint8_t* __restrict__ ncell, float min_x, float min_y, float w_inv, float h_inv, int mode,
                          unsigned long long* __restrict__ keys) {
  const int b = blockIdx.z, q = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = F.n[b];
  if (q >= nq[b] || j >= n) return;
  unsigned long long key = ~0ull;
  const psl_line_query Q = queries[(size_t)b * qcap + q];
  const psl_keyline& kl = F.kl[(size_t)b * F.cap + j];
  if (Q.flags & PSL_Q_VALID) {
    const float TH = mode == 0 ? 0.96f : 0.998f;
    const float x[3] = {Q.x1, (float)((double)__fadd_rn(Q.x1, Q.x2) / 2.0), Q.x2};
    const float y[3] = {Q.y1, (float)((double)__fadd_rn(Q.y1, Q.y2) / 2.0), Q.y2};
    float d1x = __fsub_rn(Q.x1, Q.x2), d1y = __fsub_rn(Q.y1, Q.y2);
    const float n1 = __fsqrt_rn(__fadd_rn(__fmul_rn(d1x, d1x), __fmul_rn(d1y, d1y)));
    d1x = __fdiv_rn(d1x, n1);
    d1y = __fdiv_rn(d1y, n1);
    float d2x = __fsub_rn(kl.start_x, kl.end_x), d2y = __fsub_rn(kl.start_y, kl.end_y);
    const float n2 = __fsqrt_rn(__fadd_rn(__fmul_rn(d2x, d2x), __fmul_rn(d2y, d2y)));
    d2x = __fdiv_rn(d2x, n2);
    d2y = __fdiv_rn(d2y, n2);
    const float cs = fabsf(__fadd_rn(__fmul_rn(d1x, d2x), __fmul_rn(d1y, d2y)));
    int pos = -1;
    if (!(cs < TH)) {
      const double* L = lineeq + ((size_t)b * F.cap + j) * 3;
      const uint16_t* cl = cells + ((size_t)b * F.cap + j) * kLineCells;
      const int nc = ncell[(size_t)b * F.cap + j];
      const float r = Q.radius;
      for (int s = 0; s < 3 && pos < 0; ++s) {
        const int cx0 = max(0, (int)floor((double)__fmul_rn(__fsub_rn(__fsub_rn(x[s], min_x), r), w_inv)));
        if (cx0 >= PSL_GRID_COLS) continue;
        const int cx1 = min(PSL_GRID_COLS - 1, (int)ceil((double)__fmul_rn(__fadd_rn(__fsub_rn(x[s], min_x), r), w_inv)));
        if (cx1 < 0) continue;
        const int cy0 = max(0, (int)floor((double)__fmul_rn(__fsub_rn(__fsub_rn(y[s], min_y), r), h_inv)));
        if (cy0 >= PSL_GRID_ROWS) continue;
        const int cy1 = min(PSL_GRID_ROWS - 1, (int)ceil((double)__fmul_rn(__fadd_rn(__fsub_rn(y[s], min_y), r), h_inv)));
        if (cy1 < 0) continue;
        const float dist = (float)(L[0] * (double)x[s] + L[1] * (double)y[s] + L[2]);
        if (!(fabs((double)dist) < (double)r)) continue;
        int first = 1 << 30;
        for (int c = 0; c < nc; ++c) {
          const int id = cl[c], ix = id / PSL_GRID_ROWS, iy = id - ix * PSL_GRID_ROWS;
          if (ix >= cx0 && ix <= cx1 && iy >= cy0 && iy <= cy1) first = min(first, id);
        }
        if (first < (1 << 30)) pos = s * (PSL_GRID_COLS * PSL_GRID_ROWS) + first;
      }
    }
    if (pos >= 0) {
      bool ok;
      if (mode == 0) {
        const double ang = angle2d((double)__fsub_rn(kl.e_oct_x, kl.s_oct_x), (double)__fsub_rn(kl.e_oct_y, kl.s_oct_y),
                                   (double)__fsub_rn(Q.ex, Q.sx), (double)__fsub_rn(Q.ey, Q.sy));
        const float mx = fmaxf(Q.length, kl.line_length), mn = fminf(Q.length, kl.line_length);
        ok = !(ang < cos(10.0 / 180.0 * 3.14159265358979323846)) && !((double)__fdiv_rn(mn, mx) < 0.75);
      } else {
        const double* p = lines3d + ((size_t)b * F.cap + j) * 6;
        const double vx = p[0] - p[3], vy = p[1] - p[4], vz = p[2] - p[5];
        const float dot = (float)(vx * Q.normal[0] + vy * Q.normal[1] + vz * Q.normal[2]);
        const float mag_f = (float)sqrt(vx * vx + vy * vy + vz * vz);
        const float mag_ml = (float)sqrt(Q.normal[0] * Q.normal[0] + Q.normal[1] * Q.normal[1] + Q.normal[2] * Q.normal[2]);
        const float angle = fabsf(__fdiv_rn(dot, __fmul_rn(mag_f, mag_ml)));
        ok = !((double)angle < cos(15.0 / 180.0 * 3.14159265358979323846));
      }
      if (ok) {
        const uint4* a = reinterpret_cast<const uint4*>(qdesc + ((size_t)b * qcap + q) * 32);
        const uint4* d = reinterpret_cast<const uint4*>(F.desc + ((size_t)b * F.cap + j) * 32);
        const int dist = ham256(__ldg(a), __ldg(a + 1), __ldg(d), __ldg(d + 1));
        key = ((unsigned long long)dist << 32) | ((unsigned long long)pos << 12) | (unsigned long long)j;
      }
    }
  }
  keys[((size_t)b * qcap + q) * F.cap + j] = key;
}

// The order-dependent part (LSDmatcher.cpp:137-210, :267-349): one warp per pair walks the queries in order;
// per query the lanes reduce the two smallest keys among the unclaimed lines.
__global__ void __launch_bounds__(32)
    line_proj_resolve_kernel(LineSet F, const psl_line_query* __restrict__ queries, const int32_t* __restrict__ nq,
                             int qcap, const unsigned long long* __restrict__ keys, const uint8_t* __restrict__ claimed_in,
                             uint8_t* __restrict__ claimed, int mode, float nn_ratio, int32_t* __restrict__ assign,
                             int32_t* __restrict__ nmatches) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const int n = F.n[b];
  uint8_t* cl = claimed + (size_t)b * F.cap;
  int32_t* as = assign + (size_t)b * F.cap;
  for (int i = lane; i < n; i += 32) {
    cl[i] = claimed_in ? claimed_in[(size_t)b * F.cap + i] : 0;
    as[i] = -1;
  }
  __syncwarp();
  int nm = 0;
  const int Nq = nq[b];
  for (int q = 0; q < Nq; ++q) {
    const unsigned long long* kq = keys + ((size_t)b * qcap + q) * F.cap;
    unsigned long long k1 = ~0ull, k2 = ~0ull;
    for (int j = lane; j < n; j += 32) {
      const unsigned long long k = kq[j];
      if (k == ~0ull || cl[j]) continue;
      if (k < k1) { k2 = k1; k1 = k; }
      else if (k < k2) k2 = k;
    }
    const unsigned long long best = wmin64(k1);
    if (best == ~0ull) continue;
    const unsigned long long second = wmin64(k1 == best ? k2 : k1);
    const int bestDist = (int)(best >> 32), bestIdx = (int)(best & 0xFFFu);
    if (bestDist > 95) continue;
    if (mode == 1) {
      const int bestDist2 = second == ~0ull ? 256 : (int)(second >> 32);
      const int lvl = F.kl[(size_t)b * F.cap + bestIdx].octave;
      const int lvl2 = second == ~0ull ? -1 : F.kl[(size_t)b * F.cap + (int)(second & 0xFFFu)].octave;
      if (lvl == lvl2 && (float)bestDist > __fmul_rn(nn_ratio, (float)bestDist2)) continue;
    }
    if (lane == 0) {
      as[bestIdx] = q;
      if (queries[(size_t)b * qcap + q].flags & PSL_Q_CLAIMS) cl[bestIdx] = 1;
    }
    ++nm;
    __syncwarp();
  }
  if (lane == 0) nmatches[b] = nm;
}

void launch_line_projection(const LineSet& F, const double* lineeq, const double* lines3d, const psl_line_query* queries,
                            const uint8_t* qdesc, const int32_t* nq, int qcap, int max_nq, float min_x, float min_y,
                            float w_inv, float h_inv, int mode, float nn_ratio, const uint8_t* claimed_in,
                            uint16_t* cells, uint8_t* ncell, unsigned long long* keys, uint8_t* claimed, int32_t* assign,
                            int32_t* nmatches, int B, cudaStream_t st) {
  dim3 g1((F.cap + 127) / 128, B);
  line_cells_kernel<<<g1, 128, 0, st>>>(F, w_inv, h_inv, cells, ncell);
  if (max_nq > 0) {
    dim3 g2((F.cap + 127) / 128, max_nq, B);
    line_proj_keys_kernel<<<g2, 128, 0, st>>>(F, lineeq, lines3d, queries, qdesc, nq, qcap, cells, ncell, min_x, min_y,
                                              w_inv, h_inv, mode, keys);
  }
  line_proj_resolve_kernel<<<B, 32, 0, st>>>(F, queries, nq, qcap, keys, claimed_in, claimed, mode, nn_ratio, assign,
                                             nmatches);
}

// ---------------------------------------------------------------------------------------------------
// InsectLineMatch::SearchMapInsectline (InsectlineMatch.cpp:9-60; mode 0, thread per structural line) and
// Map::AssociatePlanesByBoundary (Map.cc:204-272; mode 1: the distance threshold is carried from one structural
// line to the next, so a single thread walks them all).
// ---------------------------------------------------------------------------------------------------
__device__ void plane_assoc_one(int i, const float* planes_cam, const double* pts, const float* Tcw,
                                const float* map_planes, const uint8_t* map_bad, int n_map, float a_th, int mode,
                                float& thr, int32_t* assign, int& nm) {
  float pM[4];  // Frame::ComputeWorldPlane (Frame.cc:918-925): Tcw^T * plane, double accumulation, one rounding
  for (int k = 0; k < 4; ++k) {
    double acc = 0;
    for (int r = 0; r < 4; ++r) acc += (double)Tcw[4 * r + k] * (double)planes_cam[4 * i + r];
    pM[k] = (float)acc;
  }
  const double* P = pts + 15 * (size_t)i;
  bool found = false;
  int a = -1;
  for (int m = 0; m < n_map; ++m) {
    if (mode == 0 && map_bad && map_bad[m]) continue;
    float pW[4] = {map_planes[4 * m], map_planes[4 * m + 1], map_planes[4 * m + 2], map_planes[4 * m + 3]};
    if (mode == 1 && pW[3] < 0) { pW[0] = -pW[0]; pW[1] = -pW[1]; pW[2] = -pW[2]; pW[3] = -pW[3]; }
    const float angle = __fadd_rn(__fadd_rn(__fmul_rn(pM[0], pW[0]), __fmul_rn(pM[1], pW[1])), __fmul_rn(pM[2], pW[2]));
    if (angle > a_th || angle < -a_th) {
      float d[5];
      for (int k = 0; k < 5; ++k)
        d[k] = (float)((double)pW[0] * P[3 * k] + (double)pW[1] * P[3 * k + 1] + (double)pW[2] * P[3 * k + 2] + (double)pW[3]);
      const float dis = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(d[0], d[1]), d[2]), d[3]), d[4]), 5.f);
      if (fabsf(dis) < thr) {
        thr = dis;  // signed, as in the reference
        a = m;
        found = true;
        if (mode == 1) ++nm;
      }
    }
  }
  assign[i] = a;
  if (mode == 0 && found) ++nm;
}

__global__ void plane_assoc_kernel(const float* planes_cam, const double* pts, int n_ljl, const float* Tcw,
                                   const float* map_planes, const uint8_t* map_bad, int n_map, float d_th, float a_th,
                                   int mode, int32_t* assign, int32_t* nmatches) {
  if (mode == 0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ljl) return;
    float thr = d_th;
    int nm = 0;
    plane_assoc_one(i, planes_cam, pts, Tcw, map_planes, map_bad, n_map, a_th, 0, thr, assign, nm);
    if (nm) atomicAdd(nmatches, nm);
  } else if (blockIdx.x == 0 && threadIdx.x == 0) {
    float thr = d_th;
    int nm = 0;
    for (int i = 0; i < n_ljl; ++i) plane_assoc_one(i, planes_cam, pts, Tcw, map_planes, map_bad, n_map, a_th, 1, thr, assign, nm);
    *nmatches = nm;
  }
}

// Plane hypotheses from coplanar intersecting line pairs (Frame.cc:512-645, OldPlane :474-487).  One warp: the lanes
// evaluate 32 junctions at a time (everything but the duplicate test is independent), then the candidates are taken
// in junction order and each is compared, 32 kept planes at a time, with the planes kept so far.  The float / double
// mix follows the reference statement by statement; a degenerate pair (parallel directions) yields NaNs that fail
// the span test's `>` and pass on, as there.
__global__ void __launch_bounds__(32)
    plane_hypotheses_kernel(const psl_keyline* __restrict__ kl_un, const float* __restrict__ line_eq,
                            const double* __restrict__ lines3d, const psl_line_junction* __restrict__ js, int nj,
                            double* __restrict__ le_l, float* planes, double* __restrict__ normals,
                            int32_t* __restrict__ junction_of, int cap, int32_t* __restrict__ n_planes,
                            float* kept /* [nj*4] scratch: every kept plane, also beyond cap */) {
  const int lane = threadIdx.x;
  int np = 0;
  for (int base = 0; base < nj; base += 32) {
    const int i = base + lane;
    bool cand = false;
    float pl[4] = {0.f, 0.f, 0.f, 0.f};
    double nd[3] = {0., 0., 0.};
    if (i < nj) {
      const psl_line_junction J = js[i];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const psl_keyline k = kl_un[s == 0 ? J.l1 : J.l2];
        const double ax = (double)k.start_x, ay = (double)k.start_y, bx = (double)k.end_x, by = (double)k.end_y;
        const double c0 = __dsub_rn(ay, by), c1 = __dsub_rn(bx, ax), c2 = __dsub_rn(__dmul_rn(ax, by), __dmul_rn(ay, bx));
        const double nrm = __dsqrt_rn(__dadd_rn(__dmul_rn(c0, c0), __dmul_rn(c1, c1)));
        le_l[6 * i + 3 * s] = __ddiv_rn(c0, nrm);
        le_l[6 * i + 3 * s + 1] = __ddiv_rn(c1, nrm);
        le_l[6 * i + 3 * s + 2] = __ddiv_rn(c2, nrm);
      }
      const float* e1 = line_eq + 3 * J.l1;
      const float* e2 = line_eq + 3 * J.l2;
      const double* L1 = lines3d + 6 * J.l1;
      const double* L2 = lines3d + 6 * J.l2;
      auto zero3 = [](const double* v) { return fabs(v[0]) <= 1e-12 && fabs(v[1]) <= 1e-12 && fabs(v[2]) <= 1e-12; };
      const bool skip = (e1[0] == 0.f && e1[1] == 0.f && e1[2] == 0.f) || (e2[0] == 0.f && e2[1] == 0.f && e2[2] == 0.f) ||
                        (zero3(L1) && zero3(L1 + 3)) || (zero3(L2) && zero3(L2 + 3));
      if (!skip) {
        float pn0 = __fsub_rn(__fmul_rn(e1[1], e2[2]), __fmul_rn(e1[2], e2[1]));
        float pn1 = __fsub_rn(__fmul_rn(e1[2], e2[0]), __fmul_rn(e1[0], e2[2]));
        float pn2 = __fsub_rn(__fmul_rn(e1[0], e2[1]), __fmul_rn(e1[1], e2[0]));
        const float norm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(pn0, pn0), __fmul_rn(pn1, pn1)), __fmul_rn(pn2, pn2)));
        pn0 = __fdiv_rn(pn0, norm);
        pn1 = __fdiv_rn(pn1, norm);
        pn2 = __fdiv_rn(pn2, norm);
        nd[0] = (double)pn0; nd[1] = (double)pn1; nd[2] = (double)pn2;
        const double* P[5] = {L1, L1 + 3, L2, L2 + 3, J.cross3d};
        float d[5];
#pragma unroll
        for (int k = 0; k < 5; ++k)
          d[k] = (float)__dadd_rn(__dadd_rn(__dmul_rn(nd[0], P[k][0]), __dmul_rn(nd[1], P[k][1])), __dmul_rn(nd[2], P[k][2]));
        float dmin = 10000.f, dmax = -10000.f;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          dmin = dmin < d[k] ? dmin : d[k];
          dmax = dmax > d[k] ? dmax : d[k];
        }
        if (!((double)__fsub_rn(dmax, dmin) > 0.05)) {
          const float sum = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(d[0], d[1]), d[2]), d[3]), d[4]);
          pl[0] = pn0; pl[1] = pn1; pl[2] = pn2;
          pl[3] = __fdiv_rn(-sum, 5.f);
          if (pl[3] < 0.f) {
#pragma unroll
            for (int k = 0; k < 4; ++k) pl[k] = -pl[k];
#pragma unroll
            for (int k = 0; k < 3; ++k) nd[k] = -nd[k];
          }
          cand = true;
        }
      }
    }
    unsigned todo = __ballot_sync(0xffffffffu, cand);
    while (todo) {
      const int j = __ffs(todo) - 1;
      todo &= todo - 1;
      float q[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) q[k] = __shfl_sync(0xffffffffu, pl[k], j);
      bool old = false;
      for (int m0 = 0; m0 < np; m0 += 32) {  // OldPlane: any kept plane close in distance and direction
        const int m = m0 + lane;
        bool hit = false;
        if (m < np) {
          const float dd = __fsub_rn(q[3], kept[4 * m + 3]);
          const float angle = __fadd_rn(__fadd_rn(__fmul_rn(q[0], kept[4 * m]), __fmul_rn(q[1], kept[4 * m + 1])),
                                        __fmul_rn(q[2], kept[4 * m + 2]));
          const bool far = (double)dd > 0.2 || (double)dd < -0.2;
          const bool oblique = (double)angle < 0.9397 && (double)angle > -0.9397;
          hit = !far && !oblique;
        }
        if (__any_sync(0xffffffffu, hit)) { old = true; break; }
      }
      if (old) continue;
      if (lane == j) {
#pragma unroll
        for (int k = 0; k < 4; ++k) kept[4 * np + k] = pl[k];
        if (np < cap) {
#pragma unroll
          for (int k = 0; k < 4; ++k) planes[4 * np + k] = pl[k];
#pragma unroll
          for (int k = 0; k < 3; ++k) normals[3 * np + k] = nd[k];
          junction_of[np] = base + j;
        }
      }
      ++np;
      __syncwarp();
    }
  }
  if (lane == 0) *n_planes = np;
}

void launch_plane_hypotheses(const psl_keyline* kl_un, const float* line_eq, const double* lines3d,
                             const psl_line_junction* js, int nj, double* le_l, float* planes, double* normals,
                             int32_t* junction_of, int cap, int32_t* n_planes, float* kept, cudaStream_t st) {
  plane_hypotheses_kernel<<<1, 32, 0, st>>>(kl_un, line_eq, lines3d, js, nj, le_l, planes, normals, junction_of, cap,
                                            n_planes, kept);
}

void launch_plane_assoc(const float* planes_cam, const double* pts, int n_ljl, const float* Tcw, const float* map_planes,
                        const uint8_t* map_bad, int n_map, float d_th, float a_th, int mode, int32_t* assign,
                        int32_t* nmatches, cudaStream_t st) {
  cudaMemsetAsync(nmatches, 0, sizeof(int32_t), st);
  if (n_ljl <= 0) return;
  plane_assoc_kernel<<<(n_ljl + 63) / 64, 64, 0, st>>>(planes_cam, pts, n_ljl, Tcw, map_planes, map_bad, n_map, d_th, a_th,
                                                       mode, assign, nmatches);
}

}  // namespace psl
