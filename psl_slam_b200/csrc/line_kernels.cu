// Line front end after LSD: the per-frame merge / KeyLine stage (line_core.cuh) and the LBD descriptor
// (opencv_contrib BinaryDescriptor::compute as called at add_src/LineExtractor.cpp:349-350; source-level spec
// vendored at Thirdparty/line_descriptor/src/binary_descriptor_custom.cpp:219-261, 351-413, 1027-1373).
#include <math.h>

#include "line_kernels.cuh"
#include "orb_kernels.cuh"

namespace psl {

// ---------------------------------------------------------------------------------------------------
// clamp + optimizeAndMergeLines_lsd + top-N + KeyLines + line equations, one frame per warp.
// line_core.cuh holds the scalar statement of this stage (what the CPU suite checks against the oracle);
// here the warp runs the same decisions with the data-parallel parts spread over the lanes:
//   * the angle-sorted pair scan of MergeLines (uselongline.cpp:61-151) tests 32 partners of one segment at a
//     time; the early `break` is the first lane whose angle gap exceeds the threshold, neighbour lists are
//     appended in lane (= scan) order;
//   * connected components / sub-clusters (:153-229) stay on lane 0 (index bookkeeping only) and emit the
//     list of sub-cluster heads: every output line is fold(MergeTwoLines, head, neighbours(head)), so the
//     fp64-heavy folds (:231-262) run one sub-cluster per lane;
//   * stable index sorts are rank computations (position = number of elements that sort before).
// ---------------------------------------------------------------------------------------------------
#ifdef PSL_LSD_STATS
__device__ unsigned long long g_post_stats[16];
#define POST_T0(t) long long t = clock64()
#define POST_T1(i, t) do { const unsigned long long v__ = (unsigned long long)(clock64() - t); if (lane == 0) atomicAdd(&g_post_stats[i], v__); } while (0)
extern "C" void psl_post_stats(unsigned long long* out) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, g_post_stats, sizeof(g_post_stats));
  unsigned long long z[16] = {0};
  cudaMemcpyToSymbol(g_post_stats, z, sizeof(z));
}
#else
#define POST_T0(t) do { } while (0)
#define POST_T1(i, t) do { } while (0)
#endif

namespace linew {

using line::kNbCap;
using line::MergeScratch;
using line::Seg;
using line::ScanRec;
constexpr unsigned kFull = 0xffffffffu;

// stable rank sort of 0..n-1 by key ascending (desc = false) or descending; out[rank] = index.  Every caller sorts
// the identity permutation, so the keys are read in place, four per load (the same address for all lanes: one
// broadcast transaction); an element's rank = the keys that come before it + the equal ones with a smaller index.
__device__ void rank_sort_quadratic(uint16_t* out, int n, const float* key, bool desc, int lane) {
  const bool vec = (reinterpret_cast<uintptr_t>(key) & 15) == 0;
  for (int a = lane; a < n; a += 32) {
    const float ka = key[a];
    int r = 0;
    auto count = [&](float kb, int b) {
      const bool before = desc ? (kb > ka) : (kb < ka);
      r += (before || (!(desc ? (ka > kb) : (ka < kb)) && b < a)) ? 1 : 0;
    };
    int b = 0;
    if (vec)
      for (; b + 4 <= n; b += 4) {
        const float4 k4 = *reinterpret_cast<const float4*>(key + b);
        count(k4.x, b);
        count(k4.y, b + 1);
        count(k4.z, b + 2);
        count(k4.w, b + 3);
      }
    for (; b < n; ++b) count(key[b], b);
    out[r] = (uint16_t)a;
  }
  __syncwarp();
}

constexpr int kSortBuckets = 1024;

// The same ranks through a monotone bucketing of the keys: an element's rank = the elements in earlier buckets + the
// members of its own bucket that sort before it, so the quadratic part runs inside a bucket only.  cnt: kSortBuckets
// counters of this warp in shared memory; bkt / members: u16 [n] scratch.
__device__ void rank_sort(uint16_t* bkt, uint16_t* out, int n, const float* key, bool desc, int lane, uint32_t* cnt,
                          uint16_t* members) {
  float kmin = INFINITY, kmax = -INFINITY;
  bool bad = false;
  for (int a = lane; a < n; a += 32) {
    const float k = key[a];
    bad |= !(k == k) || fabsf(k) == INFINITY;
    kmin = fminf(kmin, k);
    kmax = fmaxf(kmax, k);
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    kmin = fminf(kmin, __shfl_xor_sync(kFull, kmin, d));
    kmax = fmaxf(kmax, __shfl_xor_sync(kFull, kmax, d));
  }
  if (n < 128 || __any_sync(kFull, bad) || !(kmax > kmin)) {
    rank_sort_quadratic(out, n, key, desc, lane);
    return;
  }
  // monotone: subtraction of a constant, multiplication by a positive constant and truncation keep the order
  const float scale = __fdiv_rn((float)(kSortBuckets - 1), __fsub_rn(kmax, kmin));
  for (int b = lane; b < kSortBuckets; b += 32) cnt[b] = 0;
  __syncwarp();
  for (int a = lane; a < n; a += 32) {
    int b = (int)__fmul_rn(__fsub_rn(key[a], kmin), scale);
    b = min(max(b, 0), kSortBuckets - 1);
    if (desc) b = kSortBuckets - 1 - b;
    bkt[a] = (uint16_t)b;
    atomicAdd(&cnt[b], 1u);
  }
  __syncwarp();
  {  // exclusive prefix over the buckets: 32 consecutive counters per lane
    uint32_t local = 0;
    for (int k = 0; k < kSortBuckets / 32; ++k) local += cnt[lane * (kSortBuckets / 32) + k];
    uint32_t inc = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(kFull, inc, d);
      if (lane >= d) inc += o;
    }
    uint32_t run = inc - local;
    for (int k = 0; k < kSortBuckets / 32; ++k) {
      const uint32_t c = cnt[lane * (kSortBuckets / 32) + k];
      cnt[lane * (kSortBuckets / 32) + k] = run;
      run += c;
    }
  }
  __syncwarp();
  for (int a = lane; a < n; a += 32) members[atomicAdd(&cnt[bkt[a]], 1u)] = (uint16_t)a;   // cnt[b] becomes the bucket's end
  __syncwarp();
  for (int a = lane; a < n; a += 32) {
    const int b = bkt[a];
    const int s0 = b ? (int)cnt[b - 1] : 0, e0 = (int)cnt[b];
    const float ka = key[a];
    int r = s0;
    for (int m = s0; m < e0; ++m) {
      const int o = members[m];
      const float kb = key[o];
      const bool before = desc ? (kb > ka) : (kb < ka);
      r += (before || (!(desc ? (ka > kb) : (ka < kb)) && o < a)) ? 1 : 0;
    }
    out[r] = (uint16_t)a;
  }
  __syncwarp();
}

// line::point_line_distance(l, x0, y0) > thr with the line's denominator sqrt(a^2 + b^2) taken from the per-line table.
// The reference compares (float)((double)num / den) with thr; the fp64 division is about forty instructions and the
// outcome is obvious for all but a sliver of the pairs, so the quotient is only formed when num is within 0.2 % of
// thr * den (tden: that product rounded to fp32, from the line's scan record) or tden is not a positive finite number.
__device__ __forceinline__ bool pld_exceeds(const Seg& l, double den, float tden, float x0, float y0, float thr) {
  const float x1 = l.v[0], y1 = l.v[1], x2 = l.v[2], y2 = l.v[3];
  const float num = fabsf(__fadd_rn(__fadd_rn(__fmul_rn(__fsub_rn(y2, y1), x0), __fmul_rn(__fsub_rn(x1, x2), y0)),
                                    __fsub_rn(__fmul_rn(x2, y1), __fmul_rn(x1, y2))));
  if (tden > 0.f && tden < 1e30f) {
    if (num > tden * 1.002f) return true;
    if (num < tden * 0.998f) return false;
  }
  return (float)((double)num / den) > thr;
}

// AngleDiff(a1, a2) > thr (uselongline.cpp:17-22: min(|a2 - a1|, pi + min - max), the second term through fp64) for
// a threshold below 0.14 rad: the fp64 term is pi - |a2 - a1| up to rounding, so it only decides when the plain
// difference is within 0.14 of pi.
__device__ __forceinline__ bool angle_gap_exceeds(float a1, float a2, float thr) {
  const float c1 = fabsf(__fsub_rn(a2, a1));
  if (c1 <= thr) return false;
  if (c1 < 3.0f && thr < 0.14f) return true;
  return line::angle_diff(a1, a2) > thr;
}

// MergeLines (uselongline.cpp:24-264) in three phases, so that the pair scan -- two thirds of the work, independent per
// row -- can run as a kernel of its own over (frame, block of 32 rows) instead of inside the one warp that owns the frame:
//   merge_prepare    angles, angle order, the per-line scan records                       (one warp per frame)
//   merge_scan_rows  the pair tests of 32 rows of the angle order                          (one warp per 32 rows)
//   merge_finish     neighbour lists in order, clusters, sub-clusters, folded merges      (one warp per frame)
__device__ void merge_prepare(const Seg* src, int n, float distance_thr, MergeScratch& S, int lane) {
  if (n <= 0) return;
  POST_T0(t_sort);
  for (int i = lane; i < n; i += 32) {
    const float dx = __fsub_rn(src[i].v[2], src[i].v[0]), dy = __fsub_rn(src[i].v[3], src[i].v[1]);
    S.angles[i] = (float)atan((double)__fdiv_rn(dy, dx));
    S.length[i] = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    S.tmp16[i] = (uint16_t)i;
    S.nb_cnt[i] = 0;
    S.code[i] = -1;
  }
  __syncwarp();
  rank_sort(S.tmp16, S.order, n, S.angles, false, lane, S.sort_cnt, S.check);
  // segments and angles in scan order, so that the pair scan reads them with plain coalesced loads instead of a
  // chain of dependent gathers (dst is free until the folds at the end)
  // one 32-byte record per line: segment, angle, denominator of PointLineDistance (it depends on the line only,
  // uselongline.cpp:5-15) — two 16-byte loads per partner off one pointer
  ScanRec* srec = S.scan;
  for (int j = lane; j < n; j += 32) {
    const int id = S.order[j];
    const Seg sj = src[id];
    const double a = (double)__fsub_rn(sj.v[3], sj.v[1]), b = (double)__fsub_rn(sj.v[0], sj.v[2]);
    ScanRec r;
    r.s = sj;
    r.angle = S.angles[id];
    r.den = sqrt(a * a + b * b);
    r.tden = (float)((double)distance_thr * r.den);
    srec[j] = r;
    S.sangles[j] = r.angle;   // (the bisection over a vertical line's far partners reads the angles alone)
  }
  __syncwarp();
  int* bcnt0 = reinterpret_cast<int*>(S.angles);   // the unsorted angles are dead from here on
  for (int j = lane; j < n; j += 32) {
    bcnt0[j] = 0;
    S.loc[S.order[j]] = (uint16_t)j;              // rank of a line in the angle order
  }
  __syncwarp();
  POST_T1(0, t_sort);
}

// rows [base, base + 32) of the pair scan; returns whether a neighbour list overflowed
__device__ int merge_scan_rows(int base, int n, float angle_thr, float distance_thr, float endpoint_threshold,
                               MergeScratch& S, int lane) {
  const ScanRec* srec = S.scan;
  // Pair scan (:75-150).  The scalar loop walks the rows i in angle order and, for each, its partners j > i up to
  // the first one whose angle gap is too large; every accepted pair appends each line to the other's neighbour list.
  // Row i's list therefore ends up as: the rows before i that accepted i (ascending), then i's own partners
  // (ascending) — i.e. all its neighbours in angle order, whatever the order of evaluation.  So the rows are
  // independent: one row per lane (32 rows at a time, coalesced loads of the sorted copies, every lane busy with
  // a partner that is inside its window), own partners into a forward list, the other side counted with an atomic
  // and put in order afterwards.
  const float gap_sq_thr = __fmul_rn(endpoint_threshold, endpoint_threshold);
  const float quarter_turn = (float)(line::kPi / 4.0);
  int* bcnt = reinterpret_cast<int*>(S.angles);   // back counters (merge_prepare zeroed them)
  int ovf = 0;
  {
    const int i = base + lane;
    if (i < n) {
      const int idx1 = S.order[i];
      const ScanRec r1 = srec[i];
      const Seg s1 = r1.s;
      const float angle1 = r1.angle;
      const bool horiz = fabsf(angle1) < quarter_turn;
      const line::AxisSeg p = line::along_axis(s1, horiz);
      const bool can_break = (double)fabsf(angle1) < (line::kPi / 2 - (double)angle_thr);
      // midpoints: the reference's 0.5 * (x1 + x2) on floats; halving is exact, so the fp64 detour is not needed
      const float mx1 = __fmul_rn(0.5f, __fadd_rn(s1.v[0], s1.v[2])), my1 = __fmul_rn(0.5f, __fadd_rn(s1.v[1], s1.v[3]));
      const double den1 = r1.den;
      const float tden1 = r1.tden;
      int fc = 0;
      // The partners are read one iteration ahead (angle, segment, denominator of the next j are requested before the
      // current one is tested): the working set of the resident frames does not fit L1, so every partner is an L2 round
      // trip that would otherwise sit in the loop's dependency chain.
      int j = i + 1;
      ScanRec r_nx = r1;
      if (j < n) r_nx = srec[j];
      while (j < n) {
        const float angle2 = r_nx.angle;
        const Seg s2 = r_nx.s;
        const double den2 = r_nx.den;
        const float tden2 = r_nx.tden;
        const int jc = j;
        ++j;
        if (j < n) r_nx = srec[j];
        if (angle_gap_exceeds(angle1, angle2, angle_thr)) {
          if (can_break) break;   // the scalar loop stops at the first partner whose angle gap is too large
          // A near-vertical segment `continue`s past such partners instead.  In the angle-sorted order they form one
          // contiguous run: |a2 - a1| grows with j, pi + a1 - a2 (the wrap-around branch of AngleDiff) shrinks, so
          // only the partners right after i and the ones at the far end of the list (the other vertical direction)
          // can pass.  Skip the run: first j whose gap is small again, by bisection on the same fp32 predicate.
          int lo = jc + 1, hi = n;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (angle_gap_exceeds(angle1, S.sangles[mid], angle_thr)) lo = mid + 1;
            else hi = mid;
          }
          j = lo;
          if (j < n) r_nx = srec[j];
          continue;
        }
        const float mx2 = __fmul_rn(0.5f, __fadd_rn(s2.v[0], s2.v[2])), my2 = __fmul_rn(0.5f, __fadd_rn(s2.v[1], s2.v[3]));
        if (pld_exceeds(s2, den2, tden2, mx1, my1, distance_thr) && pld_exceeds(s1, den1, tden1, mx2, my2, distance_thr)) continue;
        if (!line::ends_meet(p, line::along_axis(s2, horiz), horiz, gap_sq_thr)) continue;
        const int idx2 = S.order[jc];
        const int b = atomicAdd(&bcnt[idx2], 1);   // slot in idx2's list of earlier rows (unordered until the pass below)
        if (fc < kNbCap && b < kNbCap) {
          S.fw[idx1 * kNbCap + fc] = (uint16_t)idx2;
          S.nb[idx2 * kNbCap + b] = (uint16_t)idx1;
        } else {
          ovf = 1;
        }
        ++fc;
      }
      S.nb_cnt[idx1] = (uint16_t)min(fc, kNbCap);
    }
  }
  return __any_sync(kFull, ovf) ? 1 : 0;
}

// returns the number of lines written to dst
__device__ int merge_finish(const Seg* src, int n, Seg* dst, MergeScratch& S, int lane) {
  if (n <= 0) return 0;
  POST_T0(t_scan);
  int* bcnt = reinterpret_cast<int*>(S.angles);
  int ovf = 0;
  // neighbour list of a line = the earlier rows that accepted it, in angle order, then its own partners
  for (int x = lane; x < n; x += 32) {
    const int bc = min(bcnt[x], kNbCap), fc = S.nb_cnt[x];
    uint16_t* nbx = S.nb + x * kNbCap;
    for (int a = 1; a < bc; ++a) {   // insertion sort by rank (the lists hold a handful of entries)
      const uint16_t v = nbx[a];
      const uint16_t rv = S.loc[v];
      int c = a - 1;
      while (c >= 0 && S.loc[nbx[c]] > rv) { nbx[c + 1] = nbx[c]; --c; }
      nbx[c + 1] = v;
    }
    int tot = bc + fc;
    if (tot > kNbCap) { ovf = 1; tot = kNbCap; }
    for (int k = 0; bc + k < tot; ++k) nbx[bc + k] = S.fw[x * kNbCap + k];
    S.nb_cnt[x] = (uint16_t)tot;
  }
  if (ovf) S.overflow = 1;
  __syncwarp();
  S.overflow = __any_sync(kFull, S.overflow) ? 1 : 0;
  POST_T1(1, t_scan);
  POST_T0(t_bfs);
  // connected components (:153-190) and sub-clusters (:193-229): heads of the output lines, in output order
  uint16_t* heads = S.order;  // the angle order is no longer needed
  int nd = 0;
  // The scalar loop (one cluster at a time, breadth first; a round's frontier is the std::set of the uncoded neighbours of
  // the previous one, i.e. ascending) with every inner loop spread over the lanes: the members of a round are appended
  // in frontier order through a ballot prefix, their neighbours are flagged by all lanes, and the flagged range is
  // swept 32 entries at a time into the next frontier.  Flags set for lines that join in the same round are dropped by
  // the sweep exactly as the scalar code drops them.
  const unsigned lt = (1u << lane) - 1u;
  float* key2 = S.sangles;   // the scan-order angles are dead: sort keys of a cluster's members
  uint16_t* sorted = reinterpret_cast<uint16_t*>(S.angles);   // (so are the unsorted angles / back counters)
  for (int i0 = 0; i0 < n; i0 += 32) {
    for (;;) {
      const int ii = i0 + lane;
      const unsigned open = __ballot_sync(kFull, ii < n && S.code[ii] < 0);
      if (!open) break;
      const int i = i0 + __ffs(open) - 1;
      uint16_t* cl = S.tmp16;
      int cs = 1, ncheck = S.nb_cnt[i];
      if (lane == 0) { S.code[i] = 1; cl[0] = (uint16_t)i; }
      for (int k = lane; k < ncheck; k += 32) S.check[k] = S.nb[i * kNbCap + k];
      __syncwarp();
      while (ncheck > 0) {
        // (a) the frontier joins the cluster, in frontier order
        for (int c0 = 0; c0 < ncheck; c0 += 32) {
          const int c = c0 + lane;
          const int j = c < ncheck ? S.check[c] : 0;
          const bool join = c < ncheck && S.code[j] < 0;
          const unsigned bal = __ballot_sync(kFull, join);
          if (join) { S.code[j] = 1; cl[cs + __popc(bal & lt)] = (uint16_t)j; }
          cs += __popc(bal);
        }
        __syncwarp();
        // (b) flag the uncoded neighbours of the frontier
        int lo = n, hi = -1;
        for (int c = lane; c < ncheck; c += 32) {
          const int j = S.check[c];
          const int cnt = S.nb_cnt[j];
          for (int k = 0; k < cnt; ++k) {
            const int q = S.nb[j * kNbCap + k];
            if (S.code[q] < 0) { S.flag[q] = 1; lo = q < lo ? q : lo; hi = q > hi ? q : hi; }
          }
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) {
          lo = min(lo, __shfl_xor_sync(kFull, lo, d));
          hi = max(hi, __shfl_xor_sync(kFull, hi, d));
        }
        __syncwarp();
        // (c) the flagged lines, ascending, are the next frontier
        ncheck = 0;
        for (int q0 = lo; q0 <= hi; q0 += 32) {
          const int q = q0 + lane;
          const bool f = q <= hi && S.flag[q] != 0;
          if (f) S.flag[q] = 0;
          const unsigned bal = __ballot_sync(kFull, f);   // (every flagged line is still uncoded: (a) ran before (b))
          if (f) S.check[ncheck + __popc(bal & lt)] = (uint16_t)q;
          ncheck += __popc(bal);
        }
        __syncwarp();
      }
      if (cs <= 2) {  // fold(cluster) == fold(head, neighbours(head)): the only possible neighbour is the other member
        if (lane == 0) heads[nd] = cl[0];
        ++nd;
        continue;
      }
      // sub-clusters (:193-229): members by length, descending and stable; every member not yet covered heads a
      // sub-cluster and covers its neighbours (inherently sequential, a handful of steps)
      for (int k = lane; k < cs; k += 32) key2[k] = S.length[cl[k]];
      __syncwarp();
      rank_sort_quadratic(S.check, cs, key2, true, lane);   // check[r] = position in cl of the r-th longest
      for (int k = lane; k < cs; k += 32) sorted[k] = cl[S.check[k]];
      __syncwarp();
      for (int k = lane; k < cs; k += 32) { S.loc[sorted[k]] = (uint16_t)k; S.flag[k] = 0; }
      __syncwarp();
      if (lane == 0) {
        for (int j = 0; j < cs; ++j) {
          if (S.flag[j]) continue;
          const int li = sorted[j];
          for (int k = 0; k < S.nb_cnt[li]; ++k) S.flag[S.loc[S.nb[li * kNbCap + k]]] = 1;
          heads[nd++] = (uint16_t)li;
        }
      }
      nd = __shfl_sync(kFull, nd, 0);
      __syncwarp();
      for (int k = lane; k < cs; k += 32) S.flag[k] = 0;
      __syncwarp();
    }
  }
  __syncwarp();
  POST_T1(2, t_bfs);
  POST_T0(t_fold);
  for (int o = lane; o < nd; o += 32) {  // folded MergeTwoLines (:243-255), one sub-cluster per lane
    const int li = heads[o];
    Seg nl = line::merge_two(src[li], src[li]);
    for (int k = 0; k < S.nb_cnt[li]; ++k) nl = line::merge_two(nl, src[S.nb[li * kNbCap + k]]);
    dst[o] = nl;
  }
  __syncwarp();
  POST_T1(3, t_fold);
  return nd;
}

// FilterShortLines (:338-351): order-preserving compaction in place
__device__ int filter_short(Seg* lines, int n, float length_thr, int lane) {
  const float thr2 = __fmul_rn(length_thr, length_thr);
  int m = 0;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    Seg s{};
    bool keep = false;
    if (i < n) {
      s = lines[i];
      const float dx = __fsub_rn(s.v[2], s.v[0]), dy = __fsub_rn(s.v[3], s.v[1]);
      keep = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) > thr2;
    }
    const unsigned bal = __ballot_sync(kFull, keep);
    __syncwarp();
    if (keep) lines[m + __popc(bal & ((1u << lane) - 1u))] = s;  // m + rank <= i: never overtakes unread input
    m += __popc(bal);
    __syncwarp();
  }
  return m;
}

// thresholds of the two MergeLines passes (uselongline.cpp:458, :464): angle, distance, end-point gap
__device__ __forceinline__ void merge_pass_thresholds(int pass, float& angle_thr, float& distance_thr, float& endpoint_thr) {
  angle_thr = pass == 0 ? 0.05f : 0.03f;
  distance_thr = pass == 0 ? 5.f : 3.f;
  endpoint_thr = pass == 0 ? 15.f : 30.f;
}

// last part of a frame: FilterShortLines of the second pass, top-N by response, key lines and line equations
__device__ int frame_keylines(Seg* t2, int n2, int w, int h, int nfeatures, MergeScratch& S, psl_keyline* kl,
                              double* lineeq, int kl_cap, int lane) {
  n2 = filter_short(t2, n2, 50.f, lane);
  int n = n2;
  if (n2 > nfeatures) {  // LineExtractor.cpp:342-348 (stable by response, descending)
    for (int i = lane; i < n2; i += 32) {
      const double ddx = (double)__fsub_rn(t2[i].v[0], t2[i].v[2]), ddy = (double)__fsub_rn(t2[i].v[1], t2[i].v[3]);
      S.length[i] = __fdiv_rn((float)sqrt(ddx * ddx + ddy * ddy), (float)(w > h ? w : h));  // response
      S.tmp16[i] = (uint16_t)i;
    }
    __syncwarp();
    rank_sort(S.tmp16, S.order, n2, S.length, true, lane, S.sort_cnt, S.check);
    n = nfeatures;
  } else {
    for (int i = lane; i < n2; i += 32) S.order[i] = (uint16_t)i;
    __syncwarp();
  }
  if (n > kl_cap) return -1;
  for (int i = lane; i < n; i += 32) {
    psl_keyline k;
    line::make_keyline(t2[S.order[i]], n2 > nfeatures ? i : (int)S.order[i], w, h, k);
    kl[i] = k;
    // sp x ep normalised by its first two components (LineExtractor.cpp:352-363), fp64
    const double sx = k.start_x, sy = k.start_y, ex = k.end_x, ey = k.end_y;
    const double l0 = sy * 1.0 - 1.0 * ey, l1 = 1.0 * ex - sx * 1.0, l2 = sx * ey - sy * ex;
    const double nrm = sqrt(l0 * l0 + l1 * l1);
    lineeq[3 * i] = l0 / nrm; lineeq[3 * i + 1] = l1 / nrm; lineeq[3 * i + 2] = l2 / nrm;
  }
  return n;
}

}  // namespace linew

constexpr int kPostWarps = 4;  // frames per CTA (one per warp)

__device__ __forceinline__ line::MergeScratch post_scratch(const LineBuffers& L, int b, uint32_t* sort_cnt) {
  const size_t o = (size_t)b * (size_t)L.raw_cap;
  return line::MergeScratch{L.raw_cap,      L.m_angles + o, L.m_length + o, L.m_order + o, L.m_tmp16 + o, L.m_nb + o * line::kNbCap,
                            L.m_nb_cnt + o, L.m_code + o,   L.m_check + o,  L.m_loc + o,   L.m_flag + o,  0, L.m_sangles + o,
                            L.m_fw + o * line::kNbCap, sort_cnt, L.m_scan + o};
}

// The merge stage of a batch is five launches: [A] clamp + prepare pass 1, [scan 1], [B] finish pass 1 + filter +
// prepare pass 2, [scan 2], [C] finish pass 2 + filter + top-N + key lines.  A, B, C run one warp per frame; the scans
// run one warp per 32 rows of a frame's angle order.  L.m_cnt[b] = {lines of the current pass, overflow flags}.
template <int STAGE>
__global__ void __launch_bounds__(kPostWarps * 32, 8)
    line_post_kernel(LineBuffers L, int nb, int nfeatures, psl_keyline* __restrict__ kl, double* __restrict__ lineeq,
                     int cap, int32_t* __restrict__ n_out, uint32_t* __restrict__ status) {
  __shared__ uint32_t sort_cnt[kPostWarps][linew::kSortBuckets];
  const int b = blockIdx.x * kPostWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= nb) return;
  const size_t o = (size_t)b * (size_t)L.raw_cap;
  line::MergeScratch S = post_scratch(L, b, sort_cnt[threadIdx.x >> 5]);
  line::Seg* raw = reinterpret_cast<line::Seg*>(L.raw) + o;
  line::Seg *t1 = L.t1 + o, *t2 = L.t2 + o;
  int32_t* cnt = L.m_cnt + 2 * (size_t)b;
  POST_T0(t_all);
  if (STAGE == 0) {
    const int n_raw = L.n_raw[b];
    for (int i = lane; i < n_raw; i += 32) line::clamp_segment(raw[i], L.w, L.h);
    __syncwarp();
    linew::merge_prepare(raw, n_raw, 5.f, S, lane);
    if (lane == 0) { cnt[0] = n_raw; cnt[1] = 0; }
  } else if (STAGE == 1) {
    int n1 = linew::merge_finish(raw, cnt[0], t1, S, lane);
    n1 = linew::filter_short(t1, n1, 30.f, lane);
    linew::merge_prepare(t1, n1, 3.f, S, lane);
    __syncwarp();
    if (lane == 0) { cnt[0] = n1; if (S.overflow) cnt[1] = 1; }
  } else {
    const int n2 = linew::merge_finish(t1, cnt[0], t2, S, lane);
    const int n = linew::frame_keylines(t2, n2, L.w, L.h, nfeatures, S, kl + (size_t)b * cap, lineeq + (size_t)b * cap * 3,
                                        cap, lane);
    if (lane == 0) {
      if (S.overflow || cnt[1]) { atomicOr(status, kStatLineNeighbours); atomicMax(status + 1, (uint32_t)b + 1u); }
      if (n < 0) { atomicOr(status, kStatOutOverflow); atomicMax(status + 1, (uint32_t)b + 1u); }
      n_out[b] = n < 0 ? 0 : n;
    }
  }
  POST_T1(4, t_all);
}

constexpr int kScanWarps = 4;  // row blocks per CTA

__global__ void __launch_bounds__(kScanWarps * 32)
    line_scan_kernel(LineBuffers L, int nb, int pass) {
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int base = (blockIdx.x * kScanWarps + (threadIdx.x >> 5)) * 32;
  int32_t* cnt = L.m_cnt + 2 * (size_t)b;
  const int n = cnt[0];
  if (base >= n) return;
  line::MergeScratch S = post_scratch(L, b, nullptr);
  float angle_thr, distance_thr, endpoint_thr;
  linew::merge_pass_thresholds(pass, angle_thr, distance_thr, endpoint_thr);
  POST_T0(t_scan);
  const int ovf = linew::merge_scan_rows(base, n, angle_thr, distance_thr, endpoint_thr, S, lane);
  POST_T1(1, t_scan);
  if (ovf && lane == 0) cnt[1] = 1;
}

void launch_line_post(const LineBuffers& L, int nb, int nfeatures, psl_keyline* kl, double* lineeq, int cap,
                      int32_t* n_out, uint32_t* status, cudaStream_t st) {
  const int grid = (nb + kPostWarps - 1) / kPostWarps;
  const dim3 sgrid((L.raw_cap + 32 * kScanWarps - 1) / (32 * kScanWarps), nb);
  line_post_kernel<0><<<grid, kPostWarps * 32, 0, st>>>(L, nb, nfeatures, kl, lineeq, cap, n_out, status);
  line_scan_kernel<<<sgrid, kScanWarps * 32, 0, st>>>(L, nb, 0);
  line_post_kernel<1><<<grid, kPostWarps * 32, 0, st>>>(L, nb, nfeatures, kl, lineeq, cap, n_out, status);
  line_scan_kernel<<<sgrid, kScanWarps * 32, 0, st>>>(L, nb, 1);
  line_post_kernel<2><<<grid, kPostWarps * 32, 0, st>>>(L, nb, nfeatures, kl, lineeq, cap, n_out, status);
}

// ---------------------------------------------------------------------------------------------------
// LBD
// ---------------------------------------------------------------------------------------------------
constexpr int kBandW = 7, kBands = 9, kLspH = kBandW * kBands;  // widthOfBand_, NUM_OF_BANDS, heightOfLSP

__constant__ float c_gaussG[kLspH];      // gaussCoefG_ as float (the reference casts at every use)
__constant__ float c_gaussL[3 * kBandW];  // gaussCoefL_

void upload_lbd_tables() {
  // BinaryDescriptor ctor (:219-261): the integer divisions are the reference's
  double gL[3 * kBandW], gG[kLspH];
  double u = (kBandW * 3 - 1) / 2, sigma = (kBandW * 2 + 1) / 2, inv = -1 / (2 * sigma * sigma);
  for (int i = 0; i < 3 * kBandW; ++i) gL[i] = exp((i - u) * (i - u) * inv);
  u = (kBands * kBandW - 1) / 2;
  sigma = u;
  inv = -1 / (2 * sigma * sigma);
  for (int i = 0; i < kLspH; ++i) gG[i] = exp((i - u) * (i - u) * inv);
  float fL[3 * kBandW], fG[kLspH];
  for (int i = 0; i < 3 * kBandW; ++i) fL[i] = (float)gL[i];
  for (int i = 0; i < kLspH; ++i) fG[i] = (float)gG[i];
  cudaMemcpyToSymbol(c_gaussL, fL, sizeof(fL));
  cudaMemcpyToSymbol(c_gaussG, fG, sizeof(fG));
}

__device__ __forceinline__ int reflect101_1(int i, int n) {  // one reflection is enough for a 1-px halo, n >= 2
  if (n == 1) return 0;
  return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i);
}

// cv::Sobel(CV_16SC1, 3x3, BORDER_REFLECT_101) of the blurred image, dx and dy interleaved (:374-399).
// A thread owns 4 adjacent pixels: three aligned words per row (left, own, right) give the 6 bytes it needs,
// and the four (dx, dy) pairs leave as one 16-byte store.  Threads on the left / right image edge take the
// per-pixel path with the reflected column.
__device__ __forceinline__ void sobel_px(const uint8_t* r0, const uint8_t* r1, const uint8_t* r2, int x, int w, short2* out) {
  const int xm = reflect101_1(x - 1, w), xp = reflect101_1(x + 1, w);
  const int dx = ((int)r0[xp] - (int)r0[xm]) + 2 * ((int)r1[xp] - (int)r1[xm]) + ((int)r2[xp] - (int)r2[xm]);
  const int dy = ((int)r2[xm] + 2 * (int)r2[x] + (int)r2[xp]) - ((int)r0[xm] + 2 * (int)r0[x] + (int)r0[xp]);
  *out = make_short2((short)dx, (short)dy);
}

__global__ void __launch_bounds__(128)
    sobel_px_kernel(const uint8_t* __restrict__ img, int pitch, int64_t fs, int w, int h, short2* __restrict__ gxy) {
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y, b = blockIdx.z;
  if (x >= w) return;
  const uint8_t* base = img + (size_t)b * fs;
  const uint8_t* r0 = base + (size_t)reflect101_1(y - 1, h) * pitch;
  const uint8_t* r1 = base + (size_t)y * pitch;
  const uint8_t* r2 = base + (size_t)reflect101_1(y + 1, h) * pitch;
  short2* out = gxy + ((size_t)b * h + y) * w + x;
  if (x == 0 || x + 4 >= w || (w & 3)) {
    for (int k = 0; k < 4 && x + k < w; ++k) sobel_px(r0, r1, r2, x + k, w, out + k);
    return;
  }
  // bytes x-1 .. x+4 of each row
  unsigned p[3][6];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const uint8_t* row = r == 0 ? r0 : (r == 1 ? r1 : r2);
    const unsigned wl = *reinterpret_cast<const unsigned*>(row + x - 4), wc = *reinterpret_cast<const unsigned*>(row + x),
                   wr = *reinterpret_cast<const unsigned*>(row + x + 4);
    p[r][0] = wl >> 24;
    p[r][1] = wc & 0xFFu; p[r][2] = (wc >> 8) & 0xFFu; p[r][3] = (wc >> 16) & 0xFFu; p[r][4] = wc >> 24;
    p[r][5] = wr & 0xFFu;
  }
  short2 o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int dx = ((int)p[0][k + 2] - (int)p[0][k]) + 2 * ((int)p[1][k + 2] - (int)p[1][k]) + ((int)p[2][k + 2] - (int)p[2][k]);
    const int dy = ((int)p[2][k] + 2 * (int)p[2][k + 1] + (int)p[2][k + 2]) - ((int)p[0][k] + 2 * (int)p[0][k + 1] + (int)p[0][k + 2]);
    o[k] = make_short2((short)dx, (short)dy);
  }
  *reinterpret_cast<uint4*>(out) = *reinterpret_cast<const uint4*>(o);
}

// The same on whole words for images whose width is a multiple of 4: a thread owns 4 adjacent pixels and walks down
// kSobelBand rows with the three rows of its 6-pixel window (unpacked) in registers, so every image word is loaded
// once per band instead of three times, and the reflected column of the left / right image edge is a register copy.
// dx[j] = s[j+2] - s[j] with the column sums s = a + 2b + c, dy[j] = t[j] + 2 t[j+1] + t[j+2] with t = c - a
// (integer arithmetic: identical to the direct 3x3 sums).
constexpr int kSobelBand = 16;

__global__ void __launch_bounds__(128)
    sobel_kernel(const uint8_t* __restrict__ img, int pitch, int64_t fs, int w, int h, short2* __restrict__ gxy) {
  const int nw = w >> 2, nbands = (h + kSobelBand - 1) / kSobelBand;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (idx >= nw * nbands) return;
  const int band = idx / nw, xw = idx - band * nw, x = xw * 4;
  const uint8_t* base = img + (size_t)b * fs + x;
  const bool has_l = xw > 0, has_r = x + 4 < w;
  auto load_row = [&](int y, int* p) {
    const uint8_t* row = base + (size_t)reflect101_1(y, h) * pitch;
    const unsigned wc = *reinterpret_cast<const unsigned*>(row);
    p[1] = wc & 0xFFu; p[2] = (wc >> 8) & 0xFFu; p[3] = (wc >> 16) & 0xFFu; p[4] = wc >> 24;
    p[0] = has_l ? (int)(*reinterpret_cast<const unsigned*>(row - 4) >> 24) : p[2];      // x-1, reflected: x+1
    p[5] = has_r ? (int)(*reinterpret_cast<const unsigned*>(row + 4) & 0xFFu) : p[3];    // x+4, reflected: x+2
  };
  const int y0 = band * kSobelBand, y1 = min(y0 + kSobelBand, h);
  int ra[6], rb[6], rc[6];
  load_row(y0 - 1, ra);
  load_row(y0, rb);
  short2* out = gxy + ((size_t)b * h + y0) * w + x;
  for (int y = y0; y < y1; ++y, out += w) {
    load_row(y + 1, rc);
    int sc[6], t[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      sc[k] = ra[k] + 2 * rb[k] + rc[k];
      t[k] = rc[k] - ra[k];
    }
    short2 o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = make_short2((short)(sc[j + 2] - sc[j]), (short)(t[j] + 2 * t[j + 1] + t[j + 2]));
    *reinterpret_cast<uint4*>(out) = *reinterpret_cast<const uint4*>(o);
#pragma unroll
    for (int k = 0; k < 6; ++k) { ra[k] = rb[k]; rb[k] = rc[k]; }
  }
}

// computeLBD for one line (:1074-1330) + binaryConversion (:402-413, :655-668).  Block = one line:
// thread = one row of the line support region for the fp32 row sums (sequential along the line, as in the
// reference, so the sums round identically), then thread = band for the 8 band statistics (rows in hID
// order), then thread 0 for the two normalisations.
__global__ void __launch_bounds__(64)
    lbd_kernel(const short2* __restrict__ gxy, int w, int h, const psl_keyline* __restrict__ kl,
               const int32_t* __restrict__ n_kl, int cap, uint8_t* __restrict__ ldesc, float* __restrict__ lbd72) {
  const int li = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  if (li >= n_kl[b]) return;
  const psl_keyline k = kl[(size_t)b * cap + li];
  const short2* g = gxy + (size_t)b * w * h;
  __shared__ float rows[kLspH][4];
  __shared__ float des[8 * kBands];
  const float dL0 = (float)cos((double)k.angle), dL1 = (float)sin((double)k.angle);
  const float dO0 = -dL1, dO1 = dL0;
  if (t < kLspH) {
    const short lengthOfLSP = (short)k.num_pixels;
    const short halfHeight = (kLspH - 1) / 2, halfWidth = (short)((lengthOfLSP - 1) / 2);
    const short imageWidth = (short)(w - 1), imageHeight = (short)(h - 1);
    const float midX = (float)(0.5 * (double)(k.s_oct_x + k.e_oct_x)), midY = (float)(0.5 * (double)(k.s_oct_y + k.e_oct_y));
    float sCorX0 = -dL0 * (float)halfWidth + dL1 * (float)halfHeight + midX;
    float sCorY0 = -dL1 * (float)halfWidth - dL0 * (float)halfHeight + midY;
    for (int r = 0; r < t; ++r) {  // the reference steps the row origin incrementally in fp32
      sCorX0 -= dL1;
      sCorY0 += dL0;
    }
    float sCorX = sCorX0, sCorY = sCorY0;
    float pgdL = 0, ngdL = 0, pgdO = 0, ngdO = 0;
    // One sample: round() of the reference (half away from zero, on the float promoted to double) == roundf on the
    // float; the coordinates step in fp32 independently of the gradients, so the gathers of kAhead samples are issued
    // together and the four sums then take them in the reference's order.
    // (the reference rounds into a `short`; inside +/- 32767 the int below is that value, and coordinates beyond it --
    // undefined behaviour there -- cannot occur for a support region of a line inside an image of at most 16384 pixels)
    auto sample_addr = [&](float sx, float sy) -> int {
      const int xCor = min(max((int)roundf(sx), 0), (int)imageWidth);
      const int yCor = min(max((int)roundf(sy), 0), (int)imageHeight);
      return yCor * w + xCor;
    };
    auto accumulate = [&](short2 gg) {
      const float gDL = (float)gg.x * dL0 + (float)gg.y * dL1, gDO = (float)gg.x * dO0 + (float)gg.y * dO1;
      if (gDL > 0) pgdL += gDL; else ngdL -= gDL;
      if (gDO > 0) pgdO += gDO; else ngdO -= gDO;
    };
    int wID = 0;
    const int len = (int)lengthOfLSP;
    constexpr int kAhead = 8;
    for (; wID + kAhead <= len; wID += kAhead) {
      int a[kAhead];
#pragma unroll
      for (int u = 0; u < kAhead; ++u) {
        a[u] = sample_addr(sCorX, sCorY);
        sCorX += dL0;
        sCorY += dL1;
      }
      short2 gv[kAhead];
#pragma unroll
      for (int u = 0; u < kAhead; ++u) gv[u] = __ldg(g + a[u]);
#pragma unroll
      for (int u = 0; u < kAhead; ++u) accumulate(gv[u]);
    }
    for (; wID < len; ++wID) {
      accumulate(__ldg(g + sample_addr(sCorX, sCorY)));
      sCorX += dL0;
      sCorY += dL1;
    }
    const float coef = c_gaussG[t];
    rows[t][0] = coef * pgdL;
    rows[t][1] = coef * ngdL;
    rows[t][2] = coef * pgdO;
    rows[t][3] = coef * ngdO;
  }
  __syncthreads();
  if (t < kBands) {
    float pL = 0, nL = 0, pL2 = 0, nL2 = 0, pO = 0, nO = 0, pO2 = 0, nO2 = 0;
    const int h0 = max(0, kBandW * (t - 1)), h1 = min(kLspH, kBandW * (t + 2));
    for (int hID = h0; hID < h1; ++hID) {
      const int band = hID / kBandW;
      // a row adds to its own band with gaussL[r+7], to band-1 with gaussL[r+14], to band+1 with gaussL[r]
      const float c = c_gaussL[hID % kBandW + (band == t ? kBandW : band == t + 1 ? 2 * kBandW : 0)];
      const float pgdL = rows[hID][0], ngdL = rows[hID][1], pgdO = rows[hID][2], ngdO = rows[hID][3];
      const float pgdL2 = pgdL * pgdL, ngdL2 = ngdL * ngdL, pgdO2 = pgdO * pgdO, ngdO2 = ngdO * ngdO;
      pL += c * pgdL; nL += c * ngdL;
      pL2 += c * c * pgdL2; nL2 += c * c * ngdL2;
      pO += c * pgdO; nO += c * ngdO;
      pO2 += c * c * pgdO2; nO2 += c * c * ngdO2;
    }
    const float invN = (t == 0 || t == kBands - 1) ? (float)(1.0 / (kBandW * 2.0)) : (float)(1.0 / (kBandW * 3.0));
    float* d = des + t * 8;
    float q = pL * invN; d[0] = q; d[4] = (float)sqrt((double)(pL2 * invN - q * q));
    q = nL * invN; d[1] = q; d[5] = (float)sqrt((double)(nL2 * invN - q * q));
    q = pO * invN; d[2] = q; d[6] = (float)sqrt((double)(pO2 * invN - q * q));
    q = nO * invN; d[3] = q; d[7] = (float)sqrt((double)(nO2 * invN - q * q));
  }
  __syncthreads();
  if (t == 0) {
    float tempM = 0, tempS = 0;
    for (int bd = 0; bd < kBands; ++bd) {
      const float* d = des + 8 * bd;
      tempM += d[0] * d[0]; tempM += d[1] * d[1]; tempM += d[2] * d[2]; tempM += d[3] * d[3];
      tempS += d[4] * d[4]; tempS += d[5] * d[5]; tempS += d[6] * d[6]; tempS += d[7] * d[7];
    }
    tempM = (float)(1 / sqrt((double)tempM));
    tempS = (float)(1 / sqrt((double)tempS));
    for (int bd = 0; bd < kBands; ++bd) {
      float* d = des + 8 * bd;
      d[0] *= tempM; d[1] *= tempM; d[2] *= tempM; d[3] *= tempM;
      d[4] *= tempS; d[5] *= tempS; d[6] *= tempS; d[7] *= tempS;
    }
    for (int i = 0; i < 72; ++i)
      if ((double)des[i] > 0.4) des[i] = (float)0.4;
    float temp = 0;
    for (int i = 0; i < 72; ++i) temp += des[i] * des[i];
    temp = (float)(1 / sqrt((double)temp));
    for (int i = 0; i < 72; ++i) des[i] = des[i] * temp;
  }
  __syncthreads();
  const size_t row = (size_t)b * cap + li;
  if (lbd72) {
    lbd72[row * 72 + t] = des[t];
    if (t < 8) lbd72[row * 72 + 64 + t] = des[64 + t];
  }
  if (t < 32) {
    // the 32 band pairs (i, j), i < j, of `combinations` (:76-109) in its order
    constexpr unsigned char ci[32] = {0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 6, 6, 7};
    constexpr unsigned char cj[32] = {1, 2, 3, 4, 5, 6, 2, 3, 4, 5, 6, 3, 4, 5, 6, 7, 8, 4, 5, 6, 7, 8, 5, 6, 7, 8, 6, 7, 8, 7, 8, 8};
    const float* f1 = des + 8 * ci[t];
    const float* f2 = des + 8 * cj[t];
    unsigned r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (f1[i] > f2[i]) r += 1u << i;
    ldesc[row * 32 + t] = (uint8_t)r;
  }
}

void launch_lbd(const LineBuffers& L, ImgBatch in, int nb, int nfeatures, const psl_keyline* kl, const int32_t* n_kl,
                int cap, uint8_t* ldesc, float* lbd72, cudaStream_t st) {
  ImgBatchMut bl{L.blur, L.pitch, (int64_t)L.pitch * L.h, L.w, L.h};
  launch_blur7(in, bl, 0, 14, 62, 104, nb, st);  // GaussianBlur(5x5, sigma 1), computeGaussianPyramid :351-371
  if ((L.w & 3) == 0 && (L.pitch & 3) == 0) {
    const int nbands = (L.h + kSobelBand - 1) / kSobelBand;
    dim3 g1(((L.w >> 2) * nbands + 127) / 128, nb);
    sobel_kernel<<<g1, 128, 0, st>>>(L.blur, L.pitch, (int64_t)L.pitch * L.h, L.w, L.h, L.gxy);
  } else {
    dim3 g1(((L.w + 3) / 4 + 127) / 128, L.h, nb);
    sobel_px_kernel<<<g1, 128, 0, st>>>(L.blur, L.pitch, (int64_t)L.pitch * L.h, L.w, L.h, L.gxy);
  }
  dim3 g2(nfeatures < cap ? nfeatures : cap, nb);
  lbd_kernel<<<g2, 64, 0, st>>>(L.gxy, L.w, L.h, kl, n_kl, cap, ldesc, lbd72);
}

}  // namespace psl
