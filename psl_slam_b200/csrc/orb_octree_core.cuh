// DistributeOctTree (ORBextractor.cc:539-763) as a level-synchronous, CTA-cooperative
// algorithm.  One CTA owns one (frame, level).
//
// The reference walks a std::list of nodes and quarters them one at a time; what defines
// the result is (1) the list order (children are push_front'ed in UL,UR,BL,BR order),
// (2) the stable order of keys inside a node, (3) the order in which phase-2 candidates
// are expanded (descending (nKeys, address); address := creation sequence) and
// the exact early-exit point.  Here a whole round is done at once:
//   - every key of a node to be divided computes its quadrant; one CTA-wide exclusive scan
//     of packed one-hot counters over the key array gives each key its stable rank inside
//     its child (rank = S(i) - S(node.begin)) and each node its four child counts;
//   - a scan over nodes in processing order assigns creation indices; the new list is
//     [children in reverse creation order] ++ [undivided nodes in old order];
//   - phase 2 ranks the candidate nodes by (count, seq) descending, prefix-sums the list
//     growth and cuts exactly where the reference's `break` fires.
// Keys never leave their parent's [begin, begin+cnt) range, so a node is just a range.
//
// The same source compiles for the device (one thread per `tid`) and, with PSL_HOST_EMU,
// for the host (threads emulated phase by phase) so that tests can drive it without a GPU.
#pragma once
#include <stdint.h>

#ifdef PSL_HOST_EMU
#include <string.h>
#define PSL_HD
#define PSL_PHASE_BEGIN for (int tid = 0; tid < NT; ++tid) {
#define PSL_PHASE_END }
#else
#define PSL_HD __device__ __forceinline__
#define PSL_PHASE_BEGIN { const int tid = (int)threadIdx.x;
#define PSL_PHASE_END } __syncthreads();
#endif

namespace psl {
namespace octree {

struct Node {          // 16 bytes
  int16_t ulx, urx, uly, bry;
  int32_t begin;       // first key position
  int32_t cnt : 24;    // number of keys
  uint32_t last : 1;   // created in the most recent round/pass with >1 key (phase-2 candidate)
  uint32_t pad : 7;
};

// Shared-memory working set of one CTA.  NODE_CAP bounds the list length.
template <int NT, int NODE_CAP>
struct Shared {
  Node tab[2][NODE_CAP];
  unsigned long long sbeg[NODE_CAP];  // S at node begin (packed 4x16 one-hot prefix)
  unsigned long long send[NODE_CAP];  // S after node end
  int32_t seq[2][NODE_CAP];           // creation index within the round that made the node
  int32_t po[NODE_CAP];               // processing-order index of a divided node, -1 if kept
  int32_t ne_by_po[NODE_CAP];         // #non-empty children by processing order -> excl. scan
  int32_t kept_idx[NODE_CAP];         // index among kept nodes (list order)
  uint16_t childpos[NODE_CAP][4];     // new list position of each child
  unsigned long long strip_base[NT];  // per-thread strip prefix
  unsigned long long scan_tmp[NT];
  int32_t iscan_tmp[NT];
  int32_t L, prevL, n_div, total_new, n_expand, n_cand, cut, finish, phase2, error;
};

PSL_HD unsigned long long onehot(int q) { return 1ull << (16 * q); }
PSL_HD int field(unsigned long long v, int q) { return (int)((v >> (16 * q)) & 0xFFFFull); }

// ---- CTA-wide exclusive scans over one value per thread --------------------------------------
#ifdef PSL_HOST_EMU
template <int NT>
inline unsigned long long scan_u64(unsigned long long* vals, unsigned long long*) {
  unsigned long long run = 0;
  for (int t = 0; t < NT; ++t) { unsigned long long v = vals[t]; vals[t] = run; run += v; }
  return run;
}
template <int NT>
inline int scan_i32(int32_t* vals, int32_t*) {
  int run = 0;
  for (int t = 0; t < NT; ++t) { int v = vals[t]; vals[t] = run; run += v; }
  return run;
}
#else
// vals[tid] holds the input; on return vals[tid] is the exclusive prefix; returns the total.
template <int NT>
__device__ __forceinline__ unsigned long long scan_u64(unsigned long long* vals, unsigned long long* tmp) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  unsigned long long v = vals[tid], inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    unsigned long long o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) tmp[wid] = inc;
  __syncthreads();
  unsigned long long base = 0, total = 0;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) {
    unsigned long long t = tmp[w];
    if (w < wid) base += t;
    total += t;
  }
  vals[tid] = base + inc - v;
  __syncthreads();
  return total;
}
template <int NT>
__device__ __forceinline__ int scan_i32(int32_t* vals, int32_t* tmp) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int v = vals[tid], inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) tmp[wid] = inc;
  __syncthreads();
  int base = 0, total = 0;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) {
    int t = tmp[w];
    if (w < wid) base += t;
    total += t;
  }
  vals[tid] = base + inc - v;
  __syncthreads();
  return total;
}
#endif

// In-place exclusive scan of arr[0..n) (n may exceed NT); returns the total.  Uniform call.
template <int NT, int NODE_CAP>
PSL_HD int scan_array(Shared<NT, NODE_CAP>& sh, int32_t* arr, int n) {
  int carry = 0;
  for (int base = 0; base < n; base += NT) {
    // stage the chunk in the (idle) u64 scan scratch, viewed as ints
    int32_t* stage = reinterpret_cast<int32_t*>(sh.scan_tmp);
    PSL_PHASE_BEGIN
    stage[tid] = (base + tid < n) ? arr[base + tid] : 0;
    PSL_PHASE_END
    int tot = scan_i32<NT>(stage, sh.iscan_tmp);
    PSL_PHASE_BEGIN
    if (base + tid < n) arr[base + tid] = stage[tid] + carry;
    PSL_PHASE_END
    carry += tot;
  }
  return carry;
}

PSL_HD int quadrant(uint32_t key, const Node& nd) {
  // DivideNode :512-526; candidate coordinates are small exact integers, so the reference's
  // float-vs-int compares reduce to integer compares.
  const int x = (int)(key >> 20), y = (int)((key >> 8) & 0xFFFu);
  const int mx = nd.ulx + ((nd.urx - nd.ulx + 1) >> 1);  // UL.x + ceil((UR.x-UL.x)/2)  :483
  const int my = nd.uly + ((nd.bry - nd.uly + 1) >> 1);  // UL.y + ceil((BR.y-UL.y)/2)  :484
  return (x < mx ? 0 : 1) + (y < my ? 0 : 2);            // n1=UL n2=UR n3=BL n4=BR
}

// One division round.  dv decided by sh.po[p] >= 0 (processing order already assigned for
// phase 1; for phase 2 it is assigned inside after the counts are known).
// keysA/knodeA: current; keysB/knodeB: next.  cur = index of the current table.
template <int NT, int NODE_CAP>
PSL_HD void divide_round(Shared<NT, NODE_CAP>& sh, int cur, int n, const uint32_t* keysA, const uint16_t* knodeA,
                         uint32_t* keysB, uint16_t* knodeB, int N, bool phase2) {
  Node* T = sh.tab[cur];
  Node* U = sh.tab[cur ^ 1];
  const int L = sh.L;
  const int strip = (n + NT - 1) / NT;

  // (a) which nodes try to divide
  PSL_PHASE_BEGIN
  for (int p = tid; p < L; p += NT) {
    bool d = phase2 ? (T[p].last != 0) : (T[p].cnt > 1);
    sh.po[p] = d ? 0 : -1;
    sh.sbeg[p] = 0;
    sh.send[p] = 0;
  }
  PSL_PHASE_END

  // (b) packed one-hot scan over the key array
  PSL_PHASE_BEGIN
  unsigned long long acc = 0;
  const int i0 = tid * strip, i1 = (i0 + strip < n) ? i0 + strip : n;
  for (int i = i0; i < i1; ++i) {
    const int p = knodeA[i];
    if (sh.po[p] >= 0) acc += onehot(quadrant(keysA[i], T[p]));
  }
  sh.strip_base[tid] = acc;
  PSL_PHASE_END
  scan_u64<NT>(sh.strip_base, sh.scan_tmp);
  PSL_PHASE_BEGIN
  unsigned long long S = sh.strip_base[tid];
  const int i0 = tid * strip, i1 = (i0 + strip < n) ? i0 + strip : n;
  for (int i = i0; i < i1; ++i) {
    const int p = knodeA[i];
    if (sh.po[p] >= 0) {
      if (i == T[p].begin) sh.sbeg[p] = S;
      S += onehot(quadrant(keysA[i], T[p]));
      if (i == T[p].begin + T[p].cnt - 1) sh.send[p] = S;
    }
  }
  PSL_PHASE_END

  // (c) processing order + cut
  if (!phase2) {
    // creation follows list order: po = index among dividing nodes
    PSL_PHASE_BEGIN
    for (int p = tid; p < L; p += NT) sh.kept_idx[p] = sh.po[p] >= 0 ? 1 : 0;
    PSL_PHASE_END
    int ndiv = scan_array<NT, NODE_CAP>(sh, sh.kept_idx, L);
    PSL_PHASE_BEGIN
    for (int p = tid; p < L; p += NT)
      if (sh.po[p] >= 0) sh.po[p] = sh.kept_idx[p];
    if (tid == 0) { sh.n_div = ndiv; sh.cut = ndiv; }
    PSL_PHASE_END
  } else {
    // rank candidates by (cnt, seq) descending  (sort ascending + walk from the back, :684-685)
    PSL_PHASE_BEGIN
    for (int p = tid; p < L; p += NT) {
      if (sh.po[p] < 0) continue;
      const int c = T[p].cnt, s = sh.seq[cur][p];
      int r = 0;
      for (int o = 0; o < L; ++o) {
        if (!T[o].last) continue;
        const int co = T[o].cnt, so = sh.seq[cur][o];
        r += (co > c || (co == c && so > s)) ? 1 : 0;
      }
      sh.po[p] = r;
    }
    if (tid == 0) sh.n_div = 0;
    PSL_PHASE_END
  }
  // non-empty children by processing order
  PSL_PHASE_BEGIN
  for (int p = tid; p < L; p += NT)
    if (sh.po[p] >= 0) {
      const unsigned long long c = sh.send[p] - sh.sbeg[p];
      int ne = 0;
      for (int q = 0; q < 4; ++q) ne += field(c, q) ? 1 : 0;
      sh.ne_by_po[sh.po[p]] = ne;
    }
  PSL_PHASE_END
  if (phase2) {
    // count candidates, then find the first processing index where the list reaches N (:733-734)
    PSL_PHASE_BEGIN
    for (int p = tid; p < L; p += NT) sh.kept_idx[p] = sh.po[p] >= 0 ? 1 : 0;
    PSL_PHASE_END
    int ncand = scan_array<NT, NODE_CAP>(sh, sh.kept_idx, L);
    PSL_PHASE_BEGIN
    if (tid == 0) {
      int size = L, cut = ncand;
      for (int r = 0; r < ncand; ++r) {
        size += sh.ne_by_po[r] - 1;
        if (size >= N) { cut = r + 1; break; }
      }
      sh.cut = cut;
      sh.n_div = cut;
    }
    PSL_PHASE_END
    PSL_PHASE_BEGIN
    for (int p = tid; p < L; p += NT)
      if (sh.po[p] >= sh.cut) sh.po[p] = -1;  // not reached before the break: stays as is
    PSL_PHASE_END
  }
  const int ndiv = sh.n_div;
  // exclusive scan of ne over processing order -> creation index base
  const int total_new = scan_array<NT, NODE_CAP>(sh, sh.ne_by_po, ndiv);
  PSL_PHASE_BEGIN
  for (int p = tid; p < L; p += NT) sh.kept_idx[p] = sh.po[p] < 0 ? 1 : 0;
  PSL_PHASE_END
  const int kept = scan_array<NT, NODE_CAP>(sh, sh.kept_idx, L);
  const int newL = total_new + kept;
  if (newL > NODE_CAP) {
    PSL_PHASE_BEGIN
    if (tid == 0) { sh.error = 1; sh.finish = 1; }
    PSL_PHASE_END
    return;
  }

  // (d) build the new table
  PSL_PHASE_BEGIN
  if (tid == 0) sh.n_expand = 0;
  PSL_PHASE_END
  PSL_PHASE_BEGIN
  for (int p = tid; p < L; p += NT) {
    const Node nd = T[p];
    if (sh.po[p] < 0) {
      const int pos = total_new + sh.kept_idx[p];
      Node o = nd;
      o.last = 0;
      U[pos] = o;
      sh.seq[cur ^ 1][pos] = 0;
      sh.childpos[p][0] = (uint16_t)pos;
    } else {
      const unsigned long long c = sh.send[p] - sh.sbeg[p];
      const int mx = nd.ulx + ((nd.urx - nd.ulx + 1) >> 1), my = nd.uly + ((nd.bry - nd.uly + 1) >> 1);
      int cidx = sh.ne_by_po[sh.po[p]], off = nd.begin, nexp = 0;
      for (int q = 0; q < 4; ++q) {
        const int cq = field(c, q);
        if (!cq) continue;
        Node o;
        o.ulx = (q & 1) ? mx : nd.ulx;
        o.urx = (q & 1) ? nd.urx : mx;
        o.uly = (q & 2) ? my : nd.uly;
        o.bry = (q & 2) ? nd.bry : my;
        o.begin = off;
        o.cnt = cq;
        o.last = cq > 1 ? 1u : 0u;
        o.pad = 0;
        const int pos = total_new - 1 - cidx;
        U[pos] = o;
        sh.seq[cur ^ 1][pos] = cidx;
        sh.childpos[p][q] = (uint16_t)pos;
        off += cq;
        ++cidx;
        nexp += cq > 1 ? 1 : 0;
      }
      if (nexp) {
#ifdef PSL_HOST_EMU
        sh.n_expand += nexp;
#else
        atomicAdd(&sh.n_expand, nexp);
#endif
      }
    }
  }
  PSL_PHASE_END

  // (e) move the keys (stable inside every child)
  PSL_PHASE_BEGIN
  unsigned long long S = sh.strip_base[tid];
  const int i0 = tid * strip, i1 = (i0 + strip < n) ? i0 + strip : n;
  for (int i = i0; i < i1; ++i) {
    const int p = knodeA[i];
    const uint32_t key = keysA[i];
    const bool tried = phase2 ? (T[p].last != 0) : (T[p].cnt > 1);
    int q = 0;
    if (tried) q = quadrant(key, T[p]);
    if (sh.po[p] >= 0) {
      const unsigned long long c = sh.send[p] - sh.sbeg[p];
      const unsigned long long rel = S - sh.sbeg[p];
      int dst = T[p].begin + field(rel, q);
      for (int qq = 0; qq < q; ++qq) dst += field(c, qq);
      keysB[dst] = key;
      knodeB[dst] = sh.childpos[p][q];
    } else {
      keysB[i] = key;
      knodeB[i] = sh.childpos[p][0];
    }
    if (tried) S += onehot(q);
  }
  if (tid == 0) { sh.prevL = L; sh.L = newL; sh.total_new = total_new; }
  PSL_PHASE_END
}

// Full selection for one (frame, level).  keys[0]: candidates in reference order (packed),
// n of them; keys[1], knode[0..1]: scratch of the same length.  roots: n_ini, hx.
// Writes the selected keys (list order) to out[0..cap) and returns the list length
// (or -1 on an internal bound violation).
template <int NT, int NODE_CAP>
PSL_HD int select(Shared<NT, NODE_CAP>& sh, int n, uint32_t* keys0, uint32_t* keys1, uint16_t* knode0,
                  uint16_t* knode1, int n_ini, float hx, int width, int height, int N, uint32_t* out, int cap) {
  uint32_t* K[2] = {keys0, keys1};
  uint16_t* KN[2] = {knode0, knode1};
  // ---- roots (:543-586): one pseudo round with "quadrant" = root index ----------------------
  const int strip = (n + NT - 1) / NT;
  PSL_PHASE_BEGIN
  unsigned long long acc = 0;
  const int i0 = tid * strip, i1 = (i0 + strip < n) ? i0 + strip : n;
  for (int i = i0; i < i1; ++i) {
    int r = (int)((float)(int)(keys0[i] >> 20) / hx);  // vpIniNodes[kp.pt.x/hX]  :569
    if (r < 0 || r >= n_ini) { r = n_ini - 1; sh.error = 2; }
    acc += onehot(r);
  }
  sh.strip_base[tid] = acc;
  if (tid == 0) { sh.finish = 0; sh.phase2 = 0; }
  PSL_PHASE_END
  const unsigned long long tot = scan_u64<NT>(sh.strip_base, sh.scan_tmp);
  PSL_PHASE_BEGIN
  if (tid == 0) {
    int L = 0, off = 0;
    for (int r = 0; r < n_ini; ++r) {
      const int c = field(tot, r);
      sh.childpos[0][r] = (uint16_t)L;
      if (c) {  // empty roots are erased (:580-581)
        Node o;
        o.ulx = (int16_t)(int)(hx * (float)r);        // :555
        o.urx = (int16_t)(int)(hx * (float)(r + 1));  // :556
        o.uly = 0;
        o.bry = (int16_t)height;
        o.begin = off;
        o.cnt = c;
        o.last = 0;
        o.pad = 0;
        sh.tab[1][L] = o;
        sh.seq[1][L] = r;
        ++L;
      }
      sh.po[r] = off;  // root begin (reuse po as scratch)
      off += c;
    }
    sh.L = L;
    sh.prevL = L;
  }
  PSL_PHASE_END
  PSL_PHASE_BEGIN
  unsigned long long S = sh.strip_base[tid];
  const int i0 = tid * strip, i1 = (i0 + strip < n) ? i0 + strip : n;
  for (int i = i0; i < i1; ++i) {
    const uint32_t key = keys0[i];
    int r = (int)((float)(int)(key >> 20) / hx);
    if (r < 0 || r >= n_ini) r = n_ini - 1;
    const int dst = sh.po[r] + field(S, r);
    keys1[dst] = key;
    knode1[dst] = sh.childpos[0][r];
    S += onehot(r);
  }
  PSL_PHASE_END
  (void)width;
  int cur = 1;  // tables/keys index currently valid

  // ---- main loop (:597-739) -----------------------------------------------------------------
  for (int guard = 0; guard < 64 && !sh.finish; ++guard) {
    const bool p2 = sh.phase2 != 0;
    divide_round<NT, NODE_CAP>(sh, cur, n, K[cur], KN[cur], K[cur ^ 1], KN[cur ^ 1], N, p2);
    if (sh.error == 1) return -1;
    cur ^= 1;
    PSL_PHASE_BEGIN
    if (tid == 0) {
      if (sh.L >= N || sh.L == sh.prevL) sh.finish = 1;                  // :667-671 / :736-737
      else if (!p2 && sh.L + sh.n_expand * 3 > N) sh.phase2 = 1;         // :672
    }
    PSL_PHASE_END
  }

  // ---- best response per node, first wins ties (:744-760) -----------------------------------
  const int L = sh.L;
  PSL_PHASE_BEGIN
  const Node* T = sh.tab[cur];
  const uint32_t* keys = K[cur];
  for (int p = tid; p < L; p += NT) {
    const Node nd = T[p];
    uint32_t best = keys[nd.begin];
    for (int k = 1; k < nd.cnt; ++k) {
      const uint32_t c = keys[nd.begin + k];
      if ((c & 0xFFu) > (best & 0xFFu)) best = c;
    }
    if (p < cap) out[p] = best;
  }
  PSL_PHASE_END
  return L;
}

}  // namespace octree
}  // namespace psl
