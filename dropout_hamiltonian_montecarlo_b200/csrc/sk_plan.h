// Work decomposition of k_tc_bwd_sk (softmax_bwd_sk.cuh): plain C++ so that the host-side test (tests/test_sk_plan_cpu.py)
// can compile it with g++ and check that every (item, chunk) is owned by exactly one cluster.  See softmax_bwd_sk.cuh
// for the scheme.
#pragma once
#ifndef __CUDACC__
#ifndef __host__
#define __host__
#endif
#ifndef __device__
#define __device__
#endif
#ifndef __forceinline__
#define __forceinline__ inline
#endif
#endif

static constexpr int SK_MAX_PIECES = 2;   // items a cluster can touch (the host keeps the cost per cluster below the
                                          // cost of the cheapest item)

struct SkPlan {
  int on;             // 0: classic split-K partials (PartRegions), 1: this decomposition
  int kc;             // chunks of the contraction (rows of the window / 64)
  int n_nt, bn;       // feature tiles and their width (multiple of 32, <= 192)
  int n_pair;         // 256-row pair tiles of DmT
  int has_half;       // a last 128-row tile, split 64 + 64 over the two CTAs
  int wp, wh;         // cost of one chunk of a pair / half item
  int lp, lh;         // lanes per pair / half item
  int T;              // cost per cluster
  int n_clusters;
  int piece_elems;    // floats per piece: 2 CTAs x bn x 128
};

__host__ __device__ __forceinline__ int sk_n_items(const SkPlan& s) { return (s.n_pair + s.has_half) * s.n_nt; }
__host__ __device__ __forceinline__ bool sk_item_half(const SkPlan& s, int i) { return i >= s.n_pair * s.n_nt; }
__host__ __device__ __forceinline__ int sk_item_start(const SkPlan& s, int i) {
  const int np = s.n_pair * s.n_nt;
  return i < np ? i * s.wp * s.kc : np * s.wp * s.kc + (i - np) * s.wh * s.kc;
}
__host__ __device__ __forceinline__ int sk_item_of_cost(const SkPlan& s, int x) {
  const int np = s.n_pair * s.n_nt, cp = np * s.wp * s.kc;
  const int i = x < cp ? x / (s.wp * s.kc) : np + (x - cp) / (s.wh * s.kc);
  const int n = sk_n_items(s);
  return i < n ? i : n;
}

// The entries of a cluster inside one item are a contiguous range of the lane-major order: the tail of lane a0, then
// whole lanes, then the head of the last lane.  Run r (lane a0 + r) therefore covers the positions
//   [r == 0 ? b_first : 0,  r == n_runs - 1 ? b_last : entries of that lane)
// which the kernel roles evaluate from registers (an earlier version kept a run table in shared memory: its loads
// queued behind the UMMA / TMA traffic of the saturated shared-memory port and cost ~600 cycles per chunk).
struct SkPiece {
  int item, half, L, n_chunks, slot;
  int a0, n_runs, b_first, b_last;
  int q, r;          // lanes a < r hold q + 1 entries, the others q
  int b_min, b_max;  // positions covered by the runs: the piece is walked position by position, every run that holds
                     // the position contributing its chunk (all lanes of a cluster advance together)
};
struct SkWork {
  int n_pieces;
  SkPiece piece[SK_MAX_PIECES];
};
__host__ __device__ __forceinline__ void sk_run_bounds(const SkPiece& pc, int run, int* lo, int* hi) {
  *lo = run == 0 ? pc.b_first : 0;
  *hi = run == pc.n_runs - 1 ? pc.b_last : pc.q + ((pc.a0 + run) < pc.r ? 1 : 0);
}

// the entries of cluster `cl`: one piece per item it touches; pieces ordered by where they lie in the contraction
__host__ __device__ inline void sk_build(const SkPlan& s, int cl, SkWork& w) {
  w.n_pieces = 0;
  const int lo = cl * s.T, hi = lo + s.T;
  const int n_items = sk_n_items(s);
  for (int i = sk_item_of_cost(s, lo); i < n_items && w.n_pieces < SK_MAX_PIECES; ++i) {
    const bool half = sk_item_half(s, i);
    const int S = sk_item_start(s, i), wg = half ? s.wh : s.wp;
    if (S >= hi) break;
    int e0 = lo <= S ? 0 : (lo - S + wg - 1) / wg;
    int e1 = (hi - S + wg - 1) / wg;
    if (e1 > s.kc) e1 = s.kc;
    if (e0 >= e1) continue;
    SkPiece& pc = w.piece[w.n_pieces];
    pc.item = i;
    pc.half = half ? 1 : 0;
    pc.L = half ? s.lh : s.lp;
    pc.n_chunks = e1 - e0;
    pc.slot = lo < S ? 1 : 0;  // the cluster's range starts in the previous item: that one owns slot 0
    const int L = pc.L, q = s.kc / L, r = s.kc % L;
    pc.q = q, pc.r = r;
    // lane and position of entry e in lane-major order
    auto locate = [&](int e, int* a, int* b) {
      if (e < r * (q + 1)) {
        *a = e / (q + 1), *b = e % (q + 1);
      } else {
        const int e2 = e - r * (q + 1);
        *a = r + e2 / q, *b = e2 % q;
      }
    };
    int a1, b1;
    locate(e0, &pc.a0, &pc.b_first);
    locate(e1 - 1, &a1, &b1);
    pc.n_runs = a1 - pc.a0 + 1;
    pc.b_last = b1 + 1;
    pc.b_min = 1 << 30, pc.b_max = 0;
    for (int x = 0; x < pc.n_runs; ++x) {
      int rl, rh;
      sk_run_bounds(pc, x, &rl, &rh);
      if (rl < pc.b_min) pc.b_min = rl;
      if (rh > pc.b_max) pc.b_max = rh;
    }
    ++w.n_pieces;
  }
  // the piece that lies earlier in the contraction goes first (compare the centres of the chunk ranges)
  if (w.n_pieces == 2 && (w.piece[1].b_min + w.piece[1].b_max) * w.piece[1].L < (w.piece[0].b_min + w.piece[0].b_max) * w.piece[0].L) {
    const SkPiece t = w.piece[0];
    w.piece[0] = w.piece[1];
    w.piece[1] = t;
  }
}

// widest feature tile (multiple of 32, <= 192: the epilogue keeps bn/64 16-column chunks per thread) with the least padding
inline int sk_pick_bn(int n_feat) {
  int best = 0, best_tot = 1 << 30;
  for (int bn = 192; bn >= 96; bn -= 32) {
    const int tot = (n_feat + bn - 1) / bn * bn;
    if (tot < best_tot) best_tot = tot, best = bn;
  }
  return best;
}

// Plan for rows_m chain-class rows x n_feat feature rows x kc chunks on n_clusters CTA pairs.  Returns false when the
// shape does not fit the scheme (fewer than two 128-row tiles, a cluster would span more than two items, or the cost
// line does not fit 32-bit arithmetic): the caller then uses the row-slab kernels.
inline bool sk_make_plan(long long rows_m, int n_feat, int kc, int n_clusters, int wh, SkPlan* out) {
  SkPlan s{};
  const long long tiles128 = (rows_m + 127) / 128;
  if (tiles128 < 2 || kc < 1 || n_clusters < 1 || tiles128 > 4096) return false;
  s.kc = kc;
  s.bn = sk_pick_bn(n_feat);
  s.n_nt = (n_feat + s.bn - 1) / s.bn;
  s.n_pair = (int)(tiles128 / 2);
  s.has_half = (int)(tiles128 & 1);
  s.wp = 10;
  s.wh = wh < 1 ? 1 : wh;
  s.n_clusters = n_clusters;
  const long long W = (long long)(s.n_pair * s.wp + s.has_half * s.wh) * s.n_nt * s.kc;
  // every cluster must own at least one entry of every item inside its cost range (the reduce kernels enumerate an
  // item's pieces in closed form): the cost per cluster may not fall below the cost of one chunk -> fewer clusters
  const long long w_max = s.has_half && s.wh > s.wp ? s.wh : s.wp;
  if (W / n_clusters < w_max) n_clusters = (int)(W / w_max);
  if (n_clusters < 1) return false;
  s.n_clusters = n_clusters;
  const long long T = (W + n_clusters - 1) / n_clusters;
  const long long min_item = (long long)(s.has_half && s.wh < s.wp ? s.wh : s.wp) * s.kc;
  if (W >= (1LL << 30) || T > min_item || T < 1) return false;
  s.on = 1;
  s.T = (int)T;
  s.lp = (int)(((long long)s.wp * s.kc + T - 1) / T);
  s.lh = (int)(((long long)s.wh * s.kc + T - 1) / T);
  if (s.lp < 1) s.lp = 1;
  if (s.lh < 1) s.lh = 1;
  s.piece_elems = 2 * s.bn * 128;
  *out = s;
  return true;
}
