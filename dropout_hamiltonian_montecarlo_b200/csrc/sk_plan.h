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

// Item types.  The gradient block G^T[chain-class rows, feature rows] is cut into
//   SK_P  pair item       256 chain-class rows (two CTAs x 128) x bn features            M = 256, N = bn
// and, when the number of 128-row chain-class tiles is odd, the last tile is covered EITHER by
//   SK_H  half item       128 chain-class rows (two CTAs x 64) x bn features             M = 128, N = bn   (half rate)
// OR, transposed (feature rows back on M, where 256-row pairs exist again), by
//   SK_Q  transposed pair 256 feature rows (two CTAs x 128) x the 128 chain-class rows   M = 256, N = 128
//   SK_R  remainder       the 128 chain-class rows (two CTAs x 64) x the features beyond the last 256-row pair
//                                                                                        M = 128, N = bnr (small)
// cfg2 (640 x 785): 2 x 5 P items + 3 Q items + 1 R item (N = 32) instead of 5 H items: 129 cost units instead of 145.
enum { SK_P = 0, SK_H = 1, SK_Q = 2, SK_R = 3 };

struct SkPlan {
  int on;             // 0: classic split-K partials (PartRegions), 1: this decomposition
  int kc;             // chunks of the contraction (rows of the window / 64)
  int n_nt, bn;       // feature tiles of the P / H items and their width (multiple of 32, <= 192)
  int n_pair;         // 256-row pair tiles of DmT
  int odd;            // the odd last 128-row tile: 0 = none, 1 = H items, 2 = Q items (+ R)
  int n_fp, bnr;      // odd == 2: 256-feature pairs (Q items) and width of the R item (multiple of 32; 0 = no remainder)
  // item classes in cost-line order: 0 = P, 1 = H or Q, 2 = R
  int cnt[3];         // items per class
  int w[3];           // cost of one chunk of an item of the class
  int L[3];           // interleave lanes per item of the class
  int T;              // cost per cluster
  int n_clusters;
  int piece_elems;    // floats per piece (2 CTAs x columns x 128 lanes, the largest item type)
};

__host__ __device__ __forceinline__ int sk_n_items(const SkPlan& s) { return s.cnt[0] + s.cnt[1] + s.cnt[2]; }
__host__ __device__ __forceinline__ int sk_item_class(const SkPlan& s, int i) {
  return i < s.cnt[0] ? 0 : (i < s.cnt[0] + s.cnt[1] ? 1 : 2);
}
// (no dynamic indexing of the plan's arrays in device code: a kernel parameter indexed at run time is copied to local memory)
__host__ __device__ __forceinline__ int sk_class_w(const SkPlan& s, int c) { return c == 0 ? s.w[0] : (c == 1 ? s.w[1] : s.w[2]); }
__host__ __device__ __forceinline__ int sk_class_L(const SkPlan& s, int c) { return c == 0 ? s.L[0] : (c == 1 ? s.L[1] : s.L[2]); }
__host__ __device__ __forceinline__ int sk_item_type(const SkPlan& s, int i) {
  const int c = sk_item_class(s, i);
  return c == 0 ? SK_P : (c == 2 ? SK_R : (s.odd == 2 ? SK_Q : SK_H));
}
__host__ __device__ __forceinline__ int sk_item_start(const SkPlan& s, int i) {
  const int c0 = s.cnt[0] * s.w[0] * s.kc, c1 = s.cnt[1] * s.w[1] * s.kc;
  if (i < s.cnt[0]) return i * s.w[0] * s.kc;
  if (i < s.cnt[0] + s.cnt[1]) return c0 + (i - s.cnt[0]) * s.w[1] * s.kc;
  return c0 + c1 + (i - s.cnt[0] - s.cnt[1]) * s.w[2] * s.kc;
}
__host__ __device__ __forceinline__ int sk_item_of_cost(const SkPlan& s, int x) {
  const int c0 = s.cnt[0] * s.w[0] * s.kc, c1 = s.cnt[1] * s.w[1] * s.kc;
  int i;
  if (x < c0) i = x / (s.w[0] * s.kc);
  else if (x < c0 + c1) i = s.cnt[0] + (x - c0) / (s.w[1] * s.kc);
  else i = s.cnt[0] + s.cnt[1] + (x - c0 - c1) / (s.w[2] * s.kc);
  const int n = sk_n_items(s);
  return i < n ? i : n;
}

// The entries of a cluster inside one item are a contiguous range of the lane-major order: the tail of lane a0, then
// whole lanes, then the head of the last lane.  Run r (lane a0 + r) therefore covers the positions
//   [r == 0 ? b_first : 0,  r == n_runs - 1 ? b_last : entries of that lane)
// which the kernel roles evaluate from registers (an earlier version kept a run table in shared memory: its loads
// queued behind the UMMA / TMA traffic of the saturated shared-memory port and cost ~600 cycles per chunk).
struct SkPiece {
  int item, type, idx;  // idx: index inside the item's class
  int L, n_chunks, slot;
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
    const int cls = sk_item_class(s, i);
    const int S = sk_item_start(s, i), wg = sk_class_w(s, cls);
    if (S >= hi) break;
    int e0 = lo <= S ? 0 : (lo - S + wg - 1) / wg;
    int e1 = (hi - S + wg - 1) / wg;
    if (e1 > s.kc) e1 = s.kc;
    if (e0 >= e1) continue;
    SkPiece& pc = w.piece[w.n_pieces];
    pc.item = i;
    pc.type = sk_item_type(s, i);
    pc.idx = i - (cls == 0 ? 0 : (cls == 1 ? s.cnt[0] : s.cnt[0] + s.cnt[1]));
    pc.L = sk_class_L(s, cls);
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

// Cost of one 64-row chunk in units of 1/10 of a 256 x 160 pair chunk (12 MMAs of 80 cycles).  An MMA with M = 128 over
// two CTAs takes as long as one with M = 256 (measured: half items cost 0.9-1.0 pair chunks); below ~80 columns the 13
// tcgen05 instructions of a chunk take longer to ISSUE (~36 cycles each) than to execute, hence the floor.
struct SkTune {
  int wh_pct;   // cost of a half-rate item relative to a full-rate one of the same width, per cent (default 90)
  int w_floor;  // cost floor of any chunk (default 5)
  int transposed;  // 1: cover an odd tile by Q / R items where that is cheaper, 0: always H items (default)
  // measured costs of the transposed items at cfg2 (in-kernel counters, 64 chains): a Q chunk takes ~1040 cycles (its
  // MMAs need 768: N = 128 re-reads the A tile from shared memory for fewer columns, and the shared-memory port is the
  // limit), an R chunk ~790 (13 instructions and 20 KB per chunk: bound by the TMA latency over 4 stages, not by work)
  int w_q, w_r_floor;  // defaults 10, 8 (best of a 3 x 2 sweep: 142.1 us per 64-chain launch against 146.5 with H items --
                       // 3 %, inside the box-to-box spread, so H items stay the default)
};
inline int sk_chunk_cost(int n_cols, bool half_rate, const SkTune& t) {
  int w = (10 * n_cols * (half_rate ? t.wh_pct : 100) + 8000) / 16000;
  return w < t.w_floor ? t.w_floor : w;
}

// Plan for rows_m chain-class rows x n_feat feature rows x kc chunks on n_clusters CTA pairs.  Returns false when the
// shape does not fit the scheme (fewer than two 128-row tiles, a cluster would span more than two items, or the cost
// line does not fit 32-bit arithmetic): the caller then uses the row-slab kernels.
inline bool sk_make_plan(long long rows_m, int n_feat, int kc, int n_clusters, const SkTune& tune, SkPlan* out) {
  SkPlan s{};
  const long long tiles128 = (rows_m + 127) / 128;
  if (tiles128 < 2 || kc < 1 || n_clusters < 1 || tiles128 > 4096 || n_feat < 1) return false;
  s.kc = kc;
  s.bn = sk_pick_bn(n_feat);
  s.n_nt = (n_feat + s.bn - 1) / s.bn;
  s.n_pair = (int)(tiles128 / 2);
  s.cnt[0] = s.n_pair * s.n_nt;
  s.w[0] = sk_chunk_cost(s.bn, false, tune);
  s.w[1] = s.w[2] = 1;
  if (tiles128 & 1) {
    const int w_h = sk_chunk_cost(s.bn, true, tune);
    const int n_fp = n_feat / 256, rem = n_feat - 256 * n_fp, bnr = (rem + 31) / 32 * 32;
    int w_q = sk_chunk_cost(128, false, tune), w_r = sk_chunk_cost(bnr, true, tune);
    if (w_q < tune.w_q) w_q = tune.w_q;
    if (w_r < tune.w_r_floor) w_r = tune.w_r_floor;
    if (tune.transposed && bnr <= 192 && n_fp * w_q + (rem ? w_r : 0) < s.n_nt * w_h) {
      s.odd = 2;
      s.n_fp = n_fp;
      s.bnr = rem ? bnr : 0;
      s.cnt[1] = n_fp, s.w[1] = w_q;
      s.cnt[2] = rem ? 1 : 0, s.w[2] = w_r;
      if (n_fp == 0) s.cnt[1] = 0, s.w[1] = 1;
    } else {
      s.odd = 1;
      s.cnt[1] = s.n_nt, s.w[1] = w_h;
    }
  }
  long long W = 0, w_max = 1, min_item = 1LL << 40;
  for (int c = 0; c < 3; ++c) {
    if (!s.cnt[c]) continue;
    W += (long long)s.cnt[c] * s.w[c] * kc;
    if (s.w[c] > w_max) w_max = s.w[c];
    if ((long long)s.w[c] * kc < min_item) min_item = (long long)s.w[c] * kc;
  }
  // every cluster must own at least one entry of every item inside its cost range (the reduce kernels enumerate an
  // item's pieces in closed form): the cost per cluster may not fall below the cost of one chunk -> fewer clusters
  if (W / n_clusters < w_max) n_clusters = (int)(W / w_max);
  if (n_clusters < 1) return false;
  s.n_clusters = n_clusters;
  const long long T = (W + n_clusters - 1) / n_clusters;
  if (W >= (1LL << 30) || T > min_item || T < 1) return false;
  s.on = 1;
  s.T = (int)T;
  for (int c = 0; c < 3; ++c) {
    s.L[c] = (int)(((long long)s.w[c] * kc + T - 1) / T);
    if (s.L[c] < 1) s.L[c] = 1;
  }
  int cols = s.bn;  // columns of tensor memory a CTA drains: P bn, H bn/2, Q 128, R bnr/2
  if (s.odd == 2 && s.cnt[1] && cols < 128) cols = 128;
  s.piece_elems = 2 * cols * 128;
  *out = s;
  return true;
}

// (tensor work) cost of the whole plan per chunk of the contraction, same units as sk_chunk_cost
inline long long sk_plan_cost(const SkPlan& s) {
  return (long long)s.cnt[0] * s.w[0] + (long long)s.cnt[1] * s.w[1] + (long long)s.cnt[2] * s.w[2];
}
