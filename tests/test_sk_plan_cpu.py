"""Host-side check of the interleaved stream-K decomposition of the backward GEMM (csrc/sk_plan.h, used by
k_tc_bwd_sk in csrc/softmax_bwd_sk.cuh): the header is plain C++, so it is compiled here with g++ and exercised
exhaustively -- every (item, chunk) must be owned by exactly one cluster, a cluster may touch at most two items, the
reduce kernels' closed-form enumeration of an item's pieces must agree with what the clusters write, and the clusters
must advance through the contraction in lockstep (that is what keeps the operand re-use in L2)."""
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "dropout_hamiltonian_montecarlo_b200", "csrc")

CHECKER = r'''
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include <vector>
#include <algorithm>
#include "sk_plan.h"

static int check(long long rows_m, int n_feat, int kc, int n_clusters, const SkTune& tune, bool expect, double max_skew = 0.5) {
  SkPlan s;
  const bool ok = sk_make_plan(rows_m, n_feat, kc, n_clusters, tune, &s);
  if (ok != expect) { printf("FAIL make_plan(%lld,%d,%d,%d) = %d\n", rows_m, n_feat, kc, n_clusters, (int)ok); return 1; }
  if (!ok) return 0;
  const int n_items = sk_n_items(s);
  std::vector<std::vector<int>> owner(n_items, std::vector<int>(kc, -1));
  std::map<std::pair<int, int>, int> slot_of;  // (cluster, item) -> slot
  long long max_cost = 0;
  double worst_skew = 0;
  for (int cl = 0; cl < s.n_clusters; ++cl) {
    SkWork w;
    sk_build(s, cl, w);
    if (w.n_pieces > SK_MAX_PIECES) { printf("FAIL pieces\n"); return 1; }
    std::set<int> slots;
    long long cost = 0, t = 0;
    double cl_skew = 0;
    for (int pi = 0; pi < w.n_pieces; ++pi) {
      const SkPiece& pc = w.piece[pi];
      if (pc.n_runs < 1 || pc.a0 + pc.n_runs > pc.L) { printf("FAIL runs\n"); return 1; }
      if (!slots.insert(pc.slot).second || pc.slot < 0 || pc.slot >= SK_MAX_PIECES) { printf("FAIL slot\n"); return 1; }
      slot_of[{cl, pc.item}] = pc.slot;
      const int wg = s.w[sk_item_class(s, pc.item)];
      if (pc.type != sk_item_type(s, pc.item)) { printf("FAIL type\n"); return 1; }
      int n = 0;
      for (int b = pc.b_min; b < pc.b_max; ++b) {   // the walk of the kernel's three roles
        for (int ri = 0; ri < pc.n_runs; ++ri) {
          int rl, rh;
          sk_run_bounds(pc, ri, &rl, &rh);
          if (b < rl || b >= rh) continue;
          ++n;
          const int k = b * pc.L + pc.a0 + ri;
          if (k < 0 || k >= kc) { printf("FAIL chunk range\n"); return 1; }
          if (owner[pc.item][k] != -1) { printf("FAIL chunk (%d,%d) owned twice\n", pc.item, k); return 1; }
          owner[pc.item][k] = cl;
          // lockstep: chunk k of the contraction should be processed at fraction ~k/kc of the cluster's time
          const double skew = std::abs((double)(t + wg / 2) / s.T - (double)k / kc);
          cl_skew = std::max(cl_skew, skew);
          t += wg;
        }
      }
      if (n != pc.n_chunks) { printf("FAIL n_chunks\n"); return 1; }
      cost += (long long)n * wg;
    }
    max_cost = std::max(max_cost, cost);
    if (4 * cost >= s.T) worst_skew = std::max(worst_skew, cl_skew);  // (the last cluster may hold a few left-over entries)
  }
  for (int i = 0; i < n_items; ++i)
    for (int k = 0; k < kc; ++k)
      if (owner[i][k] < 0) { printf("FAIL chunk (%d,%d) not owned\n", i, k); return 1; }
  if (max_cost > s.T + std::max(s.w[0], std::max(s.w[1], s.w[2]))) { printf("FAIL balance %lld > %d\n", max_cost, s.T); return 1; }
  // the reduce side: pieces of item i live in clusters j_lo..j_hi, slot = (j*T < S_i)
  for (int i = 0; i < n_items; ++i) {
    const int S = sk_item_start(s, i), wg = s.w[sk_item_class(s, i)];
    const int j_lo = S / s.T, j_hi = (S + (kc - 1) * wg) / s.T;
    std::set<int> want;
    for (int k = 0; k < kc; ++k) want.insert(owner[i][k]);
    std::set<int> got;
    for (int j = j_lo; j <= j_hi; ++j) {
      got.insert(j);
      auto it = slot_of.find({j, i});
      if (it == slot_of.end()) { printf("FAIL reduce reads a piece nobody wrote (item %d cluster %d)\n", i, j); return 1; }
      if (it->second != (j * s.T < S ? 1 : 0)) { printf("FAIL slot formula\n"); return 1; }
    }
    if (want != got) { printf("FAIL piece set of item %d\n", i); return 1; }
  }
  // two pieces of a cluster run one after the other, so a chunk can be off by the share of the shorter piece; what
  // matters is that the band of chunks in flight stays a small part of the contraction
  if (kc >= 256 && worst_skew > max_skew) { printf("FAIL lockstep: skew %.3f (rows %lld feat %d kc %d)\n", worst_skew, rows_m, n_feat, kc); return 1; }
  return 0;
}

// every gradient element (chain-class row, feature) must belong to exactly one item
static int check_tiling(long long rows_m, int n_feat, const SkPlan& s) {
  if (s.n_pair * 256 + (s.odd ? 128 : 0) < rows_m) { printf("FAIL rows not covered\n"); return 1; }
  if (s.n_nt * s.bn < n_feat) { printf("FAIL features not covered (P/H)\n"); return 1; }
  if (s.odd == 2 && s.n_fp * 256 + s.bnr < n_feat) { printf("FAIL features not covered (Q/R)\n"); return 1; }
  if (s.odd == 2 && (s.cnt[1] != s.n_fp || s.cnt[2] != (s.bnr ? 1 : 0))) { printf("FAIL Q/R counts\n"); return 1; }
  if (s.odd == 1 && (s.cnt[1] != s.n_nt || s.cnt[2] != 0)) { printf("FAIL H counts\n"); return 1; }
  if (s.cnt[0] != s.n_pair * s.n_nt) { printf("FAIL P count\n"); return 1; }
  if (s.bn % 32 || s.bn > 192 || s.bnr % 32 || s.bnr > 192) { printf("FAIL widths\n"); return 1; }
  return 0;
}

int main() {
  int bad = 0, n = 0, n_t = 0;
  const int feats[] = {785, 2049, 25, 101, 513, 1025, 300};
  const int kcs[] = {64, 65, 100, 938, 1024, 15625};
  for (int c = 1; c <= 64; ++c)
    for (int kp : {4, 10, 16, 40})
      for (int f : feats)
        for (int kc : kcs)
          for (int pct : {50, 90, 100})
            for (int tr : {0, 1}) {
              const long long rows = (long long)c * kp;
              const SkTune tune{pct, 5, tr, tr ? 8 : 11, tr ? 5 : 8};
              SkPlan s;
              const bool ok = sk_make_plan(rows, f, kc, 74, tune, &s);
              bad += check(rows, f, kc, 74, tune, ok);
              if (ok) bad += check_tiling(rows, f, s), n_t += s.odd == 2;
              n += ok;
            }
  const SkTune dflt{90, 5, 1, 8, 5};
  // the two bench shapes must be eligible: cfg2 (64 chains x 10 classes, 785 feature rows, 938 chunks) and cfg5
  bad += check(640, 785, 938, 74, dflt, true, 0.2);
  bad += check(320, 2049, 15625, 74, dflt, true, 0.25);
  bad += check(100, 785, 938, 74, dflt, false);   // a single 128-row tile: no pair
  SkPlan s;
  sk_make_plan(640, 785, 938, 74, dflt, &s);
  printf("cfg2 plan: pairs %d tiles %d x %d odd %d Q %d R %d weights %d/%d/%d lanes %d/%d/%d T %d cost %lld\n", s.n_pair, s.n_nt, s.bn, s.odd,
         s.n_fp, s.bnr, s.w[0], s.w[1], s.w[2], s.L[0], s.L[1], s.L[2], s.T, sk_plan_cost(s));
  sk_make_plan(320, 2049, 15625, 74, dflt, &s);
  printf("cfg5 plan: pairs %d tiles %d x %d odd %d Q %d R %d weights %d/%d/%d cost %lld\n", s.n_pair, s.n_nt, s.bn, s.odd, s.n_fp, s.bnr,
         s.w[0], s.w[1], s.w[2], sk_plan_cost(s));
  printf("%s %d plans checked (%d with transposed items)\n", bad ? "FAILED" : "OK", n, n_t);
  return bad ? 1 : 0;
}
'''


def test_stream_k_plan_covers_every_chunk_once():
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "sk_check.cpp")
        exe = os.path.join(tmp, "sk_check")
        with open(src, "w") as f:
            f.write(CHECKER)
        r = subprocess.run(["g++", "-O2", "-std=c++17", "-I", CSRC, src, "-o", exe], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
        assert "OK" in r.stdout and "cfg2 plan: pairs 2 tiles 5 x 160 odd 2 Q 3 R 32 weights 10/8/5" in r.stdout, r.stdout
