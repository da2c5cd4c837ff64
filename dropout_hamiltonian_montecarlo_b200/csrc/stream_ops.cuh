// Per-row operations of the streaming schedule, shared by k_stream_update (update.cu) and the fused split-K reduce +
// update kernel (softmax_tc.cu).  hmc.py:51-54 arithmetic; see internal.cuh for the op codes.
#pragma once
#include "internal.cuh"

namespace bhmc {

// scalar bookkeeping of a row: done by exactly one thread per row, BEFORE the row's elementwise work of the phase
__device__ __forceinline__ void stream_row_scalars(const StreamUpdateArgs& a, int r, uint32_t op) {
  a.kin1[r] = 0.0;  // consumed by the previous phase's accept; re-accumulated by this phase's kinetic kernel
  if (op & OP_LATCH) {
    a.stat_cur[r] = a.stat[r];
    a.stat_new[r] = a.stat[r];
    a.kin0[(int64_t)((a.step[r] + 1) & 1) * a.C_total + r] = 0.0;  // the NEXT step's start-of-step buffer
  }
  if (op & OP_LATCH_CACHED) {
    const double s0 = a.stat_next[r];
    a.stat_cur[r] = s0;
    a.stat_new[r] = s0;
    if (a.extra_next) a.extra_cur[r] = a.extra_next[r];
    a.kin0[(int64_t)((a.step[r] + 1) & 1) * a.C_total + r] = 0.0;
  }
  if (op & OP_POST) a.stat_new[r] = a.stat[r];
}

// kicks / drift of four consecutive parameters i..i+3 of a row; returns true when q changed
__device__ __forceinline__ bool stream_apply4(const StreamUpdateArgs& a, uint32_t op, int64_t i, float pe[4], const float ge[4],
                                              float qe[4]) {
  const int pv = (op >> 8) & 15, v = (op >> 12) & 15;
  const int64_t post_off = a.off[pv], post_len = (op & OP_POST) ? a.len[pv] : 0;
  const int64_t pre_off = a.off[v], pre_len = (op & OP_PRE) ? a.len[v] : 0;
  bool moved = false;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int64_t idx = i + e;
    if (idx >= post_off && idx < post_off + post_len) pe[e] = 1.0f * pe[e] - a.a_post * ge[e];
    if (idx >= pre_off && idx < pre_off + pre_len) {
      pe[e] = pe[e] - a.a_pre * ge[e];
      qe[e] = qe[e] + a.eps * pe[e];
      moved = true;
    }
  }
  return moved;
}

}  // namespace bhmc
