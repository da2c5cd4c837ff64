// Backward GEMM of the softmax gradient with the operand roles swapped and an interleaved stream-K decomposition
// (included by softmax_tc.cu inside namespace bhmc, after k_tc_bwd2).
//
// Reference arithmetic: hamiltonian/models/cpu/softmax.py:52-60   grad_w = X^T (P - Y), grad_b = sum_n (P - Y).
//
// k_tc_gemm<MODE_BWD> / k_tc_bwd2 put the D+1 feature rows of X^T on the UMMA M dimension: 785 rows need SEVEN 128-row
// tiles (896 rows, 14 % of the tensor work and 27 MB of zero rows per launch are padding) and an odd tile count leaves
// a phantom half in the cta_group::2 kernel.  Here the chain-class rows of (P-Y)^T are the M operand and the feature
// rows of X^T the N operand:
//   G^T[C*KP, D+1] = DmT[C*KP, rows] . Xt[D+1, rows]^T
//   M = C*KP      64 chains x 10 classes = 640 rows = 5 exact tiles
//   N = D+1       785 -> 5 tiles of 160 = 800 columns (1.9 % padding)
// Every MMA is a cta_group::2 instruction (one kernel may use only one cta_group).  Item types (sk_plan.h):
//   pair item        M = 256: the two CTAs of a cluster hold two adjacent 128-row tiles of DmT and half of the X^T tile each
//   odd last 128-row tile of DmT (5 tiles at 64 chains), one of
//     half items     M = 128, split 64 + 64 over the two CTAs.  Measured: an M = 128 MMA over two CTAs takes as long
//                    as an M = 256 one (cost 0.9-1.0 of a pair chunk), i.e. the odd tile costs as much as a whole pair;
//     transposed     for THIS tile the feature rows go back on M, where pairs exist again: 3 items of 256 feature rows
//                    x 128 chain-class rows (M = 256, N = 128) + one remainder item for the 17 features beyond 768
//                    (M = 128, N = 32, bound by the issue rate of its 13 instructions per chunk): 29 cost units instead
//                    of 45.
// Per CTA and 64-row chunk 52 KB of a pair item enter shared memory instead of the 72 KB of the single-CTA kernel, whose
// main loop is bound by exactly that; the MMA thread of this kernel runs at ~930 cycles per chunk
// (12 MMAs of 80 cycles) against 1250.
//
// Work decomposition.  Items = (row tile pair | half tile) x feature tile, each with all K chunks of the row window.
// 15 items do not divide over 74 clusters, and plain split-K leaves either 14 idle clusters or a second round.  The
// chunks of an item are dealt round-robin to L "lanes" (lane a owns chunks a, a+L, a+2L, ...); the concatenation of all
// lanes of all items, weighted by the cost of a chunk (pair = wp, half = wh), is cut into n_clusters equal pieces.
// A cluster therefore owns a contiguous range of (item, lane, position) entries -- at most two items -- and walks it
// in order of the position, so that at any time ALL clusters work on the same narrow band of chunks of the contraction:
// every operand tile is fetched from HBM once and re-used out of L2 by the other items that need it, as with plain
// split-K, while every cluster gets the same amount of work (stream-K).  A cluster writes one fp32 partial tile per
// item it touched ("piece"); the reduce kernels sum the pieces of an item in cluster order (deterministic).
#pragma once

#include "sk_plan.h"

struct SkParams {
  SkPlan s;
  int stages, split3, sub_chunks;
  int a_chunk0;      // X^T slab of chunk 0 ((row0 - shift) / 64)
  int xt_rows;       // rows per X^T slab (Dt_pad)
  int dm_rows;       // rows per (P-Y)^T slab
  float* part;       // [n_clusters][2 pieces][2 CTAs][columns][128 lanes]
  long long* prof;
};

// tensor maps (all boxes 64 contraction indices wide): DmT with a 128-row box (A of pair items) and a 64-row box (A of
// half / remainder items, B of transposed items); X^T with a bn/2-row box (B of pair / half items), a 128-row box (A of
// transposed items) and a bnr/2-row box (B of the remainder item)
enum { SKM_D128 = 0, SKM_D64 = 2, SKM_XBN = 4, SKM_X128 = 6, SKM_XR = 8 };  // + 1 = the lo copy
struct SkMaps {
  CUtensorMap m[10];
};

// what the roles need to know about an item
struct SkItem {
  int a_map, b_map;       // SKM_*
  int a_rows, b_rows;     // rows of the A / B box of one CTA
  int a_row0, b_row0;     // row of this CTA's box inside the slab of chunk 0 (before the per-chunk slab offset)
  int a_slab, b_slab;     // rows per slab of the A / B operand
  int a_k0, b_k0;         // slab index of chunk 0
  bool a_lo, b_lo;        // lo copies staged (and multiplied)
  int m, n;               // UMMA shape
  int width;              // tensor-memory columns one CTA drains
};
__device__ __forceinline__ SkItem sk_item(const SkParams& p, const SkPiece& pc, int rank) {
  const SkPlan& s = p.s;
  const bool has_dlo = p.split3 != 0, has_xlo = p.split3 == 1;
  const int odd0 = 256 * s.n_pair;  // first chain-class row of the odd tile
  SkItem it;
  const bool a_is_x = pc.type == SK_Q;
  it.a_slab = a_is_x ? p.xt_rows : p.dm_rows;
  it.b_slab = a_is_x ? p.dm_rows : p.xt_rows;
  it.a_k0 = a_is_x ? p.a_chunk0 : 0;
  it.b_k0 = a_is_x ? 0 : p.a_chunk0;
  it.a_lo = a_is_x ? has_xlo : has_dlo;
  it.b_lo = a_is_x ? has_dlo : has_xlo;
  if (pc.type == SK_P) {
    const int mi = pc.idx / s.n_nt, nt = pc.idx % s.n_nt;
    it.a_map = SKM_D128, it.b_map = SKM_XBN;
    it.a_rows = 128, it.b_rows = s.bn / 2;
    it.a_row0 = 256 * mi + 128 * rank, it.b_row0 = nt * s.bn + rank * (s.bn / 2);
    it.m = 256, it.n = s.bn, it.width = s.bn;
  } else if (pc.type == SK_H) {
    it.a_map = SKM_D64, it.b_map = SKM_XBN;
    it.a_rows = 64, it.b_rows = s.bn / 2;
    it.a_row0 = odd0 + 64 * rank, it.b_row0 = pc.idx * s.bn + rank * (s.bn / 2);
    it.m = 128, it.n = s.bn, it.width = s.bn / 2;
  } else if (pc.type == SK_Q) {
    it.a_map = SKM_X128, it.b_map = SKM_D64;
    it.a_rows = 128, it.b_rows = 64;
    it.a_row0 = 256 * pc.idx + 128 * rank, it.b_row0 = odd0 + 64 * rank;
    it.m = 256, it.n = 128, it.width = 128;
  } else {  // SK_R
    it.a_map = SKM_D64, it.b_map = SKM_XR;
    it.a_rows = 64, it.b_rows = s.bnr / 2;
    it.a_row0 = odd0 + 64 * rank, it.b_row0 = 256 * s.n_fp + rank * (s.bnr / 2);
    it.m = 128, it.n = s.bnr, it.width = s.bnr / 2;
  }
  return it;
}

// stage layout (fixed offsets whatever the item type): [A hi | A lo | B hi | B lo], A slots of 128 rows, B slots of
// max(bn/2, 64) rows; a lo slot exists only when some item type stages that copy
struct SkStage {
  int a_lo_off, b_off, b_lo_off, bytes;
};
__host__ __device__ __forceinline__ SkStage sk_stage(const SkPlan& s, int split3) {
  const bool has_dlo = split3 != 0, has_xlo = split3 == 1;
  const bool a_lo = has_dlo || (has_xlo && s.odd == 2), b_lo = has_xlo || (has_dlo && s.odd == 2);
  const int a_slot = BM * BK * 2, b_rows = s.bn / 2 > 64 ? s.bn / 2 : 64, b_slot = b_rows * BK * 2;
  SkStage st;
  st.a_lo_off = a_slot;
  st.b_off = (a_lo ? 2 : 1) * a_slot;
  st.b_lo_off = st.b_off + b_slot;
  st.bytes = st.b_off + (b_lo ? 2 : 1) * b_slot;
  return st;
}

template <int EW>
__global__ void __launch_bounds__(NON_EPI_THREADS + 32 * EW, 1)
k_tc_bwd_sk(const __grid_constant__ SkMaps maps, const SkParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ SkWork work;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const bool leader = rank == 0;
  const int cl = blockIdx.x >> 1;
  const SkStage st = sk_stage(p.s, p.split3);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bar_tfull[b]), 1);
      mbar_init(smem_u32(&bar_tempty[b]), 2 * 32 * EW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    sk_build(p.s, cl, work);
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  pdl_launch_dependents();
  tcgen05_fence_before();
  __syncthreads();  // `work` is CTA-local
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();  // everything above touched only shared / tensor memory

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer (both CTAs) =====
      int stage = 0;
      uint32_t phase = 0;
      for (int pi = 0; pi < work.n_pieces; ++pi) {
        const SkPiece pc = work.piece[pi];
        const SkItem it = sk_item(p, pc, rank);
        const uint32_t tx = (uint32_t)(2 * ((it.a_lo ? 2 : 1) * it.a_rows + (it.b_lo ? 2 : 1) * it.b_rows) * BK * 2);
        const CUtensorMap* ma = &maps.m[it.a_map];
        const CUtensorMap* mb = &maps.m[it.b_map];
        for (int b = pc.b_min; b < pc.b_max; ++b) {
          for (int ri = 0; ri < pc.n_runs; ++ri) {
            int rl, rh;
            sk_run_bounds(pc, ri, &rl, &rh);
            if (b < rl || b >= rh) continue;
            const int k = b * pc.L + pc.a0 + ri;
            mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
            const uint32_t full = smem_u32(&bar_full[stage]);
            if (leader) mbar_expect_tx(full, tx);
            const uint32_t sa = smem_base + stage * st.bytes;
            const int ar = (it.a_k0 + k) * it.a_slab + it.a_row0, br = (it.b_k0 + k) * it.b_slab + it.b_row0;
            tma_load_2d_2sm(sa, ma, full, 0, ar);
            if (it.a_lo) tma_load_2d_2sm(sa + st.a_lo_off, ma + 1, full, 0, ar);
            tma_load_2d_2sm(sa + st.b_off, mb, full, 0, br);
            if (it.b_lo) tma_load_2d_2sm(sa + st.b_lo_off, mb + 1, full, 0, br);
            if (++stage == p.stages) stage = 0, phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {  // ===== MMA issuer (leader CTA; warp-uniform loop, elected lane issues) =====
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      long long t_tempty = 0, t_full = 0, t_issue = 0, t_start = clock64(), n_chunks = 0;
      for (int pi = 0; pi < work.n_pieces; ++pi) {
        const SkPiece pc = work.piece[pi];
        const SkItem im = sk_item(p, pc, 0);
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(im.n >> 3) << 17) | ((uint32_t)(im.m >> 4) << 24);
        // the item type comes out of shared memory, i.e. the compiler cannot know it is warp-uniform: it may only
        // select operands, never guard an MMA (a tcgen05.mma behind a possibly divergent branch is wrapped in an
        // elect / R2UR / BRA.ANY loop: 851 instead of 516 cycles of issue per chunk).  The branches below depend on the
        // kernel parameter split3 alone.
        const bool a_is_x = pc.type == SK_Q;
        // the issuer does not need to know WHICH chunk a stage holds, only how many the piece has: no walk here
        // (with the producer's position walk and a modulo per chunk this thread spent ~500 cycles per chunk outside
        // the barrier wait and the issue, and the tensor pipe starved behind it)
        uint32_t tmem_d = 0;
        int cs = 0;  // chunk inside the current sub-slab: the accumulator is drained every sub_chunks chunks
        for (int ci = 0; ci < pc.n_chunks; ++ci) {
          if (cs == 0) {
            const int buf = it & 1;
            const uint32_t use = (uint32_t)(it >> 1);
            long long c0 = p.prof ? clock64() : 0;
            mbar_wait(smem_u32(&bar_tempty[buf]), (use & 1u) ^ 1u);  // both CTAs' epilogues have drained this accumulator
            tcgen05_fence_after();
            if (p.prof) t_tempty += clock64() - c0;
            tmem_d = tmem_base + (uint32_t)(buf * TMEM_BUF_COLS);
          }
          long long c1 = p.prof ? clock64() : 0;
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tcgen05_fence_after();
          long long c2 = p.prof ? clock64() : 0;
          if (p.prof) t_full += c2 - c1, ++n_chunks;
          const uint32_t sa = smem_base + stage * st.bytes;
          const uint32_t first = cs > 0 ? 1u : 0u;
          const uint64_t a_hi = make_smem_desc(sa), b_hi = make_smem_desc(sa + st.b_off);
          // bf16x3: hi.hi + hi.lo + lo.hi; a lo copy that is identically zero (exact operand) is neither staged nor multiplied
          if (p.split3 == 1) {
            const uint64_t a_lod = make_smem_desc(sa + st.a_lo_off), b_lod = make_smem_desc(sa + st.b_lo_off);
#pragma unroll
            for (int ks = 0; ks < BK / UMMA_K; ++ks) {
              const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);
              umma_bf16_2sm(tmem_d, a_hi + adv, b_hi + adv, idesc, ks > 0 ? 1u : first);
              umma_bf16_2sm(tmem_d, a_hi + adv, b_lod + adv, idesc, 1u);
              umma_bf16_2sm(tmem_d, a_lod + adv, b_hi + adv, idesc, 1u);
            }
          } else if (p.split3 == 2) {  // X exact in bf16: only (P-Y)^T has a lo copy -- the A operand, or B of a transposed item
            const uint64_t a2 = make_smem_desc(a_is_x ? sa : sa + st.a_lo_off);
            const uint64_t b2 = make_smem_desc(a_is_x ? sa + st.b_lo_off : sa + st.b_off);
#pragma unroll
            for (int ks = 0; ks < BK / UMMA_K; ++ks) {
              const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);
              umma_bf16_2sm(tmem_d, a_hi + adv, b_hi + adv, idesc, ks > 0 ? 1u : first);
              umma_bf16_2sm(tmem_d, a2 + adv, b2 + adv, idesc, 1u);
            }
          } else {
#pragma unroll
            for (int ks = 0; ks < BK / UMMA_K; ++ks) {
              const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);
              umma_bf16_2sm(tmem_d, a_hi + adv, b_hi + adv, idesc, ks > 0 ? 1u : first);
            }
          }
          umma_commit_2sm(smem_u32(&bar_empty[stage]), 3);
          if (p.prof) t_issue += clock64() - c2;
          if (++stage == p.stages) stage = 0, phase ^= 1u;
          if (++cs == p.sub_chunks || ci == pc.n_chunks - 1) {
            umma_commit_2sm(smem_u32(&bar_tfull[it & 1]), 3);  // this sub-slab's accumulator is complete in both CTAs
            ++it;
            cs = 0;
          }
        }
      }
      if (p.prof && lane == 0) {
        long long* o = p.prof + (size_t)blockIdx.x * 8;
        o[0] = clock64() - t_start, o[1] = t_tempty, o[2] = t_full, o[3] = t_issue, o[4] = n_chunks, o[5] = it;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs): drain the accumulator every sub_chunks chunks into fp32 registers, store the sum
    // of a piece once.  M = 256 items: TMEM lane t = row t of this CTA's 128-row A tile, columns = the B rows.
    // M = 128 items (64 A rows per CTA): lane t holds A row t % 64, TMEM column j = B row (t / 64) * n/2 + j.
    // Both are stored as [column][lane] (coalesced); the reduce kernels undo the mapping (sk_sum_partials).
    const int ew = warp & 3, part = (warp - 4) >> 2;
    constexpr int PARTS = EW / 4;
    constexpr int MAXCH = 192 / 16 / PARTS;  // 16-column chunks per thread: <= 192 columns with 16 epilogue warps (48 fp32
                                             // accumulators; 64 would not fit the 102-register budget of the 640-thread CTA)
    const int t = ew * 32 + lane;
    int it = 0;
    for (int pi = 0; pi < work.n_pieces; ++pi) {
      const SkPiece pc = work.piece[pi];
      const int width = sk_item(p, pc, rank).width;
      float acc[MAXCH][16];
#pragma unroll
      for (int i = 0; i < MAXCH; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
      const int n_sub = (pc.n_chunks + p.sub_chunks - 1) / p.sub_chunks;
      for (int sb = 0; sb < n_sub; ++sb, ++it) {
        const int buf = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait(smem_u32(&bar_tfull[buf]), use & 1u);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * TMEM_BUF_COLS);
#pragma unroll
        for (int i = 0; i < MAXCH; ++i) {
          const int j0 = (part + i * PARTS) * 16;
          if (j0 < width) {  // warp-uniform
            uint32_t raw[16];
            tmem_ld<16>(tacc + (uint32_t)j0, raw);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[i][j] += __uint_as_float(raw[j]);
          }
        }
        tcgen05_fence_before();
        if (leader) mbar_arrive(smem_u32(&bar_tempty[buf]));
        else mbar_arrive_remote(smem_u32(&bar_tempty[buf]), 0);
      }
      float* dst = p.part + ((size_t)(cl * SK_MAX_PIECES + pc.slot) * p.s.piece_elems) + (size_t)rank * width * 128 + t;
#pragma unroll
      for (int i = 0; i < MAXCH; ++i) {
        const int j0 = (part + i * PARTS) * 16;
        if (j0 < width) {
#pragma unroll
          for (int j = 0; j < 16; ++j) dst[(size_t)(j0 + j) * 128] = acc[i][j];
        }
      }
    }
  }
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// sum over the pieces of the item that holds gradient element (feature d, chain-class column col), in cluster order
__device__ __forceinline__ float sk_sum_partials(const SkPlan& s, const float* __restrict__ part, int d, int col) {
  const int rt = col >> 7;
  int item, off, cls;
  if (rt < 2 * s.n_pair) {  // pair item: lane = row of the CTA's 128-row tile, column = feature inside the tile
    const int nt = d / s.bn, n = d - nt * s.bn;
    cls = 0;
    item = (rt >> 1) * s.n_nt + nt;
    off = ((rt & 1) * s.bn + n) * 128 + (col & 127);
  } else {
    const int cc = col & 127;  // row inside the odd tile
    int n, hb;
    if (s.odd == 2 && d < 256 * s.n_fp) {  // transposed pair: lane = feature inside the CTA's 128, column = cc
      cls = 1;
      item = s.cnt[0] + (d >> 8);
      off = ((((d >> 7) & 1) * 128) + cc) * 128 + (d & 127);
      n = 0, hb = 0;
    } else {
      // M = 128 over two CTAs: CTA rank = cc / 64; TMEM lane = cc % 64 + 64 * (n / (N/2)), column n % (N/2)
      if (s.odd == 2) {
        cls = 2, item = s.cnt[0] + s.cnt[1];
        n = d - 256 * s.n_fp, hb = s.bnr >> 1;
      } else {
        const int nt = d / s.bn;
        cls = 1, item = s.cnt[0] + nt;
        n = d - nt * s.bn, hb = s.bn >> 1;
      }
      const int hi = n / hb;
      off = (((cc >> 6) * hb) + (n - hi * hb)) * 128 + ((cc & 63) + 64 * hi);
    }
  }
  const int S = sk_item_start(s, item), wg = sk_class_w(s, cls);
  const int j_lo = S / s.T, j_hi = (S + (s.kc - 1) * wg) / s.T;
  float v = 0.f;
  int j = j_lo;
  for (; j + 4 <= j_hi + 1; j += 4) {  // loads in batches (a plain loop serialises one memory latency per piece)
    float t[4];
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const int jj = j + x;
      t[x] = part[(size_t)(jj * SK_MAX_PIECES + (jj * s.T < S ? 1 : 0)) * s.piece_elems + off];
    }
#pragma unroll
    for (int x = 0; x < 4; ++x) v += t[x];
  }
  for (; j <= j_hi; ++j) v += part[(size_t)(j * SK_MAX_PIECES + (j * s.T < S ? 1 : 0)) * s.piece_elems + off];
  return v;
}

static bool pdl_enabled();

static int launch_bwd_sk(bhmc_ctx* ctx, const SkMaps& maps, const SkParams& p) {
  const size_t smem = (size_t)p.stages * sk_stage(p.s, p.split3).bytes + 1024;
  static size_t configured = 0;
  if (smem > configured) {
    BHMC_CUDA_OK(cudaFuncSetAttribute(k_tc_bwd_sk<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * p.s.n_clusters));
  cfg.blockDim = dim3(NON_EPI_THREADS + 32 * 16);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  BHMC_CUDA_OK(cudaLaunchKernelEx(&cfg, k_tc_bwd_sk<16>, maps, p));
  ctx->launches++;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}
