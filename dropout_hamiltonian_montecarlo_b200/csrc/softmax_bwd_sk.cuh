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
// Every MMA is a cta_group::2 instruction (one kernel may use only one cta_group):
//   pair item  M = 256: the two CTAs of a cluster hold two adjacent 128-row tiles of DmT and half of the X^T tile each
//   half item  M = 128: an odd last 128-row tile is split 64 + 64 over the two CTAs (no phantom rows: the tensor cores
//                       multiply only the rows that exist)
// Per CTA and 64-row chunk 52 KB (36 KB for a half item) enter shared memory instead of the 72 KB of the single-CTA
// kernel, whose main loop is bound by exactly that.
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
  float* part;       // [n_clusters][2 pieces][2 CTAs][bn][128]
  long long* prof;
};

// tensor maps: DmT with a 128-row box (pair items) and a 64-row box (half items), X^T with a bn/2-row box
template <int EW>
__global__ void __launch_bounds__(NON_EPI_THREADS + 32 * EW, 1)
k_tc_bwd_sk(const __grid_constant__ CUtensorMap tmD_hi, const __grid_constant__ CUtensorMap tmD_lo,
            const __grid_constant__ CUtensorMap tmDh_hi, const __grid_constant__ CUtensorMap tmDh_lo,
            const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo, const SkParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_base_slot;
  __shared__ SkWork work;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const bool leader = rank == 0;
  const int cl = blockIdx.x >> 1;
  // operand split: 1 = DmT hi/lo and X^T hi/lo (3 MMAs), 2 = X exact in bf16 (no X^T lo; 2 MMAs), 0 = single pass
  const bool has_dlo = p.split3 != 0, has_xlo = p.split3 == 1;
  const int d_bytes = BM * BK * 2, xh_bytes = (p.s.bn / 2) * BK * 2;
  const int stage_bytes = (has_dlo ? 2 : 1) * d_bytes + (has_xlo ? 2 : 1) * xh_bytes;  // layout of a stage (pair item)
  const int off_dlo = d_bytes, off_x = (has_dlo ? 2 : 1) * d_bytes, off_xlo = off_x + xh_bytes;

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
        const int mi = pc.item / p.s.n_nt, nt = pc.item % p.s.n_nt;
        const int rows_cta = pc.half ? 64 : 128;
        const int drow = mi * 256 + rank * rows_cta;                 // row of this CTA's DmT tile inside a slab
        const int xrow = nt * p.s.bn + rank * (p.s.bn / 2);          // row of this CTA's half of the X^T tile
        const uint32_t tx = (uint32_t)(2 * ((has_dlo ? 2 : 1) * rows_cta * BK * 2 + (has_xlo ? 2 : 1) * xh_bytes));
        const CUtensorMap* md_hi = pc.half ? &tmDh_hi : &tmD_hi;
        const CUtensorMap* md_lo = pc.half ? &tmDh_lo : &tmD_lo;
        for (int b = pc.b_min; b < pc.b_max; ++b) {
          for (int ri = 0; ri < pc.n_runs; ++ri) {
            int rl, rh;
            sk_run_bounds(pc, ri, &rl, &rh);
            if (b < rl || b >= rh) continue;
            const int k = b * pc.L + pc.a0 + ri;
            mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
            const uint32_t full = smem_u32(&bar_full[stage]);
            if (leader) mbar_expect_tx(full, tx);
            const uint32_t sa = smem_base + stage * stage_bytes;
            const int dr = k * p.dm_rows + drow, xr = (p.a_chunk0 + k) * p.xt_rows + xrow;
            tma_load_2d_2sm(sa, md_hi, full, 0, dr);
            if (has_dlo) tma_load_2d_2sm(sa + off_dlo, md_lo, full, 0, dr);
            tma_load_2d_2sm(sa + off_x, &tmX_hi, full, 0, xr);
            if (has_xlo) tma_load_2d_2sm(sa + off_xlo, &tmX_lo, full, 0, xr);
            if (++stage == p.stages) stage = 0, phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {  // ===== MMA issuer (leader CTA; warp-uniform loop, elected lane issues) =====
      const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.s.bn >> 3) << 17);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      long long t_tempty = 0, t_full = 0, t_issue = 0, t_start = clock64(), n_chunks = 0;
      for (int pi = 0; pi < work.n_pieces; ++pi) {
        const SkPiece pc = work.piece[pi];
        const uint32_t idesc = idesc0 | ((uint32_t)((pc.half ? 128 : 256) >> 4) << 24);
        // the issuer does not need to know WHICH chunk a stage holds, only how many the piece has: no walk here
        // (with the producer's position walk and a modulo per chunk this thread spent ~500 cycles per chunk outside
        // the barrier wait and the issue, and the tensor pipe starved behind it)
        uint32_t tmem_d = 0;
        int cs = 0;  // chunk inside the current sub-slab: the accumulator is drained every sub_chunks chunks
        {
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
            const uint32_t sa = smem_base + stage * stage_bytes;
            const uint32_t first = cs > 0 ? 1u : 0u;
            const uint64_t d_hi = make_smem_desc(sa), x_hi = make_smem_desc(sa + off_x);
            if (p.split3 == 1) {
              const uint64_t d_lo = make_smem_desc(sa + off_dlo), x_lo = make_smem_desc(sa + off_xlo);
#pragma unroll
              for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);
                umma_bf16_2sm(tmem_d, d_hi + adv, x_hi + adv, idesc, ks > 0 ? 1u : first);
                umma_bf16_2sm(tmem_d, d_hi + adv, x_lo + adv, idesc, 1u);
                umma_bf16_2sm(tmem_d, d_lo + adv, x_hi + adv, idesc, 1u);
              }
            } else if (p.split3 == 2) {  // X exact in bf16: lo(X) == 0, its product is not issued
              const uint64_t d_lo = make_smem_desc(sa + off_dlo);
#pragma unroll
              for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);
                umma_bf16_2sm(tmem_d, d_hi + adv, x_hi + adv, idesc, ks > 0 ? 1u : first);
                umma_bf16_2sm(tmem_d, d_lo + adv, x_hi + adv, idesc, 1u);
              }
            } else {
#pragma unroll
              for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);
                umma_bf16_2sm(tmem_d, d_hi + adv, x_hi + adv, idesc, ks > 0 ? 1u : first);
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
      }
      if (p.prof && lane == 0) {
        long long* o = p.prof + (size_t)blockIdx.x * 8;
        o[0] = clock64() - t_start, o[1] = t_tempty, o[2] = t_full, o[3] = t_issue, o[4] = n_chunks, o[5] = it;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs): drain the accumulator every sub_chunks chunks into fp32 registers, store the sum
    // of a piece once.  TMEM lane t of a pair item = row t of this CTA's 128-row tile, columns = features of the tile;
    // of a half item (M = 128 over two CTAs): row t % 64 of this CTA's 64 rows, TMEM column j = feature
    // (t / 64) * bn/2 + j.  Both are stored as [column][lane] (coalesced); the reduce kernels undo the mapping.
    const int ew = warp & 3, part = (warp - 4) >> 2;
    constexpr int PARTS = EW / 4;
    constexpr int MAXCH = 192 / 16 / PARTS;  // 16-column chunks per thread: bn <= 192 with 16 epilogue warps (48 fp32
                                             // accumulators; 64 would not fit the 102-register budget of the 640-thread CTA)
    const int t = ew * 32 + lane;
    int it = 0;
    for (int pi = 0; pi < work.n_pieces; ++pi) {
      const SkPiece pc = work.piece[pi];
      const int width = pc.half ? p.s.bn / 2 : p.s.bn;
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
  int mi, off;
  const int nt = d / s.bn, n = d - nt * s.bn;
  bool half;
  if (rt < 2 * s.n_pair) {
    half = false;
    mi = rt >> 1;
    off = ((rt & 1) * s.bn + n) * 128 + (col & 127);
  } else {  // half item: CTA rank = row / 64; TMEM lane = row % 64 + 64 * (n / (bn/2)), column n % (bn/2)
    half = true;
    mi = s.n_pair;
    const int hb = s.bn >> 1, hi = n / hb;
    off = ((((col & 127) >> 6) * hb) + (n - hi * hb)) * 128 + ((col & 63) + 64 * hi);
  }
  const int item = mi * s.n_nt + nt;
  const int S = sk_item_start(s, item), wg = half ? s.wh : s.wp;
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

static int launch_bwd_sk(bhmc_ctx* ctx, const CUtensorMap& d_hi, const CUtensorMap& d_lo, const CUtensorMap& dh_hi,
                         const CUtensorMap& dh_lo, const CUtensorMap& x_hi, const CUtensorMap& x_lo, const SkParams& p) {
  const int stage_bytes = (p.split3 ? 2 : 1) * BM * BK * 2 + (p.split3 == 1 ? 2 : 1) * (p.s.bn / 2) * BK * 2;
  const size_t smem = (size_t)p.stages * stage_bytes + 1024;
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
  BHMC_CUDA_OK(cudaLaunchKernelEx(&cfg, k_tc_bwd_sk<16>, d_hi, d_lo, dh_hi, dh_lo, x_hi, x_lo, p));
  ctx->launches++;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}
