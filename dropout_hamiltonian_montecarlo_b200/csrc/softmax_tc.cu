// BHMC_PREC_BF16X3 / BHMC_PREC_BF16: the two contractions of the softmax-regression gradient
// on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by
// TMA into 128B-swizzled shared memory), chains batched on the GEMM N dimension.
//
// Reference arithmetic (hamiltonian/models/cpu/softmax.py):
//   forward  :38-43,63-72  Z = X.W + b, clip, row softmax / log-sum-exp      -> k_tc_gemm<MODE_FWD>
//   backward :52-60        grad = X^T (P - Y) + alpha W, sum_n (P - Y) + alpha b -> k_tc_gemm<MODE_BWD> + k_tc_reduce
//
// Layout in HBM (all bf16 operands are "K-major": the contraction index is contiguous)
//   Xa_{hi,lo} [N, Dp]        forward A, built once at bind time   (Dp = D rounded up to 8)
//   Xt_{hi,lo} [Npad/S][Dt_pad, S]  backward A = X^T plus a row of ones that yields the bias gradient, stored in
//                             slabs of S rows of X.  Default S = BK = 64 with an unpadded row stride = the fully
//                             BLOCKED layout: every TMA box (128 feature rows x one 64-row chunk) is one contiguous
//                             16 KB block.  (Plain [D+1, N] rows put every feature row of a tile into a different
//                             2 MB page at N = 1e6 and thrashed the TLB; S = 8192 slabs fixed that; blocks are
//                             another 10 % faster for launches that carry few chains.)
//   Wt_{hi,lo} [C*KP, Dp]     forward B, rebuilt from the fp32 chain state before every evaluation
//   DmT_{hi,lo}[slabs][Nt*BN, Sd]  (P - Y)^T written by the forward epilogue, backward B; same scheme as Xt
//                             (Sd = BK: [chunk][class row][64 window rows], a chain's classes are 128 B apart)
//   part [S, Mt*128, Nt*BN]   fp32 split-K partials of the backward GEMM (deterministic reduce)
// bf16x3: every fp32 value v is split v = hi + lo (two bf16); a product uses 3 MMAs
// (hi*hi + hi*lo + lo*hi) accumulated in fp32, which restores ~fp32 accuracy.
//
// Kernel structure (one persistent CTA per SM, 256 threads):
//   warp 0 : TMA producer  (cp.async.bulk.tensor -> smem ring, mbarrier expect_tx)
//   warp 1 : MMA issuer    (one thread: tcgen05.mma, tcgen05.commit -> ring "empty" / TMEM "full")
//   warp 2 : TMEM allocator
//   warps 4-7 : epilogue   (tcgen05.ld -> registers -> fused math -> global)
// Accumulators are double buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps
// the main loop of tile i+1.
#include <cuda_bf16.h>

#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "internal.cuh"
#include "philox.cuh"
#include "stream_ops.cuh"

namespace bhmc {

static constexpr int BM = 128;        // UMMA M
// K extent of one pipeline stage.  64 elements = 128 B rows (SWIZZLE_128B); 32 elements = 64 B rows (SWIZZLE_64B).
// Both are implemented and parity-tested; measured on B200 the half-size stages (twice the ring depth) are SLOWER
// (forward 222 us vs 185 us at cfg2): the main loop pays ~450 cycles per chunk that do not overlap with the MMAs,
// so fewer, larger chunks win.  Build with -DBHMC_BK=32 to reproduce.
#ifndef BHMC_BK
#define BHMC_BK 64
#endif
static constexpr int BK = BHMC_BK;
static_assert(BK == 32 || BK == 64, "BK must be 32 (SWIZZLE_64B) or 64 (SWIZZLE_128B)");
static constexpr int UMMA_K = 16;
static constexpr int MAX_STAGES = 8;
static constexpr int NON_EPI_THREADS = 128;  // warps 0-3: TMA, MMA, TMEM alloc, spare
static constexpr int TMEM_COLS = 512;
static constexpr int TMEM_BUF_COLS = 256;
static constexpr float CLIP_HI = 36.04365338911715f;
static constexpr float CLIP_LO = -708.3964185322641f;

enum { MODE_FWD = 0, MODE_BWD = 1 };

#include "tc_common.cuh"

// ---------------------------------------------------------------------------------------------
struct TcParams {
  // work decomposition
  int m_tiles, n_tiles, n_split;  // work items = n_split * m_tiles * n_tiles
  int k_chunks;                   // BK-chunks of the contraction dimension (total)
  int chunks_per_split;
  int sub_chunks;                 // accumulation chain length in BK-chunks: the TMEM accumulator is drained into
                                  // fp32 registers (round-to-nearest adds) every sub_chunks chunks (backward only)
  int BN;                         // UMMA N (multiple of 16, <= 256)
  int stages;
  int split3;                     // 1: hi/lo operands, 3 MMAs per product; 2: A is exact in bf16 (no lo copy is
                                  // staged), 2 MMAs per product; 0: single pass
  int a_k0, a_m0;                 // coordinate offsets of A in its tensor map (contraction, row)
  int b_slab, b_slab_rows;        // same for B
  int hint_a, hint_b;             // L2 eviction priority of the operand streams (make_l2_policy)
  int a_slab, a_slab_rows;        // backward A is stored in slabs of a_slab contraction indices (0 = plain matrix):
                                  // element (m, k) lives at row (k / a_slab) * a_slab_rows + m, column k % a_slab
  int pair;                       // CTA pairs (cluster of 2) sharing one operand through TMA multicast:
                                  // 0 = off, 1 = two M tiles share the B tile, 2 = two N tiles share the A tile
  // forward epilogue
  int C, K, cpt;                  // chains, classes, chains per N tile (BN = cpt*KP)
  int D;
  int64_t ld;                     // chain row stride of q
  const float* q;                 // bias lives at q[c*ld + D*K + k]
  const int32_t* labels;          // already offset to the row window
  int64_t nrows;
  int dm_slab, dm_slab_rows;      // DmT slabs: window column j of row i lives at row (j / dm_slab) * dm_slab_rows + i,
  int dm_ld;                      // column j % dm_slab; row stride dm_ld = dm_slab + 64 (not a power of two)
  int dm_tail;                    // columns after the last written row that the backward's last chunk still reads
  // Z cache: X.W (before the bias) of every (row, chain, class), fp32, transposed like DmT (same slab width / row
  // stride, its own rows-per-slab fixed when it was stored).  Forward kernels store it when zt != nullptr; the
  // evaluation after a sub-step that moved only the bias rebuilds softmax / (P-Y)^T / log-lik from it (k_softmax_from_z)
  float* zt;
  int zt_slab_rows;
  int pack_dm;                    // forward epilogue: 4-byte (row pair) stores of (P-Y)^T (dm_shift even, dm_ld even; BHMC_FWD_PACK)
  int dm_shift;                   // DmT column of window row 0 (row0 % BK: backward chunks start on absolute multiples of BK)
  __nv_bfloat16* dmt_hi;
  __nv_bfloat16* dmt_lo;          // nullptr in single-pass mode
  double* loglik;
  int write_dm;
  int skip_loglik;                // persistent minibatch kernel: no per-step log-likelihood (it is evaluated once per epoch)
  long long* prof;                // optional [grid][8] cycle counters (BHMC_PROF=1): see tools/profile_grad.py
  int debug;                      // BHMC_DEBUG_EPI (measurement only): 1 = skip the forward epilogue, 2 = skip its atomics
  // backward epilogue
  float* part;                    // [n_split, m_tiles*128, n_tiles*BN]
  // k_tc_bwd2 only: an odd last row tile is processed AFTER the main items by the same launch as a pair whose second
  // half is a phantom tile (its rows are out of range and its results are dropped), with its own row-slab split
  int prefetch;                   // k_tc_bwd2: L2 prefetch distance in chunks (0 = off)
  int left_n_split, left_cps;     // 0 = no odd tile
  float* part_left;               // [left_n_split, 128, n_tiles*BN]
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Forward epilogue of one 128 x BN accumulator tile (softmax.py:38-43,52,63-72): thread = row, the EW/4 warps of
// a TMEM lane quarter split the tile's chains.  Per chain: tcgen05.ld KP columns, + bias, clip, row max, exp2,
// sum, P - Y split to bf16 hi/lo and stored transposed, z_y - logsumexp warp-reduced into one fp64 red.
template <int KP, int EW, bool EXACT, bool FROM_Z = false>
__device__ __forceinline__ void fwd_epilogue_tile(const TcParams& p, uint32_t tacc, int mt, int nt, int part, int lane,
                                                  int t, int row_shift = 0, int live = BM) {
  // (row_shift, live): the accumulator holds only `live` rows, those of window rows mt*BM + row_shift + (0..live-1) -- the
  // half tiles of the persistent minibatch kernel (live = 64: the warps of TMEM lanes 64..127 only do the zero fill)
  constexpr int PARTS = EW / 4;
  const float L2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  const int64_t r = (int64_t)mt * BM + row_shift + t;  // row inside the window
  const bool alive = t < live;                          // warp-uniform (live is a multiple of 32)
  const bool valid = alive && r < p.nrows;
  const int y = valid ? p.labels[r] : -1;
  const int K = EXACT ? KP : p.K;  // EXACT: no padded classes, every class loop is branch-free
  // DmT position of this thread's row (chain-independent part): slab, column inside the slab
  const int dm_col = p.dm_shift + (int)r, dm_sl = dm_col / p.dm_slab;
  const int64_t dm_off = (int64_t)dm_sl * p.dm_slab_rows * p.dm_ld + (dm_col - dm_sl * p.dm_slab);
  const int64_t zt_off = (int64_t)dm_sl * p.zt_slab_rows * p.dm_ld + (dm_col - dm_sl * p.dm_slab);
  for (int cc = part; cc < p.cpt; cc += PARTS) {
    const int c = nt * p.cpt + cc;
    if (c >= p.C) break;  // warp-uniform
    const float* bias = p.q + (int64_t)c * p.ld + (int64_t)p.D * K;
    float ll = 0.f;
    if (!alive) {
    } else if constexpr (KP >= 24 && !FROM_Z) {
      // Wide class counts (KP = 24 / 40 / 64; cfg5 has K = 38): holding a whole row of logits (+ exps) per thread does
      // not fit the 96-register budget of the 640-thread CTA and spilled (ptxas: 80-156 B at KP = 40, ~900 B at
      // KP = 64).  Tensor memory is cheap to re-read, so the row is walked three times in 8-column pieces instead:
      // max, sum of exps, then P - Y.  Twice the ex2 work, no local memory.
      constexpr int CH = 8;
      const uint32_t tcol = tacc + (uint32_t)(cc * KP);
      float m = -INFINITY, zy = 0.f;
      float* zp = p.zt ? p.zt + zt_off + (int64_t)c * KP * p.dm_ld : nullptr;
#pragma unroll 1
      for (int k0 = 0; k0 < KP; k0 += CH) {
        uint32_t r8[CH];
        tmem_ld<CH>(tcol + (uint32_t)k0, r8);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const int k = k0 + j;
          if (EXACT || k < K) {
            if (zp) __stcs(zp + (int64_t)k * p.dm_ld, __uint_as_float(r8[j]));
            const float v = fminf(__uint_as_float(r8[j]) + __ldcg(bias + k), CLIP_HI);
            m = fmaxf(m, v);
            zy = (k == y) ? v : zy;
          }
        }
      }
      zy = fmaxf(zy, CLIP_LO);
      const bool all_low = m < CLIP_LO;  // every logit below the lower clip: the literal path (all classes at CLIP_LO)
      if (all_low) m = CLIP_LO;
      float ssum = 0.f;
#pragma unroll 1
      for (int k0 = 0; k0 < KP; k0 += CH) {
        uint32_t r8[CH];
        tmem_ld<CH>(tcol + (uint32_t)k0, r8);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const int k = k0 + j;
          if (EXACT || k < K) {
            const float v = all_low ? CLIP_LO : fminf(__uint_as_float(r8[j]) + __ldcg(bias + k), CLIP_HI);
            ssum += ex2_approx((v - m) * L2E);
          }
        }
      }
      const float inv = __fdividef(1.0f, ssum);
      ll = valid ? ((zy - m) - LN2 * lg2_approx(ssum)) : 0.f;
      if (p.write_dm) {
        const int64_t o = dm_off + (int64_t)c * KP * p.dm_ld;
#pragma unroll 1
        for (int k0 = 0; k0 < KP; k0 += CH) {
          uint32_t r8[CH];
          tmem_ld<CH>(tcol + (uint32_t)k0, r8);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int k = k0 + j;
            if (EXACT || k < K) {
              const float v = all_low ? CLIP_LO : fminf(__uint_as_float(r8[j]) + __ldcg(bias + k), CLIP_HI);
              const float e = ex2_approx((v - m) * L2E);
              const float d = valid ? fmaf(e, inv, (k == y) ? -1.f : 0.f) : 0.f;  // P - Y
              const __nv_bfloat16 h = __float2bfloat16_rn(d);
              p.dmt_hi[o + (int64_t)k * p.dm_ld] = h;
              if (p.split3) p.dmt_lo[o + (int64_t)k * p.dm_ld] = __float2bfloat16_rn(d - __bfloat162float(h));
            }
          }
        }
      }
    } else {
      uint32_t raw[KP];
      if constexpr (FROM_Z) {
        const float* zp = p.zt + zt_off + (int64_t)c * KP * p.dm_ld;
#pragma unroll
        for (int k = 0; k < KP; ++k) raw[k] = (EXACT || k < K) ? __float_as_uint(__ldcs(zp + (int64_t)k * p.dm_ld)) : 0u;
      } else {
        tmem_ld_cols<KP>(tacc + (uint32_t)(cc * KP), raw);
      }
      // Up to 16 classes the bias is fetched while the tcgen05.ld is in flight.  Wider class counts (KP = 24 / 40 / 64:
      // cfg5's K = 38) would keep raw[KP] + bv[KP] live at once and spill under the 96-register budget of the
      // 640-thread CTA (ptxas: 60-84 B at KP = 40), so there the bias is read where it is consumed (L1 hits).
      constexpr bool PRELOAD_BIAS = KP <= 16;
      float bv[PRELOAD_BIAS ? KP : 1];
      if constexpr (PRELOAD_BIAS) {
#pragma unroll
        for (int k = 0; k < KP; ++k) bv[k] = (EXACT || k < K) ? __ldcg(bias + k) : 0.f;
      }
      if constexpr (!FROM_Z) {
        tmem_ld_wait();
        if (p.zt) {  // keep X.W for the next evaluation (streaming stores: read back once, by another kernel)
          float* zp = p.zt + zt_off + (int64_t)c * KP * p.dm_ld;
#pragma unroll
          for (int k = 0; k < KP; ++k)
            if (EXACT || k < K) __stcs(zp + (int64_t)k * p.dm_ld, __uint_as_float(raw[k]));
        }
      }
      float z[KP];
      float m = -INFINITY, zy = 0.f;
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        // softmax.py:40-41 clips to [-708.4, 36.04].  The lower clip only matters for the label's logit
        // (exp(z - max) underflows to 0 in fp32 either way), so it is applied to zy alone below.
        float bk;
        if constexpr (PRELOAD_BIAS) bk = bv[k];
        else bk = (EXACT || k < K) ? __ldcg(bias + k) : 0.f;
        float v = fminf(__uint_as_float(raw[k]) + bk, CLIP_HI);
        if (!EXACT && k >= K) v = -INFINITY;  // padded classes
        z[k] = v;
        m = fmaxf(m, v);
        zy = (k == y) ? v : zy;
      }
      zy = fmaxf(zy, CLIP_LO);
      if (m < CLIP_LO) {  // every logit below the lower clip (diverged chain): take the slow, literal path
        m = CLIP_LO;
#pragma unroll
        for (int k = 0; k < KP; ++k)
          if (EXACT || k < K) z[k] = CLIP_LO;
      }
      float ssum = 0.f;
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        z[k] = ex2_approx((z[k] - m) * L2E);  // exp(clip(z) - max); 0 for padded classes
        ssum += z[k];
      }
      const float inv = __fdividef(1.0f, ssum);
      ll = valid ? ((zy - m) - LN2 * lg2_approx(ssum)) : 0.f;  // z_y - logsumexp(z)
      if (p.write_dm) {
        const int64_t o = dm_off + (int64_t)c * KP * p.dm_ld;
        if (!FROM_Z && KP % 2 == 0 && p.pack_dm) {
          // Rows r and r + 1 (lanes l, l ^ 1; r even) are neighbours in the transposed layout: one 4-byte store per class and
          // copy instead of two 2-byte stores.  A lane pair exchanges (hi | lo << 16) words -- the even lane stores the even
          // classes, the odd lane the odd ones -- so the store instructions of the epilogue halve (they share the
          // L1 / shared-memory port with the operand traffic of the main loop).  Needs an even dm_shift (host).
          const bool odd = lane & 1;
          const int64_t oe = o - (odd ? 1 : 0);  // the even row's column
#pragma unroll
          for (int k0 = 0; k0 < KP; k0 += 2) {
            uint32_t pk[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int k = k0 + j;
              const float d = valid ? fmaf(z[k], inv, (k == y) ? -1.f : 0.f) : 0.f;  // P - Y
              const __nv_bfloat16 h = __float2bfloat16_rn(d);
              const __nv_bfloat16 l = __float2bfloat16_rn(d - __bfloat162float(h));
              pk[j] = (uint32_t)__bfloat16_as_ushort(h) | ((uint32_t)__bfloat16_as_ushort(l) << 16);
            }
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, odd ? pk[0] : pk[1], 1);  // the partner's word of MY class
            const uint32_t mine = odd ? pk[1] : pk[0];
            const uint32_t E = odd ? recv : mine, O = odd ? mine : recv;                 // even row, odd row
            const int ks = k0 + (odd ? 1 : 0);
            if (EXACT || ks < K) {
              *reinterpret_cast<uint32_t*>(p.dmt_hi + oe + (int64_t)ks * p.dm_ld) = __byte_perm(E, O, 0x5410);
              if (p.split3) *reinterpret_cast<uint32_t*>(p.dmt_lo + oe + (int64_t)ks * p.dm_ld) = __byte_perm(E, O, 0x7632);
            }
          }
        } else {
        __nv_bfloat16* dh = p.dmt_hi + o;
        __nv_bfloat16* dl = p.dmt_lo + o;  // only dereferenced in split mode
#pragma unroll
        for (int k = 0; k < KP; ++k) {
          if (EXACT || k < K) {
            float d = valid ? fmaf(z[k], inv, (k == y) ? -1.f : 0.f) : 0.f;  // P - Y
            __nv_bfloat16 h = __float2bfloat16_rn(d);
            *dh = h;
            if (p.split3) *dl = __float2bfloat16_rn(d - __bfloat162float(h));
          }
          dh += p.dm_ld;
          dl += p.dm_ld;
        }
        }
      }
    }
    if (p.write_dm) {
      // zero what the backward chunks read but no row writes: the alignment prefix (columns before the window) and
      // the columns between the last tile row and the end of the last BK-chunk
      int zc = -1;
      if (mt == 0 && row_shift == 0 && t < p.dm_shift) zc = t;  // dm_shift, dm_tail < 64: threads 0..63 / 64..127 of the tile
      else if (mt == p.m_tiles - 1 && (live == BM || row_shift != 0) && t >= 64 && t - 64 < p.dm_tail)
        zc = p.dm_shift + p.m_tiles * BM + (t - 64);
      if (zc >= 0) {
        const int zs = zc / p.dm_slab;
        const int64_t zo = (zs * p.dm_slab_rows + (int64_t)c * KP) * p.dm_ld + (zc - zs * p.dm_slab);
        for (int k = 0; k < K; ++k) {
          p.dmt_hi[zo + (int64_t)k * p.dm_ld] = __float2bfloat16_rn(0.f);
          if (p.split3) p.dmt_lo[zo + (int64_t)k * p.dm_ld] = __float2bfloat16_rn(0.f);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ll += __shfl_xor_sync(0xffffffffu, ll, o);
    if constexpr (FROM_Z) {  // 4 warps, one chain per block: one atomic per block
      __shared__ float sll[4];
      if (lane == 0) sll[t >> 5] = ll;
      __syncthreads();
      if (t == 0) atomicAdd(p.loglik + c, quantize_addend<24>((double)sll[0] + (double)sll[1] + (double)sll[2] + (double)sll[3]));
    } else {
      if (lane == 0 && p.debug != 2 && !p.skip_loglik) atomicAdd(p.loglik + c, quantize_addend<24>((double)ll));  // order-independent
    }
  }
}

// The evaluation after a sub-step that moved only the bias (Gauss-Seidel sweep, hmc.py:50-53): X.W is unchanged, so
// the forward GEMM is replaced by this HBM-bound pass over the cached Z^T: + new bias, clip, softmax, (P-Y)^T hi/lo,
// log-likelihood -- the forward epilogue with global memory in place of tensor memory.  Block = 128 rows of ONE
// chain (grid = row tiles x chains): with 4 chains per thread the kernel ran at 1.9 TB/s, every warp waiting a full
// DRAM latency on its 10 loads (ncu: 41 % of the samples on the first use); one chain per thread and 12-16 resident
// blocks per SM keep ~4x the bytes in flight.
template <int KP, bool EXACT>
__global__ void __launch_bounds__(128, (KP <= 16 ? 12 : 4)) k_softmax_from_z(const TcParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // blockIdx.x = chain (fastest): concurrently resident blocks spread their log-likelihood atomics over all chains
  fwd_epilogue_tile<KP, 4, EXACT, true>(p, 0u, (int)blockIdx.y, (int)blockIdx.x, 0, lane, warp * 32 + lane);  // p.cpt == 1
}

// Vectorised form of k_softmax_from_z (default).  The first version (thread = one row of one chain, scalar loads and
// 2-byte stores) ran at 3.6 TB/s = 55 % of the copy peak and was ISSUE-bound (ncu r01: smsp issue active 80 %, 60 M warp
// instructions per launch) -- per element: one 4-byte load, two 2-byte stores, address arithmetic and the class
// predicates.  Here a thread owns VEC consecutive DmT columns (= window rows) of one chain: one 16-byte load per class
// (VEC = 4), one 8-byte store per class and operand half, index arithmetic paid once per VEC rows.  It walks the DmT
// COLUMNS (alignment prefix and chunk tail included: those get zeros), so no separate zero-fill is needed.
// grid = (chains, column groups); block = 128 threads = 128 * VEC columns of one chain.
template <int VEC> struct ZVec;
template <> struct ZVec<4> { using f = float4; using h = uint2; };
template <> struct ZVec<2> { using f = float2; using h = uint32_t; };
template <int KP, bool EXACT, int VEC>
__global__ void __launch_bounds__(128) k_softmax_from_z_vec(const TcParams p, int n_cols) {
  using FV = typename ZVec<VEC>::f;
  using HV = typename ZVec<VEC>::h;
  static_assert(VEC % 2 == 0, "bf16 pairs");
  // logits are kept in the log2 domain (x * log2 e): one FFMA adds the bias and scales, ex2 needs no multiply
  const float L2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
  const float HI2 = CLIP_HI * L2E, LO2 = CLIP_LO * L2E;
  const int c = blockIdx.x;
  const int K = EXACT ? KP : p.K;
  const int col0 = ((int)blockIdx.y * 128 + (int)threadIdx.x) * VEC;  // first DmT column of this thread
  float ll = 0.f;
  if (col0 < n_cols) {
    const int sl = col0 / p.dm_slab, cin = col0 - sl * p.dm_slab;  // dm_slab % 64 == 0 and col0 % VEC == 0: one slab
    const float* zp = p.zt + ((int64_t)sl * p.zt_slab_rows + (int64_t)c * KP) * p.dm_ld + cin;
    const int64_t doff = ((int64_t)sl * p.dm_slab_rows + (int64_t)c * KP) * p.dm_ld + cin;
    const float* bias = p.q + (int64_t)c * p.ld + (int64_t)p.D * K;
    int y[VEC];
    bool valid[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int r = col0 + j - p.dm_shift;
      valid[j] = r >= 0 && r < p.nrows;
      y[j] = valid[j] ? __ldg(p.labels + r) : -1;
    }
    FV raw[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) {  // all loads of the thread in flight before the first use
      if (EXACT || k < K) raw[k] = __ldcs(reinterpret_cast<const FV*>(zp));
      zp += p.dm_ld;
    }
    float z[KP][VEC];
    float m[VEC], zy[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) m[j] = -INFINITY, zy[j] = 0.f;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      if (EXACT || k < K) {
        const float b2 = __ldcg(bias + k) * L2E;
        const float* vv = reinterpret_cast<const float*>(&raw[k]);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float x = fminf(fmaf(vv[j], L2E, b2), HI2);  // (z + b) log2 e, upper clip (softmax.py:40)
          z[k][j] = x;
          m[j] = fmaxf(m[j], x);
          if (k == y[j]) zy[j] = x;
        }
      }
    }
    float inv[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      zy[j] = fmaxf(zy[j], LO2);
      const bool all_low = m[j] < LO2;  // every logit below the lower clip: the literal path (softmax.py:41)
      if (all_low) m[j] = LO2;
      float ssum = 0.f;
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        if (EXACT || k < K) {
          const float e = ex2_approx((all_low ? LO2 : z[k][j]) - m[j]);
          z[k][j] = e;
          ssum += e;
        }
      }
      // columns outside the window (alignment prefix, chunk tail) read whatever the scratch buffer held: fminf() above
      // maps a NaN to the clip value, so every e is finite and ssum >= 1; inv = 0 (and y = -1) then gives exact zeros
      inv[j] = valid[j] ? __fdividef(1.0f, ssum) : 0.f;
      ll += valid[j] ? LN2 * ((zy[j] - m[j]) - lg2_approx(ssum)) : 0.f;  // z_y - logsumexp(z)
    }
    if (p.write_dm) {
      __nv_bfloat16* dh = p.dmt_hi + doff;
      __nv_bfloat16* dl = p.dmt_lo + doff;  // only dereferenced in split mode
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        if (EXACT || k < K) {
          __nv_bfloat162 hi[VEC / 2], lo[VEC / 2];
#pragma unroll
          for (int j = 0; j < VEC; j += 2) {
            float d0 = z[k][j] * inv[j], d1 = z[k][j + 1] * inv[j + 1];  // P - Y
            if (k == y[j]) d0 -= 1.f;
            if (k == y[j + 1]) d1 -= 1.f;
            hi[j / 2] = __floats2bfloat162_rn(d0, d1);
            lo[j / 2] = __floats2bfloat162_rn(d0 - __low2float(hi[j / 2]), d1 - __high2float(hi[j / 2]));
          }
          *reinterpret_cast<HV*>(dh) = *reinterpret_cast<const HV*>(hi);
          if (p.split3) *reinterpret_cast<HV*>(dl) = *reinterpret_cast<const HV*>(lo);
        }
        dh += p.dm_ld;
        dl += p.dm_ld;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ll += __shfl_xor_sync(0xffffffffu, ll, o);
  __shared__ float sll[4];
  if ((threadIdx.x & 31) == 0) sll[threadIdx.x >> 5] = ll;
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(p.loglik + c, quantize_addend<24>((double)sll[0] + (double)sll[1] + (double)sll[2] + (double)sll[3]));
}

// EW = number of epilogue warps (multiple of 4).  Warp w may only touch TMEM lanes 32*(w%4)..+31, so the
// EW/4 warps that share a lane quarter split the tile's chains (forward) / column chunks (backward).
template <int MODE, int KP, int EW, bool EXACT>
__global__ void __launch_bounds__(NON_EPI_THREADS + 32 * EW, 1)
k_tc_gemm(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
          const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B needs 1024 B alignment
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a_bytes = BM * BK * 2, b_bytes = p.BN * BK * 2;
  const int na = p.split3 == 1 ? 2 : 1, nb = p.split3 ? 2 : 1;  // matrices staged per operand
  const int stage_bytes = na * a_bytes + nb * b_bytes;
  // work items are handed out per cluster (1 or 2 CTAs); a pair splits two adjacent M (or N) tiles
  const int csize = p.pair ? 2 : 1;
  const int rank = p.pair ? (int)cluster_ctarank() : 0;
  const int m_items = p.pair == 1 ? (p.m_tiles + 1) / 2 : p.m_tiles;
  const int n_items = p.pair == 2 ? (p.n_tiles + 1) / 2 : p.n_tiles;
  const int num_work = p.n_split * m_items * n_items;
  const int wi0 = blockIdx.x / csize, wi_step = gridDim.x / csize;
#define BHMC_DECODE_WORK(wi)                                             \
  const int s = (wi) / (m_items * n_items), rem = (wi) % (m_items * n_items); \
  const int mt = (p.pair == 1) ? 2 * (rem / n_items) + rank : rem / n_items;  \
  const int nt = (p.pair == 2) ? 2 * (rem % n_items) + rank : rem % n_items;  \
  (void)s; (void)mt; (void)nt;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), csize);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bar_tfull[b]), 1);
      mbar_init(smem_u32(&bar_tempty[b]), 32 * EW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_launch_dependents();
  tcgen05_fence_before();
  if (p.pair) cluster_sync_all();  // peer barriers must be initialised before anything is multicast into them
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();  // everything above touched only shared / tensor memory

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t pol_a = make_l2_policy(p.hint_a), pol_b = make_l2_policy(p.hint_b);
      for (int w = wi0; w < num_work && p.debug != 9 && p.debug != 10; w += wi_step) {
        BHMC_DECODE_WORK(w)
        int k_begin = s * p.chunks_per_split, k_end = min(p.k_chunks, k_begin + p.chunks_per_split);
        for (int k = k_begin; k < k_end; ++k) {
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
          uint32_t full = smem_u32(&bar_full[stage]);
          if (p.debug >= 4) {  // measurement: no TMA traffic at all
            mbar_arrive(full);
            if (++stage == p.stages) stage = 0, phase ^= 1u;
            continue;
          }
          mbar_expect_tx(full, (uint32_t)stage_bytes);
          uint32_t sa = smem_base + stage * stage_bytes;  // [A_hi | A_lo | B_hi | B_lo]
          uint32_t sb = sa + na * a_bytes;
          int ak = p.a_k0 + k * BK, am = p.a_m0 + mt * BM, bk = k * BK, bn = nt * p.BN;
          if (p.a_slab) {  // chunks never straddle a slab: a_k0 and a_slab are multiples of BK
            const int sl = ak / p.a_slab;
            ak -= sl * p.a_slab;
            am += sl * p.a_slab_rows;
          }
          if (p.b_slab) {
            const int sl = bk / p.b_slab;
            bk -= sl * p.b_slab;
            bn += sl * p.b_slab_rows;
          }
          if (p.pair == 2) {  // A tile shared by the pair: each CTA fetches 64 of its 128 rows for both
            uint32_t off = (uint32_t)rank * (BM / 2) * (BK * 2);
            tma_load_2d_mc_hint(sa + off, &tmA_hi, full, ak, am + rank * (BM / 2), 3, pol_a);
            if (na == 2) tma_load_2d_mc_hint(sa + a_bytes + off, &tmA_lo, full, ak, am + rank * (BM / 2), 3, pol_a);
          } else {
            tma_load_2d_hint(sa, &tmA_hi, full, ak, am, pol_a);
            if (na == 2) tma_load_2d_hint(sa + a_bytes, &tmA_lo, full, ak, am, pol_a);
          }
          if (p.pair == 1) {  // B tile shared by the pair: each CTA fetches BN/2 of its rows for both
            uint32_t off = (uint32_t)rank * (p.BN / 2) * (BK * 2);
            tma_load_2d_mc_hint(sb + off, &tmB_hi, full, bk, bn + rank * (p.BN / 2), 3, pol_b);
            if (p.split3) tma_load_2d_mc_hint(sb + b_bytes + off, &tmB_lo, full, bk, bn + rank * (p.BN / 2), 3, pol_b);
          } else {
            tma_load_2d_hint(sb, &tmB_hi, full, bk, bn, pol_b);
            if (p.split3) tma_load_2d_hint(sb + b_bytes, &tmB_lo, full, bk, bn, pol_b);
          }
          if (++stage == p.stages) stage = 0, phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs this loop in warp-uniform control flow and one elected lane issues: keeping the
    // operands warp-uniform lets ptxas hold descriptors in uniform registers.  (Issuing from inside a divergent
    // `if (lane == 0)` made it wrap every UTCHMMA in an ELECT / 5x R2UR / BRA.ANY loop, ~90 cycles per MMA --
    // more than the 80 cycles the MMA itself takes; measured: 1070 cycles of issue per 12-MMA chunk.)
    // instruction descriptor: D=f32, A=B=bf16, both K-major, N, M=128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const bool do_mma = !(p.debug == 3 || p.debug == 5 || p.debug == 7);
    const bool do_wait = !(p.debug == 9 || p.debug == 10);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    long long t_tempty = 0, t_full = 0, t_issue = 0, t_start = clock64(), n_chunks = 0;
    for (int w = wi0; w < num_work; w += wi_step) {
      BHMC_DECODE_WORK(w)
      const int k_begin = s * p.chunks_per_split, k_end = min(p.k_chunks, k_begin + p.chunks_per_split);
      for (int kb = k_begin; kb < k_end; kb += p.sub_chunks, ++it) {
        const int ke = min(k_end, kb + p.sub_chunks);
        const int buf = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        long long c0 = p.prof ? clock64() : 0;
        mbar_wait(smem_u32(&bar_tempty[buf]), (use & 1u) ^ 1u);  // epilogue has drained this accumulator
        tcgen05_fence_after();
        if (p.prof) t_tempty += clock64() - c0;
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * TMEM_BUF_COLS);
        for (int k = kb; k < ke; ++k) {
          long long c1 = p.prof ? clock64() : 0;
          if (do_wait) {
            mbar_wait(smem_u32(&bar_full[stage]), phase);
            tcgen05_fence_after();
          }
          long long c2 = p.prof ? clock64() : 0;
          if (p.prof) t_full += c2 - c1, ++n_chunks;
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint32_t first = (k > kb) ? 1u : 0u;
          if (do_mma) {
            if (p.split3 == 1) {
              const uint64_t a_hi = make_smem_desc(sa), a_lo = make_smem_desc(sa + a_bytes);
              const uint64_t b_hi = make_smem_desc(sa + 2 * a_bytes), b_lo = make_smem_desc(sa + 2 * a_bytes + b_bytes);
#pragma unroll
              for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);  // +32 B per K step inside the swizzle span
                umma_bf16(tmem_d, a_hi + adv, b_hi + adv, idesc, ks > 0 ? 1u : first);
                umma_bf16(tmem_d, a_hi + adv, b_lo + adv, idesc, 1u);
                umma_bf16(tmem_d, a_lo + adv, b_hi + adv, idesc, 1u);
              }
            } else if (p.split3 == 2) {  // A exact in bf16: a_lo == 0, its product is not issued
              const uint64_t a_hi = make_smem_desc(sa);
              const uint64_t b_hi = make_smem_desc(sa + a_bytes), b_lo = make_smem_desc(sa + a_bytes + b_bytes);
#pragma unroll
              for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);
                umma_bf16(tmem_d, a_hi + adv, b_hi + adv, idesc, ks > 0 ? 1u : first);
                umma_bf16(tmem_d, a_hi + adv, b_lo + adv, idesc, 1u);
              }
            } else {
              const uint64_t a_hi = make_smem_desc(sa), b_hi = make_smem_desc(sa + a_bytes);
#pragma unroll
              for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);
                umma_bf16(tmem_d, a_hi + adv, b_hi + adv, idesc, ks > 0 ? 1u : first);
              }
            }
          }
          // smem slot reusable once these MMAs retire; in pair mode the peer's producer also writes into this
          // CTA's slot, so the release is multicast to both CTAs (empty barriers count 2 arrivals)
          if (p.debug == 9) {
          } else if (p.debug == 10 || !p.pair) {
            umma_commit(smem_u32(&bar_empty[stage]));
          } else {
            umma_commit_mc(smem_u32(&bar_empty[stage]), 3);
          }
          if (p.prof) t_issue += clock64() - c2;
          if (++stage == p.stages) stage = 0, phase ^= 1u;
        }
        umma_commit(smem_u32(&bar_tfull[buf]));  // accumulator (of this sub-slab) complete
      }
    }
    if (p.prof && lane == 0) {
      long long* o = p.prof + (size_t)blockIdx.x * 8;
      o[0] = clock64() - t_start, o[1] = t_tempty, o[2] = t_full, o[3] = t_issue, o[4] = n_chunks, o[5] = it;
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp & 3;             // TMEM lane quarter this warp may access
    const int part = (warp - 4) >> 2;    // which share of the chains / column chunks
    constexpr int PARTS = EW / 4;
    const int t = ew * 32 + lane;        // accumulator row handled by this thread
    int it = 0;
    for (int w = wi0; w < num_work; w += wi_step) {
      BHMC_DECODE_WORK(w)
      const bool tile_ok = mt < p.m_tiles && nt < p.n_tiles;  // odd tile counts leave a phantom tile in the last pair
      if constexpr (MODE == MODE_BWD) {
        // backward: drain the accumulator every sub_chunks chunks into fp32 registers (short tensor-core
        // accumulation chains), store the sum once per work item (coalesced per row) to the split-K partials
        constexpr int MAXCH = 3;  // 16-column chunks per thread: BN <= 192 with 16 epilogue warps
        float acc[MAXCH][16];
#pragma unroll
        for (int i = 0; i < MAXCH; ++i)
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
        const int k_begin = s * p.chunks_per_split, k_end = min(p.k_chunks, k_begin + p.chunks_per_split);
        for (int kb = k_begin; kb < k_end; kb += p.sub_chunks, ++it) {
          const int buf = it & 1;
          const uint32_t use = (uint32_t)(it >> 1);
          mbar_wait(smem_u32(&bar_tfull[buf]), use & 1u);
          tcgen05_fence_after();
          const uint32_t tacc = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * TMEM_BUF_COLS);
#pragma unroll
          for (int i = 0; i < MAXCH; ++i) {
            const int j0 = (part + i * PARTS) * 16;
            if (j0 < p.BN) {  // warp-uniform
              uint32_t raw[16];
              tmem_ld<16>(tacc + (uint32_t)j0, raw);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) acc[i][j] += __uint_as_float(raw[j]);
            }
          }
          tcgen05_fence_before();
          mbar_arrive(smem_u32(&bar_tempty[buf]));
        }
        if (tile_ok) {
          const int64_t rows = (int64_t)p.m_tiles * BM, cols = (int64_t)p.n_tiles * p.BN;
          float* dst = p.part + ((int64_t)s * rows + (int64_t)mt * BM + t) * cols + (int64_t)nt * p.BN;
#pragma unroll
          for (int i = 0; i < MAXCH; ++i) {
            const int j0 = (part + i * PARTS) * 16;
            if (j0 < p.BN) {
#pragma unroll
              for (int v = 0; v < 4; ++v)
                *reinterpret_cast<float4*>(dst + j0 + 4 * v) =
                    make_float4(acc[i][4 * v], acc[i][4 * v + 1], acc[i][4 * v + 2], acc[i][4 * v + 3]);
            }
          }
        }
        continue;
      }
      // forward: one accumulation chain per work item
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait(smem_u32(&bar_tfull[buf]), use & 1u);
      tcgen05_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * TMEM_BUF_COLS);
      if (!tile_ok || p.debug == 1 || p.debug == 6 || p.debug == 7 || p.debug >= 9) {
        // nothing to store; only the barrier protocol below
      } else if constexpr (MODE == MODE_FWD) {
        fwd_epilogue_tile<KP, EW, EXACT>(p, tacc, mt, nt, part, lane, t);
      }
      tcgen05_fence_before();
      mbar_arrive(smem_u32(&bar_tempty[buf]));
      ++it;
    }
  }
  // ---- teardown ----
  tcgen05_fence_before();
  if (p.pair) cluster_sync_all();  // the peer may still multicast into this CTA's smem / barriers
  else __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// Forward GEMM with cta_group::2: a cluster of two CTAs computes a 256 x BN tile pair per work item.  Each CTA
// loads its own 128 rows of X (hi, lo) and only HALF of the W tile (BN/2 rows); the leader CTA's elected thread
// issues tcgen05.mma.cta_group::2 (M = 256), which reads A and the two B halves from both CTAs' shared memory and
// writes rows 0-127 / 128-255 of the accumulator into the two CTAs' tensor memory.  Per CTA and K chunk this
// moves 52 KB instead of 72 KB through L2 (the main loop is bound by operand delivery) and leaves room for 4 stages.
//   full[s]   : leader's barrier, 1 arrival (leader's expect_tx covers both CTAs' bytes; the peer's TMA credits it)
//   empty[s]  : per CTA, released by the leader's tcgen05.commit multicast to both CTAs
//   tfull[b]  : per CTA, same multicast; tempty[b]: leader's, 2 x 32 x EW arrivals (peer's threads arrive remotely)
template <int KP, int EW, bool EXACT>
__global__ void __launch_bounds__(NON_EPI_THREADS + 32 * EW, 1)
k_tc_fwd2(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
          const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_base_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const bool leader = rank == 0;
  const int na = p.split3 == 1 ? 2 : 1, nb = p.split3 ? 2 : 1;
  const int a_bytes = BM * BK * 2, bh_bytes = (p.BN / 2) * BK * 2;  // this CTA's share of the B tile
  const int stage_bytes = na * a_bytes + nb * bh_bytes;
  const int m_items = (p.m_tiles + 1) / 2;
  const int num_work = m_items * p.n_tiles;
  const int wi0 = blockIdx.x / 2, wi_step = gridDim.x / 2;

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
  }
  if (warp == 2) {  // both CTAs' warp 2 take part in the pair-wide allocation
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  pdl_launch_dependents();
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();  // everything above touched only shared / tensor memory

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = wi0; w < num_work; w += wi_step) {
        const int mt = 2 * (w / p.n_tiles) + rank, nt = w % p.n_tiles;
        for (int k = 0; k < p.k_chunks; ++k) {
          if (p.prefetch) {
            // (measurement switch, off by default) Launches that carry few chains have one or two N tiles: every X tile
            // comes straight from HBM and the MMA thread waits 811 of 1 490 cycles per chunk for operands (in-kernel
            // counters at 16 chains).  Asking for the X tile `prefetch` chunks ahead in L2 did NOT help: see the host side.
            const int cl = k + p.prefetch, wn = w + (cl / p.k_chunks) * wi_step, kn = cl % p.k_chunks;
            if (wn < num_work) {
              int pk = p.a_k0 + kn * BK, pm = p.a_m0 + (2 * (wn / p.n_tiles) + rank) * BM;
              if (p.a_slab) {
                const int sl = pk / p.a_slab;
                pk -= sl * p.a_slab;
                pm += sl * p.a_slab_rows;
              }
              tma_prefetch_2d(&tmA_hi, pk, pm);
              if (na == 2) tma_prefetch_2d(&tmA_lo, pk, pm);
            }
          }
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
          const uint32_t full = smem_u32(&bar_full[stage]);  // same offset in the leader CTA
          if (leader) mbar_expect_tx(full, (uint32_t)(2 * stage_bytes));
          const uint32_t sa = smem_base + stage * stage_bytes, sb = sa + na * a_bytes;
          int ak = p.a_k0 + k * BK, am = p.a_m0 + mt * BM;
          const int bk = k * BK, bn = nt * p.BN + rank * (p.BN / 2);
          if (p.a_slab) {  // k-chunk-major X: chunk sl starts at row sl * a_slab_rows
            const int sl = ak / p.a_slab;
            ak -= sl * p.a_slab;
            am += sl * p.a_slab_rows;
          }
          tma_load_2d_2sm(sa, &tmA_hi, full, ak, am);
          if (na == 2) tma_load_2d_2sm(sa + a_bytes, &tmA_lo, full, ak, am);
          tma_load_2d_2sm(sb, &tmB_hi, full, bk, bn);
          if (p.split3) tma_load_2d_2sm(sb + bh_bytes, &tmB_lo, full, bk, bn);
          if (++stage == p.stages) stage = 0, phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only; warp-uniform loop, elected lane issues) =====
    if (leader) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      long long t_tempty = 0, t_full = 0, t_issue = 0, t_start = clock64(), n_chunks = 0;
      for (int w = wi0; w < num_work; w += wi_step, ++it) {
        const int buf = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        long long c0 = p.prof ? clock64() : 0;
        mbar_wait(smem_u32(&bar_tempty[buf]), (use & 1u) ^ 1u);  // both CTAs' epilogues have drained this accumulator
        tcgen05_fence_after();
        if (p.prof) t_tempty += clock64() - c0;
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * TMEM_BUF_COLS);
        for (int k = 0; k < p.k_chunks; ++k) {
          long long c1 = p.prof ? clock64() : 0;
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tcgen05_fence_after();
          long long c2 = p.prof ? clock64() : 0;
          if (p.prof) t_full += c2 - c1, ++n_chunks;
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint32_t first = k > 0 ? 1u : 0u;
          if (p.split3 == 1) {
            const uint64_t a_hi = make_smem_desc(sa), a_lo = make_smem_desc(sa + a_bytes);
            const uint64_t b_hi = make_smem_desc(sa + 2 * a_bytes), b_lo = make_smem_desc(sa + 2 * a_bytes + bh_bytes);
#pragma unroll
            for (int ks = 0; ks < BK / UMMA_K; ++ks) {
              const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);
              umma_bf16_2sm(tmem_d, a_hi + adv, b_hi + adv, idesc, ks > 0 ? 1u : first);
              umma_bf16_2sm(tmem_d, a_hi + adv, b_lo + adv, idesc, 1u);
              umma_bf16_2sm(tmem_d, a_lo + adv, b_hi + adv, idesc, 1u);
            }
          } else if (p.split3 == 2) {  // A exact in bf16
            const uint64_t a_hi = make_smem_desc(sa);
            const uint64_t b_hi = make_smem_desc(sa + a_bytes), b_lo = make_smem_desc(sa + a_bytes + bh_bytes);
#pragma unroll
            for (int ks = 0; ks < BK / UMMA_K; ++ks) {
              const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);
              umma_bf16_2sm(tmem_d, a_hi + adv, b_hi + adv, idesc, ks > 0 ? 1u : first);
              umma_bf16_2sm(tmem_d, a_hi + adv, b_lo + adv, idesc, 1u);
            }
          } else {
            const uint64_t a_hi = make_smem_desc(sa), b_hi = make_smem_desc(sa + a_bytes);
#pragma unroll
            for (int ks = 0; ks < BK / UMMA_K; ++ks) {
              const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);
              umma_bf16_2sm(tmem_d, a_hi + adv, b_hi + adv, idesc, ks > 0 ? 1u : first);
            }
          }
          umma_commit_2sm(smem_u32(&bar_empty[stage]), 3);  // frees the stage in BOTH CTAs
          if (p.prof) t_issue += clock64() - c2;
          if (++stage == p.stages) stage = 0, phase ^= 1u;
        }
        umma_commit_2sm(smem_u32(&bar_tfull[buf]), 3);  // accumulator halves complete in both CTAs
      }
      if (p.prof && lane == 0) {
        long long* o = p.prof + (size_t)blockIdx.x * 8;
        o[0] = clock64() - t_start, o[1] = t_tempty, o[2] = t_full, o[3] = t_issue, o[4] = n_chunks, o[5] = it;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs, each on its own 128 rows) =====
    const int ew = warp & 3, part = (warp - 4) >> 2;
    const int t = ew * 32 + lane;
    int it = 0;
    for (int w = wi0; w < num_work; w += wi_step, ++it) {
      const int mt = 2 * (w / p.n_tiles) + rank, nt = w % p.n_tiles;
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait(smem_u32(&bar_tfull[buf]), use & 1u);
      tcgen05_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * TMEM_BUF_COLS);
      if (mt < p.m_tiles && p.debug != 1) fwd_epilogue_tile<KP, EW, EXACT>(p, tacc, mt, nt, part, lane, t);  // debug 1: measurement only
      tcgen05_fence_before();
      if (leader) mbar_arrive(smem_u32(&bar_tempty[buf]));
      else mbar_arrive_remote(smem_u32(&bar_tempty[buf]), 0);
    }
  }
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// Backward GEMM with cta_group::2.  In-kernel counters showed the single-CTA backward starved for operands (270 of
// 1264 cycles per chunk waiting on the `full` barrier with 72 KB entering every SM per chunk): here a CTA pair owns
// TWO adjacent 128-row tiles of X^T (M = 256), each CTA stages its own X^T tile and only half of the (P-Y)^T tile
// (52 KB per chunk, 4 stages).  Work item = (row-slab s, tile pair, N tile); the accumulator is drained into fp32
// registers every sub_chunks chunks by both CTAs' epilogue warps; partial tiles go to part[s][row][col].
template <int EW>
__global__ void __launch_bounds__(NON_EPI_THREADS + 32 * EW, 1)
k_tc_bwd2(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
          const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_base_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const bool leader = rank == 0;
  const int na = p.split3 == 1 ? 2 : 1, nb = p.split3 ? 2 : 1;
  const int a_bytes = BM * BK * 2, bh_bytes = (p.BN / 2) * BK * 2;
  const int stage_bytes = na * a_bytes + nb * bh_bytes;
  const int m_pairs = p.m_tiles / 2;  // p.m_tiles is even; an odd last tile is the "left" work below
  const int per_split = m_pairs * p.n_tiles;
  const int num_main = p.n_split * per_split;
  const int num_work = num_main + p.left_n_split * p.n_tiles;
  const int wi0 = blockIdx.x / 2, wi_step = gridDim.x / 2;
#define BHMC_DECODE_BWD2(wi)                                                                              \
  const bool left = (wi) >= num_main;                                                                     \
  const int wl = left ? (wi) - num_main : (wi);                                                           \
  const int s = left ? wl / p.n_tiles : wl / per_split;                                                   \
  const int rem = left ? wl % p.n_tiles : wl % per_split;                                                 \
  const int mt = (left ? p.m_tiles : 2 * (rem / p.n_tiles)) + rank, nt = rem % p.n_tiles;                 \
  const int cps_w = left ? p.left_cps : p.chunks_per_split;                                               \
  const int k_begin = s * cps_w, k_end = min(p.k_chunks, k_begin + cps_w);

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
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  pdl_launch_dependents();
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();  // everything above touched only shared / tensor memory

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer (both CTAs) =====
      int stage = 0;
      uint32_t phase = 0;
      const uint64_t pol_a = make_l2_policy(p.hint_a), pol_b = make_l2_policy(p.hint_b);
      (void)pol_a;
      (void)pol_b;
      for (int w = wi0; w < num_work; w += wi_step) {
        BHMC_DECODE_BWD2(w)
        for (int k = k_begin; k < k_end; ++k) {
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
          const uint32_t full = smem_u32(&bar_full[stage]);
          if (leader) mbar_expect_tx(full, (uint32_t)(2 * stage_bytes));
          const uint32_t sa = smem_base + stage * stage_bytes, sb = sa + na * a_bytes;
          int ak = p.a_k0 + k * BK, am = p.a_m0 + mt * BM, bk = k * BK, bn = nt * p.BN + rank * (p.BN / 2);
          if (p.a_slab) {
            const int sl = ak / p.a_slab;
            ak -= sl * p.a_slab;
            am += sl * p.a_slab_rows;
          }
          if (p.b_slab) {
            const int sl = bk / p.b_slab;
            bk -= sl * p.b_slab;
            bn += sl * p.b_slab_rows;
          }
          tma_load_2d_2sm(sa, &tmA_hi, full, ak, am);
          if (na == 2) tma_load_2d_2sm(sa + a_bytes, &tmA_lo, full, ak, am);
          tma_load_2d_2sm(sb, &tmB_hi, full, bk, bn);
          if (p.split3) tma_load_2d_2sm(sb + bh_bytes, &tmB_lo, full, bk, bn);
          if (p.prefetch && k + p.prefetch < k_end) {  // pull a later chunk of this item's operands into L2
            int pak = p.a_k0 + (k + p.prefetch) * BK, pam = p.a_m0 + mt * BM, pbk = (k + p.prefetch) * BK;
            int pbn = nt * p.BN + rank * (p.BN / 2);
            if (p.a_slab) {
              const int sl = pak / p.a_slab;
              pak -= sl * p.a_slab;
              pam += sl * p.a_slab_rows;
            }
            if (p.b_slab) {
              const int sl = pbk / p.b_slab;
              pbk -= sl * p.b_slab;
              pbn += sl * p.b_slab_rows;
            }
            tma_prefetch_2d(&tmA_hi, pak, pam);
            if (na == 2) tma_prefetch_2d(&tmA_lo, pak, pam);
            tma_prefetch_2d(&tmB_hi, pbk, pbn);
            if (p.split3) tma_prefetch_2d(&tmB_lo, pbk, pbn);
          }
          if (++stage == p.stages) stage = 0, phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {  // ===== MMA issuer (leader CTA; warp-uniform loop, elected lane issues) =====
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      long long t_tempty = 0, t_full = 0, t_issue = 0, t_start = clock64(), n_chunks = 0;
      for (int w = wi0; w < num_work; w += wi_step) {
        BHMC_DECODE_BWD2(w)
        (void)mt;
        (void)nt;
        for (int kb = k_begin; kb < k_end; kb += p.sub_chunks, ++it) {
          const int ke = min(k_end, kb + p.sub_chunks);
          const int buf = it & 1;
          const uint32_t use = (uint32_t)(it >> 1);
          long long c0 = p.prof ? clock64() : 0;
          mbar_wait(smem_u32(&bar_tempty[buf]), (use & 1u) ^ 1u);  // both CTAs' epilogues have drained this accumulator
          tcgen05_fence_after();
          if (p.prof) t_tempty += clock64() - c0;
          const uint32_t tmem_d = tmem_base + (uint32_t)(buf * TMEM_BUF_COLS);
          for (int k = kb; k < ke; ++k) {
            long long c1 = p.prof ? clock64() : 0;
            mbar_wait(smem_u32(&bar_full[stage]), phase);
            tcgen05_fence_after();
            long long c2 = p.prof ? clock64() : 0;
            if (p.prof) t_full += c2 - c1, ++n_chunks;
            const uint32_t sa = smem_base + stage * stage_bytes;
            const uint32_t first = (k > kb) ? 1u : 0u;
            if (p.split3 == 1) {
              const uint64_t a_hi = make_smem_desc(sa), a_lo = make_smem_desc(sa + a_bytes);
              const uint64_t b_hi = make_smem_desc(sa + 2 * a_bytes), b_lo = make_smem_desc(sa + 2 * a_bytes + bh_bytes);
#pragma unroll
              for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);
                umma_bf16_2sm(tmem_d, a_hi + adv, b_hi + adv, idesc, ks > 0 ? 1u : first);
                umma_bf16_2sm(tmem_d, a_hi + adv, b_lo + adv, idesc, 1u);
                umma_bf16_2sm(tmem_d, a_lo + adv, b_hi + adv, idesc, 1u);
              }
            } else if (p.split3 == 2) {  // X^T exact in bf16: no lo copy staged, 2 MMAs per product
              const uint64_t a_hi = make_smem_desc(sa);
              const uint64_t b_hi = make_smem_desc(sa + a_bytes), b_lo = make_smem_desc(sa + a_bytes + bh_bytes);
#pragma unroll
              for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);
                umma_bf16_2sm(tmem_d, a_hi + adv, b_hi + adv, idesc, ks > 0 ? 1u : first);
                umma_bf16_2sm(tmem_d, a_hi + adv, b_lo + adv, idesc, 1u);
              }
            } else {
              const uint64_t a_hi = make_smem_desc(sa), b_hi = make_smem_desc(sa + a_bytes);
#pragma unroll
              for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);
                umma_bf16_2sm(tmem_d, a_hi + adv, b_hi + adv, idesc, ks > 0 ? 1u : first);
              }
            }
            umma_commit_2sm(smem_u32(&bar_empty[stage]), 3);
            if (p.prof) t_issue += clock64() - c2;
            if (++stage == p.stages) stage = 0, phase ^= 1u;
          }
          umma_commit_2sm(smem_u32(&bar_tfull[buf]), 3);  // this sub-slab's accumulator halves are complete in both CTAs
        }
      }
      if (p.prof && lane == 0) {
        long long* o = p.prof + (size_t)blockIdx.x * 8;
        o[0] = clock64() - t_start, o[1] = t_tempty, o[2] = t_full, o[3] = t_issue, o[4] = n_chunks, o[5] = it;
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs, each on its own 128 rows) =====
    const int ew = warp & 3, part = (warp - 4) >> 2;
    constexpr int PARTS = EW / 4;
    const int t = ew * 32 + lane;
    int it = 0;
    for (int w = wi0; w < num_work; w += wi_step) {
      BHMC_DECODE_BWD2(w)
      constexpr int MAXCH = 3;  // 16-column chunks per thread: BN <= 192 with 16 epilogue warps
      float acc[MAXCH][16];
#pragma unroll
      for (int i = 0; i < MAXCH; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
      for (int kb = k_begin; kb < k_end; kb += p.sub_chunks, ++it) {
        const int buf = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait(smem_u32(&bar_tfull[buf]), use & 1u);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * TMEM_BUF_COLS);
#pragma unroll
        for (int i = 0; i < MAXCH; ++i) {
          const int j0 = (part + i * PARTS) * 16;
          if (j0 < p.BN) {  // warp-uniform
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
      const int64_t cols = (int64_t)p.n_tiles * p.BN;
      float* dst = left ? p.part_left + ((int64_t)s * BM + t) * cols + (int64_t)nt * p.BN  // [s][128 rows of the odd tile]
                        : p.part + ((int64_t)s * p.m_tiles * BM + (int64_t)mt * BM + t) * cols + (int64_t)nt * p.BN;
      if (!(left && rank == 1)) {  // the phantom half of the odd pair has nothing to keep
#pragma unroll
        for (int i = 0; i < MAXCH; ++i) {
          const int j0 = (part + i * PARTS) * 16;
          if (j0 < p.BN) {
#pragma unroll
            for (int v = 0; v < 4; ++v)
              *reinterpret_cast<float4*>(dst + j0 + 4 * v) =
                  make_float4(acc[i][4 * v], acc[i][4 * v + 1], acc[i][4 * v + 2], acc[i][4 * v + 3]);
          }
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
#undef BHMC_DECODE_BWD2
}

#include "softmax_bwd_sk.cuh"

// ---------------------------------------------------------------------------------------------
// small helper kernels
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// Exact-operand check at bind time.  If scale*X is exactly representable in bf16 for every element (scale = 1: binary /
// small-integer features; scale = 255: 8-bit pixels divided by 255, the reference's MNIST input), the lo copy of X is
// identically zero and the lo(X).hi(W) MMA of the bf16x3 scheme adds nothing: the GEMMs then run on hi(scale*X) alone
// (2 MMAs per product, a third less tensor work and operand traffic) and 1/scale is folded into W / the gradient.
// flags[0] |= 1 if some x is not bf16-exact; flags[1] |= 1 if some x is not k/255 for an integer 0 <= k <= 255
__global__ void k_detect_exact(const float* __restrict__ X, int64_t n, int* __restrict__ flags) {
  int bad1 = 0, bad255 = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = X[i];
    bad1 |= __bfloat162float(__float2bfloat16_rn(x)) != x;
    const float k = rintf(x * 255.0f);
    // the quotient as fp32 computes it, as fp64 computes it before the cast to fp32 (pixels / 255.0 in NumPy), or as
    // a multiplication by the rounded reciprocal (what GPU array libraries turn a division by a scalar into)
    bad255 |= !(k >= 0.f && k <= 255.f &&
                (x == __fdiv_rn(k, 255.0f) || x == (float)((double)k / 255.0) || x == __fmul_rn(k, 1.0f / 255.0f)));
  }
  bad1 = __syncthreads_or(bad1);
  bad255 = __syncthreads_or(bad255);
  if (threadIdx.x == 0) {
    if (bad1) atomicOr(flags, 1);
    if (bad255) atomicOr(flags + 1, 1);
  }
}

// Xa[n, d] = split(scale * X[n, d]) with zero padding d in [D, Dp).  blk_rows == 0: plain [N, Dp] rows; else the
// k-chunk-major layout [Dp/64][blk_rows][64]: the 128 rows x 64 features a forward TMA box fetches are one contiguous
// 16 KB block (with plain rows every box row sits in a different DRAM page: 128-byte pieces at a 1664-byte stride)
__global__ void k_split_rows(const float* __restrict__ X, int64_t n0, int D, int64_t Dp, __nv_bfloat16* __restrict__ hi,
                             __nv_bfloat16* __restrict__ lo, float scale, int64_t blk_rows) {
  const int64_t n = n0 + blockIdx.y;
  for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < Dp; d += (int64_t)gridDim.x * blockDim.x) {
    float v = d < D ? scale * X[n * D + d] : 0.f;
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    const int64_t o = blk_rows ? ((d >> 6) * blk_rows + n) * 64 + (d & 63) : n * Dp + d;
    hi[o] = h;
    if (lo) lo[o] = l;
  }
}

// Xt slab s = n / S holds Xt[d, n % S] = split(scale * X[n, d]) for d < D and Xt[D, n % S] = scale; everything else (rows
// D+1..Dt_pad-1, columns of rows n >= N) stays zero from the memset at bind time
__global__ void k_split_transpose(const float* __restrict__ X, int64_t N, int D, int64_t S, int64_t ld, int64_t Dt_pad,
                                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, float scale) {
  __shared__ float tile[32][33];
  int64_t n0 = (int64_t)blockIdx.x * 32;
  int d0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int64_t n = n0 + i;
    int d = d0 + threadIdx.x;
    float v = 0.f;
    if (n < N) v = d < D ? scale * X[n * D + d] : (d == D ? scale : 0.f);
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int d = d0 + i;
    int64_t n = n0 + threadIdx.x;
    if (d <= D && n < N) {
      __nv_bfloat16 h, l;
      split_bf16(tile[threadIdx.x][i], h, l);
      const int64_t o = ((n / S) * Dt_pad + d) * ld + n % S;
      hi[o] = h;
      if (lo) lo[o] = l;
    }
  }
}

// Wt[(c*KP + k), d] = split(wscale * q[c, d*K + k]) (zero for k >= K or d >= D); loglik[c] = 0
// wscale = 1 / (scale of the bound X operand, SoftmaxData::x_scale)
__global__ void k_tc_prep(const float* __restrict__ q, int64_t ld, int D, int K, int KP, int64_t Dp,
                          __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, double* __restrict__ loglik,
                          float wscale) {
  pdl_launch_dependents();
  int c = blockIdx.y;
  if (blockIdx.x == 0 && threadIdx.x == 0 && loglik) loglik[c] = 0.0;
  int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= Dp) return;
  const float* src = q + (int64_t)c * ld + d * K;
  for (int k = 0; k < KP; ++k) {
    float v = (k < K && d < D) ? wscale * src[k] : 0.f;
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    int64_t o = ((int64_t)c * KP + k) * Dp + d;
    hi[o] = h;
    if (lo) lo[o] = l;
  }
}

// g[c, d*K + k] = alpha*q[c, d*K + k] + sum_s part[s, d, c*KP + k]      (d <= D: row D is the bias gradient)
// split-K partials of the backward GEMM: rows [0, split_row) come from the cta_group::2 launch (part0), the odd last
// row tile (if any) from the single-CTA launch (part1); each region has its own number of row slabs
struct PartRegions {
  const float* part0;
  const float* part1;
  int n_split0, n_split1;
  int64_t rows0, rows1;   // rows per slab of each region
  int64_t split_row;      // first output row of region 1
  int64_t cols;
  float scale;            // 1 / (scale of the bound X operand): multiplies the summed partials and the next W operand
  SkPlan sk;              // sk.on: part0 holds the pieces of k_tc_bwd_sk (softmax_bwd_sk.cuh) instead of row slabs
};
__device__ __forceinline__ float sum_partials(const PartRegions& r, int64_t d, int64_t col) {
  if (r.sk.on) return sk_sum_partials(r.sk, r.part0, (int)d, (int)col) * r.scale;
  const bool main = d < r.split_row;
  const float* src = (main ? r.part0 + d * r.cols : r.part1 + (d - r.split_row) * r.cols) + col;
  const int ns = main ? r.n_split0 : r.n_split1;
  const int64_t slab = (main ? r.rows0 : r.rows1) * r.cols;
  // loads in batches of 8 (a plain `v += src[..]` loop serialises one DRAM latency per slab), summed in slab order
  float v = 0.f;
  int s = 0;
  for (; s + 8 <= ns; s += 8) {
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = src[(int64_t)(s + j) * slab];
#pragma unroll
    for (int j = 0; j < 8; ++j) v += t[j];
  }
  for (; s < ns; ++s) v += src[(int64_t)s * slab];
  return v * r.scale;
}

__global__ void k_tc_reduce(PartRegions r, int K, int KP, int64_t P, const float* __restrict__ q, float* __restrict__ g,
                            int64_t ld, float alpha) {
  pdl_launch_dependents();
  int c = blockIdx.y;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ld) return;
  float v = 0.f;
  if (i < P) {
    int64_t d = i / K;
    int k = (int)(i - d * K);
    v = sum_partials(r, d, (int64_t)c * KP + k);
    v += alpha * q[(int64_t)c * ld + i];
  }
  g[(int64_t)c * ld + i] = v;
}

// k_tc_reduce + the SGLD / SGD parameter update + k_tc_prep of the NEXT evaluation in one launch.  One thread owns four
// consecutive parameters (the float4 / Philox block granularity of k_sgld, so the noise is element-for-element the
// one the separate kernel draws).  g is not materialised.
template <int KIND>
__global__ void __launch_bounds__(256)
k_tc_reduce_step(PartRegions r, int D, int K, int KP, int64_t P,
                 int64_t ld, float alpha, FusedStep fs, int64_t Dp, __nv_bfloat16* __restrict__ wt_hi,
                 __nv_bfloat16* __restrict__ wt_lo, double* __restrict__ loglik) {
  pdl_launch_dependents();
  const int c = blockIdx.y;
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i == 0 && loglik) loglik[c] = 0.0;  // for the next forward pass
  if (i >= P) return;
  float* qrow = fs.q + (int64_t)c * ld + i;
  float* prow = fs.p + (int64_t)c * ld + i;
  const float4 q4 = *reinterpret_cast<const float4*>(qrow);
  float qe[4] = {q4.x, q4.y, q4.z, q4.w}, ge[4], pe[4];
  int dd[4], kk[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int64_t idx = i + e;
    dd[e] = (int)(idx / K);
    kk[e] = (int)(idx - (int64_t)dd[e] * K);
    float v = 0.f;
    if (idx < P) {
      v = sum_partials(r, dd[e], (int64_t)c * KP + kk[e]);
      v += alpha * qe[e];
    }
    ge[e] = v;
  }
  if (KIND == BHMC_KIND_SGLD) {
    float4 z;
    if (fs.z) {
      const float* r = fs.z + (int64_t)c * fs.ld_z + i;
      z.x = r[0];
      z.y = (i + 1 < P) ? r[1] : 0.f;
      z.z = (i + 2 < P) ? r[2] : 0.f;
      z.w = (i + 3 < P) ? r[3] : 0.f;
    } else {
      z = philox_normal4(fs.seed, fs.chain_id0 + c, (uint32_t)(i >> 2), fs.stream_lo, fs.stream_hi);
    }
    const float ze[4] = {z.x, z.y, z.z, z.w};
    const float s2 = 2.0f * fs.eps, h = 0.5f * fs.eps;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      pe[e] = s2 * ze[e] - h * ge[e];
      qe[e] = qe[e] + pe[e];
    }
  } else {
    const float4 m4 = *reinterpret_cast<const float4*>(prow);
    const float me[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      pe[e] = fs.gamma * me[e] - fs.eps * ge[e];
      qe[e] = qe[e] + pe[e];
    }
  }
  *reinterpret_cast<float4*>(prow) = make_float4(pe[0], pe[1], pe[2], pe[3]);
  *reinterpret_cast<float4*>(qrow) = make_float4(qe[0], qe[1], qe[2], qe[3]);
  // bf16 operand copy of the updated weights (rows of Wt for k >= K or d >= D stay zero from the first k_tc_prep)
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (i + e < P && dd[e] < D) {
      __nv_bfloat16 hb, lb;
      split_bf16(r.scale * qe[e], hb, lb);
      const int64_t o = ((int64_t)c * KP + kk[e]) * Dp + dd[e];
      wt_hi[o] = hb;
      if (wt_lo) wt_lo[o] = lb;
    }
  }
}

// k_tc_reduce + the streaming schedule's next pre-event update (k_stream_update) + (optionally) k_tc_prep of the next
// evaluation, one launch.  One thread owns four consecutive parameters of a row; g IS written (the Metropolis / begin
// kernels of event phases and g_start need it).
__global__ void __launch_bounds__(256)
k_tc_reduce_stream(PartRegions r, int D, int K, int KP, int64_t P, int64_t ld, float alpha, float* __restrict__ g,
                   StreamUpdateArgs u, int prep_next, int64_t Dp, __nv_bfloat16* __restrict__ wt_hi,
                   __nv_bfloat16* __restrict__ wt_lo, double* __restrict__ loglik) {
  pdl_launch_dependents();
  const int c = blockIdx.y;
  const uint32_t op = u.code[c];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    stream_row_scalars(u, c, op);          // latches stat (= loglik of this evaluation) where the op says so ...
    if (prep_next) loglik[c] = 0.0;        // ... then clears it for the next forward pass (k_tc_prep's job)
  }
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= ld) return;
  const int64_t o = (int64_t)c * ld + i;
  const float4 q4 = *reinterpret_cast<const float4*>(u.q + o);
  float qe[4] = {q4.x, q4.y, q4.z, q4.w}, ge[4];
  int dd[4], kk[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int64_t idx = i + e;
    dd[e] = (int)(idx / K);
    kk[e] = (int)(idx - (int64_t)dd[e] * K);
    ge[e] = idx < P ? sum_partials(r, dd[e], (int64_t)c * KP + kk[e]) + alpha * qe[e] : 0.f;
  }
  *reinterpret_cast<float4*>(g + o) = make_float4(ge[0], ge[1], ge[2], ge[3]);
  if (op & OP_LATCH) *reinterpret_cast<float4*>(u.g_start + o) = make_float4(ge[0], ge[1], ge[2], ge[3]);
  bool moved = false;
  if (op & (OP_POST | OP_PRE)) {
    const float4 p4 = *reinterpret_cast<const float4*>(u.p + o);
    float pe[4] = {p4.x, p4.y, p4.z, p4.w};
    moved = stream_apply4(u, op, i, pe, ge, qe);
    *reinterpret_cast<float4*>(u.p + o) = make_float4(pe[0], pe[1], pe[2], pe[3]);
    if (moved) *reinterpret_cast<float4*>(u.q + o) = make_float4(qe[0], qe[1], qe[2], qe[3]);
  }
  if (prep_next) {  // bf16 operand copy of the (possibly updated) weights for the next forward pass
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (i + e < P && dd[e] < D) {
        __nv_bfloat16 hb, lb;
        split_bf16(r.scale * qe[e], hb, lb);
        const int64_t w = ((int64_t)c * KP + kk[e]) * Dp + dd[e];
        wt_hi[w] = hb;
        if (wt_lo) wt_lo[w] = lb;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// bf16 matrix [outer, inner] (inner contiguous, row stride in elements), box = [box_outer, 64], 128B swizzle
static int make_map(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_elems,
                    uint32_t box_outer) {
  // The same few maps are rebuilt for every evaluation (X, X^T and the scratch operands keep their addresses; a row
  // window is a kernel parameter, not part of the map).  Minibatch steps last ~45 us with three launches, so eight
  // driver calls per step are a visible share of the host time: keep the encoded maps (BHMC_MAP_CACHE=0 disables).
  struct Key {
    const void* base;
    uint64_t inner, outer, stride;
    uint32_t box;
    bool operator==(const Key& o) const {
      return base == o.base && inner == o.inner && outer == o.outer && stride == o.stride && box == o.box;
    }
  };
  struct Entry {
    Key k;
    CUtensorMap m;
  };
  static thread_local std::vector<Entry> cache;
  static int cache_env = -1;
  if (cache_env < 0) {
    const char* e = getenv("BHMC_MAP_CACHE");
    cache_env = e ? atoi(e) : 1;
  }
  const Key key{base, inner, outer, row_stride_elems, box_outer};
  if (cache_env) {
    for (const Entry& e : cache)
      if (e.k == key) {
        *m = e.m;
        return BHMC_OK;
      }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available (driver too old?)");
    return BHMC_ERR_CUDA;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): base=%p inner=%llu outer=%llu stride=%llu box=%u", (int)r, base,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_elems, box_outer);
    return BHMC_ERR_CUDA;
  }
  if (cache_env) {
    if (cache.size() >= 64) cache.clear();  // a map only describes addresses and shapes: a stale entry is never wrong
    cache.push_back(Entry{key, *m});
  }
  return BHMC_OK;
}

// feature stride of the K-major operands: 8 elements (16 B, the TMA minimum) or 64 (128 B: every box row is one
// aligned L2 line); BHMC_DP_ALIGN overrides for A/B measurements
static int dp_align() {
  static int v = 0;
  if (!v) {
    const char* e = getenv("BHMC_DP_ALIGN");
    v = e ? atoi(e) : 64;  // measured: 128 B aligned rows are ~3 % faster (one L2 line per box row)
    if (v != 8 && v != 64) v = 64;
  }
  return v;
}

static inline int k_chunks_of(int64_t cols) { return (int)ceil_div(cols, BK); }

static int pick_kp(int K) {
  const int opts[] = {4, 8, 10, 16, 24, 40, 64};
  for (int o : opts)
    if (K <= o) return o;
  return 0;
}

// chains per N tile: BN = cpt*KP must be a multiple of 16 and <= 192 (the backward epilogue keeps BN/64 chunks of
// 16 columns in registers).  Both GEMMs are bound by operand delivery (L2 -> shared memory), so the cost of a
// tiling is the bytes one work item streams -- every N tile re-reads the 128-row A tile and reads its own BN-row B
// tile -- times the number of rounds the persistent CTAs need.  Minibatch windows have fewer work items than SMs
// (one round whatever the tiling), where the narrowest tile wins: 2.26 -> 2.58 M grad-evals/s at cfg3.
static int pick_cpt(int KP, int C, int64_t m_tiles, int sm_count) {
  int unit = 1;
  while ((unit * KP) % 16) ++unit;
  static int forced = -1;  // BHMC_CPT: chains per N tile override (A/B measurements)
  if (forced < 0) {
    const char* e = getenv("BHMC_CPT");
    forced = e ? atoi(e) : 0;
  }
  if (forced > 0 && forced % unit == 0 && forced * KP <= 192) return forced;
  int best = unit;
  double best_cost = 1e30;
  for (int cpt = unit; cpt * KP <= 192; cpt += unit) {
    const int tiles = (C + cpt - 1) / cpt;
    const double rounds = std::max(1.0, (double)m_tiles * tiles / (double)sm_count);
    const double cost = rounds * (BM + cpt * KP + 16);  // +16: per-tile epilogue / scheduling overhead
    if (cost < best_cost - 1e-9) best_cost = cost, best = cpt;
  }
  return best;
}

// Layout knobs of the backward operands (A/B measurements): slab widths in contraction indices and the padding of
// the slab row stride.  Width BK with padding 0 is the fully blocked layout: every TMA box is one contiguous
// 128-row x 128-byte block.
static int64_t env_i64(const char* name, int64_t dflt) {
  const char* e = getenv(name);
  return e ? atoll(e) : dflt;
}
static int64_t xt_slab_width() {
  static int64_t v = 0;
  if (!v) v = std::max<int64_t>(BK, round_up(env_i64("BHMC_XT_SLAB", BK), BK));
  return v;
}
static int64_t slab_pad() {
  static int64_t v = -1;
  if (v < 0) v = round_up(std::max<int64_t>(0, env_i64("BHMC_SLAB_PAD", 0)), 8);
  return v;
}

int tc_softmax_bind(bhmc_ctx* ctx, SoftmaxData& d, bool want_lo, bool detect_exact) {
  const int kp = pick_kp(d.K);
  if (!kp) {
    set_error("tensor-core path supports at most 64 classes (got %d); use BHMC_PREC_FP32", d.K);
    return BHMC_ERR_UNSUPPORTED;
  }
  // re-binding the same shape (fresh host data every step) reuses the operand buffers
  const int64_t slab = std::min<int64_t>(xt_slab_width(), round_up(d.N, 64));
  const bool reuse = d.Xa_hi && d.Kp == kp && d.Dp == round_up(d.D, dp_align()) && d.slab == slab &&
                     d.Npad == round_up(d.N, slab) && d.Dt == d.D + 1 && (d.has_lo || !want_lo);
  if (!reuse) {
    tc_softmax_release(d);
    d.Kp = kp;
    d.Dp = round_up(d.D, dp_align());
    d.slab = slab;
    d.slab_ld = slab + slab_pad();
    d.Npad = round_up(d.N, slab);
    d.Dt = d.D + 1;
    d.Dt_pad = round_up(d.Dt, BM);
    size_t a_bytes = (size_t)d.N * d.Dp * 2, t_bytes = (size_t)d.Dt_pad * (d.Npad / d.slab) * d.slab_ld * 2;
    BHMC_CUDA_OK(cudaMalloc(&d.Xa_hi, a_bytes));
    BHMC_CUDA_OK(cudaMalloc(&d.Xt_hi, t_bytes));
    if (want_lo) {
      BHMC_CUDA_OK(cudaMalloc(&d.Xa_lo, a_bytes));
      BHMC_CUDA_OK(cudaMalloc(&d.Xt_lo, t_bytes));
    }
  }
  want_lo = want_lo || d.Xa_lo != nullptr;
  d.has_lo = want_lo;
  // BHMC_XA_BLOCKED=1: k-chunk-major forward operand.  Measured neutral at cfg2 for 8..64 chains per launch (forward
  // 66.7 / 80.2 / 110.4 / 148.0 us vs 72.0 / 76.7 / 110.1 / 148.4 us with plain rows; bench 171.5 k vs 171.7 k): the
  // forward is not limited by DRAM page locality, so the plain layout stays the default.
  static int xa_blk_env = -1;
  if (xa_blk_env < 0) {
    const char* e = getenv("BHMC_XA_BLOCKED");
    xa_blk_env = e ? atoi(e) : 0;
  }
  d.xa_blocked = xa_blk_env != 0 && BK == 64 && d.Dp % 64 == 0 && (d.Dp / 64) * d.N < ((int64_t)1 << 31);
  // exact-operand check (one pass over X and one host sync per bind; BHMC_X_EXACT=0 disables it for A/B measurements)
  d.x_scale = 1.f;
  d.x_exact = false;
  static int exact_env = -1;
  if (exact_env < 0) {
    const char* e = getenv("BHMC_X_EXACT");
    exact_env = e ? atoi(e) : 1;
  }
  if (detect_exact && exact_env) {
    void* fl = nullptr;
    BHMC_TRY(ctx->get_scratch(12, 2 * sizeof(int), &fl));
    BHMC_CUDA_OK(cudaMemsetAsync(fl, 0, 2 * sizeof(int), ctx->stream));
    const int64_t n = d.N * d.D;
    const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(n, 256 * 8), (int64_t)ctx->sm_count * 16);
    k_detect_exact<<<std::max(blocks, 1u), 256, 0, ctx->stream>>>(d.X, n, (int*)fl);
    ctx->launches++;
    int flags[2] = {1, 1};
    BHMC_CUDA_OK(cudaMemcpyAsync(flags, fl, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
    BHMC_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    if (!flags[0]) d.x_exact = true;
    else if (!flags[1]) d.x_exact = true, d.x_scale = 255.f;
  }
  if (d.x_exact) want_lo = false;  // the lo copies (if allocated) are neither written nor read
  {
    dim3 grid((unsigned)std::min<int64_t>(ceil_div(d.Dp, 256), 65535), (unsigned)1);
    // rows can exceed the 65535 limit of gridDim.y -> loop over row blocks
    for (int64_t n0 = 0; n0 < d.N; n0 += 65535) {
      int64_t cnt = std::min<int64_t>(65535, d.N - n0);
      dim3 g(grid.x, (unsigned)cnt);
      k_split_rows<<<g, 256, 0, ctx->stream>>>(d.X, n0, d.D, d.Dp, (__nv_bfloat16*)d.Xa_hi,
                                               want_lo ? (__nv_bfloat16*)d.Xa_lo : nullptr, d.x_scale,
                                               d.xa_blocked ? d.N : 0);
      ctx->launches++;
    }
  }
  {
    const size_t t_bytes = (size_t)d.Dt_pad * (d.Npad / d.slab) * d.slab_ld * 2;
    BHMC_CUDA_OK(cudaMemsetAsync(d.Xt_hi, 0, t_bytes, ctx->stream));
    if (want_lo) BHMC_CUDA_OK(cudaMemsetAsync(d.Xt_lo, 0, t_bytes, ctx->stream));
    dim3 grid((unsigned)ceil_div(d.N, 32), (unsigned)ceil_div(d.Dt, 32));
    k_split_transpose<<<grid, dim3(32, 8), 0, ctx->stream>>>(d.X, d.N, d.D, d.slab, d.slab_ld, d.Dt_pad, (__nv_bfloat16*)d.Xt_hi,
                                                             want_lo ? (__nv_bfloat16*)d.Xt_lo : nullptr, d.x_scale);
    ctx->launches++;
  }
  BHMC_CUDA_OK(cudaGetLastError());
  d.tc_ready = true;
  return BHMC_OK;
}

void tc_softmax_release(SoftmaxData& d) {
  cudaFree(d.Xa_hi);
  cudaFree(d.Xa_lo);
  cudaFree(d.Xt_hi);
  cudaFree(d.Xt_lo);
  d.Xa_hi = d.Xa_lo = d.Xt_hi = d.Xt_lo = nullptr;
  d.tc_ready = false;
}

// The GEMM kernels are launched with programmatic stream serialization (see tc_common.cuh): their prologue overlaps the
// tail of the kernel before them.  Measured: cfg3 SGLD 49.7 / 48.4 -> 43.3 / 43.1 us per minibatch step, cfg2 170.9 k ->
// 173.6 k grad-evals/s; parity suite green either way.  BHMC_PDL=0 launches them plainly (A/B measurements).
static bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BHMC_PDL");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}

// BHMC_PAIR=0 disables the CTA-pair multicast (debug / A-B comparison)
static bool pairing_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BHMC_PAIR");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}

static int epilogue_warps() {
  static int ew = 0;
  if (!ew) {
    const char* e = getenv("BHMC_EPI_WARPS");
    ew = e ? atoi(e) : 16;
    if (ew != 8 && ew != 16 && ew != 24) ew = 16;
  }
  return ew;
}

template <int MODE, int KP, int EW, bool EXACT>
static int launch_gemm_ew(bhmc_ctx* ctx, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                          const CUtensorMap& b_lo, const TcParams& p) {
  int stage_bytes = (p.split3 == 1 ? 2 : 1) * BM * BK * 2 + (p.split3 ? 2 : 1) * p.BN * BK * 2;
  size_t smem = (size_t)p.stages * stage_bytes + 1024;
  static size_t configured = 0;
  if (smem > configured) {
    BHMC_CUDA_OK(cudaFuncSetAttribute(k_tc_gemm<MODE, KP, EW, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int m_items = p.pair == 1 ? (p.m_tiles + 1) / 2 : p.m_tiles;
  const int n_items = p.pair == 2 ? (p.n_tiles + 1) / 2 : p.n_tiles;
  const int work = p.n_split * m_items * n_items;
  const int csize = p.pair ? 2 : 1;
  const int grid = csize * std::min(work, ctx->sm_count / csize);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(NON_EPI_THREADS + 32 * EW);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  BHMC_CUDA_OK(cudaLaunchKernelEx(&cfg, k_tc_gemm<MODE, KP, EW, EXACT>, a_hi, a_lo, b_hi, b_lo, p));
  ctx->launches++;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

template <int KP, int EW, bool EXACT>
static int launch_fwd2_ew(bhmc_ctx* ctx, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                          const CUtensorMap& b_lo, const TcParams& p) {
  const int stage_bytes = (p.split3 == 1 ? 2 : 1) * BM * BK * 2 + (p.split3 ? 2 : 1) * (p.BN / 2) * BK * 2;
  const size_t smem = (size_t)p.stages * stage_bytes + 1024;
  static size_t configured = 0;
  if (smem > configured) {
    BHMC_CUDA_OK(cudaFuncSetAttribute(k_tc_fwd2<KP, EW, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int work = ((p.m_tiles + 1) / 2) * p.n_tiles;
  const int grid = 2 * std::min(work, ctx->sm_count / 2);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(NON_EPI_THREADS + 32 * EW);
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
  BHMC_CUDA_OK(cudaLaunchKernelEx(&cfg, k_tc_fwd2<KP, EW, EXACT>, a_hi, a_lo, b_hi, b_lo, p));
  ctx->launches++;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

static int launch_bwd2(bhmc_ctx* ctx, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                       const CUtensorMap& b_lo, const TcParams& p) {
  const int stage_bytes = (p.split3 == 1 ? 2 : 1) * BM * BK * 2 + (p.split3 ? 2 : 1) * (p.BN / 2) * BK * 2;
  const size_t smem = (size_t)p.stages * stage_bytes + 1024;
  static size_t configured = 0;
  if (smem > configured) {
    BHMC_CUDA_OK(cudaFuncSetAttribute(k_tc_bwd2<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int work = p.n_split * (p.m_tiles / 2) * p.n_tiles + p.left_n_split * p.n_tiles;
  const int grid = 2 * std::min(work, ctx->sm_count / 2);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
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
  BHMC_CUDA_OK(cudaLaunchKernelEx(&cfg, k_tc_bwd2<16>, a_hi, a_lo, b_hi, b_lo, p));
  ctx->launches++;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

template <int MODE, int KP>
static int launch_gemm(bhmc_ctx* ctx, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                       const CUtensorMap& b_lo, const TcParams& p) {
  const bool exact = p.K == KP;
  if constexpr (MODE == MODE_FWD) {
    if (p.pair == 3) {  // cta_group::2 forward
      if (exact) return launch_fwd2_ew<KP, 16, true>(ctx, a_hi, a_lo, b_hi, b_lo, p);
      return launch_fwd2_ew<KP, 16, false>(ctx, a_hi, a_lo, b_hi, b_lo, p);
    }
  }
  if (epilogue_warps() == 8) {
    if (exact) return launch_gemm_ew<MODE, KP, 8, true>(ctx, a_hi, a_lo, b_hi, b_lo, p);
    return launch_gemm_ew<MODE, KP, 8, false>(ctx, a_hi, a_lo, b_hi, b_lo, p);
  }
  if constexpr (KP <= 16 && MODE == MODE_FWD) {
    if (exact && epilogue_warps() == 24) return launch_gemm_ew<MODE, KP, 24, true>(ctx, a_hi, a_lo, b_hi, b_lo, p);
  }
  if (exact) return launch_gemm_ew<MODE, KP, 16, true>(ctx, a_hi, a_lo, b_hi, b_lo, p);
  return launch_gemm_ew<MODE, KP, 16, false>(ctx, a_hi, a_lo, b_hi, b_lo, p);
}

template <int KP>
static int launch_fwd_kp(bhmc_ctx* ctx, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi,
                         const CUtensorMap& b_lo, const TcParams& p) {
  return launch_gemm<MODE_FWD, KP>(ctx, a_hi, a_lo, b_hi, b_lo, p);
}

int tc_softmax_grad(bhmc_ctx* ctx, const SoftmaxData& d, const float* q, int C, int64_t ld, float alpha,
                    int64_t row0, int64_t nrows, float* g, double* loglik, bool split3, const FusedStep* fs, ZCache* zc,
                    int zmode, const FusedStream* fst, bool prepared) {
  BHMC_CHECK_ARG(!fst || (g && !fs && q == fst->u.q), "fused stream update: needs g and the working state the update moves");
  BHMC_CHECK_ARG(!fs || (q == fs->q && loglik), "fused step: the gradient must be evaluated at the state it updates");
  BHMC_CHECK_ARG(d.tc_ready, "tensor-core operands were not prepared at bind time (precision_mask)");
  BHMC_CHECK_ARG(!split3 || d.has_lo, "bf16x3 needs the lo operand copies (precision_mask bit 1 at bind time)");
  BHMC_CHECK_ARG(row0 >= 0 && nrows > 0 && row0 + nrows <= d.N, "row window [%lld,+%lld) outside the %lld bound rows",
                 (long long)row0, (long long)nrows, (long long)d.N);
  const int KP = d.Kp, K = d.K, D = d.D;
  const int cpt = pick_cpt(KP, C, ceil_div(nrows, BM), ctx->sm_count);
  const int BN = cpt * KP;
  const int n_tiles = (int)ceil_div(C, cpt);
  const int64_t ncols = (int64_t)C * KP;  // rows of Wt / DmT
  const int64_t Mfwd = round_up(nrows, BM);          // rows covered by the forward tiles
  // The backward contraction index is the absolute row: its BK-chunks must start on multiples of BK so that no
  // chunk straddles two Xt slabs (and TMA inner coordinates must be 16 B aligned anyway), hence (P-Y)^T is
  // written `shift` columns to the right
  const int shift = (int)(row0 % BK);
  const int dm_tail = (BK - shift) % BK;             // columns the last backward chunk reads past the last tile row
  const int64_t dm_cols = Mfwd + BK;                 // >= shift + Mfwd + dm_tail
  static int64_t dm_slab_max = 0;  // BHMC_DM_SLAB overrides the slab width (A/B measurements; multiple of 64)
  if (!dm_slab_max) {
    const char* e = getenv("BHMC_DM_SLAB");
    dm_slab_max = e ? std::max<int64_t>(64, round_up(atoll(e), 64)) : BK;
  }
  const int64_t dm_slab = std::min<int64_t>(dm_slab_max, dm_cols);  // multiple of 64 either way
  const int64_t dm_nslab = ceil_div(dm_cols, dm_slab), dm_ld = dm_slab + slab_pad();
  const int64_t dm_rows = (int64_t)n_tiles * BN;     // rows per slab (>= C*KP; the padding rows are never written
                                                     // and only feed accumulator columns nobody reads)
  const int64_t P = (int64_t)(D + 1) * K;
  // operand split: 0 = single pass, 1 = hi/lo of both operands (3 MMAs), 2 = X exact in bf16 (bind-time check): only
  // the per-evaluation operand (W / P-Y) carries a lo copy, 2 MMAs
  const int smode = split3 ? (d.x_exact ? 2 : 1) : 0;
  const int nmat = split3 ? 2 : 1, na = smode == 1 ? 2 : 1;
  const int stage_bytes = na * BM * BK * 2 + nmat * BN * BK * 2;
  int stages = std::max(2, std::min(MAX_STAGES, (int)((225 * 1024) / stage_bytes)));
  if (const char* e = getenv("BHMC_STAGES")) stages = std::max(1, std::min(stages, atoi(e)));

  // scratch: slot 1 = Wt hi|lo, slot 2 = DmT hi|lo, slot 3 = split-K partials
  void* wt = nullptr;
  size_t wt_bytes = (size_t)ncols * d.Dp * 2;
  BHMC_TRY(ctx->get_scratch(1, wt_bytes * 2, &wt));
  __nv_bfloat16* wt_hi = (__nv_bfloat16*)wt;
  __nv_bfloat16* wt_lo = (__nv_bfloat16*)((char*)wt + wt_bytes);

  // ---- Z cache -------------------------------------------------------------------------------------------------
  if (!zc || fs) zmode = ZMODE_NONE;
  const bool use_z = zmode == ZMODE_USE && zc->valid && ctx->zcache_owner == zc && zc->row0 == row0 && zc->nrows == nrows && zc->KP == KP &&
                     C <= zc->C && zc->ld == dm_ld && zc->slab == dm_slab;
  if (zmode == ZMODE_USE && !use_z) zmode = ZMODE_NONE;
  if (!use_z && !(fs && fs->wt_ready) && !prepared) {
    GroupTimer t(ctx, KG_PREP);
    dim3 grid((unsigned)ceil_div(d.Dp, 128), C);
    k_tc_prep<<<grid, 128, 0, ctx->stream>>>(q, ld, D, K, KP, d.Dp, wt_hi, split3 ? wt_lo : nullptr, loglik, 1.0f / d.x_scale);
    ctx->launches++;
  }

  void* dmt = nullptr;
  size_t dmt_bytes = (size_t)(dm_nslab * dm_rows * dm_ld) * 2;
  __nv_bfloat16 *dmt_hi = nullptr, *dmt_lo = nullptr;
  if (g || fs) {
    BHMC_TRY(ctx->get_scratch(2, dmt_bytes * 2, &dmt));
    dmt_hi = (__nv_bfloat16*)dmt;
    dmt_lo = split3 ? (__nv_bfloat16*)((char*)dmt + dmt_bytes) : nullptr;
  }

  // ---- forward: Z[rows, C*KP] = Xa[rows, Dp] . Wt^T ----
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  // forward: pairs of M tiles share the W tile (each CTA fetches half of it and multicasts)
  // forward CTA pairing: 3 = cta_group::2 MMA (each CTA stores half of the W tile), 1 = two independent MMAs sharing
  // the W tile by TMA multicast, 0 = single CTAs.  BHMC_FWD2=0 falls back from 3 to 1.
  static int fwd2 = -1;
  if (fwd2 < 0) {
    const char* e = getenv("BHMC_FWD2");
    fwd2 = e ? atoi(e) : 1;
  }
  const bool big = pairing_enabled() && Mfwd / BM >= 2 && (Mfwd / BM) * n_tiles >= ctx->sm_count;
  const int fwd_pair = big ? (fwd2 ? 3 : 1) : 0;
  const uint32_t fwd_bbox = (uint32_t)(fwd_pair ? BN / 2 : BN);
  // forward A: plain rows [N, Dp] or k-chunk-major blocks [Dp/64][N][64] (coordinates: k inside the chunk, chunk*N + row)
  const uint64_t xa_inner = d.xa_blocked ? 64 : (uint64_t)d.Dp, xa_outer = d.xa_blocked ? (uint64_t)(d.Dp / 64) * d.N : (uint64_t)d.N;
  BHMC_TRY(make_map(&a_hi, d.Xa_hi, xa_inner, xa_outer, xa_inner, BM));
  BHMC_TRY(make_map(&b_hi, wt_hi, (uint64_t)d.Dp, (uint64_t)ncols, (uint64_t)d.Dp, fwd_bbox));
  a_lo = a_hi;
  b_lo = b_hi;
  if (smode == 1) BHMC_TRY(make_map(&a_lo, d.Xa_lo, xa_inner, xa_outer, xa_inner, BM));
  if (split3) BHMC_TRY(make_map(&b_lo, wt_lo, (uint64_t)d.Dp, (uint64_t)ncols, (uint64_t)d.Dp, fwd_bbox));
  TcParams p{};
  p.m_tiles = (int)(Mfwd / BM);
  p.n_tiles = n_tiles;
  p.n_split = 1;
  p.k_chunks = (int)ceil_div(D, BK);
  p.chunks_per_split = p.k_chunks;
  p.sub_chunks = p.k_chunks;  // forward: K = D is short (13 chunks at D=784), one accumulation chain
  p.BN = BN;
  p.stages = fwd_pair == 3 ? std::max(2, std::min(MAX_STAGES, (int)((225 * 1024) / (na * BM * BK * 2 + nmat * (BN / 2) * BK * 2))))
                           : stages;
  p.split3 = smode;
  p.a_k0 = 0;
  p.a_m0 = (int)row0;
  if (d.xa_blocked) p.a_slab = 64, p.a_slab_rows = (int)d.N;
  p.pair = fwd_pair;
  {
    // L2 prefetch distance of the cta_group::2 forward kernel, in chunks (BHMC_FWD_PF, default 0 = off).  Measured on one
    // box at 8 / 16 / 32 / 64 chains per launch: 62.8 / 75.2 / 140.8 / 145.7 us without, 71.5 / 84.3 / 127.6 / 165.8 us at a
    // distance of 8 (4 and 16 alike): it helps only the two-N-tile case and costs 10-14 % everywhere else.
    static int pf_fwd = -1;
    if (pf_fwd < 0) {
      const char* e = getenv("BHMC_FWD_PF");
      pf_fwd = e ? std::max(0, atoi(e)) : 0;
    }
    p.prefetch = pf_fwd;
  }
  p.C = C;
  p.K = K;
  p.cpt = cpt;
  p.D = D;
  p.ld = ld;
  p.q = q;
  p.labels = d.labels + row0;
  p.nrows = nrows;
  p.dm_slab = (int)dm_slab;
  p.dm_ld = (int)dm_ld;
  p.dm_slab_rows = (int)dm_rows;
  p.dm_tail = dm_tail;
  p.dm_shift = shift;
  {
    // measured neutral at cfg2 (forward launches 12.4 ms of a 8-step run either way, 188.2 / 187.9 k vs 188.2 / 190.4 k
    // grad-evals/s; all 115 GPU tests pass with it): off by default
    static int pack_env = -1;  // BHMC_FWD_PACK=1: 4-byte row-pair stores
    if (pack_env < 0) {
      const char* e = getenv("BHMC_FWD_PACK");
      pack_env = e ? atoi(e) : 0;
    }
    p.pack_dm = (pack_env && shift % 2 == 0 && dm_ld % 2 == 0 && dm_slab % 2 == 0) ? 1 : 0;
  }
  p.dmt_hi = dmt_hi;
  p.dmt_lo = dmt_lo;
  p.loglik = loglik;
  p.write_dm = (g || fs) ? 1 : 0;
  p.prof = nullptr;
  static int want_prof = -1;
  if (want_prof < 0) want_prof = getenv("BHMC_PROF") ? 1 : 0;
  void* prof_dev = nullptr;
  if (want_prof) {
    BHMC_TRY(ctx->get_scratch(5, sizeof(long long) * 8 * 1024, &prof_dev));
    BHMC_CUDA_OK(cudaMemsetAsync(prof_dev, 0, sizeof(long long) * 8 * 1024, ctx->stream));
    p.prof = (long long*)prof_dev;
  }
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("BHMC_DEBUG_EPI");
      dbg = e ? atoi(e) : 0;
    }
    p.debug = dbg;
  }
  p.part = nullptr;
  // Z cache: slot 8 holds X.W (fp32, transposed) of the last full forward pass that was asked to keep it
  if (use_z || zmode == ZMODE_STORE) {
    const int zrows = use_z ? zc->slab_rows : (int)dm_rows;
    void* zbuf = nullptr;
    BHMC_TRY(ctx->get_scratch(8, sizeof(float) * (size_t)(dm_nslab * zrows * dm_ld), &zbuf));
    p.zt = (float*)zbuf;
    p.zt_slab_rows = zrows;
  }
  if (use_z) {
    GroupTimer t(ctx, KG_FWD);
    BHMC_CUDA_OK(cudaMemsetAsync(loglik, 0, sizeof(double) * C, ctx->stream));
    dim3 grid((unsigned)C, (unsigned)p.m_tiles);
    const bool exact = K == KP;
    TcParams pz = p;
    pz.cpt = 1;  // one chain per block: chain index = blockIdx.y
    static int fz_vec = -1;  // BHMC_FROM_Z_VEC=0: the scalar kernel (A/B measurements)
    if (fz_vec < 0) {
      const char* e = getenv("BHMC_FROM_Z_VEC");
      fz_vec = e ? atoi(e) : 1;
    }
    const int n_cols = k_chunks_of(nrows + shift) * BK;  // every column the backward chunks read
    bool done_vec = false;
    if (fz_vec && KP <= 40) {
#define BHMC_FROM_ZV(KPV, VEC)                                                                                \
  case KPV: {                                                                                                 \
    dim3 gv((unsigned)C, (unsigned)ceil_div(n_cols, 128 * VEC));                                              \
    if (exact) k_softmax_from_z_vec<KPV, true, VEC><<<gv, 128, 0, ctx->stream>>>(pz, n_cols);                 \
    else k_softmax_from_z_vec<KPV, false, VEC><<<gv, 128, 0, ctx->stream>>>(pz, n_cols);                      \
    done_vec = true;                                                                                          \
  } break;
      switch (KP) {
        BHMC_FROM_ZV(4, 4) BHMC_FROM_ZV(8, 4) BHMC_FROM_ZV(10, 4) BHMC_FROM_ZV(16, 4) BHMC_FROM_ZV(24, 2) BHMC_FROM_ZV(40, 2)
        default: break;
      }
#undef BHMC_FROM_ZV
    }
    if (!done_vec) {
#define BHMC_FROM_Z(KPV)                                                    \
  case KPV:                                                                 \
    if (exact) k_softmax_from_z<KPV, true><<<grid, 128, 0, ctx->stream>>>(pz); \
    else k_softmax_from_z<KPV, false><<<grid, 128, 0, ctx->stream>>>(pz);      \
    break;
    switch (KP) {
      BHMC_FROM_Z(4) BHMC_FROM_Z(8) BHMC_FROM_Z(10) BHMC_FROM_Z(16) BHMC_FROM_Z(24) BHMC_FROM_Z(40) BHMC_FROM_Z(64)
      default: set_error("unsupported KP %d", KP); return BHMC_ERR_UNSUPPORTED;
    }
#undef BHMC_FROM_Z
    }
    ctx->launches++;
    BHMC_CUDA_OK(cudaGetLastError());
  } else {
    GroupTimer t(ctx, KG_FWD);
    int rc;
    switch (KP) {
      case 4: rc = launch_fwd_kp<4>(ctx, a_hi, a_lo, b_hi, b_lo, p); break;
      case 8: rc = launch_fwd_kp<8>(ctx, a_hi, a_lo, b_hi, b_lo, p); break;
      case 10: rc = launch_fwd_kp<10>(ctx, a_hi, a_lo, b_hi, b_lo, p); break;
      case 16: rc = launch_fwd_kp<16>(ctx, a_hi, a_lo, b_hi, b_lo, p); break;
      case 24: rc = launch_fwd_kp<24>(ctx, a_hi, a_lo, b_hi, b_lo, p); break;
      case 40: rc = launch_fwd_kp<40>(ctx, a_hi, a_lo, b_hi, b_lo, p); break;
      case 64: rc = launch_fwd_kp<64>(ctx, a_hi, a_lo, b_hi, b_lo, p); break;
      default: set_error("unsupported KP %d", KP); rc = BHMC_ERR_UNSUPPORTED;
    }
    BHMC_TRY(rc);
    if (zmode == ZMODE_STORE) {
      ctx->zcache_owner = zc;
      zc->valid = true;
      zc->row0 = row0;
      zc->nrows = nrows;
      zc->C = C;
      zc->KP = KP;
      zc->slab_rows = (int)dm_rows;
      zc->slab = dm_slab;
      zc->ld = dm_ld;
    } else if (zc && !use_z) {
      zc->valid = false;  // a full pass that did not keep Z: whatever slot 8 holds belongs to older weights
    }
  }
  if (want_prof) {
    std::vector<long long> hp(8 * 148);
    BHMC_CUDA_OK(cudaMemcpyAsync(hp.data(), prof_dev, sizeof(long long) * 8 * 148, cudaMemcpyDeviceToHost, ctx->stream));
    BHMC_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    double tot = 0, te = 0, tf = 0, ti = 0, nc = 0, nt = 0;
    int n = 0;
    for (int b = 0; b < 148; ++b)
      if (hp[b * 8] > 0) tot += hp[b * 8], te += hp[b * 8 + 1], tf += hp[b * 8 + 2], ti += hp[b * 8 + 3], nc += hp[b * 8 + 4], nt += hp[b * 8 + 5], ++n;
    if (n)
      fprintf(stderr, "[bhmc prof fwd] MMA thread, mean over %d CTAs: total %.0f cyc; wait tempty %.0f; wait full %.0f; issue+commit %.0f; chunks %.0f tiles %.0f -> per chunk: full-wait %.0f issue %.0f\n", n, tot / n, te / n, tf / n, ti / n, nc / n, nt / n, tf / nc, ti / nc);
  }
  if (!g && !fs) return BHMC_OK;

  // ---- backward: G[D+1, C*KP] = Xt[D+1, rows] . DmT^T, split over row slabs ----
  const int m_tiles_all = (int)ceil_div(d.Dt, BM);
  const int k_chunks_b = (int)ceil_div(nrows + shift, BK);
  // k_tc_bwd2 (cta_group::2 MMAs: two adjacent row tiles per CTA pair, 52 KB per SM per chunk instead of 72 KB) is
  // OPT-IN (BHMC_BWD2=1, =2 also for launches with < 3 N tiles).  Timed alone it is 9 % faster than the single-CTA
  // kernel (ncu: 152.5 + 9.9 us vs 167.6 + 11.1 us at cfg2, MMA thread at ~1005 cycles per chunk = the tensor
  // roofline), but inside a long run the GPU sits at its 1 kW power cap: the busier tensor pipe plus the phantom half
  // of the odd tile pair pull the SM clock from 1.53 to 1.37 GHz and the whole step gets SLOWER (169 k -> 162 k
  // grad-evals/s, three A/B repetitions).  The run is energy-bound, not pipe-bound: removing flops helps, raising
  // utilisation does not.
  static int bwd2_env = -1;
  if (bwd2_env < 0) {
    const char* e = getenv("BHMC_BWD2");
    bwd2_env = e ? atoi(e) : 0;
  }
  // (measured at cfg2: faster from ~3 N tiles on; launches that carry few chains are better off with more row slabs)
  const bool use_bwd2 = bwd2_env && pairing_enabled() && m_tiles_all >= 2 && k_chunks_b >= 64 && BN % 16 == 0 &&
                        (n_tiles >= 3 || bwd2_env >= 2);
  const int m2 = use_bwd2 ? 2 * (m_tiles_all / 2) : 0;   // row tiles handled pairwise by k_tc_bwd2
  const int m_left = use_bwd2 ? m_tiles_all - m2 : 0;    // odd last tile: extra work items of the same launch
  const int m1 = use_bwd2 ? 0 : m_tiles_all;             // row tiles handled by the single-CTA kernel
  // <= 192 MMAs per tensor-core accumulation chain, then round-to-nearest fp32 adds (BHMC_SUB_CHUNKS overrides)
  static int sub_env = -1;
  if (sub_env < 0) {
    const char* e = getenv("BHMC_SUB_CHUNKS");
    sub_env = e ? std::max(1, atoi(e)) : 16;
  }
  static int l2hint = -1;
  if (l2hint < 0) {
    const char* e = getenv("BHMC_L2HINT");
    l2hint = e ? atoi(e) : 1;
  }
  TcParams b{};
  b.n_tiles = n_tiles;
  b.k_chunks = k_chunks_b;
  b.sub_chunks = sub_env * (64 / BK);  // expressed in 64-element units
  b.BN = BN;
  b.split3 = smode;
  b.a_k0 = (int)(row0 - shift);
  b.b_slab = (int)dm_slab;
  b.b_slab_rows = (int)dm_rows;
  b.a_slab = (int)d.slab;
  b.a_slab_rows = (int)d.Dt_pad;
  const int64_t pcol = (int64_t)n_tiles * BN;
  const uint64_t xt_rows = (uint64_t)(d.Npad / d.slab) * d.Dt_pad;

  // plan of the pair kernel: split the contraction so that one round of (pair, N tile, slab) items fills the clusters
  TcParams b2 = b;
  if (m2) {
    const int units = ctx->sm_count / 2, items = (m2 / 2) * n_tiles;
    int want = std::max(1, units / items);
    want = std::min(want, std::max(1, k_chunks_b / 4));
    b2.chunks_per_split = (int)ceil_div(k_chunks_b, want);
    b2.n_split = (int)ceil_div(k_chunks_b, b2.chunks_per_split);
    b2.m_tiles = m2;
    b2.a_m0 = 0;
    b2.pair = 3;
    const int st_bytes = na * BM * BK * 2 + nmat * (BN / 2) * BK * 2;
    b2.stages = std::max(2, std::min(MAX_STAGES, (int)((225 * 1024) / st_bytes)));
    static int pf_env = -1;  // BHMC_PF: L2 prefetch distance of the pair kernel in chunks
    if (pf_env < 0) {
      const char* e = getenv("BHMC_PF");
      pf_env = e ? std::max(0, atoi(e)) : 0;
    }
    b2.prefetch = pf_env;
    if (m_left) {  // one more item per cluster: (odd tile + phantom) x N tile x row slab
      int want_l = std::max(1, units / n_tiles);
      want_l = std::min(want_l, std::max(1, k_chunks_b / 4));
      b2.left_cps = (int)ceil_div(k_chunks_b, want_l);
      b2.left_n_split = (int)ceil_div(k_chunks_b, b2.left_cps);
    }
  } else {
    b2.n_split = 0;
  }
  // plan of the single-CTA kernel (all row tiles, or only the odd last one): plain CTAs or CTA pairs sharing the X^T
  // tile between two N tiles, whichever has the shorter critical path (rounds x chunks per item)
  TcParams b1 = b;
  if (m1) {
    b1.m_tiles = m1;
    b1.a_m0 = m2 * BM;
    auto plan = [&](int pair, int* n_split, int* cps) {
      const int units = pair ? ctx->sm_count / 2 : ctx->sm_count;
      const int items = m1 * (pair ? (n_tiles + 1) / 2 : n_tiles);
      int want = std::max(1, units / items);
      want = std::min(want, std::max(1, k_chunks_b / 4));  // keep >= 4 chunks per slab
      *cps = (int)ceil_div(k_chunks_b, want);
      *n_split = (int)ceil_div(k_chunks_b, *cps);
      const int rounds = (int)ceil_div((int64_t)items * *n_split, units);
      return (double)rounds * *cps * (pair ? 0.97 : 1.0);
    };
    int ns0, cps0, ns2 = 1, cps2 = 1;
    const double cost0 = plan(0, &ns0, &cps0);
    const double cost2 = (pairing_enabled() && n_tiles >= 2) ? plan(2, &ns2, &cps2) : 1e30;
    b1.pair = cost2 < cost0 ? 2 : 0;
    b1.n_split = b1.pair ? ns2 : ns0;
    b1.chunks_per_split = b1.pair ? cps2 : cps0;
    b1.stages = stages;
    if (l2hint) {
      const int a_reuse = b1.pair == 2 ? (n_tiles + 1) / 2 : n_tiles;
      b1.hint_a = a_reuse <= 1 ? 1 : 0;
      b1.hint_b = m1 >= 4 ? 2 : 0;
    }
  } else {
    b1.n_split = 0;
  }
  // ---- swapped operand roles + interleaved stream-K (softmax_bwd_sk.cuh): chain-class rows on M, features on N ----
  // BHMC_BWD_SK: 0 = off, 1 = where the cost model prefers it (default), 2 = wherever the shape allows (tests)
  // BHMC_SK_WH (per cent, default 90): cost of an M = 128 item relative to an M = 256 one of the same width -- measured
  // 154.9 / 146.1 / 148.6 / 153.6 us per 64-chain launch at 70 / 90 / 100 / 110; BHMC_SK_T=1: odd tile by transposed items where the plan finds them cheaper
  static int sk_env = -1;
  static SkTune sk_tune;
  if (sk_env < 0) {
    const char* e = getenv("BHMC_BWD_SK");
    sk_env = e ? atoi(e) : 1;
    const char* w = getenv("BHMC_SK_WH");
    sk_tune.wh_pct = w ? std::max(10, std::min(300, atoi(w))) : 90;
    const char* f = getenv("BHMC_SK_WFLOOR");
    sk_tune.w_floor = f ? std::max(1, std::min(20, atoi(f))) : 5;
    const char* t = getenv("BHMC_SK_T");
    sk_tune.transposed = t ? atoi(t) : 0;
    const char* q = getenv("BHMC_SK_WQ");
    sk_tune.w_q = q ? std::max(1, std::min(40, atoi(q))) : 10;
    const char* r = getenv("BHMC_SK_WR");
    sk_tune.w_r_floor = r ? std::max(1, std::min(40, atoi(r))) : 8;
  }
  SkParams sk{};
  bool use_sk = false;
  if (sk_env && pairing_enabled() && BK == 64 && d.slab == BK && d.slab_ld == BK && dm_slab == BK && dm_ld == BK &&
      k_chunks_b >= 64 && ncols >= 2 * BM) {
    SkPlan& s = sk.s;
    // tensor work per chunk in units of 1/10 of a 256 x 160 pair chunk; a 128 x 160 tile chunk of the single-CTA kernel
    // costs 5 x ~1.25 (1250 against 1005 cycles)
    if (sk_make_plan(ncols, (int)d.Dt, k_chunks_b, ctx->sm_count / 2, sk_tune, &s) &&
        (sk_env >= 2 || (double)sk_plan_cost(s) < 1.25 * 5.0 * m_tiles_all * (double)pcol / 160.0)) {
      sk.split3 = smode;
      sk.sub_chunks = b.sub_chunks;
      sk.a_chunk0 = (int)((row0 - shift) / BK);
      sk.xt_rows = (int)d.Dt_pad;
      sk.dm_rows = (int)dm_rows;
      sk.stages = std::max(2, std::min(MAX_STAGES, (int)((225 * 1024) / sk_stage(s, smode).bytes)));
      if (const char* e = getenv("BHMC_STAGES")) sk.stages = std::max(1, std::min(sk.stages, atoi(e)));
      use_sk = true;
    }
  }
  const int64_t rows2 = (int64_t)m2 * BM, rows1 = (int64_t)m1 * BM;
  void* part = nullptr;
  const size_t part2_elems = (size_t)b2.n_split * rows2 * pcol;
  const size_t part1_elems = m2 ? (size_t)b2.left_n_split * BM * pcol : (size_t)b1.n_split * rows1 * pcol;
  const size_t sk_elems = use_sk ? (size_t)sk.s.n_clusters * SK_MAX_PIECES * sk.s.piece_elems : 0;
  BHMC_TRY(ctx->get_scratch(3, sizeof(float) * std::max(part2_elems + part1_elems, sk_elems), &part));
  sk.part = (float*)part;
  b2.part = (float*)part;
  b2.part_left = (float*)part + part2_elems;
  b1.part = (float*)part + part2_elems;
  if (want_prof) {
    BHMC_CUDA_OK(cudaMemsetAsync(prof_dev, 0, sizeof(long long) * 8 * 1024, ctx->stream));
    b2.prof = (long long*)prof_dev;
    if (!m2) b1.prof = (long long*)prof_dev;
  }
  {
    GroupTimer t(ctx, KG_BWD);
    // every column a backward chunk reads is written by the forward epilogue (rows, alignment prefix, tail)
    if (use_sk) {
      SkMaps maps;
      const uint64_t dm_outer = (uint64_t)(dm_nslab * dm_rows);
      const bool has_dlo = smode != 0, has_xlo = smode == 1;
      auto dmap = [&](int at, uint32_t box) -> int {  // DmT hi (+ lo) with a `box`-row box
        BHMC_TRY(make_map(&maps.m[at], dmt_hi, (uint64_t)BK, dm_outer, (uint64_t)BK, box));
        maps.m[at + 1] = maps.m[at];
        if (has_dlo) BHMC_TRY(make_map(&maps.m[at + 1], dmt_lo, (uint64_t)BK, dm_outer, (uint64_t)BK, box));
        return BHMC_OK;
      };
      auto xmap = [&](int at, uint32_t box) -> int {  // X^T hi (+ lo)
        BHMC_TRY(make_map(&maps.m[at], d.Xt_hi, (uint64_t)BK, xt_rows, (uint64_t)BK, box));
        maps.m[at + 1] = maps.m[at];
        if (has_xlo) BHMC_TRY(make_map(&maps.m[at + 1], d.Xt_lo, (uint64_t)BK, xt_rows, (uint64_t)BK, box));
        return BHMC_OK;
      };
      BHMC_TRY(dmap(SKM_D128, BM));
      BHMC_TRY(dmap(SKM_D64, BM / 2));
      BHMC_TRY(xmap(SKM_XBN, (uint32_t)(sk.s.bn / 2)));
      BHMC_TRY(xmap(SKM_X128, BM));
      BHMC_TRY(xmap(SKM_XR, (uint32_t)(sk.s.bnr ? sk.s.bnr / 2 : 16)));
      if (want_prof) sk.prof = (long long*)prof_dev;
      BHMC_TRY(launch_bwd_sk(ctx, maps, sk));
    }
    if (m2 && !use_sk) {
      BHMC_TRY(make_map(&a_hi, d.Xt_hi, (uint64_t)d.slab, xt_rows, (uint64_t)d.slab_ld, BM));
      BHMC_TRY(make_map(&b_hi, dmt_hi, (uint64_t)dm_slab, (uint64_t)(dm_nslab * dm_rows), (uint64_t)dm_ld, (uint32_t)(BN / 2)));
      a_lo = a_hi;
      b_lo = b_hi;
      if (smode == 1) BHMC_TRY(make_map(&a_lo, d.Xt_lo, (uint64_t)d.slab, xt_rows, (uint64_t)d.slab_ld, BM));
      if (split3) BHMC_TRY(make_map(&b_lo, dmt_lo, (uint64_t)dm_slab, (uint64_t)(dm_nslab * dm_rows), (uint64_t)dm_ld, (uint32_t)(BN / 2)));
      BHMC_TRY(launch_bwd2(ctx, a_hi, a_lo, b_hi, b_lo, b2));
    }
    if (m1 && !use_sk) {
      const uint32_t abox = (uint32_t)(b1.pair ? BM / 2 : BM);
      BHMC_TRY(make_map(&a_hi, d.Xt_hi, (uint64_t)d.slab, xt_rows, (uint64_t)d.slab_ld, abox));
      BHMC_TRY(make_map(&b_hi, dmt_hi, (uint64_t)dm_slab, (uint64_t)(dm_nslab * dm_rows), (uint64_t)dm_ld, (uint32_t)BN));
      a_lo = a_hi;
      b_lo = b_hi;
      if (smode == 1) BHMC_TRY(make_map(&a_lo, d.Xt_lo, (uint64_t)d.slab, xt_rows, (uint64_t)d.slab_ld, abox));
      if (split3) BHMC_TRY(make_map(&b_lo, dmt_lo, (uint64_t)dm_slab, (uint64_t)(dm_nslab * dm_rows), (uint64_t)dm_ld, (uint32_t)BN));
      BHMC_TRY((launch_gemm_ew<MODE_BWD, 1, 16, false>(ctx, a_hi, a_lo, b_hi, b_lo, b1)));
    }
    if (want_prof) {
      std::vector<long long> hp(8 * 148);
      BHMC_CUDA_OK(cudaMemcpyAsync(hp.data(), prof_dev, sizeof(long long) * 8 * 148, cudaMemcpyDeviceToHost, ctx->stream));
      BHMC_CUDA_OK(cudaStreamSynchronize(ctx->stream));
      double tot = 0, te = 0, tf = 0, ti = 0, nc = 0, nt = 0, mx = 0;
      int n = 0;
      for (int bb = 0; bb < 148; ++bb)
        if (hp[bb * 8] > 0) {
          tot += hp[bb * 8], te += hp[bb * 8 + 1], tf += hp[bb * 8 + 2], ti += hp[bb * 8 + 3], nc += hp[bb * 8 + 4], nt += hp[bb * 8 + 5], ++n;
          mx = std::max(mx, (double)hp[bb * 8]);
        }
      const TcParams& bp = m2 ? b2 : b1;
      if (n && use_sk)
        fprintf(stderr, "[bhmc prof bwd] stream-K: pairs %d x %d tiles of %d, odd %d (Q %d, R width %d), items %d/%d/%d weights %d/%d/%d lanes %d/%d/%d, kc %d, T %d, stages %d | MMA thread, mean over %d CTAs: total %.0f cyc (max %.0f); wait tempty %.0f; wait full %.0f; issue+commit %.0f; chunks %.0f drains %.0f -> per chunk: full-wait %.0f issue %.0f\n",
                sk.s.n_pair, sk.s.n_nt, sk.s.bn, sk.s.odd, sk.s.n_fp, sk.s.bnr, sk.s.cnt[0], sk.s.cnt[1], sk.s.cnt[2], sk.s.w[0], sk.s.w[1],
                sk.s.w[2], sk.s.L[0], sk.s.L[1], sk.s.L[2], sk.s.kc, sk.s.T, sk.stages, n, tot / n, mx, te / n, tf / n, ti / n, nc / n, nt / n,
                tf / nc, ti / nc);
      else if (n)
        fprintf(stderr, "[bhmc prof bwd] %s n_split %d cps %d pair %d BN %d stages %d (+ %d odd tile: n_split %d cps %d) | MMA thread, mean over %d CTAs: total %.0f cyc (max %.0f); wait tempty %.0f; wait full %.0f; issue+commit %.0f; chunks %.0f drains %.0f -> per chunk: full-wait %.0f issue %.0f\n",
                m2 ? "cta_group::2" : "single", bp.n_split, bp.chunks_per_split, bp.pair, bp.BN, bp.stages, m_left,
                b2.left_n_split, b2.left_cps, n, tot / n, mx, te / n, tf / n, ti / n, nc / n, nt / n, tf / nc, ti / nc);
    }
    PartRegions pr{};
    pr.part0 = m2 ? b2.part : b1.part;
    pr.n_split0 = m2 ? b2.n_split : b1.n_split;
    pr.rows0 = m2 ? rows2 : rows1;
    pr.part1 = m2 ? b2.part_left : b1.part;
    pr.n_split1 = m2 ? b2.left_n_split : b1.n_split;
    pr.rows1 = m2 ? BM : rows1;
    pr.split_row = m2 ? rows2 : ((int64_t)1 << 40);
    pr.cols = pcol;
    pr.scale = 1.0f / d.x_scale;
    if (use_sk) pr.sk = sk.s, pr.part0 = sk.part;
    if (fs) {
      dim3 grid((unsigned)ceil_div(ceil_div(ld, 4), 256), C);
      if (fs->kind == BHMC_KIND_SGLD)
        k_tc_reduce_step<BHMC_KIND_SGLD><<<grid, 256, 0, ctx->stream>>>(pr, D, K, KP, P, ld, alpha, *fs, d.Dp, wt_hi,
                                                                        split3 ? wt_lo : nullptr, loglik);
      else
        k_tc_reduce_step<BHMC_KIND_SGD><<<grid, 256, 0, ctx->stream>>>(pr, D, K, KP, P, ld, alpha, *fs, d.Dp, wt_hi,
                                                                       split3 ? wt_lo : nullptr, loglik);
    } else if (fst) {
      dim3 grid((unsigned)ceil_div(ceil_div(ld, 4), 256), C);
      k_tc_reduce_stream<<<grid, 256, 0, ctx->stream>>>(pr, D, K, KP, P, ld, alpha, g, fst->u, fst->prep_next ? 1 : 0, d.Dp, wt_hi,
                                                       split3 ? wt_lo : nullptr, loglik);
    } else {
      dim3 grid((unsigned)ceil_div(ld, 256), C);
      k_tc_reduce<<<grid, 256, 0, ctx->stream>>>(pr, K, KP, P, q, g, ld, alpha);
    }
    ctx->launches++;
  }
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

#include "softmax_persist.cuh"

}  // namespace bhmc
