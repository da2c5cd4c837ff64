// Strided-batched GEMM on the 5th-generation tensor cores for the dropout MLP (hamiltonian/models/gpu/mlp.py):
//   C[z] = epilogue( A[z] (M x K) . B[z] (K x N) ),  z = chain, fp32 in / fp32 out,
// computed as bf16 hi/lo split products (3 tcgen05.mma per K step, fp32 accumulation in TMEM) or one bf16 pass.
// It takes the same GemmDesc (operand strides + fused epilogue: bias, alpha/2*W, ReLU gate, Philox / injected
// dropout + ReLU) as the CUDA-core kernel in mlp.cu, so the two paths are interchangeable per GEMM.
//
// Pipeline: (1) k_split_operand writes K-major bf16 hi/lo copies of A ([z][M][Kp]) and of B^T ([z][N][Kp]) -- the
// transposition, if any, is absorbed here through a 32x32 shared-memory tile; (2) k_tc_bgemm: persistent CTAs,
// warp 0 = TMA producer (3-D tensor maps: k, row, chain), warp 1 = MMA issuer (warp-uniform, elect.sync),
// warp 2 = TMEM allocator, 16 epilogue warps (tcgen05.ld -> fused epilogue -> fp32 stores, 64 B per thread per chunk);
// accumulators double-buffered in TMEM.  Same main loop as k_tc_gemm in softmax_tc.cu.
#include <cuda_bf16.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "internal.cuh"
#include "philox.cuh"

namespace bhmc {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int UMMA_K = 16;
static constexpr int MAX_STAGES = 8;
static constexpr int NON_EPI_THREADS = 128;
static constexpr int EW = 16;
static constexpr int TMEM_COLS = 512;
static constexpr int TMEM_BUF_COLS = 256;

#include "tc_common.cuh"
#include "mlp_common.cuh"

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

struct BgParams {
  int batch, m_tiles, n_tiles, k_chunks, BN, stages, split3;
  int a_shared, b_shared;  // operand identical for every chain (e.g. the data matrix)
  int vec;                 // epilogue may use 16-byte loads / stores (N % 4 == 0, strides % 4 == 0, 16 B aligned bases)
  int epi2;                // second epilogue form (per-warp shared-memory block behind the pipeline stages)
  int debug_epi;           // measurement only (BHMC_BG_DEBUG_EPI): 1 = the epilogue does nothing (main loop alone; results wrong)
  GemmDesc d;              // sizes + epilogue (the A/B pointers inside are unused here)
};

// Fused epilogue of one 128 x BN accumulator tile (thread = row; the EW/4 warps of a TMEM lane quarter split the 16-column
// chunks): bias, alpha/2*W, ReLU gate, Philox / injected dropout + ReLU, fp32 stores.
__device__ __forceinline__ void bg_epilogue_tile(const BgParams& p, uint32_t tacc, int z, int mt, int nt, int part, int t) {
  const GemmDesc& d = p.d;
  constexpr int PARTS = EW / 4;
  const int m = mt * BM + t;
  for (int j0 = part * 16; j0 < p.BN; j0 += 16 * PARTS) {
    uint32_t raw[16];
    tmem_ld<16>(tacc + (uint32_t)j0, raw);  // all lanes take part (.sync.aligned), stores are predicated below
    tmem_ld_wait();
    const int n0 = nt * p.BN + j0;
    if (m < d.M && n0 < d.N) {
      // Every input of the 16 outputs is fetched BEFORE the first store: with loads and stores interleaved per
      // element the compiler must keep them in program order (C may alias the inputs for all it knows), which
      // made the epilogue one dependent DRAM round trip per element (ncu: 104 us for the W1-gradient GEMM, tensor
      // pipe 14 %).  p.vec: row strides and base addresses allow 16-byte accesses.
      const int nq = min(4, (d.N - n0 + 3) >> 2);  // quads with at least one valid column
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
      if (d.bias) {
        const float* bp = d.bias + (int64_t)z * d.bias_batch + n0;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n0 + j < d.N) v[j] += __ldg(bp + j);
      }
      if (d.addsrc) {
        const float* ap = d.addsrc + (int64_t)z * d.add_batch + (int64_t)m * d.add_rs + n0;
        if (p.vec) {
          float4 t4[4];
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) t4[qd] = qd < nq ? __ldg(reinterpret_cast<const float4*>(ap) + qd) : make_float4(0, 0, 0, 0);
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            v[4 * qd] = fmaf(d.add_scale, t4[qd].x, v[4 * qd]);
            v[4 * qd + 1] = fmaf(d.add_scale, t4[qd].y, v[4 * qd + 1]);
            v[4 * qd + 2] = fmaf(d.add_scale, t4[qd].z, v[4 * qd + 2]);
            v[4 * qd + 3] = fmaf(d.add_scale, t4[qd].w, v[4 * qd + 3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + j < d.N) v[j] = fmaf(d.add_scale, __ldg(ap + j), v[j]);
        }
      }
      if (d.gate) {
        const float* gp = d.gate + (int64_t)z * d.gate_batch + (int64_t)m * d.gate_rs + n0;
        if (p.vec) {
          float4 t4[4];
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) t4[qd] = qd < nq ? __ldg(reinterpret_cast<const float4*>(gp) + qd) : make_float4(0, 0, 0, 0);
#pragma unroll
          for (int qd = 0; qd < 4; ++qd) {
            v[4 * qd] = t4[qd].x > 0.f ? v[4 * qd] * d.gate_scale : 0.f;
            v[4 * qd + 1] = t4[qd].y > 0.f ? v[4 * qd + 1] * d.gate_scale : 0.f;
            v[4 * qd + 2] = t4[qd].z > 0.f ? v[4 * qd + 2] * d.gate_scale : 0.f;
            v[4 * qd + 3] = t4[qd].w > 0.f ? v[4 * qd + 3] * d.gate_scale : 0.f;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + j < d.N) v[j] = __ldg(gp + j) > 0.f ? v[j] * d.gate_scale : 0.f;
        }
      }
      uint32_t ka = 0xFFFFu, kb = 0xFFFFu;  // keep bits of the 16 columns
      if (d.epi >= 1) {
        ka = 0, kb = 0;
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          if (qd >= nq) break;
          const int n = n0 + 4 * qd;
          uint32_t k8 = 0xF0u;
          if (!d.mask_a || (d.epi == 2 && !d.mask_b)) k8 = keep_bits8(d, z, m, n);
          uint32_t a4 = k8 & 15u, b4 = k8 >> 4;
          if (d.mask_a) {
            a4 = 0;
            for (int j = 0; j < 4; ++j)
              if (n + j < d.N && d.mask_a[(int64_t)z * d.mask_batch + (int64_t)m * d.N + n + j]) a4 |= 1u << j;
          }
          if (d.epi == 2 && d.mask_b) {
            b4 = 0;
            for (int j = 0; j < 4; ++j)
              if (n + j < d.N && d.mask_b[(int64_t)z * d.mask_batch + (int64_t)m * d.N + n + j]) b4 |= 1u << j;
          }
          ka |= a4 << (4 * qd);
          kb |= b4 << (4 * qd);
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float x = v[j];
        if (d.epi >= 1) {
          x = ((ka >> j) & 1u) ? x * d.keep_inv : 0.f;
          x = fmaxf(x, 0.f);
          if (d.epi == 2) x = ((kb >> j) & 1u) ? x * d.keep_inv : 0.f;
        }
        v[j] = x;
      }
      float* cp = d.C + (int64_t)z * d.c_batch + (int64_t)m * d.c_rs + n0;
      if (p.vec) {
#pragma unroll
        for (int qd = 0; qd < 4; ++qd)
          if (qd < nq) reinterpret_cast<float4*>(cp)[qd] = make_float4(v[4 * qd], v[4 * qd + 1], v[4 * qd + 2], v[4 * qd + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n0 + j < d.N) cp[j] = v[j];
      }
      // bf16 hi/lo operand copies for the GEMMs that consume this result (GemmDesc::ck_* / ct_*)
      if (d.ck_hi || d.ct_hi) {
        __nv_bfloat16 hb[16], lb[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          hb[j] = __float2bfloat16_rn(v[j]);
          lb[j] = __float2bfloat16_rn(v[j] - __bfloat162float(hb[j]));
        }
        if (d.ck_hi) {  // this row, 16 consecutive columns: 32 bytes per copy
          const int64_t o = (int64_t)z * d.ck_batch + (int64_t)m * d.ck_ld + n0;
          if (p.vec && (d.ck_ld & 7) == 0 && n0 + 16 <= d.N) {
            reinterpret_cast<uint4*>(d.ck_hi + o)[0] = reinterpret_cast<const uint4*>(hb)[0];
            reinterpret_cast<uint4*>(d.ck_hi + o)[1] = reinterpret_cast<const uint4*>(hb)[1];
            if (d.ck_lo) {
              reinterpret_cast<uint4*>(d.ck_lo + o)[0] = reinterpret_cast<const uint4*>(lb)[0];
              reinterpret_cast<uint4*>(d.ck_lo + o)[1] = reinterpret_cast<const uint4*>(lb)[1];
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n0 + j < d.N) {
                d.ck_hi[o + j] = hb[j];
                if (d.ck_lo) d.ck_lo[o + j] = lb[j];
              }
          }
        }
        if (d.ct_hi) {  // transposed: for every column the warp's 32 rows are 64 contiguous bytes
          const int64_t o = (int64_t)z * d.ct_batch + (int64_t)n0 * d.ct_ld + m;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + j < d.N) {
              d.ct_hi[o + (int64_t)j * d.ct_ld] = hb[j];
              if (d.ct_lo) d.ct_lo[o + (int64_t)j * d.ct_ld] = lb[j];
            }
        }
      }
    }
  }

}


// ---------------------------------------------------------------------------------------------
// Epilogue, second form (round 2, default when p.vec; BHMC_BG_EPI2=0 selects the one above).  Measured at cfg4 with the
// epilogue switched off (BHMC_BG_DEBUG_EPI=1): 72 of the 164 us the five GEMMs of an evaluation take are epilogue.  In
// the TMEM layout a thread owns a row, so every load / store instruction of a warp touches 32 different rows (32 L2
// sectors, each half used), and the inputs (alpha/2*W, ReLU gate) were fetched only after the accumulator was complete
// -- one exposed DRAM round trip per 16-column chunk.  Here
//   * the inputs of a tile are fetched BEFORE the wait on the accumulator barrier (two chunks in flight);
//   * the accumulator chunk goes through a per-warp 32 x 16 fp32 block in shared memory (quad swizzle, conflict free both
//     ways) and is processed in memory order: lane l owns columns 4(l%4)..+3 of rows 8i + l/4, i = 0..3 -- every global
//     access of a warp covers 8 rows x 64 contiguous bytes (full sectors), the bf16 K-major copies 8 x 32 bytes;
//   * the transposed bf16 copies (ct_*) read the finished block back in the row layout (64 contiguous bytes per store).
// Same arithmetic per element in the same order, same Philox keys: results are bit-identical to the first form.
struct EpiPre {
  float4 a[4];  // addsrc or gate values of the lane's four row groups
};

__device__ __forceinline__ void bg_epi2_prefetch(const BgParams& p, int z, int mt, int nt, int chunk, int lane, EpiPre& pre) {
  const GemmDesc& d = p.d;
  const float* src = d.addsrc ? d.addsrc + (int64_t)z * d.add_batch : (d.gate ? d.gate + (int64_t)z * d.gate_batch : nullptr);
  if (!src) return;
  const int64_t rs = d.addsrc ? d.add_rs : d.gate_rs;
  const int n = nt * p.BN + chunk * 16 + 4 * (lane & 3);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = mt * BM + (threadIdx.x & 96) + 8 * i + (lane >> 2);  // (threadIdx.x & 96) = 32 * TMEM lane quarter (NON_EPI_THREADS = 128)
    pre.a[i] = (m < d.M && n < d.N) ? __ldg(reinterpret_cast<const float4*>(src + (int64_t)m * rs + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <bool UPD>
__device__ __forceinline__ void bg_epi2_chunk(const BgParams& p, uint32_t tacc, int z, int mt, int nt, int chunk, int lane,
                                              float4* blk, const EpiPre& pre) {
  const GemmDesc& d = p.d;
  const int row0 = mt * BM + (threadIdx.x & 96);
  // (1) accumulator chunk -> shared memory, thread = row
  {
    uint32_t raw[16];
    tmem_ld<16>(tacc + (uint32_t)(chunk * 16), raw);
    tmem_ld_wait();
    const int sw = (lane >> 1) & 3;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      blk[lane * 4 + (q ^ sw)] = make_float4(__uint_as_float(raw[4 * q]), __uint_as_float(raw[4 * q + 1]), __uint_as_float(raw[4 * q + 2]),
                                             __uint_as_float(raw[4 * q + 3]));
  }
  __syncwarp();
  // (2) memory order: lane = (row group l/4, column quad l%4)
  const int q = lane & 3;
  const int n = nt * p.BN + chunk * 16 + 4 * q;
  const bool n_ok = n < d.N;
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (d.bias && n_ok) bias4 = __ldg(reinterpret_cast<const float4*>(d.bias + (int64_t)z * d.bias_batch + n));
  const bool want_t = d.ct_hi != nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 8 * i + (lane >> 2);
    const int m = row0 + r;
    const int slot = r * 4 + (q ^ ((r >> 1) & 3));
    float4 v = blk[slot];
    if (m < d.M && n_ok) {
      v.x += bias4.x, v.y += bias4.y, v.z += bias4.z, v.w += bias4.w;
      if (d.addsrc) {
        v.x = fmaf(d.add_scale, pre.a[i].x, v.x), v.y = fmaf(d.add_scale, pre.a[i].y, v.y);
        v.z = fmaf(d.add_scale, pre.a[i].z, v.z), v.w = fmaf(d.add_scale, pre.a[i].w, v.w);
      } else if (d.gate) {
        v.x = pre.a[i].x > 0.f ? v.x * d.gate_scale : 0.f, v.y = pre.a[i].y > 0.f ? v.y * d.gate_scale : 0.f;
        v.z = pre.a[i].z > 0.f ? v.z * d.gate_scale : 0.f, v.w = pre.a[i].w > 0.f ? v.w * d.gate_scale : 0.f;
      }
      if (d.epi >= 1) {
        uint32_t ka = 0, kb = 0xFu;
        if (!d.mask_a || (d.epi == 2 && !d.mask_b)) {
          const uint32_t k8 = keep_bits8(d, z, m, n);
          ka = k8 & 15u, kb = k8 >> 4;
        }
        if (d.mask_a) {
          const uint8_t* mp = d.mask_a + (int64_t)z * d.mask_batch + (int64_t)m * d.N + n;
          ka = (mp[0] ? 1u : 0u) | (mp[1] ? 2u : 0u) | (mp[2] ? 4u : 0u) | (mp[3] ? 8u : 0u);
        }
        if (d.epi == 2 && d.mask_b) {
          const uint8_t* mp = d.mask_b + (int64_t)z * d.mask_batch + (int64_t)m * d.N + n;
          kb = (mp[0] ? 1u : 0u) | (mp[1] ? 2u : 0u) | (mp[2] ? 4u : 0u) | (mp[3] ? 8u : 0u);
        }
        auto act = [&](float x, uint32_t bit) {
          x = (ka & bit) ? x * d.keep_inv : 0.f;
          x = fmaxf(x, 0.f);
          if (d.epi == 2) x = (kb & bit) ? x * d.keep_inv : 0.f;
          return x;
        };
        v.x = act(v.x, 1u), v.y = act(v.y, 2u), v.z = act(v.z, 4u), v.w = act(v.w, 8u);
      }
      if (UPD && d.upd_on) {
        // v is the gradient slice (acc + alpha/2 q): apply the sampler's update instead of storing it; from here on v is
        // the NEW position (the transposed operand copy below is taken from it)
        const float ge[4] = {v.x, v.y, v.z, v.w};
        float qe[4] = {pre.a[i].x, pre.a[i].y, pre.a[i].z, pre.a[i].w};
        upd_apply4(d.upd, z, d.upd_off + (int64_t)m * d.add_rs + n, ge, qe);
        v = make_float4(qe[0], qe[1], qe[2], qe[3]);
      } else {
        *reinterpret_cast<float4*>(d.C + (int64_t)z * d.c_batch + (int64_t)m * d.c_rs + n) = v;
      }
      if (d.ck_hi) {
        const int64_t o = (int64_t)z * d.ck_batch + (int64_t)m * d.ck_ld + n;
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y), h2 = __float2bfloat16_rn(v.z),
                            h3 = __float2bfloat16_rn(v.w);
        __nv_bfloat162 a = __halves2bfloat162(h0, h1), b = __halves2bfloat162(h2, h3);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&a), pk.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(d.ck_hi + o) = pk;
        if (d.ck_lo) {
          a = __halves2bfloat162(__float2bfloat16_rn(v.x - __bfloat162float(h0)), __float2bfloat16_rn(v.y - __bfloat162float(h1)));
          b = __halves2bfloat162(__float2bfloat16_rn(v.z - __bfloat162float(h2)), __float2bfloat16_rn(v.w - __bfloat162float(h3)));
          pk.x = *reinterpret_cast<uint32_t*>(&a), pk.y = *reinterpret_cast<uint32_t*>(&b);
          *reinterpret_cast<uint2*>(d.ck_lo + o) = pk;
        }
      }
    }
    if (want_t) blk[slot] = v;
  }
  __syncwarp();
  // (3) transposed bf16 copies: back to thread = row; for every column the warp's 32 rows are 64 contiguous bytes
  if (want_t) {
    const int m = row0 + lane;
    const int n0 = nt * p.BN + chunk * 16;
    const int sw = (lane >> 1) & 3;
    // fused update: a chain that does not take part in the drift keeps its position -- and its copy
    const bool moved = !UPD || !d.upd_on || (d.upd.pre_len > 0 && d.upd.it_pre < d.upd.L[z] - 1);
    if (m < d.M && moved) {
      const int64_t o = (int64_t)z * d.ct_batch + (int64_t)n0 * d.ct_ld + m;
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        if (n0 + 4 * qq >= d.N) break;
        const float4 v = blk[lane * 4 + (qq ^ sw)];
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __nv_bfloat16 h = __float2bfloat16_rn(e[j]);
          d.ct_hi[o + (int64_t)(4 * qq + j) * d.ct_ld] = h;
          if (d.ct_lo) d.ct_lo[o + (int64_t)(4 * qq + j) * d.ct_ld] = __float2bfloat16_rn(e[j] - __bfloat162float(h));
        }
      }
    }
    __syncwarp();
  }
}

// One tile: prefetch, wait for the accumulator, chunks.  `part` (0..EW/4-1) selects the warp's 16-column chunks.
template <bool UPD>
__device__ __forceinline__ void bg_epilogue_tile2(const BgParams& p, uint32_t tacc, int z, int mt, int nt, int part, int lane,
                                                  float4* blk, uint32_t bar, uint32_t parity) {
  constexpr int PARTS = EW / 4;
  const int n_chunks = (p.BN / 16 - part + PARTS - 1) / PARTS;  // chunks part, part + PARTS, ...
  EpiPre pre0, pre1;
  if (n_chunks > 0) bg_epi2_prefetch(p, z, mt, nt, part, lane, pre0);
  if (n_chunks > 1) bg_epi2_prefetch(p, z, mt, nt, part + PARTS, lane, pre1);
  mbar_wait(bar, parity);
  tcgen05_fence_after();
  if (p.debug_epi == 1) return;
  for (int k = 0; k < n_chunks; k += 2) {
    const int c0 = part + k * PARTS;
    if (nt * p.BN + c0 * 16 < p.d.N) bg_epi2_chunk<UPD>(p, tacc, z, mt, nt, c0, lane, blk, pre0);
    if (k + 2 < n_chunks) bg_epi2_prefetch(p, z, mt, nt, c0 + 2 * PARTS, lane, pre0);
    if (k + 1 < n_chunks) {
      const int c1 = c0 + PARTS;
      if (nt * p.BN + c1 * 16 < p.d.N) bg_epi2_chunk<UPD>(p, tacc, z, mt, nt, c1, lane, blk, pre1);
      if (k + 3 < n_chunks) bg_epi2_prefetch(p, z, mt, nt, c1 + 2 * PARTS, lane, pre1);
    }
  }
}

template <bool UPD>
__global__ void __launch_bounds__(NON_EPI_THREADS + 32 * EW, 1)
k_tc_bgemm(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
           const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, const BgParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_base_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a_bytes = BM * BK * 2, b_bytes = p.BN * BK * 2;
  const int nmat = p.split3 ? 2 : 1;
  const int stage_bytes = nmat * (a_bytes + b_bytes);
  const int tiles = p.m_tiles * p.n_tiles;
  const int num_work = p.batch * tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
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
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();  // programmatic dependent launch: only shared / tensor memory was touched so far

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        const int z = w / tiles, rem = w % tiles, mt = rem / p.n_tiles, nt = rem % p.n_tiles;
        const int za = p.a_shared ? 0 : z, zb = p.b_shared ? 0 : z;
        for (int k = 0; k < p.k_chunks; ++k) {
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
          const uint32_t full = smem_u32(&bar_full[stage]);
          mbar_expect_tx(full, (uint32_t)stage_bytes);
          const uint32_t sa = smem_base + stage * stage_bytes, sb = sa + nmat * a_bytes;
          tma_load_3d(sa, &tmA_hi, full, k * BK, mt * BM, za);
          if (p.split3) tma_load_3d(sa + a_bytes, &tmA_lo, full, k * BK, mt * BM, za);
          tma_load_3d(sb, &tmB_hi, full, k * BK, nt * p.BN, zb);
          if (p.split3) tma_load_3d(sb + b_bytes, &tmB_lo, full, k * BK, nt * p.BN, zb);
          if (++stage == p.stages) stage = 0, phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // warp-uniform MMA issue loop, one elected lane issues (see softmax_tc.cu)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait(smem_u32(&bar_tempty[buf]), (use & 1u) ^ 1u);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(buf * TMEM_BUF_COLS);
      for (int k = 0; k < p.k_chunks; ++k) {
        mbar_wait(smem_u32(&bar_full[stage]), phase);
        tcgen05_fence_after();
        const uint32_t sa = smem_base + stage * stage_bytes;
        const uint32_t first = k > 0 ? 1u : 0u;
        if (p.split3) {
          const uint64_t a_hi = make_smem_desc(sa), a_lo = make_smem_desc(sa + a_bytes);
          const uint64_t b_hi = make_smem_desc(sa + 2 * a_bytes), b_lo = make_smem_desc(sa + 2 * a_bytes + b_bytes);
#pragma unroll
          for (int ks = 0; ks < BK / UMMA_K; ++ks) {
            const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);
            umma_bf16(tmem_d, a_hi + adv, b_hi + adv, idesc, ks > 0 ? 1u : first);
            umma_bf16(tmem_d, a_hi + adv, b_lo + adv, idesc, 1u);
            umma_bf16(tmem_d, a_lo + adv, b_hi + adv, idesc, 1u);
          }
        } else {
          const uint64_t a_hi = make_smem_desc(sa), b_hi = make_smem_desc(sa + a_bytes);
#pragma unroll
          for (int ks = 0; ks < BK / UMMA_K; ++ks) {
            const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);
            umma_bf16(tmem_d, a_hi + adv, b_hi + adv, idesc, ks > 0 ? 1u : first);
          }
        }
        umma_commit(smem_u32(&bar_empty[stage]));
        if (++stage == p.stages) stage = 0, phase ^= 1u;
      }
      umma_commit(smem_u32(&bar_tfull[buf]));
    }
  } else if (warp >= 4) {
    const int ew = warp & 3, part = (warp - 4) >> 2;
    const int t = ew * 32 + lane;
    int it = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++it) {
      const int z = w / tiles, rem = w % tiles, mt = rem / p.n_tiles, nt = rem % p.n_tiles;
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      const uint32_t tacc = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * TMEM_BUF_COLS);
      if (p.epi2) {
        float4* blk = reinterpret_cast<float4*>(smem_raw + (smem_base - smem_u32(smem_raw)) + p.stages * stage_bytes) + (warp - 4) * 128;
        bg_epilogue_tile2<UPD>(p, tacc, z, mt, nt, part, lane, blk, smem_u32(&bar_tfull[buf]), use & 1u);
      } else {
        mbar_wait(smem_u32(&bar_tfull[buf]), use & 1u);
        tcgen05_fence_after();
        if (p.debug_epi != 1) bg_epilogue_tile(p, tacc, z, mt, nt, part, t);
      }
      tcgen05_fence_before();
      mbar_arrive(smem_u32(&bar_tempty[buf]));
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// The same GEMM with cta_group::2 (round 2).  ncu of k_tc_bgemm at cfg4: tensor pipe 29-37 %.  A 128 x 128 tile streams
// 64 KB per K chunk (A and B, hi and lo) for 12 MMAs of 64 cycles -- the ~50 B/clk an SM ingests allow a chunk every
// ~1 200 cycles where the MMAs need 768.  Here a cluster of two CTAs owns TWO adjacent 128-row tiles (M = 256) and one
// N tile up to 256 wide: every CTA stages its own A tile and only HALF of the B tile (32 KB + BN/2 rows), the leader's
// elected thread issues tcgen05.mma.cta_group::2.  At BN = 256 that is 64 KB per CTA for 12 MMAs of 128 cycles: the
// main loop is bound by the tensor pipe again.  Barrier protocol as in k_tc_fwd2 (softmax_tc.cu).
template <bool UPD>
__global__ void __launch_bounds__(NON_EPI_THREADS + 32 * EW, 1)
k_tc_bgemm2(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
            const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, const BgParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_base_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const bool leader = rank == 0;
  const int a_bytes = BM * BK * 2, bh_bytes = (p.BN / 2) * BK * 2;  // this CTA's half of the B tile
  const int nmat = p.split3 ? 2 : 1;
  const int stage_bytes = nmat * (a_bytes + bh_bytes);
  const int m_pairs = (p.m_tiles + 1) / 2;
  const int tiles = m_pairs * p.n_tiles;
  const int num_work = p.batch * tiles;
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
  pdl_wait();  // only shared / tensor memory was touched so far

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer (both CTAs) =====
      int stage = 0;
      uint32_t phase = 0;
      for (int w = wi0; w < num_work; w += wi_step) {
        const int z = w / tiles, rem = w % tiles, mt = 2 * (rem / p.n_tiles) + rank, nt = rem % p.n_tiles;
        const int za = p.a_shared ? 0 : z, zb = p.b_shared ? 0 : z;
        for (int k = 0; k < p.k_chunks; ++k) {
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
          const uint32_t full = smem_u32(&bar_full[stage]);  // same offset in the leader CTA
          if (leader) mbar_expect_tx(full, (uint32_t)(2 * stage_bytes));
          const uint32_t sa = smem_base + stage * stage_bytes, sb = sa + nmat * a_bytes;
          // rows beyond M (phantom half of an odd last pair) and beyond N are out of bounds of the map: zero fill
          tma_load_3d_2sm(sa, &tmA_hi, full, k * BK, mt * BM, za);
          if (p.split3) tma_load_3d_2sm(sa + a_bytes, &tmA_lo, full, k * BK, mt * BM, za);
          tma_load_3d_2sm(sb, &tmB_hi, full, k * BK, nt * p.BN + rank * (p.BN / 2), zb);
          if (p.split3) tma_load_3d_2sm(sb + bh_bytes, &tmB_lo, full, k * BK, nt * p.BN + rank * (p.BN / 2), zb);
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
      for (int w = wi0; w < num_work; w += wi_step, ++it) {
        const int buf = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait(smem_u32(&bar_tempty[buf]), (use & 1u) ^ 1u);  // both CTAs' epilogues have drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * TMEM_BUF_COLS);
        for (int k = 0; k < p.k_chunks; ++k) {
          mbar_wait(smem_u32(&bar_full[stage]), phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint32_t first = k > 0 ? 1u : 0u;
          if (p.split3) {
            const uint64_t a_hi = make_smem_desc(sa), a_lo = make_smem_desc(sa + a_bytes);
            const uint64_t b_hi = make_smem_desc(sa + 2 * a_bytes), b_lo = make_smem_desc(sa + 2 * a_bytes + bh_bytes);
#pragma unroll
            for (int ks = 0; ks < BK / UMMA_K; ++ks) {
              const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);
              umma_bf16_2sm(tmem_d, a_hi + adv, b_hi + adv, idesc, ks > 0 ? 1u : first);
              umma_bf16_2sm(tmem_d, a_hi + adv, b_lo + adv, idesc, 1u);
              umma_bf16_2sm(tmem_d, a_lo + adv, b_hi + adv, idesc, 1u);
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
          if (++stage == p.stages) stage = 0, phase ^= 1u;
        }
        umma_commit_2sm(smem_u32(&bar_tfull[buf]), 3);  // accumulator halves complete in both CTAs
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs, each on its own 128 rows) =====
    const int ew = warp & 3, part = (warp - 4) >> 2;
    const int t = ew * 32 + lane;
    int it = 0;
    for (int w = wi0; w < num_work; w += wi_step, ++it) {
      const int z = w / tiles, rem = w % tiles, mt = 2 * (rem / p.n_tiles) + rank, nt = rem % p.n_tiles;
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      const uint32_t tacc = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * TMEM_BUF_COLS);
      if (p.epi2) {  // rows of a phantom tile fail the m < M tests
        float4* blk = reinterpret_cast<float4*>(smem_raw + (smem_base - smem_u32(smem_raw)) + p.stages * stage_bytes) + (warp - 4) * 128;
        bg_epilogue_tile2<UPD>(p, tacc, z, mt, nt, part, lane, blk, smem_u32(&bar_tfull[buf]), use & 1u);
      } else {
        mbar_wait(smem_u32(&bar_tfull[buf]), use & 1u);
        tcgen05_fence_after();
        if (p.debug_epi != 1) bg_epilogue_tile(p, tacc, z, mt, nt, part, t);
      }
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
// Two independent batched GEMMs in ONE launch (round 2).  At cfg4 a GEMM is 256 tiles for 148 CTAs: 1.7 tiles per CTA, so
// its launch ramp, pipeline fill and the exposed epilogue of the last tile are paid per GEMM.  gW2 = dA2^T H1 and
// dA1 = dA2 W2 * gate depend on the same inputs and not on each other: as one work list of 512 tiles a CTA runs 3-4 tiles
// back to back (every epilogue but the last under the next main loop) and one ramp is gone.  Same tile code as k_tc_bgemm;
// the two problems must agree on BN, pipeline depth, split mode and epilogue form (the host checks).
struct BgGroup {
  CUtensorMap m[2][4];  // A_hi, A_lo, B_hi, B_lo of each problem
  BgParams p[2];
  int work0;            // work items of problem 0 (they come first in the list)
  int work;             // all work items
};

__global__ void __launch_bounds__(NON_EPI_THREADS + 32 * EW, 1) k_tc_bgemm_grp(const __grid_constant__ BgGroup g) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_base_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = g.p[0].BN, stages = g.p[0].stages, split3 = g.p[0].split3;
  const int a_bytes = BM * BK * 2, b_bytes = BN * BK * 2;
  const int nmat = split3 ? 2 : 1;
  const int stage_bytes = nmat * (a_bytes + b_bytes);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
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
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < g.work; w += gridDim.x) {
        const int pi = w >= g.work0 ? 1 : 0;
        const BgParams& p = g.p[pi];
        const int wl = w - (pi ? g.work0 : 0), tiles = p.m_tiles * p.n_tiles;
        const int z = wl / tiles, rem = wl % tiles, mt = rem / p.n_tiles, nt = rem % p.n_tiles;
        const int za = p.a_shared ? 0 : z, zb = p.b_shared ? 0 : z;
        for (int k = 0; k < p.k_chunks; ++k) {
          mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u);
          const uint32_t full = smem_u32(&bar_full[stage]);
          mbar_expect_tx(full, (uint32_t)stage_bytes);
          const uint32_t sa = smem_base + stage * stage_bytes, sb = sa + nmat * a_bytes;
          tma_load_3d(sa, &g.m[pi][0], full, k * BK, mt * BM, za);
          if (split3) tma_load_3d(sa + a_bytes, &g.m[pi][1], full, k * BK, mt * BM, za);
          tma_load_3d(sb, &g.m[pi][2], full, k * BK, nt * BN, zb);
          if (split3) tma_load_3d(sb + b_bytes, &g.m[pi][3], full, k * BK, nt * BN, zb);
          if (++stage == stages) stage = 0, phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int w = blockIdx.x; w < g.work; w += gridDim.x, ++it) {
      const int k_chunks = g.p[w >= g.work0 ? 1 : 0].k_chunks;
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait(smem_u32(&bar_tempty[buf]), (use & 1u) ^ 1u);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(buf * TMEM_BUF_COLS);
      for (int k = 0; k < k_chunks; ++k) {
        mbar_wait(smem_u32(&bar_full[stage]), phase);
        tcgen05_fence_after();
        const uint32_t sa = smem_base + stage * stage_bytes;
        const uint32_t first = k > 0 ? 1u : 0u;
        if (split3) {
          const uint64_t a_hi = make_smem_desc(sa), a_lo = make_smem_desc(sa + a_bytes);
          const uint64_t b_hi = make_smem_desc(sa + 2 * a_bytes), b_lo = make_smem_desc(sa + 2 * a_bytes + b_bytes);
#pragma unroll
          for (int ks = 0; ks < BK / UMMA_K; ++ks) {
            const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);
            umma_bf16(tmem_d, a_hi + adv, b_hi + adv, idesc, ks > 0 ? 1u : first);
            umma_bf16(tmem_d, a_hi + adv, b_lo + adv, idesc, 1u);
            umma_bf16(tmem_d, a_lo + adv, b_hi + adv, idesc, 1u);
          }
        } else {
          const uint64_t a_hi = make_smem_desc(sa), b_hi = make_smem_desc(sa + a_bytes);
#pragma unroll
          for (int ks = 0; ks < BK / UMMA_K; ++ks) {
            const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);
            umma_bf16(tmem_d, a_hi + adv, b_hi + adv, idesc, ks > 0 ? 1u : first);
          }
        }
        umma_commit(smem_u32(&bar_empty[stage]));
        if (++stage == stages) stage = 0, phase ^= 1u;
      }
      umma_commit(smem_u32(&bar_tfull[buf]));
    }
  } else if (warp >= 4) {
    const int ew = warp & 3, part = (warp - 4) >> 2;
    float4* blk = reinterpret_cast<float4*>(smem_raw + (smem_base - smem_u32(smem_raw)) + stages * stage_bytes) + (warp - 4) * 128;
    int it = 0;
    for (int w = blockIdx.x; w < g.work; w += gridDim.x, ++it) {
      const int pi = w >= g.work0 ? 1 : 0;
      const BgParams& p = g.p[pi];
      const int wl = w - (pi ? g.work0 : 0), tiles = p.m_tiles * p.n_tiles;
      const int z = wl / tiles, rem = wl % tiles, mt = rem / p.n_tiles, nt = rem % p.n_tiles;
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      const uint32_t tacc = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * TMEM_BUF_COLS);
      bg_epilogue_tile2<false>(p, tacc, z, mt, nt, part, lane, blk, smem_u32(&bar_tfull[buf]), use & 1u);
      tcgen05_fence_before();
      mbar_arrive(smem_u32(&bar_tempty[buf]));
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// dst_{hi,lo}[z][r][k] (k < Kp contiguous, Kp % 64 == 0) = split(src[z*sb + r*rs + k*cs]) for k < K, 0 for K <= k < Kp.
// One 32 (rows) x 64 (k) tile per block through shared memory, so both cs == 1 (row-major source) and rs == 1
// (transposed source) read coalesced; every thread writes a bf16 pair (128 B per warp and row).
template <int TR>  // tile rows (32 or 64): 64 keeps twice the loads in flight per thread
__global__ void __launch_bounds__(256) k_split_operand(const float* __restrict__ src, int64_t sb, int64_t rs, int64_t cs,
                                                       int R, int K, int64_t Kp, __nv_bfloat16* __restrict__ hi,
                                                       __nv_bfloat16* __restrict__ lo) {
  __shared__ float tile[TR][65];
  pdl_launch_dependents();  // lets a k_tc_bgemm launched with the attribute run its prologue under this kernel
  const int z = blockIdx.z;
  const int r0 = blockIdx.y * TR, k0 = blockIdx.x * 64;
  const float* s = src + (int64_t)z * sb;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  if (cs == 1 || rs != 1) {
#pragma unroll
    for (int i = ty; i < TR; i += 8) {
      const int r = r0 + i;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = k0 + tx + 32 * h;
        tile[i][tx + 32 * h] = (r < R && k < K) ? s[(int64_t)r * rs + (int64_t)k * cs] : 0.f;
      }
    }
  } else {  // rows are the contiguous index of the source: read along r
#pragma unroll
    for (int i = ty; i < 64; i += 8) {
      const int k = k0 + i;
#pragma unroll
      for (int h = 0; h < TR / 32; ++h) {
        const int r = r0 + tx + 32 * h;
        tile[tx + 32 * h][i] = (r < R && k < K) ? s[(int64_t)r * rs + (int64_t)k * cs] : 0.f;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = ty; i < TR; i += 8) {
    const int r = r0 + i, k = k0 + 2 * tx;
    if (r < R) {
      const float v0 = tile[i][2 * tx], v1 = tile[i][2 * tx + 1];
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
      const int64_t o = ((int64_t)z * R + r) * Kp + k;
      __nv_bfloat162 hv;
      hv.x = h0, hv.y = h1;
      *reinterpret_cast<__nv_bfloat162*>(hi + o) = hv;
      if (lo) {
        __nv_bfloat162 lv;
        lv.x = __float2bfloat16_rn(v0 - __bfloat162float(h0));
        lv.y = __float2bfloat16_rn(v1 - __bfloat162float(h1));
        *reinterpret_cast<__nv_bfloat162*>(lo + o) = lv;
      }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
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

// bf16 [Z][R][Kp] (K contiguous), box [1][box_rows][64], 128B swizzle, OOB rows / k read as zero
static int make_map3(CUtensorMap* m, const void* base, uint64_t K, uint64_t R, uint64_t Z, uint64_t Kp, uint32_t box_rows,
                     uint64_t zs = 0) {  // zs: elements between two batch entries (0 = R * Kp, densely packed)
  if (zs == 0) zs = R * Kp;
  // an encoded map is a pure function of these arguments and the scratch operands keep their addresses: the 20 driver
  // calls per MLP evaluation are served from a small cache (BHMC_MAP_CACHE=0 disables)
  struct Entry {
    const void* base;
    uint64_t K, R, Z, Kp, zs;
    uint32_t box;
    CUtensorMap m;
  };
  static thread_local std::vector<Entry> cache;
  static int cache_env = -1;
  if (cache_env < 0) {
    const char* e = getenv("BHMC_MAP_CACHE");
    cache_env = e ? atoi(e) : 1;
  }
  if (cache_env) {
    for (const Entry& e : cache)
      if (e.base == base && e.K == K && e.R == R && e.Z == Z && e.Kp == Kp && e.zs == zs && e.box == box_rows) {
        *m = e.m;
        return BHMC_OK;
      }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return BHMC_ERR_CUDA;
  }
  cuuint64_t dims[3] = {K, R, Z};
  cuuint64_t strides[2] = {Kp * 2, zs * 2};
  cuuint32_t box[3] = {(cuuint32_t)BK, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d) failed (%d): K=%llu R=%llu Z=%llu Kp=%llu box=%u", (int)r, (unsigned long long)K,
              (unsigned long long)R, (unsigned long long)Z, (unsigned long long)Kp, box_rows);
    return BHMC_ERR_CUDA;
  }
  if (cache_env) {
    if (cache.size() >= 64) cache.clear();
    cache.push_back(Entry{base, K, R, Z, Kp, zs, box_rows, *m});
  }
  return BHMC_OK;
}

// K-major bf16 hi/lo copy of a strided fp32 matrix: dst[z][r][k] = split(src[z*sb + r*rs + k*cs]), k < K (rows of Kp elements)
int tc_split_rows(bhmc_ctx* ctx, const float* src, int64_t sb, int64_t rs, int64_t cs, int R, int K, int64_t Kp, int Z,
                  __nv_bfloat16* hi, __nv_bfloat16* lo) {
  GroupTimer t(ctx, KG_PREP);
  dim3 g((unsigned)(Kp / 64), (unsigned)ceil_div(R, 32), (unsigned)Z);
  k_split_operand<32><<<g, 256, 0, ctx->stream>>>(src, sb, rs, cs, R, K, Kp, hi, lo);
  ctx->launches++;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

// Preparation of one batched GEMM; scratch slots 1..2 of the context hold the bf16 operand copies that are not handed in.
struct BgLaunch {  // one prepared batched GEMM: operand copies made, tensor maps encoded, kernel parameters filled
  CUtensorMap m[4];
  BgParams p;
  size_t smem;
  int work, grid;
  bool pair, upd;
  bool used_a, used_b;  // the A / B operand copy lives in scratch slot 1 / 2 of the context (valid until the next prepare)
};
static int bg_launch(bhmc_ctx* ctx, const BgLaunch& L);

static int bg_prepare(bhmc_ctx* ctx, const GemmDesc& d, int batch, bool split3, BgLaunch* L) {
  const int64_t Kp = round_up(d.K, 64);
  const bool a_shared = d.a_batch == 0, b_shared = d.b_batch == 0;
  const int za = a_shared ? 1 : batch, zb = b_shared ? 1 : batch;
  const size_t a_elems = (size_t)za * d.M * Kp, b_elems = (size_t)zb * d.N * Kp;
  // operands already split by their producer (GemmDesc::a_hi / b_hi) skip the split launch
  // (x_zs != 0: the copy is a strided view -- row pitch x_kp >= K, x_zs elements between chains -- e.g. the operand mirror
  // that the sampler's update kernel fills; K beyond the row is zero-filled by TMA, both strides are multiples of 16 bytes)
  const bool a_pre = d.a_hi != nullptr && (d.a_kp == Kp || (d.a_zs && d.a_kp >= d.K && d.a_kp % 8 == 0 && d.a_zs % 8 == 0)) && (!split3 || d.a_lo);
  const bool b_pre = d.b_hi != nullptr && (d.b_kp == Kp || (d.b_zs && d.b_kp >= d.K && d.b_kp % 8 == 0 && d.b_zs % 8 == 0)) && (!split3 || d.b_lo);
  void *sa = nullptr, *sb = nullptr;
  if (!a_pre) BHMC_TRY(ctx->get_scratch(1, a_elems * 2 * 2, &sa));
  if (!b_pre) BHMC_TRY(ctx->get_scratch(2, b_elems * 2 * 2, &sb));
  const __nv_bfloat16 *a_hi = a_pre ? d.a_hi : (const __nv_bfloat16*)sa, *a_lo = a_pre ? d.a_lo : a_hi + a_elems;
  const __nv_bfloat16 *b_hi = b_pre ? d.b_hi : (const __nv_bfloat16*)sb, *b_lo = b_pre ? d.b_lo : b_hi + b_elems;
  if (!a_pre || !b_pre) {
    GroupTimer t(ctx, KG_PREP);
    static int tr_env = -1;  // BHMC_SPLIT_TR: tile rows of the split kernel (32 or 64; A/B measurements)
    if (tr_env < 0) {
      const char* e = getenv("BHMC_SPLIT_TR");
      tr_env = e && atoi(e) == 64 ? 64 : 32;  // measured at cfg4: 44.3 k (32) vs 43.2 k (64) grad-evals/s
    }
    __nv_bfloat16 *wa_hi = (__nv_bfloat16*)sa, *wa_lo = wa_hi + a_elems, *wb_hi = (__nv_bfloat16*)sb, *wb_lo = wb_hi + b_elems;
    // B is (k, n) with strides (b_rs, b_cs): its K-major copy has rows n
    if (tr_env == 64) {
      dim3 ga((unsigned)(Kp / 64), (unsigned)ceil_div(d.M, 64), (unsigned)za), gb((unsigned)(Kp / 64), (unsigned)ceil_div(d.N, 64), (unsigned)zb);
      if (!a_pre) k_split_operand<64><<<ga, 256, 0, ctx->stream>>>(d.A, d.a_batch, d.a_rs, d.a_cs, d.M, d.K, Kp, wa_hi, split3 ? wa_lo : nullptr);
      if (!b_pre) k_split_operand<64><<<gb, 256, 0, ctx->stream>>>(d.B, d.b_batch, d.b_cs, d.b_rs, d.N, d.K, Kp, wb_hi, split3 ? wb_lo : nullptr);
    } else {
      dim3 ga((unsigned)(Kp / 64), (unsigned)ceil_div(d.M, 32), (unsigned)za), gb((unsigned)(Kp / 64), (unsigned)ceil_div(d.N, 32), (unsigned)zb);
      if (!a_pre) k_split_operand<32><<<ga, 256, 0, ctx->stream>>>(d.A, d.a_batch, d.a_rs, d.a_cs, d.M, d.K, Kp, wa_hi, split3 ? wa_lo : nullptr);
      if (!b_pre) k_split_operand<32><<<gb, 256, 0, ctx->stream>>>(d.B, d.b_batch, d.b_cs, d.b_rs, d.N, d.K, Kp, wb_hi, split3 ? wb_lo : nullptr);
    }
    ctx->launches += (a_pre ? 0 : 1) + (b_pre ? 0 : 1);
  }
  BgParams p{};
  p.batch = batch;
  p.BN = d.N >= 128 ? 128 : (int)round_up(d.N, 16);
  {
    // BHMC_BG_BN=auto: N-tile width from a rounds x operand-bytes model (a tile streams BM + BN rows per K chunk; the
    // persistent CTAs need ceil(items / SMs) rounds); BHMC_BG_BN=<n>: forced width for N >= n.  Default: 128.
    static int bn_env = -2;
    if (bn_env == -2) {
      const char* e = getenv("BHMC_BG_BN");
      bn_env = !e ? 0 : (e[0] == 'a' ? -1 : atoi(e));
    }
    const int m_tiles = (int)ceil_div(d.M, BM);
    if (bn_env > 0 && bn_env % 16 == 0 && bn_env <= 256 && d.N >= bn_env) p.BN = bn_env;
    // default: 128 when it divides N; otherwise the model picks (cfg4: N = 784 of the W1 gradient -> 35.0 -> 32.4 us,
    // while it loses 2-3 us on each of the N = 512 GEMMs)
    if ((bn_env == -1 || (bn_env == 0 && d.N % 128 != 0)) && d.N >= 128) {
      double best = 1e30;
      for (int bn = 128; bn <= 256; bn += 16) {
        const int stg = (int)((225 * 1024) / ((split3 ? 2 : 1) * (BM * BK * 2 + bn * BK * 2)));
        if (stg < 2) continue;
        const int64_t items = (int64_t)batch * m_tiles * ceil_div(d.N, bn);
        const double rounds = (double)ceil_div(items, (int64_t)ctx->sm_count);
        const double cost = rounds * (BM + bn + 32) * (stg == 2 ? 1.1 : 1.0);
        if (cost < best - 1e-9) best = cost, p.BN = bn;
      }
    }
  }
  // cta_group::2 kernel (BHMC_BG2=0: off): two row tiles per cluster and an N tile chosen by a cost model -- a chunk
  // takes max(12 MMAs of BN/2 cycles, (32 KB + BN/2 rows) at ~52 B/clk of ingest), the clusters need
  // ceil(items / clusters) rounds, plus the epilogue of a tile per round
  static int bg2_env = -1;
  if (bg2_env < 0) {
    const char* e = getenv("BHMC_BG2");
    bg2_env = e ? atoi(e) : 0;
  }
  p.m_tiles = (int)ceil_div(d.M, BM);
  p.k_chunks = (int)ceil_div(d.K, BK);
  const bool pair = bg2_env && p.m_tiles >= 2 && d.N >= 64;
  if (pair) {
    const int clusters = ctx->sm_count / 2, m_pairs = (p.m_tiles + 1) / 2;
    double best = 1e30;
    for (int bn = 64; bn <= 256; bn += 16) {
      const int st_b = (split3 ? 2 : 1) * (BM * BK * 2 + (bn / 2) * BK * 2);
      if ((225 * 1024) / st_b < 2) continue;
      const int64_t items = (int64_t)batch * m_pairs * ceil_div(d.N, bn);
      const double rounds = (double)ceil_div(items, (int64_t)clusters);
      const double chunk = std::max(6.0 * bn * (split3 ? 1.0 : 1.0 / 3.0), (double)st_b / 52.0);
      const double cost = rounds * (p.k_chunks * chunk + 400.0 + 12.0 * bn);
      if (cost < best - 1e-9) best = cost, p.BN = bn;
    }
    if (bg2_env >= 64 && bg2_env % 16 == 0 && bg2_env <= 256) p.BN = bg2_env;  // BHMC_BG2=<n>: forced N tile (A/B)
  }
  p.n_tiles = (int)ceil_div(d.N, p.BN);
  p.split3 = split3 ? 1 : 0;
  p.a_shared = a_shared;
  p.b_shared = b_shared;
  p.d = d;
  {
    static int dbg_env = -1;
    if (dbg_env < 0) {
      const char* e = getenv("BHMC_BG_DEBUG_EPI");
      dbg_env = e ? atoi(e) : 0;
    }
    p.debug_epi = dbg_env;
  }
  {
    auto al16 = [](const void* q) { return ((uintptr_t)q & 15u) == 0; };
    bool v = d.N % 4 == 0 && al16(d.C) && d.c_batch % 4 == 0 && d.c_rs % 4 == 0;
    if (d.addsrc) v = v && al16(d.addsrc) && d.add_batch % 4 == 0 && d.add_rs % 4 == 0;
    if (d.gate) v = v && al16(d.gate) && d.gate_batch % 4 == 0 && d.gate_rs % 4 == 0;
    p.vec = v ? 1 : 0;
  }
  const uint32_t b_box = (uint32_t)(pair ? p.BN / 2 : p.BN);
  const int stage_bytes = (split3 ? 2 : 1) * (BM * BK * 2 + (int)b_box * BK * 2);
  static int epi2_env = -1;
  if (epi2_env < 0) {
    const char* e = getenv("BHMC_BG_EPI2");
    epi2_env = e ? atoi(e) : 1;
  }
  const int epi_bytes = EW * 2048;  // per-warp 32 x 16 fp32 blocks of the second epilogue form
  p.epi2 = (epi2_env && p.vec && (!d.ck_hi || (d.ck_ld % 4 == 0 && d.ck_batch % 4 == 0 && (((uintptr_t)d.ck_hi | (uintptr_t)d.ck_lo) & 7u) == 0)) &&
            (!d.bias || (d.bias_batch % 4 == 0 && ((uintptr_t)d.bias & 15u) == 0)) && !(d.addsrc && d.gate)) ? 1 : 0;
  if (d.upd_on && (!p.epi2 || !d.addsrc || d.gate || d.bias || d.epi != 0 || d.ck_hi)) {
    set_error("fused sampler update needs the second epilogue form of k_tc_bgemm on a plain gradient GEMM");
    return BHMC_ERR_STATE;
  }
  p.stages = std::max(2, std::min(MAX_STAGES, (int)((225 * 1024 - (p.epi2 ? epi_bytes : 0)) / stage_bytes)));
  CUtensorMap &mA_hi = L->m[0], &mA_lo = L->m[1], &mB_hi = L->m[2], &mB_lo = L->m[3];
  const uint64_t a_pitch = a_pre ? (uint64_t)d.a_kp : (uint64_t)Kp, b_pitch = b_pre ? (uint64_t)d.b_kp : (uint64_t)Kp;
  const uint64_t a_zs = a_pre ? (uint64_t)d.a_zs : 0, b_zs = b_pre ? (uint64_t)d.b_zs : 0;
  BHMC_TRY(make_map3(&mA_hi, a_hi, (uint64_t)d.K, (uint64_t)d.M, (uint64_t)za, a_pitch, BM, a_zs));
  BHMC_TRY(make_map3(&mB_hi, b_hi, (uint64_t)d.K, (uint64_t)d.N, (uint64_t)zb, b_pitch, b_box, b_zs));
  if (split3) {
    BHMC_TRY(make_map3(&mA_lo, a_lo, (uint64_t)d.K, (uint64_t)d.M, (uint64_t)za, a_pitch, BM, a_zs));
    BHMC_TRY(make_map3(&mB_lo, b_lo, (uint64_t)d.K, (uint64_t)d.N, (uint64_t)zb, b_pitch, b_box, b_zs));
  } else {
    mA_lo = mA_hi;
    mB_lo = mB_hi;
  }
  L->smem = (size_t)p.stages * stage_bytes + 1024 + (p.epi2 ? epi_bytes : 0);
  L->work = batch * (pair ? (p.m_tiles + 1) / 2 : p.m_tiles) * p.n_tiles;
  L->grid = pair ? 2 * std::min(L->work, ctx->sm_count / 2) : std::min(L->work, ctx->sm_count);
  L->pair = pair;
  L->upd = d.upd_on != 0;
  L->used_a = !a_pre, L->used_b = !b_pre;
  L->p = p;
  return BHMC_OK;
}

static bool pdl_mlp_enabled() {
  // programmatic dependent launch of the batched GEMM: its prologue runs under the split kernel in front of it
  // (MLP parity tests green with it; cfg4 43.8 k -> 45.4 k grad-evals/s in one A/B pair).  BHMC_PDL_MLP=0: plain launch.
  static int pdl_env = -1;
  if (pdl_env < 0) {
    const char* e = getenv("BHMC_PDL_MLP");
    pdl_env = e ? atoi(e) : 1;
  }
  return pdl_env != 0;
}

static int bg_launch(bhmc_ctx* ctx, const BgLaunch& L) {
  const bool pair = L.pair;
  const size_t smem = L.smem;
  static size_t configured = 0, configured2 = 0;
  if (!pair && smem > configured) {
    BHMC_CUDA_OK(cudaFuncSetAttribute(k_tc_bgemm<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BHMC_CUDA_OK(cudaFuncSetAttribute(k_tc_bgemm<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  if (pair && smem > configured2) {
    BHMC_CUDA_OK(cudaFuncSetAttribute(k_tc_bgemm2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BHMC_CUDA_OK(cudaFuncSetAttribute(k_tc_bgemm2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured2 = smem;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)L.grid);
  cfg.blockDim = dim3(NON_EPI_THREADS + 32 * EW);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pair) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_mlp_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  // the instantiation with the fused sampler update carries Philox / Box-Muller code in its epilogue: only where asked for
  if (pair) BHMC_CUDA_OK(L.upd ? cudaLaunchKernelEx(&cfg, k_tc_bgemm2<true>, L.m[0], L.m[1], L.m[2], L.m[3], L.p)
                               : cudaLaunchKernelEx(&cfg, k_tc_bgemm2<false>, L.m[0], L.m[1], L.m[2], L.m[3], L.p));
  else BHMC_CUDA_OK(L.upd ? cudaLaunchKernelEx(&cfg, k_tc_bgemm<true>, L.m[0], L.m[1], L.m[2], L.m[3], L.p)
                          : cudaLaunchKernelEx(&cfg, k_tc_bgemm<false>, L.m[0], L.m[1], L.m[2], L.m[3], L.p));
  ctx->launches++;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

// C[z] = epilogue(A[z] . B[z]) for z < batch, operands described by d (fp32, arbitrary strides).
int tc_bgemm(bhmc_ctx* ctx, const GemmDesc& d, int batch, bool split3) {
  BgLaunch L;
  BHMC_TRY(bg_prepare(ctx, d, batch, split3, &L));
  return bg_launch(ctx, L);
}

// Two GEMMs that do not depend on each other: one grouped launch when their tile geometry agrees (k_tc_bgemm_grp), else two.
int tc_bgemm_two(bhmc_ctx* ctx, const GemmDesc& d0, const GemmDesc& d1, int batch, bool split3) {
  // Measured at cfg4 (16 chains): the grouped launch takes 45.4 us where the two launches take 22.2 + 24.5 us, and the step
  // gets 2.4 % SLOWER (62.8 vs 64.4 k grad-evals/s): with 3-4 tiles per CTA the epilogues do run under the next main loop,
  // but they contend with it (global loads / stores and the shared-memory port) instead of hiding behind it.  Off by default.
  static int grp_env = -1;  // BHMC_BG_GROUP=1: one grouped launch
  if (grp_env < 0) {
    const char* e = getenv("BHMC_BG_GROUP");
    grp_env = e ? atoi(e) : 0;
  }
  BgLaunch L0, L1;
  BHMC_TRY(bg_prepare(ctx, d0, batch, split3, &L0));
  {
    // the second preparation must not overwrite an operand copy of the first in the shared scratch slots
    const bool a1_scratch = !(d1.a_hi && (!split3 || d1.a_lo)), b1_scratch = !(d1.b_hi && (!split3 || d1.b_lo));
    if (!grp_env || (L0.used_a && a1_scratch) || (L0.used_b && b1_scratch)) {
      BHMC_TRY(bg_launch(ctx, L0));
      return tc_bgemm(ctx, d1, batch, split3);
    }
  }
  BHMC_TRY(bg_prepare(ctx, d1, batch, split3, &L1));
  if ((L0.used_a && L1.used_a) || (L0.used_b && L1.used_b)) {
    set_error("grouped GEMM launch: both problems claimed the same operand scratch slot");
    return BHMC_ERR_STATE;
  }
  const bool ok = grp_env && !L0.pair && !L1.pair && !L0.upd && !L1.upd && L0.p.epi2 && L1.p.epi2 && L0.p.BN == L1.p.BN &&
                  L0.p.stages == L1.p.stages && L0.p.split3 == L1.p.split3 && L0.smem == L1.smem;
  if (!ok) {
    BHMC_TRY(bg_launch(ctx, L0));
    return bg_launch(ctx, L1);
  }
  static BgGroup g;  // 3 KB: filled per call, passed by value to the launch
  for (int i = 0; i < 4; ++i) g.m[0][i] = L0.m[i], g.m[1][i] = L1.m[i];
  g.p[0] = L0.p, g.p[1] = L1.p;
  g.work0 = L0.work, g.work = L0.work + L1.work;
  static size_t configured = 0;
  if (L0.smem > configured) {
    BHMC_CUDA_OK(cudaFuncSetAttribute(k_tc_bgemm_grp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L0.smem));
    configured = L0.smem;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)std::min(g.work, ctx->sm_count));
  cfg.blockDim = dim3(NON_EPI_THREADS + 32 * EW);
  cfg.dynamicSmemBytes = L0.smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (pdl_mlp_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  BHMC_CUDA_OK(cudaLaunchKernelEx(&cfg, k_tc_bgemm_grp, g));
  ctx->launches++;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

}  // namespace bhmc
