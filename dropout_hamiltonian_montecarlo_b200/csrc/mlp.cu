// Dropout MLP (n_in -> n_mid -> n_mid -> n_out) log-posterior gradient for a batch of chains with
// chain-private weights.  Reference: hamiltonian/models/gpu/mlp.py
//   :19-31  MyNetwork: h = relu(dropout(l1 x)); h = relu(dropout(l2 h)); y = l3(dropout(h)), ratio 0.1,
//           Chainer L.Linear stores W as (out, in); Chainer dropout is inverted (x * mask / (1-ratio))
//   :47-64  grad = d(mean softmax-CE)/d theta + alpha/2 * theta
//   :66-82  log_likelihood = the (positive) mean CE loss; NLP = loss + log_prior,
//   :40-45  log_prior = -alpha/2 sum_v |theta_v|^2 / dim_v
// Parameter row (Chainer namedparams order): /l1/W [mid,in] | /l1/b | /l2/W [mid,mid] | /l2/b | /l3/W [out,mid] | /l3/b.
//
// The eight GEMMs per evaluation are strided-batched over chains with fused epilogues (bias + dropout + ReLU,
// ReLU-gate of the back-propagated signal, alpha/2*W).  BHMC_PREC_FP32 runs them as fp32 FMA on CUDA cores
// (k_mlp_gemm below); BHMC_PREC_BF16X3 / BF16 send the five large ones to tcgen05 (tc_bgemm.cu).
// Because relu(a*m) > 0 implies the keep-mask m == 1, the backward pass needs only the stored activations
// (H1, H2d), not the masks:  dA2 = dH2d * [H2d>0] / keep^2,  dA1 = dH1 * [H1>0] / keep.
#include <cuda_bf16.h>

#include <new>

#include <stdlib.h>

#include <algorithm>

#include "internal.cuh"
#include "philox.cuh"

namespace bhmc {

static constexpr int TM = 64, TN = 64, TK = 16;

#include "mlp_common.cuh"

// tc_bgemm.cu: the same GEMM + epilogue on the tensor cores (bf16 hi/lo split or single bf16 pass)
int tc_bgemm(bhmc_ctx* ctx, const GemmDesc& d, int batch, bool split3);
int tc_bgemm_two(bhmc_ctx* ctx, const GemmDesc& d0, const GemmDesc& d1, int batch, bool split3);
int tc_split_rows(bhmc_ctx* ctx, const float* src, int64_t sb, int64_t rs, int64_t cs, int R, int K, int64_t Kp, int Z,
                  __nv_bfloat16* hi, __nv_bfloat16* lo);

__global__ void __launch_bounds__(256) k_mlp_gemm(GemmDesc d) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int c = blockIdx.z, t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const float* A = d.A + (int64_t)c * d.a_batch;
  const float* B = d.B + (int64_t)c * d.b_batch;
  const bool a_kfast = d.a_cs == 1, b_nfast = d.b_cs == 1;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < d.K; k0 += TK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int m, k;
      if (a_kfast) m = t >> 2, k = (t & 3) * 4 + e;
      else k = t >> 4, m = (t & 15) * 4 + e;
      float v = 0.f;
      if (m0 + m < d.M && k0 + k < d.K) v = A[(int64_t)(m0 + m) * d.a_rs + (int64_t)(k0 + k) * d.a_cs];
      As[k][m] = v;
      int n, kb;
      if (b_nfast) kb = t >> 4, n = (t & 15) * 4 + e;
      else n = t >> 2, kb = (t & 3) * 4 + e;
      float w = 0.f;
      if (n0 + n < d.N && k0 + kb < d.K) w = B[(int64_t)(k0 + kb) * d.b_rs + (int64_t)(n0 + n) * d.b_cs];
      Bs[kb][n] = w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) a[e] = As[kk][ty * 4 + e], b[e] = Bs[kk][tx * 4 + e];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int nb = n0 + tx * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= d.M) continue;
    uint32_t ka = 15u, kb = 15u;
    if (d.epi >= 1 && nb < d.N) {
      if (!d.mask_a || (d.epi == 2 && !d.mask_b)) {
        const uint32_t k8 = keep_bits8(d, c, m, nb);
        ka = k8 & 15u, kb = k8 >> 4;
      }
      if (d.mask_a) {
        ka = 0;
        for (int j = 0; j < 4; ++j)
          if (nb + j < d.N && d.mask_a[(int64_t)c * d.mask_batch + (int64_t)m * d.N + nb + j]) ka |= 1u << j;
      }
      if (d.epi == 2 && d.mask_b) {
        kb = 0;
        for (int j = 0; j < 4; ++j)
          if (nb + j < d.N && d.mask_b[(int64_t)c * d.mask_batch + (int64_t)m * d.N + nb + j]) kb |= 1u << j;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = nb + j;
      if (n >= d.N) continue;
      float v = acc[i][j];
      if (d.bias) v += d.bias[(int64_t)c * d.bias_batch + n];
      if (d.addsrc) v += d.add_scale * d.addsrc[(int64_t)c * d.add_batch + (int64_t)m * d.add_rs + n];
      if (d.gate) v = (d.gate[(int64_t)c * d.gate_batch + (int64_t)m * d.gate_rs + n] > 0.f) ? v * d.gate_scale : 0.f;
      if (d.epi >= 1) {
        v = ((ka >> j) & 1u) ? v * d.keep_inv : 0.f;
        v = fmaxf(v, 0.f);
        if (d.epi == 2) v = ((kb >> j) & 1u) ? v * d.keep_inv : 0.f;
      }
      d.C[(int64_t)c * d.c_batch + (int64_t)m * d.c_rs + n] = v;
      if (d.ck_hi) split_store_pair(v, d.ck_hi, d.ck_lo, (int64_t)c * d.ck_batch + (int64_t)m * d.ck_ld + n);
      if (d.ct_hi) split_store_pair(v, d.ct_hi, d.ct_lo, (int64_t)c * d.ct_batch + (int64_t)n * d.ct_ld + m);
    }
  }
}

// ---- skinny GEMMs of the n_out-wide output layer (N, M or K = n_out ~ 10) -------------------------------------------
// The 64x64 tile kernel above wastes 84 % of a tile on them (ncu: 46 / 51 / 25 us at 16 chains); each gets a kernel
// shaped like its own data movement.  Plain-epilogue subsets only (what mlp.cu needs): run_gemm() falls back to
// k_mlp_gemm for anything else.
static constexpr int SKINNY = 16;

// N <= 16, A rows contiguous in k (a_cs == 1), B(k, n) contiguous in k (b_rs == 1); + bias.  Block = 32 rows (8 warps
// x 4 rows); lanes split k (coalesced 128-byte loads of the A rows and of the B columns, which stay in L1 -- a few KB
// per chain), butterfly reduction, lane n stores column n.  No staging phase: every load of the k loop is independent.
__global__ void __launch_bounds__(256) k_mlp_gemm_small_n(GemmDesc d) {
  const int c = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* A = d.A + (int64_t)c * d.a_batch;
  const float* B = d.B + (int64_t)c * d.b_batch;
  const int m0 = (blockIdx.x * 8 + warp) * 4;
  float acc[4][SKINNY];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int n = 0; n < SKINNY; ++n) acc[r][n] = 0.f;
#pragma unroll 2
  for (int k = lane; k < d.K; k += 32) {
    float a[4], b[SKINNY];
#pragma unroll
    for (int r = 0; r < 4; ++r) a[r] = (m0 + r < d.M) ? __ldg(A + (int64_t)(m0 + r) * d.a_rs + k) : 0.f;
#pragma unroll
    for (int n = 0; n < SKINNY; ++n) b[n] = (n < d.N) ? __ldg(B + (int64_t)n * d.b_cs + k) : 0.f;
#pragma unroll
    for (int n = 0; n < SKINNY; ++n)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r][n] = fmaf(a[r], b[n], acc[r][n]);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float mine = 0.f;
#pragma unroll
    for (int n = 0; n < SKINNY; ++n) {
      float v = acc[r][n];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == n) mine = v;
    }
    if (m0 + r < d.M && lane < d.N) {
      if (d.bias) mine += d.bias[(int64_t)c * d.bias_batch + lane];
      d.C[(int64_t)c * d.c_batch + (int64_t)(m0 + r) * d.c_rs + lane] = mine;
    }
  }
}

// M <= 16, B(k, n) contiguous in n (b_cs == 1); + add_scale * addsrc.  Block = 32 columns x 16 k-groups; A (K x M, a few
// KB) is staged in shared memory as [k][m]; every thread keeps the M partial sums of its column over every 16th k, the
// groups are combined in a fixed order through the same shared memory.
static constexpr int SM_GROUPS = 16;
__global__ void __launch_bounds__(32 * SM_GROUPS) k_mlp_gemm_small_m(GemmDesc d) {
  extern __shared__ float sm_buf[];  // max(K*M, SM_GROUPS*M*33) floats
  const int c = blockIdx.y, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  const float* A = d.A + (int64_t)c * d.a_batch;
  const float* B = d.B + (int64_t)c * d.b_batch;
  for (int i = threadIdx.x; i < d.K * d.M; i += 32 * SM_GROUPS) {
    int k, m;
    if (d.a_rs == 1) k = i / d.M, m = i - k * d.M;  // source is [k][m]-contiguous
    else m = i / d.K, k = i - m * d.K;
    sm_buf[k * d.M + m] = A[(int64_t)m * d.a_rs + (int64_t)k * d.a_cs];
  }
  __syncthreads();
  float acc[SKINNY];
#pragma unroll
  for (int m = 0; m < SKINNY; ++m) acc[m] = 0.f;
  if (n < d.N) {
#pragma unroll 4
    for (int k = ty; k < d.K; k += SM_GROUPS) {
      const float b = B[(int64_t)k * d.b_rs + n];
      const float* a = sm_buf + k * d.M;
#pragma unroll
      for (int m = 0; m < SKINNY; ++m)
        if (m < d.M) acc[m] = fmaf(a[m], b, acc[m]);
    }
  }
  __syncthreads();  // A is no longer needed: the buffer now holds the partial sums [group][m][33]
#pragma unroll
  for (int m = 0; m < SKINNY; ++m)
    if (m < d.M) sm_buf[(ty * d.M + m) * 33 + tx] = acc[m];
  __syncthreads();
  if (n < d.N) {
    for (int m = ty; m < d.M; m += SM_GROUPS) {
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < SM_GROUPS; ++i) v += sm_buf[(i * d.M + m) * 33 + tx];
      if (d.addsrc) v = fmaf(d.add_scale, d.addsrc[(int64_t)c * d.add_batch + (int64_t)m * d.add_rs + n], v);
      d.C[(int64_t)c * d.c_batch + (int64_t)m * d.c_rs + n] = v;
    }
  }
}

// K <= 16, B(k, n) contiguous in n (b_cs == 1), N % 4 == 0, 16-byte aligned rows; * gate.  HBM-bound: a thread produces
// four consecutive columns of four rows (K float4 loads of B from L1/L2 shared by the rows, 4K broadcast loads of A,
// four float4 stores).
__global__ void __launch_bounds__(128) k_mlp_gemm_small_k(GemmDesc d) {
  const int c = blockIdx.z, m0 = blockIdx.y * 4;
  const int n = (blockIdx.x * 128 + threadIdx.x) * 4;
  if (n >= d.N) return;
  const float* A = d.A + (int64_t)c * d.a_batch;
  const float* B = d.B + (int64_t)c * d.b_batch + n;
  float4 acc[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < SKINNY; ++k) {
    if (k < d.K) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(B + (int64_t)k * d.b_rs));
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float a = (m0 + r < d.M) ? __ldg(A + (int64_t)(m0 + r) * d.a_rs + (int64_t)k * d.a_cs) : 0.f;
        acc[r].x = fmaf(a, b.x, acc[r].x), acc[r].y = fmaf(a, b.y, acc[r].y);
        acc[r].z = fmaf(a, b.z, acc[r].z), acc[r].w = fmaf(a, b.w, acc[r].w);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int m = m0 + r;
    if (m >= d.M) break;
    float4 v = acc[r];
    if (d.gate) {
      const float4 gt = __ldg(reinterpret_cast<const float4*>(d.gate + (int64_t)c * d.gate_batch + (int64_t)m * d.gate_rs + n));
      v.x = gt.x > 0.f ? v.x * d.gate_scale : 0.f;
      v.y = gt.y > 0.f ? v.y * d.gate_scale : 0.f;
      v.z = gt.z > 0.f ? v.z * d.gate_scale : 0.f;
      v.w = gt.w > 0.f ? v.w * d.gate_scale : 0.f;
    }
    *reinterpret_cast<float4*>(d.C + (int64_t)c * d.c_batch + (int64_t)m * d.c_rs + n) = v;
    if (d.ck_hi) {  // K-major operand copy of the row: four consecutive bf16 = one 8-byte store (ck_ld % 4 == 0, n % 4 == 0)
      const float ve[4] = {v.x, v.y, v.z, v.w};
      __nv_bfloat16 hb[4], lb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        hb[j] = __float2bfloat16_rn(ve[j]);
        lb[j] = __float2bfloat16_rn(ve[j] - __bfloat162float(hb[j]));
      }
      const int64_t o = (int64_t)c * d.ck_batch + (int64_t)m * d.ck_ld + n;
      *reinterpret_cast<uint2*>(d.ck_hi + o) = *reinterpret_cast<const uint2*>(hb);
      if (d.ck_lo) *reinterpret_cast<uint2*>(d.ck_lo + o) = *reinterpret_cast<const uint2*>(lb);
    }
    // (no transposed copy here: this kernel walks rows, a column-major write would be 2 bytes per 32-byte sector)
  }
}

static int run_gemm(bhmc_ctx* ctx, const GemmDesc& d, int C) {
  auto al16 = [](const void* q) { return ((uintptr_t)q & 15u) == 0; };
  const bool plain = d.epi == 0 && !d.mask_a && !d.mask_b;
  static int skinny_env = -1;  // BHMC_MLP_SKINNY=0: generic tile kernel for every shape (A/B measurements)
  if (skinny_env < 0) {
    const char* e = getenv("BHMC_MLP_SKINNY");
    skinny_env = e ? atoi(e) : 1;
  }
  if (skinny_env && plain && d.N <= SKINNY && d.M > SKINNY && d.a_cs == 1 && d.b_rs == 1 && !d.addsrc && !d.gate) {
    k_mlp_gemm_small_n<<<dim3((unsigned)ceil_div(d.M, 32), C), 256, 0, ctx->stream>>>(d);
  } else if (skinny_env && plain && d.M <= SKINNY && d.N > SKINNY && d.b_cs == 1 && !d.bias && !d.gate &&
             (size_t)d.K * d.M * sizeof(float) <= 44 * 1024) {
    const size_t sm = sizeof(float) * std::max((size_t)d.K * d.M, (size_t)SM_GROUPS * d.M * 33);
    k_mlp_gemm_small_m<<<dim3((unsigned)ceil_div(d.N, 32), C), 32 * SM_GROUPS, sm, ctx->stream>>>(d);
  } else if (skinny_env && plain && d.K <= SKINNY && d.M <= 4 * 65535 && d.b_cs == 1 && d.N % 4 == 0 && !d.bias && !d.addsrc && al16(d.B) &&
             d.b_batch % 4 == 0 && d.b_rs % 4 == 0 && al16(d.C) && d.c_batch % 4 == 0 && d.c_rs % 4 == 0 &&
             (!d.gate || (al16(d.gate) && d.gate_batch % 4 == 0 && d.gate_rs % 4 == 0)) && !d.ct_hi &&
             (!d.ck_hi || (d.ck_ld % 4 == 0 && d.ck_batch % 4 == 0))) {
    k_mlp_gemm_small_k<<<dim3((unsigned)ceil_div(d.N, 512), (unsigned)ceil_div(d.M, 4), C), 128, 0, ctx->stream>>>(d);
  } else {
    dim3 grid((unsigned)ceil_div(d.M, TM), (unsigned)ceil_div(d.N, TN), (unsigned)C);
    k_mlp_gemm<<<grid, 256, 0, ctx->stream>>>(d);
  }
  ctx->launches++;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

// softmax cross-entropy over n_out logits: Z <- (softmax(Z) - onehot)/B ; loss[c] += (lse - z_y)/B   (mlp.py:57)
__global__ void __launch_bounds__(256) k_mlp_loss(float* __restrict__ Z, int B, int n_out, const int32_t* __restrict__ y,
                                                  double* __restrict__ loss, int write_grad) {
  const int c = blockIdx.y;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  double my = 0.0;
  if (r < B) {
    float* z = Z + ((int64_t)c * B + r) * n_out;
    float m = -INFINITY;
    for (int k = 0; k < n_out; ++k) m = fmaxf(m, z[k]);
    float s = 0.f;
    for (int k = 0; k < n_out; ++k) s += expf(z[k] - m);
    const int yy = y[r];
    my = ((double)m + (double)logf(s) - (double)z[yy]) / (double)B;
    if (write_grad) {
      const float inv = 1.0f / s, ib = 1.0f / (float)B;
      for (int k = 0; k < n_out; ++k) z[k] = (expf(z[k] - m) * inv - (k == yy ? 1.f : 0.f)) * ib;
    }
  }
  __shared__ double sm[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) my += __shfl_xor_sync(0xffffffffu, my, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = my;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < 8; ++w) tot += sm[w];
    atomicAdd(loss + c, quantize_addend<40>(tot));  // mean CE (< 8192): order-independent, see internal.cuh
  }
}

// bias gradient: out[c, n] = sum_b S[c, b, n] + alpha/2 * bias[c, n]
// block = 32 columns x 32 row groups of one chain (a thread per column walking all B rows left 64 blocks on the GPU and
// took 40 us: one DRAM latency per row); each thread sums every 32nd row with four independent accumulators, the
// groups are combined in a fixed order through shared memory (deterministic)
struct UpdOpt {  // optional sampler update applied where a gradient value is produced (GemmDesc::upd, mlp_common.cuh)
  int on;
  UpdateArgs u;
};
__global__ void __launch_bounds__(1024) k_mlp_colsum(const float* __restrict__ S, int B, int N, int64_t s_batch,
                                                     const float* __restrict__ q, int64_t ld, int64_t b_off,
                                                     float half_alpha, float* __restrict__ g, const UpdOpt uo) {
  __shared__ float part[32][33];
  const int c = blockIdx.y;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (n < N) {
    const float* s = S + (int64_t)c * s_batch + n;
    int b = ty;
    for (; b + 96 < B; b += 128) {
      a0 += s[(int64_t)b * N];
      a1 += s[(int64_t)(b + 32) * N];
      a2 += s[(int64_t)(b + 64) * N];
      a3 += s[(int64_t)(b + 96) * N];
    }
    for (; b < B; b += 32) a0 += s[(int64_t)b * N];
  }
  part[ty][tx] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (ty == 0 && n < N) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) acc += part[i][tx];
    const float gv = acc + half_alpha * q[(int64_t)c * ld + b_off + n];
    if (uo.on) upd_apply1(uo.u, c, b_off + n, gv);
    else g[(int64_t)c * ld + b_off + n] = gv;
  }
}


// ---------------------------------------------------------------------------------------------
// The n_out-wide head of the network in one pass over H2d (round 2).  Until now five launches read the same 16 MB:
// Z = H2d W3^T + b3 (k_mlp_gemm_small_n), loss / dZ (k_mlp_loss), gW3 = dZ^T H2d (k_mlp_gemm_small_m), gb3
// (k_mlp_colsum), dA2 = dZ W3 * gate (k_mlp_gemm_small_k), plus the transposed split of dA2 and the column sums of dA2
// for gb2 -- 68 + 8 + 7 us of a 330 us evaluation (ncu launch list).  A block owns 32 rows of one chain: the rows of H2d
// and W3 sit in shared memory, logits / softmax / dZ are computed once, and everything that depends on them leaves the
// block: the loss, dA2 as bf16 hi/lo in BOTH orientations (the operands of the dA1 and gW2 GEMMs; dA2 itself is never
// needed in fp32), and per-block partial sums of gW3, gb3, gb2 that k_mlp_head_reduce adds in block order (deterministic)
// together with alpha/2 * theta.  mlp.py:28-31,57-63 (F.softmax_cross_entropy mean over the batch, Linear backward).
template <int NO>  // classes padded to a multiple of 4, <= 16
__global__ void __launch_bounds__(256, 2) k_mlp_head(const float* __restrict__ H2d, int64_t act, int B, int n_mid, int n_out,
                                                  const float* __restrict__ q, int64_t ld, int64_t oW3, int64_t ob3,
                                                  const int32_t* __restrict__ y, float gate_scale, double* __restrict__ loss,
                                                  __nv_bfloat16* __restrict__ k_hi, __nv_bfloat16* __restrict__ k_lo, int64_t k_batch,
                                                  int64_t k_ld, __nv_bfloat16* __restrict__ t_hi, __nv_bfloat16* __restrict__ t_lo,
                                                  int64_t t_batch, int64_t t_ld, float* __restrict__ part, int n_rb) {
  extern __shared__ __align__(16) float sm_head[];
  const int pitch = n_mid + 4;              // floats; + 16 bytes: a column of 32 rows spreads over the banks
  float* Hs = sm_head;                      // [32][pitch]
  float* Ws = Hs + 32 * pitch;              // [NO][n_mid], rows >= n_out zero
  float* dZs = Ws + NO * n_mid;             // [32][NO]
  __shared__ double red[8];
  const int c = blockIdx.y, rb = blockIdx.x, m0 = rb * 32;
  const int rows = min(32, B - m0);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int nq = n_mid >> 2;  // quads per row
  const float* Hg = H2d + (int64_t)c * act + (int64_t)m0 * n_mid;
  const float* W3 = q + (int64_t)c * ld + oW3;
  for (int i = t; i < 32 * nq; i += 256) {
    const int r = i / nq, k4 = i - r * nq;
    const float4 v = r < rows ? __ldg(reinterpret_cast<const float4*>(Hg + (int64_t)r * n_mid) + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(Hs + r * pitch + 4 * k4) = v;
  }
  for (int i = t; i < NO * nq; i += 256) {
    const int o = i / nq, k4 = i - o * nq;
    // W3 rows start at oW3 + o*n_mid: 16-byte aligned when oW3 % 4 == 0 (checked by the host)
    const float4 v = o < n_out ? __ldg(reinterpret_cast<const float4*>(W3 + (int64_t)o * n_mid) + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(Ws + o * n_mid + 4 * k4) = v;
  }
  __syncthreads();
  // ---- logits of rows 4*warp .. +3 (lanes split the contraction), softmax, loss, dZ ----
  {
    float acc[4][NO];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int o = 0; o < NO; ++o) acc[r][o] = 0.f;
    for (int k4 = lane; k4 < nq; k4 += 32) {
      float4 h[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) h[r] = *reinterpret_cast<const float4*>(Hs + (4 * warp + r) * pitch + 4 * k4);
#pragma unroll
      for (int o = 0; o < NO; ++o) {
        const float4 w = *reinterpret_cast<const float4*>(Ws + o * n_mid + 4 * k4);
#pragma unroll
        for (int r = 0; r < 4; ++r)
          acc[r][o] = fmaf(h[r].x, w.x, fmaf(h[r].y, w.y, fmaf(h[r].z, w.z, fmaf(h[r].w, w.w, acc[r][o]))));
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int o = 0; o < NO; ++o)
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) acc[r][o] += __shfl_xor_sync(0xffffffffu, acc[r][o], s);
    double my = 0.0;
    if (lane < 4) {  // lane r finishes row 4*warp + r
      const int r = 4 * warp + lane;
      float z[NO];
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
        if (rr == lane) {
#pragma unroll
          for (int o = 0; o < NO; ++o) z[o] = acc[rr][o];
        }
      if (r < rows) {
        const float* b3 = q + (int64_t)c * ld + ob3;
        float mx = -INFINITY;
#pragma unroll
        for (int o = 0; o < NO; ++o)
          if (o < n_out) z[o] += __ldg(b3 + o), mx = fmaxf(mx, z[o]);
        float ssum = 0.f;
#pragma unroll
        for (int o = 0; o < NO; ++o)
          if (o < n_out) ssum += expf(z[o] - mx);
        const int yy = y[m0 + r];
        float zy = 0.f;
#pragma unroll
        for (int o = 0; o < NO; ++o)
          if (o == yy) zy = z[o];
        my = ((double)mx + (double)logf(ssum) - (double)zy) / (double)B;
        const float inv = 1.0f / ssum, ib = 1.0f / (float)B;
#pragma unroll
        for (int o = 0; o < NO; ++o) dZs[r * NO + o] = o < n_out ? (expf(z[o] - mx) * inv - (o == yy ? 1.f : 0.f)) * ib : 0.f;
      } else {
#pragma unroll
        for (int o = 0; o < NO; ++o) dZs[r * NO + o] = 0.f;
      }
    }
    my += __shfl_xor_sync(0xffffffffu, my, 1);
    my += __shfl_xor_sync(0xffffffffu, my, 2);
    if (lane == 0) red[warp] = my;
  }
  __syncthreads();
  if (t == 0) {
    double tot = 0.0;
    for (int w = 0; w < 8; ++w) tot += red[w];
    atomicAdd(loss + c, quantize_addend<40>(tot));  // mean CE: order-independent, see internal.cuh
  }
  float* pW = part + ((int64_t)c * n_rb + rb) * ((int64_t)(NO + 2) * n_mid + NO);  // [NO][n_mid] gW3, [2][n_mid] gb2 (row halves), [NO] gb3
  if (t < NO) {
    float sacc = 0.f;
    for (int r = 0; r < 32; ++r) sacc += dZs[r * NO + t];
    pW[(int64_t)(NO + 2) * n_mid + t] = sacc;
  }
  // ---- gW3 partial = dZ^T H2d of the block's rows (thread = column quad x half of the classes); must read H before the
  //      next phase overwrites it ----
  const int half = t >> 7;  // classes half*NO/2 .. +NO/2 here; rows half, half + 2, ... below
  for (int k4 = t & 127; k4 < nq; k4 += 128) {
    float4 gacc[NO / 2];
#pragma unroll
    for (int j = 0; j < NO / 2; ++j) gacc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < rows; ++r) {
      const float4 h = *reinterpret_cast<const float4*>(Hs + r * pitch + 4 * k4);
#pragma unroll
      for (int j = 0; j < NO / 2; ++j) {
        const float dz = dZs[r * NO + half * (NO / 2) + j];
        gacc[j].x = fmaf(dz, h.x, gacc[j].x), gacc[j].y = fmaf(dz, h.y, gacc[j].y);
        gacc[j].z = fmaf(dz, h.z, gacc[j].z), gacc[j].w = fmaf(dz, h.w, gacc[j].w);
      }
    }
#pragma unroll
    for (int j = 0; j < NO / 2; ++j) *reinterpret_cast<float4*>(pW + (int64_t)(half * (NO / 2) + j) * n_mid + 4 * k4) = gacc[j];
  }
  __syncthreads();
  // ---- dA2 = (dZ W3) * [H2d > 0] * gate_scale, thread = column quad: K-major bf16 copies (a warp stores 256 contiguous
  //      bytes per row), column sums (gb2).  The 16 bytes of H a thread has just read are replaced by the bf16 hi | lo
  //      quad of dA2, for the transposed copies below ----
  for (int k4 = t & 127; k4 < nq; k4 += 128) {
    float4 w[NO];
#pragma unroll
    for (int o = 0; o < NO; ++o) w[o] = *reinterpret_cast<const float4*>(Ws + o * n_mid + 4 * k4);
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = half; r < rows; r += 2) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int o4 = 0; o4 < NO / 4; ++o4) {
        const float4 dz4 = *reinterpret_cast<const float4*>(dZs + r * NO + 4 * o4);  // broadcast
        const float dze[4] = {dz4.x, dz4.y, dz4.z, dz4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int o = 4 * o4 + j;
          v.x = fmaf(dze[j], w[o].x, v.x), v.y = fmaf(dze[j], w[o].y, v.y), v.z = fmaf(dze[j], w[o].z, v.z), v.w = fmaf(dze[j], w[o].w, v.w);
        }
      }
      float4* hslot = reinterpret_cast<float4*>(Hs + r * pitch + 4 * k4);
      const float4 h = *hslot;
      v.x = h.x > 0.f ? v.x * gate_scale : 0.f, v.y = h.y > 0.f ? v.y * gate_scale : 0.f;
      v.z = h.z > 0.f ? v.z * gate_scale : 0.f, v.w = h.w > 0.f ? v.w * gate_scale : 0.f;
      cs.x += v.x, cs.y += v.y, cs.z += v.z, cs.w += v.w;
      const __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y), h2 = __float2bfloat16_rn(v.z),
                          h3 = __float2bfloat16_rn(v.w);
      const int64_t o = (int64_t)c * k_batch + (int64_t)(m0 + r) * k_ld + 4 * k4;
      __nv_bfloat162 a = __halves2bfloat162(h0, h1), b = __halves2bfloat162(h2, h3);
      uint4 both;
      both.x = *reinterpret_cast<uint32_t*>(&a), both.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(k_hi + o) = make_uint2(both.x, both.y);
      a = __halves2bfloat162(__float2bfloat16_rn(v.x - __bfloat162float(h0)), __float2bfloat16_rn(v.y - __bfloat162float(h1)));
      b = __halves2bfloat162(__float2bfloat16_rn(v.z - __bfloat162float(h2)), __float2bfloat16_rn(v.w - __bfloat162float(h3)));
      both.z = *reinterpret_cast<uint32_t*>(&a), both.w = *reinterpret_cast<uint32_t*>(&b);
      if (k_lo) *reinterpret_cast<uint2*>(k_lo + o) = make_uint2(both.z, both.w);
      *reinterpret_cast<uint4*>(hslot) = both;  // hi quad | lo quad
    }
    *reinterpret_cast<float4*>(pW + (int64_t)(NO + half) * n_mid + 4 * k4) = cs;  // gb2: one slot per row half
  }
  __syncthreads();
  // ---- transposed bf16 copies of dA2 (operand of gW2 = dA2^T H1): lane = row, so the 32 rows of a column are 64
  //      contiguous bytes; the values come back from the H block (row pitch + 16 bytes: 4 wavefronts per 16-byte load) ----
  if (lane < rows) {
    for (int k4 = warp; k4 < nq; k4 += 8) {
      const uint4 both = *reinterpret_cast<const uint4*>(Hs + lane * pitch + 4 * k4);
      const int64_t o0 = (int64_t)c * t_batch + (int64_t)(4 * k4) * t_ld + m0 + lane;
      const uint32_t hw[2] = {both.x, both.y}, lw[2] = {both.z, both.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint16_t hb = (uint16_t)(hw[j >> 1] >> (16 * (j & 1))), lb = (uint16_t)(lw[j >> 1] >> (16 * (j & 1)));
        reinterpret_cast<uint16_t*>(t_hi)[o0 + (int64_t)j * t_ld] = hb;
        if (t_lo) reinterpret_cast<uint16_t*>(t_lo)[o0 + (int64_t)j * t_ld] = lb;
      }
    }
  }
}

// g[W3] = sum_blocks partial + alpha/2 W3, likewise gb3 and gb2 (block order: deterministic)
template <int NO>
__global__ void __launch_bounds__(256) k_mlp_head_reduce(const float* __restrict__ part, int n_rb, int n_mid, int n_out,
                                                         const float* __restrict__ q, int64_t ld, int64_t oW3, int64_t ob3,
                                                         int64_t ob2, float half_alpha, float* __restrict__ g, const UpdOpt uo) {
  const int c = blockIdx.y;
  if (uo.on && blockIdx.x == 0 && threadIdx.x == 0 && uo.u.stat_new && uo.u.post_len > 0 && uo.u.it_post < uo.u.L[c] - 1)
    uo.u.stat_new[c] = uo.u.stat[c];  // k_hmc_update's latch of the evaluation's scalar (the head kernel has finished)
  const int64_t stride = (int64_t)(NO + 2) * n_mid + NO;
  const float* pc = part + (int64_t)c * n_rb * stride;
  const int i = blockIdx.x * 256 + threadIdx.x;
  const int nW = n_out * n_mid;
  if (i < nW) {
    float acc = 0.f;
    for (int b = 0; b < n_rb; ++b) acc += pc[b * stride + i];  // rows o < n_out of the [NO][n_mid] block are contiguous
    const float gv = acc + half_alpha * q[(int64_t)c * ld + oW3 + i];
    if (uo.on) upd_apply1(uo.u, c, oW3 + i, gv);
    else g[(int64_t)c * ld + oW3 + i] = gv;
  } else if (i < nW + n_mid) {
    const int n = i - nW;
    float acc = 0.f;
    for (int b = 0; b < n_rb; ++b) acc += pc[b * stride + (int64_t)NO * n_mid + n] + pc[b * stride + (int64_t)(NO + 1) * n_mid + n];
    const float gv = acc + half_alpha * q[(int64_t)c * ld + ob2 + n];
    if (uo.on) upd_apply1(uo.u, c, ob2 + n, gv);
    else g[(int64_t)c * ld + ob2 + n] = gv;
  } else if (i < nW + n_mid + n_out) {
    const int o = i - nW - n_mid;
    float acc = 0.f;
    for (int b = 0; b < n_rb; ++b) acc += pc[b * stride + (int64_t)(NO + 2) * n_mid + o];
    const float gv = acc + half_alpha * q[(int64_t)c * ld + ob3 + o];
    if (uo.on) upd_apply1(uo.u, c, ob3 + o, gv);
    else g[(int64_t)c * ld + ob3 + o] = gv;
  }
}

template <int NO>
static int launch_head(bhmc_ctx* ctx, const float* H2d, int64_t act, int B, int n_mid, int n_out, const float* q, int64_t ld,
                       int64_t oW3, int64_t ob3, int64_t ob2, const int32_t* y, float gate_scale, float half_alpha, double* loss,
                       __nv_bfloat16* k_hi, __nv_bfloat16* k_lo, int64_t k_batch, int64_t k_ld, __nv_bfloat16* t_hi,
                       __nv_bfloat16* t_lo, int64_t t_batch, int64_t t_ld, float* part, int C, float* g, const UpdOpt& uo) {
  const int n_rb = (int)ceil_div(B, 32);
  const size_t smem = sizeof(float) * ((size_t)32 * (n_mid + 4) + (size_t)NO * n_mid + 32 * NO);
  static size_t configured = 0;
  if (smem > configured) {
    BHMC_CUDA_OK(cudaFuncSetAttribute(k_mlp_head<NO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  k_mlp_head<NO><<<dim3((unsigned)n_rb, (unsigned)C), 256, smem, ctx->stream>>>(H2d, act, B, n_mid, n_out, q, ld, oW3, ob3, y, gate_scale,
                                                                             loss, k_hi, k_lo, k_batch, k_ld, t_hi, t_lo, t_batch,
                                                                             t_ld, part, n_rb);
  const int n_items = n_out * n_mid + n_mid + n_out;
  k_mlp_head_reduce<NO><<<dim3((unsigned)ceil_div(n_items, 256), (unsigned)C), 256, 0, ctx->stream>>>(part, n_rb, n_mid, n_out, q, ld, oW3,
                                                                                                   ob3, ob2, half_alpha, g, uo);
  ctx->launches += 2;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

struct MlpModel : ModelBase {
  int64_t N = 0;
  int n_in = 0, n_mid = 0, n_out = 0;
  float alpha = 0.f, ratio = 0.1f;
  const float* X = nullptr;
  const int32_t* labels = nullptr;
  float* X_owned = nullptr;
  int32_t* y_owned = nullptr;
  const uint8_t* masks = nullptr;  // injected [3][C][B][n_mid] or nullptr
  uint64_t seed = 0;
  int64_t chain_id0 = 0;
  uint32_t eval_id = 0;
  int64_t oW1, ob1, oW2, ob2, oW3, ob3;
  float* logits_sink = nullptr;  // predict: grad() stops after the forward pass and leaves the logits here
  // bf16 hi/lo operand copies of the current row window of X (K-major for X W1^T, transposed for dA1^T X): they do not
  // depend on q, and a sampler evaluates the same minibatch window many times in a row (every leapfrog iteration of an
  // sghmc.step) -- split once per window instead of once per evaluation (round 2; BHMC_MLP_XCACHE=0: per evaluation)
  // operand mirror: bf16 hi/lo image of the positions [C, mir_ld], filled by the sampler's update kernel (ModelBase::
  // operand_mirror); W1 and W2 are read from it as strided K-major views when it describes the q being evaluated
  __nv_bfloat16 *mir_hi = nullptr, *mir_lo = nullptr;
  int64_t mir_ld = 0;
  int mir_C = 0;
  const float* mir_fresh = nullptr;  // the buffer whose image the mirror holds, valid for the next grad() only
  // fused sampler update (grad_fused_update): the gradient-producing kernels apply it; W2^T of the NEW position is written
  // by the W2-gradient epilogue into w2t and serves the next evaluation's dA1 GEMM (same one-shot validity as the mirror)
  const UpdateArgs* fused_upd = nullptr;
  __nv_bfloat16* w2t = nullptr;
  size_t w2t_bytes = 0;
  const float* w2t_fresh = nullptr;
  __nv_bfloat16* xw = nullptr;
  size_t xw_bytes = 0;
  const float* xw_src = nullptr;
  int64_t xw_row0 = -1, xw_rows = -1;
  int xw_split3 = -1;

  ~MlpModel() override {
    cudaFree(X_owned);
    cudaFree(y_owned);
    cudaFree(xw);
    cudaFree(mir_hi);
    cudaFree(w2t);
  }
  int64_t default_rows() const override { return N; }

  bool operand_mirror(int C, int64_t ld, int prec, __nv_bfloat16** hi, __nv_bfloat16** lo, int64_t* mld) override {
    static int env = -1;  // BHMC_MLP_MIRROR=0: the weights are split by launches of their own in every evaluation
    if (env < 0) {
      const char* e = getenv("BHMC_MLP_MIRROR");
      env = e ? atoi(e) : 1;
    }
    // strided TMA views need 16-byte aligned rows: every weight matrix starts at a multiple of 8 elements and has rows of
    // a multiple of 8 elements
    if (!env || prec == BHMC_PREC_FP32 || n_in % 8 || n_mid % 8 || oW1 % 8 || oW2 % 8 || n_in < 64 || n_mid < 64) return false;
    const int64_t want_ld = round_up(ld, 8);
    if (C > mir_C || want_ld != mir_ld) {
      cudaFree(mir_hi);
      mir_hi = mir_lo = nullptr, mir_C = 0, mir_fresh = nullptr;
      if (cudaMalloc(&mir_hi, sizeof(__nv_bfloat16) * 2 * (size_t)C * want_ld) != cudaSuccess) {
        cudaGetLastError();
        return false;
      }
      mir_lo = mir_hi + (size_t)C * want_ld;
      mir_C = C, mir_ld = want_ld;
    }
    *hi = mir_hi, *lo = mir_lo, *mld = mir_ld;
    return true;
  }
  void mirror_written(const float* q) override { mir_fresh = q; }

  // Gradient at q followed by the sampler's update u (whole parameter vector: closing kick of iteration u.it_post, then --
  // if u.pre_len > 0 -- the drift of iteration u.it_pre), applied by the kernels that produce the gradient slices: no
  // update launch, no gradient round trip through HBM, and the bf16 operand copies of the new weights (mirror, W2^T) come
  // from the same epilogues.  BHMC_ERR_UNSUPPORTED before anything is launched = the caller runs grad() + update.
  int grad_fused_update(const float* q, int C, int64_t ld, int64_t row0, int64_t nrows, int prec, double* stat,
                        const UpdateArgs& u) override {
    // Measured (cfg4, 16 chains): 52.7 k grad-evals/s with it against 65 k without -- the GEMM epilogues are the
    // bottleneck of these launches (1.7 tiles per CTA, last tile exposed), and Philox / Box-Muller noise plus the p / q
    // traffic in there cost more than the HBM-bound update launch they replace.  Off by default; parity-tested both ways.
    static int env = -1;  // BHMC_MLP_FUSE_UPD=1: on
    if (env < 0) {
      const char* e = getenv("BHMC_MLP_FUSE_UPD");
      env = e ? atoi(e) : 0;
    }
    const bool whole = u.post_off == 0 && u.post_len >= P && (u.pre_len == 0 || (u.pre_off == 0 && u.pre_len >= P));
    if (!env || prec == BHMC_PREC_FP32 || !whole || u.q != q || !u.p || !u.L || !u.g || !u.mir_hi || u.mir_hi != mir_hi || logits_sink ||
        nrows < 64 || n_mid < 64 || n_in < 64 || n_out > 16 || n_mid % 8 || n_in % 8 || oW1 % 8 || oW2 % 8 || ld % 4)
      return BHMC_ERR_UNSUPPORTED;
    {
      const char* e1 = getenv("BHMC_MLP_FUSE");
      const char* e2 = getenv("BHMC_MLP_HEAD");
      const char* e3 = getenv("BHMC_BG_EPI2");
      if ((e1 && !atoi(e1)) || (e2 && !atoi(e2)) || (e3 && !atoi(e3))) return BHMC_ERR_UNSUPPORTED;
      const size_t head_smem = sizeof(float) * ((size_t)32 * (n_mid + 4) + (size_t)round_up(n_out, 4) * n_mid + 32 * round_up(n_out, 4));
      if (head_smem > (size_t)200 * 1024) return BHMC_ERR_UNSUPPORTED;
    }
    const int64_t kp_mid = round_up(n_mid, 64);
    const size_t need = sizeof(__nv_bfloat16) * 2 * (size_t)C * n_mid * kp_mid;
    if (need > w2t_bytes) {
      cudaFree(w2t);
      w2t = nullptr, w2t_bytes = 0, w2t_fresh = nullptr;
      if (cudaMalloc(&w2t, need) != cudaSuccess) {
        cudaGetLastError();
        return BHMC_ERR_UNSUPPORTED;
      }
      w2t_bytes = need;
    }
    fused_upd = &u;
    const int rc = grad(q, C, ld, row0, nrows, prec, const_cast<float*>(u.g), stat, 0);
    fused_upd = nullptr;
    if (rc == BHMC_OK) mir_fresh = q, w2t_fresh = q;  // both images describe the position the epilogues wrote
    return rc;
  }

  // NLP = loss + log_prior, log_prior = -alpha/2 sum_v |theta_v|^2/dim_v  (mlp.py:40-45,80-82)
  void energy_coeffs(int64_t, double* a, double* b, double* cv) const override {
    *a = 1.0;
    *b = 0.0;
    for (int v = 0; v < n_vars; ++v) cv[v] = -0.5 * (double)alpha / (double)var_len[v];
  }

  int grad(const float* q, int C, int64_t ld, int64_t row0, int64_t nrows, int prec, float* g, double* stat,
           uint32_t) override {
    BHMC_CHECK_ARG(X && (labels || logits_sink), "mlp model has no bound data");
    BHMC_CHECK_ARG(row0 >= 0 && nrows > 0 && row0 + nrows <= N, "row window outside the bound rows");
    const int B = (int)nrows;
    const float keep_inv = 1.0f / (1.0f - ratio), keep_prob = 1.0f - ratio;
    const int64_t act = (int64_t)B * n_mid;
    void* sc = nullptr;
    BHMC_TRY(ctx->get_scratch(7, sizeof(float) * (size_t)C * (4 * act + (size_t)B * n_out), &sc));
    float* H1 = (float*)sc;
    float* H2d = H1 + (size_t)C * act;
    float* dA2 = H2d + (size_t)C * act;
    float* dA1 = dA2 + (size_t)C * act;
    float* Z = dA1 + (size_t)C * act;
    const float* Xb = X + row0 * n_in;
    const uint32_t ev = eval_id++;
    const int64_t mstride = (int64_t)B * n_mid;  // per-chain stride inside one injected mask layer
    auto base = [&]() {
      GemmDesc d{};
      d.keep_inv = keep_inv;
      d.keep_prob = keep_prob;
      d.keep_thr16 = (uint32_t)lrint((double)keep_prob * 65536.0);
      d.seed = seed;
      d.chain_id0 = chain_id0;
      d.eval_id = ev;
      d.mask_batch = mstride;
      return d;
    };
    // the five large GEMMs go to the tensor cores unless the fp32 CUDA-core path is requested; the three GEMMs
    // that touch the n_out-wide logits (N, M or K = n_out ~ 10) stay on CUDA cores
    const bool use_tc = prec != BHMC_PREC_FP32, split3 = prec == BHMC_PREC_BF16X3;
    // Producer-written operand copies (GemmDesc::ck_* / ct_*, round 2): H1, dA2 and dA1 reach their consumer GEMMs as
    // bf16 hi/lo matrices written by the epilogue that computed them, in both orientations where two consumers
    // contract over different indices -- four of the eight large split launches of an evaluation disappear.  The
    // values are the split of the same fp32 numbers, so the result is bit-identical (BHMC_MLP_FUSE=0: separate splits).
    static int fuse_env = -1;
    if (fuse_env < 0) {
      const char* e = getenv("BHMC_MLP_FUSE");
      fuse_env = e ? atoi(e) : 1;
    }
    const bool fuse = fuse_env && use_tc && g && B >= 64 && n_mid >= 64 && n_in >= 64;
    const int64_t kp_mid = round_up(n_mid, 64), kp_b = round_up(B, 64);
    const int64_t e_k = (int64_t)B * kp_mid, e_t = (int64_t)n_mid * kp_b;  // elements per chain of a K-major / transposed copy
    __nv_bfloat16 *H1k = nullptr, *H1t = nullptr, *dA2k = nullptr, *dA2t = nullptr, *dA1t = nullptr;
    const int64_t lo_k = (int64_t)C * e_k, lo_t = (int64_t)C * e_t;  // offset of the lo copy inside a buffer
    // the n_out-wide head in one pass over H2d (k_mlp_head; BHMC_MLP_HEAD=0: the five separate launches)
    static int head_env = -1;
    if (head_env < 0) {
      const char* e = getenv("BHMC_MLP_HEAD");
      head_env = e ? atoi(e) : 1;
    }
    const int NO = (int)round_up(n_out, 4), n_rb = (int)ceil_div(B, 32);
    const size_t head_smem = sizeof(float) * ((size_t)32 * (n_mid + 4) + (size_t)NO * n_mid + 32 * NO);
    const bool head = head_env && fuse && !logits_sink && n_out <= 16 && n_mid % 4 == 0 && oW3 % 4 == 0 && ld % 4 == 0 &&
                      head_smem <= (size_t)200 * 1024;
    const size_t part_floats = head ? (size_t)C * n_rb * ((size_t)(NO + 2) * n_mid + NO) : 0;
    float* head_part = nullptr;
    if (fuse) {
      void* cb = nullptr;
      const size_t bf_bytes = sizeof(__nv_bfloat16) * 2 * (size_t)(2 * lo_k + 3 * lo_t);  // multiple of 256 (kp_* % 64 == 0)
      BHMC_TRY(ctx->get_scratch(15, bf_bytes + sizeof(float) * part_floats, &cb));
      H1k = (__nv_bfloat16*)cb;
      dA2k = H1k + 2 * lo_k;
      H1t = dA2k + 2 * lo_k;
      dA2t = H1t + 2 * lo_t;
      dA1t = dA2t + 2 * lo_t;
      head_part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(cb) + bf_bytes);
    }
    const int64_t kp_in = round_up(n_in, 64);
    const int64_t e_xk = (int64_t)B * kp_in, e_xt = (int64_t)n_in * kp_b;  // elements of the K-major / transposed copy of the window
    static int xcache_env = -1;
    if (xcache_env < 0) {
      const char* e = getenv("BHMC_MLP_XCACHE");
      xcache_env = e ? atoi(e) : 1;
    }
    const bool xcache = xcache_env && fuse;
    if (xcache) {
      const size_t need = sizeof(__nv_bfloat16) * 2 * (size_t)(e_xk + e_xt);
      if (need > xw_bytes) {
        cudaFree(xw);
        xw = nullptr, xw_bytes = 0, xw_row0 = -1;
        BHMC_CUDA_OK(cudaMalloc(&xw, need));
        xw_bytes = need;
      }
      if (xw_src != X || xw_row0 != row0 || xw_rows != nrows || xw_split3 != (int)split3) {
        BHMC_TRY(tc_split_rows(ctx, Xb, 0, n_in, 1, B, n_in, kp_in, 1, xw, split3 ? xw + e_xk : nullptr));
        BHMC_TRY(tc_split_rows(ctx, Xb, 0, 1, n_in, n_in, B, kp_b, 1, xw + 2 * e_xk, split3 ? xw + 2 * e_xk + e_xt : nullptr));
        xw_src = X, xw_row0 = row0, xw_rows = nrows, xw_split3 = (int)split3;
      }
    }
    auto out_k = [&](GemmDesc& gd, __nv_bfloat16* buf) {
      gd.ck_hi = buf, gd.ck_lo = split3 ? buf + lo_k : nullptr, gd.ck_batch = e_k, gd.ck_ld = kp_mid;
    };
    auto out_t = [&](GemmDesc& gd, __nv_bfloat16* buf) {
      gd.ct_hi = buf, gd.ct_lo = split3 ? buf + lo_t : nullptr, gd.ct_batch = e_t, gd.ct_ld = kp_b;
    };
    auto run_gemm = [&](bhmc_ctx* cx, const GemmDesc& gd, int batch) -> int {
      if (use_tc && gd.M >= 64 && gd.N >= 64 && gd.K >= 64) return tc_bgemm(cx, gd, batch, split3);
      return bhmc::run_gemm(cx, gd, batch);
    };
    // the image is valid for exactly the evaluation that follows the update launch that wrote it
    const bool mirror = mir_fresh == q && mir_hi && C <= mir_C && use_tc && fuse;
    mir_fresh = nullptr;
    const bool fu = fused_upd != nullptr;
    const bool w2t_ok = w2t_fresh == q && w2t && use_tc && fuse && sizeof(__nv_bfloat16) * 2 * (size_t)C * n_mid * kp_mid <= w2t_bytes;
    w2t_fresh = nullptr;
    const int64_t e_w2t = (int64_t)n_mid * kp_mid;  // elements per chain of the W2^T copy; the lo half follows C_alloc chains
    const int64_t w2t_lo = (int64_t)(w2t_bytes / (sizeof(__nv_bfloat16) * 2));
    UpdOpt uo{};
    if (fu) {
      if (!(fuse && head)) {
        set_error("fused sampler update requested on a path without producer-written operand copies");
        return BHMC_ERR_STATE;
      }
      uo.on = 1;
      uo.u = *fused_upd;
    }
    BHMC_CUDA_OK(cudaMemsetAsync(stat, 0, sizeof(double) * C, ctx->stream));
    {
      GroupTimer t(ctx, KG_FWD);
      // H1 = relu(dropout(X W1^T + b1))
      GemmDesc d = base();
      d.A = Xb, d.a_batch = 0, d.a_rs = n_in, d.a_cs = 1;
      d.B = q + oW1, d.b_batch = ld, d.b_rs = 1, d.b_cs = n_in;
      d.C = H1, d.c_batch = act, d.c_rs = n_mid;
      d.M = B, d.N = n_mid, d.K = n_in;
      d.bias = q + ob1, d.bias_batch = ld;
      d.epi = 1, d.layer_a = 0, d.mask_a = masks ? masks : nullptr;
      if (fuse) out_k(d, H1k), out_t(d, H1t);
      if (xcache) d.a_hi = xw, d.a_lo = split3 ? xw + e_xk : nullptr, d.a_kp = kp_in;
      if (mirror) d.b_hi = mir_hi + oW1, d.b_lo = split3 ? mir_lo + oW1 : nullptr, d.b_kp = n_in, d.b_zs = mir_ld;
      BHMC_TRY(run_gemm(ctx, d, C));
      // H2d = dropout(relu(dropout(H1 W2^T + b2)))
      d = base();
      d.A = H1, d.a_batch = act, d.a_rs = n_mid, d.a_cs = 1;
      d.B = q + oW2, d.b_batch = ld, d.b_rs = 1, d.b_cs = n_mid;
      d.C = H2d, d.c_batch = act, d.c_rs = n_mid;
      d.M = B, d.N = n_mid, d.K = n_mid;
      d.bias = q + ob2, d.bias_batch = ld;
      d.epi = 2, d.layer_a = 1, d.layer_b = 2;
      d.mask_a = masks ? masks + (size_t)C * mstride : nullptr;
      d.mask_b = masks ? masks + 2 * (size_t)C * mstride : nullptr;
      if (fuse) d.a_hi = H1k, d.a_lo = split3 ? H1k + lo_k : nullptr, d.a_kp = kp_mid;
      if (mirror) d.b_hi = mir_hi + oW2, d.b_lo = split3 ? mir_lo + oW2 : nullptr, d.b_kp = n_mid, d.b_zs = mir_ld;
      BHMC_TRY(run_gemm(ctx, d, C));
      if (!(head && g)) {
        // Z = H2d W3^T + b3
        d = base();
        d.A = H2d, d.a_batch = act, d.a_rs = n_mid, d.a_cs = 1;
        d.B = q + oW3, d.b_batch = ld, d.b_rs = 1, d.b_cs = n_mid;
        d.C = Z, d.c_batch = (int64_t)B * n_out, d.c_rs = n_out;
        d.M = B, d.N = n_out, d.K = n_mid;
        d.bias = q + ob3, d.bias_batch = ld;
        BHMC_TRY(run_gemm(ctx, d, C));
        if (logits_sink) {
          BHMC_CUDA_OK(cudaMemcpyAsync(logits_sink, Z, sizeof(float) * (size_t)C * B * n_out, cudaMemcpyDeviceToDevice, ctx->stream));
          return BHMC_OK;
        }
        dim3 grid((unsigned)ceil_div(B, 256), C);
        k_mlp_loss<<<grid, 256, 0, ctx->stream>>>(Z, B, n_out, labels + row0, stat, g ? 1 : 0);
        ctx->launches++;
      }
    }
    if (!g) return BHMC_OK;
    GroupTimer t(ctx, KG_BWD);
    const float ha = 0.5f * alpha;
    GemmDesc d = base();
    if (head) {
      // loss, dZ, gW3, gb3, gb2 and dA2 (bf16 hi/lo, both orientations) in one pass over H2d
      __nv_bfloat16 *klo = split3 ? dA2k + lo_k : nullptr, *tlo = split3 ? dA2t + lo_t : nullptr;
      const float gs = keep_inv * keep_inv;
#define BHMC_HEAD(NOV)                                                                                                          \
  BHMC_TRY(launch_head<NOV>(ctx, H2d, act, B, n_mid, n_out, q, ld, oW3, ob3, ob2, labels + row0, gs, ha, stat, dA2k, klo, e_k, kp_mid, \
                            dA2t, tlo, e_t, kp_b, head_part, C, g, uo))
      if (NO == 4) BHMC_HEAD(4);
      else if (NO == 8) BHMC_HEAD(8);
      else if (NO == 12) BHMC_HEAD(12);
      else BHMC_HEAD(16);
#undef BHMC_HEAD
    } else {
    // gW3 = dZ^T H2d + alpha/2 W3 ; gb3
    d.A = Z, d.a_batch = (int64_t)B * n_out, d.a_rs = 1, d.a_cs = n_out;  // (m=o, k=b)
    d.B = H2d, d.b_batch = act, d.b_rs = n_mid, d.b_cs = 1;
    d.C = g + oW3, d.c_batch = ld, d.c_rs = n_mid;
    d.M = n_out, d.N = n_mid, d.K = B;
    d.addsrc = q + oW3, d.add_batch = ld, d.add_rs = n_mid, d.add_scale = ha;
    BHMC_TRY(run_gemm(ctx, d, C));
    k_mlp_colsum<<<dim3((unsigned)ceil_div(n_out, 32), C), 1024, 0, ctx->stream>>>(Z, B, n_out, (int64_t)B * n_out, q, ld, ob3, ha, g, UpdOpt{});
    // dA2 = (dZ W3) * [H2d > 0] / keep^2
    d = base();
    d.A = Z, d.a_batch = (int64_t)B * n_out, d.a_rs = n_out, d.a_cs = 1;
    d.B = q + oW3, d.b_batch = ld, d.b_rs = n_mid, d.b_cs = 1;
    d.C = dA2, d.c_batch = act, d.c_rs = n_mid;
    d.M = B, d.N = n_mid, d.K = n_out;
    d.gate = H2d, d.gate_batch = act, d.gate_rs = n_mid, d.gate_scale = keep_inv * keep_inv;
    if (fuse) out_k(d, dA2k);  // its transposed copy (for gW2) stays a split launch: see k_mlp_gemm_small_k
    BHMC_TRY(run_gemm(ctx, d, C));
    }
    auto make_gW2 = [&](GemmDesc& d) {  // gW2 = dA2^T H1 + alpha/2 W2 ; gb2
      d = base();
      d.A = dA2, d.a_batch = act, d.a_rs = 1, d.a_cs = n_mid;
      d.B = H1, d.b_batch = act, d.b_rs = n_mid, d.b_cs = 1;
      d.C = g + oW2, d.c_batch = ld, d.c_rs = n_mid;
      d.M = n_mid, d.N = n_mid, d.K = B;
      d.addsrc = q + oW2, d.add_batch = ld, d.add_rs = n_mid, d.add_scale = ha;
      if (fuse) d.b_hi = H1t, d.b_lo = split3 ? H1t + lo_t : nullptr, d.b_kp = kp_b;
      if (head) d.a_hi = dA2t, d.a_lo = split3 ? dA2t + lo_t : nullptr, d.a_kp = kp_b;
      if (fu) {  // the epilogue updates W2 and leaves W2^T of the new position for the next evaluation's dA1 GEMM
        d.upd_on = 1, d.upd_off = oW2, d.upd = *fused_upd;
        d.ct_hi = w2t, d.ct_lo = split3 ? w2t + w2t_lo : nullptr, d.ct_batch = e_w2t, d.ct_ld = kp_mid;
      }
    };
    auto do_gW2 = [&]() -> int {
      GemmDesc d;
      make_gW2(d);
      BHMC_TRY(run_gemm(ctx, d, C));
      if (!head) k_mlp_colsum<<<dim3((unsigned)ceil_div(n_mid, 32), C), 1024, 0, ctx->stream>>>(dA2, B, n_mid, act, q, ld, ob2, ha, g, UpdOpt{});
      return BHMC_OK;
    };
    auto make_dA1 = [&](GemmDesc& d) {  // dA1 = (dA2 W2) * [H1 > 0] / keep
      d = base();
      d.A = dA2, d.a_batch = act, d.a_rs = n_mid, d.a_cs = 1;
      d.B = q + oW2, d.b_batch = ld, d.b_rs = n_mid, d.b_cs = 1;
      d.C = dA1, d.c_batch = act, d.c_rs = n_mid;
      d.M = B, d.N = n_mid, d.K = n_mid;
      d.gate = H1, d.gate_batch = act, d.gate_rs = n_mid, d.gate_scale = keep_inv;
      if (fuse) {
        d.a_hi = dA2k, d.a_lo = split3 ? dA2k + lo_k : nullptr, d.a_kp = kp_mid;
        out_t(d, dA1t);
      }
      if (w2t_ok) d.b_hi = w2t, d.b_lo = split3 ? w2t + w2t_lo : nullptr, d.b_kp = kp_mid;  // written by the previous fused evaluation
    };
    auto do_dA1 = [&]() -> int {
      GemmDesc d;
      make_dA1(d);
      BHMC_TRY(run_gemm(ctx, d, C));
      return BHMC_OK;
    };
    // fused update: the W2-gradient epilogue overwrites W2 and its transposed copy, which the dA1 GEMM reads -> dA1 first
    if (fu) {
      BHMC_TRY(do_dA1());
      BHMC_TRY(do_gW2());
    } else if (use_tc && head && n_mid >= 64 && B >= 64) {
      // both on the tensor cores and independent of each other: one grouped launch (tc_bgemm_two; BHMC_BG_GROUP=0: two)
      GemmDesc d0, d1;
      make_gW2(d0);
      make_dA1(d1);
      BHMC_TRY(tc_bgemm_two(ctx, d0, d1, C, split3));
    } else {
      BHMC_TRY(do_gW2());
      BHMC_TRY(do_dA1());
    }
    // gW1 = dA1^T X + alpha/2 W1 ; gb1
    d = base();
    d.A = dA1, d.a_batch = act, d.a_rs = 1, d.a_cs = n_mid;
    d.B = Xb, d.b_batch = 0, d.b_rs = n_in, d.b_cs = 1;
    d.C = g + oW1, d.c_batch = ld, d.c_rs = n_in;
    d.M = n_mid, d.N = n_in, d.K = B;
    d.addsrc = q + oW1, d.add_batch = ld, d.add_rs = n_in, d.add_scale = ha;
    if (fuse) d.a_hi = dA1t, d.a_lo = split3 ? dA1t + lo_t : nullptr, d.a_kp = kp_b;
    if (xcache) d.b_hi = xw + 2 * e_xk, d.b_lo = split3 ? xw + 2 * e_xk + e_xt : nullptr, d.b_kp = kp_b;
    if (fu) d.upd_on = 1, d.upd_off = oW1, d.upd = *fused_upd;
    BHMC_TRY(run_gemm(ctx, d, C));
    k_mlp_colsum<<<dim3((unsigned)ceil_div(n_mid, 32), C), 1024, 0, ctx->stream>>>(dA1, B, n_mid, act, q, ld, ob1, ha, g, uo);
    ctx->launches += head ? 1 : 3;
    // padding columns of g (ld > P) are never read by the update kernels beyond P; keep them finite
    BHMC_CUDA_OK(cudaGetLastError());
    return BHMC_OK;
  }
};

ModelBase* mlp_model_new(bhmc_ctx* ctx, int64_t n_rows, int n_in, int n_mid, int n_out, float alpha, float ratio,
                         uint64_t seed, int64_t chain_id0) {
  auto* m = new (std::nothrow) MlpModel();
  if (!m) return nullptr;
  m->ctx = ctx;
  m->N = n_rows;
  m->n_in = n_in;
  m->n_mid = n_mid;
  m->n_out = n_out;
  m->alpha = alpha;
  m->ratio = ratio;
  m->seed = seed;
  m->chain_id0 = chain_id0;
  int64_t lens[6] = {(int64_t)n_mid * n_in, n_mid, (int64_t)n_mid * n_mid, n_mid, (int64_t)n_out * n_mid, n_out};
  int64_t off = 0;
  m->n_vars = 6;
  for (int v = 0; v < 6; ++v) {
    m->var_off[v] = off;
    m->var_len[v] = lens[v];
    off += lens[v];
  }
  m->P = off;
  m->oW1 = m->var_off[0], m->ob1 = m->var_off[1], m->oW2 = m->var_off[2], m->ob2 = m->var_off[3];
  m->oW3 = m->var_off[4], m->ob3 = m->var_off[5];
  return m;
}

int mlp_model_bind(ModelBase* mb, const float* X, const int32_t* labels, int is_host) {
  auto* m = dynamic_cast<MlpModel*>(mb);
  BHMC_CHECK_ARG(m && X && labels, "not an mlp model / NULL data");
  m->xw_row0 = -1;  // operand copies of the previous window are stale
  if (is_host) {
    size_t xb = sizeof(float) * (size_t)m->N * m->n_in, yb = sizeof(int32_t) * (size_t)m->N;
    if (!m->X_owned) BHMC_CUDA_OK(cudaMalloc(&m->X_owned, xb));
    if (!m->y_owned) BHMC_CUDA_OK(cudaMalloc(&m->y_owned, yb));
    BHMC_CUDA_OK(cudaMemcpyAsync(m->X_owned, X, xb, cudaMemcpyHostToDevice, m->ctx->stream));
    BHMC_CUDA_OK(cudaMemcpyAsync(m->y_owned, labels, yb, cudaMemcpyHostToDevice, m->ctx->stream));
    m->X = m->X_owned;
    m->labels = m->y_owned;
  } else {
    m->X = X;
    m->labels = labels;
  }
  return BHMC_OK;
}

// row softmax / argmax of the logits [C, B, n_out] (mlp.py:84-95)
__global__ void __launch_bounds__(256) k_mlp_predict(const float* __restrict__ Z, int B, int n_out, float* __restrict__ probs,
                                                     int32_t* __restrict__ lab) {
  const int c = blockIdx.y;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B) return;
  const float* z = Z + ((int64_t)c * B + r) * n_out;
  float m = -INFINITY;
  int am = 0;
  for (int k = 0; k < n_out; ++k)
    if (z[k] > m) m = z[k], am = k;
  if (lab) lab[(int64_t)c * B + r] = am;
  if (probs) {
    float s = 0.f;
    for (int k = 0; k < n_out; ++k) s += expf(z[k] - m);
    const float inv = 1.0f / s;
    for (int k = 0; k < n_out; ++k) probs[((int64_t)c * B + r) * n_out + k] = expf(z[k] - m) * inv;
  }
}

// mlp.predict (mlp.py:84-95): forward pass of every chain on caller rows X_dev [nrows, n_in] -- dropout stays ON, as in
// the reference (no train=False there) -- then softmax / argmax.  probs_dev [C, nrows, n_out], labels_dev [C, nrows].
int mlp_model_predict(ModelBase* mb, const float* q, int C, int64_t ld, const float* X_dev, int64_t nrows, int prec,
                      float* probs_dev, int32_t* labels_dev) {
  auto* m = dynamic_cast<MlpModel*>(mb);
  BHMC_CHECK_ARG(m && q && X_dev && nrows > 0 && (probs_dev || labels_dev), "bad argument");
  void* zb = nullptr;
  BHMC_TRY(m->ctx->get_scratch(10, sizeof(float) * (size_t)C * nrows * m->n_out, &zb));
  const float* X0 = m->X;
  const int64_t N0 = m->N;
  const uint8_t* masks0 = m->masks;  // injected masks are shaped for the training batch: predict draws Philox masks
  m->X = X_dev;
  m->N = nrows;
  m->masks = nullptr;
  m->logits_sink = (float*)zb;
  double* stat = nullptr;
  void* sb = nullptr;
  int rc = m->ctx->get_scratch(11, sizeof(double) * C, &sb);
  stat = (double*)sb;
  if (rc == BHMC_OK) rc = m->grad(q, C, ld, 0, nrows, prec, nullptr, stat, 0);
  m->X = X0;
  m->N = N0;
  m->masks = masks0;
  m->logits_sink = nullptr;
  BHMC_TRY(rc);
  dim3 grid((unsigned)ceil_div(nrows, 256), C);
  k_mlp_predict<<<grid, 256, 0, m->ctx->stream>>>((const float*)zb, (int)nrows, m->n_out, probs_dev, labels_dev);
  m->ctx->launches++;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

int mlp_model_set_masks(ModelBase* mb, const uint8_t* masks_dev) {
  auto* m = dynamic_cast<MlpModel*>(mb);
  BHMC_CHECK_ARG(m, "not an mlp model");
  m->masks = masks_dev;
  return BHMC_OK;
}

}  // namespace bhmc
