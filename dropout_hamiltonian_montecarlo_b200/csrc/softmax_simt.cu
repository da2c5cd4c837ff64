// BHMC_PREC_FP32: the softmax-regression potential and its gradient in plain fp32 FMA on
// CUDA cores.  This is the exact-fp32 *checker* path (the tensor-core path in softmax_tc.cu is
// the product); it also serves predict() where the row matrix is caller-supplied.
//
// Reference arithmetic (hamiltonian/models/cpu/softmax.py):
//   :38-43  Z = X.W + b, clipped to [-708.396.., 36.0436..]; P = softmax_row(Z)
//   :45-61  grad_W = X^T (P - Y) + alpha W ; grad_b = sum_n (P - Y) + alpha b   (sum over rows)
//   :63-72  LL = sum_n (Z[n, y_n] - logsumexp(Z[n, :]))
// Columns of all chains are concatenated: j = c*K + k.  The bias is folded in as feature D
// (x = 1), whose weight row is the chain's bias slice, since P = (D+1)*K is laid out
// [weights(D,K) | bias(K)].
#include "internal.cuh"

namespace bhmc {

static constexpr float CLIP_HI = 36.04365338911715f;    // -log(eps64)
static constexpr float CLIP_LO = -708.3964185322641f;   // -log(1/tiny64 - 1); below fp32 exp range anyway

__global__ void k_init_grad(const float* __restrict__ q, float* __restrict__ g, int64_t ld, int64_t P, float alpha,
                            double* __restrict__ ll) {
  int c = blockIdx.y;
  if (blockIdx.x == 0 && threadIdx.x == 0 && ll) ll[c] = 0.0;
  if (!g) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ld; i += (int64_t)gridDim.x * blockDim.x)
    g[(int64_t)c * ld + i] = (i < P) ? alpha * q[(int64_t)c * ld + i] : 0.f;
}

static constexpr int BM = 64, BN = 64, BK = 16;

// Z[r, j] = sum_{d<=D} Xaug[r, d] * Wcat[d, j]
__global__ void __launch_bounds__(256) k_sgemm_fwd(const float* __restrict__ X, int64_t nrows, int D, int K,
                                                   const float* __restrict__ q, int64_t ld, int Ncols,
                                                   float* __restrict__ Z) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  int64_t m0 = (int64_t)blockIdx.x * BM;
  int n0 = blockIdx.y * BN;
  float acc[4][4] = {};
  // B gather coordinates of this thread: 4 columns at k-row (t/16)
  int bk = t >> 4, bn = (t & 15) * 4;
  int64_t bbase[4];
  bool bok[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    int j = n0 + bn + e;
    bok[e] = j < Ncols;
    int c = bok[e] ? j / K : 0, k = bok[e] ? j % K : 0;
    bbase[e] = (int64_t)c * ld + k;
  }
  int am = t >> 2, ak = (t & 3) * 4;
  for (int k0 = 0; k0 <= D; k0 += BK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int d = k0 + ak + e;
      int64_t r = m0 + am;
      float v = 0.f;
      if (r < nrows) v = (d < D) ? X[r * D + d] : (d == D ? 1.f : 0.f);
      As[ak + e][am] = v;
    }
    {
      int d = k0 + bk;
#pragma unroll
      for (int e = 0; e < 4; ++e) Bs[bk][bn + e] = (bok[e] && d <= D) ? q[bbase[e] + (int64_t)d * K] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
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
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t r = m0 + ty * 4 + i;
    if (r >= nrows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int col = n0 + tx * 4 + j;
      if (col < Ncols) Z[r * Ncols + col] = acc[i][j];
    }
  }
}

// one thread per (row, chain): clip, softmax, D = P - Y in place, LL per chain.
// mode 0: grad (Z <- P - Y, ll += ...), mode 1: predict (probs/labels out)
__global__ void __launch_bounds__(256) k_softmax_rows(float* __restrict__ Z, int64_t nrows, int K, int Ncols,
                                                      const int32_t* __restrict__ labels, double* __restrict__ ll,
                                                      int mode, float* __restrict__ probs,
                                                      int32_t* __restrict__ pred) {
  int c = blockIdx.y;
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double my = 0.0;
  if (r < nrows) {
    float* z = Z + r * Ncols + (int64_t)c * K;
    float m = -INFINITY;
    int am = 0;
    for (int k = 0; k < K; ++k) {
      float v = fmaxf(fminf(z[k], CLIP_HI), CLIP_LO);
      if (v > m) m = v, am = k;
    }
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += expf(fmaxf(fminf(z[k], CLIP_HI), CLIP_LO) - m);
    float inv = 1.0f / s;
    if (mode == 0) {
      int y = labels[r];
      float zy = fmaxf(fminf(z[y], CLIP_HI), CLIP_LO);
      my = (double)(zy - m) - (double)logf(s);
      for (int k = 0; k < K; ++k) {
        float p = expf(fmaxf(fminf(z[k], CLIP_HI), CLIP_LO) - m) * inv;
        z[k] = p - (k == y ? 1.f : 0.f);
      }
    } else {
      if (probs)
        for (int k = 0; k < K; ++k)
          probs[((int64_t)c * nrows + r) * K + k] = expf(fmaxf(fminf(z[k], CLIP_HI), CLIP_LO) - m) * inv;
      if (pred) pred[(int64_t)c * nrows + r] = am;
    }
  }
  if (mode == 0) {
    __shared__ double sm[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my += __shfl_xor_sync(0xffffffffu, my, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = my;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < 8; ++w) tot += sm[w];
      atomicAdd(ll + c, quantize_addend<24>(tot));  // order-independent: see internal.cuh
    }
  }
}

// g[c, d*K + k] += sum_r Xaug[r, d] * Dm[r, j]   (row slab per blockIdx.z, fp32 atomics)
__global__ void __launch_bounds__(256) k_sgemm_bwd(const float* __restrict__ X, int64_t nrows, int D, int K,
                                                   const float* __restrict__ Dm, int Ncols, int64_t slab,
                                                   float* __restrict__ g, int64_t ld) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  int d0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  int64_t r_begin = (int64_t)blockIdx.z * slab, r_end = min(nrows, r_begin + slab);
  float acc[4][4] = {};
  int lk = t >> 4, lm = (t & 15) * 4;
  for (int64_t r0 = r_begin; r0 < r_end; r0 += BK) {
    int64_t r = r0 + lk;
    bool rok = r < r_end;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int d = d0 + lm + e;
      float v = 0.f;
      if (rok) v = (d < D) ? X[r * D + d] : (d == D ? 1.f : 0.f);
      As[lk][lm + e] = v;
      int j = n0 + lm + e;
      Bs[lk][lm + e] = (rok && j < Ncols) ? Dm[r * Ncols + j] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
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
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int d = d0 + ty * 4 + i;
    if (d > D) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int col = n0 + tx * 4 + j;
      if (col < Ncols) atomicAdd(g + (int64_t)(col / K) * ld + (int64_t)d * K + (col % K), acc[i][j]);
    }
  }
}

int simt_softmax_grad(bhmc_ctx* ctx, const SoftmaxData& d, const float* q, int C, int64_t ld, float alpha,
                      int64_t row0, int64_t nrows, float* g, double* loglik) {
  BHMC_CHECK_ARG(d.X && d.labels, "softmax model has no bound data");
  BHMC_CHECK_ARG(row0 >= 0 && nrows > 0 && row0 + nrows <= d.N, "row window [%lld,+%lld) outside the %lld bound rows",
                 (long long)row0, (long long)nrows, (long long)d.N);
  int64_t P = (int64_t)(d.D + 1) * d.K;
  int Ncols = C * d.K;
  void* zbuf = nullptr;
  BHMC_TRY(ctx->get_scratch(0, sizeof(float) * (size_t)nrows * Ncols, &zbuf));
  float* Z = (float*)zbuf;
  const float* X = d.X + row0 * d.D;
  {
    GroupTimer t(ctx, KG_PREP);
    dim3 grid((unsigned)std::min<int64_t>(ceil_div(ld, 256), 64), C);
    k_init_grad<<<grid, 256, 0, ctx->stream>>>(q, g, ld, P, alpha, loglik);
    ctx->launches++;
  }
  {
    GroupTimer t(ctx, KG_FWD);
    dim3 grid((unsigned)ceil_div(nrows, BM), (unsigned)ceil_div(Ncols, BN));
    k_sgemm_fwd<<<grid, 256, 0, ctx->stream>>>(X, nrows, d.D, d.K, q, ld, Ncols, Z);
    dim3 g2((unsigned)ceil_div(nrows, 256), C);
    k_softmax_rows<<<g2, 256, 0, ctx->stream>>>(Z, nrows, d.K, Ncols, d.labels + row0, loglik, 0, nullptr, nullptr);
    ctx->launches += 2;
  }
  if (g) {
    GroupTimer t(ctx, KG_BWD);
    int gx = (int)ceil_div(d.D + 1, BM), gy = (int)ceil_div(Ncols, BN);
    int64_t want = std::max<int64_t>(1, (4 * ctx->sm_count) / std::max(1, gx * gy));
    int64_t slab = round_up(std::max<int64_t>(ceil_div(nrows, want), 64), BK);
    int S = (int)ceil_div(nrows, slab);
    dim3 grid(gx, gy, S);
    k_sgemm_bwd<<<grid, 256, 0, ctx->stream>>>(X, nrows, d.D, d.K, Z, Ncols, slab, g, ld);
    ctx->launches++;
  }
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

int simt_softmax_predict(bhmc_ctx* ctx, int D, int K, const float* q, int C, int64_t ld, const float* X,
                         int64_t nrows, float* probs, int32_t* labels) {
  int Ncols = C * K;
  void* zbuf = nullptr;
  BHMC_TRY(ctx->get_scratch(0, sizeof(float) * (size_t)nrows * Ncols, &zbuf));
  float* Z = (float*)zbuf;
  GroupTimer t(ctx, KG_FWD);
  dim3 grid((unsigned)ceil_div(nrows, BM), (unsigned)ceil_div(Ncols, BN));
  k_sgemm_fwd<<<grid, 256, 0, ctx->stream>>>(X, nrows, D, K, q, ld, Ncols, Z);
  dim3 g2((unsigned)ceil_div(nrows, 256), C);
  k_softmax_rows<<<g2, 256, 0, ctx->stream>>>(Z, nrows, K, Ncols, nullptr, nullptr, 1, probs, labels);
  ctx->launches += 2;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

}  // namespace bhmc
