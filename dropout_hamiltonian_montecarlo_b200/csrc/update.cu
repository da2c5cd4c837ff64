// Fused elementwise kernels of the sampler inner loop (HBM-bound, no tensor cores):
//   momentum draw + kinetic energy            hamiltonian/inference/cpu/hmc.py:41,74-87
//   leapfrog kicks / drift (Gauss-Seidel)     hmc.py:49-54
//   SGHMC friction + noise                    hamiltonian/inference/cpu/sghmc.py:28-34
//   SGLD step                                 hamiltonian/inference/cpu/sgld.py:31-46
//   Metropolis accept / select / sample sink  hmc.py:58-71,112-114
//   heavy-ball SGD                            hamiltonian/inference/cpu/sgd.py:38-41
// Chain state is [C, ld] fp32 with ld % 4 == 0, so every row is 16-byte aligned and each
// thread moves one float4 per array.  grid = (ceil(P/4/256), C): coalesced, chain-uniform
// control flow (activity mask, accept decision) is block-uniform.
#include <cuda_bf16.h>
#include "internal.cuh"
#include "philox.cuh"
#include "stream_ops.cuh"

namespace bhmc {

static constexpr int TPB = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void block_atomic_add(double v, double* dst) {
  __shared__ double sm[TPB / 32];
  v = warp_sum(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sm[w] = v;
  __syncthreads();
  if (w == 0) {
    v = (l < TPB / 32) ? sm[l] : 0.0;
    v = warp_sum(v);
    if (l == 0) atomicAdd(dst, quantize_addend<24>(v));  // order-independent: see internal.cuh
  }
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// streaming store: samples are written once and read back by the host only
__device__ __forceinline__ void st4_cs(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

__device__ __forceinline__ float4 load_z(const float* z, int64_t ld_z, int c, int64_t i, int64_t P) {
  // injected tapes are compact ([C, P], P arbitrary) -> scalar loads
  const float* r = z + (int64_t)c * ld_z + i;
  float4 v;
  v.x = r[0];
  v.y = (i + 1 < P) ? r[1] : 0.f;
  v.z = (i + 2 < P) ? r[2] : 0.f;
  v.w = (i + 3 < P) ? r[3] : 0.f;
  return v;
}

// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_hmc_begin(BeginArgs a) {
  int r = blockIdx.y;                    // working row
  int c = a.perm ? a.perm[r] : r;        // chain held by this row
  int t = 0;
  if (a.code) {  // streaming schedule: block-uniform opt-out, per-row step
    const uint32_t op = a.code[r];
    if (!(op & OP_BEGIN)) return;
    t = a.step[r] + ((op & OP_FINISH) ? 1 : 0);
    a.kin0 += (int64_t)(t & 1) * a.C_total;
    if (a.z) a.z += (int64_t)t * a.z_step_stride;
    a.stream_lo += (uint32_t)t;
  }
  int64_t i = ((int64_t)blockIdx.x * TPB + threadIdx.x) * 4;
  double k = 0.0;
  if (i < a.P) {
    int64_t oc = (int64_t)c * a.ld + i, orow = (int64_t)r * a.ld + i;
    float4 z = a.z ? load_z(a.z, a.ld_z, c, i, a.P)
                   : philox_normal4(a.seed, a.chain_id0 + c, (uint32_t)(i >> 2), a.stream_lo, a.stream_hi);
    if (i + 1 >= a.P) z.y = 0.f;
    if (i + 2 >= a.P) z.z = 0.f;
    if (i + 3 >= a.P) z.w = 0.f;
    st4(a.q_new + orow, ld4(a.q + oc));
    st4(a.p0 + oc, z);
    st4(a.p_new + orow, z);
    if (a.g_start && (a.code[r] & OP_FINISH)) {  // streaming: the gradient at the start point is already known
      if (a.acc_flag[r]) st4(a.g_start + orow, ld4(a.g + orow));   // accepted: it is the last evaluation
      else st4(a.g + orow, ld4(a.g_start + orow));                  // rejected: it is the previous start point's
    }
    k = 0.5 * ((double)z.x * z.x + (double)z.y * z.y + (double)z.z * z.z + (double)z.w * z.w);
  }
  block_atomic_add(k, a.kin0 + r);
}

int launch_hmc_begin(bhmc_ctx* ctx, const BeginArgs& a) {
  GroupTimer t(ctx, KG_UPDATE);
  if (!a.code) BHMC_CUDA_OK(cudaMemsetAsync(a.kin0, 0, sizeof(double) * a.C, ctx->stream));
  dim3 grid((unsigned)ceil_div(ceil_div(a.P, 4), TPB), a.C);
  k_hmc_begin<<<grid, TPB, 0, ctx->stream>>>(a);
  ctx->launches++;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

// ---------------------------------------------------------------------------------------
template <bool NOISE>
__global__ void __launch_bounds__(TPB) k_hmc_update(UpdateArgs a) {
  int c = blockIdx.y;
  int Lc = a.L[c];
  bool act_post = a.post_len > 0 && a.it_post < Lc - 1;
  bool act_pre = a.pre_len > 0 && a.it_pre < Lc - 1;
  if (blockIdx.x == 0 && threadIdx.x == 0 && act_post && a.stat_new) a.stat_new[c] = a.stat[c];
  if (!act_post && !act_pre) return;
  int64_t i = ((int64_t)blockIdx.x * TPB + threadIdx.x) * 4;
  if (i >= a.P) return;
  bool hit_post = act_post && i < a.post_off + a.post_len && i + 4 > a.post_off;
  bool hit_pre = act_pre && i < a.pre_off + a.pre_len && i + 4 > a.pre_off;
  if (!hit_post && !hit_pre) return;
  int64_t o = (int64_t)c * a.ld + i;
  float4 p4 = ld4(a.p + o);
  float4 g4 = ld4(a.g + o);
  float4 q4 = hit_pre ? ld4(a.q + o) : make_float4(0, 0, 0, 0);
  float4 z4 = make_float4(0, 0, 0, 0);
  if (NOISE && hit_post) {
    int chain = a.perm ? a.perm[c] : c;  // noise streams are keyed by chain, not by working row
    z4 = a.z ? load_z(a.z, a.ld_z, chain, i, a.P)
             : philox_normal4(a.seed, a.chain_id0 + chain, (uint32_t)(i >> 2), a.stream_lo, a.stream_hi);
  }
  float pe[4] = {p4.x, p4.y, p4.z, p4.w}, ge[4] = {g4.x, g4.y, g4.z, g4.w}, qe[4] = {q4.x, q4.y, q4.z, q4.w};
  float ze[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    int64_t idx = i + e;
    if (act_post && idx >= a.post_off && idx < a.post_off + a.post_len) {
      float v = a.f_post * pe[e] - a.a_post * ge[e];
      if (NOISE) v += a.n_post * ze[e];
      pe[e] = v;
    }
    if (act_pre && idx >= a.pre_off && idx < a.pre_off + a.pre_len) {
      pe[e] = pe[e] - a.a_pre * ge[e];
      qe[e] = qe[e] + a.eps * pe[e];
    }
  }
  st4(a.p + o, make_float4(pe[0], pe[1], pe[2], pe[3]));
  if (hit_pre) {
    st4(a.q + o, make_float4(qe[0], qe[1], qe[2], qe[3]));
    if (a.mir_hi) {  // bf16 hi/lo image of the new position (operand mirror of the model)
      const int64_t om = (int64_t)c * a.mir_ld + i;
      __nv_bfloat16 h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        h[e] = __float2bfloat16_rn(qe[e]);
        l[e] = __float2bfloat16_rn(qe[e] - __bfloat162float(h[e]));
      }
      *reinterpret_cast<uint2*>(a.mir_hi + om) = *reinterpret_cast<const uint2*>(h);
      if (a.mir_lo) *reinterpret_cast<uint2*>(a.mir_lo + om) = *reinterpret_cast<const uint2*>(l);
    }
  }
}

int launch_hmc_update(bhmc_ctx* ctx, const UpdateArgs& a) {
  GroupTimer t(ctx, KG_UPDATE);
  dim3 grid((unsigned)ceil_div(ceil_div(a.P, 4), TPB), a.C);
  if (a.n_post != 0.f)
    k_hmc_update<true><<<grid, TPB, 0, ctx->stream>>>(a);
  else
    k_hmc_update<false><<<grid, TPB, 0, ctx->stream>>>(a);
  ctx->launches++;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

// ---------------------------------------------------------------------------------------
// Streaming schedule: the same kick / drift arithmetic as k_hmc_update (hmc.py:51-54), but every row executes the op
// its own trajectory has reached.  Thread 0 of a row's first block also does the row's scalar bookkeeping.
__global__ void __launch_bounds__(TPB) k_stream_update(StreamUpdateArgs a) {
  const int r = blockIdx.y;
  const uint32_t op = a.code[r];
  if (blockIdx.x == 0 && threadIdx.x == 0) stream_row_scalars(a, r, op);
  const int64_t i = ((int64_t)blockIdx.x * TPB + threadIdx.x) * 4;
  if (i >= a.P) return;
  if (op & OP_LATCH) st4(a.g_start + (int64_t)r * a.ld + i, ld4(a.g + (int64_t)r * a.ld + i));
  if (!(op & (OP_POST | OP_PRE))) return;
  const int pv = (op >> 8) & 15, v = (op >> 12) & 15;
  const int64_t post_off = a.off[pv], post_len = (op & OP_POST) ? a.len[pv] : 0;
  const int64_t pre_off = a.off[v], pre_len = (op & OP_PRE) ? a.len[v] : 0;
  const bool hit_post = post_len > 0 && i < post_off + post_len && i + 4 > post_off;
  const bool hit_pre = pre_len > 0 && i < pre_off + pre_len && i + 4 > pre_off;
  if (!hit_post && !hit_pre) return;
  const int64_t o = (int64_t)r * a.ld + i;
  float4 p4 = ld4(a.p + o);
  float4 g4 = ld4(a.g + o);
  float4 q4 = hit_pre ? ld4(a.q + o) : make_float4(0, 0, 0, 0);
  float pe[4] = {p4.x, p4.y, p4.z, p4.w}, ge[4] = {g4.x, g4.y, g4.z, g4.w}, qe[4] = {q4.x, q4.y, q4.z, q4.w};
  stream_apply4(a, op, i, pe, ge, qe);
  st4(a.p + o, make_float4(pe[0], pe[1], pe[2], pe[3]));
  if (hit_pre) st4(a.q + o, make_float4(qe[0], qe[1], qe[2], qe[3]));
}

int launch_stream_update(bhmc_ctx* ctx, const StreamUpdateArgs& a) {
  GroupTimer t(ctx, KG_UPDATE);
  dim3 grid((unsigned)ceil_div(ceil_div(a.P, 4), TPB), a.rows);
  k_stream_update<<<grid, TPB, 0, ctx->stream>>>(a);
  ctx->launches++;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

__global__ void __launch_bounds__(TPB) k_stream_kinetic(const float* p, int64_t ld, int64_t P, const uint32_t* code,
                                                        double* kin) {
  const int r = blockIdx.y;
  if (!(code[r] & OP_FINISH)) return;
  const int64_t i = ((int64_t)blockIdx.x * TPB + threadIdx.x) * 4;
  double k = 0.0;
  if (i < P) {
    float4 v = ld4(p + (int64_t)r * ld + i);
    k = (double)v.x * v.x;
    if (i + 1 < P) k += (double)v.y * v.y;
    if (i + 2 < P) k += (double)v.z * v.z;
    if (i + 3 < P) k += (double)v.w * v.w;
    k *= 0.5;
  }
  block_atomic_add(k, kin + r);
}

int launch_stream_kinetic(bhmc_ctx* ctx, const float* p, int64_t ld, int64_t P, int rows, const uint32_t* code,
                          double* kin) {
  GroupTimer t(ctx, KG_UPDATE);
  dim3 grid((unsigned)ceil_div(ceil_div(P, 4), TPB), rows);
  k_stream_kinetic<<<grid, TPB, 0, ctx->stream>>>(p, ld, P, code, kin);
  ctx->launches++;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_kinetic(const float* p, int64_t ld, int64_t P, double* kin) {
  int c = blockIdx.y;
  int64_t i = ((int64_t)blockIdx.x * TPB + threadIdx.x) * 4;
  double k = 0.0;
  if (i < P) {
    float4 v = ld4(p + (int64_t)c * ld + i);
    k = (double)v.x * v.x;
    if (i + 1 < P) k += (double)v.y * v.y;
    if (i + 2 < P) k += (double)v.z * v.z;
    if (i + 3 < P) k += (double)v.w * v.w;
    k *= 0.5;
  }
  block_atomic_add(k, kin + c);
}

int launch_kinetic(bhmc_ctx* ctx, const float* p, int64_t ld, int64_t P, int C, double* kin) {
  GroupTimer t(ctx, KG_UPDATE);
  BHMC_CUDA_OK(cudaMemsetAsync(kin, 0, sizeof(double) * C, ctx->stream));
  dim3 grid((unsigned)ceil_div(ceil_div(P, 4), TPB), C);
  k_kinetic<<<grid, TPB, 0, ctx->stream>>>(p, ld, P, kin);
  ctx->launches++;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

struct VarLayout {
  int n;
  int64_t off[BHMC_MAX_VARS], len[BHMC_MAX_VARS];
};

__global__ void __launch_bounds__(TPB) k_sumsq(const float* q, int64_t ld, VarLayout vl, double* out) {
  int c = blockIdx.y, v = blockIdx.z;
  int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
  double s = 0.0;
  for (; i < vl.len[v]; i += (int64_t)gridDim.x * TPB) {
    float x = q[(int64_t)c * ld + vl.off[v] + i];
    s += (double)x * x;
  }
  block_atomic_add(s, out + (int64_t)c * vl.n + v);
}

int launch_sumsq(bhmc_ctx* ctx, const float* q, int64_t ld, int C, int n_vars, const int64_t* off,
                 const int64_t* len, double* out) {
  GroupTimer t(ctx, KG_UPDATE);
  VarLayout vl;
  vl.n = n_vars;
  int64_t mx = 1;
  for (int v = 0; v < n_vars; ++v) {
    vl.off[v] = off[v];
    vl.len[v] = len[v];
    if (len[v] > mx) mx = len[v];
  }
  BHMC_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(double) * C * n_vars, ctx->stream));
  int64_t bx = ceil_div(mx, TPB * 8);
  if (bx > 64) bx = 64;
  dim3 grid((unsigned)bx, C, n_vars);
  k_sumsq<<<grid, TPB, 0, ctx->stream>>>(q, ld, vl, out);
  ctx->launches++;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

struct CvArr {
  int n;
  double cv[BHMC_MAX_VARS];
};
__global__ void k_prior_combine(const double* sumsq, CvArr cv, double* out, int C, const uint32_t* code, uint32_t flag) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (code && !(code[c] & flag)) return;
  double e = 0.0;
  for (int v = 0; v < cv.n; ++v) e += cv.cv[v] * sumsq[(int64_t)c * cv.n + v];
  out[c] = e;
}

// out[c] = sum_v cv[v] * |q_v[c]|^2 : the quadratic log-prior part of the Metropolis energy
// (models/gpu/softmax.py:29-39, models/gpu/mlp.py:40-45)
int launch_prior_energy(bhmc_ctx* ctx, const float* q, int64_t ld, int C, int n_vars, const int64_t* off,
                        const int64_t* len, const double* cv, double* sumsq_scratch, double* out, const uint32_t* code,
                        uint32_t flag) {
  BHMC_TRY(launch_sumsq(ctx, q, ld, C, n_vars, off, len, sumsq_scratch));
  CvArr a;
  a.n = n_vars;
  for (int v = 0; v < n_vars; ++v) a.cv[v] = cv[v];
  k_prior_combine<<<(C + 127) / 128, 128, 0, ctx->stream>>>(sumsq_scratch, a, out, C, code, flag);
  ctx->launches++;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

// ---------------------------------------------------------------------------------------
// Metropolis test (hmc.py:60-63,67-71) + state select + sample sink.  The decision is
// recomputed identically by every block of a chain (a handful of double ops).
__global__ void __launch_bounds__(TPB) k_accept(AcceptArgs a) {
  int r = blockIdx.y;                    // working row (energies, proposal)
  int c = a.perm ? a.perm[r] : r;        // chain (state, draws, outputs)
  if (a.code) {  // streaming schedule: block-uniform opt-out, per-row step
    if (!(a.code[r] & OP_FINISH)) return;
    const int64_t t = a.step[r];
    a.kin0 += (t & 1) * a.C_total;
    a.stream_lo += (uint32_t)t;
    if (a.u) a.u += t * a.C_total;
    if (a.sample) a.sample += t * a.C_total * a.P;
    if (a.loss) a.loss += t * a.C_total;
    if (a.accept_prob) a.accept_prob += t * a.C_total;
    if (a.accepted) a.accepted += t * a.C_total;
  }
  double u_cur = a.ea * a.stat_cur[r] + a.eb, u_new = a.ea * a.stat_new[r] + a.eb;
  if (a.extra_cur) {
    u_cur += a.extra_cur[r];
    u_new += a.extra_new[r];
  }
  double e_cur = u_cur + a.kin0[r], e_new = u_new + a.kin1[r];
  double x = exp(e_cur - e_new);
  // Python's builtin min(1, x): returns x only when x < 1 (so NaN -> 1), hmc.py:70
  double A = (x < 1.0) ? x : 1.0;
  if (a.reject_nan && !(isfinite(e_new))) A = 0.0;
  double u = a.u ? a.u[c] : philox_uniform(a.seed, a.chain_id0 + c, a.stream_lo, a.stream_hi);
  bool acc = isfinite(A) && (u < A);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (a.stat_next) a.stat_next[r] = acc ? a.stat_new[r] : a.stat_cur[r];
    if (a.extra_next && a.extra_cur) a.extra_next[r] = acc ? a.extra_new[r] : a.extra_cur[r];
    if (a.acc_flag) a.acc_flag[r] = acc ? 1 : 0;
    if (a.accept_prob) a.accept_prob[c] = A;
    if (a.accepted) a.accepted[c] = acc ? 1 : 0;
    if (a.loss) a.loss[c] = acc ? u_new : u_cur;
  }
  int64_t i = ((int64_t)blockIdx.x * TPB + threadIdx.x) * 4;
  if (i >= a.P) return;
  int64_t oc = (int64_t)c * a.ld + i, orow = (int64_t)r * a.ld + i;
  float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (acc) {
    qv = ld4(a.q_new + orow);
    st4(a.q + oc, qv);
    float4 pv = ld4(a.p_new + orow);
    st4(a.p_out + oc, make_float4(a.p_sign * pv.x, a.p_sign * pv.y, a.p_sign * pv.z, a.p_sign * pv.w));
  } else if (a.sample) {
    qv = ld4(a.q + oc);
  }
  if (a.sample) {
    float* s = a.sample + (int64_t)c * a.P + i;  // compact rows: scalar stores at the tail
    if (((a.P & 3) == 0)) {
      st4_cs(s, qv);
    } else {
      s[0] = qv.x;
      if (i + 1 < a.P) s[1] = qv.y;
      if (i + 2 < a.P) s[2] = qv.z;
      if (i + 3 < a.P) s[3] = qv.w;
    }
  }
}

int launch_accept(bhmc_ctx* ctx, const AcceptArgs& a) {
  GroupTimer t(ctx, KG_UPDATE);
  dim3 grid((unsigned)ceil_div(ceil_div(a.P, 4), TPB), a.C);
  k_accept<<<grid, TPB, 0, ctx->stream>>>(a);
  ctx->launches++;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_sgld(SgldArgs a) {
  int c = blockIdx.y;
  int64_t i = ((int64_t)blockIdx.x * TPB + threadIdx.x) * 4;
  if (i >= a.P) return;
  int64_t o = (int64_t)c * a.ld + i;
  float4 z = a.z ? load_z(a.z, a.ld_z, c, i, a.P)
                 : philox_normal4(a.seed, a.chain_id0 + c, (uint32_t)(i >> 2), a.stream_lo, a.stream_hi);
  float4 g = ld4(a.g + o), q = ld4(a.q + o);
  float s = 2.0f * a.eps, h = 0.5f * a.eps;
  float4 p = make_float4(s * z.x - h * g.x, s * z.y - h * g.y, s * z.z - h * g.z, s * z.w - h * g.w);
  st4(a.p + o, p);
  st4(a.q + o, make_float4(q.x + p.x, q.y + p.y, q.z + p.z, q.w + p.w));
}

int launch_sgld_update(bhmc_ctx* ctx, const SgldArgs& a) {
  GroupTimer t(ctx, KG_UPDATE);
  dim3 grid((unsigned)ceil_div(ceil_div(a.P, 4), TPB), a.C);
  k_sgld<<<grid, TPB, 0, ctx->stream>>>(a);
  ctx->launches++;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

__global__ void __launch_bounds__(TPB) k_sgd(float* q, float* m, const float* g, int64_t ld, int64_t P,
                                             float gamma, float eps) {
  int c = blockIdx.y;
  int64_t i = ((int64_t)blockIdx.x * TPB + threadIdx.x) * 4;
  if (i >= P) return;
  int64_t o = (int64_t)c * ld + i;
  float4 mv = ld4(m + o), gv = ld4(g + o), qv = ld4(q + o);
  mv = make_float4(gamma * mv.x - eps * gv.x, gamma * mv.y - eps * gv.y, gamma * mv.z - eps * gv.z,
                   gamma * mv.w - eps * gv.w);
  st4(m + o, mv);
  st4(q + o, make_float4(qv.x + mv.x, qv.y + mv.y, qv.z + mv.z, qv.w + mv.w));
}

int launch_sgd_update(bhmc_ctx* ctx, float* q, float* m, const float* g, int64_t ld, int64_t P, int C,
                      float gamma, float eps) {
  GroupTimer t(ctx, KG_UPDATE);
  dim3 grid((unsigned)ceil_div(ceil_div(P, 4), TPB), C);
  k_sgd<<<grid, TPB, 0, ctx->stream>>>(q, m, g, ld, P, gamma, eps);
  ctx->launches++;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) k_copy_rows(const float* src, int64_t ld_src, float* dst, int64_t ld_dst,
                                                   int64_t P) {
  int c = blockIdx.y;
  for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < P; i += (int64_t)gridDim.x * TPB)
    dst[(int64_t)c * ld_dst + i] = src[(int64_t)c * ld_src + i];
}

int launch_copy_rows(bhmc_ctx* ctx, const float* src, int64_t ld_src, float* dst, int64_t ld_dst, int64_t P,
                     int C) {
  dim3 grid((unsigned)ceil_div(P, TPB), C);
  k_copy_rows<<<grid, TPB, 0, ctx->stream>>>(src, ld_src, dst, ld_dst, P);
  ctx->launches++;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

__global__ void __launch_bounds__(TPB) k_philox_normal(float* out, int64_t P, int64_t ld, uint64_t seed,
                                                       int64_t chain_id0, uint32_t slo, uint32_t shi) {
  int c = blockIdx.y;
  int64_t i = ((int64_t)blockIdx.x * TPB + threadIdx.x) * 4;
  if (i >= P) return;
  float4 z = philox_normal4(seed, chain_id0 + c, (uint32_t)(i >> 2), slo, shi);
  float* r = out + (int64_t)c * ld + i;
  r[0] = z.x;
  if (i + 1 < P) r[1] = z.y;
  if (i + 2 < P) r[2] = z.z;
  if (i + 3 < P) r[3] = z.w;
}

int launch_philox_normal(bhmc_ctx* ctx, float* out, int C, int64_t P, int64_t ld, uint64_t seed,
                         int64_t chain_id0, uint32_t slo, uint32_t shi) {
  dim3 grid((unsigned)ceil_div(ceil_div(P, 4), TPB), C);
  k_philox_normal<<<grid, TPB, 0, ctx->stream>>>(out, P, ld, seed, chain_id0, slo, shi);
  ctx->launches++;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

__global__ void k_affine(const double* in, double a, double b, const double* extra, double* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a * in[i] + b + (extra ? extra[i] : 0.0);
}

int launch_affine(bhmc_ctx* ctx, const double* in, double a, double b, const double* extra, double* out, int n) {
  k_affine<<<(n + 127) / 128, 128, 0, ctx->stream>>>(in, a, b, extra, out, n);
  ctx->launches++;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

}  // namespace bhmc
