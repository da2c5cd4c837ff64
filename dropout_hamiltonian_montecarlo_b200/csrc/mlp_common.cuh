// Shared by mlp.cu (fp32 CUDA-core GEMMs) and tc_bgemm.cu (tcgen05 batched GEMMs): the description of one
// strided-batched GEMM with its fused epilogue, and the Philox dropout mask both paths must agree on.
// Include inside `namespace bhmc` after internal.cuh and philox.cuh.
#pragma once

struct GemmDesc {
  const float* A;  // element (m,k) at A + c*a_batch + m*a_rs + k*a_cs
  int64_t a_batch, a_rs, a_cs;
  const float* B;  // element (k,n) at B + c*b_batch + k*b_rs + n*b_cs
  int64_t b_batch, b_rs, b_cs;
  float* C;  // element (m,n) at C + c*c_batch + m*c_rs + n
  int64_t c_batch, c_rs;
  int M, N, K;
  const float* bias;  // + bias[c*bias_batch + n]
  int64_t bias_batch;
  const float* addsrc;  // + add_scale * addsrc[c*add_batch + m*add_rs + n]
  int64_t add_batch, add_rs;
  float add_scale;
  const float* gate;  // *= gate_scale * [gate[c*gate_batch + m*gate_rs + n] > 0]
  int64_t gate_batch, gate_rs;
  float gate_scale;
  int epi;  // 0: linear; 1: relu(v*mask_a*keep_inv); 2: relu(v*mask_a*keep_inv)*mask_b*keep_inv
  float keep_inv, keep_prob;
  uint32_t keep_thr16;  // round(keep_prob * 65536): threshold of the 16-bit Bernoulli draws (set by the host)
  const uint8_t* mask_a;  // injected keep-masks [c][m][n] (tests) or nullptr -> Philox
  const uint8_t* mask_b;
  int64_t mask_batch;
  uint64_t seed;
  int64_t chain_id0;
  uint32_t eval_id, layer_a, layer_b;
  // Operand copies written by the producer (round 2).  The tensor-core GEMMs read K-major bf16 hi/lo copies of their
  // fp32 operands; k_split_operand made them in a launch of its own, twice for an activation that one consumer contracts
  // over its columns and another over its rows (10 launches, 25 % of an evaluation).  A GEMM whose result feeds another
  // GEMM now writes the copies from its epilogue, in the layout(s) the consumers need:
  //   ck_*: element (m, n) at ck + c*ck_batch + m*ck_ld + n   (K-major for a consumer that contracts over n)
  //   ct_*: element (m, n) at ct + c*ct_batch + n*ct_ld + m   (K-major for a consumer that contracts over m)
  // (lo == nullptr in single-pass mode; padding beyond N / M inside ck_ld / ct_ld is zeroed once by the owner)
  __nv_bfloat16 *ck_hi, *ck_lo, *ct_hi, *ct_lo;
  int64_t ck_batch, ck_ld, ct_batch, ct_ld;
  // ...and a consumer is handed the copies instead of the fp32 matrix: [c][rows][kp] with the contraction index
  // contiguous; tc_bgemm then skips its split of that operand
  const __nv_bfloat16 *a_hi, *a_lo, *b_hi, *b_lo;
  int64_t a_kp, b_kp;
  int64_t a_zs, b_zs;  // 0: densely packed [c][rows][kp]; else elements between two chains of a strided view
  // Sampler update applied by the epilogue (round 2, ModelBase::grad_fused_update): the tile is a slice of the gradient
  // (C = g + upd_off, addsrc = q + upd_off with add_scale = alpha/2), and instead of storing it the epilogue performs
  // what k_hmc_update (update.cu) would do with it -- kick with friction / noise, drift -- on p and q, writes the bf16
  // image of the new position into the operand mirror and, through ct_*, its transposed copy.
  int upd_on;
  int64_t upd_off;
  UpdateArgs upd;
};

__device__ __forceinline__ void split_store_pair(float v, __nv_bfloat16* hi, __nv_bfloat16* lo, int64_t o) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[o] = h;
  if (lo) lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// Dropout keep bits of units n..n+3 (n % 4 == 0) of row m: ONE Philox4x32-10 call per quad.  Word e of the output serves
// unit n+e: its low 16 bits decide the first mask of the epilogue (layer_a), its high 16 bits the second one (layer_b, the
// dropout in front of l3 that the H2 epilogue applies as well) -- P(keep) = keep_thr16 / 65536 (0.9 -> 0.899994).  Returns
// ka | kb << 4.  (Round 2, first form: one call per mask with 32-bit draws and an fp64 threshold computed per call -- ncu
// showed the H2 epilogue issue-bound on 8 Philox calls per 16 elements.)  Keyed by position, chain, evaluation and
// layer_a only: independent of tile geometry, precision path and sharding.
__device__ __forceinline__ uint32_t keep_bits8(const GemmDesc& d, int c, int m, int n) {
  const U4 r = philox4x32_10(U4{(uint32_t)(((int64_t)m * d.N + n) >> 2), (uint32_t)(d.chain_id0 + c), d.eval_id,
                                 TAG_DROPOUT | d.layer_a},
                             (uint32_t)d.seed, (uint32_t)(d.seed >> 32));
  const uint32_t t = d.keep_thr16;
  return ((r.x & 0xFFFFu) < t ? 1u : 0u) | ((r.y & 0xFFFFu) < t ? 2u : 0u) | ((r.z & 0xFFFFu) < t ? 4u : 0u) |
         ((r.w & 0xFFFFu) < t ? 8u : 0u) | ((r.x >> 16) < t ? 16u : 0u) | ((r.y >> 16) < t ? 32u : 0u) |
         ((r.z >> 16) < t ? 64u : 0u) | ((r.w >> 16) < t ? 128u : 0u);
}

// ---- the arithmetic of k_hmc_update (update.cu) for callers that hold the gradient in registers ----
__device__ __forceinline__ float4 upd_noise4(const UpdateArgs& a, int c, int64_t i) {
  const int chain = a.perm ? a.perm[c] : c;  // noise streams are keyed by chain, not by working row
  if (a.z) {
    const float* r = a.z + (int64_t)chain * a.ld_z + i;
    return make_float4(r[0], (i + 1 < a.P) ? r[1] : 0.f, (i + 2 < a.P) ? r[2] : 0.f, (i + 3 < a.P) ? r[3] : 0.f);
  }
  return philox_normal4(a.seed, a.chain_id0 + chain, (uint32_t)(i >> 2), a.stream_lo, a.stream_hi);
}
// parameters i .. i+3 (i % 4 == 0) of working row c: ge = their gradient, qe = their position (in: old, out: new).
// Returns true when the position moved (the row takes part in the pre part).
__device__ __forceinline__ bool upd_apply4(const UpdateArgs& a, int c, int64_t i, const float ge[4], float qe[4]) {
  const int Lc = a.L[c];
  const bool act_post = a.post_len > 0 && a.it_post < Lc - 1;
  const bool act_pre = a.pre_len > 0 && a.it_pre < Lc - 1;
  if (!act_post && !act_pre) return false;
  const int64_t o = (int64_t)c * a.ld + i;
  const float4 p4 = *reinterpret_cast<const float4*>(a.p + o);
  float pe[4] = {p4.x, p4.y, p4.z, p4.w};
  float ze[4] = {0.f, 0.f, 0.f, 0.f};
  if (a.n_post != 0.f && act_post) {
    const float4 z4 = upd_noise4(a, c, i);
    ze[0] = z4.x, ze[1] = z4.y, ze[2] = z4.z, ze[3] = z4.w;
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int64_t idx = i + e;
    if (act_post && idx >= a.post_off && idx < a.post_off + a.post_len) {
      float v = a.f_post * pe[e] - a.a_post * ge[e];
      if (a.n_post != 0.f) v += a.n_post * ze[e];
      pe[e] = v;
    }
    if (act_pre && idx >= a.pre_off && idx < a.pre_off + a.pre_len) {
      pe[e] = pe[e] - a.a_pre * ge[e];
      qe[e] = qe[e] + a.eps * pe[e];
    }
  }
  *reinterpret_cast<float4*>(a.p + o) = make_float4(pe[0], pe[1], pe[2], pe[3]);
  if (act_pre) {
    *reinterpret_cast<float4*>(a.q + o) = make_float4(qe[0], qe[1], qe[2], qe[3]);
    if (a.mir_hi) {
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
  return act_pre;
}
// one parameter (bias vectors, the small last layer): same arithmetic, same noise value as the float4 form
__device__ __forceinline__ void upd_apply1(const UpdateArgs& a, int c, int64_t i, float g) {
  const int Lc = a.L[c];
  const bool act_post = a.post_len > 0 && a.it_post < Lc - 1 && i >= a.post_off && i < a.post_off + a.post_len;
  const bool act_pre = a.pre_len > 0 && a.it_pre < Lc - 1 && i >= a.pre_off && i < a.pre_off + a.pre_len;
  if (!act_post && !act_pre) return;
  const int64_t o = (int64_t)c * a.ld + i;
  float pv = a.p[o];
  if (act_post) {
    float v = a.f_post * pv - a.a_post * g;
    if (a.n_post != 0.f) {
      const float4 z4 = upd_noise4(a, c, i & ~(int64_t)3);
      const int e = (int)(i & 3);
      v += a.n_post * (e == 0 ? z4.x : e == 1 ? z4.y : e == 2 ? z4.z : z4.w);
    }
    pv = v;
  }
  if (act_pre) {
    pv = pv - a.a_pre * g;
    const float qn = a.q[o] + a.eps * pv;
    a.q[o] = qn;
    if (a.mir_hi) {
      const __nv_bfloat16 h = __float2bfloat16_rn(qn);
      a.mir_hi[(int64_t)c * a.mir_ld + i] = h;
      if (a.mir_lo) a.mir_lo[(int64_t)c * a.mir_ld + i] = __float2bfloat16_rn(qn - __bfloat162float(h));
    }
  }
  a.p[o] = pv;
}
