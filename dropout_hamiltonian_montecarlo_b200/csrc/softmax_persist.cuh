// Persistent minibatch kernel: MANY SGLD / SGD steps of the softmax model in ONE launch.
// Included by softmax_tc.cu (inside namespace bhmc, after its kernels and host helpers).
//
// Reference path: sgmcmc.sample's minibatch loop (inference/cpu/sgmcmc.py:55-77) calling sgld.step (sgld.py:31-46) /
// sgd.fit's loop (sgd.py:36-41), each step = softmax.grad on a 500-row window (models/cpu/softmax.py:45-61) + update.
//
// Why: at BASELINE config 3 (128 chains, 500-row windows, 784 x 10) one step is ~2 us of tensor work per GEMM, but the
// three-launch form of round 1 (forward GEMM, backward GEMM, fused reduce/update/prep) cost 41-43 us per step: each
// launch paid its own ramp (launch latency even with programmatic dependent launch, barrier init, TMEM allocation,
// pipeline fill, tail) for 13 / 9 K-chunks of work, and the split-K reduce existed only to occupy more SMs.
// Here one cooperative launch of one CTA per SM runs a whole epoch:
//   phase F  forward items (row tile x chain tile): TMA -> tcgen05.mma -> softmax epilogue -> (P-Y)^T hi/lo  (as k_tc_gemm)
//   grid barrier
//   phase B  backward items (feature tile x chain tile), NO split-K: the item owns its [128 features x cpt chains]
//            block of the gradient, so its epilogue applies the parameter update directly from tensor memory
//            (g = acc/s + alpha q; SGLD: p = 2 eps z - eps/2 g, q += p with Philox or injected z; SGD: heavy ball)
//            and writes the bf16 hi/lo operand copy of the new weights for the next step's forward pass
//   grid barrier
// No gradient, no split-K partials and no per-step operand preparation ever reach HBM; mbarrier rings, TMEM and the
// warp roles live across steps.  Data written with ordinary stores in one phase ((P-Y)^T, W^T operand) and read by
// TMA (async proxy) in the next is ordered by fence.proxy.async on both sides of the grid barrier.

struct PersistParams {
  int D, K, C, cpt, BN, n_tiles;  // chains per N tile, UMMA N, chain tiles
  int k_chunks_f;                 // BK-chunks of the feature dimension
  int m_tiles_b;                  // 128-row tiles of the D+1 gradient rows (row D = bias)
  int split3, stages;
  int64_t ld, Dp;
  float* q;
  float* p;
  const int32_t* labels;          // label of bound row 0
  int64_t batch, row_first;       // step j works on rows [row_first + j*batch, +batch)
  int n_steps;
  const float* eps;               // [n_steps] step size of every step (the schedule is the host's, sgmcmc.py:72-73)
  float gamma, alpha, scale;      // scale = 1 / (scale of the bound X operand)
  int kind;                       // BHMC_KIND_SGLD / BHMC_KIND_SGD
  const float* z;                 // injected N(0,1) tape [n_steps][C, ld_z] or nullptr (Philox)
  int64_t ld_z, z_step_stride;
  uint64_t seed;
  int64_t chain_id0;
  uint64_t step0;                 // global index of step 0 (Philox stream)
  __nv_bfloat16 *wt_hi, *wt_lo, *dmt_hi, *dmt_lo;
  int dm_rows;                    // rows per (P-Y)^T slab = n_tiles * BN
  int xt_rows;                    // rows per X^T slab (Dt_pad)
  unsigned int* bar;              // grid barrier counter (zeroed before the launch)
  unsigned int* flags;            // [2][n_tiles] monotonic item counters per chain tile (forward / backward), zeroed before the
                                  // launch; nullptr = grid-wide barriers between the phases (BHMC_PERSIST_FLAGS=0)
  int two_cta;                    // k_sg_persistent2: cta_group::2 items (needs the chain-tile flags)
  int half_f;                     // forward items are HALF row tiles (64 rows of the window): twice the items, 36 KB instead of
                                  // 52 KB per chunk and item; the MMA stays M = 128, the upper accumulator half is ignored
  int pair;                       // clusters of two CTAs work on the same row / feature tile and adjacent chain tiles: each
                                  // fetches half of the shared X (X^T) tile and multicasts it to both (BHMC_PERSIST_PAIR)
  int prefetch;                   // L2 prefetch of the next phase's X window (BHMC_PERSIST_PF, default on)
  long long* prof;                // optional [grid][8] cycle counters of one epilogue thread (BHMC_PROF=1): per step
                                  // phase F work, wait at barrier 1, phase B work, wait at barrier 2 (sums over the steps)
};

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Grid-wide barrier of a cooperative launch (every CTA resident).  Monotonic counter: the n-th barrier waits for
// n * gridDim.x arrivals.  Bounded spin: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int target) {
  fence_proxy_async_all();  // this thread's ordinary stores must be visible to TMA reads issued after the barrier
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    const long long t0 = clock64();
    while (true) {
      unsigned int v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (v >= target) break;
      if (clock64() - t0 > 4000000000LL) {
        printf("bhmc: grid barrier timed out (block %d, %u of %u)\n", blockIdx.x, v, target);
        __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
  fence_proxy_async_all();
}

// Per-chain-tile dependencies instead of grid-wide barriers (round 2).  Chain tiles never exchange data: the backward items
// of chain tile nt need the (P-Y)^T rows of ITS forward items only, and the next step's forward items of nt need the W^T
// rows / bias of ITS backward items only.  Every finished item bumps a counter of its chain tile; the TMA producer of a
// dependent item waits for the count of the step.  Same fences as the grid barrier (generic stores -> TMA reads).
// One counter per 128-byte line: with the 32 counters of a launch in ONE line every polling producer and every signalling
// atomic met in the same L2 slot, and single runs took 57-84 us per step instead of 24.5 (measured).  The poll backs off.
static constexpr int FLAG_PAD = 32;
__device__ __forceinline__ void flag_wait(const unsigned int* f, unsigned int target) {  // one thread
  const long long t0 = clock64();
  while (true) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    if (v >= target) break;
    __nanosleep(40);
    if (clock64() - t0 > 4000000000LL) {
      printf("bhmc: chain-tile flag timed out (block %d, %u of %u)\n", blockIdx.x, v, target);
      __trap();
    }
  }
  fence_proxy_async_all();
}
// all 32 * EW epilogue threads of the CTA call it after the item's last store
template <int EW>
__device__ __forceinline__ void flag_signal(unsigned int* f) {
  fence_proxy_async_all();
  asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
  if (threadIdx.x == NON_EPI_THREADS) {
    __threadfence();  // (red.release.gpu instead of fence + relaxed atomic: measured identical, 24.4-24.7 us per step)
    atomicAdd(f, 1u);
  }
}
// (A per-warp form -- every epilogue warp fences and bumps the counter itself, no CTA-wide barrier in front of the release --
// was measured: the signals cost the same 5.2 k cycles per step and the step got slower and erratic, 26-52 us against
// 24.5: sixteen times the atomics on the line the consumers poll.)

// Update epilogue of one backward tile.  Tensor memory hands a thread one gradient ROW (feature d; d == D is the bias
// row) with the KP classes of a chain, while a chain's parameters are stored row-major (i = d*K + k): a warp's 32 rows
// are ONE contiguous block of 32*KP floats.  The first version updated q / p straight from the row layout -- every
// load / store instruction of a warp touched 32 addresses 4*KP bytes apart, ~340 L1 wavefronts per (warp, chain), and
// the phase took 44 k cycles per step against 20 k for the (nearly empty) last row tile (in-kernel timers).  Now the
// gradient rows go through a per-warp shared-memory block and the update runs in the memory order: one float4 per lane
// and instruction (coalesced q / p traffic, one Philox block per float4 -- the keys of k_sgld, element for element);
// the new weights go back through the same block for the K-major bf16 operand copy, which IS row-ordered.
// The noise (SGLD) / momentum (SGD) operand of the update does not depend on the gradient: it is produced while the
// epilogue warps would otherwise wait for the accumulator (Philox + Box-Muller were ~4 k of the ~10 k cycles of a
// backward item's epilogue, on the critical path of every step).  Slot [ci][fi]: chain part + ci*PARTS of the tile,
// float4 index 32*fi + lane of the warp's block -- the loops of sg_update_tile.
template <int KP>
struct SgPre {
  static constexpr int NF = (32 * KP / 4 + 31) / 32;
  float4 z[2][NF];  // at most two chains per warp (cpt <= 8 with EW = 16)
};
template <int KP, int EW, int KIND>
__device__ __forceinline__ void sg_update_pre(const PersistParams& p, int mt, int nt, int part, int ew, int lane, int step,
                                              SgPre<KP>& pre) {
  constexpr int PARTS = EW / 4;
  constexpr int NV = 32 * KP / 4;
  const int64_t P = (int64_t)(p.D + 1) * KP;
  const int64_t i0w = (int64_t)(mt * BM + ew * 32) * KP;
  const uint64_t gstep = p.step0 + (uint64_t)step;
  const uint32_t slo = (uint32_t)gstep, shi = TAG_NOISE | (uint32_t)((gstep >> 32) & 0xffffff);
#pragma unroll
  for (int ci = 0; ci < 2; ++ci) {
    const int cc = part + ci * PARTS, c = nt * p.cpt + cc;
#pragma unroll
    for (int fi = 0; fi < SgPre<KP>::NF; ++fi) {
      const int f = 32 * fi + lane;
      const int64_t gi = i0w + 4 * f;
      float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cc < p.cpt && c < p.C && f < NV && gi < P) {
        if (KIND == BHMC_KIND_SGLD) {
          if (p.z) {
            const float* zr = p.z + (int64_t)step * p.z_step_stride + (int64_t)c * p.ld_z + gi;
            z4.x = __ldcs(zr);
            z4.y = gi + 1 < P ? __ldcs(zr + 1) : 0.f;
            z4.z = gi + 2 < P ? __ldcs(zr + 2) : 0.f;
            z4.w = gi + 3 < P ? __ldcs(zr + 3) : 0.f;
          } else {
            z4 = philox_normal4(p.seed, p.chain_id0 + c, (uint32_t)(gi >> 2), slo, shi);
          }
        } else {
          z4 = __ldcg(reinterpret_cast<const float4*>(p.p + (int64_t)c * p.ld + gi));  // heavy-ball momentum (this thread wrote it)
        }
      }
      pre.z[ci][fi] = z4;
    }
  }
}

template <int KP, int EW, int KIND, bool PRE = false>
__device__ __forceinline__ void sg_update_tile(const PersistParams& p, uint32_t tacc, int mt, int nt, int part, int ew, int lane,
                                               int step, float eps, float* S, const SgPre<KP>* pre = nullptr) {
  constexpr int PARTS = EW / 4;
  constexpr int NV = 32 * KP / 4;  // float4s of a warp's block
  const int d = mt * BM + ew * 32 + lane;
  const int64_t P = (int64_t)(p.D + 1) * KP;             // KP == K (exact class count)
  const int64_t i0w = (int64_t)(mt * BM + ew * 32) * KP;  // first parameter of the warp's block (multiple of 32)
  const uint64_t gstep = p.step0 + (uint64_t)step;
  const uint32_t slo = (uint32_t)gstep, shi = TAG_NOISE | (uint32_t)((gstep >> 32) & 0xffffff);
#pragma unroll 2
  for (int ci = 0; ci < (PRE ? 2 : 64); ++ci) {
    const int cc = part + ci * PARTS;
    if (cc >= p.cpt) break;
    const int c = nt * p.cpt + cc;
    if (c >= p.C) break;  // warp-uniform
    uint32_t raw[KP];
    tmem_ld_cols<KP>(tacc + (uint32_t)(cc * KP), raw);
    tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < KP; ++k) S[lane * KP + k] = __uint_as_float(raw[k]) * p.scale;  // X^T (P - Y), row layout
    __syncwarp();
    float* qc = p.q + (int64_t)c * p.ld;
    float* pc = p.p + (int64_t)c * p.ld;
#pragma unroll
    for (int f0 = 0; f0 < NV; f0 += 32) {
      const int f = f0 + lane;
      const int64_t gi = i0w + 4 * f;
      if (f < NV && gi < P) {
        const float4 q4 = __ldcg(reinterpret_cast<const float4*>(qc + gi));  // L2: another SM wrote it one step ago
        const float4 g4 = *reinterpret_cast<const float4*>(S + 4 * f);
        float4 z4;
        if (PRE) {
          z4 = pre->z[ci < 2 ? ci : 0][f0 / 32];
        } else if (KIND == BHMC_KIND_SGLD) {
          if (p.z) {
            const float* zr = p.z + (int64_t)step * p.z_step_stride + (int64_t)c * p.ld_z + gi;
            z4.x = __ldcs(zr);
            z4.y = gi + 1 < P ? __ldcs(zr + 1) : 0.f;
            z4.z = gi + 2 < P ? __ldcs(zr + 2) : 0.f;
            z4.w = gi + 3 < P ? __ldcs(zr + 3) : 0.f;
          } else {
            z4 = philox_normal4(p.seed, p.chain_id0 + c, (uint32_t)(gi >> 2), slo, shi);
          }
        } else {
          z4 = __ldcg(reinterpret_cast<const float4*>(pc + gi));  // heavy-ball momentum
        }
        const float qe[4] = {q4.x, q4.y, q4.z, q4.w}, ge[4] = {g4.x, g4.y, g4.z, g4.w}, ze[4] = {z4.x, z4.y, z4.z, z4.w};
        float pe[4], qn[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const float g = fmaf(p.alpha, qe[x], ge[x]);  // + alpha q
          pe[x] = KIND == BHMC_KIND_SGLD ? (2.0f * eps) * ze[x] - (0.5f * eps) * g   // sgld.py:31-46
                                         : p.gamma * ze[x] - eps * g;                // sgd.py:40
          if (gi + x >= P) pe[x] = 0.f;  // padding of the parameter row
          qn[x] = qe[x] + pe[x];
        }
        *reinterpret_cast<float4*>(pc + gi) = make_float4(pe[0], pe[1], pe[2], pe[3]);
        *reinterpret_cast<float4*>(qc + gi) = make_float4(qn[0], qn[1], qn[2], qn[3]);
        *reinterpret_cast<float4*>(S + 4 * f) = make_float4(qn[0], qn[1], qn[2], qn[3]);
      }
    }
    __syncwarp();
    if (d < p.D) {  // operand copy of the new weights for the next forward pass (the bias is added in the epilogue)
      __nv_bfloat16* wh = p.wt_hi + ((int64_t)c * KP) * p.Dp + d;
      __nv_bfloat16* wl = p.wt_lo ? p.wt_lo + ((int64_t)c * KP) * p.Dp + d : nullptr;
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        __nv_bfloat16 hb, lb;
        split_bf16(p.scale * S[lane * KP + k], hb, lb);
        wh[(int64_t)k * p.Dp] = hb;
        if (wl) wl[(int64_t)k * p.Dp] = lb;
      }
    }
    __syncwarp();  // the block is reused by the warp's next chain
  }
}

template <int KP, int EW>
__global__ void __launch_bounds__(NON_EPI_THREADS + 32 * EW, 1)
k_sg_persistent(const __grid_constant__ CUtensorMap tmXa_hi, const __grid_constant__ CUtensorMap tmXa_lo,
                const __grid_constant__ CUtensorMap tmWt_hi, const __grid_constant__ CUtensorMap tmWt_lo,
                const __grid_constant__ CUtensorMap tmXt_hi, const __grid_constant__ CUtensorMap tmXt_lo,
                const __grid_constant__ CUtensorMap tmDm_hi, const __grid_constant__ CUtensorMap tmDm_lo, const PersistParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_base_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a_bytes = BM * BK * 2, b_bytes = p.BN * BK * 2;
  const int na = p.split3 == 1 ? 2 : 1, nb = p.split3 ? 2 : 1;
  const int stage_bytes = na * a_bytes + nb * b_bytes;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), p.pair ? 2 : 1);  // pair: the peer's producer writes into this slot as well
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
  tcgen05_fence_before();
  __syncthreads();
  if (p.pair) cluster_sync_all();  // peer barriers must be initialised before anything is multicast into them
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  const int rank = p.pair ? (int)cluster_ctarank() : 0;

  // ring / accumulator bookkeeping (lives across phases and steps); every role has its own variables.
  // What bounds a phase (in-kernel counters, BHMC_PROF=1): the MMA warp of phase F spends 15.6 k cycles per item and waits
  // only 4.3 k of them for operands -- 11.3 k cycles for 156 MMAs = 72 cycles per MMA where the tensor core needs 40
  // (M = 128, N = 80, K = 16).  Both operands come from shared memory: 4 KB of A + 2.5 KB of B per MMA at 128 B/clk is 52
  // cycles, and the TMA writes of the next chunks (52 KB = 406 cycles per chunk) share the port.  Narrow-N bf16x3 items
  // are bound by the shared-memory port, not by HBM, L2, the request path (pair multicast: no gain) or the bytes per
  // chunk (half tiles: no gain).
  int p_stage = 0;      // producer (warp 0, lane 0)
  uint32_t p_phase = 0;
  int m_stage = 0;      // MMA issuer (warp 1, warp-uniform)
  uint32_t m_phase = 0;
  int m_it = 0;
  int it = 0;           // epilogue warps
  unsigned int n_bar = 0;

  // one work item of either phase, seen from the three roles
  auto produce = [&](const CUtensorMap* a_hi, const CUtensorMap* a_lo, const CUtensorMap* b_hi, const CUtensorMap* b_lo,
                     int n_chunks, int a_k0, int a_dk, int a_m0, int a_dm, int b_k0, int b_dk, int b_n0, int b_dn,
                     int a_box_bytes) {
    // chunk k: A box at (a_k0 + k*a_dk, a_m0 + k*a_dm), B box at (b_k0 + k*b_dk, b_n0 + k*b_dn); a_box_bytes < a_bytes: the
    // map's box is a half tile (rows 64..127 of the slot keep whatever they held -- their accumulator rows are ignored)
    for (int k = 0; k < n_chunks; ++k) {
      mbar_wait(smem_u32(&bar_empty[p_stage]), p_phase ^ 1u);
      const uint32_t full = smem_u32(&bar_full[p_stage]);
      mbar_expect_tx(full, (uint32_t)(na * a_box_bytes + nb * b_bytes));
      const uint32_t sa = smem_base + p_stage * stage_bytes, sb = sa + na * a_bytes;
      if (p.pair) {  // the A tile is the same for both CTAs of the cluster: fetch 64 of its 128 rows for both
        const uint32_t off = (uint32_t)rank * (BM / 2) * (BK * 2);
        tma_load_2d_mc(sa + off, a_hi, full, a_k0 + k * a_dk, a_m0 + k * a_dm + rank * (BM / 2), 3);
        if (na == 2) tma_load_2d_mc(sa + a_bytes + off, a_lo, full, a_k0 + k * a_dk, a_m0 + k * a_dm + rank * (BM / 2), 3);
      } else {
        tma_load_2d(sa, a_hi, full, a_k0 + k * a_dk, a_m0 + k * a_dm);
        if (na == 2) tma_load_2d(sa + a_bytes, a_lo, full, a_k0 + k * a_dk, a_m0 + k * a_dm);
      }
      tma_load_2d(sb, b_hi, full, b_k0 + k * b_dk, b_n0 + k * b_dn);
      if (p.split3) tma_load_2d(sb + b_bytes, b_lo, full, b_k0 + k * b_dk, b_n0 + k * b_dn);
      if (++p_stage == p.stages) p_stage = 0, p_phase ^= 1u;
    }
  };
  long long mma_wait[2] = {0, 0}, mma_all[2] = {0, 0};  // BHMC_PROF: cycles the MMA warp waits for operands / spends per phase
  int which_phase = 0;
  auto issue = [&](int n_chunks) {  // whole warp, warp-uniform; the elected lane issues
    const int buf = m_it & 1;
    const uint32_t use = (uint32_t)(m_it >> 1);
    const long long t_in = p.prof ? clock64() : 0;
    mbar_wait(smem_u32(&bar_tempty[buf]), (use & 1u) ^ 1u);
    tcgen05_fence_after();
    const uint32_t tmem_d = tmem_base + (uint32_t)(buf * TMEM_BUF_COLS);
    for (int k = 0; k < n_chunks; ++k) {
      const long long t_w = p.prof ? clock64() : 0;
      mbar_wait(smem_u32(&bar_full[m_stage]), m_phase);
      if (p.prof) mma_wait[which_phase] += clock64() - t_w;
      tcgen05_fence_after();
      const uint32_t sa = smem_base + m_stage * stage_bytes;
      const uint32_t first = k > 0 ? 1u : 0u;
      if (p.split3 == 1) {
        const uint64_t a_hi = make_smem_desc(sa), a_lo = make_smem_desc(sa + a_bytes);
        const uint64_t b_hi = make_smem_desc(sa + 2 * a_bytes), b_lo = make_smem_desc(sa + 2 * a_bytes + b_bytes);
#pragma unroll
        for (int ks = 0; ks < BK / UMMA_K; ++ks) {
          const uint64_t adv = (uint64_t)((ks * UMMA_K * 2) >> 4);
          umma_bf16(tmem_d, a_hi + adv, b_hi + adv, idesc, ks > 0 ? 1u : first);
          umma_bf16(tmem_d, a_hi + adv, b_lo + adv, idesc, 1u);
          umma_bf16(tmem_d, a_lo + adv, b_hi + adv, idesc, 1u);
        }
      } else if (p.split3 == 2) {  // X exact in bf16: no lo copy of the X operand
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
      if (p.pair) umma_commit_mc(smem_u32(&bar_empty[m_stage]), 3);  // frees the slot in both CTAs
      else umma_commit(smem_u32(&bar_empty[m_stage]));
      if (++m_stage == p.stages) m_stage = 0, m_phase ^= 1u;
    }
    umma_commit(smem_u32(&bar_tfull[buf]));
    if (p.prof) mma_all[which_phase] += clock64() - t_in;
    ++m_it;
  };

  const int ew = warp & 3, part = (warp - 4) >> 2, t = ew * 32 + lane;
  // per-warp block of 32 x KP floats behind the stage ring (update epilogue)
  float* epi_S = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + (size_t)p.stages * stage_bytes) +
                 (warp >= 4 ? (warp - 4) * 32 * KP : 0);
  const int m_tiles_f = (int)((p.batch + BM - 1) / BM);
  const int f_rows = p.half_f ? BM / 2 : BM;  // window rows of a forward item
  const int items_f = (p.half_f ? 2 : 1) * m_tiles_f * p.n_tiles, items_b = p.m_tiles_b * p.n_tiles;
  // forward-epilogue parameters that do not change from step to step
  TcParams pf{};
  pf.K = p.K, pf.C = p.C, pf.cpt = p.cpt, pf.D = p.D, pf.ld = p.ld, pf.q = p.q, pf.nrows = p.batch;
  pf.dm_slab = BK, pf.dm_slab_rows = p.dm_rows, pf.dm_ld = BK, pf.dmt_hi = p.dmt_hi, pf.dmt_lo = p.dmt_lo;
  pf.split3 = p.split3, pf.write_dm = 1, pf.m_tiles = m_tiles_f, pf.skip_loglik = 1;

  const bool timer = p.prof && threadIdx.x == NON_EPI_THREADS;  // first epilogue thread
  long long tF = 0, tW1 = 0, tB = 0, tW2 = 0, t_prev = timer ? clock64() : 0;
  for (int step = 0; step < p.n_steps; ++step) {
    const int64_t row0 = p.row_first + (int64_t)step * p.batch;
    const int shift = (int)(row0 % BK);
    const int k_chunks_b = (int)((p.batch + shift + BK - 1) / BK);
    // ---------------- phase F: Z = X_window . W^T, softmax, (P - Y)^T ----------------
    if (warp == 0) {
      if (lane == 0) {
        // Every step works on a NEW window of X: its tiles come from HBM in 128-byte pieces of 128 different rows, and
        // with 13 / 9 dependent chunks per item and 3 stages the phases were bound by that latency (in-kernel timers:
        // 19 k / 18 k cycles per phase for ~5 k cycles of MMA work).  The window of the next phase is known one phase
        // ahead, so one CTA per row tile asks for it in L2 while the current phase computes (BHMC_PERSIST_PF=0: off).
        if (p.prefetch)
          for (int w = blockIdx.x; w < items_b; w += gridDim.x)
            if (w % p.n_tiles == 0) {  // X^T window of this step's backward phase
              const int slab0 = (int)((row0 - shift) / BK), mt = w / p.n_tiles;
              for (int k = 0; k < k_chunks_b; ++k)
                for (int h = 0; h <= p.pair; ++h) {  // pair mode: the boxes are 64 rows high
                  tma_prefetch_2d(&tmXt_hi, 0, (slab0 + k) * p.xt_rows + mt * BM + h * (BM / 2));
                  if (na == 2) tma_prefetch_2d(&tmXt_lo, 0, (slab0 + k) * p.xt_rows + mt * BM + h * (BM / 2));
                }
            }
        for (int w = blockIdx.x; w < items_f; w += gridDim.x) {
          const int mt = w / p.n_tiles, nt = w % p.n_tiles;  // half_f: mt counts half tiles
          if (p.flags && step > 0) flag_wait(p.flags + FLAG_PAD * (p.n_tiles + nt), (unsigned int)(p.m_tiles_b * step));  // W^T / bias of step - 1
          produce(&tmXa_hi, &tmXa_lo, &tmWt_hi, &tmWt_lo, p.k_chunks_f, 0, BK, (int)row0 + mt * f_rows, 0, 0, BK, nt * p.BN, 0,
                  p.half_f ? a_bytes / 2 : a_bytes);
        }
      }
    } else if (warp == 1) {
      which_phase = 0;
      for (int w = blockIdx.x; w < items_f; w += gridDim.x) issue(p.k_chunks_f);
    } else if (warp >= 4) {
      pf.labels = p.labels + row0;
      pf.dm_shift = shift;
      pf.pack_dm = 0;  // (row-pair stores of (P-Y)^T: BHMC_FWD_PACK in the per-launch path, measured neutral)
      pf.dm_tail = (BK - shift) % BK;
      for (int w = blockIdx.x; w < items_f; w += gridDim.x, ++it) {
        const int mt = w / p.n_tiles, nt = w % p.n_tiles;
        const int buf = it & 1;
        mbar_wait(smem_u32(&bar_tfull[buf]), (uint32_t)(it >> 1) & 1u);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * TMEM_BUF_COLS);
        if (p.half_f) fwd_epilogue_tile<KP, EW, true>(pf, tacc, mt >> 1, nt, part, lane, t, (mt & 1) * (BM / 2), BM / 2);
        else fwd_epilogue_tile<KP, EW, true>(pf, tacc, mt, nt, part, lane, t);
        tcgen05_fence_before();
        mbar_arrive(smem_u32(&bar_tempty[buf]));
        if (p.flags) flag_signal<EW>(p.flags + FLAG_PAD * nt);
      }
    }
    if (timer) { const long long t = clock64(); tF += t - t_prev; t_prev = t; }
    if (!p.flags) grid_barrier(p.bar, ++n_bar * gridDim.x);
    if (timer) { const long long t = clock64(); tW1 += t - t_prev; t_prev = t; }
    // ---------------- phase B: G = X_window^T (P - Y), update, next W^T operand ----------------
    if (warp == 0) {
      if (lane == 0) {
        const int slab0 = (int)((row0 - shift) / BK);  // X^T is stored in blocks of BK rows of X: [slab][xt_rows][BK]
        if (p.prefetch && step + 1 < p.n_steps)
          for (int w = blockIdx.x; w < items_f; w += gridDim.x)
            if (w % p.n_tiles == 0) {  // X window of the next step's forward phase
              const int mt = w / p.n_tiles;
              for (int k = 0; k < p.k_chunks_f; ++k)
                for (int h = 0; h <= p.pair; ++h) {
                  tma_prefetch_2d(&tmXa_hi, k * BK, (int)(row0 + p.batch) + mt * f_rows + h * (BM / 2));
                  if (na == 2) tma_prefetch_2d(&tmXa_lo, k * BK, (int)(row0 + p.batch) + mt * f_rows + h * (BM / 2));
                }
            }
        for (int w = blockIdx.x; w < items_b; w += gridDim.x) {
          const int mt = w / p.n_tiles, nt = w % p.n_tiles;
          if (p.flags) flag_wait(p.flags + FLAG_PAD * nt, (unsigned int)((items_f / p.n_tiles) * (step + 1)));  // (P-Y)^T of this step
          produce(&tmXt_hi, &tmXt_lo, &tmDm_hi, &tmDm_lo, k_chunks_b, 0, 0, slab0 * p.xt_rows + mt * BM, p.xt_rows, 0, 0,
                  nt * p.BN, p.dm_rows, a_bytes);
        }
      }
    } else if (warp == 1) {
      which_phase = 1;
      for (int w = blockIdx.x; w < items_b; w += gridDim.x) issue(k_chunks_b);
    } else if (warp >= 4) {
      const float eps = p.eps[step];
      for (int w = blockIdx.x; w < items_b; w += gridDim.x, ++it) {
        const int mt = w / p.n_tiles, nt = w % p.n_tiles;
        const int buf = it & 1;
        mbar_wait(smem_u32(&bar_tfull[buf]), (uint32_t)(it >> 1) & 1u);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * TMEM_BUF_COLS);
        if (p.kind == BHMC_KIND_SGLD) sg_update_tile<KP, EW, BHMC_KIND_SGLD>(p, tacc, mt, nt, part, ew, lane, step, eps, epi_S);
        else sg_update_tile<KP, EW, BHMC_KIND_SGD>(p, tacc, mt, nt, part, ew, lane, step, eps, epi_S);
        tcgen05_fence_before();
        mbar_arrive(smem_u32(&bar_tempty[buf]));
        if (p.flags) flag_signal<EW>(p.flags + FLAG_PAD * (p.n_tiles + nt));
      }
    }
    if (timer) { const long long t = clock64(); tB += t - t_prev; t_prev = t; }
    if (!p.flags) grid_barrier(p.bar, ++n_bar * gridDim.x);
    if (timer) { const long long t = clock64(); tW2 += t - t_prev; t_prev = t; }
  }
  if (timer) {
    long long* o = p.prof + (size_t)blockIdx.x * 8;
    o[0] = tF, o[1] = tW1, o[2] = tB, o[3] = tW2, o[4] = p.n_steps;
  }
  if (p.prof && warp == 1 && lane == 0) {  // MMA warp: operand waits / total per phase, packed (wait << 32 | total), in k-cycles... plain sums
    long long* o = p.prof + (size_t)blockIdx.x * 8;
    o[5] = mma_wait[0], o[6] = mma_all[0], o[7] = mma_wait[1];
  }
  tcgen05_fence_before();
  __syncthreads();
  if (p.pair) cluster_sync_all();  // the peer's last commits still arrive on this CTA's barriers
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------
// The same epoch loop with cta_group::2 items (round 2, BHMC_PERSIST_2CTA).  The phases of k_sg_persistent are bound by
// the shared-memory port (72 cycles per M = 128, N = 80 MMA: 4 KB of A + 2.5 KB of B read per MMA, plus the TMA writes of
// the next chunks).  Here a cluster of two CTAs owns TWO adjacent row tiles (phase F) / feature tiles (phase B) of one
// chain tile: the leader's elected thread issues tcgen05.mma.cta_group::2 (M = 256), every CTA stages its own A tile and
// only HALF of the chain-side operand (40 of 80 rows) -- 5.25 KB read per MMA and CTA, 42 KB staged per chunk instead
// of 52 KB.  An odd tile count gets a phantom tile (phase B: rows 896.. of X^T -- the next slab's rows or TMA zero fill --
// whose gradient rows lie beyond the parameters and are skipped by the update).  Barrier protocol as in k_tc_fwd2;
// dependencies between the phases through the chain-tile flags only (every CTA, phantom ones included, counts).
template <int KP, int EW>
__global__ void __launch_bounds__(NON_EPI_THREADS + 32 * EW, 1)
k_sg_persistent2(const __grid_constant__ CUtensorMap tmXa_hi, const __grid_constant__ CUtensorMap tmXa_lo,
                 const __grid_constant__ CUtensorMap tmWt_hi, const __grid_constant__ CUtensorMap tmWt_lo,
                 const __grid_constant__ CUtensorMap tmXt_hi, const __grid_constant__ CUtensorMap tmXt_lo,
                 const __grid_constant__ CUtensorMap tmDm_hi, const __grid_constant__ CUtensorMap tmDm_lo, const PersistParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_base_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const bool leader = rank == 0;
  const int a_bytes = BM * BK * 2, bh_bytes = (p.BN / 2) * BK * 2;  // this CTA's half of the chain-side operand
  const int na = p.split3 == 1 ? 2 : 1, nb = p.split3 ? 2 : 1;
  const int stage_bytes = na * a_bytes + nb * bh_bytes;
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
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

  int p_stage = 0;  // producer
  uint32_t p_phase = 0;
  int m_stage = 0;  // MMA issuer (leader)
  uint32_t m_phase = 0;
  int m_it = 0;
  int it = 0;       // epilogue warps

  auto produce = [&](const CUtensorMap* a_hi, const CUtensorMap* a_lo, const CUtensorMap* b_hi, const CUtensorMap* b_lo,
                     int n_chunks, int a_k0, int a_dk, int a_m0, int a_dm, int b_k0, int b_dk, int b_n0, int b_dn) {
    for (int k = 0; k < n_chunks; ++k) {
      mbar_wait(smem_u32(&bar_empty[p_stage]), p_phase ^ 1u);
      const uint32_t full = smem_u32(&bar_full[p_stage]);  // same offset in the leader CTA
      if (leader) mbar_expect_tx(full, (uint32_t)(2 * stage_bytes));
      const uint32_t sa = smem_base + p_stage * stage_bytes, sb = sa + na * a_bytes;
      tma_load_2d_2sm(sa, a_hi, full, a_k0 + k * a_dk, a_m0 + k * a_dm);
      if (na == 2) tma_load_2d_2sm(sa + a_bytes, a_lo, full, a_k0 + k * a_dk, a_m0 + k * a_dm);
      tma_load_2d_2sm(sb, b_hi, full, b_k0 + k * b_dk, b_n0 + k * b_dn);
      if (p.split3) tma_load_2d_2sm(sb + bh_bytes, b_lo, full, b_k0 + k * b_dk, b_n0 + k * b_dn);
      if (++p_stage == p.stages) p_stage = 0, p_phase ^= 1u;
    }
  };
  auto issue = [&](int n_chunks) {  // leader's warp 1, warp-uniform; the elected lane issues
    const int buf = m_it & 1;
    const uint32_t use = (uint32_t)(m_it >> 1);
    mbar_wait(smem_u32(&bar_tempty[buf]), (use & 1u) ^ 1u);  // both CTAs' epilogues have drained this accumulator
    tcgen05_fence_after();
    const uint32_t tmem_d = tmem_base + (uint32_t)(buf * TMEM_BUF_COLS);
    for (int k = 0; k < n_chunks; ++k) {
      mbar_wait(smem_u32(&bar_full[m_stage]), m_phase);
      tcgen05_fence_after();
      const uint32_t sa = smem_base + m_stage * stage_bytes;
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
      } else if (p.split3 == 2) {  // X exact in bf16: no lo copy of the X operand
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
      umma_commit_2sm(smem_u32(&bar_empty[m_stage]), 3);  // frees the stage in both CTAs
      if (++m_stage == p.stages) m_stage = 0, m_phase ^= 1u;
    }
    umma_commit_2sm(smem_u32(&bar_tfull[buf]), 3);  // accumulator halves complete in both CTAs
    ++m_it;
  };

  const int ew = warp & 3, part = (warp - 4) >> 2, t = ew * 32 + lane;
  float* epi_S = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + (size_t)p.stages * stage_bytes) +
                 (warp >= 4 ? (warp - 4) * 32 * KP : 0);
  const int m_tiles_f = (int)((p.batch + BM - 1) / BM);
  const int pairs_f = (m_tiles_f + 1) / 2, pairs_b = (p.m_tiles_b + 1) / 2;
  const int items_f = pairs_f * p.n_tiles, items_b = pairs_b * p.n_tiles;  // pair items
  const int wi0 = blockIdx.x / 2, wi_step = gridDim.x / 2;
  TcParams pf{};
  pf.K = p.K, pf.C = p.C, pf.cpt = p.cpt, pf.D = p.D, pf.ld = p.ld, pf.q = p.q, pf.nrows = p.batch;
  pf.dm_slab = BK, pf.dm_slab_rows = p.dm_rows, pf.dm_ld = BK, pf.dmt_hi = p.dmt_hi, pf.dmt_lo = p.dmt_lo;
  pf.split3 = p.split3, pf.write_dm = 1, pf.m_tiles = m_tiles_f, pf.skip_loglik = 1;

  // BHMC_PROF=1: cycles of the first epilogue thread per segment (sums over the steps) and of the producer's flag waits
  const bool timer = p.prof && threadIdx.x == NON_EPI_THREADS;
  long long tp[7] = {0, 0, 0, 0, 0, 0, 0}, t_prev = timer ? clock64() : 0, t_flag = 0;
  auto lap = [&](int i) {
    if (timer) {
      const long long t = clock64();
      tp[i] += t - t_prev;
      t_prev = t;
    }
  };
  for (int step = 0; step < p.n_steps; ++step) {
    const int64_t row0 = p.row_first + (int64_t)step * p.batch;
    const int shift = (int)(row0 % BK);
    const int k_chunks_b = (int)((p.batch + shift + BK - 1) / BK);
    // ---------------- phase F ----------------
    if (warp == 0) {
      if (lane == 0)
        for (int w = wi0; w < items_f; w += wi_step) {
          const int mt = 2 * (w / p.n_tiles) + rank, nt = w % p.n_tiles;
          const long long tf0 = p.prof ? clock64() : 0;
          if (step > 0) flag_wait(p.flags + FLAG_PAD * (p.n_tiles + nt), (unsigned int)(2 * pairs_b * step));  // W^T / bias of step - 1
          if (p.prof) t_flag += clock64() - tf0;
          produce(&tmXa_hi, &tmXa_lo, &tmWt_hi, &tmWt_lo, p.k_chunks_f, 0, BK, (int)row0 + mt * BM, 0, 0, BK,
                  nt * p.BN + rank * (p.BN / 2), 0);
        }
    } else if (warp == 1) {
      if (leader)
        for (int w = wi0; w < items_f; w += wi_step) issue(p.k_chunks_f);
    } else if (warp >= 4) {
      pf.labels = p.labels + row0;
      pf.dm_shift = shift;
      pf.dm_tail = (BK - shift) % BK;
      for (int w = wi0; w < items_f; w += wi_step, ++it) {
        const int mt = 2 * (w / p.n_tiles) + rank, nt = w % p.n_tiles;
        const int buf = it & 1;
        lap(6);
        mbar_wait(smem_u32(&bar_tfull[buf]), (uint32_t)(it >> 1) & 1u);
        lap(0);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * TMEM_BUF_COLS);
        if (mt < m_tiles_f) fwd_epilogue_tile<KP, EW, true>(pf, tacc, mt, nt, part, lane, t);
        tcgen05_fence_before();
        if (leader) mbar_arrive(smem_u32(&bar_tempty[buf]));
        else mbar_arrive_remote(smem_u32(&bar_tempty[buf]), 0);
        lap(1);
        flag_signal<EW>(p.flags + FLAG_PAD * nt);
        lap(2);
      }
    }
    // ---------------- phase B ----------------
    if (warp == 0) {
      if (lane == 0) {
        const int slab0 = (int)((row0 - shift) / BK);
        for (int w = wi0; w < items_b; w += wi_step) {
          const int mt = 2 * (w / p.n_tiles) + rank, nt = w % p.n_tiles;
          const long long tf0 = p.prof ? clock64() : 0;
          flag_wait(p.flags + FLAG_PAD * nt, (unsigned int)(2 * pairs_f * (step + 1)));  // (P-Y)^T of this step
          if (p.prof) t_flag += clock64() - tf0;
          produce(&tmXt_hi, &tmXt_lo, &tmDm_hi, &tmDm_lo, k_chunks_b, 0, 0, slab0 * p.xt_rows + mt * BM, p.xt_rows, 0, 0,
                  nt * p.BN + rank * (p.BN / 2), p.dm_rows);
        }
      }
    } else if (warp == 1) {
      if (leader)
        for (int w = wi0; w < items_b; w += wi_step) issue(k_chunks_b);
    } else if (warp >= 4) {
      const float eps = p.eps[step];
      for (int w = wi0; w < items_b; w += wi_step, ++it) {
        const int mt = 2 * (w / p.n_tiles) + rank, nt = w % p.n_tiles;
        const int buf = it & 1;
        SgPre<KP> pre;  // noise / momentum of the update, produced while the accumulator is still being computed
        lap(6);
        if (mt < p.m_tiles_b) {
          if (p.kind == BHMC_KIND_SGLD) sg_update_pre<KP, EW, BHMC_KIND_SGLD>(p, mt, nt, part, ew, lane, step, pre);
          else sg_update_pre<KP, EW, BHMC_KIND_SGD>(p, mt, nt, part, ew, lane, step, pre);
        }
        lap(3);
        mbar_wait(smem_u32(&bar_tfull[buf]), (uint32_t)(it >> 1) & 1u);
        lap(4);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * TMEM_BUF_COLS);
        if (mt < p.m_tiles_b) {  // (a phantom tile has no gradient rows)
          if (p.kind == BHMC_KIND_SGLD)
            sg_update_tile<KP, EW, BHMC_KIND_SGLD, true>(p, tacc, mt, nt, part, ew, lane, step, eps, epi_S, &pre);
          else
            sg_update_tile<KP, EW, BHMC_KIND_SGD, true>(p, tacc, mt, nt, part, ew, lane, step, eps, epi_S, &pre);
        }
        tcgen05_fence_before();
        if (leader) mbar_arrive(smem_u32(&bar_tempty[buf]));
        else mbar_arrive_remote(smem_u32(&bar_tempty[buf]), 0);
        lap(5);
        flag_signal<EW>(p.flags + FLAG_PAD * (p.n_tiles + nt));
        lap(2);
      }
    }
  }
  if (timer) {  // [0] wait F accumulator, [1] F epilogue, [2] flag signals, [3] noise, [4] wait B accumulator, [5] update epilogue, [6] rest
    long long* o = p.prof + (size_t)blockIdx.x * 8;
    for (int i = 0; i < 7; ++i) o[i] = tp[i];
  }
  if (p.prof && warp == 0 && lane == 0) p.prof[(size_t)blockIdx.x * 8 + 7] = t_flag;
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

template <int KP>
static int launch_sg_persistent(bhmc_ctx* ctx, const CUtensorMap* maps, const PersistParams& p, int grid, size_t smem) {
  static size_t configured = 0;
  if (smem > configured) {
    BHMC_CUDA_OK(cudaFuncSetAttribute(k_sg_persistent<KP, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BHMC_CUDA_OK(cudaFuncSetAttribute(k_sg_persistent2<KP, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(NON_EPI_THREADS + 32 * 16);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[2];
  // BHMC_PERSIST_NOCOOP=1 (profiling only): ncu 2025.2.1 fails the cooperative + cluster launch of k_sg_persistent2 with
  // LaunchFailed before the kernel runs (grid and block reported as 0); without the attribute the launch profiles.  The grid
  // is at most one CTA per SM, so on an otherwise idle device (ncu serialises kernels) every CTA is still resident.
  static int nocoop = -1;
  if (nocoop < 0) {
    const char* e = getenv("BHMC_PERSIST_NOCOOP");
    nocoop = (e && atoi(e) != 0) ? 1 : 0;
  }
  int na = 0;
  if (!nocoop) {
    attr[na].id = cudaLaunchAttributeCooperative;  // every CTA resident: the grid barrier / flag waits cannot deadlock
    attr[na].val.cooperative = 1;
    ++na;
  }
  if (p.pair || p.two_cta) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (p.two_cta)
    BHMC_CUDA_OK(cudaLaunchKernelEx(&cfg, k_sg_persistent2<KP, 16>, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6],
                                    maps[7], p));
  else
    BHMC_CUDA_OK(cudaLaunchKernelEx(&cfg, k_sg_persistent<KP, 16>, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6],
                                    maps[7], p));
  ctx->launches++;
  BHMC_CUDA_OK(cudaGetLastError());
  return BHMC_OK;
}

// n_steps consecutive minibatch steps (windows row_first + j*batch) of SGLD / SGD in one launch.  BHMC_ERR_UNSUPPORTED
// when this shape / layout has no persistent kernel (the caller falls back to the per-step launches).
int tc_softmax_sg_persistent(bhmc_ctx* ctx, const SoftmaxData& d, int C, int64_t ld, float alpha, bool split3, const FusedStep& fs,
                             int64_t row_first, int64_t batch, int n_steps, const float* eps_dev, int64_t z_step_stride,
                             uint64_t step0) {
  if (!d.tc_ready || (split3 && !d.has_lo)) return BHMC_ERR_UNSUPPORTED;
  const int KP = d.Kp, K = d.K, D = d.D;
  if (K != KP || !(KP == 4 || KP == 8 || KP == 10 || KP == 16)) return BHMC_ERR_UNSUPPORTED;
  if (d.xa_blocked || d.slab != BK || d.slab_ld != BK || BK != 64) return BHMC_ERR_UNSUPPORTED;
  if (batch < 1 || row_first < 0 || row_first + (int64_t)n_steps * batch > d.N || n_steps < 1) return BHMC_ERR_UNSUPPORTED;
  if (row_first + (int64_t)n_steps * batch + BM >= ((int64_t)1 << 31)) return BHMC_ERR_UNSUPPORTED;  // 32-bit TMA coordinates
  int cpt = 1;
  while ((cpt * KP) % 16) ++cpt;  // narrowest chain tile: more work items than a wide one, and every step is latency-bound
  const int BN = cpt * KP, n_tiles = (int)ceil_div(C, cpt);
  const int64_t Mfwd = round_up(batch, BM);
  if (Mfwd / BM * n_tiles > 4 * ctx->sm_count) return BHMC_ERR_UNSUPPORTED;  // big windows: the throughput kernels win
  const int smode = split3 ? (d.x_exact ? 2 : 1) : 0;
  const int nmat = split3 ? 2 : 1, na = smode == 1 ? 2 : 1;
  // cta_group::2 items (k_sg_persistent2): every CTA stages half of the chain-side operand
  static int two_env = -1;
  if (two_env < 0) {
    const char* e = getenv("BHMC_PERSIST_2CTA");
    two_env = e ? atoi(e) : 1;  // measured at cfg3: 28.5-28.7 -> 25.1-25.9 us per step (4.47 -> 4.94-5.09 M grad-evals/s)
  }
  const int pairs_max = std::max(((int)(Mfwd / BM) + 1) / 2, ((int)ceil_div(d.Dt, BM) + 1) / 2);
  const bool two_cta = two_env && (BN / 2) % 8 == 0 && 2 * pairs_max * n_tiles <= (ctx->sm_count & ~1);
  const int stage_bytes = na * BM * BK * 2 + nmat * (two_cta ? BN / 2 : BN) * BK * 2;
  const size_t epi_bytes = (size_t)16 * 32 * KP * sizeof(float);  // update epilogue: one 32 x KP block per epilogue warp
  const int stages = std::max(2, std::min(MAX_STAGES, (int)((225 * 1024 - epi_bytes) / stage_bytes)));
  const size_t smem = (size_t)stages * stage_bytes + 1024 + epi_bytes;
  const int64_t ncols = (int64_t)C * KP;
  const int64_t dm_rows = (int64_t)n_tiles * BN, dm_nslab = ceil_div(Mfwd + BK, BK);
  void *wt = nullptr, *dmt = nullptr, *bar = nullptr;
  const size_t wt_bytes = (size_t)ncols * d.Dp * 2, dmt_bytes = (size_t)(dm_nslab * dm_rows * BK) * 2;
  BHMC_TRY(ctx->get_scratch(1, wt_bytes * 2, &wt));
  BHMC_TRY(ctx->get_scratch(2, dmt_bytes * 2, &dmt));
  const size_t bar_bytes = 256 + sizeof(unsigned int) * 2 * (size_t)n_tiles * FLAG_PAD;  // grid-barrier counter | chain-tile item counters (one per line)
  BHMC_TRY(ctx->get_scratch(13, bar_bytes, &bar));
  BHMC_CUDA_OK(cudaMemsetAsync(bar, 0, bar_bytes, ctx->stream));
  __nv_bfloat16* wt_hi = (__nv_bfloat16*)wt;
  __nv_bfloat16* wt_lo = (__nv_bfloat16*)((char*)wt + wt_bytes);
  {  // operand copy of the start point (later steps write it from the update epilogue)
    GroupTimer t(ctx, KG_PREP);
    dim3 grid((unsigned)ceil_div(d.Dp, 128), C);
    k_tc_prep<<<grid, 128, 0, ctx->stream>>>(fs.q, ld, D, K, KP, d.Dp, wt_hi, split3 ? wt_lo : nullptr, nullptr, 1.0f / d.x_scale);
    ctx->launches++;
  }
  CUtensorMap maps[8];
  const uint64_t xt_rows_total = (uint64_t)(d.Npad / d.slab) * d.Dt_pad;
  // CTA pairs (round 2): the phases are bound by what one SM can ingest per chunk (32 KB of X or X^T hi/lo + 20 KB of the
  // chain-side operand for 12 MMAs of N = 80); two CTAs on adjacent chain tiles need the SAME X tile, so each fetches half
  // of it and multicasts -- 36 KB requested per CTA and chunk instead of 52.  Needs an even number of chain tiles and CTAs.
  static int pair_env = -1;
  if (pair_env < 0) {
    const char* e = getenv("BHMC_PERSIST_PAIR");
    pair_env = e ? atoi(e) : 0;  // measured (cfg3, BHMC_PROF=1): phase F 22.1 k -> 23.0 k cycles, phase B 23.6 k -> 25.1 k, 4.15 -> 3.90 M
                                 // grad-evals/s: the chunk time is not set by what a CTA REQUESTS (36 KB instead of 52 KB) but
                                 // by what lands in its shared memory -- off by default, kept for A/B
  }
  const int items_all = std::max((int)(Mfwd / BM) * n_tiles, (int)ceil_div(d.Dt, BM) * n_tiles);
  const int grid_all = std::min(items_all, ctx->sm_count);
  const bool pair = !two_cta && pair_env && n_tiles % 2 == 0 && grid_all % 2 == 0 && grid_all >= 2;
  // half row tiles in the forward phase (default): at cfg3 the phase has 64 items for 148 SMs and each is bound by what
  // one SM ingests per chunk (32 KB of X hi/lo + 20 KB of W^T); half tiles make 128 items of 16 + 20 KB
  static int half_env = -1;
  if (half_env < 0) {
    const char* e = getenv("BHMC_PERSIST_HALF");
    half_env = e ? atoi(e) : 0;  // measured: 128 items of 36 KB per chunk take as long as 64 items of 52 KB (phase F 20.1 k vs
                                 // 19.5-22 k cycles; 4.17 vs 4.14-4.22 M grad-evals/s) -- see the note on the MMA warp below
  }
  const bool half_f = half_env && !pair && !two_cta && 2 * (int)(Mfwd / BM) * n_tiles <= ctx->sm_count;
  const uint32_t abox = pair ? BM / 2 : BM;
  const uint32_t abox_f = (pair || half_f) ? BM / 2 : BM;
  BHMC_TRY(make_map(&maps[0], d.Xa_hi, (uint64_t)d.Dp, (uint64_t)d.N, (uint64_t)d.Dp, abox_f));
  maps[1] = maps[0];
  if (smode == 1) BHMC_TRY(make_map(&maps[1], d.Xa_lo, (uint64_t)d.Dp, (uint64_t)d.N, (uint64_t)d.Dp, abox_f));
  const uint32_t bbox = (uint32_t)(two_cta ? BN / 2 : BN);
  BHMC_TRY(make_map(&maps[2], wt_hi, (uint64_t)d.Dp, (uint64_t)ncols, (uint64_t)d.Dp, bbox));
  maps[3] = maps[2];
  if (split3) BHMC_TRY(make_map(&maps[3], wt_lo, (uint64_t)d.Dp, (uint64_t)ncols, (uint64_t)d.Dp, bbox));
  BHMC_TRY(make_map(&maps[4], d.Xt_hi, (uint64_t)d.slab, xt_rows_total, (uint64_t)d.slab_ld, abox));
  maps[5] = maps[4];
  if (smode == 1) BHMC_TRY(make_map(&maps[5], d.Xt_lo, (uint64_t)d.slab, xt_rows_total, (uint64_t)d.slab_ld, abox));
  __nv_bfloat16* dmt_hi = (__nv_bfloat16*)dmt;
  __nv_bfloat16* dmt_lo = split3 ? (__nv_bfloat16*)((char*)dmt + dmt_bytes) : nullptr;
  BHMC_TRY(make_map(&maps[6], dmt_hi, (uint64_t)BK, (uint64_t)(dm_nslab * dm_rows), (uint64_t)BK, bbox));
  maps[7] = maps[6];
  if (split3) BHMC_TRY(make_map(&maps[7], dmt_lo, (uint64_t)BK, (uint64_t)(dm_nslab * dm_rows), (uint64_t)BK, bbox));
  PersistParams p{};
  p.D = D, p.K = K, p.C = C, p.cpt = cpt, p.BN = BN, p.n_tiles = n_tiles;
  p.k_chunks_f = (int)ceil_div(D, BK);
  p.m_tiles_b = (int)ceil_div(d.Dt, BM);
  p.split3 = smode, p.stages = stages;
  p.ld = ld, p.Dp = d.Dp;
  p.q = fs.q, p.p = fs.p;
  p.labels = d.labels;
  p.batch = batch, p.row_first = row_first, p.n_steps = n_steps;
  p.eps = eps_dev;
  p.gamma = fs.gamma, p.alpha = alpha, p.scale = 1.0f / d.x_scale;
  p.kind = fs.kind;
  p.z = fs.z, p.ld_z = fs.ld_z, p.z_step_stride = z_step_stride;
  p.seed = fs.seed, p.chain_id0 = fs.chain_id0, p.step0 = step0;
  p.wt_hi = wt_hi, p.wt_lo = split3 ? wt_lo : nullptr, p.dmt_hi = dmt_hi, p.dmt_lo = dmt_lo;
  p.dm_rows = (int)dm_rows, p.xt_rows = (int)d.Dt_pad;
  p.bar = (unsigned int*)bar;
  p.pair = pair ? 1 : 0;
  p.half_f = half_f ? 1 : 0;
  static int flags_env = -1;
  if (flags_env < 0) {
    const char* e = getenv("BHMC_PERSIST_FLAGS");
    flags_env = e ? atoi(e) : 1;
  }
  p.flags = (flags_env || two_cta) ? reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(bar) + 256) : nullptr;
  p.two_cta = two_cta ? 1 : 0;
  const int items = std::max((half_f ? 2 : 1) * (int)(Mfwd / BM) * n_tiles, p.m_tiles_b * n_tiles);
  const int grid = two_cta ? 2 * pairs_max * n_tiles : std::min(items, ctx->sm_count);
  static int want_prof = -1, want_pf = -1;
  if (want_prof < 0) want_prof = getenv("BHMC_PROF") ? 1 : 0;
  if (want_pf < 0) {
    const char* e = getenv("BHMC_PERSIST_PF");
    want_pf = e ? atoi(e) : 1;
  }
  p.prefetch = want_pf;
  void* prof_dev = nullptr;
  if (want_prof) {
    BHMC_TRY(ctx->get_scratch(5, sizeof(long long) * 8 * 1024, &prof_dev));
    BHMC_CUDA_OK(cudaMemsetAsync(prof_dev, 0, sizeof(long long) * 8 * 1024, ctx->stream));
    p.prof = (long long*)prof_dev;
  }
  int rc = BHMC_ERR_UNSUPPORTED;
  {
    GroupTimer t(ctx, KG_FWD);
    switch (KP) {
      case 4: rc = launch_sg_persistent<4>(ctx, maps, p, grid, smem); break;
      case 8: rc = launch_sg_persistent<8>(ctx, maps, p, grid, smem); break;
      case 10: rc = launch_sg_persistent<10>(ctx, maps, p, grid, smem); break;
      case 16: rc = launch_sg_persistent<16>(ctx, maps, p, grid, smem); break;
    }
  }
  if (want_prof && rc == BHMC_OK && two_cta) {
    std::vector<long long> hp(8 * 148);
    BHMC_CUDA_OK(cudaMemcpyAsync(hp.data(), prof_dev, sizeof(long long) * 8 * 148, cudaMemcpyDeviceToHost, ctx->stream));
    BHMC_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    double sF[8] = {0}, sA[8] = {0};
    int nF = 0, nA = 0;
    const int f_ctas = 2 * (((int)(Mfwd / BM) + 1) / 2) * n_tiles;  // CTAs that own a forward item
    for (int b = 0; b < grid && b < 148; ++b) {
      for (int j = 0; j < 8; ++j) (b < f_ctas ? sF[j] : sA[j]) += (double)hp[b * 8 + j] / n_steps;
      (b < f_ctas ? nF : nA)++;
    }
    const char* names[8] = {"wait F acc", "F epilogue", "flag signals", "noise", "wait B acc", "update epilogue", "rest", "producer flag waits"};
    for (int g2 = 0; g2 < 2; ++g2) {
      const int n = g2 ? nA : nF;
      if (!n) continue;
      fprintf(stderr, "[bhmc prof persist2] %d CTAs %s, cycles per step:", n, g2 ? "with a backward item only" : "with forward + backward items");
      for (int j = 0; j < 8; ++j) fprintf(stderr, " %s %.0f;", names[j], (g2 ? sA[j] : sF[j]) / n);
      fprintf(stderr, "\n");
    }
  } else if (want_prof && rc == BHMC_OK) {
    std::vector<long long> hp(8 * 148);
    BHMC_CUDA_OK(cudaMemcpyAsync(hp.data(), prof_dev, sizeof(long long) * 8 * 148, cudaMemcpyDeviceToHost, ctx->stream));
    BHMC_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    double s[4] = {0, 0, 0, 0}, mx[4] = {0, 0, 0, 0}, mn[4] = {1e30, 1e30, 1e30, 1e30};
    int n = 0;
    {  // MMA warp of the CTAs that own a forward item: operand wait and total per step in phase F, operand wait in phase B
      double w0 = 0, a0 = 0, w1 = 0;
      int m = 0;
      for (int b = 0; b < grid && b < 148; ++b)
        if (hp[b * 8 + 6] > 0) w0 += (double)hp[b * 8 + 5], a0 += (double)hp[b * 8 + 6], w1 += (double)hp[b * 8 + 7], ++m;
      if (m)
        fprintf(stderr, "[bhmc prof persist] MMA warp (%d CTAs with forward items), cycles per step: phase F operand wait %.0f of %.0f, phase B operand wait %.0f\n",
                m, w0 / m / n_steps, a0 / m / n_steps, w1 / m / n_steps);
    }
    for (int b = 0; b < grid && b < 148; ++b) {
      if (hp[b * 8 + 4] <= 0) continue;
      ++n;
      for (int j = 0; j < 4; ++j) {
        const double v = (double)hp[b * 8 + j] / (double)hp[b * 8 + 4];
        s[j] += v, mx[j] = std::max(mx[j], v), mn[j] = std::min(mn[j], v);
      }
    }
    if (n)
      fprintf(stderr, "[bhmc prof persist] %d CTAs, %d steps, items F %d B %d, BN %d stages %d | cycles per step, mean (min..max) over CTAs: phase F %.0f (%.0f..%.0f); barrier 1 %.0f (%.0f..%.0f); phase B %.0f (%.0f..%.0f); barrier 2 %.0f (%.0f..%.0f)\n",
              n, n_steps, (int)(Mfwd / BM) * n_tiles, p.m_tiles_b * n_tiles, BN, stages, s[0] / n, mn[0], mx[0], s[1] / n, mn[1], mx[1], s[2] / n,
              mn[2], mx[2], s[3] / n, mn[3], mx[3]);
  }
  return rc;
}
