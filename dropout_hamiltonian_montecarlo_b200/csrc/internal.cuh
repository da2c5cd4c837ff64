// Internal declarations shared by the translation units of libbhmc.so.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/bhmc.h"

namespace bhmc {

void set_error(const char* fmt, ...);

#define BHMC_CUDA_OK(expr)                                                                 \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      ::bhmc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                        __LINE__);                                                         \
      return BHMC_ERR_CUDA;                                                                \
    }                                                                                      \
  } while (0)

#define BHMC_CHECK_ARG(cond, ...)   \
  do {                              \
    if (!(cond)) {                  \
      ::bhmc::set_error(__VA_ARGS__); \
      return BHMC_ERR_ARG;          \
    }                               \
  } while (0)

#define BHMC_TRY(expr)          \
  do {                          \
    int rc__ = (expr);          \
    if (rc__ != BHMC_OK) return rc__; \
  } while (0)

enum KernelGroup { KG_FWD = 0, KG_BWD = 1, KG_PREP = 2, KG_UPDATE = 3, KG_COUNT = 4 };

struct EventPair {
  cudaEvent_t a, b;
};

}  // namespace bhmc

struct bhmc_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  int64_t launches = 0;
  // optional per-group device timing (CUDA events on the launching stream)
  int timing = 0;  // 0 off, 1 every kernel group, 2 only the GEMM groups (KG_FWD, KG_BWD)
  std::vector<bhmc::EventPair> pool[bhmc::KG_COUNT];
  size_t used[bhmc::KG_COUNT] = {0, 0, 0, 0};
  double ms_acc[bhmc::KG_COUNT] = {0, 0, 0, 0};
  int64_t n_acc[bhmc::KG_COUNT] = {0, 0, 0, 0};
  // sampled timing: only every timing_stride-th launch group of a kind is bracketed by events; units_acc sums the
  // work units (chains) of exactly the bracketed launches, so time / units stays consistent under ragged launches
  int timing_stride = 1;
  int64_t seen[bhmc::KG_COUNT] = {0, 0, 0, 0};
  bool sampled[bhmc::KG_COUNT] = {false, false, false, false};
  double units_acc[bhmc::KG_COUNT] = {0, 0, 0, 0};
  int64_t cur_units = 0;  // set by the callers of model->grad (rows of this launch)
  // grow-only device scratch
  void* scratch[16] = {nullptr};
  size_t scratch_bytes[16] = {0};
  // pinned staging for small per-step host->device uploads
  void* pinned = nullptr;
  size_t pinned_bytes = 0;
  cudaEvent_t pinned_ev = nullptr;
  bool pinned_inflight = false;

  const void* zcache_owner = nullptr;  // ZCache descriptor whose data scratch slot 8 currently holds
  int get_scratch(int slot, size_t bytes, void** out);
  int get_pinned(size_t bytes, void** out);
  void begin_group(int group);
  void end_group(int group);
  int flush_timing();
};

namespace bhmc {

struct GroupTimer {
  bhmc_ctx* ctx;
  int group;
  GroupTimer(bhmc_ctx* c, int g) : ctx(c), group(g) { ctx->begin_group(g); }
  ~GroupTimer() { ctx->end_group(group); }
};

// Order-independent accumulation of energy terms (log-likelihood, kinetic energy, prior energy).  Several blocks add
// their partial sums to one fp64 cell with atomics, in whatever order they arrive, and fp64 addition is not
// associative: round 1's energies differed in the last bits from run to run, and with them -- rarely -- an accept
// decision.  Every addend is now rounded to a multiple of 2^-FRAC first.  Sums of such numbers are EXACT in fp64 while
// they stay below 2^(53-FRAC), so the result no longer depends on the order: same inputs, same bits.  FRAC = 24: exact up
// to |sum| < 5.4e8 (a log-likelihood of 1e6 rows is ~4e6), the rounding adds < 6e-8 per partial (relative 1e-10 of such
// a sum; the fp32 per-row terms carry 1e-7).  Beyond the range it degrades to ordinary rounding, never to a wrong value.
template <int FRAC = 24>
__device__ __forceinline__ double quantize_addend(double x) {
  constexpr double S = (double)(1ull << FRAC);
  return rint(x * S) * (1.0 / S);
}

inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Minibatch samplers (SGLD / SGD) on small windows are bound by launch latency, not throughput: the split-K reduce,
// the parameter update and the operand preparation of the NEXT evaluation are one kernel (k_tc_reduce_step).
struct FusedStep {
  int kind;            // BHMC_KIND_SGLD: p = 2 eps z - eps/2 g, q += p (sgld.py:31-46); BHMC_KIND_SGD: m = gamma m - eps g, q += m
  float* q;            // chain state, updated in place (the gradient is evaluated at its value on entry)
  float* p;            // SGLD: momentum out; SGD: heavy-ball momentum in/out
  float eps, gamma;
  const float* z;      // injected N(0,1) tape [C, ld_z] or nullptr (Philox)
  int64_t ld_z;
  uint64_t seed;
  int64_t chain_id0;
  uint32_t stream_lo, stream_hi;
  bool wt_ready;       // the previous fused step already wrote the bf16 operand copy of q and zeroed loglik
};

// ---------------------------------------------------------------------------------------
// model interface used by the sampler drivers
// ---------------------------------------------------------------------------------------
enum : uint32_t {
  GRAD_HINT_CHEAP_MOVE = 1u,  // since the previous evaluation of these rows only a cheap_slice() moved
  GRAD_HINT_KEEP = 2u,        // the NEXT evaluation of these rows will follow a cheap move: keep what it can reuse
  GRAD_HINT_PREPARED = 4u,    // the previous fused evaluation already prepared this evaluation's operands from q
};
struct FusedStream;
struct UpdateArgs;
struct ModelBase {
  bhmc_ctx* ctx = nullptr;
  int64_t P = 0;
  int n_vars = 0;
  int64_t var_off[BHMC_MAX_VARS] = {0};
  int64_t var_len[BHMC_MAX_VARS] = {0};
  virtual ~ModelBase() {}
  // g may be nullptr (log-lik only). stat[c] (double, device) receives the model's scalar.
  // hint (GRAD_HINT_*): what the caller knows about q relative to the previous evaluation of the same rows
  virtual int grad(const float* q, int C, int64_t ld, int64_t row0, int64_t nrows, int prec, float* g,
                   double* stat, uint32_t hint = 0) = 0;
  // true when moving only parameters [off, off+len) leaves the expensive part of the forward pass unchanged
  // (softmax: the slice lies inside the bias) -> the next evaluation may be requested with GRAD_HINT_CHEAP_MOVE
  virtual bool cheap_slice(int64_t, int64_t) const { return false; }
  // potential used by the Metropolis test: U = a*stat + b (+ sum_v cv[v]*|q_v|^2)
  virtual void energy_coeffs(int64_t nrows, double* a, double* b, double* cv) const = 0;
  virtual int64_t default_rows() const { return 0; }
  virtual int64_t n_features() const { return 0; }
  // Operand mirror (round 2).  A model that reads bf16 hi/lo copies of its parameters may hand the sampler a [C, *mld]
  // image to be filled by an update launch that rewrites EVERY parameter of the rows it moves (joint sweep); the sampler
  // then reports the buffer it wrote with mirror_written(q), and the model may use the image for the next grad(q, ...)
  // only -- any other call sequence makes it split the parameters itself, as before.
  virtual bool operand_mirror(int, int64_t, int, __nv_bfloat16**, __nv_bfloat16**, int64_t*) { return false; }
  virtual void mirror_written(const float*) {}
  // gradient at q followed by the lockstep sampler's whole-vector update (closing kick of one iteration, drift of the
  // next), applied by the model's own kernels; BHMC_ERR_UNSUPPORTED (before anything is launched) = grad() + update launch
  virtual int grad_fused_update(const float*, int, int64_t, int64_t, int64_t, int, double*, const UpdateArgs&) {
    return BHMC_ERR_UNSUPPORTED;
  }
  // gradient at fs.q followed by the sampler's parameter update in the same launch sequence (no g materialised);
  // BHMC_ERR_UNSUPPORTED = caller falls back to grad() + the separate update kernel
  virtual int grad_fused_step(int, int64_t, int64_t, int64_t, int, double*, const FusedStep&) { return BHMC_ERR_UNSUPPORTED; }
  // n_steps minibatch steps (gradient + update each) in one launch; same fallback convention
  virtual int sg_steps_persistent(int, int64_t, int, const FusedStep&, int64_t, int64_t, int, const float*, int64_t, uint64_t) {
    return BHMC_ERR_UNSUPPORTED;
  }
  // gradient (g IS written) + the streaming schedule's next update in the reduce launch; same fallback convention
  virtual int grad_fused_stream(const float*, int, int64_t, int64_t, int64_t, int, float*, double*, uint32_t,
                                const FusedStream&) {
    return BHMC_ERR_UNSUPPORTED;
  }
  // sgd.fit_dropout: gradient on rows [row0, row0+nrows) multiplied elementwise by a Bernoulli(keep) mask
  // (mask != nullptr: injected [nrows, D] keep flags; else Philox keyed by (seed, stream))
  virtual int grad_input_dropout(const float*, int, int64_t, int64_t, int64_t, int, float*, double*, const uint8_t*,
                                 float, uint64_t, uint32_t) {
    set_error("this model has no input-dropout gradient (sgd.fit_dropout)");
    return BHMC_ERR_UNSUPPORTED;
  }
};

// ---- update.cu --------------------------------------------------------------------------
struct UpdateArgs {
  float* q;
  float* p;
  const float* g;
  int64_t ld, P;
  int C;
  // post: p = f_post*p - a_post*g + n_post*z   on slice [post_off, post_off+post_len)
  int64_t post_off, post_len;
  int it_post;
  float f_post, a_post, n_post;
  // pre: p -= a_pre*g ; q += eps*p             on slice [pre_off, pre_off+pre_len)
  int64_t pre_off, pre_len;
  int it_pre;
  float a_pre, eps;
  const int32_t* L;  // per-chain path length (device); chain active for iteration it iff it < L-1
  // noise for the post part: injected (device, [C, ld_z] rows) or Philox
  const float* z;
  int64_t ld_z;
  uint64_t seed;
  int64_t chain_id0;
  uint32_t stream_lo, stream_hi;
  // latch of the model scalar produced by the gradient evaluation that preceded this update
  const double* stat;  // [C]
  double* stat_new;    // [C] written where the chain was active for it_post
  const int32_t* perm; // row -> chain for the noise stream (nullptr = identity)
  // bf16 hi/lo image of the moved positions, element for element ([C, mir_ld]; ModelBase::operand_mirror): a model whose
  // gradient GEMMs read bf16 copies of the parameters gets them from the kernel that wrote the parameters
  __nv_bfloat16* mir_hi;
  __nv_bfloat16* mir_lo;
  int64_t mir_ld;
};
int launch_hmc_update(bhmc_ctx* ctx, const UpdateArgs& a);

// ---- streaming schedule (asynchronous chains): every working row follows its OWN position in its own sequence of
// HMC transitions; one 32-bit op code per (phase, row) says what the row does between two gradient launches.
enum : uint32_t {
  OP_LATCH = 1u,    // the previous launch evaluated the step's start point (first step of a call only):
                    // stat_cur = stat_new = stat, g_start = g
  OP_POST = 2u,     // closing kick of variable (code >> 8) & 15; also latches stat_new = stat
  OP_PRE = 4u,      // opening half kick + drift of variable (code >> 12) & 15
  OP_FINISH = 8u,   // kinetic energy + Metropolis test of step op_step[row]
  OP_BEGIN = 16u,   // momentum draw / proposal := state for step op_step[row] + (FINISH ? 1 : 0); after a FINISH the
                    // gradient at the new start point is NOT re-evaluated: accepted -> g_start = g, rejected -> g = g_start
  OP_LATCH_CACHED = 32u,  // start-of-step statistics taken from what the Metropolis test left: stat_cur = stat_new =
                          // stat_next, extra_cur = extra_next
};
struct StreamUpdateArgs {
  float* q;
  float* p;
  const float* g;
  int64_t ld, P;
  int rows;
  const uint32_t* code;   // [rows] ops of this phase
  const int32_t* step;    // [rows] step index the row's op refers to
  int n_vars;
  int64_t off[BHMC_MAX_VARS], len[BHMC_MAX_VARS];
  float a_pre, a_post, eps;
  const double* stat;
  double *stat_cur, *stat_new;
  double* kin0;           // [2][C_total] kinetic energy at the start of the step, buffer = step parity
  double* kin1;           // [C_total] zeroed here for the kinetic kernel of this phase
  int C_total;
  float* g_start;         // [rows, ld] gradient at the start point of the row's current step
  const double* stat_next;   // written by the Metropolis test: statistic / prior energy of the state it selected
  const double* extra_next;  // (may be nullptr)
  double* extra_cur;
};
int launch_stream_update(bhmc_ctx* ctx, const StreamUpdateArgs& a);
// Streaming schedule, fused tail of a gradient evaluation: the split-K reduce also executes the NEXT phase's pre-event
// update (u: its op codes / rows) and, when the following launch is a full forward pass of unchanged rows, writes the
// bf16 operand copy of the updated weights (prep_next) -- 3 launches per phase instead of 5.
struct FusedStream {
  StreamUpdateArgs u;
  bool prep_next;
};
int launch_stream_kinetic(bhmc_ctx* ctx, const float* p, int64_t ld, int64_t P, int rows, const uint32_t* code,
                          double* kin);

struct BeginArgs {
  const float* q;
  float* q_new;
  float* p0;
  float* p_new;
  int64_t ld, P;
  int C;
  const float* z;  // injected N(0,1) [C,P] compact (ld_z) or nullptr
  int64_t ld_z;
  uint64_t seed;
  int64_t chain_id0;
  uint32_t stream_lo, stream_hi;
  double* kin0;  // [C] zeroed by the launcher, receives 0.5*sum p^2
  // ragged-trajectory compaction: working row r holds chain perm[r] (chains sorted by path length,
  // longest first, so the active set is always a prefix); nullptr = identity.  q / p0 / z are chain-indexed,
  // q_new / p_new / kin0 are row-indexed.
  const int32_t* perm;
  // streaming schedule: only rows whose op has OP_BEGIN take part; their step index is step[r] (+1 after a FINISH);
  // stream_lo is the run's first step, z / kin0 are offset by the step (kin0: parity buffers of C_total doubles)
  const uint32_t* code;
  const int32_t* step;
  int64_t z_step_stride;
  int C_total;
  float* g;                 // streaming: working gradient and its start-of-step copy (see OP_BEGIN)
  float* g_start;
  const int32_t* acc_flag;  // [rows] decision of the Metropolis test that preceded this begin
};
int launch_hmc_begin(bhmc_ctx* ctx, const BeginArgs& a);

int launch_kinetic(bhmc_ctx* ctx, const float* p, int64_t ld, int64_t P, int C, double* kin);
// per-variable sum of squares: out[c*n_vars + v]
int launch_sumsq(bhmc_ctx* ctx, const float* q, int64_t ld, int C, int n_vars, const int64_t* off,
                 const int64_t* len, double* out);

int launch_prior_energy(bhmc_ctx* ctx, const float* q, int64_t ld, int C, int n_vars, const int64_t* off,
                        const int64_t* len, const double* cv, double* sumsq_scratch, double* out,
                        const uint32_t* code = nullptr, uint32_t flag = 0);  // code: write out[r] only where code[r] & flag

struct AcceptArgs {
  float* q;            // in/out: current state, overwritten by the proposal where accepted
  const float* q_new;
  float* p_out;        // in: p0 ; out: accepted ? sign*p_new : p0
  const float* p_new;
  int64_t ld, P;
  int C;
  float p_sign;        // -1 for HMC (hmc.py:58-59), +1 for SGHMC
  const double* stat_cur;
  const double* stat_new;
  const double* extra_cur;  // optional sum_v cv*|q_v|^2 terms (may be nullptr)
  const double* extra_new;
  double ea, eb;       // U = ea*stat + eb + extra
  const double* kin0;
  const double* kin1;
  const double* u;     // [C] injected uniforms (device) or nullptr -> Philox
  uint64_t seed;
  int64_t chain_id0;
  uint32_t stream_lo, stream_hi;
  int reject_nan;
  float* sample;       // [C, P] compact or nullptr
  double* loss;        // [C] or nullptr
  double* accept_prob; // [C] or nullptr
  int32_t* accepted;   // [C] or nullptr
  const int32_t* perm; // row -> chain (see BeginArgs); q / p_out / u / outputs are chain-indexed
  // streaming schedule: only rows with OP_FINISH; t = step[r] offsets the Philox stream, u, the outputs (t*C_total
  // rows) and selects the kin0 parity buffer
  const uint32_t* code;
  const int32_t* step;
  int C_total;
  double* stat_next;   // [rows] streaming: statistic / prior energy of the selected state, decision flag
  double* extra_next;
  int32_t* acc_flag;
};
int launch_accept(bhmc_ctx* ctx, const AcceptArgs& a);

struct SgldArgs {
  float* q;
  float* p;
  const float* g;
  int64_t ld, P;
  int C;
  float eps;      // p = 2*eps*z - 0.5*eps*g ; q += p   (sgld.py:31-46)
  const float* z; // injected or nullptr
  int64_t ld_z;
  uint64_t seed;
  int64_t chain_id0;
  uint32_t stream_lo, stream_hi;
};
int launch_sgld_update(bhmc_ctx* ctx, const SgldArgs& a);
// heavy-ball: m = gamma*m - eps*g ; q += m  (sgd.py:40-41)
int launch_sgd_update(bhmc_ctx* ctx, float* q, float* m, const float* g, int64_t ld, int64_t P, int C,
                      float gamma, float eps);
int launch_copy_rows(bhmc_ctx* ctx, const float* src, int64_t ld_src, float* dst, int64_t ld_dst, int64_t P,
                     int C);
int launch_philox_normal(bhmc_ctx* ctx, float* out, int C, int64_t P, int64_t ld, uint64_t seed,
                         int64_t chain_id0, uint32_t stream_lo, uint32_t stream_hi);
int launch_affine(bhmc_ctx* ctx, const double* in, double a, double b, const double* extra, double* out, int n);

// ---- softmax_simt.cu / softmax_tc.cu -----------------------------------------------------
struct SoftmaxData {
  int64_t N = 0;
  int D = 0, K = 0;
  const float* X = nullptr;        // [N, D] fp32 (bound or owned)
  const int32_t* labels = nullptr; // [N]
  bool owned = false;
  // tensor-core operand copies (built by tc_bind): bf16 hi / lo
  int Kp = 0;                       // classes padded per chain (even)
  int64_t Dp = 0;                   // feature stride of Xa (multiple of 8 elements)
  int64_t Npad = 0;                 // rows covered by the Xt slabs (n_slabs * slab)
  int64_t slab = 0;                 // rows of X per Xt slab (multiple of 64)
  int64_t slab_ld = 0;              // row stride of an Xt slab: slab + 64 (a power-of-two stride would map the 128
                                    // rows of a tile onto the same L2 sets)
  int64_t Dt_pad = 0;               // Xt rows per slab (D+1 rounded up to the 128-row tile, zero filled)
  void* Xa_hi = nullptr;            // [N, Dp]   K-major A of the forward GEMM
  void* Xa_lo = nullptr;
  void* Xt_hi = nullptr;            // [Npad/slab][Dt_pad, slab] K-major A of the backward GEMM (row D = ones)
  void* Xt_lo = nullptr;
  int64_t Dt = 0;                   // D+1 rows (ones row feeds the bias gradient)
  bool has_lo = false;
  bool tc_ready = false;
  // exact-operand detection at bind time (softmax_tc.cu:k_detect_exact): the bf16 operand copies hold x_scale * X; when
  // that is exact for every element (x_exact) bf16x3 needs no lo copy of X and runs 2 MMAs per product instead of 3
  float x_scale = 1.f;
  bool x_exact = false;
  bool xa_blocked = false;          // Xa is stored k-chunk-major: [Dp/64][N][64] (one contiguous 16 KB block per TMA box)
};

int simt_softmax_grad(bhmc_ctx* ctx, const SoftmaxData& d, const float* q, int C, int64_t ld, float alpha,
                      int64_t row0, int64_t nrows, float* g, double* loglik);
int simt_softmax_predict(bhmc_ctx* ctx, int D, int K, const float* q, int C, int64_t ld, const float* X,
                         int64_t nrows, float* probs, int32_t* labels);

int tc_softmax_bind(bhmc_ctx* ctx, SoftmaxData& d, bool want_lo, bool detect_exact = true);
void tc_softmax_release(SoftmaxData& d);
// Z cache (X.W of the last full forward pass, scratch slot 8) and how an evaluation may use it
struct ZCache {
  bool valid = false;
  int64_t row0 = 0, nrows = 0;
  int C = 0, KP = 0, slab_rows = 0;
  int64_t slab = 0, ld = 0;
};
enum { ZMODE_NONE = 0, ZMODE_STORE = 1, ZMODE_USE = 2 };  // USE falls back to a full pass when the cache does not fit
int tc_softmax_grad(bhmc_ctx* ctx, const SoftmaxData& d, const float* q, int C, int64_t ld, float alpha,
                    int64_t row0, int64_t nrows, float* g, double* loglik, bool split3, const FusedStep* fs = nullptr,
                    ZCache* zc = nullptr, int zmode = ZMODE_NONE, const FusedStream* fst = nullptr, bool prepared = false);

// n_steps consecutive minibatch steps (rows row_first + j*batch) of SGLD / SGD in ONE cooperative launch
// (softmax_persist.cuh); eps_dev[n_steps] = step size of every step; fs.z (if any) advances by z_step_stride per step.
// BHMC_ERR_UNSUPPORTED = no persistent kernel for this shape: use the per-step path.
int tc_softmax_sg_persistent(bhmc_ctx* ctx, const SoftmaxData& d, int C, int64_t ld, float alpha, bool split3, const FusedStep& fs,
                             int64_t row_first, int64_t batch, int n_steps, const float* eps_dev, int64_t z_step_stride,
                             uint64_t step0);

// ---- mlp.cu -----------------------------------------------------------------------------------
ModelBase* mlp_model_new(bhmc_ctx* ctx, int64_t n_rows, int n_in, int n_mid, int n_out, float alpha, float ratio,
                         uint64_t seed, int64_t chain_id0);
int mlp_model_bind(ModelBase* m, const float* X, const int32_t* labels, int is_host);
int mlp_model_set_masks(ModelBase* m, const uint8_t* masks_dev);
int mlp_model_predict(ModelBase* m, const float* q, int C, int64_t ld, const float* X_dev, int64_t nrows, int prec,
                      float* probs_dev, int32_t* labels_dev);

}  // namespace bhmc
