/*
 * bhmc.h -- C ABI of libbhmc.so: the B200 (sm_100a) implementation of the sampler hot path of
 * sherna90/dropout_hamiltonian_montecarlo (HMC / SGLD / SGHMC / SGD over batched chains).
 *
 * The reference has no FFI today: its plugin boundary is two Python duck-typed seams
 * (SURVEY.md 8(b)).  Each entry point below names the reference interface it replaces
 * (paths relative to the reference root).  INTEGRATION.md shows the ctypes stub a maintainer
 * of the reference would add.
 *
 * Conventions
 *   - every function returns an int status: 0 = ok, <0 = error (see BHMC_ERR_*);
 *     bhmc_last_error() returns a human readable message for the calling thread's last error.
 *   - "dev" pointers are caller-owned CUDA device pointers (e.g. torch tensors); the
 *     library never frees them and only keeps those documented as "bound".
 *   - chain state is fp32, one row per chain: q[c*ld + i], i < P, ld >= P, ld % 4 == 0.
 *     For the softmax model P = (D+1)*K: 'weights' [D,K] row-major first, then 'bias' [K]
 *     (the flattening of the reference's start_p dict {'weights','bias'}).
 *   - all work is enqueued on the context's stream and is asynchronous unless stated.
 *   - a context is not re-entrant; different contexts may be used from different threads.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef BHMC_H_
#define BHMC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BHMC_OK 0
#define BHMC_ERR_ARG (-1)
#define BHMC_ERR_CUDA (-2)
#define BHMC_ERR_STATE (-3)
#define BHMC_ERR_UNSUPPORTED (-4)
#define BHMC_ERR_NOMEM (-5)

/* arithmetic used for the two contractions X.W and X^T.(P-Y) */
#define BHMC_PREC_FP32 0   /* fp32 FMA on CUDA cores (exact-fp32 checker path)                */
#define BHMC_PREC_BF16X3 1 /* tcgen05 bf16 hi/lo split, 3 MMAs per product, fp32 accumulate    */
#define BHMC_PREC_BF16 2   /* tcgen05 single bf16 pass, fp32 accumulate ("fast")               */

/* sampler kinds -- hamiltonian/inference/cpu/{hmc,sgld,sghmc,sgd}.py */
#define BHMC_KIND_HMC 0
#define BHMC_KIND_SGLD 1
#define BHMC_KIND_SGHMC 2
#define BHMC_KIND_SGD 3

/* softmax log-prior variants: models/cpu/softmax.py:22-30 vs models/gpu/softmax.py:29-39 */
#define BHMC_PRIOR_CPU 0
#define BHMC_PRIOR_GPU 1

#define BHMC_MAX_VARS 8

typedef struct bhmc_ctx bhmc_ctx;
typedef struct bhmc_model bhmc_model;
typedef struct bhmc_sampler bhmc_sampler;

int bhmc_version(void);
const char* bhmc_last_error(void);

/* ---- context: one per (device, stream) ------------------------------------------------- */
int bhmc_ctx_create(int device, void* cuda_stream, bhmc_ctx** out);
int bhmc_ctx_destroy(bhmc_ctx* ctx);
int bhmc_ctx_sync(bhmc_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t bhmc_ctx_launch_count(const bhmc_ctx* ctx);
/* time (ms, CUDA events on the context stream) and launch count of the named kernel group
 * since the last reset; groups: 0 = grad fwd, 1 = grad bwd, 2 = prep, 3 = update/accept */
int bhmc_ctx_kernel_time(bhmc_ctx* ctx, int group, double* ms, int64_t* launches);
/* enable: 0 = off, 1 = every kernel group, 2 = the two GEMM groups only (fewer event records on the stream) */
int bhmc_ctx_timing(bhmc_ctx* ctx, int enable);
/* bracket only every stride-th launch of a group with events (default 1); bhmc_ctx_kernel_units returns the work
 * units (chains) of exactly the bracketed launches of the sampler drivers since the last bhmc_ctx_timing call */
int bhmc_ctx_timing_stride(bhmc_ctx* ctx, int stride);
int bhmc_ctx_kernel_units(bhmc_ctx* ctx, int group, double* units);

/* ---- model protocol (seam 1): replaces softmax.grad / log_likelihood /
 *      negative_log_posterior, hamiltonian/models/cpu/softmax.py:45-79 ---------------------- */
int bhmc_softmax_create(bhmc_ctx* ctx, int64_t n_rows, int32_t n_features, int32_t n_classes,
                        float alpha, int32_t prior_variant, bhmc_model** out);
/* logistic regression, hamiltonian/models/cpu/logistic.py:15-72: parameters weights[D] (the reference's [D,1]) and
 * bias[1]; labels are 0/1.  grad = X^T(sigmoid(z) - y) + alpha*theta (:24-40), log-lik = sum y log s + (1-y) log(1-s)
 * (:64-72), log-prior = dim/2 log(alpha/2pi) - alpha/2 |theta|^2 (:15-21) so the Metropolis energy carries the quadratic
 * term.  Evaluated by the softmax kernels as a two-class softmax with class 0 pinned to zero; bind data with
 * bhmc_softmax_bind_data / _host. */
int bhmc_logistic_create(bhmc_ctx* ctx, int64_t n_rows, int32_t n_features, float alpha, bhmc_model** out);
/* bind X [n_rows, n_features] fp32 row-major and integer labels [n_rows] (the argmax of the
 * reference's one-hot y_train).  The library builds its own resident bf16 hi/lo operand copies
 * (precision_mask: bit i set = prepare BHMC_PREC_i).  X_dev/labels_dev stay bound (needed by
 * BHMC_PREC_FP32) until the model is destroyed or re-bound. */
int bhmc_softmax_bind_data(bhmc_model* m, const float* X_dev, const int32_t* labels_dev,
                           int32_t precision_mask);
/* same from HOST memory: the library owns the device copies (this is the path bench.py's
 * end-to-end number goes through: the H2D copy happens inside this call) */
int bhmc_softmax_bind_data_host(bhmc_model* m, const float* X_host, const int32_t* labels_host,
                                int32_t precision_mask);
/* What the last bind found about X (tensor-core precisions only): *exact = 1 if x_scale * X is exactly representable in
 * bf16 for every element (x_scale = 1: binary / small-integer features; 255: 8-bit pixels / 255, the reference's MNIST
 * input, hamiltonian/utils.py scaling).  BHMC_PREC_BF16X3 then needs no lo copy of X and issues 2 MMAs per product
 * instead of 3, with 1/x_scale folded into the W operand and the gradient -- same result up to fp32 rounding.
 * Environment BHMC_X_EXACT=0 turns the check off. */
int bhmc_softmax_operand_info(bhmc_model* m, int32_t* exact, float* x_scale);
/* 2-D (d-dimensional) Gaussian target, hamiltonian/models/cpu/mvn_gaussian.py:14-31.
 * cov_inv: [dim,dim] row-major (host), logdet = log det(cov). */
int bhmc_mvn_create(bhmc_ctx* ctx, int32_t dim, const double* mu_host, const double* cov_inv_host,
                    double logdet, bhmc_model** out);
/* Dropout MLP n_in -> n_mid -> n_mid -> n_out with chain-private weights, hamiltonian/models/gpu/mlp.py:19-82.
 * Chain row = /l1/W [mid,in] | /l1/b | /l2/W [mid,mid] | /l2/b | /l3/W [out,mid] | /l3/b (Chainer namedparams
 * order, W stored (out,in)).  grad = d(mean CE)/d theta + alpha/2 theta (mlp.py:63); the model scalar is the mean CE
 * loss (mlp.py:66-78); NLP = loss - alpha/2 sum_v |theta_v|^2/dim_v (mlp.py:40-45,80-82).  Dropout keep-masks
 * (ratio on the two hidden pre-activations and on the last hidden output, mlp.py:29-31) come from Philox keyed
 * (seed, global chain id, evaluation counter, layer) -- a fresh mask per evaluation like Chainer -- or from
 * bhmc_mlp_set_masks (tests).  labels are int32 class ids (mlp.py:52). */
int bhmc_mlp_create(bhmc_ctx* ctx, int64_t n_rows, int32_t n_in, int32_t n_mid, int32_t n_out, float alpha,
                    float dropout_ratio, uint64_t seed, int64_t chain_id0, bhmc_model** out);
int bhmc_mlp_bind_data(bhmc_model* m, const float* X, const int32_t* labels, int32_t is_host);
/* injected keep-masks, DEVICE uint8 [3][n_chains][batch_rows][n_mid] (NULL = back to Philox) */
int bhmc_mlp_set_masks(bhmc_model* m, const uint8_t* masks_dev);
/* mlp.predict (hamiltonian/models/gpu/mlp.py:84-95) on caller rows X_dev [nrows, n_in]: forward pass of every chain with
 * dropout ON (the reference never switches it off), then softmax / argmax.  probs_dev [n_chains, nrows, n_out] fp32
 * and/or labels_dev [n_chains, nrows] int32 (either may be NULL). */
int bhmc_mlp_predict(bhmc_model* m, const float* q_dev, int32_t n_chains, int64_t ld, const float* X_dev, int64_t nrows,
                     int32_t precision, float* probs_dev, int32_t* labels_dev);
/* row-sharded data (multi-GPU full-batch HMC): this model holds n_rows of n_global_rows; energies are
 * normalised by the global count (softmax.py:79 divides by the size of the whole X_train) and the
 * log-prior constant uses alpha_global (the model itself was created with alpha_global / n_ranks so that the
 * all-reduced gradient carries the prior term exactly once) */
int bhmc_model_set_global_rows(bhmc_model* m, int64_t n_global_rows, float alpha_global);
int bhmc_model_destroy(bhmc_model* m);
int64_t bhmc_model_n_params(const bhmc_model* m);
int32_t bhmc_model_n_vars(const bhmc_model* m);
/* offsets / lengths of the variables inside a chain row, in canonical order */
int bhmc_model_var_layout(const bhmc_model* m, int64_t* offsets, int64_t* lengths);

/* grad of the potential for C chains over rows [row0, row0+nrows): g[c] = X^T(P-Y) + alpha q
 * (a SUM over rows, softmax.py:54-60); stat[c] = log-likelihood (softmax.py:63-72). */
int bhmc_model_grad(bhmc_model* m, const float* q_dev, int32_t n_chains, int64_t ld, int64_t row0,
                    int64_t nrows, int32_t precision, float* g_dev, double* loglik_dev);
int bhmc_model_loglik(bhmc_model* m, const float* q_dev, int32_t n_chains, int64_t ld, int64_t row0,
                      int64_t nrows, int32_t precision, double* loglik_dev);
/* negative_log_posterior (softmax.py:74-79): nlp[c] = -(LL + log_prior)/nrows */
int bhmc_model_nlp(bhmc_model* m, const float* q_dev, int32_t n_chains, int64_t ld, int64_t row0,
                   int64_t nrows, int32_t precision, double* nlp_dev);
/* softmax.net / predict (softmax.py:38-43,82-89) on caller rows X_dev [nrows, D]:
 * probs_dev [n_chains, nrows, K] fp32 (may be NULL), labels_dev [n_chains, nrows] (may be NULL) */
int bhmc_softmax_predict(bhmc_model* m, const float* q_dev, int32_t n_chains, int64_t ld,
                         const float* X_dev, int64_t nrows, float* probs_dev, int32_t* labels_dev);

/* ---- counter-based RNG (Philox4x32-10), exposed for tests ------------------------------ */
/* out[c*ld + i] = N(0,1) keyed by (seed, chain_id0 + c, stream, i) */
int bhmc_philox_normal(bhmc_ctx* ctx, float* out_dev, int32_t n_chains, int64_t P, int64_t ld,
                       uint64_t seed, int64_t chain_id0, uint32_t stream_lo, uint32_t stream_hi);
/* host-side uniform in [0,1) of the same generator (used for path length / accept draws) */
double bhmc_philox_uniform_host(uint64_t seed, int64_t chain_id, uint32_t stream_lo, uint32_t stream_hi);
/* raw Philox4x32-10 block (host), for known-answer tests against the Random123 vectors */
void bhmc_philox4x32_host(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]);

/* ---- sampler (seam 2): replaces hmc.step/sample (inference/cpu/hmc.py:39-119),
 *      sgmcmc.sample + sgld.step / sghmc.step (sgmcmc.py:40-86, sgld.py:31-46,
 *      sghmc.py:19-39) and sgd.fit (sgd.py:25-45) ---------------------------------------- */
typedef struct bhmc_sampler_config {
  int32_t kind;          /* BHMC_KIND_*                                                      */
  int32_t n_chains;      /* chains batched on this device                                    */
  int64_t chain_id0;     /* global id of chain 0 (Philox streams are keyed by global id)     */
  uint64_t seed;
  int32_t precision;     /* BHMC_PREC_*                                                      */
  int32_t n_sweep;       /* number of Gauss-Seidel groups (hmc.py:50); 1 = joint             */
  int64_t sweep_off[BHMC_MAX_VARS]; /* slice of the chain row moved by group i               */
  int64_t sweep_len[BHMC_MAX_VARS];
  int32_t shared_path;   /* 1: one path-length draw per step shared by all chains            */
  int32_t leapfrog;      /* 0: reference kicks (eps/2 then eps, hmc.py:51-54); 1: eps/2,eps/2 */
  int32_t sghmc_descent; /* 0: literal '+eps*grad' (sghmc.py:34); 1: '-eps*grad'             */
  int32_t reject_nan;    /* 0: literal builtin-min semantics (NaN -> A=1, accepted); 1: reject */
  int32_t reserved[4];
} bhmc_sampler_config;

int bhmc_sampler_create(bhmc_ctx* ctx, bhmc_model* model, const bhmc_sampler_config* cfg,
                        bhmc_sampler** out);
int bhmc_sampler_destroy(bhmc_sampler* s);
/* Hook invoked (on the host, in stream order) after every gradient / log-likelihood evaluation of the
 * drivers with the device buffers it produced: g_dev [n_rows_active, ld] (NULL for log-lik only) and
 * stat_dev [n_rows_active].  Row-sharded runs use it to enqueue the NCCL all-reduce (sum) of both on the
 * context stream.  Return non-zero to abort the run. */
typedef int (*bhmc_grad_hook)(void* user, float* g_dev, double* stat_dev, int32_t n_rows_active, int64_t ld);
int bhmc_sampler_set_grad_hook(bhmc_sampler* s, bhmc_grad_hook hook, void* user);
int64_t bhmc_sampler_ld(const bhmc_sampler* s);
/* device pointers of the resident chain state ([n_chains, ld] fp32): 0=q 1=p 2=grad */
int bhmc_sampler_state_ptr(bhmc_sampler* s, int32_t which, float** out_dev);
/* copy state in/out; src/dst are [n_chains, P] fp32 compact; is_host selects memcpy kind */
int bhmc_sampler_set_q(bhmc_sampler* s, const float* src, int32_t is_host);
int bhmc_sampler_get(bhmc_sampler* s, int32_t which, float* dst, int32_t is_host);

typedef struct bhmc_hmc_run {
  int32_t n_steps;
  double step_size;       /* epsilon                                                         */
  double path_length;
  int64_t row0, nrows;    /* data window (full batch for HMC)                                */
  int64_t step0;          /* global index of the first step (Philox stream counter)          */
  /* injected draws (parity mode); NULL = Philox.  z_*: DEVICE fp32, u_*: HOST fp64          */
  const float* z_momentum_dev; /* [n_steps, n_chains, P]  N(0,1), hmc.py:41,82-87            */
  const double* u_path_host;   /* [n_steps, n_chains]     np.random.rand(), hmc.py:46         */
  const double* u_accept_host; /* [n_steps, n_chains]     np.random.rand(), hmc.py:61         */
  const float* z_noise_dev;    /* SGHMC only: [n_steps, Lmax-1, n_chains, P], sghmc.py:31     */
  int64_t z_noise_iters;       /* Lmax-1 rows available per step in z_noise_dev               */
  /* outputs, DEVICE, all optional (NULL = skip)                                            */
  float* samples_dev;     /* [n_steps, n_chains, P] q after each step (hmc.py:113-114)       */
  double* loss_dev;       /* [n_steps, n_chains]  NLP(q) after each step (hmc.py:112)        */
  double* accept_prob_dev;/* [n_steps, n_chains]                                             */
  int32_t* accepted_dev;  /* [n_steps, n_chains]                                             */
  /* outputs, HOST                                                                          */
  int64_t n_grad_evals;   /* chain-gradient evaluations applied (masked chains excluded)     */
  int64_t n_grad_launched;/* chain-gradient evaluations computed (incl. masked)              */
  /* schedule of the chains across the n_steps transitions of this call:
   *   BHMC_SCHED_AUTO (0)      streaming for HMC with per-chain path lengths and n_steps >= 2, else lockstep
   *   BHMC_SCHED_LOCKSTEP (1)  every chain starts transition t together; short trajectories idle (compacted away)
   *   BHMC_SCHED_STREAMING (2) asynchronous chains: a chain starts its next transition as soon as its own trajectory
   *                            ends, so every gradient launch carries all unfinished chains.  Same draws (Philox keyed
   *                            by chain and step, or the same injected tapes), same arithmetic -> same samples. */
  int32_t schedule;
  int32_t n_phases;       /* out: gradient launches issued                                   */
} bhmc_hmc_run;
#define BHMC_SCHED_AUTO 0
#define BHMC_SCHED_LOCKSTEP 1
#define BHMC_SCHED_STREAMING 2
/* run n_steps transitions of HMC (kind HMC) or SGHMC (kind SGHMC) for all chains */
int bhmc_sampler_hmc_run(bhmc_sampler* s, bhmc_hmc_run* run);

typedef struct bhmc_sg_run {
  int32_t epochs;          /* sampling epochs (one stored sample per epoch, sgmcmc.py:79-81)   */
  int32_t burnin;          /* burn-in epochs at constant step size (sgmcmc.py:55-63)          */
  int64_t batch_size;      /* sequential windows, remainder dropped (sgmcmc.py:34-38)         */
  int64_t n_rows;          /* rows of the bound data to iterate over                           */
  double step_size;        /* eps0; schedule eps0/(1+j*eps0) after batch j (sgmcmc.py:73,88)   */
  double gamma;            /* SGD only: heavy-ball momentum (sgd.py:40)                        */
  int64_t step0;
  const float* z_dev;      /* [(burnin+epochs)*n_batches, n_chains, P] N(0,1) or NULL          */
  float* samples_dev;      /* [epochs, n_chains, P]                                            */
  double* logp_dev;        /* [epochs, n_chains] NLP(q, last batch) (sgmcmc.py:79)             */
  int64_t n_grad_evals;    /* out */
  double final_step_size;  /* out */
  /* sgd.fit_dropout (sgd.py:47-70): the gradient of minibatch k sees X_batch * Z_k, Z ~ Bernoulli(dropout_keep)
   * (no rescaling), and the per-epoch loss is -log_likelihood(theta, last batch) on the unmasked rows (:67).
   * dropout_keep = 0 disables.  mask_dev: injected keep flags [(burnin+epochs)*n_batches, batch_size, D] uint8,
   * or NULL for in-kernel Philox draws keyed (seed, minibatch, row, feature). */
  double dropout_keep;
  const uint8_t* mask_dev;
  /* Chunked calls (one epoch, or a bounded number of minibatches' worth of injected noise, per call): the step size of
   * this call's FIRST minibatch.  0 = step_size.  The reference re-assigns eps after every batch of a sampling epoch
   * (sgmcmc.py:72-73), so batch 0 of epoch e >= 1 runs at lr(n_batches - 1): pass the previous call's final_step_size. */
  double first_step_size;
  /* SGD only.  0 (default): the heavy-ball momentum starts from zero, as every sgd.fit / fit_dropout call of the
   * reference does (sgd.py:35,57: momentum = zeros_like(par)); 1: continue from the momentum of the previous call. */
  int32_t keep_momentum;
  int32_t reserved_;
} bhmc_sg_run;
/* SGLD (kind SGLD) / SGD (kind SGD) epochs over sequential minibatches */
int bhmc_sampler_sg_run(bhmc_sampler* s, bhmc_sg_run* run);

/* ---- multi-GPU, rows sharded (full-batch HMC on large N; no reference counterpart -- the reference only has host
 *      multiprocessing, hamiltonian/inference/cpu/hmc_multicore.py).  Every rank binds ITS rows and runs ALL chains;
 *      after each evaluation the gradient and the log-likelihoods are summed over the ranks by ONE grouped NCCL
 *      all-reduce enqueued by the C driver on the context stream (no host callback).  libnccl.so.2 is resolved with
 *      dlopen at first use; single-GPU use never loads it. ------------------------------------------------------ */
#define BHMC_COMM_ID_BYTES 128
typedef struct bhmc_comm bhmc_comm;
/* rank 0: create the rendezvous id (ncclGetUniqueId) and hand its 128 bytes to the other ranks by any means */
int bhmc_comm_unique_id(uint8_t* id_out);
/* collective over all ranks: ncclCommInitRank on the context's device */
int bhmc_comm_create(bhmc_ctx* ctx, const uint8_t* id, int32_t rank, int32_t world, bhmc_comm** out);
/* adopt a communicator the host program already owns (ncclComm_t); it is not destroyed by bhmc_comm_destroy */
int bhmc_comm_wrap(bhmc_ctx* ctx, void* nccl_comm, int32_t rank, int32_t world, bhmc_comm** out);
int bhmc_comm_destroy(bhmc_comm* c);
int32_t bhmc_comm_world(const bhmc_comm* c);
int bhmc_nccl_version(void); /* 0 when libnccl cannot be loaded */
/* sum over ranks, in place: g_dev [g_count] fp32 and stat_dev [n_stat] fp64 (either may be NULL); one grouped call */
int bhmc_allreduce_grad(bhmc_ctx* ctx, void* nccl_comm, float* g_dev, int64_t g_count, double* stat_dev, int32_t n_stat);
int bhmc_comm_allreduce(bhmc_comm* c, float* g_dev, int64_t g_count, double* stat_dev, int32_t n_stat);
/* from now on every gradient / log-likelihood evaluation of this sampler's drivers is followed by the all-reduce of
 * the rows it produced (NULL detaches).  The model must have been created with alpha / world and told the global row
 * count (bhmc_model_set_global_rows) so that the prior term and the energies come out once. */
int bhmc_sampler_set_row_comm(bhmc_sampler* s, bhmc_comm* comm);

/* ---- test hook, host only (no device needed): compiles the streaming schedule for path lengths L[n_steps][n_chains].
 * Call with code1 == NULL to get n_phases (gradient launches J) and n_grad_evals, then with code1/code2/step1/step2 of
 * (J+1)*n_chains entries each (ops of the update before / after the begins of a phase and the transition index they
 * refer to), perm[n_chains] (row -> chain), rows_el[J+1], rows_grad[J+1] (active rows per phase / launch). */
int bhmc_stream_plan_host(const int32_t* L, int32_t n_chains, int32_t n_steps, int32_t n_sweep, int64_t* n_phases,
                          int64_t* n_grad_evals, uint32_t* code1, uint32_t* code2, int32_t* step1, int32_t* step2,
                          int32_t* perm, int32_t* rows_el, int32_t* rows_grad);

/* ---- measurement helper: time one fused update kernel in isolation (CUDA events on the context stream).
 * which: 0 = HMC kick+drift (20 B/param), 1 = SGHMC friction+Philox noise+drift (20 B/param),
 *        2 = SGLD with Philox noise (16 B/param), 3 = accept/select + sample sink (20 B/param),
 *        4 = momentum draw (Philox) + proposal copy + kinetic energy (16 B/param).
 * Allocates its own [n_chains, P] buffers; returns the mean milliseconds per launch over reps. */
int bhmc_bench_update(bhmc_ctx* ctx, int32_t which, int32_t n_chains, int64_t P, int32_t reps, double* ms_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* BHMC_H_ */
