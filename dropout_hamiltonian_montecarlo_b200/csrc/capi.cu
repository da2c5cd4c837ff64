// C ABI of libbhmc.so (include/bhmc.h): context, model handles and the sampler drivers.
// The drivers restate the *control flow* of the reference samplers
//   hmc.step / hmc.sample          hamiltonian/inference/cpu/hmc.py:39-119
//   sghmc.step                     hamiltonian/inference/cpu/sghmc.py:19-39
//   sgmcmc.sample + sgld.step      hamiltonian/inference/cpu/sgmcmc.py:40-89, sgld.py:31-46
//   sgd.fit                        hamiltonian/inference/cpu/sgd.py:25-45
// for C chains at once, enqueueing kernels on one stream with no host round-trip inside a step.
#include <nvtx3/nvToolsExt.h>
#include <stdarg.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <cmath>
#include <new>

#include "internal.cuh"
#include "philox.cuh"

namespace bhmc {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace bhmc

using namespace bhmc;

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
int bhmc_ctx::get_scratch(int slot, size_t bytes, void** out) {
  if (bytes > scratch_bytes[slot]) {
    // grow-only; free synchronises with outstanding work on the device
    if (scratch[slot]) BHMC_CUDA_OK(cudaFree(scratch[slot]));
    scratch[slot] = nullptr;
    scratch_bytes[slot] = 0;
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&scratch[slot], want);
    if (e != cudaSuccess) {
      set_error("cudaMalloc of %zu scratch bytes failed: %s", want, cudaGetErrorString(e));
      return BHMC_ERR_NOMEM;
    }
    scratch_bytes[slot] = want;
  }
  *out = scratch[slot];
  return BHMC_OK;
}

int bhmc_ctx::get_pinned(size_t bytes, void** out) {
  if (bytes > pinned_bytes) {
    // page-locking is slow and erratic (measured: up to 0.7 s for a few MB on a busy host) and a re-allocation lands
    // inside whatever call needed the larger buffer: start with 32 MB and grow geometrically
    bytes = std::max(bytes + bytes / 2, (size_t)32 << 20);
    if (pinned) {
      BHMC_CUDA_OK(cudaStreamSynchronize(stream));
      BHMC_CUDA_OK(cudaFreeHost(pinned));
    }
    pinned = nullptr;
    pinned_bytes = 0;
    BHMC_CUDA_OK(cudaMallocHost(&pinned, bytes));
    pinned_bytes = bytes;
  }
  *out = pinned;
  return BHMC_OK;
}

// NVTX ranges (tracing).  The header-only NVTX v3 resolves the injection library lazily: without a profiler
// attached a push / pop is a load and a branch.  BHMC_NVTX=1 switches them on: one range per library run
// (bhmc_sampler_hmc_run / bhmc_sampler_sg_run) and one per kernel group of a gradient evaluation (forward GEMM, backward
// GEMM + reduce, operand preparation, update), so that `ncu --nvtx --nvtx-include "bhmc.bwd/"` or a timeline tool can cut
// the launch stream the way DESIGN.md section 4 names it.
static bool nvtx_on() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BHMC_NVTX");
    v = e ? (atoi(e) != 0) : 0;
  }
  return v != 0;
}
struct NvtxRange {
  bool on;
  explicit NvtxRange(const char* name) : on(nvtx_on()) {
    if (on) nvtxRangePushA(name);
  }
  ~NvtxRange() {
    if (on) nvtxRangePop();
  }
};
static const char* const kGroupNames[bhmc::KG_COUNT] = {"bhmc.fwd", "bhmc.bwd", "bhmc.prep", "bhmc.update"};

void bhmc_ctx::begin_group(int g) {
  if (nvtx_on()) nvtxRangePushA(kGroupNames[g]);
  sampled[g] = false;
  if (!timing || (timing == 2 && g >= KG_PREP)) return;
  if ((seen[g]++ % timing_stride) != 0) return;
  sampled[g] = true;
  units_acc[g] += (double)cur_units;
  if (used[g] == pool[g].size()) {
    // flushing synchronises the stream and reads every event back (cudaEventElapsedTime, ~10 us each): the GPU
    // idles for 0.1-0.7 s, which would land inside whatever region is being timed -> only when the pool is huge
    if (pool[g].size() >= 65536) {
      flush_timing();
    } else {
      EventPair ep;
      cudaEventCreate(&ep.a);
      cudaEventCreate(&ep.b);
      pool[g].push_back(ep);
    }
  }
  cudaEventRecord(pool[g][used[g]].a, stream);
}

void bhmc_ctx::end_group(int g) {
  if (nvtx_on()) nvtxRangePop();
  if (!sampled[g]) return;
  cudaEventRecord(pool[g][used[g]].b, stream);
  used[g]++;
}

int bhmc_ctx::flush_timing() {
  BHMC_CUDA_OK(cudaStreamSynchronize(stream));
  for (int g = 0; g < KG_COUNT; ++g) {
    for (size_t i = 0; i < used[g]; ++i) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, pool[g][i].a, pool[g][i].b) == cudaSuccess) {
        ms_acc[g] += ms;
        n_acc[g]++;
      }
    }
    used[g] = 0;
  }
  return BHMC_OK;
}

extern "C" {

int bhmc_version(void) { return 100; }
const char* bhmc_last_error(void) { return g_err; }

int bhmc_ctx_create(int device, void* cuda_stream, bhmc_ctx** out) {
  BHMC_CHECK_ARG(out, "out is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("no CUDA device available (%s): libbhmc has no CPU fallback", cudaGetErrorString(e));
    return BHMC_ERR_CUDA;
  }
  BHMC_CHECK_ARG(device >= 0 && device < n, "device %d out of range (%d devices)", device, n);
  BHMC_CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  BHMC_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("device %d is sm_%d%d; libbhmc is built for sm_100a (B200) only", device, prop.major, prop.minor);
    return BHMC_ERR_UNSUPPORTED;
  }
  bhmc_ctx* c = new (std::nothrow) bhmc_ctx();
  if (!c) return BHMC_ERR_NOMEM;
  c->device = device;
  c->stream = (cudaStream_t)cuda_stream;
  c->sm_count = prop.multiProcessorCount;
  *out = c;
  return BHMC_OK;
}

int bhmc_ctx_destroy(bhmc_ctx* ctx) {
  if (!ctx) return BHMC_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < 16; ++i) cudaFree(ctx->scratch[i]);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->pinned_ev) cudaEventDestroy(ctx->pinned_ev);
  for (int g = 0; g < KG_COUNT; ++g)
    for (auto& ep : ctx->pool[g]) {
      cudaEventDestroy(ep.a);
      cudaEventDestroy(ep.b);
    }
  delete ctx;
  return BHMC_OK;
}

int bhmc_ctx_sync(bhmc_ctx* ctx) {
  BHMC_CHECK_ARG(ctx, "ctx is NULL");
  BHMC_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return BHMC_OK;
}

int64_t bhmc_ctx_launch_count(const bhmc_ctx* ctx) { return ctx ? ctx->launches : 0; }

int bhmc_ctx_timing(bhmc_ctx* ctx, int enable) {
  BHMC_CHECK_ARG(ctx, "ctx is NULL");
  BHMC_TRY(ctx->flush_timing());
  ctx->timing = enable;  // 1 = every kernel group; 2 = the GEMM groups only (every event record between two kernels
                         // costs ~1-2 us of overlap on the stream: half the records, half the perturbation)
  for (int g = 0; g < KG_COUNT; ++g) ctx->ms_acc[g] = 0, ctx->n_acc[g] = 0, ctx->seen[g] = 0, ctx->units_acc[g] = 0;
  if (ctx->timing) {
    // create the whole event pool now: cudaEventCreate costs microseconds and would otherwise be paid lazily
    // inside the region being timed (8 events per gradient evaluation)
    BHMC_CUDA_OK(cudaSetDevice(ctx->device));
    for (int g = 0; g < KG_COUNT; ++g)
      while (ctx->pool[g].size() < 8192) {
        EventPair ep;
        BHMC_CUDA_OK(cudaEventCreate(&ep.a));
        BHMC_CUDA_OK(cudaEventCreate(&ep.b));
        ctx->pool[g].push_back(ep);
      }
  }
  return BHMC_OK;
}

int bhmc_ctx_timing_stride(bhmc_ctx* ctx, int stride) {
  BHMC_CHECK_ARG(ctx && stride >= 1, "bad argument");
  ctx->timing_stride = stride;
  return BHMC_OK;
}

int bhmc_ctx_kernel_units(bhmc_ctx* ctx, int group, double* units) {
  BHMC_CHECK_ARG(ctx && units && group >= 0 && group < KG_COUNT, "bad argument");
  *units = ctx->units_acc[group];
  return BHMC_OK;
}

int bhmc_ctx_kernel_time(bhmc_ctx* ctx, int group, double* ms, int64_t* launches) {
  BHMC_CHECK_ARG(ctx && group >= 0 && group < KG_COUNT, "bad group");
  BHMC_TRY(ctx->flush_timing());
  if (ms) *ms = ctx->ms_acc[group];
  if (launches) *launches = ctx->n_acc[group];
  return BHMC_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// models
// ------------------------------------------------------------------------------------------------
struct bhmc_model {
  ModelBase* impl = nullptr;
};

namespace bhmc {

// logistic regression (models/cpu/logistic.py) rides on the softmax kernels: sigmoid(z) is the class-1 probability of
// a two-class softmax whose class-0 weights and bias are pinned to zero.  The chain state keeps the reference layout
// (weights[D], bias) and is expanded to / contracted from the two-class layout around every evaluation.
__global__ void k_logistic_expand(const float* __restrict__ q, int64_t ld, int D, float* __restrict__ q2, int64_t ld2) {
  const int c = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // index into the two-class row
  if (i >= ld2) return;
  float v = 0.f;
  if ((i & 1) && (i >> 1) <= D) v = q[(int64_t)c * ld + (i >> 1)];  // weights[d,1] = w[d]; bias[1] = b
  q2[(int64_t)c * ld2 + i] = v;
}
__global__ void k_logistic_contract(const float* __restrict__ g2, int64_t ld2, int D, float* __restrict__ g, int64_t ld) {
  const int c = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ld) return;
  g[(int64_t)c * ld + i] = i <= D ? g2[(int64_t)c * ld2 + 2 * i + 1] : 0.f;
}

// Xb[r, j] = X[row0 + r, j] * keep(r, j): the masked minibatch of sgd.fit_dropout (sgd.py:60-61)
__global__ void k_input_dropout(const float* __restrict__ X, int D, int64_t n, const uint8_t* __restrict__ mask, float keep,
                                uint64_t seed, uint32_t stream, float* __restrict__ out) {
  const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  float k[4];
  if (mask) {
#pragma unroll
    for (int j = 0; j < 4; ++j) k[j] = (i4 + j < n && mask[i4 + j]) ? 1.f : 0.f;
  } else {
    U4 c{(uint32_t)(i4 >> 2), (uint32_t)((uint64_t)i4 >> 34), stream, TAG_DROPOUT | 0x800000u};
    U4 r = philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    k[0] = u01_open(r.x) < keep ? 1.f : 0.f;
    k[1] = u01_open(r.y) < keep ? 1.f : 0.f;
    k[2] = u01_open(r.z) < keep ? 1.f : 0.f;
    k[3] = u01_open(r.w) < keep ? 1.f : 0.f;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (i4 + j < n) out[i4 + j] = X[i4 + j] * k[j];
}

struct SoftmaxModel : ModelBase {
  SoftmaxData d;
  SoftmaxData db;            // masked-minibatch view used by grad_input_dropout (owns its operand copies)
  float* Xb = nullptr;       // [batch, D] fp32 masked rows
  int64_t Xb_rows = 0;
  bool logistic = false;
  float alpha = 0.f;
  int prior = BHMC_PRIOR_CPU;
  float* X_owned = nullptr;
  int32_t* y_owned = nullptr;

  ~SoftmaxModel() override {
    tc_softmax_release(d);
    tc_softmax_release(db);
    cudaFree(Xb);
    cudaFree(X_owned);
    cudaFree(y_owned);
  }
  int64_t default_rows() const override { return d.N; }
  int64_t n_features() const override { return d.D; }
  int64_t global_rows = 0;
  float alpha_energy = -1.f;  // alpha of the log-prior constant when it differs from the gradient's (row shards)

  ZCache zcache;
  bool cheap_slice(int64_t off, int64_t) const override { return !logistic && off >= (int64_t)d.D * d.K; }  // bias only

  int grad(const float* q, int C, int64_t ld, int64_t row0, int64_t nrows, int prec, float* g, double* stat,
           uint32_t hint) override {
    if (logistic) {
      const int64_t ld2 = round_up(2 * ((int64_t)d.D + 1), 4);
      void* buf = nullptr;
      BHMC_TRY(ctx->get_scratch(7, sizeof(float) * 2 * (size_t)C * ld2, &buf));
      float* q2 = (float*)buf;
      float* g2 = g ? q2 + (size_t)C * ld2 : nullptr;
      k_logistic_expand<<<dim3((unsigned)ceil_div(ld2, 256), C), 256, 0, ctx->stream>>>(q, ld, d.D, q2, ld2);
      ctx->launches++;
      BHMC_TRY(grad_softmax(q2, C, ld2, row0, nrows, prec, g2, stat));
      if (g) {
        k_logistic_contract<<<dim3((unsigned)ceil_div(ld, 256), C), 256, 0, ctx->stream>>>(g2, ld2, d.D, g, ld);
        ctx->launches++;
      }
      BHMC_CUDA_OK(cudaGetLastError());
      return BHMC_OK;
    }
    return grad_softmax(q, C, ld, row0, nrows, prec, g, stat, hint);
  }
  int grad_input_dropout(const float* q, int C, int64_t ld, int64_t row0, int64_t nrows, int prec, float* g,
                         double* stat, const uint8_t* mask, float keep, uint64_t seed, uint32_t stream) override {
    BHMC_CHECK_ARG(row0 >= 0 && nrows > 0 && row0 + nrows <= d.N, "row window outside the bound rows");
    if (Xb_rows < nrows) {
      cudaFree(Xb);
      Xb = nullptr;
      BHMC_CUDA_OK(cudaMalloc(&Xb, sizeof(float) * (size_t)nrows * d.D));
      Xb_rows = nrows;
    }
    const int64_t n = nrows * d.D;
    k_input_dropout<<<(unsigned)ceil_div(ceil_div(n, 4), 256), 256, 0, ctx->stream>>>(d.X + row0 * d.D, d.D, n, mask, keep,
                                                                                   seed, stream, Xb);
    ctx->launches++;
    BHMC_CUDA_OK(cudaGetLastError());
    db.N = nrows;
    db.D = d.D;
    db.K = d.K;
    db.X = Xb;
    db.labels = d.labels + row0;
    // no exact-operand check on the masked rows (it would cost a host sync per minibatch; x / keep is not exact anyway)
    if (prec != BHMC_PREC_FP32) BHMC_TRY(tc_softmax_bind(ctx, db, prec == BHMC_PREC_BF16X3, false));
    std::swap(d, db);  // evaluate on the masked view through the ordinary (logistic-aware) path
    const int rc = grad(q, C, ld, 0, nrows, prec, g, stat, 0);
    std::swap(d, db);
    return rc;
  }
  int grad_fused_step(int C, int64_t ld, int64_t row0, int64_t nrows, int prec, double* stat, const FusedStep& fs) override {
    if (logistic || prec == BHMC_PREC_FP32) return BHMC_ERR_UNSUPPORTED;
    return tc_softmax_grad(ctx, d, fs.q, C, ld, alpha, row0, nrows, nullptr, stat, prec == BHMC_PREC_BF16X3, &fs);
  }
  int sg_steps_persistent(int C, int64_t ld, int prec, const FusedStep& fs, int64_t row_first, int64_t batch, int n_steps,
                          const float* eps_dev, int64_t z_step_stride, uint64_t step0) override {
    if (logistic || prec == BHMC_PREC_FP32) return BHMC_ERR_UNSUPPORTED;
    zcache.valid = false;
    return tc_softmax_sg_persistent(ctx, d, C, ld, alpha, prec == BHMC_PREC_BF16X3, fs, row_first, batch, n_steps, eps_dev,
                                    z_step_stride, step0);
  }
  int grad_fused_stream(const float* q, int C, int64_t ld, int64_t row0, int64_t nrows, int prec, float* g, double* stat,
                        uint32_t hint, const FusedStream& fst) override {
    if (logistic || prec == BHMC_PREC_FP32) return BHMC_ERR_UNSUPPORTED;
    return grad_softmax(q, C, ld, row0, nrows, prec, g, stat, hint, &fst);
  }
  int grad_softmax(const float* q, int C, int64_t ld, int64_t row0, int64_t nrows, int prec, float* g, double* stat,
                   uint32_t hint = 0, const FusedStream* fst = nullptr) {
    static int zc_env = -1;  // BHMC_ZCACHE=0 disables the X.W cache (A/B measurements)
    if (zc_env < 0) {
      const char* e = getenv("BHMC_ZCACHE");
      zc_env = e ? atoi(e) : 1;
    }
    const int zmode = !zc_env ? ZMODE_NONE : (hint & GRAD_HINT_CHEAP_MOVE) ? ZMODE_USE : (hint & GRAD_HINT_KEEP) ? ZMODE_STORE : ZMODE_NONE;
    switch (prec) {
      case BHMC_PREC_FP32: return simt_softmax_grad(ctx, d, q, C, ld, alpha, row0, nrows, g, stat);
      case BHMC_PREC_BF16X3:
        return tc_softmax_grad(ctx, d, q, C, ld, alpha, row0, nrows, g, stat, true, nullptr, &zcache, zmode, fst,
                               (hint & GRAD_HINT_PREPARED) != 0);
      case BHMC_PREC_BF16:
        return tc_softmax_grad(ctx, d, q, C, ld, alpha, row0, nrows, g, stat, false, nullptr, &zcache, zmode, fst,
                               (hint & GRAD_HINT_PREPARED) != 0);
    }
    set_error("unknown precision %d", prec);
    return BHMC_ERR_ARG;
  }
  // NLP = -(LL + log_prior)/n  (softmax.py:74-79)
  void energy_coeffs(int64_t nrows, double* a, double* b, double* cv) const override {
    if (global_rows > 0) nrows = global_rows;  // row-sharded: the all-reduced LL covers all rows
    *a = -1.0 / (double)nrows;
    double lp = 0.0;
    for (int v = 0; v < n_vars; ++v) {
      cv[v] = 0.0;
      const double al = (double)(alpha_energy > 0.f ? alpha_energy : alpha);
      if (logistic) {  // logistic.py:15-21: dim/2 log(alpha/2pi) - alpha/2 |theta_v|^2
        lp += 0.5 * (double)var_len[v] * std::log(al / (2.0 * M_PI));
        cv[v] = 0.5 * al / (double)nrows;
      } else if (prior == BHMC_PRIOR_CPU)  // softmax.py:22-30: -(dim/2 log 2pi - dim/2 log alpha)
        lp -= 0.5 * (double)var_len[v] * std::log(2.0 * M_PI) - 0.5 * (double)var_len[v] * std::log((double)(alpha_energy > 0.f ? alpha_energy : alpha));
      else  // models/gpu/softmax.py:29-39: -alpha/2 |theta_v|^2 / dim_v
        cv[v] = 0.5 * (double)(alpha_energy > 0.f ? alpha_energy : alpha) / ((double)var_len[v] * (double)nrows);
    }
    *b = -lp / (double)nrows;
  }
};

// d-dimensional Gaussian target: one thread per chain (mvn_gaussian.py:14-31)
__global__ void k_mvn(const float* __restrict__ q, int64_t ld, int dim, const double* __restrict__ mu,
                      const double* __restrict__ cinv, float* __restrict__ g, double* __restrict__ stat, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float* x = q + (int64_t)c * ld;
  double quad = 0.0;
  for (int j = 0; j < dim; ++j) {
    double s = 0.0;
    for (int i = 0; i < dim; ++i) s += ((double)x[i] - mu[i]) * cinv[i * dim + j];  // (x-mu) Sigma^-1
    if (g) g[(int64_t)c * ld + j] = (float)s;
    quad += s * ((double)x[j] - mu[j]);
  }
  if (stat) stat[c] = quad;
}

struct MvnModel : ModelBase {
  int dim = 0;
  double logdet = 0.0;
  double* dev = nullptr;  // mu | cov_inv
  ~MvnModel() override { cudaFree(dev); }
  int grad(const float* q, int C, int64_t ld, int64_t, int64_t, int, float* g, double* stat, uint32_t) override {
    GroupTimer t(ctx, KG_FWD);
    k_mvn<<<(C + 127) / 128, 128, 0, ctx->stream>>>(q, ld, dim, dev, dev + dim, g, stat, C);
    ctx->launches++;
    BHMC_CUDA_OK(cudaGetLastError());
    return BHMC_OK;
  }
  // U = 0.5 (d log 2pi + log det + quad)
  void energy_coeffs(int64_t, double* a, double* b, double* cv) const override {
    *a = 0.5;
    *b = 0.5 * (dim * std::log(2.0 * M_PI) + logdet);
    cv[0] = 0.0;
  }
};

}  // namespace bhmc

extern "C" {

int bhmc_softmax_create(bhmc_ctx* ctx, int64_t n_rows, int32_t n_features, int32_t n_classes, float alpha,
                        int32_t prior_variant, bhmc_model** out) {
  BHMC_CHECK_ARG(ctx && out, "ctx/out is NULL");
  BHMC_CHECK_ARG(n_rows > 0 && n_features > 0 && n_classes > 0, "bad shape N=%lld D=%d K=%d", (long long)n_rows, n_features,
                 n_classes);
  BHMC_CHECK_ARG(n_rows < (1LL << 31) - 256, "row count %lld exceeds the 32-bit TMA coordinate range", (long long)n_rows);
  auto* m = new (std::nothrow) SoftmaxModel();
  if (!m) return BHMC_ERR_NOMEM;
  m->ctx = ctx;
  m->d.N = n_rows;
  m->d.D = n_features;
  m->d.K = n_classes;
  m->alpha = alpha;
  m->prior = prior_variant;
  m->P = (int64_t)(n_features + 1) * n_classes;
  m->n_vars = 2;
  m->var_off[0] = 0;
  m->var_len[0] = (int64_t)n_features * n_classes;
  m->var_off[1] = m->var_len[0];
  m->var_len[1] = n_classes;
  bhmc_model* h = new bhmc_model();
  h->impl = m;
  *out = h;
  return BHMC_OK;
}

int bhmc_logistic_create(bhmc_ctx* ctx, int64_t n_rows, int32_t n_features, float alpha, bhmc_model** out) {
  BHMC_TRY(bhmc_softmax_create(ctx, n_rows, n_features, 2, alpha, BHMC_PRIOR_CPU, out));
  auto* m = static_cast<SoftmaxModel*>((*out)->impl);
  m->logistic = true;
  m->P = (int64_t)n_features + 1;
  m->var_len[0] = n_features;
  m->var_off[1] = n_features;
  m->var_len[1] = 1;
  return BHMC_OK;
}

static int softmax_of(bhmc_model* m, SoftmaxModel** out) {
  BHMC_CHECK_ARG(m && m->impl, "model is NULL");
  auto* s = dynamic_cast<SoftmaxModel*>(m->impl);
  BHMC_CHECK_ARG(s, "model is not a softmax model");
  *out = s;
  return BHMC_OK;
}

static int bind_common(SoftmaxModel* s, int32_t mask) {
  s->zcache.valid = false;  // new data: whatever X.W the cache holds belongs to the old rows
  if (mask & ((1 << BHMC_PREC_BF16X3) | (1 << BHMC_PREC_BF16)))
    BHMC_TRY(tc_softmax_bind(s->ctx, s->d, (mask & (1 << BHMC_PREC_BF16X3)) != 0));
  return BHMC_OK;
}

int bhmc_softmax_bind_data(bhmc_model* m, const float* X_dev, const int32_t* labels_dev, int32_t precision_mask) {
  SoftmaxModel* s;
  BHMC_TRY(softmax_of(m, &s));
  BHMC_CHECK_ARG(X_dev && labels_dev, "X/labels is NULL");
  BHMC_CUDA_OK(cudaSetDevice(s->ctx->device));
  s->d.X = X_dev;
  s->d.labels = labels_dev;
  return bind_common(s, precision_mask);
}

int bhmc_softmax_bind_data_host(bhmc_model* m, const float* X_host, const int32_t* labels_host, int32_t precision_mask) {
  SoftmaxModel* s;
  BHMC_TRY(softmax_of(m, &s));
  BHMC_CHECK_ARG(X_host && labels_host, "X/labels is NULL");
  BHMC_CUDA_OK(cudaSetDevice(s->ctx->device));
  size_t xb = sizeof(float) * (size_t)s->d.N * s->d.D, yb = sizeof(int32_t) * (size_t)s->d.N;
  if (!s->X_owned) BHMC_CUDA_OK(cudaMalloc(&s->X_owned, xb));
  if (!s->y_owned) BHMC_CUDA_OK(cudaMalloc(&s->y_owned, yb));
  BHMC_CUDA_OK(cudaMemcpyAsync(s->X_owned, X_host, xb, cudaMemcpyHostToDevice, s->ctx->stream));
  BHMC_CUDA_OK(cudaMemcpyAsync(s->y_owned, labels_host, yb, cudaMemcpyHostToDevice, s->ctx->stream));
  s->d.X = s->X_owned;
  s->d.labels = s->y_owned;
  return bind_common(s, precision_mask);
}

int bhmc_softmax_operand_info(bhmc_model* m, int32_t* exact, float* x_scale) {
  SoftmaxModel* s;
  BHMC_TRY(softmax_of(m, &s));
  BHMC_CHECK_ARG(s->d.tc_ready, "no tensor-core operands bound");
  if (exact) *exact = s->d.x_exact ? 1 : 0;
  if (x_scale) *x_scale = s->d.x_scale;
  return BHMC_OK;
}

int bhmc_mvn_create(bhmc_ctx* ctx, int32_t dim, const double* mu_host, const double* cov_inv_host, double logdet,
                    bhmc_model** out) {
  BHMC_CHECK_ARG(ctx && out && mu_host && cov_inv_host, "NULL argument");
  BHMC_CHECK_ARG(dim > 0 && dim <= 64, "dim %d out of range [1,64]", dim);
  auto* m = new (std::nothrow) MvnModel();
  if (!m) return BHMC_ERR_NOMEM;
  m->ctx = ctx;
  m->dim = dim;
  m->logdet = logdet;
  m->P = dim;
  m->n_vars = 1;
  m->var_off[0] = 0;
  m->var_len[0] = dim;
  BHMC_CUDA_OK(cudaMalloc(&m->dev, sizeof(double) * (dim + dim * dim)));
  BHMC_CUDA_OK(cudaMemcpy(m->dev, mu_host, sizeof(double) * dim, cudaMemcpyHostToDevice));
  BHMC_CUDA_OK(cudaMemcpy(m->dev + dim, cov_inv_host, sizeof(double) * dim * dim, cudaMemcpyHostToDevice));
  bhmc_model* h = new bhmc_model();
  h->impl = m;
  *out = h;
  return BHMC_OK;
}

int bhmc_mlp_create(bhmc_ctx* ctx, int64_t n_rows, int32_t n_in, int32_t n_mid, int32_t n_out, float alpha,
                    float dropout_ratio, uint64_t seed, int64_t chain_id0, bhmc_model** out) {
  BHMC_CHECK_ARG(ctx && out, "ctx/out is NULL");
  BHMC_CHECK_ARG(n_rows > 0 && n_in > 0 && n_mid > 0 && n_out > 0, "bad MLP shape");
  BHMC_CHECK_ARG(dropout_ratio >= 0.f && dropout_ratio < 1.f, "dropout ratio %f out of [0,1)", dropout_ratio);
  ModelBase* m = mlp_model_new(ctx, n_rows, n_in, n_mid, n_out, alpha, dropout_ratio, seed, chain_id0);
  if (!m) return BHMC_ERR_NOMEM;
  bhmc_model* h = new bhmc_model();
  h->impl = m;
  *out = h;
  return BHMC_OK;
}

int bhmc_mlp_bind_data(bhmc_model* m, const float* X, const int32_t* labels, int32_t is_host) {
  BHMC_CHECK_ARG(m && m->impl, "model is NULL");
  BHMC_CUDA_OK(cudaSetDevice(m->impl->ctx->device));
  return mlp_model_bind(m->impl, X, labels, is_host);
}

int bhmc_mlp_set_masks(bhmc_model* m, const uint8_t* masks_dev) {
  BHMC_CHECK_ARG(m && m->impl, "model is NULL");
  return mlp_model_set_masks(m->impl, masks_dev);
}

int bhmc_mlp_predict(bhmc_model* m, const float* q, int32_t C, int64_t ld, const float* X_dev, int64_t nrows, int32_t prec,
                     float* probs_dev, int32_t* labels_dev) {
  BHMC_CHECK_ARG(m && m->impl, "model is NULL");
  BHMC_CUDA_OK(cudaSetDevice(m->impl->ctx->device));
  return mlp_model_predict(m->impl, q, C, ld, X_dev, nrows, prec, probs_dev, labels_dev);
}

int bhmc_model_set_global_rows(bhmc_model* m, int64_t n_global_rows, float alpha_global) {
  SoftmaxModel* s;
  BHMC_TRY(softmax_of(m, &s));
  BHMC_CHECK_ARG(n_global_rows >= s->d.N, "global row count %lld is smaller than the local %lld", (long long)n_global_rows,
                 (long long)s->d.N);
  s->global_rows = n_global_rows;
  s->alpha_energy = alpha_global;
  return BHMC_OK;
}

int bhmc_model_destroy(bhmc_model* m) {
  if (!m) return BHMC_OK;
  if (m->impl) {
    cudaStreamSynchronize(m->impl->ctx->stream);
    delete m->impl;
  }
  delete m;
  return BHMC_OK;
}

int64_t bhmc_model_n_params(const bhmc_model* m) { return (m && m->impl) ? m->impl->P : -1; }
int32_t bhmc_model_n_vars(const bhmc_model* m) { return (m && m->impl) ? m->impl->n_vars : -1; }

int bhmc_model_var_layout(const bhmc_model* m, int64_t* offsets, int64_t* lengths) {
  BHMC_CHECK_ARG(m && m->impl && offsets && lengths, "NULL argument");
  for (int v = 0; v < m->impl->n_vars; ++v) offsets[v] = m->impl->var_off[v], lengths[v] = m->impl->var_len[v];
  return BHMC_OK;
}

static int check_state(const bhmc_model* m, const float* q, int C, int64_t ld) {
  BHMC_CHECK_ARG(m && m->impl && q, "NULL argument");
  BHMC_CHECK_ARG(C > 0 && C <= 65535, "n_chains %d out of range [1,65535]", C);
  BHMC_CHECK_ARG(ld >= m->impl->P && ld % 4 == 0, "ld %lld must be >= P=%lld and a multiple of 4", (long long)ld,
                 (long long)m->impl->P);
  BHMC_CHECK_ARG(((uintptr_t)q & 15) == 0, "chain state must be 16-byte aligned");
  return BHMC_OK;
}

int bhmc_model_grad(bhmc_model* m, const float* q, int32_t C, int64_t ld, int64_t row0, int64_t nrows, int32_t prec,
                    float* g, double* loglik) {
  BHMC_TRY(check_state(m, q, C, ld));
  BHMC_CHECK_ARG(g && loglik, "g/loglik is NULL");
  BHMC_CUDA_OK(cudaSetDevice(m->impl->ctx->device));
  return m->impl->grad(q, C, ld, row0, nrows, prec, g, loglik);
}

int bhmc_model_loglik(bhmc_model* m, const float* q, int32_t C, int64_t ld, int64_t row0, int64_t nrows, int32_t prec,
                      double* loglik) {
  BHMC_TRY(check_state(m, q, C, ld));
  BHMC_CHECK_ARG(loglik, "loglik is NULL");
  BHMC_CUDA_OK(cudaSetDevice(m->impl->ctx->device));
  return m->impl->grad(q, C, ld, row0, nrows, prec, nullptr, loglik);
}

int bhmc_model_nlp(bhmc_model* m, const float* q, int32_t C, int64_t ld, int64_t row0, int64_t nrows, int32_t prec,
                   double* nlp) {
  BHMC_TRY(check_state(m, q, C, ld));
  BHMC_CHECK_ARG(nlp, "nlp is NULL");
  ModelBase* mb = m->impl;
  BHMC_CUDA_OK(cudaSetDevice(mb->ctx->device));
  BHMC_TRY(mb->grad(q, C, ld, row0, nrows, prec, nullptr, nlp));
  double a, b, cv[BHMC_MAX_VARS];
  mb->energy_coeffs(nrows, &a, &b, cv);
  bool need_extra = false;
  for (int v = 0; v < mb->n_vars; ++v) need_extra |= cv[v] != 0.0;
  if (!need_extra) return launch_affine(mb->ctx, nlp, a, b, nullptr, nlp, C);
  void* ss = nullptr;
  BHMC_TRY(mb->ctx->get_scratch(4, sizeof(double) * C * (mb->n_vars + 1), &ss));
  double* sumsq = (double*)ss;
  double* extra = sumsq + (size_t)C * mb->n_vars;
  BHMC_TRY(launch_prior_energy(mb->ctx, q, ld, C, mb->n_vars, mb->var_off, mb->var_len, cv, sumsq, extra));
  return launch_affine(mb->ctx, nlp, a, b, extra, nlp, C);
}

int bhmc_softmax_predict(bhmc_model* m, const float* q, int32_t C, int64_t ld, const float* X_dev, int64_t nrows,
                         float* probs, int32_t* labels) {
  SoftmaxModel* s;
  BHMC_TRY(softmax_of(m, &s));
  BHMC_TRY(check_state(m, q, C, ld));
  BHMC_CHECK_ARG(X_dev && nrows > 0, "X is NULL / no rows");
  BHMC_CUDA_OK(cudaSetDevice(s->ctx->device));
  return simt_softmax_predict(s->ctx, s->d.D, s->d.K, q, C, ld, X_dev, nrows, probs, labels);
}

int bhmc_philox_normal(bhmc_ctx* ctx, float* out, int32_t C, int64_t P, int64_t ld, uint64_t seed, int64_t chain_id0,
                       uint32_t slo, uint32_t shi) {
  BHMC_CHECK_ARG(ctx && out && C > 0 && P > 0 && ld >= P, "bad argument");
  BHMC_CUDA_OK(cudaSetDevice(ctx->device));
  return launch_philox_normal(ctx, out, C, P, ld, seed, chain_id0, slo, shi);
}

double bhmc_philox_uniform_host(uint64_t seed, int64_t chain_id, uint32_t slo, uint32_t shi) {
  return philox_uniform(seed, chain_id, slo, shi);
}

void bhmc_philox4x32_host(const uint32_t counter[4], const uint32_t key[2], uint32_t out[4]) {
  U4 r = philox4x32_10(U4{counter[0], counter[1], counter[2], counter[3]}, key[0], key[1]);
  out[0] = r.x, out[1] = r.y, out[2] = r.z, out[3] = r.w;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// sampler
// ------------------------------------------------------------------------------------------------
struct bhmc_sampler {
  bhmc_ctx* ctx = nullptr;
  ModelBase* model = nullptr;
  bhmc_sampler_config cfg;
  int64_t P = 0, ld = 0;
  float* state = nullptr;  // q | p | g | q_new | p_new  (each [C, ld])
  float *q = nullptr, *p = nullptr, *g = nullptr, *q_new = nullptr, *p_new = nullptr;
  double* scal = nullptr;  // stat | stat_cur | stat_new | kin0 | kin1 | extra_cur | extra_new | sumsq[C*nv] | u[C]
  int32_t* Ldev = nullptr;
  bhmc_grad_hook hook = nullptr;
  void* hook_user = nullptr;
  bhmc_comm* row_comm = nullptr;  // rows sharded over ranks: all-reduce after every evaluation (comm.cu)
  // gradient (or log-lik only when g == nullptr) of the first `rows` working rows, then the optional hook
  // evaluation whose reduce launch also runs the streaming schedule's next update; BHMC_ERR_UNSUPPORTED = not fused
  int eval_fused_stream(const float* q, int rows, int64_t row0, int64_t nrows, float* g, double* stat, uint32_t hint,
                        const FusedStream& fst) {
    if (hook || row_comm) return BHMC_ERR_UNSUPPORTED;  // the row-shard all-reduce must see g before any update uses it
    ctx->cur_units = rows;
    return model->grad_fused_stream(q, rows, ld, row0, nrows, cfg.precision, g, stat, hint, fst);
  }
  int eval_fused_update(const float* q, int rows, int64_t row0, int64_t nrows, double* stat, const bhmc::UpdateArgs& u) {
    ctx->cur_units = rows;
    return model->grad_fused_update(q, rows, ld, row0, nrows, cfg.precision, stat, u);
  }
  int eval(const float* q, int rows, int64_t row0, int64_t nrows, float* g, double* stat, uint32_t hint = 0) {
    ctx->cur_units = rows;
    BHMC_TRY(model->grad(q, rows, ld, row0, nrows, cfg.precision, g, stat, hint));
    if (row_comm) BHMC_TRY(bhmc_comm_allreduce(row_comm, g, g ? (int64_t)rows * ld : 0, stat, rows));
    if (hook && hook(hook_user, g, stat, rows, ld) != 0) {
      bhmc::set_error("gradient hook reported a failure");
      return BHMC_ERR_STATE;
    }
    return BHMC_OK;
  }
};

extern "C" {

int bhmc_sampler_create(bhmc_ctx* ctx, bhmc_model* model, const bhmc_sampler_config* cfg, bhmc_sampler** out) {
  BHMC_CHECK_ARG(ctx && model && model->impl && cfg && out, "NULL argument");
  BHMC_CHECK_ARG(cfg->n_chains > 0 && cfg->n_chains <= 65535, "n_chains %d out of range [1,65535]", cfg->n_chains);
  BHMC_CHECK_ARG(cfg->kind >= BHMC_KIND_HMC && cfg->kind <= BHMC_KIND_SGD, "unknown sampler kind %d", cfg->kind);
  BHMC_CHECK_ARG(cfg->n_sweep >= 1 && cfg->n_sweep <= BHMC_MAX_VARS, "n_sweep %d out of range", cfg->n_sweep);
  ModelBase* mb = model->impl;
  for (int i = 0; i < cfg->n_sweep; ++i)
    BHMC_CHECK_ARG(cfg->sweep_off[i] >= 0 && cfg->sweep_len[i] > 0 && cfg->sweep_off[i] + cfg->sweep_len[i] <= mb->P,
                   "sweep group %d [%lld,+%lld) outside the %lld parameters", i, (long long)cfg->sweep_off[i],
                   (long long)cfg->sweep_len[i], (long long)mb->P);
  BHMC_CUDA_OK(cudaSetDevice(ctx->device));
  auto* s = new (std::nothrow) bhmc_sampler();
  if (!s) return BHMC_ERR_NOMEM;
  s->ctx = ctx;
  s->model = mb;
  s->cfg = *cfg;
  s->P = mb->P;
  s->ld = round_up(mb->P, 4);
  int C = cfg->n_chains;
  size_t row = (size_t)C * s->ld;
  cudaError_t e = cudaMalloc(&s->state, sizeof(float) * row * 5);
  if (e != cudaSuccess) {
    delete s;
    set_error("cudaMalloc of chain state (%zu bytes) failed: %s", sizeof(float) * row * 5, cudaGetErrorString(e));
    return BHMC_ERR_NOMEM;
  }
  BHMC_CUDA_OK(cudaMemsetAsync(s->state, 0, sizeof(float) * row * 5, ctx->stream));
  s->q = s->state;
  s->p = s->state + row;
  s->g = s->state + 2 * row;
  s->q_new = s->state + 3 * row;
  s->p_new = s->state + 4 * row;
  size_t nscal = (size_t)C * (8 + mb->n_vars);
  BHMC_CUDA_OK(cudaMalloc(&s->scal, sizeof(double) * nscal));
  BHMC_CUDA_OK(cudaMemsetAsync(s->scal, 0, sizeof(double) * nscal, ctx->stream));
  *out = s;
  return BHMC_OK;
}

int bhmc_sampler_destroy(bhmc_sampler* s) {
  if (!s) return BHMC_OK;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  cudaFree(s->state);
  cudaFree(s->scal);
  cudaFree(s->Ldev);
  delete s;
  return BHMC_OK;
}

int64_t bhmc_sampler_ld(const bhmc_sampler* s) { return s ? s->ld : -1; }

int bhmc_sampler_set_grad_hook(bhmc_sampler* s, bhmc_grad_hook hook, void* user) {
  BHMC_CHECK_ARG(s, "sampler is NULL");
  s->hook = hook;
  s->hook_user = user;
  return BHMC_OK;
}

int bhmc_sampler_set_row_comm(bhmc_sampler* s, bhmc_comm* comm) {
  BHMC_CHECK_ARG(s, "sampler is NULL");
  s->row_comm = comm;
  return BHMC_OK;
}

int bhmc_sampler_state_ptr(bhmc_sampler* s, int32_t which, float** out) {
  BHMC_CHECK_ARG(s && out && which >= 0 && which <= 2, "bad argument");
  *out = which == 0 ? s->q : which == 1 ? s->p : s->g;
  return BHMC_OK;
}

int bhmc_sampler_set_q(bhmc_sampler* s, const float* src, int32_t is_host) {
  BHMC_CHECK_ARG(s && src, "NULL argument");
  BHMC_CUDA_OK(cudaSetDevice(s->ctx->device));
  BHMC_CUDA_OK(cudaMemcpy2DAsync(s->q, sizeof(float) * s->ld, src, sizeof(float) * s->P, sizeof(float) * s->P,
                                 s->cfg.n_chains, is_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice,
                                 s->ctx->stream));
  return BHMC_OK;
}

int bhmc_sampler_get(bhmc_sampler* s, int32_t which, float* dst, int32_t is_host) {
  BHMC_CHECK_ARG(s && dst && which >= 0 && which <= 2, "bad argument");
  BHMC_CUDA_OK(cudaSetDevice(s->ctx->device));
  const float* src = which == 0 ? s->q : which == 1 ? s->p : s->g;
  BHMC_CUDA_OK(cudaMemcpy2DAsync(dst, sizeof(float) * s->P, src, sizeof(float) * s->ld, sizeof(float) * s->P,
                                 s->cfg.n_chains, is_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice,
                                 s->ctx->stream));
  if (is_host) BHMC_CUDA_OK(cudaStreamSynchronize(s->ctx->stream));
  return BHMC_OK;
}

// Host-only planner of the streaming schedule (no CUDA): everything the launch loop needs, from the path lengths.
struct StreamPlan {
  int64_t J = 0;                      // gradient launches; elementwise phases are 0..J
  int64_t n_grad_evals = 0;           // chain-evaluations that move a chain (+ one start point per chain)
  std::vector<int32_t> perm;          // row -> chain, most total work first
  std::vector<int> rows_el, rows_grad;  // active rows of elementwise phase j / gradient launch j (prefixes)
  std::vector<char> ev_finish, ev_begin;
};
static int64_t stream_slots(int Lt, int nsw) { return (int64_t)std::max(Lt - 1, 1) * nsw; }
// pass 1: sizes and row order
static void stream_plan_order(const int32_t* L, int C, int n_steps, int nsw, StreamPlan* pl) {
  std::vector<int64_t> T(C, 1);  // launches per chain: 1 (start point of its first transition) + max(L-1,1)*nsw each
  pl->n_grad_evals = 0;
  for (int c = 0; c < C; ++c) {
    pl->n_grad_evals += 1;
    for (int t = 0; t < n_steps; ++t) {
      T[c] += stream_slots(L[(size_t)t * C + c], nsw);
      pl->n_grad_evals += (int64_t)std::max(L[(size_t)t * C + c] - 1, 0) * nsw;
    }
  }
  pl->perm.resize(C);
  for (int c = 0; c < C; ++c) pl->perm[c] = c;
  std::stable_sort(pl->perm.begin(), pl->perm.end(), [&](int32_t x, int32_t y) { return T[x] > T[y]; });
  pl->J = T[pl->perm[0]];
}
// pass 2: op tables, [J+1][C] each, zero-initialised by the caller.  code1 = update before the Metropolis tests of a
// phase, code2 = update after the begins of the phase (first half kick + drift of the transitions that start in it)
static void stream_plan_ops(const int32_t* L, int C, int n_steps, int nsw, StreamPlan* pl, uint32_t* code1, uint32_t* code2,
                            int32_t* step1, int32_t* step2) {
  const int64_t J = pl->J;
  pl->rows_el.assign(J + 2, 0);
  pl->rows_grad.assign(J + 1, 0);
  pl->ev_finish.assign(J + 1, 0);
  pl->ev_begin.assign(J + 1, 0);
  for (int r = 0; r < C; ++r) {
    const int c = pl->perm[r];
    code1[0 * C + r] = OP_BEGIN;  // phase 0: begin transition 0; launch 0 evaluates its start point
    step1[0 * C + r] = 0;
    pl->ev_begin[0] = 1;
    int64_t j = 1;
    for (int t = 0; t < n_steps; ++t) {
      const int Lt = L[(size_t)t * C + c];
      const int iters = std::max(Lt - 1, 0);
      // phase j: first sub-step of transition t.  t == 0: in the ordinary update (after the start-point launch);
      // t > 0: in the post-begin update of the phase in which transition t-1 was finished
      uint32_t* first_code = t == 0 ? code1 : code2;
      int32_t* first_step = t == 0 ? step1 : step2;
      first_code[j * C + r] |= (t == 0 ? OP_LATCH : OP_LATCH_CACHED) | (iters > 0 ? OP_PRE | (0u << 12) : 0u);
      first_step[j * C + r] = t;
      const int64_t ns = stream_slots(Lt, nsw);
      for (int64_t k = 1; k < ns; ++k) {  // remaining sub-steps: closing kick of the previous + opening of this
        const int v = (int)(k % nsw), pv = v == 0 ? nsw - 1 : v - 1;
        if (iters > 0) code1[(j + k) * C + r] |= OP_POST | ((uint32_t)pv << 8) | OP_PRE | ((uint32_t)v << 12);
        step1[(j + k) * C + r] = t;
      }
      j += ns;
      // phase j: closing kick of the last variable, Metropolis test, begin of transition t+1
      if (iters > 0) code1[j * C + r] |= OP_POST | ((uint32_t)(nsw - 1) << 8);
      code1[j * C + r] |= OP_FINISH | (t + 1 < n_steps ? OP_BEGIN : 0u);
      step1[j * C + r] = t;
      pl->ev_finish[j] = 1;
      if (t + 1 < n_steps) pl->ev_begin[j] = 1;
    }
    // j == T[c]: the chain takes part in elementwise phases 0..T and gradient launches 0..T-1
    for (int64_t k = 0; k <= j; ++k) pl->rows_el[k] = r + 1;  // rows sorted by T descending -> prefix
    for (int64_t k = 0; k < j; ++k) pl->rows_grad[k] = r + 1;
  }
}

// Test hook (host only, no device needed): compiles the plan for path lengths L[n_steps][n_chains].  Call with
// code1 == NULL to get the sizes (n_phases = J), then with four [(J+1)*n_chains] arrays, perm[n_chains] and
// rows_el[J+1] / rows_grad[J+1].
extern "C" int bhmc_stream_plan_host(const int32_t* L, int32_t n_chains, int32_t n_steps, int32_t n_sweep, int64_t* n_phases,
                                     int64_t* n_grad_evals, uint32_t* code1, uint32_t* code2, int32_t* step1, int32_t* step2,
                                     int32_t* perm, int32_t* rows_el, int32_t* rows_grad) {
  BHMC_CHECK_ARG(L && n_chains > 0 && n_steps > 0 && n_sweep > 0 && n_sweep <= BHMC_MAX_VARS && n_phases, "bad argument");
  StreamPlan pl;
  stream_plan_order(L, n_chains, n_steps, n_sweep, &pl);
  *n_phases = pl.J;
  if (n_grad_evals) *n_grad_evals = pl.n_grad_evals;
  if (!code1) return BHMC_OK;
  BHMC_CHECK_ARG(code2 && step1 && step2 && perm && rows_el && rows_grad, "output array is NULL");
  const size_t n = (size_t)(pl.J + 1) * n_chains;
  memset(code1, 0, sizeof(uint32_t) * n);
  memset(code2, 0, sizeof(uint32_t) * n);
  memset(step1, 0, sizeof(int32_t) * n);
  memset(step2, 0, sizeof(int32_t) * n);
  stream_plan_ops(L, n_chains, n_steps, n_sweep, &pl, code1, code2, step1, step2);
  for (int c = 0; c < n_chains; ++c) perm[c] = pl.perm[c];
  for (int64_t j = 0; j <= pl.J; ++j) rows_el[j] = pl.rows_el[j], rows_grad[j] = j < pl.J ? pl.rows_grad[j] : 0;
  return BHMC_OK;
}

// ---- HMC, streaming schedule (asynchronous chains) -----------------------------------------------------------
// Path lengths are fresh uniforms per chain and step (hmc.py:46), so in lockstep the chains of one transition finish
// at very different times (E[L]/max L ~ 0.5 for 64 chains).  Chains are independent, so nothing forces them to start
// transition t+1 together: here every chain walks through its own sequence of transitions back to back, and every
// gradient launch carries all chains that still have work.  The host knows every L in advance (host Philox == device
// Philox), so it compiles the whole call into op codes per (phase, row); rows are sorted by total work so that the
// rows still busy at the end of the call form a prefix.  Draws are keyed by (chain, step): the decisions are the ones
// the lockstep schedule takes.
//
// The gradient at the start point of transition t+1 is not re-evaluated: after an accepted proposal it is the last
// evaluation of transition t, after a rejection it is the start-point gradient of transition t (kept in g_start).
// Every transition therefore occupies (L-1)*nsw launches -- a multiple of the sweep length -- so ALL rows move the
// same variable before a given launch: launch j follows a move of sweep group (j-1) mod nsw.  That uniformity is
// what lets the model reuse X.W after bias-only moves (GRAD_HINT_CHEAP_MOVE) for whole launches.  A transition with
// L <= 1 (no leapfrog iteration, A = 1) idles for one sweep so that the alternation holds.
static int hmc_run_streaming(bhmc_sampler* s, bhmc_hmc_run* run, const std::vector<int32_t>& L /* [n_steps][C] */) {
  const bhmc_sampler_config& cfg = s->cfg;
  bhmc_ctx* ctx = s->ctx;
  ModelBase* mb = s->model;
  const int C = cfg.n_chains, nsw = cfg.n_sweep, n_steps = run->n_steps;
  const int64_t P = s->P, ld = s->ld;
  const int64_t nrows = run->nrows > 0 ? run->nrows : mb->default_rows();
  const float eps = (float)run->step_size;
  double* stat = s->scal;
  double* stat_cur = stat + C;
  double* stat_new = stat + 2 * C;
  double* kin1 = stat + 4 * C;
  double* extra_cur = stat + 5 * C;
  double* extra_new = stat + 6 * C;
  double ea, eb, cv[BHMC_MAX_VARS];
  mb->energy_coeffs(nrows, &ea, &eb, cv);
  bool need_extra = false;
  for (int v = 0; v < mb->n_vars; ++v) need_extra |= cv[v] != 0.0;
  struct timespec ts_setup;
  clock_gettime(CLOCK_MONOTONIC, &ts_setup);

  StreamPlan pl;
  stream_plan_order(L.data(), C, n_steps, nsw, &pl);
  run->n_grad_evals = pl.n_grad_evals;
  const int64_t J = pl.J;
  const size_t n_ops = (size_t)(J + 1) * C;
  const size_t code_bytes = round_up(sizeof(uint32_t) * n_ops, 64), perm_bytes = round_up(sizeof(int32_t) * C, 64);
  const size_t u_bytes = sizeof(double) * (size_t)n_steps * C;
  if (ctx->pinned_inflight) {
    BHMC_CUDA_OK(cudaEventSynchronize(ctx->pinned_ev));
    ctx->pinned_inflight = false;
  }
  void* pin = nullptr;
  const size_t tab_bytes = 4 * code_bytes + perm_bytes;
  BHMC_TRY(ctx->get_pinned(tab_bytes + u_bytes, &pin));
  uint32_t* code1_h = (uint32_t*)pin;
  uint32_t* code2_h = (uint32_t*)((char*)pin + code_bytes);
  int32_t* step1_h = (int32_t*)((char*)pin + 2 * code_bytes);
  int32_t* step2_h = (int32_t*)((char*)pin + 3 * code_bytes);
  int32_t* perm_h = (int32_t*)((char*)pin + 4 * code_bytes);
  double* uacc_h = (double*)((char*)pin + tab_bytes);
  memset(pin, 0, 4 * code_bytes);
  stream_plan_ops(L.data(), C, n_steps, nsw, &pl, code1_h, code2_h, step1_h, step2_h);
  for (int r = 0; r < C; ++r) perm_h[r] = pl.perm[r];
  const std::vector<int>& rows_el = pl.rows_el;
  const std::vector<int>& rows_grad = pl.rows_grad;
  const std::vector<char>& ev_finish = pl.ev_finish;
  const std::vector<char>& ev_begin = pl.ev_begin;
  if (run->u_accept_host) memcpy(uacc_h, run->u_accept_host, u_bytes);
  void* dev = nullptr;
  BHMC_TRY(ctx->get_scratch(6, std::max(tab_bytes + u_bytes, (size_t)32 << 20), &dev));
  const uint32_t* code1_d = (const uint32_t*)dev;
  const uint32_t* code2_d = (const uint32_t*)((char*)dev + code_bytes);
  const int32_t* step1_d = (const int32_t*)((char*)dev + 2 * code_bytes);
  const int32_t* step2_d = (const int32_t*)((char*)dev + 3 * code_bytes);
  const int32_t* perm_d = (const int32_t*)((char*)dev + 4 * code_bytes);
  const double* uacc_d = (const double*)((char*)dev + tab_bytes);
  BHMC_CUDA_OK(cudaMemcpyAsync(dev, pin, tab_bytes + (run->u_accept_host ? u_bytes : 0), cudaMemcpyHostToDevice, ctx->stream));
  if (!ctx->pinned_ev) BHMC_CUDA_OK(cudaEventCreateWithFlags(&ctx->pinned_ev, cudaEventDisableTiming));
  BHMC_CUDA_OK(cudaEventRecord(ctx->pinned_ev, ctx->stream));
  ctx->pinned_inflight = true;
  // per-row scalars of this schedule: kin0 parity buffers [2C], prior-energy scratch [C*(n_vars+1)], stat_next [C],
  // extra_next [C], acc_flag [C]; and the start-of-step gradients g_start [C, ld]
  void* kbuf = nullptr;
  const size_t n_dbl = (size_t)2 * C + (size_t)C * (mb->n_vars + 1) + 2 * (size_t)C;
  BHMC_TRY(ctx->get_scratch(4, sizeof(double) * n_dbl + sizeof(int32_t) * C, &kbuf));
  double* kin0 = (double*)kbuf;
  double* extra_tmp = kin0 + 2 * (size_t)C;
  double* stat_next = extra_tmp + (size_t)C * (mb->n_vars + 1);
  double* extra_next = stat_next + C;
  int32_t* acc_flag = (int32_t*)(extra_next + C);
  BHMC_CUDA_OK(cudaMemsetAsync(kin0, 0, sizeof(double) * 2 * C, ctx->stream));
  void* gbuf = nullptr;
  BHMC_TRY(ctx->get_scratch(9, sizeof(float) * (size_t)C * ld, &gbuf));
  float* g_start = (float*)gbuf;

  StreamUpdateArgs u{};
  u.q = s->q_new;
  u.p = s->p_new;
  u.g = s->g;
  u.ld = ld;
  u.P = P;
  u.n_vars = nsw;
  for (int v = 0; v < nsw; ++v) u.off[v] = cfg.sweep_off[v], u.len[v] = cfg.sweep_len[v];
  u.a_pre = 0.5f * eps;                         // hmc.py:51
  u.a_post = cfg.leapfrog ? 0.5f * eps : eps;   // hmc.py:54 applies a FULL eps kick
  u.eps = eps;
  u.stat = stat;
  u.stat_cur = stat_cur;
  u.stat_new = stat_new;
  u.kin0 = kin0;
  u.kin1 = kin1;
  u.C_total = C;
  u.g_start = g_start;
  u.stat_next = stat_next;
  u.extra_next = need_extra ? extra_next : nullptr;
  u.extra_cur = extra_cur;

  bool cheap[BHMC_MAX_VARS];
  for (int v = 0; v < nsw; ++v) cheap[v] = mb->cheap_slice(cfg.sweep_off[v], cfg.sweep_len[v]);
  bool kept = false;  // the previous launch kept what a cheap move can reuse, for every active row
  static int fuse_env = -1;  // BHMC_FUSED_STREAM=0: separate reduce / update / prep launches (A/B measurements)
  if (fuse_env < 0) {
    const char* e = getenv("BHMC_FUSED_STREAM");
    fuse_env = e ? atoi(e) : 1;
  }
  bool try_fuse = fuse_env != 0;
  bool update_done = false;  // this phase's pre-event update already ran inside the previous launch's reduce
  bool prepared = false;     // ... and so did the operand preparation of this phase's launch

  run->n_grad_launched = 0;
  static int prof_host = -1;
  if (prof_host < 0) prof_host = getenv("BHMC_PROF_HOST") ? 1 : 0;
  struct timespec ts0, ts1;
  if (prof_host) {
    clock_gettime(CLOCK_MONOTONIC, &ts0);
    fprintf(stderr, "[bhmc prof host] streaming run: setup (op table, staging buffers, upload) %.2f ms\n",
            1e3 * ((ts0.tv_sec - ts_setup.tv_sec) + 1e-9 * (ts0.tv_nsec - ts_setup.tv_nsec)));
  }
  for (int64_t j = 0; j <= J; ++j) {
    const int rows = rows_el[j];
    const uint32_t* cj = code1_d + j * C;
    const int32_t* sj = step1_d + j * C;
    if (j > 0 && !update_done) {
      u.rows = rows;
      u.code = cj;
      u.step = sj;
      BHMC_TRY(launch_stream_update(ctx, u));
    }
    update_done = false;
    if (ev_finish[j]) {
      BHMC_TRY(launch_stream_kinetic(ctx, s->p_new, ld, P, rows, cj, kin1));
      if (need_extra)
        BHMC_TRY(launch_prior_energy(ctx, s->q_new, ld, rows, mb->n_vars, mb->var_off, mb->var_len, cv, extra_tmp + C,
                                     extra_new, cj, OP_FINISH));
      AcceptArgs a{};
      a.q = s->q;
      a.q_new = s->q_new;
      a.p_out = s->p;
      a.p_new = s->p_new;
      a.ld = ld;
      a.P = P;
      a.C = rows;
      a.p_sign = -1.f;  // hmc.py:58-59
      a.stat_cur = stat_cur;
      a.stat_new = stat_new;
      a.ea = ea;
      a.eb = eb;
      a.extra_cur = need_extra ? extra_cur : nullptr;
      a.extra_new = need_extra ? extra_new : nullptr;
      a.kin0 = kin0;
      a.kin1 = kin1;
      a.u = run->u_accept_host ? uacc_d : nullptr;
      a.seed = cfg.seed;
      a.chain_id0 = cfg.chain_id0;
      a.stream_lo = (uint32_t)run->step0;
      a.stream_hi = TAG_ACCEPT;
      a.reject_nan = cfg.reject_nan;
      a.sample = run->samples_dev;
      a.loss = run->loss_dev;
      a.accept_prob = run->accept_prob_dev;
      a.accepted = run->accepted_dev;
      a.perm = perm_d;
      a.code = cj;
      a.step = sj;
      a.C_total = C;
      a.stat_next = stat_next;
      a.extra_next = need_extra ? extra_next : nullptr;
      a.acc_flag = acc_flag;
      BHMC_TRY(launch_accept(ctx, a));
    }
    if (j == J) break;
    if (ev_begin[j]) {
      BeginArgs b{};
      b.q = s->q;
      b.q_new = s->q_new;
      b.p0 = s->p;
      b.p_new = s->p_new;
      b.ld = ld;
      b.P = P;
      b.C = rows_grad[j];
      b.z = run->z_momentum_dev;
      b.ld_z = P;
      b.z_step_stride = (int64_t)C * P;
      b.seed = cfg.seed;
      b.chain_id0 = cfg.chain_id0;
      b.stream_lo = (uint32_t)run->step0;
      b.stream_hi = TAG_MOMENTUM;
      b.kin0 = kin0;
      b.perm = perm_d;
      b.code = cj;
      b.step = sj;
      b.C_total = C;
      b.g = s->g;
      b.g_start = g_start;
      b.acc_flag = acc_flag;
      BHMC_TRY(launch_hmc_begin(ctx, b));
      if (j == 0) {
        if (need_extra)  // quadratic log-prior part of U(q) at the very first start point
          BHMC_TRY(launch_prior_energy(ctx, s->q_new, ld, rows_grad[j], mb->n_vars, mb->var_off, mb->var_len, cv, extra_tmp + C,
                                       extra_cur, cj, OP_BEGIN));
      } else {  // first half kick + drift of the transitions that just began
        u.rows = rows_grad[j];
        u.code = code2_d + j * C;
        u.step = step2_d + j * C;
        BHMC_TRY(launch_stream_update(ctx, u));
      }
    }
    // what moved since the previous launch: nothing before launch 0 (start points), else sweep group (j-1) % nsw for
    // every row; a begin in this phase re-seats q_new, which a cheap move alone would not tell the model
    uint32_t hint = 0;
    if (j > 0) {
      const int v = (int)((j - 1) % nsw), vn = (int)(j % nsw);
      if (kept && cheap[v] && !ev_begin[j]) hint |= GRAD_HINT_CHEAP_MOVE;
      const bool keep_next = cheap[vn] && j + 1 < J;  // a cheap evaluation leaves the kept data valid, a full one renews it
      if (keep_next) hint |= GRAD_HINT_KEEP;
      kept = keep_next;
    } else {
      kept = cheap[0];
      if (kept) hint |= GRAD_HINT_KEEP;
    }
    if (prepared && !ev_begin[j]) hint |= GRAD_HINT_PREPARED;
    prepared = false;
    bool fused = false;
    if (try_fuse) {
      // the next phase's pre-event update runs on exactly the rows of this launch (rows_el[j+1] == rows_grad[j])
      FusedStream fst{};
      fst.u = u;
      fst.u.rows = rows_grad[j];
      fst.u.code = code1_d + (j + 1) * C;
      fst.u.step = step1_d + (j + 1) * C;
      // operand preparation of launch j+1: only if it is a full forward pass (not a cheap move served by the cache)
      // and no begin re-seats rows in between; the lo operand copy sits behind the hi copy of ALL rows of a launch,
      // so the two launches must carry the same number of rows
      const int vn = (int)(j % nsw);
      const bool next_cheap = j + 1 < J && kept && cheap[vn] && !ev_begin[j + 1];
      fst.prep_next = j + 1 < J && !next_cheap && !ev_begin[j + 1] && rows_grad[j + 1] == rows_grad[j];
      const int rc = s->eval_fused_stream(s->q_new, rows_grad[j], run->row0, nrows, s->g, stat, hint, fst);
      if (rc == BHMC_OK) {
        fused = true;
        update_done = true;
        prepared = fst.prep_next;
      } else if (rc != BHMC_ERR_UNSUPPORTED) {
        return rc;
      } else {
        try_fuse = false;
      }
    }
    if (!fused) BHMC_TRY(s->eval(s->q_new, rows_grad[j], run->row0, nrows, s->g, stat, hint));
    run->n_grad_launched += rows_grad[j];
  }
  run->n_phases = (int32_t)J;
  if (prof_host) {
    clock_gettime(CLOCK_MONOTONIC, &ts1);
    const double host_s = (ts1.tv_sec - ts0.tv_sec) + 1e-9 * (ts1.tv_nsec - ts0.tv_nsec);
    fprintf(stderr, "[bhmc prof host] streaming run: %lld phases enqueued in %.1f ms of host time (%.1f us per phase, "
                    "asynchronous: the GPU may still be running)\n", (long long)J, 1e3 * host_s, 1e6 * host_s / (double)J);
  }
  return BHMC_OK;
}

// ---- HMC / SGHMC ---------------------------------------------------------------------------------
int bhmc_sampler_hmc_run(bhmc_sampler* s, bhmc_hmc_run* run) {
  NvtxRange nvtx_run("bhmc_sampler_hmc_run");
  BHMC_CHECK_ARG(s && run, "NULL argument");
  const bhmc_sampler_config& cfg = s->cfg;
  BHMC_CHECK_ARG(cfg.kind == BHMC_KIND_HMC || cfg.kind == BHMC_KIND_SGHMC, "sampler kind %d cannot run hmc_run", cfg.kind);
  BHMC_CHECK_ARG(run->n_steps >= 0 && run->step_size > 0 && run->path_length >= 0, "bad run parameters");
  bhmc_ctx* ctx = s->ctx;
  ModelBase* mb = s->model;
  BHMC_CUDA_OK(cudaSetDevice(ctx->device));
  const int C = cfg.n_chains, nsw = cfg.n_sweep;
  const bool sghmc = cfg.kind == BHMC_KIND_SGHMC;
  const int64_t P = s->P, ld = s->ld;
  const int64_t nrows = run->nrows > 0 ? run->nrows : mb->default_rows();
  const float eps = (float)run->step_size;
  double* stat = s->scal;
  double* stat_cur = stat + C;
  double* stat_new = stat + 2 * C;
  double* kin0 = stat + 3 * C;
  double* kin1 = stat + 4 * C;
  double* extra_cur = stat + 5 * C;
  double* extra_new = stat + 6 * C;
  double* u_dev = stat + 7 * C;
  double* sumsq = stat + 8 * C;
  double ea, eb, cv[BHMC_MAX_VARS];
  mb->energy_coeffs(nrows, &ea, &eb, cv);
  bool need_extra = false;
  for (int v = 0; v < mb->n_vars; ++v) need_extra |= cv[v] != 0.0;

  // path lengths for every step and chain: L = ceil(2 u path/eps), hmc.py:46 (host side: it sizes the launch loop)
  const int n_steps = run->n_steps;
  if (n_steps == 0) return BHMC_OK;
  if (ctx->pinned_inflight) {
    BHMC_CUDA_OK(cudaEventSynchronize(ctx->pinned_ev));
    ctx->pinned_inflight = false;
  }
  // pinned staging: [L sorted | perm | u_accept] per step
  void* pin = nullptr;
  const size_t l_bytes = ((sizeof(int32_t) * (size_t)n_steps * C + 63) / 64) * 64;
  const size_t u_bytes = sizeof(double) * (size_t)n_steps * C;
  BHMC_TRY(ctx->get_pinned(2 * l_bytes + u_bytes, &pin));
  int32_t* Lh = (int32_t*)pin;
  int32_t* perm_h = (int32_t*)((char*)pin + l_bytes);
  double* uacc_h = (double*)((char*)pin + 2 * l_bytes);
  std::vector<int> lmax(n_steps, 0);
  std::vector<int> Ltmp(C);
  std::vector<int32_t> Lraw((size_t)n_steps * C);
  for (int t = 0; t < n_steps; ++t) {
    uint32_t step = (uint32_t)(run->step0 + t);
    for (int c = 0; c < C; ++c) {
      double u;
      if (run->u_path_host)
        u = run->u_path_host[(size_t)t * C + c];
      else
        u = philox_uniform(cfg.seed, cfg.shared_path ? -1 : cfg.chain_id0 + c, step, TAG_PATH);
      double Ld = std::ceil(2.0 * u * run->path_length / run->step_size);
      Lraw[(size_t)t * C + c] = Ld > 1e9 ? 1000000000 : (int)Ld;
    }
  }
  {
    static int sched_env = -1;  // BHMC_SCHEDULE=1|2 overrides AUTO (A/B measurements)
    if (sched_env < 0) {
      const char* e = getenv("BHMC_SCHEDULE");
      sched_env = e ? atoi(e) : 0;
    }
    int sched = run->schedule ? run->schedule : sched_env;
    const bool can_stream = !sghmc && C > 1;
    if (sched == BHMC_SCHED_AUTO) sched = (can_stream && !cfg.shared_path && n_steps >= 2) ? BHMC_SCHED_STREAMING : BHMC_SCHED_LOCKSTEP;
    BHMC_CHECK_ARG(sched != BHMC_SCHED_STREAMING || !sghmc, "the streaming schedule is implemented for HMC only");
    if (sched == BHMC_SCHED_STREAMING) {
      // The op tables hold 16 bytes per (phase, row) in pinned memory and again on the device; a long call at a small
      // step size (many leapfrog iterations per transition) must not turn into multi-GB allocations or an error:
      // over the budget the call is cut into consecutive streaming sub-calls (draws are keyed by the absolute step
      // index, so the samples are the same), and a single transition that alone exceeds it runs in lockstep, which
      // needs no table.  BHMC_STREAM_OPS_MAX overrides the budget (tests).
      static int64_t ops_max = 0;
      if (!ops_max) {
        const char* e = getenv("BHMC_STREAM_OPS_MAX");
        ops_max = e ? std::max<int64_t>(1, atoll(e)) : ((int64_t)4 << 20);
      }
      StreamPlan pl;
      stream_plan_order(Lraw.data(), C, n_steps, nsw, &pl);
      if ((pl.J + 1) * (int64_t)C <= ops_max) return hmc_run_streaming(s, run, Lraw);
      std::vector<int64_t> T(C);
      int64_t tot_evals = 0, tot_launched = 0, tot_phases = 0;
      int t0 = 0;
      while (t0 < n_steps) {
        std::fill(T.begin(), T.end(), 1);
        int m = 0;
        for (int t = t0; t < n_steps; ++t, ++m) {
          int64_t mx = 0;
          for (int c = 0; c < C; ++c) mx = std::max(mx, T[c] + stream_slots(Lraw[(size_t)t * C + c], nsw));
          if ((mx + 1) * (int64_t)C > ops_max) break;
          for (int c = 0; c < C; ++c) T[c] += stream_slots(Lraw[(size_t)t * C + c], nsw);
        }
        bhmc_hmc_run sub = *run;
        sub.n_steps = std::max(m, 1);
        sub.schedule = m >= 2 ? BHMC_SCHED_STREAMING : BHMC_SCHED_LOCKSTEP;
        sub.step0 = run->step0 + t0;
        const size_t oc = (size_t)t0 * C, op = oc * (size_t)P;
        if (run->z_momentum_dev) sub.z_momentum_dev = run->z_momentum_dev + op;
        if (run->u_path_host) sub.u_path_host = run->u_path_host + oc;
        if (run->u_accept_host) sub.u_accept_host = run->u_accept_host + oc;
        if (run->samples_dev) sub.samples_dev = run->samples_dev + op;
        if (run->loss_dev) sub.loss_dev = run->loss_dev + oc;
        if (run->accept_prob_dev) sub.accept_prob_dev = run->accept_prob_dev + oc;
        if (run->accepted_dev) sub.accepted_dev = run->accepted_dev + oc;
        BHMC_TRY(bhmc_sampler_hmc_run(s, &sub));
        tot_evals += sub.n_grad_evals;
        tot_launched += sub.n_grad_launched;
        tot_phases += sub.n_phases;
        t0 += sub.n_steps;
      }
      run->n_grad_evals = tot_evals;
      run->n_grad_launched = tot_launched;
      run->n_phases = (int32_t)std::min<int64_t>(tot_phases, INT32_MAX);
      return BHMC_OK;
    }
  }
  bool ragged = false;
  run->n_grad_evals = 0;
  run->n_grad_launched = 0;
  run->n_phases = 0;
  for (int t = 0; t < n_steps; ++t) {
    int32_t* Lt = Lh + (size_t)t * C;
    int32_t* pt = perm_h + (size_t)t * C;
    for (int c = 0; c < C; ++c) {
      Ltmp[c] = Lraw[(size_t)t * C + c];
      pt[c] = c;
      run->n_grad_evals += 1 + (int64_t)std::max(Ltmp[c] - 1, 0) * nsw;
      if (run->u_accept_host) uacc_h[(size_t)t * C + c] = run->u_accept_host[(size_t)t * C + c];
    }
    // ragged trajectories: sort chains by path length (longest first) so that the chains still moving at
    // leapfrog iteration `it` are always the first n_act(it) working rows -> masked chains cost nothing
    std::stable_sort(pt, pt + C, [&](int32_t x, int32_t y) { return Ltmp[x] > Ltmp[y]; });
    for (int r = 0; r < C; ++r) {
      Lt[r] = Ltmp[pt[r]];
      ragged |= pt[r] != r;
    }
    lmax[t] = Lt[0];
    BHMC_CHECK_ARG(!(sghmc && run->z_noise_dev) || std::max(lmax[t] - 1, 0) <= run->z_noise_iters,
                   "step %d needs %d noise iterations but the tape holds %lld", t, lmax[t] - 1, (long long)run->z_noise_iters);
  }
  void* ldev = nullptr;
  BHMC_TRY(ctx->get_scratch(6, 2 * l_bytes + u_bytes, &ldev));
  int32_t* Ld_all = (int32_t*)ldev;
  int32_t* perm_all = (int32_t*)((char*)ldev + l_bytes);
  double* uacc_d = (double*)((char*)ldev + 2 * l_bytes);
  BHMC_CUDA_OK(cudaMemcpyAsync(ldev, pin, 2 * l_bytes + (run->u_accept_host ? u_bytes : 0), cudaMemcpyHostToDevice,
                               ctx->stream));
  if (!ctx->pinned_ev) BHMC_CUDA_OK(cudaEventCreateWithFlags(&ctx->pinned_ev, cudaEventDisableTiming));
  BHMC_CUDA_OK(cudaEventRecord(ctx->pinned_ev, ctx->stream));
  ctx->pinned_inflight = true;
  (void)u_dev;

  for (int t = 0; t < n_steps; ++t) {
    const uint32_t step = (uint32_t)(run->step0 + t);
    const int32_t* Lc = Ld_all + (size_t)t * C;                         // sorted, row-indexed
    const int32_t* Lhost = Lh + (size_t)t * C;
    const int32_t* perm = ragged ? perm_all + (size_t)t * C : nullptr;  // row -> chain
    // rows still moving at leapfrog iteration `it` (L sorted descending): it < L-1
    auto n_act = [&](int it) {
      int n = 0;
      while (n < C && it < Lhost[n] - 1) ++n;
      return n;
    };
    // 1. momentum ~ N(0,1), proposal := current (hmc.py:40-44)
    BeginArgs b{};
    b.q = s->q;
    b.q_new = s->q_new;
    b.p0 = s->p;
    b.p_new = s->p_new;
    b.ld = ld;
    b.P = P;
    b.C = C;
    b.z = run->z_momentum_dev ? run->z_momentum_dev + (size_t)t * C * P : nullptr;
    b.ld_z = P;
    b.seed = cfg.seed;
    b.chain_id0 = cfg.chain_id0;
    b.stream_lo = step;
    b.stream_hi = TAG_MOMENTUM;
    b.kin0 = kin0;
    b.perm = perm;
    BHMC_TRY(launch_hmc_begin(ctx, b));
    if (need_extra)  // quadratic log-prior part of U(q) (row-indexed like the other energies)
      BHMC_TRY(launch_prior_energy(ctx, s->q_new, ld, C, mb->n_vars, mb->var_off, mb->var_len, cv, sumsq, extra_cur));
    // 2. gradient at the current point (hmc.py:47); its log-likelihood doubles as NLP(q)
    // evaluation hints: a sub-step that moved only a cheap slice (softmax: the bias) lets the model reuse X.W of
    // the previous evaluation, provided that one was asked to keep it
    bool cheap[BHMC_MAX_VARS];
    for (int v = 0; v < nsw; ++v) cheap[v] = mb->cheap_slice(cfg.sweep_off[v], cfg.sweep_len[v]);
    const int iters_pre = std::max(lmax[t] - 1, 0);
    bool kept = iters_pre > 0 && cheap[0];
    BHMC_TRY(s->eval(s->q_new, C, run->row0, nrows, s->g, stat, kept ? GRAD_HINT_KEEP : 0u));
    run->n_grad_launched += C;
    BHMC_CUDA_OK(cudaMemcpyAsync(stat_cur, stat, sizeof(double) * C, cudaMemcpyDeviceToDevice, ctx->stream));
    BHMC_CUDA_OK(cudaMemcpyAsync(stat_new, stat, sizeof(double) * C, cudaMemcpyDeviceToDevice, ctx->stream));
    // 3. leapfrog: (L-1) Gauss-Seidel sweeps (hmc.py:49-54 / sghmc.py:28-34)
    const int iters = std::max(lmax[t] - 1, 0);
    UpdateArgs u{};
    u.q = s->q_new;
    u.p = s->p_new;
    u.g = s->g;
    u.ld = ld;
    u.P = P;
    u.L = Lc;
    u.eps = eps;
    u.seed = cfg.seed;
    u.chain_id0 = cfg.chain_id0;
    u.stat = stat;
    u.stat_new = stat_new;
    u.perm = perm;
    if (sghmc) {
      u.f_post = 1.0f - eps;
      u.a_post = cfg.sghmc_descent ? eps : -eps;  // literal: p = (1-eps)p + eps*grad + r (sghmc.py:34)
      u.n_post = 2.0f * eps;                      // N(0, std = 2 eps), sghmc.py:31
      u.a_pre = 0.f;
    } else {
      u.f_post = 1.0f;
      u.a_post = cfg.leapfrog ? 0.5f * eps : eps;  // hmc.py:54 applies a FULL eps kick
      u.n_post = 0.f;
      u.a_pre = 0.5f * eps;                        // hmc.py:51
    }
    int rows_prev = 0;  // rows that took part in the previous sub-step (superset of the current ones)
    // joint sweep: a model may apply the update inside its gradient kernels (ModelBase::grad_fused_update); then the
    // update launch in front of the next evaluation has already happened
    const bool whole_sweep = nsw == 1 && cfg.sweep_off[0] == 0 && cfg.sweep_len[0] >= P && !s->row_comm && !s->hook;
    bool upd_applied = false;
    for (int it = 0; it < iters; ++it) {
      const int rows = n_act(it);
      for (int v = 0; v < nsw; ++v) {
        bool first = (it == 0 && v == 0);
        int pv = v == 0 ? nsw - 1 : v - 1, pit = v == 0 ? it - 1 : it;
        u.post_off = first ? 0 : cfg.sweep_off[pv];
        u.post_len = first ? 0 : cfg.sweep_len[pv];
        u.it_post = pit;
        u.pre_off = cfg.sweep_off[v];
        u.pre_len = cfg.sweep_len[v];
        u.it_pre = it;
        u.C = first ? rows : rows_prev;
        if (sghmc && !first) {
          u.z = run->z_noise_dev ? run->z_noise_dev + ((size_t)t * run->z_noise_iters + pit) * C * P : nullptr;
          u.ld_z = P;
          u.stream_lo = step;
          u.stream_hi = TAG_NOISE | (uint32_t)((pit * nsw + pv) & 0xffffff);
        }
        // joint sweep: the launch rewrites every parameter of the rows that move -> it can fill the model's operand mirror
        u.mir_hi = nullptr, u.mir_lo = nullptr, u.mir_ld = 0;
        if (!upd_applied) {  // (else: the previous evaluation applied the closing kick of it - 1 and the drift of it)
          const bool mirrored = whole_sweep && mb->operand_mirror(C, ld, cfg.precision, &u.mir_hi, &u.mir_lo, &u.mir_ld);
          BHMC_TRY(launch_hmc_update(ctx, u));
          if (mirrored) mb->mirror_written(s->q_new);
          u.mir_hi = nullptr, u.mir_lo = nullptr;
        }
        upd_applied = false;
        const bool last = it == iters - 1 && v == nsw - 1;
        const int vn = v == nsw - 1 ? 0 : v + 1;
        uint32_t hint = (kept && cheap[v]) ? GRAD_HINT_CHEAP_MOVE : 0u;
        const bool keep_next = !last && cheap[vn] && (cheap[v] ? kept : true);
        if (keep_next) hint |= GRAD_HINT_KEEP;
        kept = keep_next;
        if (whole_sweep) {  // evaluation + closing kick of `it` + drift of `it + 1` inside the model's gradient kernels?
          bhmc::UpdateArgs f = u;
          f.post_off = cfg.sweep_off[0], f.post_len = cfg.sweep_len[0], f.it_post = it;
          f.pre_off = cfg.sweep_off[0], f.pre_len = it + 1 < iters ? cfg.sweep_len[0] : 0, f.it_pre = it + 1;
          f.C = rows;
          if (sghmc) {
            f.z = run->z_noise_dev ? run->z_noise_dev + ((size_t)t * run->z_noise_iters + it) * C * P : nullptr;
            f.ld_z = P;
            f.stream_lo = step;
            f.stream_hi = TAG_NOISE | (uint32_t)((it * nsw) & 0xffffff);
          }
          if (mb->operand_mirror(C, ld, cfg.precision, &f.mir_hi, &f.mir_lo, &f.mir_ld)) {
            const int rc = s->eval_fused_update(s->q_new, rows, run->row0, nrows, stat, f);
            if (rc == BHMC_OK) upd_applied = true, kept = false;
            else if (rc != BHMC_ERR_UNSUPPORTED) return rc;
          }
        }
        if (!upd_applied) BHMC_TRY(s->eval(s->q_new, rows, run->row0, nrows, s->g, stat, hint));
        run->n_grad_launched += rows;
        rows_prev = rows;
      }
    }
    if (iters > 0 && !upd_applied) {  // closing kick of the last variable (unless the last evaluation applied it)
      u.post_off = cfg.sweep_off[nsw - 1];
      u.post_len = cfg.sweep_len[nsw - 1];
      u.it_post = iters - 1;
      u.pre_len = 0;
      u.pre_off = 0;
      u.it_pre = 0;
      u.C = rows_prev;
      if (sghmc) {
        u.z = run->z_noise_dev ? run->z_noise_dev + ((size_t)t * run->z_noise_iters + (iters - 1)) * C * P : nullptr;
        u.ld_z = P;
        u.stream_lo = step;
        u.stream_hi = TAG_NOISE | (uint32_t)((((iters - 1) * nsw) + nsw - 1) & 0xffffff);
      }
      BHMC_TRY(launch_hmc_update(ctx, u));
    }
    // 4. Metropolis test (hmc.py:58-63): energies = NLP + 0.5|p|^2, momentum flipped for HMC only
    BHMC_TRY(launch_kinetic(ctx, s->p_new, ld, P, C, kin1));
    if (need_extra)
      BHMC_TRY(launch_prior_energy(ctx, s->q_new, ld, C, mb->n_vars, mb->var_off, mb->var_len, cv, sumsq, extra_new));
    AcceptArgs a{};
    a.q = s->q;
    a.q_new = s->q_new;
    a.p_out = s->p;
    a.p_new = s->p_new;
    a.ld = ld;
    a.P = P;
    a.C = C;
    a.p_sign = sghmc ? 1.f : -1.f;
    a.stat_cur = stat_cur;
    a.stat_new = stat_new;
    a.ea = ea;
    a.eb = eb;
    a.extra_cur = need_extra ? extra_cur : nullptr;
    a.extra_new = need_extra ? extra_new : nullptr;
    a.kin0 = kin0;
    a.kin1 = kin1;
    a.u = run->u_accept_host ? uacc_d + (size_t)t * C : nullptr;
    a.seed = cfg.seed;
    a.chain_id0 = cfg.chain_id0;
    a.stream_lo = step;
    a.stream_hi = TAG_ACCEPT;
    a.reject_nan = cfg.reject_nan;
    a.sample = run->samples_dev ? run->samples_dev + (size_t)t * C * P : nullptr;
    a.loss = run->loss_dev ? run->loss_dev + (size_t)t * C : nullptr;
    a.accept_prob = run->accept_prob_dev ? run->accept_prob_dev + (size_t)t * C : nullptr;
    a.accepted = run->accepted_dev ? run->accepted_dev + (size_t)t * C : nullptr;
    a.perm = perm;
    BHMC_TRY(launch_accept(ctx, a));
  }
  return BHMC_OK;
}

// ---- SGLD / SGD ------------------------------------------------------------------------------------
int bhmc_sampler_sg_run(bhmc_sampler* s, bhmc_sg_run* run) {
  NvtxRange nvtx_run("bhmc_sampler_sg_run");
  BHMC_CHECK_ARG(s && run, "NULL argument");
  const bhmc_sampler_config& cfg = s->cfg;
  BHMC_CHECK_ARG(cfg.kind == BHMC_KIND_SGLD || cfg.kind == BHMC_KIND_SGD, "sampler kind %d cannot run sg_run", cfg.kind);
  bhmc_ctx* ctx = s->ctx;
  ModelBase* mb = s->model;
  BHMC_CUDA_OK(cudaSetDevice(ctx->device));
  const int C = cfg.n_chains;
  const int64_t P = s->P, ld = s->ld;
  const int64_t n_rows = run->n_rows > 0 ? run->n_rows : mb->default_rows();
  BHMC_CHECK_ARG(run->batch_size > 0 && run->batch_size <= n_rows, "batch_size %lld must be in [1, %lld]",
                 (long long)run->batch_size, (long long)n_rows);
  BHMC_CHECK_ARG(run->epochs >= 0 && run->burnin >= 0 && run->step_size > 0, "bad run parameters");
  // sequential windows, remainder dropped (sgmcmc.py:34-38)
  const int64_t nb = (n_rows - run->batch_size) / run->batch_size + 1;
  const double num_batches = std::ceil((double)n_rows / (double)run->batch_size);  // sgmcmc.py:50
  const double decay = run->step_size / num_batches;                                // sgmcmc.py:51
  double* stat = s->scal;
  double eps = run->first_step_size > 0.0 ? run->first_step_size : run->step_size;
  int64_t k = 0;  // global minibatch counter (indexes the injected tape and the Philox stream)
  run->n_grad_evals = 0;
  const bool sgd = cfg.kind == BHMC_KIND_SGD;
  // sgd.py:35,57: every fit() / fit_dropout() call starts from momentum = zeros_like(par); the sampler object (and
  // its momentum buffer) outlives the call, so it is cleared here unless the caller chains calls on purpose
  if (sgd && !run->keep_momentum)
    BHMC_CUDA_OK(cudaMemsetAsync(s->p, 0, sizeof(float) * (size_t)C * ld, ctx->stream));
  // fused gradient + update + next operand preparation (tensor-core softmax path, single GPU): 3 launches per
  // minibatch instead of 5.  BHMC_FUSED_STEP=0 keeps the separate kernels (A/B measurements).
  static int fused_env = -1;
  if (fused_env < 0) {
    const char* e = getenv("BHMC_FUSED_STEP");
    fused_env = e ? atoi(e) : 1;
  }
  bool try_fused = fused_env && !s->hook && !s->row_comm && !(run->dropout_keep > 0.0);
  bool wt_ready = false;
  // Persistent path (softmax_persist.cuh): a whole epoch of minibatch steps in ONE cooperative launch -- forward,
  // grid barrier, backward with the update applied from tensor memory, grid barrier, per step.  The step sizes of
  // every step of the call are computed here (the schedule is host arithmetic, sgmcmc.py:67-73) and uploaded once.
  // BHMC_PERSIST=0 keeps the per-step launches (A/B measurements).
  static int persist_env = -1;
  if (persist_env < 0) {
    const char* e = getenv("BHMC_PERSIST");
    persist_env = e ? atoi(e) : 1;
  }
  bool try_persist = try_fused && persist_env && nb < (1 << 30);
  const float* eps_dev = nullptr;
  if (try_persist) {
    const int64_t total = (int64_t)(run->burnin + run->epochs) * nb;
    if (total > 0) {
      if (ctx->pinned_inflight) {
        BHMC_CUDA_OK(cudaEventSynchronize(ctx->pinned_ev));
        ctx->pinned_inflight = false;
      }
      void *pin = nullptr, *dev = nullptr;
      BHMC_TRY(ctx->get_pinned(sizeof(float) * (size_t)total, &pin));
      BHMC_TRY(ctx->get_scratch(14, sizeof(float) * (size_t)total, &dev));
      float* eh = (float*)pin;
      double ee = eps;
      int64_t kk = 0;
      for (int e = 0; e < run->burnin + run->epochs; ++e)
        for (int64_t j = 0; j < nb; ++j, ++kk) {
          eh[kk] = sgd ? (float)run->step_size : (float)ee;
          if (!sgd && e >= run->burnin) ee = run->step_size * (1.0 / (1.0 + (double)j * decay * num_batches));
        }
      BHMC_CUDA_OK(cudaMemcpyAsync(dev, pin, sizeof(float) * (size_t)total, cudaMemcpyHostToDevice, ctx->stream));
      if (!ctx->pinned_ev) BHMC_CUDA_OK(cudaEventCreateWithFlags(&ctx->pinned_ev, cudaEventDisableTiming));
      BHMC_CUDA_OK(cudaEventRecord(ctx->pinned_ev, ctx->stream));
      ctx->pinned_inflight = true;
      eps_dev = (const float*)dev;
    }
  }
  for (int e = 0; e < run->burnin + run->epochs; ++e) {
    const bool sampling = e >= run->burnin;
    if (try_persist && eps_dev) {
      FusedStep fs{};
      fs.kind = cfg.kind;
      fs.q = s->q;
      fs.p = s->p;
      fs.gamma = (float)run->gamma;
      fs.z = run->z_dev ? run->z_dev + (size_t)k * C * P : nullptr;
      fs.ld_z = P;
      fs.seed = cfg.seed;
      fs.chain_id0 = cfg.chain_id0;
      ctx->cur_units = C;
      const int rc = mb->sg_steps_persistent(C, ld, cfg.precision, fs, 0, run->batch_size, (int)nb, eps_dev + k, (int64_t)C * P,
                                             (uint64_t)(run->step0 + k));
      if (rc == BHMC_OK) {
        run->n_grad_evals += (int64_t)C * nb;
        if (!sgd && sampling) eps = run->step_size * (1.0 / (1.0 + (double)(nb - 1) * decay * num_batches));
        k += nb;
        wt_ready = false;
        goto epoch_end;
      }
      if (rc != BHMC_ERR_UNSUPPORTED) return rc;
      try_persist = false;
    }
    for (int64_t j = 0; j < nb; ++j, ++k) {
      const int64_t row0 = j * run->batch_size;
      if (try_fused) {
        FusedStep fs{};
        fs.kind = cfg.kind;
        fs.q = s->q;
        fs.p = s->p;
        fs.eps = sgd ? (float)run->step_size : (float)eps;
        fs.gamma = (float)run->gamma;
        fs.z = run->z_dev ? run->z_dev + (size_t)k * C * P : nullptr;
        fs.ld_z = P;
        fs.seed = cfg.seed;
        fs.chain_id0 = cfg.chain_id0;
        fs.stream_lo = (uint32_t)(run->step0 + k);
        fs.stream_hi = TAG_NOISE | (uint32_t)(((uint64_t)(run->step0 + k) >> 32) & 0xffffff);
        fs.wt_ready = wt_ready;
        const int rc = mb->grad_fused_step(C, ld, row0, run->batch_size, cfg.precision, stat, fs);
        if (rc == BHMC_OK) {
          wt_ready = true;
          run->n_grad_evals += C;
          if (!sgd && sampling) eps = run->step_size * (1.0 / (1.0 + (double)j * decay * num_batches));
          continue;
        }
        if (rc != BHMC_ERR_UNSUPPORTED) return rc;
        try_fused = false;
      }
      if (run->dropout_keep > 0.0) {
        BHMC_CHECK_ARG(sgd && !s->hook && !s->row_comm, "input dropout is sgd.fit_dropout only (single GPU)");
        const uint8_t* mk = run->mask_dev ? run->mask_dev + (size_t)k * run->batch_size * mb->n_features() : nullptr;
        BHMC_TRY(mb->grad_input_dropout(s->q, C, ld, row0, run->batch_size, cfg.precision, s->g, stat, mk,
                                        (float)run->dropout_keep, cfg.seed, (uint32_t)(run->step0 + k)));
      } else {
        BHMC_TRY(s->eval(s->q, C, row0, run->batch_size, s->g, stat));
      }
      run->n_grad_evals += C;
      if (sgd) {
        BHMC_TRY(launch_sgd_update(ctx, s->q, s->p, s->g, ld, P, C, (float)run->gamma, (float)run->step_size));
      } else {
        SgldArgs a{};
        a.q = s->q;
        a.p = s->p;
        a.g = s->g;
        a.ld = ld;
        a.P = P;
        a.C = C;
        a.eps = (float)eps;
        a.z = run->z_dev ? run->z_dev + (size_t)k * C * P : nullptr;
        a.ld_z = P;
        a.seed = cfg.seed;
        a.chain_id0 = cfg.chain_id0;
        a.stream_lo = (uint32_t)(run->step0 + k);
        a.stream_hi = TAG_NOISE | (uint32_t)(((uint64_t)(run->step0 + k) >> 32) & 0xffffff);
        BHMC_TRY(launch_sgld_update(ctx, a));
        // step size re-assigned AFTER batch j, sampling epochs only (sgmcmc.py:72-73,88-89)
        if (sampling) eps = run->step_size * (1.0 / (1.0 + (double)j * decay * num_batches));
      }
    }
  epoch_end:
    if (sampling) {
      const int i = e - run->burnin;
      if (run->logp_dev) {  // NLP(q, last batch), sgmcmc.py:79 / sgd.py:42
        const int64_t row0 = (nb - 1) * run->batch_size;
        double a, b, cv[BHMC_MAX_VARS];
        mb->energy_coeffs(run->batch_size, &a, &b, cv);
        BHMC_TRY(s->eval(s->q, C, row0, run->batch_size, nullptr, stat));
        wt_ready = false;  // that forward accumulated into stat: the next fused step must re-zero it (k_tc_prep)
        if (run->dropout_keep > 0.0) {  // sgd.py:67: loss = -log_likelihood(par, last batch), no prior, no 1/n
          BHMC_TRY(launch_affine(ctx, stat, -1.0, 0.0, nullptr, run->logp_dev + (size_t)i * C, C));
          if (run->samples_dev) BHMC_TRY(launch_copy_rows(ctx, s->q, ld, run->samples_dev + (size_t)i * C * P, P, P, C));
          continue;
        }
        bool need_extra = false;
        for (int v = 0; v < mb->n_vars; ++v) need_extra |= cv[v] != 0.0;
        double* extra = nullptr;
        if (need_extra) {
          extra = s->scal + 5 * (size_t)C;
          BHMC_TRY(launch_prior_energy(ctx, s->q, ld, C, mb->n_vars, mb->var_off, mb->var_len, cv, s->scal + 8 * (size_t)C, extra));
        }
        BHMC_TRY(launch_affine(ctx, stat, a, b, extra, run->logp_dev + (size_t)i * C, C));
      }
      if (run->samples_dev) BHMC_TRY(launch_copy_rows(ctx, s->q, ld, run->samples_dev + (size_t)i * C * P, P, P, C));
    }
  }
  run->final_step_size = eps;
  return BHMC_OK;
}

}  // extern "C"

extern "C" int bhmc_bench_update(bhmc_ctx* ctx, int32_t which, int32_t C, int64_t P, int32_t reps, double* ms_per_launch) {
  BHMC_CHECK_ARG(ctx && ms_per_launch && C > 0 && P > 0 && reps > 0 && which >= 0 && which <= 4, "bad argument");
  BHMC_CUDA_OK(cudaSetDevice(ctx->device));
  const int64_t ld = round_up(P, 4);
  const size_t row = (size_t)C * ld;
  float* buf = nullptr;
  double* scal = nullptr;
  int32_t* L = nullptr;
  BHMC_CUDA_OK(cudaMalloc(&buf, sizeof(float) * row * 6));
  BHMC_CUDA_OK(cudaMalloc(&scal, sizeof(double) * C * 8));
  BHMC_CUDA_OK(cudaMalloc(&L, sizeof(int32_t) * C));
  BHMC_CUDA_OK(cudaMemsetAsync(buf, 0, sizeof(float) * row * 6, ctx->stream));
  BHMC_CUDA_OK(cudaMemsetAsync(scal, 0, sizeof(double) * C * 8, ctx->stream));
  std::vector<int32_t> Lh(C, 1000000);
  BHMC_CUDA_OK(cudaMemcpyAsync(L, Lh.data(), sizeof(int32_t) * C, cudaMemcpyHostToDevice, ctx->stream));
  float *q = buf, *p = buf + row, *g = buf + 2 * row, *qn = buf + 3 * row, *pn = buf + 4 * row, *smp = buf + 5 * row;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int rc = BHMC_OK;
  for (int r = -3; r < reps && rc == BHMC_OK; ++r) {  // 3 warm-up launches
    if (r == 0) cudaEventRecord(e0, ctx->stream);
    if (which == 0 || which == 1) {
      UpdateArgs u{};
      u.q = q, u.p = p, u.g = g, u.ld = ld, u.P = P, u.C = C, u.L = L, u.eps = 1e-3f;
      u.post_off = 0, u.post_len = P, u.it_post = 0, u.pre_off = 0, u.pre_len = P, u.it_pre = 0;
      u.f_post = which ? 0.999f : 1.f, u.a_post = 1e-3f, u.n_post = which ? 2e-3f : 0.f, u.a_pre = which ? 0.f : 5e-4f;
      u.seed = 1, u.stream_lo = (uint32_t)r, u.stream_hi = TAG_NOISE;
      rc = launch_hmc_update(ctx, u);
    } else if (which == 2) {
      SgldArgs a{};
      a.q = q, a.p = p, a.g = g, a.ld = ld, a.P = P, a.C = C, a.eps = 1e-5f, a.seed = 1, a.stream_lo = (uint32_t)r, a.stream_hi = TAG_NOISE;
      rc = launch_sgld_update(ctx, a);
    } else if (which == 3) {
      AcceptArgs a{};
      a.q = q, a.q_new = qn, a.p_out = p, a.p_new = pn, a.ld = ld, a.P = P, a.C = C, a.p_sign = -1.f;
      a.stat_cur = scal, a.stat_new = scal + C, a.ea = 0.0, a.eb = 0.0, a.kin0 = scal + 2 * C, a.kin1 = scal + 3 * C;
      a.u = scal + 4 * C;  // u = 0 < A = 1 -> every chain accepts
      a.sample = smp;
      rc = launch_accept(ctx, a);
    } else {
      BeginArgs b{};
      b.q = q, b.q_new = qn, b.p0 = p, b.p_new = pn, b.ld = ld, b.P = P, b.C = C, b.seed = 1, b.stream_lo = (uint32_t)r;
      b.stream_hi = TAG_MOMENTUM, b.kin0 = scal + 2 * C;
      rc = launch_hmc_begin(ctx, b);
    }
  }
  cudaEventRecord(e1, ctx->stream);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  *ms_per_launch = (double)ms / reps;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  cudaFree(scal);
  cudaFree(L);
  return rc;
}
