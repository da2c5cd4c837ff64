"""ctypes binding of ``csrc/libbhmc.so`` (C ABI declared in ``include/bhmc.h``).

There is no CPU fallback and no alternative backend: if the library is missing this module
raises at import of the first symbol; if no B200 is visible ``bhmc_ctx_create`` fails and the
error is raised as ``BhmcError``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libbhmc.so")

BHMC_MAX_VARS = 8
PREC = {"fp32": 0, "bf16x3": 1, "bf16": 2}
KIND = {"hmc": 0, "sgld": 1, "sghmc": 2, "sgd": 3}
PRIOR = {"cpu": 0, "gpu": 1}


class BhmcError(RuntimeError):
    pass


class SamplerConfig(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("n_chains", C.c_int32), ("chain_id0", C.c_int64), ("seed", C.c_uint64),
        ("precision", C.c_int32), ("n_sweep", C.c_int32),
        ("sweep_off", C.c_int64 * BHMC_MAX_VARS), ("sweep_len", C.c_int64 * BHMC_MAX_VARS),
        ("shared_path", C.c_int32), ("leapfrog", C.c_int32), ("sghmc_descent", C.c_int32),
        ("reject_nan", C.c_int32), ("reserved", C.c_int32 * 4),
    ]


class HmcRun(C.Structure):
    _fields_ = [
        ("n_steps", C.c_int32), ("step_size", C.c_double), ("path_length", C.c_double),
        ("row0", C.c_int64), ("nrows", C.c_int64), ("step0", C.c_int64),
        ("z_momentum_dev", C.c_void_p), ("u_path_host", C.c_void_p), ("u_accept_host", C.c_void_p),
        ("z_noise_dev", C.c_void_p), ("z_noise_iters", C.c_int64),
        ("samples_dev", C.c_void_p), ("loss_dev", C.c_void_p), ("accept_prob_dev", C.c_void_p),
        ("accepted_dev", C.c_void_p),
        ("n_grad_evals", C.c_int64), ("n_grad_launched", C.c_int64),
        ("schedule", C.c_int32), ("n_phases", C.c_int32),
    ]


class SgRun(C.Structure):
    _fields_ = [
        ("epochs", C.c_int32), ("burnin", C.c_int32), ("batch_size", C.c_int64), ("n_rows", C.c_int64),
        ("step_size", C.c_double), ("gamma", C.c_double), ("step0", C.c_int64),
        ("z_dev", C.c_void_p), ("samples_dev", C.c_void_p), ("logp_dev", C.c_void_p),
        ("n_grad_evals", C.c_int64), ("final_step_size", C.c_double),
        ("dropout_keep", C.c_double), ("mask_dev", C.c_void_p),
        ("first_step_size", C.c_double), ("keep_momentum", C.c_int32), ("reserved_", C.c_int32),
    ]


_PROTOS = {
    # name: (restype, argtypes)
    "bhmc_version": (C.c_int, []),
    "bhmc_last_error": (C.c_char_p, []),
    "bhmc_ctx_create": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "bhmc_ctx_destroy": (C.c_int, [C.c_void_p]),
    "bhmc_ctx_sync": (C.c_int, [C.c_void_p]),
    "bhmc_ctx_launch_count": (C.c_int64, [C.c_void_p]),
    "bhmc_ctx_kernel_time": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "bhmc_ctx_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "bhmc_ctx_timing_stride": (C.c_int, [C.c_void_p, C.c_int]),
    "bhmc_ctx_kernel_units": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double)]),
    "bhmc_softmax_create": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_int32,
                                      C.POINTER(C.c_void_p)]),
    "bhmc_logistic_create": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.POINTER(C.c_void_p)]),
    "bhmc_softmax_bind_data": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "bhmc_softmax_bind_data_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "bhmc_softmax_operand_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float)]),
    "bhmc_mvn_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.POINTER(C.c_void_p)]),
    "bhmc_mlp_create": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float,
                                  C.c_uint64, C.c_int64, C.POINTER(C.c_void_p)]),
    "bhmc_mlp_bind_data": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "bhmc_mlp_set_masks": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bhmc_mlp_predict": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int64, C.c_int32,
                                   C.c_void_p, C.c_void_p]),
    "bhmc_model_set_global_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_float]),
    "bhmc_model_destroy": (C.c_int, [C.c_void_p]),
    "bhmc_model_n_params": (C.c_int64, [C.c_void_p]),
    "bhmc_model_n_vars": (C.c_int32, [C.c_void_p]),
    "bhmc_model_var_layout": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "bhmc_model_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                  C.c_void_p, C.c_void_p]),
    "bhmc_model_loglik": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                    C.c_void_p]),
    "bhmc_model_nlp": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                                 C.c_void_p]),
    "bhmc_softmax_predict": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int64,
                                       C.c_void_p, C.c_void_p]),
    "bhmc_philox_normal": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_uint64, C.c_int64,
                                     C.c_uint32, C.c_uint32]),
    "bhmc_philox_uniform_host": (C.c_double, [C.c_uint64, C.c_int64, C.c_uint32, C.c_uint32]),
    "bhmc_philox4x32_host": (None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "bhmc_bench_update": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.POINTER(C.c_double)]),
    "bhmc_sampler_create": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(SamplerConfig), C.POINTER(C.c_void_p)]),
    "bhmc_sampler_destroy": (C.c_int, [C.c_void_p]),
    "bhmc_sampler_set_grad_hook": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "bhmc_sampler_ld": (C.c_int64, [C.c_void_p]),
    "bhmc_sampler_state_ptr": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]),
    "bhmc_sampler_set_q": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "bhmc_sampler_get": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]),
    "bhmc_sampler_hmc_run": (C.c_int, [C.c_void_p, C.POINTER(HmcRun)]),
    "bhmc_sampler_sg_run": (C.c_int, [C.c_void_p, C.POINTER(SgRun)]),
    "bhmc_comm_unique_id": (C.c_int, [C.c_void_p]),
    "bhmc_comm_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "bhmc_comm_wrap": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "bhmc_comm_destroy": (C.c_int, [C.c_void_p]),
    "bhmc_comm_world": (C.c_int32, [C.c_void_p]),
    "bhmc_nccl_version": (C.c_int, []),
    "bhmc_allreduce_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32]),
    "bhmc_comm_allreduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32]),
    "bhmc_sampler_set_row_comm": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bhmc_stream_plan_host": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int64),
                                        C.POINTER(C.c_int64), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
}

GRAD_HOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64)

_lib = None


def exported_symbols():
    """Names every entry point include/bhmc.h declares (used by the CPU ABI test)."""
    return sorted(_PROTOS)


def lib():
    """Load libbhmc.so once; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BhmcError(
                "%s not found: build it with `python -m dropout_hamiltonian_montecarlo_b200.build` "
                "(nvcc, sm_100a). There is no CPU or PyTorch fallback." % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(handle, name)  # AttributeError = header/library drift
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise BhmcError("libbhmc error %d: %s" % (rc, lib().bhmc_last_error().decode("utf-8", "replace")))
