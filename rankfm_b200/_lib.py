"""ctypes binding of ``librankfm_b200.so`` (C ABI declared in ``include/rankfm_b200.h``).

Loading never falls back to a CPU implementation: if the library is missing it is built with nvcc, and if no CUDA
device is present every compute call raises ``RuntimeError`` (RFM_ERR_NO_DEVICE).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "librankfm_b200.so")

RFM_OK, RFM_ERR_ARG, RFM_ERR_CUDA, RFM_ERR_NO_DEVICE, RFM_ERR_NCCL, RFM_ERR_NONFINITE, RFM_ERR_UNSUPPORTED = range(7)
SCHEDULE = {"constant": 0, "invscaling": 1}
ORDER_FEISTEL, ORDER_HOST = 0, 1
SAMPLER_PHILOX, SAMPLER_MT = 0, 1
SCHED_PARALLEL, SCHED_SERIAL = 0, 1

EXPORTS = [
    "rfm_version", "rfm_last_error", "rfm_device_count", "rfm_nccl_unique_id", "rfm_comm_release_all", "rfm_host_register", "rfm_host_unregister", "rfm_debug_philox", "rfm_debug_feistel", "rfm_debug_feat8", "rfm_debug_tau_rank", "rfm_trim_device_cache",
    "rfm_prep_index_ids", "rfm_prep_user_items", "rfm_synth_zipf",
    "rfm_fit", "rfm_last_fit_phases", "rfm_predict", "rfm_recommend", "rfm_similar",
    "rfm_session_create", "rfm_session_train", "rfm_session_set_epoch_callback", "rfm_session_set_weights", "rfm_session_download",
    "rfm_session_snapshot", "rfm_session_restore", "rfm_session_timer_start", "rfm_session_timer_stop",
    "rfm_session_predict", "rfm_session_recommend", "rfm_session_time_predict", "rfm_session_time_recommend",
    "rfm_session_trace_enable", "rfm_session_trace_read", "rfm_session_debug_gemm", "rfm_session_recommend_stats", "rfm_session_recommend_retried", "rfm_session_attach_csr", "rfm_session_similar", "rfm_session_similar_batch", "rfm_session_evaluate", "rfm_session_flush_l2", "rfm_session_launch_count", "rfm_session_exchange_path", "rfm_session_destroy",
]


class Problem(C.Structure):
    _fields_ = [
        ("interactions", C.c_void_p), ("sample_weight", C.c_void_p), ("n_interactions", C.c_int64),
        ("csr_indptr", C.c_void_p), ("csr_indices", C.c_void_p),
        ("x_uf", C.c_void_p), ("x_if", C.c_void_p),
        ("w_i", C.c_void_p), ("w_if", C.c_void_p), ("v_u", C.c_void_p), ("v_i", C.c_void_p), ("v_uf", C.c_void_p), ("v_if", C.c_void_p),
        ("U", C.c_int32), ("I", C.c_int32), ("P", C.c_int32), ("Q", C.c_int32), ("F", C.c_int32),
        ("alpha", C.c_float), ("beta", C.c_float), ("learning_rate", C.c_float), ("learning_exponent", C.c_float),
        ("schedule", C.c_int32), ("max_samples", C.c_int32),
        ("order", C.c_int32), ("sampler", C.c_int32), ("sched", C.c_int32),
        ("mt_seed", C.c_uint32), ("seed", C.c_uint64), ("max_rejects", C.c_int32), ("device", C.c_int32),
        ("rank", C.c_int32), ("world", C.c_int32), ("nccl_id", C.c_void_p),
        ("user_lo", C.c_int32), ("user_hi", C.c_int32), ("epoch_offset", C.c_int32),
    ]


class EpochStats(C.Structure):
    _fields_ = [
        ("log_likelihood", C.c_double), ("penalty", C.c_double), ("draws", C.c_int64), ("finite", C.c_int32 * 6),
        ("eta", C.c_float), ("kernel_ms", C.c_float), ("sync_ms", C.c_float),
    ]


EPOCH_CALLBACK = C.CFUNCTYPE(None, C.c_int32, C.POINTER(EpochStats), C.c_void_p)

_lib = None


def lib():
    """load (building first if the .so is absent or stale) and declare prototypes"""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    variant = os.environ.get("RANKFM_B200_LIB")          # experiments: a variant build of the library (rankfm_b200/build.py --out=...)
    if variant:
        L = C.CDLL(variant)
    else:
        if _build.needs_build():
            _build.build()
        L = C.CDLL(LIB_PATH)
    L.rfm_version.restype = C.c_char_p
    L.rfm_last_error.restype = C.c_char_p
    L.rfm_device_count.restype = C.c_int
    pp, vp, i32, i64 = C.POINTER(Problem), C.c_void_p, C.c_int32, C.c_int64
    L.rfm_nccl_unique_id.argtypes = [vp]
    L.rfm_host_register.argtypes = [vp, C.c_uint64]
    L.rfm_host_unregister.argtypes = [vp]
    L.rfm_debug_philox.argtypes = [C.c_uint32] * 6 + [vp]
    L.rfm_debug_feistel.argtypes = [i64, C.c_uint64, i32, i64, i64, vp]
    L.rfm_debug_feat8.argtypes = [i32, i32, i32, C.c_uint32, vp]
    L.rfm_debug_tau_rank.argtypes = [i32, i32, C.c_float]
    L.rfm_prep_index_ids.argtypes = [vp, i64, i32, vp, vp, vp]
    L.rfm_prep_user_items.argtypes = [vp, i64, i32, i32, i32, vp, vp]
    L.rfm_synth_zipf.argtypes = [i32, i32, i64, C.c_double, C.c_double, C.c_uint64, C.c_uint64, i32, i32, i32, vp, vp, vp, vp]
    L.rfm_fit.argtypes = [pp, i32, vp, vp]
    L.rfm_last_fit_phases.argtypes = [vp]
    L.rfm_predict.argtypes = [pp, vp, i64, vp]
    L.rfm_recommend.argtypes = [pp, vp, i64, i32, i32, vp]
    L.rfm_similar.argtypes = [pp, i32, i32, i32, vp]
    L.rfm_session_create.argtypes = [pp, C.POINTER(vp)]
    L.rfm_session_train.argtypes = [vp, i32, vp, vp]
    L.rfm_session_set_epoch_callback.argtypes = [vp, EPOCH_CALLBACK, vp]
    L.rfm_session_set_weights.argtypes = [vp] + [vp] * 6
    L.rfm_session_download.argtypes = [vp] + [vp] * 6
    L.rfm_session_snapshot.argtypes = [vp]
    L.rfm_session_restore.argtypes = [vp]
    L.rfm_session_timer_start.argtypes = [vp]
    L.rfm_session_timer_stop.argtypes = [vp, vp]
    L.rfm_session_predict.argtypes = [vp, vp, i64, vp]
    L.rfm_session_recommend.argtypes = [vp, vp, i64, i32, i32, vp]
    L.rfm_session_time_predict.argtypes = [vp, vp, i64, i32, vp]
    L.rfm_session_time_recommend.argtypes = [vp, vp, i64, i32, i32, i32, vp, vp]
    L.rfm_session_trace_enable.argtypes = [vp]
    L.rfm_session_trace_read.argtypes = [vp, vp]
    L.rfm_session_debug_gemm.argtypes = [vp, vp, i64, vp]
    L.rfm_session_recommend_stats.argtypes = [vp, vp, vp]
    L.rfm_session_recommend_retried.argtypes = [vp, vp]
    L.rfm_session_attach_csr.argtypes = [vp, vp, vp]
    L.rfm_session_similar.argtypes = [vp, i32, i32, i32, vp]
    L.rfm_session_similar_batch.argtypes = [vp, i32, vp, i64, i32, vp]
    L.rfm_session_evaluate.argtypes = [vp, vp, i64, i32, i32, vp, vp, vp, vp, vp]
    L.rfm_session_flush_l2.argtypes = [vp]
    L.rfm_session_launch_count.argtypes = [vp, vp]
    L.rfm_session_exchange_path.argtypes = [vp, vp]
    L.rfm_comm_release_all.argtypes = []
    L.rfm_session_destroy.argtypes = [vp]
    for name in EXPORTS:
        if name not in ("rfm_version", "rfm_last_error"):
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def last_error():
    return lib().rfm_last_error().decode("utf-8", "replace")


def check(rc):
    """map a status code to the exception type the reference raises for the same condition"""
    if rc == RFM_OK:
        return
    msg = last_error()
    if rc == RFM_ERR_NONFINITE:
        raise AssertionError(msg)                      # reference: assert_finite, _rankfm.pyx:95-103
    if rc == RFM_ERR_ARG:
        raise ValueError(msg)
    raise RuntimeError("rankfm_b200: %s (code %d)" % (msg, rc))


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def as_buffer(a, dtype, ndim, name):
    """the Cython memoryviews of the reference reject wrong dtypes / non-contiguous input with ValueError"""
    if not isinstance(a, np.ndarray) or a.dtype != dtype or a.ndim != ndim or not a.flags.c_contiguous:
        raise ValueError("Buffer dtype mismatch or not C-contiguous for [%s]: expected %s ndim=%d" % (name, np.dtype(dtype).name, ndim))
    return a
