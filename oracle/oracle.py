"""ctypes front-end of the CPU oracle (``rankfm_oracle.c``) and loader of the compiled reference (``oracle/_ref``).

TEST / BASELINE INFRASTRUCTURE ONLY -- see the header of ``rankfm_oracle.c``.  The product package
``rankfm_b200`` never imports this module; it exists so that ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs can check / time the hot path on the host.

The functions keep the reference's positional signatures (``_rankfm.pyx:122-142,345-355,393-406``) so a test can
call ``oracle._fit(...)``, the reference's ``_fit(...)`` and ``rankfm_b200._rankfm._fit(...)`` interchangeably.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "rankfm_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["gcc", "-O2", "-fPIC", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-shared", src, "-o", so, "-lm"], check=True)
    return so


class _FitArgs(C.Structure):
    _fields_ = [
        ("interactions", C.c_void_p), ("sample_weight", C.c_void_p), ("indptr", C.c_void_p), ("indices", C.c_void_p),
        ("x_uf", C.c_void_p), ("x_if", C.c_void_p),
        ("w_i", C.c_void_p), ("w_if", C.c_void_p), ("v_u", C.c_void_p), ("v_i", C.c_void_p), ("v_uf", C.c_void_p), ("v_if", C.c_void_p),
        ("N", C.c_int64),
        ("U", C.c_int32), ("I", C.c_int32), ("P", C.c_int32), ("Q", C.c_int32), ("F", C.c_int32),
        ("alpha", C.c_float), ("beta", C.c_float), ("learning_rate", C.c_float), ("learning_exponent", C.c_float),
        ("schedule", C.c_int32), ("max_samples", C.c_int32), ("epochs", C.c_int32),
        ("perms", C.c_void_p), ("sampler", C.c_int32), ("mt_seed", C.c_uint32), ("seed", C.c_uint64),
        ("epoch_offset", C.c_int32), ("max_rejects", C.c_int32),
        ("out_ll", C.c_void_p), ("out_draws", C.c_void_p), ("out_neg", C.c_void_p),
    ]


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_fit.argtypes = [C.POINTER(_FitArgs)]
        _LIB.orc_fit.restype = C.c_int
        _LIB.orc_feistel_perm.argtypes = [C.c_int64, C.c_int64, C.c_uint64, C.c_int]
        _LIB.orc_feistel_perm.restype = C.c_int64
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _chk(a, dtype, ndim):
    assert isinstance(a, np.ndarray) and a.dtype == dtype and a.ndim == ndim and a.flags.c_contiguous, \
        "Buffer dtype mismatch / not C-contiguous"
    return a


def user_items_to_csr(user_items, U):
    """dict {u: sorted int32 array} -> (indptr int64[U+1], indices int32[nnz]); mirrors _rankfm.pyx:201-212"""
    if hasattr(user_items, "indptr") and hasattr(user_items, "indices"):
        return np.ascontiguousarray(user_items.indptr, dtype=np.int64), np.ascontiguousarray(user_items.indices, dtype=np.int32)
    lens = np.fromiter((len(user_items[u]) for u in range(U)), dtype=np.int64, count=U)
    indptr = np.zeros(U + 1, dtype=np.int64)
    np.cumsum(lens, out=indptr[1:])
    indices = np.concatenate([np.asarray(user_items[u], dtype=np.int32) for u in range(U)]) if U else np.zeros(0, np.int32)
    return indptr, np.ascontiguousarray(indices, dtype=np.int32)


def mt_stream(seed, n):
    out = np.empty(n, dtype=np.uint32)
    lib().orc_mt_stream(C.c_uint32(seed), _p(out), C.c_int(n))
    return out


def philox4x32(ctr, key):
    out = np.empty(4, dtype=np.uint32)
    lib().orc_philox4x32(*[C.c_uint32(int(c)) for c in ctr], *[C.c_uint32(int(k)) for k in key], _p(out))
    return out


def feistel_perm(N, seed, epoch):
    f = lib().orc_feistel_perm
    return np.fromiter((f(r, N, seed, epoch) for r in range(N)), dtype=np.int64, count=N)


def fit_ex(interactions, sample_weight, user_items, x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if,
           alpha, beta, learning_rate, learning_schedule, learning_exponent, max_samples, epochs,
           perms=None, sampler="mt", mt_seed=1492, seed=0, epoch_offset=0, max_rejects=0, want_neg=False):
    """sequential SGD on the host.  ``perms`` = int32 [epochs, N] row order per epoch (what ``np.random.shuffle``
    produced in the reference) or None for the Feistel order; returns dict(ll, draws, neg)"""
    _chk(interactions, np.int32, 2); _chk(sample_weight, np.float32, 1)
    for a, nd in ((x_uf, 2), (x_if, 2), (w_i, 1), (w_if, 1), (v_u, 2), (v_i, 2), (v_uf, 2), (v_if, 2)):
        _chk(a, np.float32, nd)
    if learning_schedule not in ("constant", "invscaling"):
        raise ValueError('unknown [learning_schedule]')
    N, (U, F), I, P, Q = interactions.shape[0], v_u.shape, v_i.shape[0], v_uf.shape[0], v_if.shape[0]
    indptr, indices = user_items_to_csr(user_items, U)
    out_ll = np.zeros(epochs, dtype=np.float32)
    out_draws = np.zeros(epochs, dtype=np.int64)
    out_neg = np.zeros((epochs, N), dtype=np.int32) if want_neg else None
    if perms is not None:
        perms = np.ascontiguousarray(perms, dtype=np.int32)
        assert perms.shape == (epochs, N)
    a = _FitArgs(_p(interactions), _p(sample_weight), _p(indptr), _p(indices), _p(x_uf), _p(x_if),
                 _p(w_i), _p(w_if), _p(v_u), _p(v_i), _p(v_uf), _p(v_if), N, U, I, P, Q, F,
                 alpha, beta, learning_rate, learning_exponent, 0 if learning_schedule == "constant" else 1,
                 max_samples, epochs, _p(perms) if perms is not None else None,
                 0 if sampler == "mt" else 1, mt_seed, seed, epoch_offset, max_rejects,
                 _p(out_ll), _p(out_draws), _p(out_neg) if want_neg else None)
    rc = lib().orc_fit(C.byref(a))
    if rc != 0:
        raise RuntimeError("oracle fit failed rc=%d" % rc)
    return {"ll": out_ll, "draws": out_draws, "neg": out_neg}


def _fit(interactions, sample_weight, user_items, x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if,
         alpha, beta, learning_rate, learning_schedule, learning_exponent, max_samples, epochs, verbose):
    """reference-identical ``_fit`` (_rankfm.pyx:122-342): consumes ``np.random.shuffle`` like the reference does"""
    N = interactions.shape[0]
    shuffle_index = np.arange(N, dtype=np.int32)
    perms = np.empty((epochs, N), dtype=np.int32)
    for e in range(epochs):
        np.random.shuffle(shuffle_index)
        perms[e] = shuffle_index
    out = fit_ex(interactions, sample_weight, user_items, x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if,
                 alpha, beta, learning_rate, learning_schedule, learning_exponent, max_samples, epochs, perms=perms)
    assert_finite(w_i, w_if, v_u, v_i, v_uf, v_if)
    if verbose:
        for e in range(epochs):
            print("\ntraining epoch:", e)
            print("log likelihood:", out["ll"][e])
    return out


def assert_finite(w_i, w_if, v_u, v_i, v_uf, v_if):
    """_rankfm.pyx:95-103"""
    assert np.isfinite(np.sum(w_i)), "item weights [w_i] are not finite - try decreasing feature/sample_weight magnitudes"
    assert np.isfinite(np.sum(w_if)), "item feature weights [w_if] are not finite - try decreasing feature/sample_weight magnitudes"
    assert np.isfinite(np.sum(v_u)), "user factors [v_u] are not finite - try decreasing feature/sample_weight magnitudes"
    assert np.isfinite(np.sum(v_i)), "item factors [v_i] are not finite - try decreasing feature/sample_weight magnitudes"
    assert np.isfinite(np.sum(v_uf)), "user-feature factors [v_uf] are not finite - try decreasing feature/sample_weight magnitudes"
    assert np.isfinite(np.sum(v_if)), "item-feature factors [v_if] are not finite - try decreasing feature/sample_weight magnitudes"


def reg_penalty(alpha, beta, w_i, w_if, v_u, v_i, v_uf, v_if):
    """_rankfm.pyx:106-116"""
    return float(sum(np.sum(c * np.square(w)) for c, w in
                     ((alpha, w_i), (alpha, v_u), (alpha, v_i), (beta, w_if), (beta, v_uf), (beta, v_if))))


def _predict(pairs, x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if):
    """_rankfm.pyx:345-390"""
    _chk(pairs, np.float32, 2)
    N, (U, F), I, P, Q = pairs.shape[0], v_u.shape, v_i.shape[0], v_uf.shape[0], v_if.shape[0]
    scores = np.empty(N, dtype=np.float32)
    lib().orc_predict(_p(pairs), C.c_int64(N), _p(x_uf), _p(x_if), _p(w_i), _p(w_if), _p(v_u), _p(v_i), _p(v_uf), _p(v_if),
                      C.c_int(U), C.c_int(I), C.c_int(P), C.c_int(Q), C.c_int(F), _p(scores))
    return scores


def scores_user(u, x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if):
    (U, F), I, P, Q = v_u.shape, v_i.shape[0], v_uf.shape[0], v_if.shape[0]
    scores = np.empty(I, dtype=np.float32)
    lib().orc_scores_user(C.c_int(int(u)), _p(x_uf), _p(x_if), _p(w_i), _p(w_if), _p(v_u), _p(v_i), _p(v_uf), _p(v_if),
                          C.c_int(U), C.c_int(I), C.c_int(P), C.c_int(Q), C.c_int(F), _p(scores))
    return scores


def _recommend(users, user_items, n_items, filter_previous, x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if):
    """_rankfm.pyx:393-460: all-item scores in C, ranking with the same NumPy call the reference makes (:444)"""
    _chk(users, np.float32, 1)
    rec_items = np.empty((users.shape[0], n_items), dtype=np.float32)
    for row, u_flt in enumerate(users):
        if np.isnan(u_flt):
            rec_items[row] = np.nan
            continue
        u = int(u_flt)
        ranked = np.argsort(scores_user(u, x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if))[::-1]
        if filter_previous:
            ranked = ranked[~np.isin(ranked, user_items[u])]
        sel = ranked[:n_items]
        rec_items[row, :len(sel)] = sel
        rec_items[row, len(sel):] = np.nan      # the reference leaves this tail uninitialised (:445)
    return rec_items


# ---------------------------------------------------------------------------------------------------------------
# the compiled reference itself (oracle/_ref), when it has been built
# ---------------------------------------------------------------------------------------------------------------
def load_reference(build_if_possible=True):
    """returns the reference's compiled ``rankfm._rankfm`` module (functions _fit/_predict/_recommend) or None"""
    sys.path.insert(0, HERE)
    try:
        import build_ref
    finally:
        sys.path.pop(0)
    so = build_ref.build(verbose=False) if build_if_possible else build_ref.ref_so_path()
    if not so or not os.path.exists(so):
        return None
    ref_root = os.path.join(HERE, "_ref")
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    import importlib
    return importlib.import_module("rankfm._rankfm")


def load_reference_class():
    """the reference's pure-Python ``RankFM`` class; only importable where /root/reference exists"""
    if load_reference() is None:
        return None
    try:
        from rankfm.rankfm import RankFM
        return RankFM
    except Exception:
        return None
