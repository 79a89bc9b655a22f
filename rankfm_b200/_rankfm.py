"""Drop-in replacement of the reference's Cython module ``rankfm._rankfm``.

Exports ``_fit``, ``_predict`` and ``_recommend`` with the reference's positional signatures
(``rankfm/_rankfm.pyx:122-142``, ``:345-355``, ``:393-406``; imported at ``rankfm/rankfm.py:8``) and forwards them
through ctypes to the hand-written sm_100a kernels in ``librankfm_b200.so``.  Host code stays NumPy; there is no
PyTorch and no CPU fallback on this path.

Execution modes of ``_fit`` (``set_mode`` / env ``RANKFM_B200_MODE``):

``production`` (default)
    Hogwild SGD over the whole GPU, Philox negatives, on-device per-epoch permutation.  Statistically equivalent
    to the reference; not the same trajectory.
``replay``
    One lane group walks the positives strictly in the order ``np.random.shuffle`` produces (consuming NumPy's
    global RNG exactly like ``_rankfm.pyx:227``) and draws negatives from the reference's MT19937 stream seeded
    1492 (``:182``): the same trajectory as the reference up to float reassociation.  Parity/test mode.
"""
import ctypes as C
import os
import zlib

import numpy as np

from . import _lib
from ._lib import Problem, EpochStats, as_buffer, check, ptr

WEIGHT_NAMES = ("w_i", "w_if", "v_u", "v_i", "v_uf", "v_if")
_MODE = os.environ.get("RANKFM_B200_MODE", "production")
_SEED = int(os.environ.get("RANKFM_B200_SEED", "1492"))
_DEVICE = int(os.environ.get("RANKFM_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
_COMM = {"rank": 0, "world": 1, "nccl_id": None, "user_range": None}
last_stats = None       # list of per-epoch dicts of the most recent _fit
# Resident TRAINING session (the reference's warm-start contract is `fit_partial`, rankfm.py:269-327):
# `_fit` keeps the session of its last call -- interactions, sample weights, user_items CSR, membership bitmap, side
# features and (multi-GPU) the communicator stay in HBM -- and a following `_fit` on the SAME data and hyper-parameters
# only uploads the weight arrays it is given, trains, and writes them back.  Any other data builds a new session.
# RANKFM_B200_RESIDENT_TRAIN=0 (or set_resident_training(False)) restores one throw-away session per call.
_RESIDENT_TRAIN = os.environ.get("RANKFM_B200_RESIDENT_TRAIN", "1") == "1"
_training = {"key": None, "sess": None, "hits": 0, "builds": 0}
_EPOCHS = {"done": 0}   # epochs trained by earlier production-mode `_fit` calls: offsets the Philox / Feistel keys of new sessions
# opt-in: keep one scoring session (packed weights, bf16 item operand, user_items CSR) resident in HBM across
# _predict / _recommend calls.  The stateless functions of the reference re-read the weight arrays on
# every call; with this switch on they are re-uploaded only when a DIFFERENT set of arrays is passed or after `_fit`
# (which trains them in place) -- a caller who edits the arrays in place himself must call `drop_resident()`.
_RESIDENT = os.environ.get("RANKFM_B200_RESIDENT", "0") == "1"
_scoring = {"key": None, "sess": None, "csr": None, "uploads": 0}


def set_mode(mode):
    global _MODE
    assert mode in ("production", "replay"), "[mode] must be in ('production', 'replay')"
    _MODE = mode


def get_mode():
    return _MODE


def set_resident(flag):
    global _RESIDENT
    _RESIDENT = bool(flag)
    if not _RESIDENT:
        drop_resident()


def drop_resident():
    """forget the resident scoring session (its weights are stale)"""
    if _scoring["sess"] is not None:
        _scoring["sess"].close()
    _scoring.update(key=None, sess=None, csr=None)


def _resident_session(weights, user_items=None):
    """the cached scoring session for exactly these weight arrays (same buffers, same shapes), created on first use"""
    key = tuple((a.ctypes.data, a.shape) for a in weights) + (_DEVICE,)
    if _scoring["key"] != key:
        drop_resident()
        keep = []
        _scoring.update(key=key, sess=Session(_problem(*weights, keep), keep), csr=None)
        _scoring["uploads"] += 1
    sess = _scoring["sess"]
    if user_items is not None:
        from_csr = hasattr(user_items, "indptr") and hasattr(user_items, "indices")
        indptr, indices = user_items_to_csr(user_items, weights[4].shape[0])
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        csr_key = (_fingerprint(indptr, from_csr), _fingerprint(indices, from_csr))
        if _scoring["csr"] != csr_key:
            sess.attach_csr(indptr, indices)          # copied to the device: nothing of the old or new CSR has to stay alive
            _scoring["csr"] = csr_key
    return sess


def set_resident_training(flag):
    global _RESIDENT_TRAIN
    _RESIDENT_TRAIN = bool(flag)
    if not _RESIDENT_TRAIN:
        drop_training()


def drop_training():
    """free the resident training session (its HBM goes back to the library's block cache)"""
    if _training["sess"] is not None:
        _training["sess"].close()
    _training.update(key=None, sess=None)


def _fingerprint(a, with_address=True):
    """cheap identity of a (large, read-only) input array: shape, dtype, [address,] CRC of <= 64k strided samples"""
    if a is None:
        return None
    flat = a.reshape(-1)
    step = max(1, flat.size // 65536)
    crc = zlib.crc32(np.ascontiguousarray(flat[::step]).tobytes()) if flat.size else 0
    return (a.ctypes.data if with_address else 0, a.shape, a.dtype.str, crc)


def set_seed(seed):
    global _SEED
    _SEED = int(seed)
    _EPOCHS["done"] = 0
    drop_training()


def set_device(device):
    global _DEVICE
    _DEVICE = int(device)


def set_comm(rank, world, nccl_id, user_range=None):
    """multi-GPU: one process per GPU; ``nccl_id`` = the 128 bytes of ``nccl_unique_id()`` made on rank 0 and
    broadcast by the caller.  Each rank then passes ITS shard of the interactions (see ``shard_by_user``) and owns the
    users ``user_range = (lo, hi)`` (default: the smallest range that covers the shard's users): only those rows of
    ``v_u`` are trained and written back on this rank, the item side is identical on every rank after every epoch.
    The library keeps the communicator of an id for all later calls (``release_comms`` frees it)."""
    drop_training()
    _COMM.update(rank=int(rank), world=int(world), nccl_id=None if nccl_id is None else np.frombuffer(bytes(nccl_id), dtype=np.uint8).copy(),
                 user_range=None if user_range is None else (int(user_range[0]), int(user_range[1])))


def release_comms():
    """destroy the cached communicators no session uses any more (collective over the ranks of each job)"""
    drop_training()
    check(_lib.lib().rfm_comm_release_all())


def nccl_unique_id():
    out = np.zeros(128, dtype=np.uint8)
    check(_lib.lib().rfm_nccl_unique_id(ptr(out)))
    return out.tobytes()


def pin(*arrays):
    """page-lock NumPy buffers in place (cudaHostRegister) so `_fit/_predict/_recommend` upload them at PCIe speed"""
    for a in arrays:
        if a is not None and a.nbytes:
            check(_lib.lib().rfm_host_register(ptr(a), a.nbytes))


def unpin(*arrays):
    for a in arrays:
        if a is not None and a.nbytes:
            _lib.lib().rfm_host_unregister(ptr(a))


def device_count():
    return _lib.lib().rfm_device_count()


class UserItems(dict):
    """``user_items`` (``rankfm/rankfm.py:174``: ``{user_idx: sorted int32 array of item_idx}``) backed by one CSR.

    Behaves like the reference's dict (``d[u]``, ``keys()``, ``items()``, ``len``) but the per-user arrays are views
    created on first access, and ``_fit``/``_recommend`` read ``indptr``/``indices`` directly instead of walking
    ``U`` Python objects like ``_rankfm.pyx:201-212`` does."""

    def __init__(self, indptr, indices):
        super().__init__()
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32)
        self._n = len(self.indptr) - 1

    def __missing__(self, u):
        ui = int(u)
        if ui != u or not 0 <= ui < self._n:
            raise KeyError(u)
        view = self.indices[self.indptr[ui]:self.indptr[ui + 1]]
        dict.__setitem__(self, ui, view)
        return view

    def __contains__(self, u):
        try:
            return 0 <= int(u) < self._n and int(u) == u
        except (TypeError, ValueError):
            return False

    def __len__(self):
        return self._n

    def __iter__(self):
        return iter(range(self._n))

    def keys(self):
        return range(self._n)

    def values(self):
        return (self[u] for u in range(self._n))

    def items(self):
        return ((u, self[u]) for u in range(self._n))

    def get(self, u, default=None):
        return self[u] if u in self else default

    def __reduce__(self):
        return (UserItems, (self.indptr, self.indices))

    @classmethod
    def from_interactions(cls, interactions, n_users, n_items=None):
        """CSR of the items of each user, sorted ascending, duplicates kept (like ``rankfm.py:174``).  One radix sort on the
        GPU (``rfm_prep_user_items``) when a device is present and the input is large enough to pay for the copies, else
        one NumPy key sort -- same result either way (``tests/test_gpu_parity.py::test_device_prep_matches_host_prep``)."""
        inter = np.asarray(interactions)
        if len(inter) >= _PREP_DEVICE_MIN and device_count() > 0:
            n_items = int(inter[:, 1].max()) + 1 if n_items is None else int(n_items)
            return cls(*prep_user_items(inter, n_users, n_items))
        return cls.from_interactions_host(inter, n_users)

    @classmethod
    def from_interactions_host(cls, interactions, n_users):
        inter = np.asarray(interactions)
        n_items = int(inter[:, 1].max()) + 1 if len(inter) else 1
        keys = inter[:, 0].astype(np.int64) * n_items + inter[:, 1]        # one key sort instead of a 2-key lexsort
        keys.sort()                                                         # equal keys are identical: stability is moot
        counts = np.bincount(inter[:, 0], minlength=n_users)
        indptr = np.zeros(n_users + 1, dtype=np.int64)
        np.cumsum(counts, out=indptr[1:])
        return cls(indptr, (keys % n_items).astype(np.int32))


_PREP_DEVICE_MIN = int(os.environ.get("RANKFM_B200_PREP_DEVICE_MIN", 200_000))   # below this the host sort is faster than the PCIe round trip


def prep_user_items(interactions, n_users, n_items):
    """``user_items`` as CSR (indptr int64 [U+1], indices int32 [N]) by a device radix sort (``rankfm.py:165-174``)"""
    inter = np.ascontiguousarray(interactions, dtype=np.int32)
    indptr = np.empty(n_users + 1, dtype=np.int64)
    indices = np.empty(len(inter), dtype=np.int32)
    check(_lib.lib().rfm_prep_user_items(ptr(inter), len(inter), int(n_users), int(n_items), _DEVICE, ptr(indptr), ptr(indices)))
    return indptr, indices


def prep_index_ids(ids):
    """(sorted unique ids, int32 index of every id in them) for integer ids, on the device: ``np.unique`` + the Series maps of
    ``rankfm.py:114-128,150-155``"""
    ids64 = np.ascontiguousarray(ids, dtype=np.int64)
    uniq = np.empty(len(ids64), dtype=np.int64)
    index = np.empty(len(ids64), dtype=np.int32)
    n = C.c_int64()
    check(_lib.lib().rfm_prep_index_ids(ptr(ids64), len(ids64), _DEVICE, ptr(uniq), C.byref(n), ptr(index)))
    return uniq[:n.value].copy(), index


def user_items_to_csr(user_items, n_users):
    if hasattr(user_items, "indptr") and hasattr(user_items, "indices"):
        return user_items.indptr, user_items.indices
    lens = np.fromiter((len(user_items[u]) for u in range(n_users)), dtype=np.int64, count=n_users)
    indptr = np.zeros(n_users + 1, dtype=np.int64)
    np.cumsum(lens, out=indptr[1:])
    if n_users:
        indices = np.ascontiguousarray(np.concatenate([np.asarray(user_items[u], dtype=np.int32) for u in range(n_users)]), dtype=np.int32)
    else:
        indices = np.zeros(0, dtype=np.int32)
    return indptr, indices


def _problem(x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if, keep):
    for a, nd, name in ((x_uf, 2, "x_uf"), (x_if, 2, "x_if"), (w_i, 1, "w_i"), (w_if, 1, "w_if"),
                        (v_u, 2, "v_u"), (v_i, 2, "v_i"), (v_uf, 2, "v_uf"), (v_if, 2, "v_if")):
        as_buffer(a, np.float32, nd, name)
    # The library takes U, I, P, Q, F from the weight arrays and reads x_uf as [U,P], x_if as [I,Q]: every shape has to
    # agree before a raw pointer crosses the C ABI.  The reference never reads a feature block whose matrix is all zero
    # (`x_uf_any` / `x_if_any`, _rankfm.pyx:193-194), so `fit(user_features=...)` followed by `fit_partial()` without
    # features -- x_uf back to zeros [U,1] while v_uf stays [P,F] (rankfm.py:199,236) -- is legal there: an all-zero
    # matrix of the wrong width is replaced by zeros of the right one, anything else is an error.
    (U, F), I, P, Q = v_u.shape, v_i.shape[0], v_uf.shape[0], v_if.shape[0]
    if v_i.shape[1] != F or v_uf.shape[1] != F or v_if.shape[1] != F:
        raise ValueError("v_u, v_i, v_uf, v_if must have the same number of factors")
    if w_i.shape[0] != I or w_if.shape[0] != Q:
        raise ValueError("w_i / w_if do not match v_i [%d] / v_if [%d]" % (I, Q))
    if x_uf.shape[0] != U or x_if.shape[0] != I:
        raise ValueError("x_uf / x_if rows do not match v_u [%d] / v_i [%d]" % (U, I))
    if x_uf.shape[1] != P:
        if x_uf.any():
            raise ValueError("x_uf has %d feature columns but v_uf has %d rows" % (x_uf.shape[1], P))
        x_uf = np.zeros((U, P), dtype=np.float32)
    if x_if.shape[1] != Q:
        if x_if.any():
            raise ValueError("x_if has %d feature columns but v_if has %d rows" % (x_if.shape[1], Q))
        x_if = np.zeros((I, Q), dtype=np.float32)
    p = Problem()
    p.x_uf, p.x_if = ptr(x_uf), ptr(x_if)
    p.w_i, p.w_if, p.v_u, p.v_i, p.v_uf, p.v_if = ptr(w_i), ptr(w_if), ptr(v_u), ptr(v_i), ptr(v_uf), ptr(v_if)
    p.U, p.F = v_u.shape
    p.I, p.P, p.Q = v_i.shape[0], v_uf.shape[0], v_if.shape[0]
    p.max_samples = 1
    p.device = _DEVICE
    p.rank, p.world = 0, 1
    keep.extend([x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if])
    return p


def fit_problem(interactions, sample_weight, user_items, x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if,
                alpha, beta, learning_rate, learning_schedule, learning_exponent, max_samples,
                mode=None, seed=None, keep=None, user_range=None, epoch_offset=0):
    """build the ``rfm_problem`` for a training call; ``keep`` collects the arrays whose memory it points into"""
    keep = [] if keep is None else keep
    as_buffer(interactions, np.int32, 2, "interactions")
    as_buffer(sample_weight, np.float32, 1, "sample_weight")
    if learning_schedule not in _lib.SCHEDULE:
        raise ValueError('unknown [learning_schedule]')                       # _rankfm.pyx:225
    p = _problem(x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if, keep)
    indptr, indices = user_items_to_csr(user_items, p.U)
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    keep.extend([interactions, sample_weight, indptr, indices])
    p.interactions, p.sample_weight, p.n_interactions = ptr(interactions), ptr(sample_weight), interactions.shape[0]
    p.csr_indptr, p.csr_indices = ptr(indptr), ptr(indices)
    p.alpha, p.beta, p.learning_rate, p.learning_exponent = alpha, beta, learning_rate, learning_exponent
    p.schedule = _lib.SCHEDULE[learning_schedule]
    p.max_samples = max_samples
    mode = mode or _MODE
    if mode == "replay":
        p.order, p.sampler, p.sched = _lib.ORDER_HOST, _lib.SAMPLER_MT, _lib.SCHED_SERIAL
    else:
        p.order, p.sampler, p.sched = _lib.ORDER_FEISTEL, _lib.SAMPLER_PHILOX, _lib.SCHED_PARALLEL
    p.mt_seed = 1492                                                           # _rankfm.pyx:182
    p.seed = _SEED if seed is None else int(seed)
    p.max_rejects = 0
    p.epoch_offset = int(epoch_offset)
    p.rank, p.world = _COMM["rank"], _COMM["world"]
    if p.world > 1:
        keep.append(_COMM["nccl_id"])
        p.nccl_id = ptr(_COMM["nccl_id"])
        user_range = user_range or _COMM["user_range"]
        if user_range is None:                  # the smallest range covering this shard's users; rows of users without
            users = interactions[:, 0]          # interactions never change, so nothing is lost outside it
            user_range = (int(users.min()), int(users.max()) + 1) if len(users) else (0, 1)
    if user_range is not None:
        p.user_lo, p.user_hi = int(user_range[0]), int(user_range[1])
        if len(interactions) and (interactions[:, 0].min() < p.user_lo or interactions[:, 0].max() >= p.user_hi):
            raise ValueError("interactions of users outside this rank's user range [%d, %d)" % (p.user_lo, p.user_hi))
    return p


def _stats_list(stats, epochs):
    return [dict(log_likelihood=s.log_likelihood, penalty=s.penalty, draws=s.draws, finite=list(s.finite), eta=s.eta,
                 kernel_ms=s.kernel_ms, sync_ms=s.sync_ms) for s in stats[:epochs]]


def fit_ex(interactions, sample_weight, user_items, x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if,
           alpha, beta, learning_rate, learning_schedule, learning_exponent, max_samples, epochs,
           mode=None, perms=None, seed=None, order=None, sampler=None, sched=None, max_rejects=0, user_range=None, epoch_offset=0,
           on_epoch=None):
    """``_fit`` with the execution knobs exposed (tests, bench).  ``perms`` int32 [epochs, N] is required when the
    order is HOST; ``order``/``sampler``/``sched`` override what ``mode`` selects.  Returns the per-epoch stats;
    raises like ``_fit``."""
    keep = []
    p = fit_problem(interactions, sample_weight, user_items, x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if,
                    alpha, beta, learning_rate, learning_schedule, learning_exponent, max_samples, mode=mode, seed=seed, keep=keep,
                    user_range=user_range, epoch_offset=epoch_offset)
    if order is not None:
        p.order = order
    if sampler is not None:
        p.sampler = sampler
    if sched is not None:
        p.sched = sched
    p.max_rejects = max_rejects
    if p.order == _lib.ORDER_HOST:
        perms = np.ascontiguousarray(perms, dtype=np.int32)
        assert perms.shape == (epochs, interactions.shape[0]), "[perms] must be int32 [epochs, N]"
    else:
        perms = None
    if on_epoch is not None:                 # per-epoch reporting needs a session (the one-shot call has no callback argument)
        sess = Session(p, keep)
        try:
            try:
                return sess.train(epochs, perms, on_epoch=on_epoch)
            finally:
                sess.download(w_i, w_if, v_u, v_i, v_uf, v_if)
        finally:
            sess.close()
    stats = (EpochStats * epochs)()
    rc = _lib.lib().rfm_fit(C.byref(p), epochs, ptr(perms), C.cast(stats, C.c_void_p))
    out = _stats_list(stats, epochs)
    check(rc)
    return out


def _print_epoch(epoch, log_likelihood, penalty):
    print("\ntraining epoch:", epoch)                                          # _rankfm.pyx:332-336
    print("log likelihood:", round(float(np.float32(log_likelihood - penalty)), 2))


def _training_session(interactions, sample_weight, user_items, x_uf, x_if, weights, hyper, max_samples):
    """the resident training session for exactly this data / these hyper-parameters, (re)built when anything differs.
    -> (session, fresh): `fresh` sessions were created from `weights` and need no further upload"""
    from_csr = hasattr(user_items, "indptr") and hasattr(user_items, "indices")
    indptr, indices = user_items_to_csr(user_items, weights[2].shape[0])
    key = (_fingerprint(interactions), _fingerprint(sample_weight), _fingerprint(indptr, from_csr), _fingerprint(indices, from_csr),
           _fingerprint(x_uf), _fingerprint(x_if), tuple(w.shape for w in weights), tuple(hyper), int(max_samples),
           _SEED, _DEVICE, _COMM["rank"], _COMM["world"], None if _COMM["nccl_id"] is None else _COMM["nccl_id"].tobytes(), _COMM["user_range"])
    if _training["key"] == key and _training["sess"] is not None:
        _training["hits"] += 1
        return _training["sess"], False
    drop_training()
    keep = []
    user_items = user_items if from_csr else UserItems(indptr, indices)
    p = fit_problem(interactions, sample_weight, user_items, x_uf, x_if, *weights, *hyper, max_samples, mode="production", keep=keep,
                    epoch_offset=_EPOCHS["done"])
    _training.update(key=key, sess=Session(p, keep))
    _training["builds"] += 1
    return _training["sess"], True


def _fit(interactions, sample_weight, user_items, x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if,
         alpha, beta, learning_rate, learning_schedule, learning_exponent, max_samples, epochs, verbose):
    """train in place on the GPU -- same contract as the reference's ``_fit`` (``_rankfm.pyx:122-342``): the six weight
    arrays are updated in place, nothing is returned, ``AssertionError`` if weights go non-finite."""
    global last_stats
    drop_resident()                                                            # the weights are about to change in place
    hyper = (alpha, beta, learning_rate, learning_schedule, learning_exponent)
    weights = (w_i, w_if, v_u, v_i, v_uf, v_if)
    if _MODE == "replay":
        N = interactions.shape[0]
        shuffle_index = np.arange(N, dtype=np.int32)                           # _rankfm.pyx:197
        perms = np.empty((epochs, N), dtype=np.int32)
        for e in range(epochs):
            np.random.shuffle(shuffle_index)                                   # _rankfm.pyx:227 (cumulative)
            perms[e] = shuffle_index
        last_stats = fit_ex(interactions, sample_weight, user_items, x_uf, x_if, *weights, *hyper, max_samples, epochs, perms=perms)
        if verbose:
            for e, st in enumerate(last_stats):
                _print_epoch(e, st["log_likelihood"], st["penalty"])
        return
    # production: the log-likelihood of every epoch is printed as the epoch completes (_rankfm.pyx:332-336)
    on_epoch = (lambda e, st: _print_epoch(e, st["log_likelihood"], st["penalty"])) if verbose else None
    if not _RESIDENT_TRAIN:
        try:
            last_stats = fit_ex(interactions, sample_weight, user_items, x_uf, x_if, *weights, *hyper, max_samples, epochs,
                                epoch_offset=_EPOCHS["done"], on_epoch=on_epoch)
        finally:
            _EPOCHS["done"] += int(epochs)
        return
    as_buffer(interactions, np.int32, 2, "interactions")
    as_buffer(sample_weight, np.float32, 1, "sample_weight")
    if learning_schedule not in _lib.SCHEDULE:
        raise ValueError('unknown [learning_schedule]')                       # _rankfm.pyx:225
    sess, fresh = _training_session(interactions, sample_weight, user_items, x_uf, x_if, weights, hyper, max_samples)
    try:
        if not fresh:
            for a, nd, name in zip(weights, (1, 1, 2, 2, 2, 2), WEIGHT_NAMES):
                as_buffer(a, np.float32, nd, name)
            sess.set_weights(*weights)                                         # warm start from the caller's arrays
        try:
            last_stats = sess.train(epochs, on_epoch=on_epoch)
        finally:
            _EPOCHS["done"] += int(epochs)
            last_stats = sess.last_stats
            sess.download(*weights)          # like the reference, weights are written back even when they went non-finite
    except Exception:
        drop_training()
        raise


def _predict(pairs, x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if):
    """scores of (user_idx, item_idx) pairs given as float32, NaN = unknown id (``_rankfm.pyx:345-390``)"""
    as_buffer(pairs, np.float32, 2, "pairs")
    if _RESIDENT:
        return _resident_session((x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if)).predict(pairs)
    keep = []
    p = _problem(x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if, keep)
    scores = np.empty(pairs.shape[0], dtype=np.float32)
    check(_lib.lib().rfm_predict(C.byref(p), ptr(pairs), pairs.shape[0], ptr(scores)))
    return scores


def _recommend(users, user_items, n_items, filter_previous, x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if):
    """top-``n_items`` item indexes (as float32) per user (``_rankfm.pyx:393-460``)"""
    as_buffer(users, np.float32, 1, "users")
    if _RESIDENT:
        sess = _resident_session((x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if), user_items if filter_previous else None)
        return sess.recommend(users, int(n_items), bool(filter_previous))
    keep = []
    p = _problem(x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if, keep)
    if filter_previous:
        indptr, indices = user_items_to_csr(user_items, p.U)
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        keep.extend([indptr, indices])
        p.csr_indptr, p.csr_indices = ptr(indptr), ptr(indices)
    rec_items = np.empty((users.shape[0], n_items), dtype=np.float32)
    check(_lib.lib().rfm_recommend(C.byref(p), ptr(users), users.shape[0], int(n_items), int(bool(filter_previous)), ptr(rec_items)))
    return rec_items


def _similar(which, index, n, x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if):
    """top-n most similar rows by latent inner product (``rankfm.py:405-454``); which=0 items, 1 users"""
    if _RESIDENT:
        return _resident_session((x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if)).similar(which, index, n)
    keep = []
    p = _problem(x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if, keep)
    out = np.empty(n, dtype=np.int32)
    check(_lib.lib().rfm_similar(C.byref(p), int(which), int(index), int(n), ptr(out)))
    return out


def _scoring_session(weights, user_items=None):
    """a scoring session for these weights: the resident one when that switch is on, else a throw-away one
    -> (session, close_after_use)"""
    if _RESIDENT:
        return _resident_session(weights, user_items), False
    keep = []
    sess = Session(_problem(*weights, keep), keep)
    if user_items is not None:
        indptr, indices = user_items_to_csr(user_items, weights[4].shape[0])
        sess.attach_csr(np.ascontiguousarray(indptr, dtype=np.int64), np.ascontiguousarray(indices, dtype=np.int32))
    return sess, True


def _evaluate(users, test_indptr, test_items, n_test, k, filter_previous, user_items, x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if, want_hits=False):
    """hold-out metrics of ``rankfm/evaluation.py:9-143`` from the device top-k (``rfm_session_evaluate``)"""
    sess, close = _scoring_session((x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if), user_items if filter_previous else None)
    try:
        return sess.evaluate(users, test_indptr, test_items, n_test, k, filter_previous, want_hits)
    finally:
        if close:
            sess.close()


def _similar_batch(which, indexes, n, x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if):
    """``_similar`` for many query rows at once -> int32 [len(indexes), n] (-1 = no such row)"""
    sess, close = _scoring_session((x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if))
    try:
        return sess.similar_batch(which, indexes, n)
    finally:
        if close:
            sess.close()


def shard_by_user(interactions, sample_weight, n_users, rank, world):
    """user-range partition with ~equal interaction counts: rank r gets the rows of users in [b_r, b_{r+1})"""
    counts = np.bincount(interactions[:, 0], minlength=n_users)
    cum = np.cumsum(counts)
    bounds = np.searchsorted(cum, cum[-1] * np.arange(1, world) / world, side="left")
    bounds = np.concatenate([[0], bounds + 1, [n_users]])
    lo, hi = bounds[rank], bounds[rank + 1]
    sel = (interactions[:, 0] >= lo) & (interactions[:, 0] < hi)
    return np.ascontiguousarray(interactions[sel]), np.ascontiguousarray(sample_weight[sel]), (int(lo), int(hi))


def allgather_user_rows(v_u, user_range):
    """multi-process jobs (torch.distributed initialised by the caller, any backend): after a sharded ``_fit`` every
    rank holds the trained rows of ITS users only; this assembles the full user table in every rank's ``v_u`` in place.
    Control plane, outside the hot path: one broadcast per rank of the row range it owns."""
    import torch
    import torch.distributed as dist
    ranges = [None] * dist.get_world_size()
    dist.all_gather_object(ranges, (int(user_range[0]), int(user_range[1])))
    for src, (lo, hi) in enumerate(ranges):
        if hi > lo:
            dist.broadcast(torch.from_numpy(v_u[lo:hi]), src=src)
    return v_u


class Session:
    """resident-HBM session used by ``bench.py`` and the ``RankFM`` class: upload once, train/score many times"""

    def __init__(self, problem, keep):
        self._keep = keep
        self._p = problem
        self.last_stats = None
        h = C.c_void_p()
        check(_lib.lib().rfm_session_create(C.byref(problem), C.byref(h)))
        self._h = h

    def train(self, epochs, perms=None, on_epoch=None):
        """`on_epoch(epoch, stats_dict)` is called after every epoch (one stream synchronisation each)"""
        stats = (EpochStats * epochs)()
        cb = None
        if on_epoch is not None:
            cb = _lib.EPOCH_CALLBACK(lambda e, st, _user: on_epoch(int(e), _stats_list([st.contents], 1)[0]))
            check(_lib.lib().rfm_session_set_epoch_callback(self._h, cb, None))
        try:
            rc = _lib.lib().rfm_session_train(self._h, epochs, ptr(perms), C.cast(stats, C.c_void_p))
        finally:
            if cb is not None:
                _lib.lib().rfm_session_set_epoch_callback(self._h, _lib.EPOCH_CALLBACK(), None)
        self.last_stats = _stats_list(stats, epochs)
        check(rc)
        return self.last_stats

    def exchange_path(self):
        """multi-GPU: 0 = single GPU, 1 = fused peer-memory kernel, 2 = ncclAllReduce fallback"""
        path = C.c_int32()
        check(_lib.lib().rfm_session_exchange_path(self._h, C.byref(path)))
        return path.value

    def set_weights(self, w_i, w_if, v_u, v_i, v_uf, v_if):
        check(_lib.lib().rfm_session_set_weights(self._h, *[ptr(a) for a in (w_i, w_if, v_u, v_i, v_uf, v_if)]))

    def download(self, w_i, w_if, v_u, v_i, v_uf, v_if):
        check(_lib.lib().rfm_session_download(self._h, *[ptr(a) for a in (w_i, w_if, v_u, v_i, v_uf, v_if)]))

    def snapshot(self):
        check(_lib.lib().rfm_session_snapshot(self._h))

    def restore(self):
        check(_lib.lib().rfm_session_restore(self._h))

    def timer_start(self):
        check(_lib.lib().rfm_session_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        check(_lib.lib().rfm_session_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def predict(self, pairs):
        scores = np.empty(pairs.shape[0], dtype=np.float32)
        check(_lib.lib().rfm_session_predict(self._h, ptr(pairs), pairs.shape[0], ptr(scores)))
        return scores

    def recommend(self, users, n_items, filter_previous=False):
        rec = np.empty((users.shape[0], n_items), dtype=np.float32)
        check(_lib.lib().rfm_session_recommend(self._h, ptr(users), users.shape[0], n_items, int(filter_previous), ptr(rec)))
        return rec

    def time_predict(self, pairs, iters=10):
        ms = C.c_float()
        check(_lib.lib().rfm_session_time_predict(self._h, ptr(pairs), pairs.shape[0], iters, C.byref(ms)))
        return ms.value

    def time_recommend(self, users, n_items, filter_previous=False, iters=3):
        ms, gemm = C.c_float(), C.c_float()
        check(_lib.lib().rfm_session_time_recommend(self._h, ptr(users), users.shape[0], n_items, int(filter_previous), iters, C.byref(ms), C.byref(gemm)))
        return ms.value, gemm.value

    def trace_enable(self):
        check(_lib.lib().rfm_session_trace_enable(self._h))

    def trace_read(self):
        out = np.empty((self._p.n_interactions, 2), dtype=np.int32)
        check(_lib.lib().rfm_session_trace_read(self._h, ptr(out)))
        return out

    def debug_gemm(self, users):
        out = np.empty((users.shape[0], self._p.I), dtype=np.float32)
        check(_lib.lib().rfm_session_debug_gemm(self._h, ptr(users), users.shape[0], ptr(out)))
        return out

    def similar(self, which, index, n):
        out = np.empty(n, dtype=np.int32)
        check(_lib.lib().rfm_session_similar(self._h, int(which), int(index), int(n), ptr(out)))
        return out

    def similar_batch(self, which, indexes, n):
        indexes = np.ascontiguousarray(indexes, dtype=np.int32)
        out = np.empty((len(indexes), n), dtype=np.int32)
        check(_lib.lib().rfm_session_similar_batch(self._h, int(which), ptr(indexes), len(indexes), int(n), ptr(out)))
        return out

    def evaluate(self, users, test_indptr, test_items, n_test, k, filter_previous=False, want_hits=False):
        """the five hold-out metrics (hit rate, reciprocal rank, DCG, precision, recall) from the device top-k"""
        users = np.ascontiguousarray(users, dtype=np.float32)
        test_indptr = np.ascontiguousarray(test_indptr, dtype=np.int64)
        test_items = np.ascontiguousarray(test_items, dtype=np.int32)
        n_test = np.ascontiguousarray(n_test, dtype=np.int32)
        out = np.zeros(5, dtype=np.float64)
        hits = np.empty((len(users), k), dtype=np.uint8) if want_hits else None
        check(_lib.lib().rfm_session_evaluate(self._h, ptr(users), len(users), int(k), int(bool(filter_previous)), ptr(test_indptr), ptr(test_items),
                                              ptr(n_test), ptr(out), ptr(hits)))
        return out, hits

    def attach_csr(self, indptr, indices):
        check(_lib.lib().rfm_session_attach_csr(self._h, ptr(indptr), ptr(indices)))

    def recommend_stats(self):
        """(rows served by the tensor-core recommend path, rows of those redone on the exact path)"""
        rows, redo = C.c_int64(), C.c_int64()
        check(_lib.lib().rfm_session_recommend_stats(self._h, C.byref(rows), C.byref(redo)))
        return rows.value, redo.value

    def recommend_retried(self):
        """rows served a second time with the provable row threshold (the estimated one came out too high for them)"""
        n = C.c_int64()
        check(_lib.lib().rfm_session_recommend_retried(self._h, C.byref(n)))
        return n.value

    def flush_l2(self):
        check(_lib.lib().rfm_session_flush_l2(self._h))

    def launch_count(self):
        n = C.c_int64()
        check(_lib.lib().rfm_session_launch_count(self._h, C.byref(n)))
        return n.value

    def close(self):
        if self._h:
            _lib.lib().rfm_session_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
