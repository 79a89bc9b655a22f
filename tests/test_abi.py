"""CPU: the C-ABI library loads, exports every symbol include/rankfm_b200.h declares, and refuses to compute
without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from rankfm_b200 import _lib, _rankfm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rankfm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rfm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 20
    for name in syms:
        assert hasattr(lib, name), "librankfm_b200.so does not export %s" % name
    assert set(_lib.EXPORTS) == set(syms)


def test_version_and_struct_layout(lib):
    assert b"sm_100a" in lib.rfm_version()
    # the ctypes mirror must match the C struct: spot-check via sizeof of a C compile
    src = '#include "rankfm_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu", sizeof(rfm_problem), sizeof(rfm_epoch_stats));}'
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")], check=True)
        a, b = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()
    assert int(a) == C.sizeof(_lib.Problem) and int(b) == C.sizeof(_lib.EpochStats)


def test_no_silent_cpu_fallback(lib):
    if lib.rfm_device_count() > 0:
        pytest.skip("a GPU is visible")
    w = dict(w_i=np.zeros(3, np.float32), w_if=np.zeros(1, np.float32), v_u=np.zeros((2, 4), np.float32),
             v_i=np.zeros((3, 4), np.float32), v_uf=np.zeros((1, 4), np.float32), v_if=np.zeros((1, 4), np.float32))
    x_uf, x_if = np.zeros((2, 1), np.float32), np.zeros((3, 1), np.float32)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        _rankfm._predict(np.zeros((1, 2), np.float32), x_uf, x_if, *w.values())
    with pytest.raises(RuntimeError, match="no CUDA device"):
        _rankfm._fit(np.zeros((1, 2), np.int32), np.ones(1, np.float32), {0: np.array([0], np.int32), 1: np.array([], np.int32)},
                     x_uf, x_if, *w.values(), 0.01, 0.1, 0.1, 'constant', 0.25, 1, 1, False)


def test_buffer_validation_mirrors_cython():
    w = dict(w_i=np.zeros(3, np.float32), w_if=np.zeros(1, np.float32), v_u=np.zeros((2, 4), np.float32),
             v_i=np.zeros((3, 4), np.float32), v_uf=np.zeros((1, 4), np.float32), v_if=np.zeros((1, 4), np.float32))
    x_uf, x_if = np.zeros((2, 1), np.float32), np.zeros((3, 1), np.float32)
    with pytest.raises(ValueError):          # float64 pairs: Cython raises "Buffer dtype mismatch"
        _rankfm._predict(np.zeros((1, 2), np.float64), x_uf, x_if, *w.values())
    with pytest.raises(ValueError, match="learning_schedule"):
        _rankfm._fit(np.zeros((1, 2), np.int32), np.ones(1, np.float32), {0: np.array([0], np.int32), 1: np.array([], np.int32)},
                     x_uf, x_if, *w.values(), 0.01, 0.1, 0.1, 'bogus', 0.25, 1, 1, False)


def test_rng_contract_with_oracle(lib):
    from oracle import oracle
    rng = np.random.default_rng(0)
    for _ in range(50):
        c = rng.integers(0, 2**32, 4, dtype=np.uint64); k = rng.integers(0, 2**32, 2, dtype=np.uint64)
        out = np.zeros(4, np.uint32)
        assert lib.rfm_debug_philox(*[int(x) for x in c], *[int(x) for x in k], _lib.ptr(out)) == 0
        assert np.array_equal(out, oracle.philox4x32(c, k))
    for n, seed, epoch in ((1, 5, 0), (7, 5, 1), (1000, 2**40 + 17, 3), (65537, 99, 12)):
        out = np.zeros(n, np.int64)
        assert lib.rfm_debug_feistel(n, seed, epoch, 0, n, _lib.ptr(out)) == 0
        assert np.array_equal(out, oracle.feistel_perm(n, seed, epoch))


def test_estimated_threshold_rank_rule(lib):
    """`tau_rank` (rfm_kernels.h): the m-th largest of a 1-in-k sample estimates the n'-th best of the whole catalogue;
    m solves k m - z k sqrt(m) = n' (negative-binomial head room), never exceeds n', and degenerates to n' without a sample"""
    rank = lib.rfm_debug_tau_rank
    for want in (20, 56, 216, 1024):
        assert rank(want, 1, 4.5) == want and rank(want, 8, 0.0) == want
        prev = want
        for k in (2, 4, 8, 16):
            m = rank(want, k, 4.5)
            assert 1 <= m <= prev                                   # a sparser sample looks at a smaller rank
            assert k * m >= want                                    # expected number of catalogue items above the threshold
            if m < want:                                            # not clamped: z sigma (k sqrt(m)) below the expectation still reaches n'
                assert k * m - 4.5 * k * np.sqrt(m) >= want - k
            assert rank(want, k, 0.001) <= m <= rank(want, k, 9.0)
            prev = m
    assert rank(216, 8, 4.5) == 64 and rank(216, 16, 4.5) == 45     # the values DESIGN.md / the simulation quote


def test_user_items_csr_view():
    X = np.array([[0, 3], [2, 1], [0, 1], [2, 5], [2, 1]], np.int32)
    ui = _rankfm.UserItems.from_interactions(X, 3)
    assert ui[0].tolist() == [1, 3] and ui[1].tolist() == [] and ui[2].tolist() == [1, 1, 5]
    assert len(ui) == 3 and list(ui.keys()) == [0, 1, 2] and 2 in ui and 3 not in ui
    assert [(u, v.tolist()) for u, v in ui.items()] == [(0, [1, 3]), (1, []), (2, [1, 1, 5])]
    ptr, idx = _rankfm.user_items_to_csr({0: np.array([1, 3]), 1: np.array([], np.int32), 2: np.array([1, 1, 5])}, 3)
    assert ptr.tolist() == ui.indptr.tolist() and idx.tolist() == ui.indices.tolist()
    import pickle
    ui2 = pickle.loads(pickle.dumps(ui))
    assert ui2[2].tolist() == [1, 1, 5]


def test_shard_by_user_partitions_everything():
    rng = np.random.default_rng(1)
    X = np.stack([rng.integers(0, 50, 1000), rng.integers(0, 20, 1000)], 1).astype(np.int32)
    sw = rng.uniform(size=1000).astype(np.float32)
    seen = 0
    for r in range(4):
        xs, ws, (lo, hi) = _rankfm.shard_by_user(X, sw, 50, r, 4)
        assert ((xs[:, 0] >= lo) & (xs[:, 0] < hi)).all() and len(xs) == len(ws)
        assert 150 < len(xs) < 350
        seen += len(xs)
    assert seen == 1000


def test_problem_validates_shapes_before_any_pointer_crosses_the_abi():
    """x_uf / x_if are read as [U,P] / [I,Q] with P, Q taken from v_uf / v_if -- mismatching widths must never
    reach the library.  All-zero matrices of another width are what `fit(features)` then `fit_partial()` without features
    produces (rankfm.py:199,236; the reference never reads them: `x_uf_any`, _rankfm.pyx:193): they are replaced by zeros
    of the right width; anything else is an error."""
    import numpy as np
    import pytest
    from rankfm_b200 import _rankfm
    U, I, F, P, Q = 5, 4, 3, 2, 3
    f32 = np.float32
    w = dict(w_i=np.zeros(I, f32), w_if=np.zeros(Q, f32), v_u=np.zeros((U, F), f32), v_i=np.zeros((I, F), f32), v_uf=np.zeros((P, F), f32), v_if=np.zeros((Q, F), f32))
    order = ('w_i', 'w_if', 'v_u', 'v_i', 'v_uf', 'v_if')
    keep = []
    p = _rankfm._problem(np.zeros((U, 1), f32), np.zeros((I, 1), f32), *[w[k] for k in order], keep)
    assert (p.U, p.I, p.P, p.Q, p.F) == (U, I, P, Q, F)
    assert keep[0].shape == (U, P) and keep[1].shape == (I, Q) and not keep[0].any() and not keep[1].any()
    with pytest.raises(ValueError):
        _rankfm._problem(np.ones((U, 1), f32), np.zeros((I, Q), f32), *[w[k] for k in order], [])
    with pytest.raises(ValueError):
        _rankfm._problem(np.zeros((U, P), f32), np.ones((I, Q + 1), f32), *[w[k] for k in order], [])
    with pytest.raises(ValueError):
        _rankfm._problem(np.zeros((U + 1, P), f32), np.zeros((I, Q), f32), *[w[k] for k in order], [])
    bad = dict(w, w_i=np.zeros(I + 1, f32))
    with pytest.raises(ValueError):
        _rankfm._problem(np.zeros((U, P), f32), np.zeros((I, Q), f32), *[bad[k] for k in order], [])
    bad = dict(w, v_if=np.zeros((Q, F + 1), f32))
    with pytest.raises(ValueError):
        _rankfm._problem(np.zeros((U, P), f32), np.zeros((I, Q), f32), *[bad[k] for k in order], [])
