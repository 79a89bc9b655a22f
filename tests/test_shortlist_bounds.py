"""The candidate-selection arithmetic of the tensor-core recommend path, restated in NumPy float32
(``oracle/shortlist_bounds.py``), against its defining property on the CPU: whatever part of the catalogue pass 1 looks at,
every item whose (bf16-GEMM) score reaches the row's n'-th best score passes the pass-2 gate -- for random, tied, clustered,
huge- and tiny-magnitude inputs.  The GPU parity tests (test_gpu_parity.py) check the kernels' end result against the exact
path; this pins the rounding analysis behind the slack term and the conservative radix bucket."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import shortlist_bounds as sb


def _holds(dot, bias, want, fraction, head):
    dot, bias = np.asarray(dot, np.float32), np.asarray(bias, np.float32)
    cand, tau = sb.candidates(dot, bias, want, fraction, head)
    score = (dot + bias).astype(np.float32)                       # what shortlist_kernel ranks the candidates by
    if want > len(score):
        return True
    s_want = np.sort(score)[len(score) - want]
    must = np.flatnonzero(score >= s_want)
    assert tau <= s_want, (tau, s_want)
    missing = np.setdiff1d(must, cand)
    assert len(missing) == 0, (missing[:5], score[missing[:5]], tau, s_want)
    return True


@pytest.mark.parametrize("fraction,head", [(1, True), (2, True), (4, True), (8, True), (4, False)])
@pytest.mark.parametrize("bias_scale,dot_scale", [(0.3, 0.1), (0.0, 0.1), (0.5, 0.5), (1e4, 1e-3), (1e-6, 1e-6), (3.0, 1e3)])
def test_gate_keeps_the_best_items_random(fraction, head, bias_scale, dot_scale):
    rng = np.random.default_rng(int(1000 * bias_scale) + fraction)
    for n_items, want in ((20000, 216), (5000, 36), (1000, 56), (130, 20)):
        for _ in range(3):
            dot = (rng.standard_normal(n_items) * dot_scale).astype(np.float32)
            bias = (rng.standard_normal(n_items) * bias_scale).astype(np.float32)
            _holds(dot, bias, want, fraction, head)


def test_gate_with_ties_and_degenerate_catalogues():
    rng = np.random.default_rng(0)
    n = 4096
    _holds(np.full(n, 0.37, np.float32), np.full(n, 0.25, np.float32), 36, 4, True)              # everything tied
    _holds(np.zeros(n, np.float32), np.zeros(n, np.float32), 36, 1, True)
    dot = rng.standard_normal(n).astype(np.float32); dot[100:400] = dot[100]                       # 300 duplicates of one item
    bias = rng.standard_normal(n).astype(np.float32); bias[100:400] = bias[100]
    _holds(dot, bias, 56, 2, True)
    bias = np.sort(rng.standard_normal(n).astype(np.float32))                                      # biases already sorted ascending
    _holds(rng.standard_normal(n).astype(np.float32), bias, 216, 8, True)
    _holds(-np.abs(rng.standard_normal(n)).astype(np.float32) * 1e5, bias, 100, 4, True)          # all scores negative, large
    _holds(rng.standard_normal(200).astype(np.float32), rng.standard_normal(200).astype(np.float32), 216, 4, True)   # want > catalogue


def test_fewer_bounds_than_wanted_opens_the_gate():
    """-inf threshold: every position passes, the kernel's slots overflow and the row goes to the exact path"""
    rng = np.random.default_rng(1)
    dot, bias = rng.standard_normal(1000).astype(np.float32), rng.standard_normal(1000).astype(np.float32)
    cand, tau = sb.candidates(dot, bias, 200, fraction=8, head=True)       # 8 tiles -> 1 visited -> 16 bounds < 200
    assert tau == -np.inf and len(cand) == 1000


@settings(max_examples=150, deadline=None)
@given(st.integers(8, 600), st.integers(1, 64), st.sampled_from([1, 2, 4, 8]), st.booleans(), st.integers(0, 2 ** 32 - 1),
       st.sampled_from([1e-30, 1e-3, 1.0, 1e6, 1e20]), st.sampled_from([0.0, 1e-3, 1.0, 1e6, 1e20]))
def test_gate_property(n_items, want, fraction, head, seed, dot_scale, bias_scale):
    rng = np.random.default_rng(seed)
    dot = (rng.standard_normal(n_items) * dot_scale).astype(np.float32)
    bias = (rng.standard_normal(n_items) * bias_scale).astype(np.float32)
    if seed % 3 == 0:                                                       # quantise: many exact ties
        dot = np.round(dot / np.float32(max(dot_scale, 1e-30)) * 2).astype(np.float32) * np.float32(dot_scale) / 2
        bias = np.round(bias * 2).astype(np.float32) / 2
    _holds(dot, bias, want, fraction, head)


def test_padding_sentinel_bounds_the_score_range():
    """padded positions carry bias -1e30: the path assumes real scores stay above that (weights that large are non-finite
    within an epoch and `_fit` raises long before); below it a pad would out-rank real items in pass 1"""
    dot = np.zeros(8, np.float32)
    bias = np.full(8, -2e30, np.float32)
    cand, tau = sb.candidates(dot, bias, 6)
    assert -1.001e30 < tau <= sb.PAD_BIAS and len(cand) == 0          # the pads set the threshold, every real item is lost


def test_key_order_and_bucket_edge():
    x = np.array([-np.inf, -3e38, -1.0, -1e-30, -0.0, 0.0, 1e-30, 1.0, 3e38, np.inf], np.float32)
    k = sb.ord_key(x)
    assert np.all(np.diff(k.astype(np.int64)) >= 0)
    assert np.array_equal(sb.key_to_float(k).view(np.uint32), x.view(np.uint32))
    for v in (1.2345678, -1.2345678, 1e-20, -7e20):
        edge = sb.key_to_float(sb.ord_key(np.float32(v)) & np.uint32(0xffffff00))[()]
        assert edge <= np.float32(v) and abs(edge - np.float32(v)) <= abs(np.float32(v)) * 2.0 ** -15
