"""CPU: pin the oracle (oracle/rankfm_oracle.c) against the reference's known answers and golden vectors."""
import numpy as np
import pytest

from helpers import CSRItems, KERNEL_CASES, WEIGHTS, golden_fit_args, load_golden
from oracle import oracle

# init_genrand(1492); genrand_int32() x6 -- verified against the reference's mt19937ar.c
MT_KAT_1492 = [1679283159, 3061641750, 3575273037, 870970970, 501062895, 3875167081]


def test_mt19937_known_answer():
    assert oracle.mt_stream(1492, 6).tolist() == MT_KAT_1492


def test_mt19937_matches_numpy_legacy_seeding():
    bg = np.random.MT19937()
    bg._legacy_seeding(1492)
    assert np.array_equal(oracle.mt_stream(1492, 5000), bg.random_raw(5000).astype(np.uint32))


@pytest.mark.parametrize("case", KERNEL_CASES)
def test_fit_matches_reference_golden(case):
    g = load_golden(case)
    args, w, _ = golden_fit_args(g)
    oracle.fit_ex(*args, g['epochs'], perms=g['perms'], sampler="mt", mt_seed=1492)
    for k in WEIGHTS:
        # the reference is built with -ffast-math (setup.py:25): last-ulp wobble only
        np.testing.assert_allclose(w[k], g[k + '_ref'], rtol=2e-5, atol=2e-6, err_msg=k)


@pytest.mark.parametrize("case", KERNEL_CASES)
def test_predict_matches_reference_golden(case):
    g = load_golden(case)
    w = [g[k + '_ref'] for k in WEIGHTS]
    scores = oracle._predict(np.ascontiguousarray(g['pairs']), g['x_uf'], g['x_if'], *w)
    assert np.array_equal(np.isnan(scores), np.isnan(g['scores']))
    np.testing.assert_allclose(scores, g['scores'], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("case", KERNEL_CASES)
@pytest.mark.parametrize("filt", [False, True])
def test_recommend_matches_reference_golden(case, filt):
    g = load_golden(case)
    _, _, ui = golden_fit_args(g)
    w = [g[k + '_ref'] for k in WEIGHTS]
    rec = oracle._recommend(g['users'], ui, 10, filt, g['x_uf'], g['x_if'], *w)
    want = g['rec_filtered'] if filt else g['rec']
    assert np.array_equal(rec, want, equal_nan=True)


def test_oracle_against_compiled_reference_when_present():
    ref = oracle.load_reference()
    if ref is None:
        pytest.skip("oracle/_ref not built on this machine")
    g = load_golden('warp_feat')
    args, w, _ = golden_fit_args(g)
    args2, w2, _ = golden_fit_args(g)
    np.random.seed(5)
    args = args[:2] + (dict(args[2]),) + args[3:]          # the Cython signature wants an exact dict
    ref._fit(*args, 2, False)
    np.random.seed(5)
    oracle._fit(*args2, 2, False)
    for k in WEIGHTS:
        np.testing.assert_allclose(w2[k], w[k], rtol=2e-5, atol=2e-6, err_msg=k)


def test_feistel_is_a_permutation():
    for n in (1, 2, 3, 17, 1000, 4097):
        p = oracle.feistel_perm(n, seed=1234, epoch=2)
        assert sorted(p.tolist()) == list(range(n))
    assert not np.array_equal(oracle.feistel_perm(1000, 1234, 0), oracle.feistel_perm(1000, 1234, 1))


def test_philox_sampler_is_deterministic_and_order_free():
    g = load_golden('bpr_f16')
    args, w1, _ = golden_fit_args(g)
    out1 = oracle.fit_ex(*args, 2, perms=None, sampler="philox", seed=77, want_neg=True)
    args, w2, _ = golden_fit_args(g)
    out2 = oracle.fit_ex(*args, 2, perms=None, sampler="philox", seed=77, want_neg=True)
    assert np.array_equal(out1['neg'], out2['neg']) and np.array_equal(w1['v_i'], w2['v_i'])
    # BPR: the negative of a row depends only on (row, epoch, seed), not on the visiting order
    N = len(g['interactions'])
    args, _, _ = golden_fit_args(g)
    out3 = oracle.fit_ex(*args, 2, perms=g['perms'][:2], sampler="philox", seed=77, want_neg=True)
    for e in range(2):
        by_row_1 = np.empty(N, np.int64); by_row_1[oracle.feistel_perm(N, 77, e)] = out1['neg'][e]
        by_row_3 = np.empty(N, np.int64); by_row_3[g['perms'][e]] = out3['neg'][e]
        assert np.array_equal(by_row_1, by_row_3)


def test_oracle_replay_at_cfg1_named_size_matches_reference_golden():
    """BASELINE.json configs[0] (10k x 5k, 100k interactions, factors=16, bpr, 5 epochs) at its NAMED size: the oracle's
    sequential replay of the reference's row order + MT19937 stream against weights / predict() / recommend() minted from
    the unmodified reference (tests/golden/make_golden.py::cfg1_case)"""
    from helpers import cfg1_case, rel_err, topk_overlap
    p, g = cfg1_case()
    w = {k: v.copy() for k, v in p['w'].items()}
    ui = CSRItems(p['indptr'], p['indices'])
    oracle.fit_ex(p['X'], p['sw'], ui, p['x_uf'], p['x_if'], *[w[k] for k in WEIGHTS], *p['hyper'], 1, p['epochs'], perms=p['perms'])
    for k in ('w_i', 'v_u', 'v_i'):
        assert rel_err(w[k], g[k + '_ref']) < 1e-4, k          # fast-math vs strict IEEE over 437k sequential steps
    scores = oracle._predict(p['pairs'], p['x_uf'], p['x_if'], *[w[k] for k in WEIGHTS])
    assert rel_err(scores, g['scores']) < 1e-4
    rec = oracle._recommend(p['users'], ui, 10, True, p['x_uf'], p['x_if'], *[w[k] for k in WEIGHTS])
    assert topk_overlap(rec, g['rec_filtered']) >= 0.99
