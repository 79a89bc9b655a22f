"""GPU parity tests: the CUDA hot path (through the C ABI / ctypes shim) against the oracle and the golden vectors
minted from the unmodified reference.  Tolerances are written next to each comparison.

  * replay schedule (serial, MT19937 negatives, the reference's row order)  -> same trajectory as the reference:
    trained weights and predict() scores within 1e-4 relative (BASELINE.json north_star tolerance)
  * serial Philox/Feistel schedule -> same trajectory as the oracle run with the same Philox/Feistel contract
  * parallel (Hogwild) schedule    -> statistical parity (log-likelihood, ranking quality, draws)
"""
import numpy as np
import pytest

from helpers import (KERNEL_CASES, WEIGHTS, CSRItems, assert_topn_exact_up_to_rounding, cfg1_case, csr_of, features, golden_fit_args,
                     init_weights, load_golden, rel_err, topk_overlap, zipf_interactions)
from oracle import oracle
from rankfm_b200 import _lib, _rankfm

pytestmark = pytest.mark.gpu

PREDICT_RTOL = 1e-4      # north_star: predict() within 1e-4 relative of the reference for identical seeds/epochs


def _weights(g, which):
    return [np.ascontiguousarray(g[k + '_' + which]) for k in WEIGHTS]


# ---------------------------------------------------------------------------------------------------------------
# training: exact-trajectory modes
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", KERNEL_CASES)
def test_replay_fit_matches_reference_golden(gpu_lib, case):
    g = load_golden(case)
    args, w, _ = golden_fit_args(g)
    stats = _rankfm.fit_ex(*args, g['epochs'], mode="replay", perms=g['perms'])
    for k in WEIGHTS:
        assert rel_err(w[k], g[k + '_ref']) < 1e-4, k
    # the trained model scores like the reference's
    scores = _rankfm._predict(np.ascontiguousarray(g['pairs']), g['x_uf'], g['x_if'], *[w[k] for k in WEIGHTS])
    assert np.array_equal(np.isnan(scores), np.isnan(g['scores']))
    assert rel_err(scores, g['scores']) < PREDICT_RTOL
    # and consumed the reference's number of MT19937 draws (same `sampled` per positive)
    args_o, _, _ = golden_fit_args(g)
    out = oracle.fit_ex(*args_o, g['epochs'], perms=g['perms'], sampler="mt")
    assert [s['draws'] for s in stats] == out['draws'].tolist()
    np.testing.assert_allclose([s['log_likelihood'] for s in stats], out['ll'], rtol=2e-4)


@pytest.mark.parametrize("case", KERNEL_CASES)
def test_serial_philox_fit_matches_oracle(gpu_lib, case):
    g = load_golden(case)
    args, w, _ = golden_fit_args(g)
    stats = _rankfm.fit_ex(*args, g['epochs'], seed=4242, order=_lib.ORDER_FEISTEL, sampler=_lib.SAMPLER_PHILOX,
                           sched=_lib.SCHED_SERIAL, max_rejects=64)
    args_o, wo, _ = golden_fit_args(g)
    out = oracle.fit_ex(*args_o, g['epochs'], perms=None, sampler="philox", seed=4242, max_rejects=64)
    for k in WEIGHTS:
        assert rel_err(w[k], wo[k]) < 1e-4, k
    assert [s['draws'] for s in stats] == out['draws'].tolist()


@pytest.mark.parametrize("bloom", ["1", "0"])
@pytest.mark.parametrize("case", ["warp_f20", "warp_feat", "warp_if_only", "bpr_f16"])
def test_serial_philox_without_bitmap_matches_oracle(gpu_lib, case, bloom, monkeypatch):
    """membership through the CSR (register-resident short lists, (G+1)-ary search for long ones) instead of the bitmap;
    bloom = "1": the one-word-per-entry filter dismisses most candidates before the search (large catalogues' default) --
    draw for draw the same negatives either way, the filter has no false negatives and its positives are verified"""
    monkeypatch.setenv("RANKFM_B200_BITMAP_MB", "0")
    monkeypatch.setenv("RANKFM_B200_BLOOM", bloom)
    g = load_golden(case)
    deg = np.diff(g['indptr'])
    assert deg.max() > 32 and deg.min() <= 16            # both the listed and the searched path are exercised
    args, w, _ = golden_fit_args(g)
    stats = _rankfm.fit_ex(*args, g['epochs'], seed=99, order=_lib.ORDER_FEISTEL, sampler=_lib.SAMPLER_PHILOX, sched=_lib.SCHED_SERIAL, max_rejects=64)
    args_o, wo, _ = golden_fit_args(g)
    out = oracle.fit_ex(*args_o, g['epochs'], perms=None, sampler="philox", seed=99, max_rejects=64)
    for k in WEIGHTS:
        assert rel_err(w[k], wo[k]) < 1e-4, k
    assert [s['draws'] for s in stats] == out['draws'].tolist()
    # and the Hogwild schedule runs through the same code
    args_p, wp, _ = golden_fit_args(g)
    sp = _rankfm.fit_ex(*args_p, g['epochs'], mode="production", seed=99)
    np.testing.assert_allclose([s['draws'] for s in sp], out['draws'], rtol=0.1)
    assert all(all(s['finite']) for s in sp)


@pytest.mark.parametrize("case,sampler", [("warp_f20", "mt"), ("warp_feat", "mt"), ("warp_f20", "philox"), ("bpr_f16", "philox")])
def test_sampler_is_draw_for_draw_identical_to_the_oracle(gpu_lib, case, sampler):
    """per position of the epoch: same negative item, same number of draws as the oracle (first epoch, where both
    start from identical weights; later epochs are covered by the weight comparisons above)"""
    g = load_golden(case)
    args, w, _ = golden_fit_args(g)
    keep = []
    mt = sampler == "mt"
    prob = _rankfm.fit_problem(*args, mode="replay" if mt else "production", seed=31337, keep=keep)
    if not mt:
        prob.sched = _lib.SCHED_SERIAL
    sess = _rankfm.Session(prob, keep)
    sess.trace_enable()
    sess.train(1, perms=np.ascontiguousarray(g['perms'][:1]) if mt else None)
    trace = sess.trace_read()
    sess.close()
    args_o, _, _ = golden_fit_args(g)
    out = oracle.fit_ex(*args_o, 1, perms=g['perms'][:1] if mt else None, sampler=sampler, seed=31337, max_rejects=0 if mt else 64, want_neg=True)
    assert np.array_equal(trace[:, 0], out['neg'][0])
    assert trace[:, 1].sum() == out['draws'][0]


def test_replay_is_deterministic(gpu_lib):
    g = load_golden('warp_f20')
    runs = []
    for _ in range(2):
        args, w, _ = golden_fit_args(g)
        _rankfm.fit_ex(*args, 2, mode="replay", perms=g['perms'][:2])
        runs.append(w)
    for k in WEIGHTS:
        assert np.array_equal(runs[0][k], runs[1][k])


def test_fit_partial_continues_the_mt_stream_from_1492(gpu_lib):
    """the reference re-seeds MT19937 at every `_fit` call (_rankfm.pyx:182): two 1-epoch calls == the oracle's two calls"""
    g = load_golden('bpr_f16')
    args, w, _ = golden_fit_args(g)
    _rankfm.fit_ex(*args, 1, mode="replay", perms=g['perms'][:1])
    _rankfm.fit_ex(*args, 1, mode="replay", perms=g['perms'][1:2])
    args_o, wo, _ = golden_fit_args(g)
    oracle.fit_ex(*args_o, 1, perms=g['perms'][:1])
    oracle.fit_ex(*args_o, 1, perms=g['perms'][1:2])
    for k in WEIGHTS:
        assert rel_err(w[k], wo[k]) < 1e-4, k


def test_replay_fit_at_cfg1_named_size_matches_reference_golden(gpu_lib):
    """BASELINE.json configs[0] at its NAMED size (10k x 5k, 100k interactions, factors=16, bpr, 5 epochs): the serial
    replay kernel against weights / predict() / recommend() minted from the unmodified reference -- north_star tolerance"""
    p, g = cfg1_case()
    w = {k: v.copy() for k, v in p['w'].items()}
    ui = CSRItems(p['indptr'], p['indices'])
    _rankfm.fit_ex(p['X'], p['sw'], ui, p['x_uf'], p['x_if'], *[w[k] for k in WEIGHTS], *p['hyper'], 1, p['epochs'], mode="replay", perms=p['perms'])
    for k in ('w_i', 'v_u', 'v_i'):
        assert rel_err(w[k], g[k + '_ref']) < 1e-4, k
    scores = _rankfm._predict(p['pairs'], p['x_uf'], p['x_if'], *[w[k] for k in WEIGHTS])
    assert rel_err(scores, g['scores']) < PREDICT_RTOL
    rec = _rankfm._recommend(p['users'], ui, 10, True, p['x_uf'], p['x_if'], *[w[k] for k in WEIGHTS])
    assert topk_overlap(rec, g['rec_filtered']) >= 0.99


# ---------------------------------------------------------------------------------------------------------------
# training: production (parallel) schedule -- statistical parity
# ---------------------------------------------------------------------------------------------------------------
def test_consecutive_fit_calls_do_not_replay_the_same_randomness(gpu_lib):
    """every `_fit` call used to restart the Philox / Feistel keys at epoch 0, so a loop of fit_partial(epochs=1)
    showed every positive the same negative in the same order each call.  The key now advances with the epochs trained so
    far (`epoch_offset`, kept by the plug-in across calls and by a resident session across trains)."""
    g = load_golden('bpr_f16')
    traces = []
    for offset in (0, 0, 1):
        args, w, _ = golden_fit_args(g)
        keep = []
        sess = _rankfm.Session(_rankfm.fit_problem(*args, mode="production", seed=5, keep=keep, epoch_offset=offset), keep)
        sess.trace_enable()
        sess.train(1)
        traces.append(sess.trace_read().copy())
        if offset == 1:                      # one more epoch on the same session: yet another set of negatives
            sess.train(1)
            traces.append(sess.trace_read().copy())
        sess.close()
    N = len(g['interactions'])
    # BPR: the negative of a position depends only on (row, epoch key): same offset -> same draws (up to Hogwild-free identity)
    assert np.mean(traces[0][:, 0] == traces[1][:, 0]) > 0.999
    assert np.mean(traces[0][:, 0] == traces[2][:, 0]) < 0.1 and np.mean(traces[2][:, 0] == traces[3][:, 0]) < 0.1
    # and through the plug-in: two `_fit` calls advance the module's epoch counter
    _rankfm.set_seed(5)
    args, w, _ = golden_fit_args(g)
    _rankfm._fit(*args, 2, False)
    assert _rankfm._EPOCHS["done"] == 2
    _rankfm._fit(*args, 1, False)
    assert _rankfm._EPOCHS["done"] == 3 and _rankfm._training["hits"] >= 1
    _rankfm.drop_training()


def test_resident_training_session_across_fit_calls(gpu_lib):
    """the fit_partial pattern: `_fit` on the same data keeps interactions / CSR / bitmap in HBM and only moves the weights;
    other data, other hyper-parameters or edited inputs build a new session; results match stateless calls statistically"""
    X = zipf_interactions(500, 300, 15000, seed=9)
    U, I = int(X[:, 0].max()) + 1, int(X[:, 1].max()) + 1
    ui = CSRItems(*csr_of(X, U))
    sw = np.ones(len(X), np.float32)
    x_uf, x_if = features(U, I, 0, 0)
    hyper = (0.01, 0.1, 0.1, 'invscaling', 0.25, 10)
    _rankfm.set_seed(3)
    b0, h0 = _rankfm._training["builds"], _rankfm._training["hits"]
    w = init_weights(U, I, 12, seed=4)
    for _ in range(3):
        _rankfm._fit(X, sw, ui, x_uf, x_if, *[w[k] for k in WEIGHTS], *hyper, 2, False)
    assert _rankfm._training["builds"] == b0 + 1 and _rankfm._training["hits"] == h0 + 2
    ll_resident = _rankfm.last_stats[-1]['log_likelihood']
    # stateless calls from the same start reach the same place (Hogwild tolerance)
    _rankfm.set_resident_training(False)
    w2 = init_weights(U, I, 12, seed=4)
    for _ in range(3):
        _rankfm._fit(X, sw, ui, x_uf, x_if, *[w2[k] for k in WEIGHTS], *hyper, 2, False)
    _rankfm.set_resident_training(True)
    assert abs(ll_resident / _rankfm.last_stats[-1]['log_likelihood'] - 1) < 0.05
    for k in ('v_u', 'v_i', 'w_i'):
        assert abs(np.linalg.norm(w[k]) / np.linalg.norm(w2[k]) - 1) < 0.05, k
    # a changed hyper-parameter, or interactions edited in place, must not hit the cached session
    b1 = _rankfm._training["builds"]
    _rankfm._fit(X, sw, ui, x_uf, x_if, *[w[k] for k in WEIGHTS], 0.02, *hyper[1:], 1, False)
    assert _rankfm._training["builds"] == b1 + 1
    X[::7, 1] = (X[::7, 1] + 1) % I
    ui2 = CSRItems(*csr_of(X, U))
    _rankfm._fit(X, sw, ui2, x_uf, x_if, *[w[k] for k in WEIGHTS], 0.02, *hyper[1:], 1, False)
    assert _rankfm._training["builds"] == b1 + 2
    _rankfm.drop_training()


def test_fit_partial_without_the_features_of_the_first_fit(gpu_lib):
    """fit(user_features, item_features) then fit_partial() without them resets x_uf / x_if to zeros [U,1] / [I,1]
    while v_uf / v_if keep their [P,F] / [Q,F] shapes (rankfm.py:199,211,236): the reference never reads the zero blocks
    (`x_uf_any`, _rankfm.pyx:193-194); here the widths must be reconciled before a pointer crosses the ABI"""
    from rankfm_b200 import RankFM
    rng = np.random.default_rng(0)
    U, I = 60, 40
    X = np.unique(np.stack([rng.integers(0, U, 900), rng.integers(0, I, 900)], 1), axis=0)
    X = np.unique(np.concatenate([X, np.stack([np.arange(U), rng.integers(0, I, U)], 1), np.stack([rng.integers(0, U, I), np.arange(I)], 1)]), axis=0)
    uf = np.concatenate([np.arange(U)[:, None], rng.random((U, 3))], 1)
    itf = np.concatenate([np.arange(I)[:, None], rng.random((I, 5))], 1)
    model = RankFM(factors=6, loss='warp', max_samples=4).fit(X, user_features=uf, item_features=itf, epochs=2)
    v_uf_before = model.v_uf.copy()
    model.fit_partial(X[:200], epochs=2)
    assert model.x_uf.shape == (U, 1) and model.v_uf.shape == (3, 6) and model.v_if.shape == (5, 6)
    assert np.array_equal(model.v_uf, v_uf_before)                       # inactive blocks are left untouched, like the reference
    assert np.isfinite(model.predict(X[:50])).all()
    assert model.recommend(np.arange(5), n_items=3).shape == (5, 3)

def _auc(w, X, indptr, indices, n=20000, seed=3):
    """P(score(u, observed i) > score(u, random unobserved j)) under the trained model"""
    rng = np.random.default_rng(seed)
    rows = rng.integers(0, len(X), n)
    u, i = X[rows, 0], X[rows, 1]
    j = rng.integers(0, w['v_i'].shape[0], n)
    keep = np.array([jj not in set(indices[indptr[uu]:indptr[uu + 1]].tolist()) for uu, jj in zip(u[:2000], j[:2000])])
    u, i, j = u[:2000][keep], i[:2000][keep], j[:2000][keep]
    s = lambda it: w['w_i'][it] + np.einsum('nf,nf->n', w['v_u'][u], w['v_i'][it])
    return float(np.mean(s(i) > s(j)))


@pytest.mark.parametrize("loss,max_samples,F", [("bpr", 1, 16), ("warp", 10, 20)])
def test_parallel_fit_statistical_parity(gpu_lib, loss, max_samples, F):
    U, I = 2000, 1000
    X = zipf_interactions(U, I, 60000, seed=42)
    U, I = int(X[:, 0].max()) + 1, int(X[:, 1].max()) + 1
    indptr, indices = csr_of(X, U)
    ui = CSRItems(indptr, indices)
    sw = np.ones(len(X), np.float32)
    x_uf, x_if = features(U, I, 0, 0)
    epochs = 5
    hyper = (0.01, 0.1, 0.1, 'invscaling', 0.25, max_samples)
    wg = init_weights(U, I, F, seed=0)
    stats = _rankfm.fit_ex(X, sw, ui, x_uf, x_if, *[wg[k] for k in WEIGHTS], *hyper, epochs, mode="production", seed=7)
    wo = init_weights(U, I, F, seed=0)
    rng = np.random.RandomState(0)
    perms = np.stack([rng.permutation(len(X)) for _ in range(epochs)]).astype(np.int32)
    out = oracle.fit_ex(X, sw, ui, x_uf, x_if, *[wo[k] for k in WEIGHTS], *hyper, epochs, perms=perms, sampler="mt")
    ll_g, ll_o = np.array([s['log_likelihood'] for s in stats]), out['ll'].astype(np.float64)
    # epoch log-likelihoods track the sequential reference: the very first epoch lags a little (up to 1/8 of an
    # epoch is in flight with stale weights while the item biases learn fastest), later epochs agree within 3 %
    np.testing.assert_allclose(ll_g[:1], ll_o[:1], rtol=0.10)
    np.testing.assert_allclose(ll_g[1:], ll_o[1:], rtol=0.03)
    assert ll_g[-1] > ll_g[0]
    draws_g, draws_o = np.array([s['draws'] for s in stats], float), out['draws'].astype(float)
    np.testing.assert_allclose(draws_g, draws_o, rtol=0.05)
    if loss == "bpr":
        assert (draws_g == len(X)).all()
    # ranking quality of the two trained models agrees
    auc_g, auc_o = _auc(wg, X, indptr, indices), _auc(wo, X, indptr, indices)
    assert abs(auc_g - auc_o) < 0.02 and auc_g > 0.7
    # weights have the same scale
    for k in ('v_u', 'v_i', 'w_i'):
        assert abs(np.linalg.norm(wg[k]) / np.linalg.norm(wo[k]) - 1) < 0.05, k


@pytest.mark.parametrize("F,P,Q", [(20, 8, 8), (64, 8, 8), (64, 3, 0), (128, 0, 5), (200, 8, 8), (33, 7, 2)])
def test_feat8_side_feature_code_matches_the_generic_code(gpu_lib, F, P, Q):
    """the P, Q <= 8 specialisation of the production kernel's side-feature math (csrc/rfm_feat8.cuh: compile-time unrolled
    column loops, multi-value reduce-scatter butterfly, two-FMA chain updates) against the generic run-time loops on the
    same pseudo-random rows (device self-test, one warp): hoisted user vectors a[] / b[], the feature parameters after
    one gradient step and the row deltas agree to float32 reassociation"""
    import ctypes as C
    out = (C.c_float * 5)()
    for seed in (1, 2, 3):
        _lib.check(_lib.lib().rfm_debug_feat8(F, P, Q, seed, out))
        d_a, d_b, d_gp, d_rows, moved = list(out)
        assert d_a < 2e-6 and d_b < 2e-5 and d_gp < 2e-6 and d_rows < 2e-6, (F, P, Q, seed, list(out))
        assert moved > 1e-4                                   # the step really changed the parameters


@pytest.mark.parametrize("F,P,Q,feat8,chain", [(12, 5, 6, "1", "warp"), (64, 8, 8, "1", "warp"), (64, 8, 8, "1", "group"), (64, 8, 8, "0", "warp"),
                                               (40, 3, 8, "1", "warp"), (128, 8, 8, "1", "warp"), (64, 8, 8, "1", "race")])
def test_parallel_fit_with_features_statistical_parity(gpu_lib, F, P, Q, feat8, chain, monkeypatch):
    """(64, 8, 8) is BASELINE.json configs[2]'s row shape.  Default: the feat8 code path on half-width lane groups (G = 8, two
    quads per lane) with ONE chain per warp that advances by the winner group's update, applied by all 32 lanes; chain =
    "race": the same chain with every group storing its update (the last writer wins; round-2 second version); chain =
    "group": wide groups, one chain per lane group (round-2 first version); feat8 = "0": the generic run-time loops with
    group-private chains"""
    monkeypatch.setenv("RANKFM_B200_FEAT8", feat8)
    monkeypatch.setenv("RANKFM_B200_CHAIN", chain)
    U, I = 1500, 800
    X = zipf_interactions(U, I, 40000, seed=5)
    U, I = int(X[:, 0].max()) + 1, int(X[:, 1].max()) + 1
    indptr, indices = csr_of(X, U)
    ui = CSRItems(indptr, indices)
    sw = np.random.default_rng(1).uniform(0.5, 1.5, len(X)).astype(np.float32)
    x_uf, x_if = features(U, I, P, Q)
    hyper = (0.01, 0.1, 0.05, 'constant', 0.25, 5)
    wg = init_weights(U, I, F, P, Q, seed=2)
    stats = _rankfm.fit_ex(X, sw, ui, x_uf, x_if, *[wg[k] for k in WEIGHTS], *hyper, 4, mode="production", seed=9)
    wo = init_weights(U, I, F, P, Q, seed=2)
    perms = np.stack([np.random.RandomState(e).permutation(len(X)) for e in range(4)]).astype(np.int32)
    out = oracle.fit_ex(X, sw, ui, x_uf, x_if, *[wo[k] for k in WEIGHTS], *hyper, 4, perms=perms, sampler="mt")
    ll_g, ll_o = np.array([s['log_likelihood'] for s in stats]), out['ll'].astype(np.float64)
    np.testing.assert_allclose(ll_g[:1], ll_o[:1], rtol=0.10)
    np.testing.assert_allclose(ll_g[1:], ll_o[1:], rtol=0.03)
    for k in ('w_i', 'v_u', 'v_i'):
        assert abs(np.linalg.norm(wg[k]) / np.linalg.norm(wo[k]) - 1) < 0.10, k
    # Feature parameters: in the reference they are an exponential moving average over the last ~1/(2*beta*eta) = 100
    # interactions (every positive decays and pushes all of them), i.e. mostly sampling noise around a small mean.  The
    # production schedule runs one such chain per warp and returns their fold, which keeps the mean and sheds the noise:
    # finite, and never larger than the reference's noisy sample.
    for k in ('w_if', 'v_uf', 'v_if'):
        assert np.isfinite(wg[k]).all() and 0 < np.linalg.norm(wg[k]) < 1.5 * np.linalg.norm(wo[k]), k
    # ... and the two models rank alike
    pairs = np.ascontiguousarray(X[:5000].astype(np.float32))
    sg = _rankfm._predict(pairs, x_uf, x_if, *[wg[k] for k in WEIGHTS])
    so = oracle._predict(pairs, x_uf, x_if, *[wo[k] for k in WEIGHTS])
    assert np.corrcoef(sg, so)[0, 1] > 0.9


def test_fit_size_independent_properties_at_cfg2_shape(gpu_lib):
    """BASELINE.json configs[1] at full size: MovieLens-1M shape, factors=20, warp, max_samples=20"""
    X = zipf_interactions(6040, 3706, 1_600_000, seed=42, a_u=0.6, a_i=1.0)[:1_000_000]
    U, I = int(X[:, 0].max()) + 1, int(X[:, 1].max()) + 1
    indptr, indices = csr_of(X, U)
    ui = CSRItems(indptr, indices)
    sw = np.ones(len(X), np.float32)
    x_uf, x_if = features(U, I, 0, 0)
    w = init_weights(U, I, 20, seed=0)
    w0 = {k: v.copy() for k, v in w.items()}
    stats = _rankfm.fit_ex(X, sw, ui, x_uf, x_if, *[w[k] for k in WEIGHTS], 0.01, 0.1, 0.1, 'invscaling', 0.25, 20, 4, mode="production", seed=1)
    N = len(X)
    for s in stats:
        assert N <= s['draws'] <= 20 * N and all(s['finite']) and np.isfinite(s['log_likelihood'])
    ll = [s['log_likelihood'] for s in stats]
    assert ll[-1] > ll[0] and -N * 0.8 < ll[0] < -N * 0.3          # starts near N*log(0.5)
    assert stats[1]['eta'] == np.float32(0.1 / 2 ** 0.25)
    for k in ('w_i', 'v_u', 'v_i'):
        assert np.isfinite(w[k]).all() and not np.array_equal(w[k], w0[k])
    assert np.array_equal(w['v_uf'], w0['v_uf']) and np.array_equal(w['w_if'], w0['w_if'])
    # every user row moved (each user has >= 1 interaction): the epoch permutation visits every row
    moved = np.abs(w['v_u'] - w0['v_u']).max(axis=1) > 0
    assert moved.all()
    # the same 4 epochs on the sequential CPU oracle (MT19937 negatives, a different but equally random row order):
    # per-epoch log-likelihood and draw counts of the Hogwild run track it at the benchmark's full size
    wo = {k: v.copy() for k, v in w0.items()}
    perms = np.stack([np.random.RandomState(e).permutation(N) for e in range(4)]).astype(np.int32)
    out = oracle.fit_ex(X, sw, ui, x_uf, x_if, *[wo[k] for k in WEIGHTS], 0.01, 0.1, 0.1, 'invscaling', 0.25, 20, 4, perms=perms, sampler="mt")
    np.testing.assert_allclose(ll[:1], out['ll'][:1].astype(np.float64), rtol=0.10)
    np.testing.assert_allclose(ll[1:], out['ll'][1:].astype(np.float64), rtol=0.03)
    np.testing.assert_allclose([s['draws'] for s in stats], out['draws'].astype(np.float64), rtol=0.05)
    for k in ('w_i', 'v_u', 'v_i'):
        assert abs(np.linalg.norm(w[k]) / np.linalg.norm(wo[k]) - 1) < 0.05, k
    # a zero-learning-rate-like run leaves weights (almost) untouched: linearity in eta
    w2 = {k: v.copy() for k, v in w0.items()}
    _rankfm.fit_ex(X, sw, ui, x_uf, x_if, *[w2[k] for k in WEIGHTS], 0.01, 0.1, 1e-9, 'constant', 0.25, 20, 1, mode="production", seed=1)
    assert np.abs(w2['v_u'] - w0['v_u']).max() < 1e-6


# ---------------------------------------------------------------------------------------------------------------
# edge cases
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("F", [1, 2, 3, 10, 33, 64, 128, 130, 300])
def test_replay_fit_any_factor_count(gpu_lib, F):
    U, I = 40, 30
    X = zipf_interactions(U, I, 400, seed=F)
    U, I = int(X[:, 0].max()) + 1, int(X[:, 1].max()) + 1
    indptr, indices = csr_of(X, U)
    ui = CSRItems(indptr, indices)
    sw = np.ones(len(X), np.float32)
    x_uf, x_if = features(U, I, 0, 0)
    perms = np.stack([np.random.RandomState(e).permutation(len(X)) for e in range(2)]).astype(np.int32)
    wg, wo = init_weights(U, I, F, seed=1), init_weights(U, I, F, seed=1)
    hyper = (0.01, 0.1, 0.1, 'constant', 0.25, 4)
    _rankfm.fit_ex(X, sw, ui, x_uf, x_if, *[wg[k] for k in WEIGHTS], *hyper, 2, mode="replay", perms=perms)
    oracle.fit_ex(X, sw, ui, x_uf, x_if, *[wo[k] for k in WEIGHTS], *hyper, 2, perms=perms)
    for k in WEIGHTS:
        assert rel_err(wg[k], wo[k]) < 1e-4, k


def test_tiny_and_ragged_inputs(gpu_lib):
    # one interaction, a user who has seen every item but one, duplicate interactions
    X = np.array([[0, 0], [1, 0], [1, 1], [1, 2], [1, 3], [2, 4], [2, 4]], np.int32)
    U, I, F = 3, 5, 4
    indptr, indices = csr_of(X, U)
    ui = CSRItems(indptr, indices)
    sw = np.ones(len(X), np.float32)
    x_uf, x_if = features(U, I, 0, 0)
    perms = np.stack([np.random.RandomState(e).permutation(len(X)) for e in range(3)]).astype(np.int32)
    wg, wo = init_weights(U, I, F, seed=1), init_weights(U, I, F, seed=1)
    hyper = (0.01, 0.1, 0.1, 'constant', 0.25, 3)
    _rankfm.fit_ex(X, sw, ui, x_uf, x_if, *[wg[k] for k in WEIGHTS], *hyper, 3, mode="replay", perms=perms)
    oracle.fit_ex(X, sw, ui, x_uf, x_if, *[wo[k] for k in WEIGHTS], *hyper, 3, perms=perms)
    for k in WEIGHTS:
        assert rel_err(wg[k], wo[k]) < 1e-4, k
    # production mode on the same tiny problem just has to run and stay finite
    wp = init_weights(U, I, F, seed=1)
    stats = _rankfm.fit_ex(X, sw, ui, x_uf, x_if, *[wp[k] for k in WEIGHTS], *hyper, 3, mode="production")
    assert all(all(s['finite']) for s in stats)
    # user 1 can only ever draw item 4
    X1 = X[1:5]
    w1 = init_weights(U, I, F, seed=1)
    before = w1['v_i'].copy()
    _rankfm.fit_ex(X1, np.ones(4, np.float32), ui, x_uf, x_if, *[w1[k] for k in WEIGHTS], *hyper, 1, mode="production")
    assert np.abs(w1['v_i'][4] - before[4]).max() > 0


def test_nonfinite_weights_raise_like_the_reference(gpu_lib):
    g = load_golden('bpr_f16')
    args, w, _ = golden_fit_args(g)
    args = list(args)
    args[1] = np.full_like(g['sample_weight'], 1e30)
    args[13] = 1e10      # learning_rate
    with pytest.raises(AssertionError, match="are not finite - try decreasing feature/sample_weight magnitudes"):
        _rankfm.fit_ex(*args, 2, mode="production")
    assert not np.isfinite(w['v_i']).all()        # weights are written back in the non-finite state, like the reference


@pytest.mark.parametrize("case", ["bpr_f16", "warp_feat"])
def test_epoch_end_penalty_and_verbose_line_match_the_reference_formulas(gpu_lib, case, capsys):
    """epoch end (`_rankfm.pyx:328-336`): the penalty reported with the last epoch is the reference's `reg_penalty` (`:106-116`:
    alpha * sum(w_i^2, v_u^2, v_i^2) + beta * sum(w_if^2, v_uf^2, v_if^2)) of the weights the call returns, and a verbose fit
    prints, per epoch, the two lines of `:332-336` with round(log-likelihood - penalty, 2)"""
    g = load_golden(case)
    args, w, _ = golden_fit_args(g)
    stats = _rankfm.fit_ex(*args, g['epochs'], mode="replay", perms=g['perms'])
    alpha, beta = float(g['hyper'][0]), float(g['hyper'][1])
    sq = lambda a: float(np.sum(np.square(a.astype(np.float64))))
    want = alpha * (sq(w['w_i']) + sq(w['v_u']) + sq(w['v_i'])) + beta * (sq(w['w_if']) + sq(w['v_uf']) + sq(w['v_if']))
    assert stats[-1]['penalty'] == pytest.approx(want, rel=1e-5)
    # the log-likelihood of the replay run is the reference's (oracle) to float32 accumulation noise
    args_o, wo, _ = golden_fit_args(g)
    out = oracle.fit_ex(*args_o, g['epochs'], perms=g['perms'])
    assert stats[-1]['log_likelihood'] == pytest.approx(float(out['ll'][-1]), rel=2e-4)
    # verbose: one "training epoch" / "log likelihood" pair per epoch, printed by the plug-in like the reference
    args_v, _, _ = golden_fit_args(g)
    capsys.readouterr()
    _rankfm._fit(*args_v, 2, True)
    lines = [l for l in capsys.readouterr().out.splitlines() if l.strip()]
    assert [l.split(":")[0] for l in lines] == ["training epoch", "log likelihood"] * 2
    assert lines[0] == "training epoch: 0" and lines[2] == "training epoch: 1"
    ll_printed = float(lines[3].split(":")[1])
    assert ll_printed == pytest.approx(round(_rankfm.last_stats[-1]['log_likelihood'] - _rankfm.last_stats[-1]['penalty'], 2), abs=0.011)
    _rankfm.drop_training()


# ---------------------------------------------------------------------------------------------------------------
# predict / recommend / similar
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", KERNEL_CASES)
def test_predict_matches_reference_golden(gpu_lib, case):
    g = load_golden(case)
    scores = _rankfm._predict(np.ascontiguousarray(g['pairs']), g['x_uf'], g['x_if'], *_weights(g, 'ref'))
    assert scores.dtype == np.float32 and scores.shape == g['scores'].shape
    assert np.array_equal(np.isnan(scores), np.isnan(g['scores']))
    assert rel_err(scores, g['scores']) < 1e-5
    assert _rankfm._predict(np.zeros((0, 2), np.float32), g['x_uf'], g['x_if'], *_weights(g, 'ref')).shape == (0,)


@pytest.mark.parametrize("case", KERNEL_CASES)
@pytest.mark.parametrize("filt", [False, True])
def test_recommend_matches_reference_golden(gpu_lib, case, filt):
    g = load_golden(case)
    _, _, ui = golden_fit_args(g)
    rec = _rankfm._recommend(g['users'], ui, 10, filt, g['x_uf'], g['x_if'], *_weights(g, 'ref'))
    want = g['rec_filtered'] if filt else g['rec']
    assert rec.dtype == np.float32 and rec.shape == want.shape
    assert np.array_equal(np.isnan(rec), np.isnan(want))
    assert topk_overlap(rec, want) >= 0.99                       # north_star: top-k set overlap >= 0.99
    assert np.mean(rec[~np.isnan(want)] == want[~np.isnan(want)]) >= 0.98      # and (ties aside) the same order
    if filt:
        for row, u in zip(rec, g['users']):
            if not np.isnan(u):
                assert not set(row.astype(int).tolist()) & set(ui[int(u)].tolist())


def test_recommend_more_than_available_and_large_n(gpu_lib):
    g = load_golden('bpr_f16')
    _, _, ui = golden_fit_args(g)
    I = g['v_i_ref'].shape[0]
    users = np.array([0, 1, np.nan, 2], np.float32)
    rec = _rankfm._recommend(users, ui, I, True, g['x_uf'], g['x_if'], *_weights(g, 'ref'))
    for row, u in zip(rec, users):
        if np.isnan(u):
            assert np.isnan(row).all()
            continue
        seen = len(set(ui[int(u)].tolist()))
        assert np.isnan(row[I - seen:]).all() and not np.isnan(row[:I - seen]).any()
        assert len(set(row[:I - seen].tolist())) == I - seen
    full = _rankfm._recommend(users[:1], ui, I, False, g['x_uf'], g['x_if'], *_weights(g, 'ref'))
    scores = oracle.scores_user(0, g['x_uf'], g['x_if'], *_weights(g, 'ref'))
    assert np.all(np.diff(scores[full[0].astype(int)]) <= 1e-6)       # sorted by score, descending


def test_similar_items_and_users(gpu_lib):
    g = load_golden('warp_feat')
    w = dict(zip(WEIGHTS, _weights(g, 'ref')))
    for which, v, x, vf in ((0, w['v_i'], g['x_if'], w['v_if']), (1, w['v_u'], g['x_uf'], w['v_uf'])):
        rep = v.astype(np.float64) + x.astype(np.float64) @ vf.astype(np.float64)
        for idx in (0, 7, len(v) - 1):
            sims = rep @ rep[idx]
            # float32 rounding bound of one inner product (the kernel reduces F + P (Q) products in float32)
            tol = 8 * np.finfo(np.float32).eps * float(np.max(np.abs(rep) @ np.abs(rep[idx])))
            sims[idx] = -np.inf
            got = _rankfm._similar(which, idx, 8, g['x_uf'], g['x_if'], *w.values())
            assert idx not in got
            # exact top-8: only candidates closer than float32 rounding may swap (the reference sorts float32 sums too, in
            # BLAS order: rankfm.py:421-424)
            assert_topn_exact_up_to_rounding(got, sims, 8, tol)


def test_similar_batch_matches_single_queries(gpu_lib):
    """the batched all-items variant returns, row by row, what the one-query call returns"""
    g = load_golden('warp_feat')
    w = dict(zip(WEIGHTS, _weights(g, 'ref')))
    for which, rows in ((0, len(w['v_i'])), (1, len(w['v_u']))):
        queries = np.arange(rows, dtype=np.int32)                 # every row of the table
        got = _rankfm._similar_batch(which, queries, 6, g['x_uf'], g['x_if'], *w.values())
        assert got.shape == (rows, 6) and (got >= 0).all()
        for idx in (0, 3, rows // 2, rows - 1):
            assert np.array_equal(got[idx], _rankfm._similar(which, idx, 6, g['x_uf'], g['x_if'], *w.values()))
            assert idx not in got[idx]
    # through the class: a DataFrame of item ids indexed by item id
    from rankfm_b200 import RankFM
    ga = load_golden('api_warp_feat')
    model = RankFM(factors=5, loss='warp', max_samples=6, learning_schedule='invscaling')
    model.fit(ga['interactions'], user_features=ga['user_features'], item_features=ga['item_features'], epochs=1)
    table = model.similar_items_batch(n_items=4)
    assert table.shape == (len(model.item_id), 4) and np.array_equal(table.index.values, model.item_id.values)
    some = model.item_id.values[5]
    assert np.array_equal(table.loc[some].values, model.similar_items(some, 4))


# ---------------------------------------------------------------------------------------------------------------
# the public class, end to end, against the reference's class (golden minted through rankfm.rankfm.RankFM)
# ---------------------------------------------------------------------------------------------------------------
def test_rankfm_class_replay_matches_reference_class(gpu_lib):
    from rankfm_b200 import RankFM
    g = load_golden('api_warp_feat')
    _rankfm.set_mode("replay")
    try:
        model = RankFM(factors=5, loss='warp', max_samples=6, learning_schedule='invscaling')
        np.random.seed(int(g['seed']))
        model.fit(g['interactions'], user_features=g['user_features'], item_features=g['item_features'],
                  sample_weight=g['sample_weight'], epochs=3)
    finally:
        _rankfm.set_mode("production")
    for k in WEIGHTS:
        assert rel_err(getattr(model, k), g[k + '_ref']) < 1e-4, k
    scores = model.predict(g['pairs'])
    assert np.array_equal(np.isnan(scores), np.isnan(g['scores'])) and rel_err(scores, g['scores']) < PREDICT_RTOL
    rec = model.recommend(g['users'], n_items=7).values.astype(np.float64)
    assert np.array_equal(np.isnan(rec), np.isnan(g['rec'])) and topk_overlap(rec, g['rec']) >= 0.99
    rec_f = model.recommend(g['users'], n_items=7, filter_previous=True).values.astype(np.float64)
    assert topk_overlap(rec_f, g['rec_filtered']) >= 0.99
    # similar_*: exact top-5 of the float64 similarities of THIS model (trained to within 1e-4 of the reference's), up to
    # float32 rounding at near-ties; and the reference's own answer may differ from it only by such near-ties
    iid = np.unique(g['interactions'][:, 1]); uid = np.unique(g['interactions'][:, 0])
    for ids, query, rep32, got, ref in ((iid, 3, model.v_i + model.x_if @ model.v_if, model.similar_items(iid[3], 5), g['sim_items']),
                                        (uid, 7, model.v_u + model.x_uf @ model.v_uf, model.similar_users(uid[7], 5), g['sim_users'])):
        rep = rep32.astype(np.float64)
        sims = rep @ rep[query]
        tol = 8 * np.finfo(np.float32).eps * float(np.max(np.abs(rep) @ np.abs(rep[query]))) + 2e-4 * float(np.max(np.abs(sims)))   # + the 1e-4 trajectory tolerance
        sims[query] = -np.inf
        index_of = {v: k for k, v in enumerate(ids.tolist())}
        assert_topn_exact_up_to_rounding([index_of[v] for v in got.tolist()], sims, 5, tol)
        assert_topn_exact_up_to_rounding([index_of[v] for v in ref.tolist()], sims, 5, tol)


# ---------------------------------------------------------------------------------------------------------------
# tensor-core (tcgen05) candidate generation for recommend
# ---------------------------------------------------------------------------------------------------------------
def _scoring_session(U, I, F, P, Q, seed):
    rng0 = np.random.default_rng(seed + 99)
    X = np.concatenate([np.stack([rng0.integers(0, U, 4 * U), rng0.integers(0, I, 4 * U)], 1),
                        np.stack([np.arange(U), rng0.integers(0, I, U)], 1), np.stack([rng0.integers(0, U, I), np.arange(I)], 1)])
    X = np.unique(X, axis=0).astype(np.int32)              # every user and every item occurs: U and I are exactly as asked
    indptr, indices = csr_of(X, U)
    ui = CSRItems(indptr, indices)
    x_uf, x_if = features(U, I, P, Q, seed=seed)
    w = init_weights(U, I, F, P, Q, seed=seed, sigma=0.3)
    w['w_i'][:] = np.random.default_rng(seed).normal(0, 0.5, I).astype(np.float32)
    if Q:
        w['w_if'][:] = np.random.default_rng(seed + 1).normal(0, 0.3, Q).astype(np.float32)
        w['v_if'] *= 10
    if P:
        w['v_uf'] *= 10
    keep = []
    prob = _rankfm.fit_problem(X, np.ones(len(X), np.float32), ui, x_uf, x_if, *[w[k] for k in WEIGHTS], 0.01, 0.1, 0.1, 'constant', 0.25, 1, keep=keep)
    return _rankfm.Session(prob, keep), w, ui, x_uf, x_if, U, I


@pytest.mark.parametrize("msub", ["2", "1"])
@pytest.mark.parametrize("F,P,Q,I", [(16, 0, 0, 1000), (128, 0, 0, 1500), (20, 3, 0, 700), (40, 0, 5, 1100), (100, 4, 6, 900)])
def test_tcgen05_gemm_scores_match_fp32(gpu_lib, F, P, Q, I, msub, monkeypatch):
    """bf16 x bf16 -> fp32 tensor-core scores (TMA + tcgen05.mma + TMEM) against the fp32 oracle utility; with one and
    with two 128-row user sub-tiles per CTA (the second only applies while K <= 128)"""
    monkeypatch.setenv("RANKFM_B200_GEMM_MSUB", msub)
    sess, w, ui, x_uf, x_if, U, I = _scoring_session(300, I, F, P, Q, seed=F)
    users = np.array([0, 1, 5, U - 1, 17, 200, 131, 128, 255, 256, 299], np.float32)
    S = sess.debug_gemm(users)
    sess.close()
    assert S.shape == (len(users), I)
    for r, u in enumerate(users.astype(int)):
        ref = oracle.scores_user(u, x_uf, x_if, *[w[k] for k in WEIGHTS])
        scale = float(np.abs(ref).mean())
        # bf16 operands: ~2^-8 relative per factor product
        assert np.abs(S[r] - ref).max() < 0.03 * max(scale, 1.0), (r, np.abs(S[r] - ref).max(), scale)
        assert np.corrcoef(S[r], ref)[0, 1] > 0.9995


@pytest.mark.parametrize("msub,stride,subset", [("2", "4", "head"), ("2", "1", "head"), ("1", "4", "head"), ("1", "1", "head"), ("2", "8", "head"),
                                                ("2", "4", "stride"), ("1", "2", "stride")])
@pytest.mark.parametrize("F,P,Q", [(32, 0, 0), (20, 2, 3)])
@pytest.mark.parametrize("filt", [False, True])
def test_recommend_tensor_core_path_matches_exact_path(gpu_lib, F, P, Q, filt, msub, stride, subset, monkeypatch):
    """tcgen05 candidate GEMM (user sub-tiles per CTA x pass-1 tile fraction x which tiles) + exact re-score against the
    exact fp32 path"""
    monkeypatch.setenv("RANKFM_B200_GEMM_MSUB", msub)
    monkeypatch.setenv("RANKFM_B200_TAU_STRIDE", stride)
    monkeypatch.setenv("RANKFM_B200_TAU_SUBSET", subset)
    sess, w, ui, x_uf, x_if, U, I = _scoring_session(600, 60000, F, P, Q, seed=7 + F)
    rng = np.random.default_rng(0)
    users = rng.integers(0, U, 300).astype(np.float32)
    users[5] = np.nan
    monkeypatch.setenv("RANKFM_B200_RECOMMEND", "exact")
    exact = sess.recommend(users, 20, filt)
    monkeypatch.setenv("RANKFM_B200_RECOMMEND", "tc")
    fast = sess.recommend(users, 20, filt)
    tc_rows, tc_redone = sess.recommend_stats()
    sess.close()
    assert tc_rows > 0 and tc_redone <= tc_rows // 20, (tc_rows, tc_redone)      # the tensor-core path did serve the rows
    assert np.array_equal(np.isnan(fast), np.isnan(exact))
    assert topk_overlap(fast, exact) >= 0.99                      # north_star: top-k set overlap >= 0.99
    assert np.mean(fast[~np.isnan(exact)] == exact[~np.isnan(exact)]) >= 0.97
    if filt:
        for row, u in zip(fast, users):
            if not np.isnan(u):
                assert not set(row.astype(int).tolist()) & set(ui[int(u)].tolist())


def _sparse_scoring_session(U, I, F, seed, mutate=None):
    """like _scoring_session, but with short user histories (the planner sends users with long ones to the exact path)"""
    rng = np.random.default_rng(seed)
    w = init_weights(U, I, F, seed=seed, sigma=0.3)
    w['w_i'][:] = rng.normal(0, 0.5, I).astype(np.float32)
    if mutate:
        mutate(w)
    x_uf, x_if = features(U, I, 0, 0)
    X = np.unique(np.stack([rng.integers(0, U, 10 * U), rng.integers(0, I, 10 * U)], 1), axis=0).astype(np.int32)
    indptr, indices = csr_of(X, U)
    ui = CSRItems(indptr, indices)
    keep = []
    prob = _rankfm.fit_problem(X, np.ones(len(X), np.float32), ui, x_uf, x_if, *[w[k] for k in WEIGHTS], 0.01, 0.1, 0.1, 'constant', 0.25, 1, keep=keep)
    return _rankfm.Session(prob, keep), ui


def test_recommend_tensor_core_eighth_of_catalogue(gpu_lib, monkeypatch):
    """a catalogue large enough for pass 1 to visit only the highest-bias 1/8 of the item tiles"""
    U = 400
    sess, ui = _sparse_scoring_session(U, 130000, 24, seed=3)
    users = np.arange(0, U, dtype=np.float32)
    monkeypatch.setenv("RANKFM_B200_RECOMMEND", "exact")
    exact = sess.recommend(users, 20, True)
    monkeypatch.setenv("RANKFM_B200_RECOMMEND", "tc")
    fast = sess.recommend(users, 20, True)
    tc_rows, tc_redone = sess.recommend_stats()
    sess.close()
    assert tc_rows == U and tc_redone == 0, (tc_rows, tc_redone)
    assert topk_overlap(fast, exact) >= 0.99
    assert np.mean(fast == exact) >= 0.97
    for row, u in zip(fast, users):
        assert not set(row.astype(int).tolist()) & set(ui[int(u)].tolist())


@pytest.mark.parametrize("mode,z,stride,ew", [("head", "4.5", "32", "8"), ("head", "4.5", "32", "16"), ("estimate", "4.5", "8", "16"), ("estimate", "4.5", "16", "8"),
                                              ("estimate", "4.5", "2", "8"), ("estimate", "0.001", "8", "16"), ("estimate", "0.001", "4", "8")])
@pytest.mark.parametrize("filt", [False, True])
def test_recommend_speculative_thresholds_equal_the_conservative_one(gpu_lib, mode, z, stride, ew, filt, monkeypatch):
    """The row threshold from a small head subset ("head", the default) or estimated from a 1-in-k sample of the item
    tiles ("estimate") must never change a result (rfm_api.cu, tau_mode): rows it does not serve are detected by the
    shortlist kernel and served again with the conservative threshold.  z = 0.001 removes the estimate's head room, so
    most rows take that way.  Also covers 8 vs 16 epilogue warps (different candidate layouts, identical rows)."""
    U = 600
    monkeypatch.setenv("RANKFM_B200_TAU_TAIL", "4")                 # lets a 130 k-item catalogue through the estimate
    monkeypatch.setenv("RANKFM_B200_RECOMMEND", "tc")
    sess, ui = _sparse_scoring_session(U, 130000, 24, seed=5)
    users = np.arange(U, dtype=np.float32)
    users[17] = np.nan
    monkeypatch.setenv("RANKFM_B200_TAU_MODE", "safe")
    safe = sess.recommend(users, 20, filt)
    assert sess.recommend_retried() == 0
    monkeypatch.setenv("RANKFM_B200_GEMM_EW", ew)
    monkeypatch.setenv("RANKFM_B200_TAU_MODE", mode)
    monkeypatch.setenv("RANKFM_B200_TAU_Z", z)
    monkeypatch.setenv("RANKFM_B200_TAU_STRIDE", stride)
    spec = sess.recommend(users, 20, filt)
    retried = sess.recommend_retried()
    tc_rows, tc_redone = sess.recommend_stats()
    monkeypatch.setenv("RANKFM_B200_RECOMMEND", "exact")
    exact = sess.recommend(users, 20, filt)
    sess.close()
    assert np.array_equal(spec, safe, equal_nan=True)
    assert tc_rows == 2 * U and tc_redone <= U // 50, (tc_rows, tc_redone)
    if z == "0.001":
        assert retried >= U // 10, retried                           # the detection + second serving was exercised
    else:
        assert retried <= U // 20, retried
    assert topk_overlap(spec, exact) >= 0.99
    assert np.mean(spec[~np.isnan(exact)] == exact[~np.isnan(exact)]) >= 0.99


def test_recommend_rows_beyond_the_shortlist_staging_are_served_again(gpu_lib, monkeypatch):
    """speculative modes stage only the candidates they expect per row (RANKFM_B200_TC_STAGE, default 2,048); a row with
    more is flagged like an overflowing slot and served again conservatively -- same rows either way"""
    U = 600
    monkeypatch.setenv("RANKFM_B200_RECOMMEND", "tc")
    sess, ui = _sparse_scoring_session(U, 130000, 24, seed=8)
    users = np.arange(U, dtype=np.float32)
    monkeypatch.setenv("RANKFM_B200_TAU_MODE", "safe")
    safe = sess.recommend(users, 100, True)
    monkeypatch.setenv("RANKFM_B200_TAU_MODE", "head")
    monkeypatch.setenv("RANKFM_B200_TC_STAGE", "256")               # every row collects at least its n' = 216 + (items seen) candidates
    head = sess.recommend(users, 100, True)
    retried = sess.recommend_retried()
    sess.close()
    assert np.array_equal(head, safe)
    assert retried > 0, retried


def test_recommend_head_threshold_falls_back_on_a_catalogue_without_bias_signal(gpu_lib, monkeypatch):
    """every item bias equal: the head subset is an arbitrary 1/16 of the catalogue, its threshold is loose, rows collect
    ~16 n' candidates and slots overflow -- those rows are served again conservatively, the results are the conservative
    ones, and once more than 1/32 of a call's rows needed that the session stops speculating"""
    U = 600

    def mutate(w):
        w['w_i'][:] = 0.25
    monkeypatch.setenv("RANKFM_B200_RECOMMEND", "tc")
    sess, ui = _sparse_scoring_session(U, 130000, 24, seed=6, mutate=mutate)
    users = np.arange(U, dtype=np.float32)
    monkeypatch.setenv("RANKFM_B200_TAU_MODE", "safe")
    safe = sess.recommend(users, 100, False)
    monkeypatch.setenv("RANKFM_B200_TAU_MODE", "head")
    head = sess.recommend(users, 100, False)
    first = sess.recommend_retried()
    again = sess.recommend(users, 100, False)
    second = sess.recommend_retried()
    tc_rows, tc_redone = sess.recommend_stats()
    sess.close()
    assert np.array_equal(head, safe) and np.array_equal(again, safe)
    assert first > 0, first
    if first * 32 > U:
        assert second == first                                       # the second call did not speculate
    assert tc_rows == 3 * U and tc_redone <= U // 20, (tc_rows, tc_redone)


def test_recommend_estimated_threshold_is_switched_off_when_it_fails(gpu_lib, monkeypatch):
    """several batches, estimate without head room: the first batch shows that too many rows fall short, the remaining
    batches (and the session's next call) use the conservative threshold; every row equals the conservative result"""
    U = 45000
    monkeypatch.setenv("RANKFM_B200_TAU_TAIL", "2")
    monkeypatch.setenv("RANKFM_B200_RECOMMEND", "tc")
    sess, ui = _sparse_scoring_session(U, 33000, 16, seed=12)
    users = np.arange(U, dtype=np.float32)
    monkeypatch.setenv("RANKFM_B200_TAU_MODE", "safe")
    safe = sess.recommend(users, 10, True)
    monkeypatch.setenv("RANKFM_B200_TAU_MODE", "estimate")
    monkeypatch.setenv("RANKFM_B200_TAU_Z", "0.001")
    est = sess.recommend(users, 10, True)
    first = sess.recommend_retried()
    again = sess.recommend(users[:5000], 10, True)
    second = sess.recommend_retried()
    sess.close()
    assert np.array_equal(est, safe)
    assert np.array_equal(again, safe[:5000])
    assert 0 < first < 30000, first                                  # only the first batch (<= 18,944 rows) was served twice
    assert second == first                                           # the next call started with the provable threshold


def test_recommend_streams_finished_batches_to_the_host(gpu_lib, monkeypatch):
    """a large all-tensor-core call copies each finished batch into the caller's buffer while later batches compute
    (rfm_api.cu HostSink); rows rewritten afterwards (second servings) are copied again.  Same rows as small calls, which
    take the single-copy path."""
    U, n = 45000, 100
    monkeypatch.setenv("RANKFM_B200_STREAM_MIN", str(1 << 20))      # default: from 16 M result floats
    monkeypatch.setenv("RANKFM_B200_RECOMMEND", "tc")
    monkeypatch.setenv("RANKFM_B200_TAU_TAIL", "2")
    sess, ui = _sparse_scoring_session(U, 33000, 16, seed=13)
    users = np.arange(U, dtype=np.float32)
    users[[7, 20000, 44999]] = np.nan
    monkeypatch.setenv("RANKFM_B200_TAU_MODE", "safe")
    parts = np.concatenate([sess.recommend(users[a:a + 9000], n, True) for a in range(0, U, 9000)])
    full = sess.recommend(users, n, True)
    assert np.array_equal(full, parts, equal_nan=True)
    # estimated threshold, verified once with head room, then with less: the first-batch check is off, some rows of every
    # batch fall short, are served again after the batch copies and must be copied again (a few: row by row; many: everything)
    monkeypatch.setenv("RANKFM_B200_TAU_MODE", "estimate")
    monkeypatch.setenv("RANKFM_B200_TAU_Z", "6")
    assert np.array_equal(sess.recommend(users, n, True), parts, equal_nan=True)
    before = sess.recommend_retried()
    seen = []
    for z in ("2.0", "1.0", "0.001"):
        monkeypatch.setenv("RANKFM_B200_TAU_Z", z)
        again = sess.recommend(users, n, True)
        seen.append(sess.recommend_retried() - before)
        before = sess.recommend_retried()
        assert np.array_equal(again, parts, equal_nan=True), z
    sess.close()
    assert seen[-1] > U // 10, seen


def test_recommend_tensor_core_many_batches(gpu_lib, monkeypatch):
    """more users than one wave of CTAs holds: several batches back to back (targets uploaded once, redo flags read once);
    rows from every batch must match the exact path"""
    U = 45000
    sess, ui = _sparse_scoring_session(U, 33000, 16, seed=11)
    users = np.arange(U, dtype=np.float32)
    users[[7, 20000, 44999]] = np.nan
    monkeypatch.setenv("RANKFM_B200_RECOMMEND", "tc")
    fast = sess.recommend(users, 10, True)
    tc_rows, tc_redone = sess.recommend_stats()
    sample = np.concatenate([np.arange(0, 300), np.arange(18800, 19100), np.arange(37700, 38000), np.arange(U - 300, U)])
    monkeypatch.setenv("RANKFM_B200_RECOMMEND", "exact")
    exact = sess.recommend(users[sample], 10, True)
    sess.close()
    assert tc_rows == U and tc_redone <= U // 100, (tc_rows, tc_redone)
    assert np.array_equal(np.isnan(fast[sample]), np.isnan(exact))
    assert topk_overlap(fast[sample], exact) >= 0.99
    assert np.mean(fast[sample][~np.isnan(exact)] == exact[~np.isnan(exact)]) >= 0.97
    assert np.isnan(fast[[7, 20000, 44999]]).all()


def test_recommend_tensor_core_long_histories_use_the_wide_tier(gpu_lib, monkeypatch):
    """filter_previous with long user histories: the shortlist n' = 2 n + 16 + (items seen) of these users exceeds the narrow
    tier (256) -- round 1 sent them to the exact path; the wide tier (n' <= 1024, 2048-entry shortlist) keeps them on the
    tensor cores.  Users beyond 1024 still take the exact path inside the same call; every row must match the exact path."""
    rng = np.random.default_rng(21)
    U, I, F, n = 600, 40000, 24, 100
    w = init_weights(U, I, F, seed=21, sigma=0.3)
    w['w_i'][:] = rng.normal(0, 0.5, I).astype(np.float32)
    x_uf, x_if = features(U, I, 0, 0)
    deg = np.concatenate([np.full(200, 5), rng.integers(100, 750, 380), np.full(20, 1500)])       # narrow / wide / exact users
    rng.shuffle(deg)
    # long histories are made of LIKELY recommendations (high-bias items), so filtering really removes top candidates
    popular = np.argsort(-w['w_i'])[:4000]
    X = np.concatenate([np.stack([np.full(d, u), rng.choice(popular, d, replace=False)], 1) for u, d in enumerate(deg)]).astype(np.int32)
    indptr, indices = csr_of(X, U)
    ui = CSRItems(indptr, indices)
    keep = []
    prob = _rankfm.fit_problem(X, np.ones(len(X), np.float32), ui, x_uf, x_if, *[w[k] for k in WEIGHTS], 0.01, 0.1, 0.1, 'constant', 0.25, 1, keep=keep)
    sess = _rankfm.Session(prob, keep)
    users = np.arange(U, dtype=np.float32)
    monkeypatch.setenv("RANKFM_B200_RECOMMEND", "tc")
    fast = sess.recommend(users, n, True)
    tc_rows, tc_redone = sess.recommend_stats()
    monkeypatch.setenv("RANKFM_B200_RECOMMEND", "exact")
    exact = sess.recommend(users, n, True)
    sess.close()
    need = 2 * n + 16 + deg
    assert tc_rows == int(np.sum(need <= 1024)) and tc_rows > int(np.sum(need <= 256)) > 0, (tc_rows, np.sum(need <= 256), np.sum(need <= 1024))
    assert tc_redone <= tc_rows // 20
    seen = [set(indices[indptr[u]:indptr[u + 1]].tolist()) for u in range(U)]
    assert all(not (set(fast[u].astype(int).tolist()) & seen[u]) for u in range(U))            # nothing seen is recommended
    assert topk_overlap(fast, exact) >= 0.999
    assert np.mean(fast == exact) >= 0.99


def test_recommend_tensor_core_path_matches_the_oracle_on_a_thousand_users(gpu_lib, monkeypatch):
    """north_star: recommend() top-k set overlap >= 0.99 -- the tcgen05 path against the ORACLE's `_recommend` (the C
    restatement of `_rankfm.pyx:393-460`, all-item fp32 scoring + full sort), 1,024 users x 40,000 items, with and without
    filter_previous"""
    rng = np.random.default_rng(33)
    U, I, F = 1024, 40000, 24
    w = init_weights(U, I, F, seed=33, sigma=0.3)
    w['w_i'][:] = rng.normal(0, 0.5, I).astype(np.float32)
    x_uf, x_if = features(U, I, 0, 0)
    X = np.unique(np.stack([rng.integers(0, U, 20 * U), rng.integers(0, I, 20 * U)], 1), axis=0).astype(np.int32)
    indptr, indices = csr_of(X, U)
    ui = CSRItems(indptr, indices)
    users = np.arange(U, dtype=np.float32)
    monkeypatch.setenv("RANKFM_B200_RECOMMEND", "tc")
    for filt in (False, True):
        got = _rankfm._recommend(users, ui, 20, filt, x_uf, x_if, *[w[k] for k in WEIGHTS])
        want = oracle._recommend(users, ui, 20, filt, x_uf, x_if, *[w[k] for k in WEIGHTS])
        assert topk_overlap(got, want) >= 0.99, filt
        assert np.mean(got == want) >= 0.97, filt                  # order too, up to float32 near-ties


@pytest.mark.parametrize("case", ["flat_bias", "all_tied"])
def test_recommend_tensor_core_degenerate_scores(gpu_lib, case, monkeypatch):
    """flat_bias: every item bias equal (bias order degenerates to item order) -> still served by the tensor-core path;
    all_tied: every item identical -> every candidate ties at the cut, slots overflow, rows are redone on the exact path
    and the answer is the exact path's (largest item indexes first, like the reference's reversed argsort)"""
    U, I, F = 200, 40000, 16

    def mutate(w):
        w['w_i'][:] = 0.25
        if case == "all_tied":
            w['v_i'][:] = w['v_i'][0]
    sess, _ = _sparse_scoring_session(U, I, F, seed=1, mutate=mutate)
    users = np.arange(U, dtype=np.float32)
    monkeypatch.setenv("RANKFM_B200_RECOMMEND", "exact")
    exact = sess.recommend(users, 10, False)
    monkeypatch.setenv("RANKFM_B200_RECOMMEND", "tc")
    fast = sess.recommend(users, 10, False)
    tc_rows, tc_redone = sess.recommend_stats()
    sess.close()
    if case == "all_tied":
        assert tc_redone == tc_rows == U
        assert np.array_equal(fast, exact)
        assert np.array_equal(exact[0], np.arange(I - 1, I - 11, -1, dtype=np.float32))
    else:
        assert tc_redone == 0
        assert topk_overlap(fast, exact) >= 0.99


def test_resident_scoring_session_across_calls(gpu_lib):
    """opt-in resident mode: predict/recommend reuse ONE upload of the model until `_fit` changes it"""
    from rankfm_b200 import RankFM
    rng = np.random.default_rng(4)
    X = np.unique(np.stack([rng.integers(0, 300, 6000), rng.integers(100, 500, 6000)], 1), axis=0)
    model = RankFM(factors=8, loss='warp', max_samples=5)
    np.random.seed(1)
    model.fit(X, epochs=2)
    users = list(range(0, 300, 7)) + [10_000]                        # one unknown user
    base_p, base_r = model.predict(X[:500]), model.recommend(users, 5, True)
    base_s = model.similar_items(X[0, 1], 5), model.similar_users(X[0, 0], 5)
    _rankfm.set_resident(True)
    try:
        n0 = _rankfm._scoring["uploads"]
        p1, r1 = model.predict(X[:500]), model.recommend(users, 5, True)
        p2, r2 = model.predict(X[:500]), model.recommend(users, 5, False)
        assert np.array_equal(model.similar_items(X[0, 1], 5), base_s[0]) and np.array_equal(model.similar_users(X[0, 0], 5), base_s[1])
        assert _rankfm._scoring["uploads"] == n0 + 1
        assert np.array_equal(p1, base_p) and np.array_equal(p2, base_p)
        assert r1.equals(base_r) and not r2.equals(base_r)
        model.fit_partial(X, epochs=1)                                  # trains in place -> the resident copy is dropped
        p3 = model.predict(X[:500])
        assert _rankfm._scoring["uploads"] == n0 + 2 and not np.allclose(p3, p1)
        _rankfm.set_resident(False)
        assert np.array_equal(model.predict(X[:500]), p3)
    finally:
        _rankfm.set_resident(False)


# ---------------------------------------------------------------------------------------------------------------
# model quality at the benchmark workload: the Hogwild-trained model ranks held-out interactions like the reference's
# ---------------------------------------------------------------------------------------------------------------
def _topk_metrics(w, train_csr, test_pairs, U, k=10):
    """hit-rate@k and precision@k over users with held-out items, training items filtered (evaluation.py semantics)"""
    indptr, indices = train_csr
    scores = w['v_u'] @ w['v_i'].T + w['w_i'][None, :]
    for u in range(U):
        scores[u, indices[indptr[u]:indptr[u + 1]]] = -np.inf
    top = np.argpartition(-scores, k, axis=1)[:, :k]
    test_sets = {}
    for u, i in test_pairs:
        test_sets.setdefault(int(u), set()).add(int(i))
    hits = np.array([len(test_sets[u] & set(top[u].tolist())) for u in sorted(test_sets)])
    return float((hits > 0).mean()), float((hits / k).mean())


def test_cfg2_model_quality_matches_sequential_reference(gpu_lib):
    """BASELINE.json configs[1] (MovieLens-1M shape, factors=20, warp, max_samples=20, 20 epochs, invscaling): train on
    90 %, rank the held-out 10 %.  Production (Hogwild/Philox) vs the sequential oracle (MT19937): same quality."""
    X = zipf_interactions(6040, 3706, 1_600_000, seed=42)[:1_000_000]
    U, I = int(X[:, 0].max()) + 1, int(X[:, 1].max()) + 1
    rng = np.random.default_rng(5)
    test_mask = rng.random(len(X)) < 0.1
    Xtr, Xte = np.ascontiguousarray(X[~test_mask]), X[test_mask]
    indptr, indices = csr_of(Xtr, U)
    ui = CSRItems(indptr, indices)
    sw = np.ones(len(Xtr), np.float32)
    x_uf, x_if = features(U, I, 0, 0)
    hyper = (0.01, 0.1, 0.1, 'invscaling', 0.25, 20)
    epochs = 20
    wg = init_weights(U, I, 20, seed=0)
    _rankfm.fit_ex(Xtr, sw, ui, x_uf, x_if, *[wg[k] for k in WEIGHTS], *hyper, epochs, mode="production", seed=11)
    wo = init_weights(U, I, 20, seed=0)
    perms = np.stack([np.random.RandomState(e).permutation(len(Xtr)) for e in range(epochs)]).astype(np.int32)
    oracle.fit_ex(Xtr, sw, ui, x_uf, x_if, *[wo[k] for k in WEIGHTS], *hyper, epochs, perms=perms, sampler="mt")
    hr_g, pr_g = _topk_metrics(wg, (indptr, indices), Xte, U)
    hr_o, pr_o = _topk_metrics(wo, (indptr, indices), Xte, U)
    # the two models are equally good rankers (absolute 0.02 on hit rate, relative 5 % on precision)
    assert abs(hr_g - hr_o) < 0.02, (hr_g, hr_o)
    assert abs(pr_g - pr_o) < 0.05 * pr_o + 0.002, (pr_g, pr_o)
    assert hr_g > 0.3
