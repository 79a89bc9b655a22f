"""API-contract tests of the ``RankFM`` class -- the same contract the reference pins in ``tests/test_rankfm.py``
(shapes, dtypes, NaN counts, id membership, exception types; the reference asserts no numeric values there).

Two backends run the same assertions:
  * ``oracle`` (CPU, always runs): the class's native calls are monkey-patched to the CPU oracle, so the HOST logic
    (id indexing, feature tables, cold-start handling, DataFrame assembly) is covered without a GPU;
  * ``cuda``  (``-m gpu``): the real ctypes -> CUDA path.
"""
import numpy as np
import pandas as pd
import pytest

import rankfm_b200.rankfm as rankfm_mod
from rankfm_b200 import _rankfm
from rankfm_b200.evaluation import all_metrics, discounted_cumulative_gain, diversity, hit_rate, precision, recall, reciprocal_rank
from rankfm_b200.rankfm import RankFM

PAIRS = [(1, 1), (1, 3), (1, 5), (2, 1), (2, 2), (2, 6), (3, 3), (3, 6), (3, 4)]
TRAIN_INT = pd.DataFrame(PAIRS, columns=['user_id', 'item_id'], dtype=np.int32)
TRAIN_STR = pd.DataFrame([('XYZ'[u - 1], 'ABCDEF'[i - 1]) for u, i in PAIRS], columns=['user_id', 'item_id'])
TRAIN_NP = np.array(PAIRS)
TRAIN_RATING = pd.DataFrame([(u, i, 3) for u, i in PAIRS], columns=['user_id', 'item_id', 'rating'], dtype=np.int32)
VALID_DISJOINT = pd.DataFrame([(1, 1), (1, 3), (1, 5), (2, 1), (2, 2), (2, 7), (4, 3), (4, 7), (4, 4)], columns=['user_id', 'item_id'], dtype=np.int32)

UF_ROWS = [(1, 0, 1, 5, 3.14), (2, 1, 0, 6, 2.72), (3, 0, 0, 4, 1.62)]
IF_ROWS = [(1, 0, 1, 5, 3.14), (2, 1, 0, 6, 2.72), (3, 0, 0, 4, 1.62), (4, 1, 1, 3, 1.05), (5, 1, 0, 6, 0.33), (6, 0, 0, 0, 0.00)]
UF_PD = pd.DataFrame(UF_ROWS, columns=['user_id', 'bin_1', 'bin_2', 'int', 'cnt'])
IF_PD = pd.DataFrame(IF_ROWS, columns=['item_id', 'bin_1', 'bin_2', 'int', 'cnt'])
UF_NP, IF_NP = np.array(UF_ROWS), np.array(IF_ROWS)
UF_NO_ID = UF_PD.drop(columns='user_id')
IF_NO_ID = IF_PD.drop(columns='item_id')
UF_STR = UF_PD.assign(int=list("ABC"))
IF_STR = IF_PD.assign(int=list("ABCAFG"))
TRAIN_USERS = np.array([1, 2, 3])
VALID_USERS = np.array([1, 2, 4, 5])


def _oracle_similar(which, index, n, x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if):
    rep = (v_i + x_if @ v_if) if which == 0 else (v_u + x_uf @ v_uf)
    sims = rep @ rep[index]
    order = [k for k in np.argsort(-sims, kind='stable') if k != index][:n]
    return np.array(order, dtype=np.int32)


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request, monkeypatch):
    if request.param == "oracle":
        from oracle import oracle
        monkeypatch.setattr(rankfm_mod, "_fit", oracle._fit)
        monkeypatch.setattr(rankfm_mod, "_predict", oracle._predict)
        monkeypatch.setattr(rankfm_mod, "_recommend", oracle._recommend)
        monkeypatch.setattr(rankfm_mod, "_similar", _oracle_similar)
    else:
        assert _rankfm.device_count() > 0
    return request.param


@pytest.mark.parametrize("interactions,user_features,item_features", [
    (TRAIN_INT, None, None), (TRAIN_STR, None, None), (TRAIN_NP, None, None),
    (TRAIN_INT, UF_PD, None), (TRAIN_INT, None, IF_PD), (TRAIN_INT, UF_PD, IF_PD), (TRAIN_INT, UF_NP, IF_NP)])
def test_fit_accepts_the_reference_input_types(backend, interactions, user_features, item_features, capsys):
    model = RankFM(factors=2)
    assert model.fit(interactions, user_features, item_features, epochs=2, verbose=True) is model
    assert model.is_fit
    out = capsys.readouterr().out
    assert out.count("training epoch:") == 2 and out.count("log likelihood:") == 2
    assert model.v_u.shape == (3, 2) and model.v_i.shape == (6, 2) and model.w_i.dtype == np.float32
    assert model.interactions.dtype == np.int32 and model.interactions.shape == (9, 2)
    assert sorted(model.user_items[0].tolist()) == [0, 2, 4]
    assert model.x_uf.shape == ((3, 4) if user_features is not None else (3, 1))
    assert model.x_if.shape == ((6, 4) if item_features is not None else (6, 1))


def test_fit_rejects_bad_inputs(backend):
    with pytest.raises(AssertionError):
        RankFM(factors=2).fit(TRAIN_RATING)
    with pytest.raises(KeyError):
        RankFM(factors=2).fit(TRAIN_INT, user_features=UF_NO_ID)
    with pytest.raises(ValueError):
        RankFM(factors=2).fit(TRAIN_INT, user_features=UF_STR)
    with pytest.raises(KeyError):
        RankFM(factors=2).fit(TRAIN_INT, item_features=IF_NO_ID)
    with pytest.raises(ValueError):
        RankFM(factors=2).fit(TRAIN_INT, item_features=IF_STR)
    with pytest.raises(AssertionError):
        RankFM(factors=2).fit(TRAIN_INT, epochs=0)
    with pytest.raises(AssertionError):
        RankFM(factors=2).fit(TRAIN_INT, sample_weight=np.ones(3))


@pytest.mark.parametrize("kwargs", [dict(factors=0), dict(factors=2.0), dict(loss='hinge'), dict(max_samples=0), dict(alpha=1),
                                    dict(beta=-0.1), dict(sigma=0.0), dict(learning_rate=1), dict(learning_schedule='cosine'),
                                    dict(learning_exponent=0.0)])
def test_constructor_validation(kwargs):
    with pytest.raises(AssertionError):
        RankFM(**kwargs)


def test_predict_contract(backend):
    model = RankFM(factors=2).fit(TRAIN_INT)
    scores = model.predict(TRAIN_INT)
    assert scores.shape == (9,) and scores.dtype == np.float32 and not np.isnan(scores).any()
    scores = model.predict(VALID_DISJOINT, cold_start='nan')
    assert scores.shape == (9,) and scores.dtype == np.float32 and np.isnan(scores).sum() == 4
    scores = model.predict(VALID_DISJOINT, cold_start='drop')
    assert scores.shape == (5,) and not np.isnan(scores).any()
    with pytest.raises(ValueError):
        model.predict(TRAIN_INT, cold_start='keep')
    with pytest.raises(AssertionError):
        RankFM(factors=2).predict(TRAIN_INT)


def test_recommend_contract(backend):
    model = RankFM(factors=2).fit(TRAIN_INT)
    recs = model.recommend(TRAIN_USERS, n_items=3)
    assert isinstance(recs, pd.DataFrame) and recs.shape == (3, 3)
    assert np.array_equal(recs.index.values, TRAIN_USERS)
    assert recs.isin(TRAIN_INT['item_id'].values).all().all()
    assert all(np.issubdtype(t, np.integer) for t in recs.dtypes)

    recs = model.recommend(TRAIN_USERS, n_items=3, filter_previous=True)
    long = recs.stack().reset_index().drop('level_1', axis=1)
    long.columns = ['user_id', 'item_id']
    assert recs.shape == (3, 3) and pd.merge(TRAIN_INT, long, on=['user_id', 'item_id'], how='inner').empty

    recs = model.recommend(VALID_USERS, n_items=3, cold_start='nan')
    assert recs.shape == (4, 3) and sorted(recs.index.values) == sorted(VALID_USERS)
    assert recs.loc[[4, 5]].isnull().all().all() and recs.dropna().isin(TRAIN_INT['item_id'].values).all().all()

    recs = model.recommend(VALID_USERS, n_items=3, cold_start='drop')
    assert recs.shape == (2, 3) and sorted(recs.index.values) == [1, 2]
    with pytest.raises(ValueError):
        model.recommend(TRAIN_USERS, cold_start='keep')


def test_recommend_string_ids(backend):
    model = RankFM(factors=2).fit(TRAIN_STR)
    recs = model.recommend(['X', 'Q'], n_items=2)
    assert recs.shape == (2, 2) and recs.loc['Q'].isnull().all() and recs.loc['X'].isin(list("ABCDEF")).all()


def test_similar_contract(backend):
    model = RankFM(factors=2).fit(TRAIN_INT)
    similar = model.similar_items(1, n_items=3)
    assert similar.shape == (3,) and np.isin(similar, TRAIN_INT['item_id'].unique()).all() and 1 not in similar
    similar = model.similar_users(1, n_users=2)
    assert similar.shape == (2,) and np.isin(similar, TRAIN_INT['user_id'].unique()).all() and 1 not in similar
    with pytest.raises(AssertionError):
        model.similar_items(99, n_items=3)
    with pytest.raises(AssertionError):
        model.similar_users(9, n_users=1)


def test_fit_partial_warm_start(backend):
    model = RankFM(factors=2).fit(TRAIN_INT)
    v_before = model.v_i.copy()
    model.fit_partial(pd.DataFrame([(1, 2), (3, 1)], columns=['user_id', 'item_id'], dtype=np.int32), epochs=1)
    assert model.user_items[0].tolist() == [0, 1, 2, 4] and model.user_items[2].tolist() == [0, 2, 3, 5]
    assert model.interactions.shape == (2, 2) and not np.array_equal(model.v_i, v_before)
    with pytest.raises(ValueError):       # unknown ids cannot be added by a warm start (the reference fails on the NaN cast)
        model.fit_partial(pd.DataFrame([(9, 2)], columns=['user_id', 'item_id'], dtype=np.int32))


def test_model_pickles(backend):
    import pickle
    model = RankFM(factors=2).fit(TRAIN_INT)
    clone = pickle.loads(pickle.dumps(model))
    assert np.array_equal(clone.v_u, model.v_u) and clone.user_items[1].tolist() == model.user_items[1].tolist()
    assert np.allclose(clone.predict(TRAIN_INT), model.predict(TRAIN_INT))


def test_evaluation_metrics(backend):
    model = RankFM(factors=2).fit(TRAIN_INT, epochs=3)
    test = pd.DataFrame([(1, 2), (1, 4), (2, 3), (3, 1), (7, 1)], columns=['user_id', 'item_id'])
    k = 3
    recs = model.recommend([1, 2, 3], n_items=k, filter_previous=True)
    truth = {1: {2, 4}, 2: {3}, 3: {1}}
    hits = {u: [int(i in truth[u]) for i in recs.loc[u]] for u in truth}
    assert hit_rate(model, test, k, True) == pytest.approx(np.mean([max(h) for h in hits.values()]))
    assert precision(model, test, k, True) == pytest.approx(np.mean([sum(h) / k for h in hits.values()]))
    assert recall(model, test, k, True) == pytest.approx(np.mean([sum(hits[u]) / len(truth[u]) for u in truth]))
    assert reciprocal_rank(model, test, k, True) == pytest.approx(np.mean([1 / (h.index(1) + 1) if 1 in h else 0 for h in hits.values()]))
    assert discounted_cumulative_gain(model, test, k, True) == pytest.approx(
        np.mean([sum(x / np.log2(r + 2) for r, x in enumerate(h)) for h in hits.values()]))
    both = all_metrics(model, test, k, True)
    assert both["hit_rate"] == pytest.approx(hit_rate(model, test, k, True)) and both["recall"] == pytest.approx(recall(model, test, k, True))
    assert both["reciprocal_rank"] == pytest.approx(reciprocal_rank(model, test, k, True))
    assert both["discounted_cumulative_gain"] == pytest.approx(discounted_cumulative_gain(model, test, k, True))
    assert both["precision"] == pytest.approx(precision(model, test, k, True))
    div = diversity(model, test, k, True)
    assert list(div.columns) == ['item_id', 'cnt_users', 'pct_users'] and div['cnt_users'].sum() == 3 * k and len(div) == 6


def _reference_trained_model(g):
    """our class carrying the model the REFERENCE trained (tests/golden/make_golden.py::eval_case): same id maps (sorted
    unique ids of the training interactions), the reference's weights"""
    model = RankFM(factors=8, loss='warp', max_samples=5, learning_schedule='invscaling')
    model._init_all(g['train'])
    for k in ('w_i', 'w_if', 'v_u', 'v_i', 'v_uf', 'v_if'):
        assert getattr(model, k).shape == g[k + '_ref'].shape, k
        setattr(model, k, np.ascontiguousarray(g[k + '_ref']))
    model.is_fit = True
    return model


def test_evaluation_matches_the_reference_evaluation_module(backend):
    """the five ranking metrics + diversity against values computed by the REFERENCE's `rankfm/evaluation.py:9-175` through
    the reference's own class on a 1,500-user model (golden eval_ref.npz): same model, same test set, same numbers"""
    import os
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval_ref.npz")))
    model = _reference_trained_model(g)
    test = g['test']
    fns = dict(hit_rate=hit_rate, reciprocal_rank=reciprocal_rank, dcg=discounted_cumulative_gain, precision=precision, recall=recall)
    for k in (5, 10):
        for filt in (False, True):
            tag = "_k%d_%s" % (k, "filt" if filt else "all")
            for name, fn in fns.items():
                assert fn(model, test, k=k, filter_previous=filt) == pytest.approx(float(g[name + tag]), rel=1e-12, abs=1e-15), name + tag
            both = all_metrics(model, test, k=k, filter_previous=filt)
            assert both["hit_rate"] == pytest.approx(float(g["hit_rate" + tag]), rel=1e-12)
            assert both["discounted_cumulative_gain"] == pytest.approx(float(g["dcg" + tag]), rel=1e-12)
            assert both["recall"] == pytest.approx(float(g["recall" + tag]), rel=1e-12)
    if backend == "cuda":
        # the fused device evaluation (rfm_session_evaluate) and the host reduction of the recommend() table agree
        import rankfm_b200.evaluation as ev
        assert ev._device_path(model)
        host = ev._hit_rate_host(model, test, 10, True), ev._recall_host(model, test, 10, True), ev._discounted_cumulative_gain_host(model, test, 10, True)
        dev = ev._device_metrics(model, test, 10, True)
        assert (dev["hit_rate"], dev["recall"], dev["discounted_cumulative_gain"]) == pytest.approx(host, rel=1e-12)
    div = diversity(model, test, k=10, filter_previous=True)
    ours = dict(zip(div['item_id'].values.tolist(), div['cnt_users'].values.tolist()))
    ref = dict(zip(g['diversity_item_id'].tolist(), g['diversity_cnt_users'].tolist()))
    assert ours == ref                                                       # the order among equal counts is unspecified (unstable sort)
    assert np.array_equal(div['cnt_users'].values, g['diversity_cnt_users']) and np.allclose(div['pct_users'].values, g['diversity_pct_users'], rtol=1e-12)


def test_fit_partial_on_the_same_input_keeps_the_prepared_arrays(backend):
    """a loop of fit_partial() on the same interactions must not redo the id lookups / the user_items union every call; any
    other input (another buffer, or the same buffer edited in place) is prepared afresh"""
    rng = np.random.default_rng(5)       # 200 users x 100 items (a 3 x 6 toy repeated 50x is no workload for a Hogwild schedule)
    X = np.unique(np.stack([rng.integers(1, 201, 3000), rng.integers(1, 101, 3000)], 1), axis=0).astype(np.int64)
    X = np.concatenate([X, np.stack([np.arange(1, 201), rng.integers(1, 101, 200)], 1), np.stack([rng.integers(1, 201, 100), np.arange(1, 101)], 1)])
    model = RankFM(factors=2).fit(X, epochs=1)
    prepared, items = model.interactions, model.user_items
    model.fit_partial(X, epochs=1)
    assert model.interactions is prepared and model.user_items is items
    X[0] = (3, 1)                                       # edited in place: same address, other content
    model.fit_partial(X, epochs=1)
    assert model.interactions is not prepared and model.interactions[0].tolist() == [2, 0]      # ids 1..N -> indexes 0..N-1
    prepared = model.interactions
    model.fit_partial(X.copy(), epochs=1)               # equal content in another buffer: prepared again (the stamp includes the address)
    assert model.interactions is not prepared
    w = np.ones(len(X), dtype=np.float32)
    model.fit_partial(X, sample_weight=w, epochs=1)     # other sample weights: prepared again
    assert model.sample_weight is not None and np.array_equal(model.sample_weight, w)
