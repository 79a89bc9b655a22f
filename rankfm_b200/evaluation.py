"""Hold-out ranking metrics with the reference's signatures (``rankfm/evaluation.py:9-175``).

On a GPU the metrics never leave the device until they are five numbers: the test interactions become a
CSR of item indexes per test user, ``rfm_session_evaluate`` runs the recommendation pass (all-item scoring + top-k, the hot
path), tests every recommendation against the user's test items and reduces hit rate, reciprocal rank, DCG, precision and
recall in one kernel.  Without a device (or for a model object that is not this package's ``RankFM``) each metric makes
ONE ``model.recommend(..., cold_start='drop')`` call and reduces the hit matrix with NumPy -- same numbers either way
(``tests/test_api_contract.py::test_evaluation_matches_the_reference_evaluation_module`` pins both on values computed by the
reference's own module).
"""
import numpy as np
import pandas as pd

from rankfm_b200.utils import get_data, lookup_ids

_METRICS = ("hit_rate", "reciprocal_rank", "discounted_cumulative_gain", "precision", "recall")


def _device_path(model):
    """the fused device evaluation applies to this package's RankFM on its CUDA back end"""
    try:
        import rankfm_b200.rankfm as impl
        from rankfm_b200 import _rankfm
        return isinstance(model, impl.RankFM) and impl._recommend is _rankfm._recommend and _rankfm.device_count() > 0
    except Exception:
        return False


def _device_metrics(model, test_interactions, k, filter_previous):
    """-> dict of the five metrics, computed by ``rfm_session_evaluate``"""
    from rankfm_b200 import _rankfm
    assert model.is_fit, "you must fit the model prior to evaluating hold-out metrics"
    test = _test_frame(test_interactions).drop_duplicates()
    u_idx = lookup_ids(test['user_id'].values, model.user_id.values)
    i_idx = lookup_ids(test['item_id'].values, model.item_id.values)
    known = u_idx >= 0                                                   # cold_start='drop': users the model has never seen
    users = np.unique(u_idx[known])
    row = np.searchsorted(users, u_idx[known])
    n_test = np.bincount(row, minlength=len(users))                      # distinct test items per user, unknown items included
    both = known & (i_idx >= 0)
    key = np.sort(np.searchsorted(users, u_idx[both]).astype(np.int64) * len(model.item_id) + i_idx[both])
    indptr = np.zeros(len(users) + 1, dtype=np.int64)
    np.cumsum(np.bincount(key // len(model.item_id), minlength=len(users)), out=indptr[1:])
    out, _ = _rankfm._evaluate(users.astype(np.float32), indptr, (key % len(model.item_id)).astype(np.int32), n_test.astype(np.int32), int(k),
                               bool(filter_previous), model.user_items, *model._weights())
    return dict(zip(_METRICS, out.tolist()))


def _test_frame(test_interactions):
    return pd.DataFrame(get_data(test_interactions), columns=['user_id', 'item_id'])


def _hits(model, test_interactions, k, filter_previous):
    """-> (hit matrix bool [n_common_users, k], number of distinct test items of each of those users)"""
    assert model.is_fit, "you must fit the model prior to evaluating hold-out metrics"
    test = _test_frame(test_interactions).drop_duplicates()
    test_users = test['user_id'].unique()
    recs = model.recommend(users=test_users, n_items=k, filter_previous=filter_previous, cold_start='drop')
    common = recs.index.values
    row_of = pd.Series(np.arange(len(common)), index=common)
    test = test[test['user_id'].isin(common)]
    rows = row_of.reindex(test['user_id'].values).values.astype(np.int64)
    n_test = np.bincount(rows, minlength=len(common))
    # (row, item) membership test through a joint key
    long = recs.stack().reset_index()
    long.columns = ['user_id', 'rank', 'item_id']
    pairs = pd.MultiIndex.from_frame(test[['user_id', 'item_id']])
    hit_flat = pd.MultiIndex.from_frame(long[['user_id', 'item_id']]).isin(pairs)
    hits = np.zeros((len(common), k), dtype=bool)
    hits[row_of.reindex(long['user_id'].values).values.astype(np.int64), long['rank'].values.astype(np.int64)] = hit_flat
    return hits, n_test


def hit_rate(model, test_interactions, k=10, filter_previous=False):
    if _device_path(model):
        return np.float64(_device_metrics(model, test_interactions, k, filter_previous)["hit_rate"])
    return _hit_rate_host(model, test_interactions, k, filter_previous)


def _hit_rate_host(model, test_interactions, k=10, filter_previous=False):
    """share of test users with at least one relevant item among their top-k (``evaluation.py:9-33``)"""
    hits, _ = _hits(model, test_interactions, k, filter_previous)
    return np.mean(hits.any(axis=1).astype(int))


def reciprocal_rank(model, test_interactions, k=10, filter_previous=False):
    if _device_path(model):
        return np.float64(_device_metrics(model, test_interactions, k, filter_previous)["reciprocal_rank"])
    return _reciprocal_rank_host(model, test_interactions, k, filter_previous)


def _reciprocal_rank_host(model, test_interactions, k=10, filter_previous=False):
    """mean of 1/rank of the first relevant recommendation, 0 when none (``evaluation.py:36-61``)"""
    hits, _ = _hits(model, test_interactions, k, filter_previous)
    first = np.argmax(hits, axis=1)
    return np.mean(np.where(hits.any(axis=1), 1.0 / (first + 1), 0.0))


def discounted_cumulative_gain(model, test_interactions, k=10, filter_previous=False):
    if _device_path(model):
        return np.float64(_device_metrics(model, test_interactions, k, filter_previous)["discounted_cumulative_gain"])
    return _discounted_cumulative_gain_host(model, test_interactions, k, filter_previous)


def _discounted_cumulative_gain_host(model, test_interactions, k=10, filter_previous=False):
    """mean of sum over relevant ranks r (0-based) of 1/log2(r+2) (``evaluation.py:64-89``)"""
    hits, _ = _hits(model, test_interactions, k, filter_previous)
    gains = 1.0 / np.log2(np.arange(hits.shape[1]) + 2)
    return np.mean((hits * gains).sum(axis=1))


def precision(model, test_interactions, k=10, filter_previous=False):
    if _device_path(model):
        return np.float64(_device_metrics(model, test_interactions, k, filter_previous)["precision"])
    return _precision_host(model, test_interactions, k, filter_previous)


def _precision_host(model, test_interactions, k=10, filter_previous=False):
    """mean share of the k recommendations that are relevant (``evaluation.py:92-116``)"""
    hits, _ = _hits(model, test_interactions, k, filter_previous)
    return np.mean(hits.sum(axis=1) / hits.shape[1])


def recall(model, test_interactions, k=10, filter_previous=False):
    if _device_path(model):
        return np.float64(_device_metrics(model, test_interactions, k, filter_previous)["recall"])
    return _recall_host(model, test_interactions, k, filter_previous)


def _recall_host(model, test_interactions, k=10, filter_previous=False):
    """mean share of a user's test items that were recommended (``evaluation.py:119-143``)"""
    hits, n_test = _hits(model, test_interactions, k, filter_previous)
    return np.mean(hits.sum(axis=1) / n_test)


def all_metrics(model, test_interactions, k=10, filter_previous=False):
    """the five ranking metrics from ONE recommendation pass (the reference needs five passes, one per metric:
    `examples/movielens.ipynb:1387`); returns a dict keyed by the metric function names"""
    if _device_path(model):
        return _device_metrics(model, test_interactions, k, filter_previous)
    hits, n_test = _hits(model, test_interactions, k, filter_previous)
    any_hit = hits.any(axis=1)
    gains = 1.0 / np.log2(np.arange(hits.shape[1]) + 2)
    return {
        "hit_rate": float(np.mean(any_hit.astype(int))),
        "reciprocal_rank": float(np.mean(np.where(any_hit, 1.0 / (np.argmax(hits, axis=1) + 1), 0.0))),
        "discounted_cumulative_gain": float(np.mean((hits * gains).sum(axis=1))),
        "precision": float(np.mean(hits.sum(axis=1) / hits.shape[1])),
        "recall": float(np.mean(hits.sum(axis=1) / n_test)),
    }


def diversity(model, test_interactions, k=10, filter_previous=False):
    """count / share of users each item is recommended to (``evaluation.py:146-175``)"""
    assert model.is_fit, "you must fit the model prior to evaluating hold-out metrics"
    test_users = _test_frame(test_interactions)['user_id'].unique()
    recs = model.recommend(users=test_users, n_items=k, filter_previous=filter_previous, cold_start='drop')
    n_users = len(recs.index)
    counts = pd.Series(recs.values.ravel()).value_counts()
    user_counts = counts.reindex(model.item_id.values, fill_value=0).rename('cnt_users').rename_axis('item_id')
    user_counts = user_counts.to_frame().sort_values('cnt_users', ascending=False, kind='stable').reset_index()
    user_counts['pct_users'] = user_counts['cnt_users'] / n_users
    return user_counts
