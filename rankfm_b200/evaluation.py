"""Hold-out ranking metrics with the reference's signatures (``rankfm/evaluation.py:9-175``).

Each metric makes ONE ``model.recommend(..., cold_start='drop')`` call for all test users known to the model -- the
recommendation pass (all-item scoring + top-k) is the GPU hot path -- and reduces the hit matrix with NumPy
instead of per-user Python set operations.
"""
import numpy as np
import pandas as pd

from rankfm_b200.utils import get_data


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
    """share of test users with at least one relevant item among their top-k (``evaluation.py:9-33``)"""
    hits, _ = _hits(model, test_interactions, k, filter_previous)
    return np.mean(hits.any(axis=1).astype(int))


def reciprocal_rank(model, test_interactions, k=10, filter_previous=False):
    """mean of 1/rank of the first relevant recommendation, 0 when none (``evaluation.py:36-61``)"""
    hits, _ = _hits(model, test_interactions, k, filter_previous)
    first = np.argmax(hits, axis=1)
    return np.mean(np.where(hits.any(axis=1), 1.0 / (first + 1), 0.0))


def discounted_cumulative_gain(model, test_interactions, k=10, filter_previous=False):
    """mean of sum over relevant ranks r (0-based) of 1/log2(r+2) (``evaluation.py:64-89``)"""
    hits, _ = _hits(model, test_interactions, k, filter_previous)
    gains = 1.0 / np.log2(np.arange(hits.shape[1]) + 2)
    return np.mean((hits * gains).sum(axis=1))


def precision(model, test_interactions, k=10, filter_previous=False):
    """mean share of the k recommendations that are relevant (``evaluation.py:92-116``)"""
    hits, _ = _hits(model, test_interactions, k, filter_previous)
    return np.mean(hits.sum(axis=1) / hits.shape[1])


def recall(model, test_interactions, k=10, filter_previous=False):
    """mean share of a user's test items that were recommended (``evaluation.py:119-143``)"""
    hits, n_test = _hits(model, test_interactions, k, filter_previous)
    return np.mean(hits.sum(axis=1) / n_test)


def all_metrics(model, test_interactions, k=10, filter_previous=False):
    """the five ranking metrics from ONE recommendation pass (the reference needs five passes, one per metric:
    `examples/movielens.ipynb:1387`); returns a dict keyed by the metric function names"""
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
