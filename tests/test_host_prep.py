"""Host-side data preparation of the `RankFM` mirror (`rankfm/rankfm.py:100-177` in the reference): id maps and the
user_items CSR, on the CPU."""
import numpy as np
import pandas as pd
import pytest

from rankfm_b200 import RankFM
from rankfm_b200._rankfm import UserItems
from rankfm_b200.utils import lookup_ids, unique_ids


@pytest.mark.parametrize("dtype", [np.int64, np.int32, np.uint32, np.int16])
def test_unique_and_lookup_match_numpy_and_pandas_for_integer_ids(dtype):
    rng = np.random.default_rng(0)
    hi = min(3_000_000, np.iinfo(dtype).max)
    col = rng.integers(5, hi, 100_000).astype(dtype)
    uniq = unique_ids(col)
    assert np.array_equal(uniq, np.unique(col)) and uniq.dtype == np.unique(col).dtype
    query = rng.integers(-10, hi + 100, 5_000)
    assert np.array_equal(lookup_ids(query, uniq), pd.Index(uniq).get_indexer(pd.Index(query)))


def test_unique_and_lookup_other_id_types():
    col = np.array(['b', 'a', 'c', 'a'], dtype=object)
    assert list(unique_ids(col)) == ['a', 'b', 'c']
    assert list(lookup_ids(np.array(['c', 'z'], dtype=object), unique_ids(col))) == [2, -1]
    sparse = np.array([-5, 3, -5, 10 ** 12])                                    # range too wide for a dense table
    assert np.array_equal(unique_ids(sparse), np.unique(sparse))
    assert list(lookup_ids(np.array([10 ** 12, -5, 0]), np.unique(sparse))) == [2, 0, -1]
    assert list(lookup_ids(np.array([1.0, 2.0, np.nan]), np.array([1, 2, 3]))) == [0, 1, -1]    # float queries: pandas path
    assert len(unique_ids(np.zeros(0, np.int64))) == 0


def test_user_items_from_interactions_keeps_duplicates_sorted():
    X = np.array([[2, 5], [0, 3], [2, 1], [0, 3], [1, 9], [2, 5]], dtype=np.int32)
    ui = UserItems.from_interactions(X, 4)
    assert list(ui.indptr) == [0, 2, 3, 6, 6]
    assert list(ui[0]) == [3, 3] and list(ui[1]) == [9] and list(ui[2]) == [1, 5, 5] and list(ui[3]) == []


def test_init_all_maps_ids_like_the_reference():
    """`rankfm.py:115-128`: sorted unique ids, index = rank of the id; interactions become int32 index pairs"""
    X = np.array([[30, 7], [10, 9], [30, 9], [20, 7], [10, 8]])
    m = RankFM(factors=2)
    np.random.seed(0)
    m._init_all(X)
    assert list(m.user_id) == [10, 20, 30] and list(m.item_id) == [7, 8, 9]
    assert m.interactions.dtype == np.int32 and m.interactions.tolist() == [[2, 0], [0, 2], [2, 2], [1, 0], [0, 1]]
    assert list(m.user_to_index.loc[[30, 10]]) == [2, 0] and list(m.index_to_item.loc[[2, 0]]) == [9, 7]
    assert {u: list(m.user_items[u]) for u in range(3)} == {0: [1, 2], 1: [0], 2: [0, 2]}
    assert m.v_u.shape == (3, 2) and m.v_i.shape == (3, 2) and m.w_i.shape == (3,) and m.sample_weight.tolist() == [1.0] * 5
    with pytest.raises(ValueError):
        m.is_fit = True
        m._init_interactions(np.array([[40, 7]]), None)                      # unseen user id on a warm start
