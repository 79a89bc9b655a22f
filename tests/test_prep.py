"""Device-side data preparation (rankfm_b200/csrc/rfm_prep.cu) against the NumPy restatement of what the
reference does with pandas on the host (`rankfm.py:114-177`): bit-exact, these are integer sorts."""
import numpy as np
import pytest

from rankfm_b200 import _rankfm
from rankfm_b200.synthetic import zipf_interactions_device


def _random_interactions(rng, U, I, n, dup=True):
    X = np.stack([rng.integers(0, U, n), rng.integers(0, I, n)], 1).astype(np.int32)
    if dup:
        X = np.concatenate([X, X[rng.integers(0, n, n // 10)]])       # duplicates are kept, like rankfm.py:174
    rng.shuffle(X)
    return np.ascontiguousarray(X)


@pytest.mark.gpu
@pytest.mark.parametrize("U,I,n", [(1, 1, 1), (7, 5, 40), (3000, 1500, 50_000), (200_000, 90_000, 1_500_000)])
def test_device_prep_matches_host_prep(gpu_lib, U, I, n):
    rng = np.random.default_rng(U + n)
    X = _random_interactions(rng, U, I, n)
    want = _rankfm.UserItems.from_interactions_host(X, U)
    indptr, indices = _rankfm.prep_user_items(X, U, I)
    assert indptr.dtype == np.int64 and indices.dtype == np.int32
    assert np.array_equal(indptr, want.indptr) and np.array_equal(indices, want.indices)
    # users without interactions (ragged / empty rows) keep empty segments
    assert indptr[0] == 0 and indptr[-1] == len(X) and (np.diff(indptr) >= 0).all()


@pytest.mark.gpu
def test_device_prep_rejects_out_of_range_indexes(gpu_lib):
    X = np.array([[0, 0], [5, 1]], np.int32)
    with pytest.raises(ValueError):
        _rankfm.prep_user_items(X, 3, 4)


@pytest.mark.gpu
def test_from_interactions_takes_the_device_path_for_large_inputs(gpu_lib):
    rng = np.random.default_rng(3)
    X = _random_interactions(rng, 50_000, 20_000, 400_000)
    a = _rankfm.UserItems.from_interactions(X, 50_000)           # >= _PREP_DEVICE_MIN: device radix sort
    b = _rankfm.UserItems.from_interactions_host(X, 50_000)
    assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)


@pytest.mark.gpu
@pytest.mark.parametrize("n,span", [(1, 10), (1000, 50), (300_000, 10**12)])
def test_device_index_ids_matches_numpy(gpu_lib, n, span):
    rng = np.random.default_rng(n)
    ids = rng.integers(-span, span, n)                            # negative ids too: the order is numeric
    uniq, index = _rankfm.prep_index_ids(ids)
    want_uniq, want_index = np.unique(ids, return_inverse=True)
    assert np.array_equal(uniq, want_uniq) and np.array_equal(index, want_index.astype(np.int32))


@pytest.mark.gpu
def test_synthetic_generator_properties(gpu_lib):
    U, I, N = 20_000, 5_000, 300_000
    X, nu, ni = zipf_interactions_device(U, I, N, seed=7)
    assert X.shape == (N, 2) and X.dtype == np.int32
    keys = X[:, 0].astype(np.int64) * I + X[:, 1]
    assert len(np.unique(keys)) == N                              # de-duplicated
    # re-indexed to the observed uniques: ids are exactly 0 .. n_observed-1
    assert nu == len(np.unique(X[:, 0])) == X[:, 0].max() + 1 and ni == len(np.unique(X[:, 1])) == X[:, 1].max() + 1
    # same seed -> same data; another seed -> other data
    X2, _, _ = zipf_interactions_device(U, I, N, seed=7)
    X3, _, _ = zipf_interactions_device(U, I, N, seed=8)
    assert np.array_equal(X, X2) and not np.array_equal(X, X3)
    # Zipf(1.0) items: the most popular item takes about 1/H_I of the draws (less after de-duplication), far above uniform
    top = np.bincount(X[:, 1]).max() / N
    assert 0.02 < top < 0.2
    # popularity is not index-ordered
    counts = np.bincount(X[:, 1], minlength=ni)
    assert abs(np.corrcoef(np.arange(ni), counts)[0, 1]) < 0.1
    # without re-indexing and with an offset: ids stay inside [offset, offset + U) x [0, I), shared item permutation
    Y, _, _ = zipf_interactions_device(U, I, N, seed=9, offset_users=1000, perm_seed=42, reindex=False)
    Z, _, _ = zipf_interactions_device(U, I, N, seed=10, offset_users=1000, perm_seed=42, reindex=False)
    assert Y[:, 0].min() >= 1000 and Y[:, 0].max() < 1000 + U and Y[:, 1].max() < I
    assert np.argmax(np.bincount(Y[:, 1], minlength=I)) == np.argmax(np.bincount(Z[:, 1], minlength=I))


@pytest.mark.gpu
def test_rankfm_class_device_prep_matches_host_prep(gpu_lib, monkeypatch):
    """`RankFM._init_all` on the device path (id maps + user_items by radix sort) builds exactly the objects of the host path"""
    from rankfm_b200.rankfm import RankFM
    rng = np.random.default_rng(11)
    uid = rng.choice(10**9, 40_000, replace=False)
    iid = rng.choice(10**7, 9_000, replace=False)
    inter = np.stack([uid[rng.integers(0, len(uid), 300_000)], iid[rng.integers(0, len(iid), 300_000)]], 1)

    def build(device):
        monkeypatch.setattr(_rankfm, "_PREP_DEVICE_MIN", 1000 if device else 10**12)
        m = RankFM(factors=4)
        np.random.seed(0)
        m._init_all(inter)
        return m
    a, b = build(True), build(False)
    assert np.array_equal(a.user_id.values, b.user_id.values) and np.array_equal(a.item_id.values, b.item_id.values)
    assert a.user_id.dtype == b.user_id.dtype
    assert np.array_equal(a.interactions, b.interactions) and a.interactions.dtype == np.int32
    assert np.array_equal(a.user_items.indptr, b.user_items.indptr) and np.array_equal(a.user_items.indices, b.user_items.indices)
    assert np.array_equal(a.v_u, b.v_u)                         # same np.random draws either way
