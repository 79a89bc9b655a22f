"""``RankFM`` model class: the reference's public surface (``rankfm/rankfm.py:11-454``) on top of the B200 kernels.

Same constructor arguments / defaults / validation, same public attributes (``user_id``, ``item_to_index``,
``interactions``, ``user_items``, ``w_i`` ... ``v_if``, ``is_fit``), same methods and return types
(``fit``, ``fit_partial``, ``predict``, ``recommend``, ``similar_items``, ``similar_users``).  Host-side data
preparation is re-hosted (NumPy sorting instead of pandas ``groupby``), the native calls go to
``rankfm_b200._rankfm`` (ctypes -> CUDA) instead of the Cython module.  Weights live in NumPy arrays owned by the
model, exactly as in the reference, so pickling, attribute access and warm starts keep working.
"""
import numpy as np
import pandas as pd

from rankfm_b200 import _rankfm
from rankfm_b200._rankfm import _fit, _predict, _recommend, _similar, _similar_batch, UserItems
from rankfm_b200.utils import get_data, lookup_ids, unique_ids

_LOSSES = ('bpr', 'warp')
_SCHEDULES = ('constant', 'invscaling')


def _positive(value, kind):
    return isinstance(value, kind) and value > 0


class RankFM():
    """Factorization Machines for Ranking Problems with Implicit Feedback Data (B200-native hot path)"""

    def __init__(self, factors=10, loss='bpr', max_samples=10, alpha=0.01, beta=0.1, sigma=0.1, learning_rate=0.1,
                 learning_schedule='constant', learning_exponent=0.25):
        """hyper-parameters as in the reference (``rankfm.py:14-49``), validated with the same assertions"""
        assert isinstance(factors, int) and factors >= 1, "[factors] must be a positive integer"
        assert isinstance(loss, str) and loss in _LOSSES, "[loss] must be in ('bpr', 'warp')"
        assert _positive(max_samples, int), "[max_samples] must be a positive integer"
        assert _positive(alpha, float), "[alpha] must be a positive float"
        assert _positive(beta, float), "[beta] must be a positive float"
        assert _positive(sigma, float), "[sigma] must be a positive float"
        assert _positive(learning_rate, float), "[learning_rate] must be a positive float"
        assert isinstance(learning_schedule, str) and learning_schedule in _SCHEDULES, "[learning_schedule] must be in ('constant', 'invscaling')"
        assert _positive(learning_exponent, float), "[learning_exponent] must be a positive float"

        self.factors, self.loss, self.max_samples = factors, loss, max_samples
        self.alpha, self.beta, self.sigma = alpha, beta, sigma
        self.learning_rate, self.learning_schedule, self.learning_exponent = learning_rate, learning_schedule, learning_exponent
        self._reset_state()

    # ------------------------------------------------------------------ state ------------------------------------

    _STATE = ('user_id', 'item_id', 'user_idx', 'item_idx', 'index_to_user', 'index_to_item', 'user_to_index',
              'item_to_index', 'interactions', 'sample_weight', 'user_items', 'x_uf', 'x_if', 'w_i', 'w_if',
              'v_u', 'v_i', 'v_uf', 'v_if', '_prepared_stamp')

    def _reset_state(self):
        """clear every fitted attribute (``rankfm.py:60-97``)"""
        for name in self._STATE:
            setattr(self, name, None)
        self.is_fit = False

    @staticmethod
    def _check_interactions(interactions):
        assert isinstance(interactions, (np.ndarray, pd.DataFrame)), "[interactions] must be np.ndarray or pd.dataframe"
        assert interactions.shape[1] == 2, "[interactions] should be: [user_id, item_id]"

    @staticmethod
    def _lookup(ids, known):
        """index of each id in the unique array ``known`` (-1 when absent); any id dtype"""
        return lookup_ids(ids, known)

    def _lookup_known(self, ids, known):
        """`_lookup` for a long column of integer ids: the distinct ids and their positions come from one device radix sort
        (rfm_prep_index_ids), only the distinct ones are then matched against the model's (sorted) id list"""
        ids = np.asarray(ids)
        if len(ids) >= _rankfm._PREP_DEVICE_MIN and ids.dtype.kind in "iu" and known.dtype.kind in "iu" and _rankfm.device_count() > 0:
            distinct, where = _rankfm.prep_index_ids(ids)
            pos = np.searchsorted(known, distinct)
            pos = np.where((pos < len(known)) & (known[np.minimum(pos, len(known) - 1)] == distinct), pos, -1).astype(np.int64)
            return pos[where]
        return self._lookup(ids, known)

    def _init_all(self, interactions, user_features=None, item_features=None, sample_weight=None):
        """first fit: build the id <-> index maps, then interactions, features, weights (``rankfm.py:100-137``)"""
        self._check_interactions(interactions)
        raw = get_data(interactions)
        pairs = None
        if (len(raw) >= _rankfm._PREP_DEVICE_MIN and raw[:, 0].dtype.kind in "iu" and raw[:, 1].dtype.kind in "iu" and _rankfm.device_count() > 0):
            # integer ids, large input: sorted unique ids + the index of every id by device radix sort (rfm_prep_index_ids)
            # instead of np.unique + a hash lookup per column
            users, u_idx = _rankfm.prep_index_ids(raw[:, 0])
            items, i_idx = _rankfm.prep_index_ids(raw[:, 1])
            self.user_id = pd.Series(users.astype(raw[:, 0].dtype, copy=False))
            self.item_id = pd.Series(items.astype(raw[:, 1].dtype, copy=False))
            pairs = np.ascontiguousarray(np.stack([u_idx, i_idx], axis=1), dtype=np.int32)
        else:
            self.user_id = pd.Series(unique_ids(raw[:, 0]))     # sorted, like np.unique
            self.item_id = pd.Series(unique_ids(raw[:, 1]))
        self.index_to_user, self.index_to_item = self.user_id, self.item_id
        self.user_to_index = pd.Series(data=self.index_to_user.index, index=self.index_to_user.values)
        self.item_to_index = pd.Series(data=self.index_to_item.index, index=self.index_to_item.values)
        self.user_idx = np.arange(len(self.user_id), dtype=np.int32)
        self.item_idx = np.arange(len(self.item_id), dtype=np.int32)
        self._init_interactions(interactions, sample_weight, pairs=pairs)
        self._init_features(user_features, item_features)
        self._init_weights(user_features, item_features)

    def _init_interactions(self, interactions, sample_weight, pairs=None):
        """ids -> int32 index pairs, sample weights, per-user observed item sets (``rankfm.py:140-177``); ``pairs`` = the
        index pairs when the caller already has them (first fit on the device path)"""
        self._check_interactions(interactions)
        if pairs is None:
            raw = get_data(interactions)
            u = self._lookup_known(raw[:, 0], self.user_id.values)
            i = self._lookup_known(raw[:, 1], self.item_id.values)
            if (u < 0).any() or (i < 0).any():
                # the reference fails in `.astype(np.int32)` on the NaN produced by the id map (rankfm.py:154-155)
                raise ValueError("[interactions] contains user/item identifiers that were not present in the initial fit")
            pairs = np.ascontiguousarray(np.stack([u, i], axis=1), dtype=np.int32)

        if sample_weight is not None:
            assert isinstance(sample_weight, (np.ndarray, pd.Series)), "[sample_weight] must be np.ndarray or pd.series"
            assert sample_weight.ndim == 1, "[sample_weight] must a vector (ndim=1)"
            assert len(sample_weight) == len(interactions), "[sample_weight] must have the same length as [interactions]"
            self.sample_weight = np.ascontiguousarray(get_data(sample_weight), dtype=np.float32)
        else:
            self.sample_weight = np.ones(len(pairs), dtype=np.float32)

        n_users = len(self.user_idx)
        if self.is_fit and self.user_items is not None:
            # warm start: union of the old and the new item sets of every user (rankfm.py:170-172)
            old_ptr, old_idx = _rankfm.user_items_to_csr(self.user_items, n_users)
            old_users = np.repeat(np.arange(n_users, dtype=np.int64), np.diff(old_ptr))
            keys = np.concatenate([old_users * len(self.item_idx) + old_idx, pairs[:, 0].astype(np.int64) * len(self.item_idx) + pairs[:, 1]])
            keys = np.unique(keys)
            merged = np.stack([keys // len(self.item_idx), keys % len(self.item_idx)], axis=1)
            self.user_items = UserItems.from_interactions(merged, n_users, len(self.item_idx))
        else:
            self.user_items = UserItems.from_interactions(pairs, n_users, len(self.item_idx))
        self.interactions = pairs

    @staticmethod
    def _input_stamp(interactions, sample_weight):
        """identity of the training input: (address, shape, dtype, CRC32) of the interaction and weight buffers -- the whole
        buffer up to 64 MB, 2^20 strided samples beyond; None when the container has no single buffer to stamp"""
        import zlib
        out = []
        for obj in (interactions, sample_weight):
            if obj is None:
                out.append(None)
                continue
            a = get_data(obj)
            if not isinstance(a, np.ndarray) or a.dtype == object:
                return None
            flat = a.reshape(-1) if a.flags.c_contiguous else np.ascontiguousarray(a).reshape(-1)
            sample = flat if flat.nbytes <= (64 << 20) else np.ascontiguousarray(flat[::max(1, flat.size >> 20)])
            out.append((a.__array_interface__['data'][0], a.shape, a.dtype.str, zlib.crc32(sample.view(np.uint8))))
        return tuple(out)

    def _feature_matrix(self, features, known_ids, n_rows, what):
        """[id, f_1..f_n] table -> float32 matrix row-ordered by index (``rankfm.py:189-211``)"""
        if features is None:
            return np.zeros([n_rows, 1], dtype=np.float32)
        table = pd.DataFrame(features.copy())
        index = self._lookup(table.iloc[:, 0].values, known_ids)
        if len(index) != n_rows or not np.array_equal(np.sort(index), np.arange(n_rows)):
            raise KeyError('the {0}s in [{0}_features] do not match the {0}s in [interactions]'.format(what))
        ordered = table.iloc[np.argsort(index, kind='stable'), 1:]
        return np.ascontiguousarray(ordered, dtype=np.float32)       # ValueError for non-numeric columns

    def _init_features(self, user_features=None, item_features=None):
        self.x_uf = self._feature_matrix(user_features, self.user_id.values, len(self.user_idx), 'user')
        self.x_if = self._feature_matrix(item_features, self.item_id.values, len(self.item_idx), 'item')

    def _init_weights(self, user_features=None, item_features=None):
        """zeros for the scalar weights, N(0, sigma) factors, drawn from NumPy's global RNG in the reference's order
        (``rankfm.py:223-244``: v_u, v_i, then v_uf / v_if only when features are given) so that
        ``np.random.seed(s)`` reproduces the reference's initial state bit for bit"""
        n_users, n_items, F = len(self.user_idx), len(self.item_idx), self.factors
        self.w_i = np.zeros(n_items, dtype=np.float32)
        self.w_if = np.zeros(self.x_if.shape[1], dtype=np.float32)
        self.v_u = np.random.normal(loc=0, scale=self.sigma, size=(n_users, F)).astype(np.float32)
        self.v_i = np.random.normal(loc=0, scale=self.sigma, size=(n_items, F)).astype(np.float32)
        feature_scale = (self.alpha / self.beta) * self.sigma
        if user_features is not None:
            self.v_uf = np.random.normal(loc=0, scale=feature_scale, size=[self.x_uf.shape[1], F]).astype(np.float32)
        else:
            self.v_uf = np.zeros([self.x_uf.shape[1], F], dtype=np.float32)
        if item_features is not None:
            self.v_if = np.random.normal(loc=0, scale=feature_scale, size=[self.x_if.shape[1], F]).astype(np.float32)
        else:
            self.v_if = np.zeros([self.x_if.shape[1], F], dtype=np.float32)

    def _weights(self):
        return (self.x_uf, self.x_if, self.w_i, self.w_if, self.v_u, self.v_i, self.v_uf, self.v_if)

    # ------------------------------------------------------------------ public API -------------------------------

    def fit(self, interactions, user_features=None, item_features=None, sample_weight=None, epochs=1, verbose=False):
        """reset the model and train from scratch (``rankfm.py:252-266``); returns self"""
        self._reset_state()
        self.fit_partial(interactions, user_features, item_features, sample_weight, epochs, verbose)
        return self

    def fit_partial(self, interactions, user_features=None, item_features=None, sample_weight=None, epochs=1, verbose=False):
        """train, resuming from the current weights when already fit (``rankfm.py:269-327``); returns self"""
        assert isinstance(epochs, int) and epochs >= 1, "[epochs] must be a positive integer"
        assert isinstance(verbose, bool), "[verbose] must be a boolean value"

        if self.is_fit:
            # warm start (rankfm.py:269-327).  A loop of fit_partial() on the SAME interactions -- the usual way to train
            # with per-epoch evaluation -- would redo the id lookups and the user_items union every call (two hash joins
            # and a sort in the reference); identical input (same buffer, same shape, same CRC) keeps the prepared arrays,
            # which also lets the plug-in reuse its resident training session
            stamp = self._input_stamp(interactions, sample_weight)
            if stamp is None or stamp != getattr(self, "_prepared_stamp", None):
                self._init_interactions(interactions, sample_weight)
                self._prepared_stamp = stamp
            self._init_features(user_features, item_features)
        else:
            self._init_all(interactions, user_features, item_features, sample_weight)
            self._prepared_stamp = self._input_stamp(interactions, sample_weight)

        if self.loss == 'bpr':
            max_samples = 1                      # BPR == one negative per positive (rankfm.py:294-295)
        elif self.loss == 'warp':
            max_samples = self.max_samples
        else:
            raise ValueError('[loss] function not recognized')

        # in place on the model's arrays, like the reference's Cython call (rankfm.py:301-324)
        _fit(self.interactions, self.sample_weight, self.user_items, self.x_uf, self.x_if,
             self.w_i, self.w_if, self.v_u, self.v_i, self.v_uf, self.v_if,
             self.alpha, self.beta, self.learning_rate, self.learning_schedule, self.learning_exponent,
             max_samples, epochs, verbose)
        self.is_fit = True
        return self

    def predict(self, pairs, cold_start='nan'):
        """pointwise utilities of (user_id, item_id) pairs (``rankfm.py:330-364``)"""
        assert isinstance(pairs, (np.ndarray, pd.DataFrame)), "[pairs] must be np.ndarray or pd.dataframe"
        assert pairs.shape[1] == 2, "[pairs] should be: [user_id, item_id]"
        assert self.is_fit, "you must fit the model prior to generating predictions"

        raw = get_data(pairs)
        index_pairs = np.empty((len(raw), 2), dtype=np.float32)
        for col, known in ((0, self.user_id.values), (1, self.item_id.values)):
            idx = self._lookup(raw[:, col], known)
            index_pairs[:, col] = np.where(idx < 0, np.nan, idx)
        scores = _predict(index_pairs, *self._weights())

        if cold_start == 'nan':
            return scores
        elif cold_start == 'drop':
            return scores[~np.isnan(scores)]
        else:
            raise ValueError("param [cold_start] must be set to either 'nan' or 'drop'")

    def recommend(self, users, n_items=10, filter_previous=False, cold_start='nan'):
        """top-N item ids per user as a DataFrame indexed by user id (``rankfm.py:367-402``)"""
        assert getattr(users, '__iter__', False), "[users] must be an iterable (e.g. list, array, series)"
        assert self.is_fit, "you must fit the model prior to generating recommendations"

        idx = self._lookup(np.asarray(pd.Series(users).values), self.user_id.values)
        user_idx = np.ascontiguousarray(np.where(idx < 0, np.nan, idx), dtype=np.float32)
        rec_idx = _recommend(user_idx, self.user_items, n_items, filter_previous, *self._weights())
        # index -> id through the same pandas alignment rules as the reference (NaN index -> NaN id; dtype follows)
        rec_ids = self.index_to_item.reindex(rec_idx.ravel()).to_numpy().reshape(rec_idx.shape)
        rec_items = pd.DataFrame(rec_ids, index=users)

        if cold_start == 'nan':
            return rec_items
        elif cold_start == 'drop':
            return rec_items.dropna(how='any')
        else:
            raise ValueError("param [cold_start] must be set to either 'nan' or 'drop'")

    def _most_similar(self, which, index, n, index_to_id):
        n_rows = len(index_to_id)
        top = _similar(which, int(index), min(int(n), max(n_rows - 1, 1)), *self._weights())
        top = top[top >= 0][:max(int(n), 0)]
        return pd.Series(top).map(index_to_id).values

    def similar_items(self, item_id, n_items=10):
        """most similar items by latent inner product, the query excluded (``rankfm.py:405-428``)"""
        assert item_id in self.item_id.values, "you must select an [item_id] present in the training data"
        assert self.is_fit, "you must fit the model prior to generating similarities"
        return self._most_similar(0, self.item_to_index.loc[item_id], n_items, self.index_to_item)

    def similar_items_batch(self, item_ids=None, n_items=10):
        """``similar_items`` for many items in ONE call (default: every item) -> DataFrame indexed by item id, ``n_items``
        columns of item ids, most similar first (the batched all-items variant of ``rankfm.py:405-428``)"""
        assert self.is_fit, "you must fit the model prior to generating similarities"
        ids = self.item_id.values if item_ids is None else np.asarray(list(item_ids))
        index = self._lookup(ids, self.item_id.values)
        assert (index >= 0).all(), "you must select [item_ids] present in the training data"
        n = min(int(n_items), max(len(self.item_id) - 1, 1))
        top = _similar_batch(0, index.astype(np.int32), n, *self._weights())
        out = np.where(top >= 0, self.index_to_item.values[np.maximum(top, 0)], None) if (top < 0).any() else self.index_to_item.values[top]
        return pd.DataFrame(out, index=ids)

    def similar_users(self, user_id, n_users=10):
        """most similar users by latent inner product, the query excluded (``rankfm.py:431-454``)"""
        assert user_id in self.user_id.values, "you must select an [user_id] present in the training data"
        assert self.is_fit, "you must fit the model prior to generating similarities"
        return self._most_similar(1, self.user_to_index.loc[user_id], n_users, self.index_to_user)
