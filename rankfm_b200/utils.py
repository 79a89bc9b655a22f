"""Small helpers shared by the host-side API mirror (`rankfm_b200.rankfm`, `rankfm_b200.evaluation`)."""
import numpy as np
import pandas as pd

_PANDAS_CONTAINERS = (pd.DataFrame, pd.Series)


def get_data(obj):
    """Return the ndarray behind `obj`.

    Accepts the three containers the reference API accepts (DataFrame, Series, ndarray; `rankfm/utils.py:5-18`);
    pandas objects hand back `.values`, an ndarray is returned untouched, anything else is a `TypeError`.
    """
    if isinstance(obj, _PANDAS_CONTAINERS):
        return obj.values
    if isinstance(obj, np.ndarray):
        return obj
    raise TypeError("input data must be in either pd.dataframe/pd.series or np.ndarray format")


def unique_ids(ids):
    """sorted unique identifiers (`np.unique`), with a linear-time path for integer ids of moderate range"""
    ids = np.asarray(ids)
    if ids.dtype.kind in "iu" and ids.size:
        lo, hi = int(ids.min()), int(ids.max())
        if hi - lo <= max(8 * ids.size, 1 << 20):
            present = np.zeros(hi - lo + 1, dtype=bool)
            present[ids - lo] = True
            return (np.flatnonzero(present) + lo).astype(ids.dtype)
    return np.unique(ids)


def lookup_ids(ids, known):
    """index of each id in the sorted unique array `known` (-1 when absent), any id dtype.

    Integer ids of moderate range go through a dense table (two linear passes); everything else through a pandas hash
    index, like the `pd.Series` maps of the reference (`rankfm/rankfm.py:118-128`)."""
    ids, known = np.asarray(ids), np.asarray(known)
    if ids.dtype.kind in "iu" and known.dtype.kind in "iu" and known.size:
        lo, hi = int(known[0]), int(known[-1])
        if hi - lo <= max(8 * known.size, 1 << 20):
            table = np.full(hi - lo + 1, -1, dtype=np.int64)
            table[known - lo] = np.arange(known.size, dtype=np.int64)
            ids64 = ids.astype(np.int64)
            inside = (ids64 >= lo) & (ids64 <= hi)
            out = np.full(ids.shape, -1, dtype=np.int64)
            out[inside] = table[ids64[inside] - lo]
            return out
    return pd.Index(known).get_indexer(pd.Index(ids)).astype(np.int64)
