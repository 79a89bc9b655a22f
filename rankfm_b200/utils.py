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
