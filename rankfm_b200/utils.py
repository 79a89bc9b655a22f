"""general utility functions (API mirror of the reference's ``rankfm/utils.py:5-18``)"""


def get_data(obj):
    """numeric data of a DataFrame / Series / ndarray"""
    kind = obj.__class__.__name__
    if kind in ('DataFrame', 'Series'):
        return obj.values
    if kind == 'ndarray':
        return obj
    raise TypeError("input data must be in either pd.dataframe/pd.series or np.ndarray format")
