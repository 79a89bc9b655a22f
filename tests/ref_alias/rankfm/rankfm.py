"""alias of the reference's `rankfm/rankfm.py` module name (see the package docstring)"""
from rankfm_b200.rankfm import RankFM  # noqa: F401
