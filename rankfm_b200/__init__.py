"""rankfm_b200 -- B200-native (sm_100a CUDA) implementation of the RankFM hot path behind the reference's API.

``from rankfm_b200 import RankFM`` is the drop-in for ``from rankfm.rankfm import RankFM``;
``rankfm_b200._rankfm`` is the drop-in for the reference's Cython module ``rankfm._rankfm``.
"""
__version__ = "0.1.0"


def __getattr__(name):
    if name == "RankFM":
        from rankfm_b200.rankfm import RankFM
        return RankFM
    raise AttributeError(name)
