"""alias of the reference's `rankfm/evaluation.py` module name (see the package docstring)"""
from rankfm_b200.evaluation import (hit_rate, reciprocal_rank, discounted_cumulative_gain, precision, recall,  # noqa: F401
                                    diversity)
