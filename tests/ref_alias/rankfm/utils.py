"""alias of the reference's `rankfm/utils.py` module name (see the package docstring)"""
from rankfm_b200.utils import get_data  # noqa: F401
