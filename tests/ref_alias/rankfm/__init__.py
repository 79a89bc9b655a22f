"""`rankfm` alias package (TEST INFRASTRUCTURE): lets the reference's own test suite -- `tests/test_rankfm.py`, which does
`from rankfm.rankfm import RankFM` / `from rankfm.evaluation import ...` (reference `tests/test_rankfm.py:5-6`) -- run
UNCHANGED against rankfm_b200.  tests/test_reference_suite.py puts this directory on PYTHONPATH.

RANKFM_ALIAS_BACKEND=oracle routes the three native calls of the class to the CPU oracle (host-logic coverage on machines
without a GPU, like the `backend` fixture of tests/test_api_contract.py); the default is the real ctypes -> CUDA path."""
import os as _os

import rankfm_b200.rankfm as _impl

if _os.environ.get("RANKFM_ALIAS_BACKEND") == "oracle":
    import numpy as _np
    from oracle import oracle as _oracle

    def _oracle_similar(which, index, n, x_uf, x_if, w_i, w_if, v_u, v_i, v_uf, v_if):
        rep = (v_i + x_if @ v_if) if which == 0 else (v_u + x_uf @ v_uf)
        sims = rep @ rep[index]
        return _np.array([k for k in _np.argsort(-sims, kind='stable') if k != index][:n], dtype=_np.int32)

    _impl._fit, _impl._predict, _impl._recommend, _impl._similar = _oracle._fit, _oracle._predict, _oracle._recommend, _oracle_similar
