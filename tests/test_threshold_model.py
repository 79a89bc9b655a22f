"""CPU model of the row-threshold rules of the tensor-core recommend path (rfm_api.cu `tau_mode`).

Not a kernel test (those are in test_gpu_parity.py): a NumPy restatement of what pass 1 + the threshold select compute --
block bounds over a subset of the bias-ordered item tiles, the m-th largest as the row threshold -- used to check the CLAIMS
the design rests on: the head subset is as tight as the whole catalogue when item biases carry part of the ranking and
collects ~k n' candidates when they carry none (the case the fallback exists for); the sampled estimate with z = 4.5 sigma
of head room does not fall short and stays within a few n'.  The rank rule itself is the library's (`rfm_debug_tau_rank`).
"""
import numpy as np
import pytest

from rankfm_b200 import _lib

TILE, BLOCK = 128, 8


def candidates_per_row(I, F, rows, k, want, sig_b, mode, z=4.5, seed=0):
    """(threshold rank, rows that fell short, candidate counts) for `rows` random users over a random catalogue"""
    rng = np.random.default_rng(seed)
    V = rng.normal(0, 0.1, (I, F)).astype(np.float32)
    b = rng.normal(0, sig_b, I).astype(np.float32) if sig_b > 0 else np.zeros(I, np.float32)
    order = np.argsort(-b, kind='stable')                           # the catalogue in descending bias order
    V, b = V[order], b[order]
    S = rng.normal(0, 0.1, (rows, F)).astype(np.float32) @ V.T + b[None, :]
    n_tiles = I // TILE
    tiles = np.arange(0, (n_tiles + k - 1) // k) if mode == "head" else np.arange(0, n_tiles, k)
    cols = (tiles[:, None] * TILE + np.arange(TILE)[None, :]).ravel()
    m = want if mode == "head" else _lib.lib().rfm_debug_tau_rank(want, k, z)
    bounds = S[:, cols].reshape(rows, -1, BLOCK).max(axis=2)        # one bound per 8-item block of the subset
    tau = np.partition(bounds, -m, axis=1)[:, -m]
    counts = (S >= tau[:, None]).sum(axis=1)
    return m, int((counts < want).sum()), counts


@pytest.fixture(scope="module")
def lib():
    return _lib.lib()


def test_head_subset_is_tight_when_biases_carry_ranking_signal(lib):
    want = 116
    _, short, full = candidates_per_row(65536, 32, 48, 1, want, 0.3, "head")
    _, short32, head = candidates_per_row(65536, 32, 48, 16, want, 0.3, "head")
    assert short == 0 and short32 == 0                              # a provable bound never falls short
    # 1/16 of the tiles gives the threshold the whole catalogue gives (both pay the same price for one bound per 8-item
    # block: in bias order a row's best items share blocks)
    assert head.mean() <= 1.05 * full.mean() and head.max() <= 1.05 * full.max() and head.max() < 8 * want


def test_head_subset_is_loose_without_bias_signal(lib):
    """the documented failure mode: ~k n' candidates per row -> slots overflow -> flagged rows, conservative second serving"""
    want = 116
    _, short, head = candidates_per_row(65536, 32, 48, 16, want, 0.0, "head")
    assert short == 0
    assert head.mean() > 6 * want


def test_sampled_estimate_keeps_its_head_room(lib):
    want = 116
    for sig_b in (0.0, 0.1):
        m, short, est = candidates_per_row(65536, 32, 96, 8, want, sig_b, "estimate", z=4.5, seed=1)
        assert m < want and short == 0, (sig_b, m, short)
        assert est.mean() < 8 * want
    # without head room about every second row falls short: what the shortlist kernel's check (flag 2) is for
    _, short, _ = candidates_per_row(65536, 32, 96, 8, want, 0.0, "estimate", z=0.001, seed=1)
    assert short > 10
