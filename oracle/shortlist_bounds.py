"""CPU restatement (NumPy, float32) of the candidate-selection arithmetic of the tensor-core recommend path
(``rankfm_b200/csrc/rfm_gemm.cu``: ``score_filter_kernel`` passes 1 and 2, ``row_threshold_reg_kernel``).

TEST INFRASTRUCTURE ONLY (see the header of ``rankfm_oracle.c``): nothing in ``rankfm_b200`` imports this module.  It
exists to check, on the CPU and for adversarial inputs, the one property the GPU path rests on: **every item among a
row's n' best scores passes the pass-2 gate**, whatever subset of the catalogue pass 1 looked at, with every float32
rounding the kernels perform.  The reference has no counterpart -- it scores and sorts every item
(``_rankfm.pyx:440-456``); the final ranking of the GPU path is exact fp32 either way, this is about never losing a
candidate before that.

Layout mirrored from the kernels: positions are items in DESCENDING bias order, tiles of 128 positions, chunks of 32,
blocks of 8; padded positions carry dot = 0 and bias = -1e30 (so real scores are assumed to stay above -1e30: see
``test_padding_sentinel_bounds_the_score_range``).
"""
import numpy as np

TILE, CHUNK, BLOCK = 128, 32, 8
PAD_BIAS = np.float32(-1e30)
SLACK = np.float32(4.8e-7)


def ord_key(x):
    """monotone uint32 image of float32 (``ord_key`` in rfm_gemm.cu)"""
    b = np.asarray(x, np.float32).view(np.uint32)
    return np.where(b & np.uint32(0x80000000), ~b, b | np.uint32(0x80000000)).astype(np.uint32)


def key_to_float(k):
    k = np.asarray(k, np.uint32)
    return np.where(k & np.uint32(0x80000000), k & np.uint32(0x7fffffff), ~k).astype(np.uint32).view(np.float32)


def bias_order(bias):
    """position -> item, stable descending (CUB SortPairsDescending is stable: equal biases stay in item order)"""
    return np.argsort(-np.asarray(bias, np.float32), kind="stable")


def pad(dot_sorted, bias_sorted):
    n = len(dot_sorted)
    n_pad = (n + TILE - 1) // TILE * TILE
    d = np.zeros(n_pad, np.float32); d[:n] = dot_sorted
    b = np.full(n_pad, PAD_BIAS, np.float32); b[:n] = bias_sorted
    return d, b


def block_bounds(dot, bias, fraction=1, head=True):
    """pass 1: max(dot over an 8-item block) + the block's smallest (= last) bias, for the visited tiles"""
    n_tiles = len(dot) // TILE
    visited = (n_tiles + fraction - 1) // fraction
    tiles = np.arange(visited) if head else np.arange(visited) * fraction
    idx = (tiles[:, None] * TILE + np.arange(TILE)[None, :]).ravel()
    m = dot[idx].reshape(-1, BLOCK).max(axis=1)
    return (m + bias[idx].reshape(-1, BLOCK)[:, BLOCK - 1]).astype(np.float32)


def row_threshold(bounds, want):
    """n'-th largest bound, lowered to the edge of its 24-bit radix bucket (three 8-bit passes); -inf if too few bounds"""
    if want > len(bounds):
        return np.float32(-np.inf)
    k = np.sort(ord_key(bounds))[len(bounds) - want]
    return key_to_float(np.uint32(k) & np.uint32(0xffffff00))[()]


def pass2_gate(dot, bias, tau):
    """boolean mask of the positions pass 2 appends: dot >= (tau - largest bias of the 32-item chunk) - slack"""
    bmax = bias.reshape(-1, CHUNK)[:, 0]
    tau = np.float32(tau)
    with np.errstate(over="ignore", invalid="ignore"):
        thr = ((tau - bmax).astype(np.float32) - (SLACK * (np.abs(tau) + np.abs(bmax)).astype(np.float32)).astype(np.float32)).astype(np.float32)
    return dot >= np.repeat(thr, CHUNK)


def candidates(dot_by_item, bias_by_item, want, fraction=1, head=True):
    """(candidate item ids, tau) for one user row given its bf16-GEMM dot products and the fp32 biases"""
    order = bias_order(bias_by_item)
    dot, bias = pad(np.asarray(dot_by_item, np.float32)[order], np.asarray(bias_by_item, np.float32)[order])
    tau = row_threshold(block_bounds(dot, bias, fraction, head), want)
    gate = pass2_gate(dot, bias, tau)
    gate[len(order):] = False                                            # padded positions carry no item
    return order[np.flatnonzero(gate[:len(order)])], tau
