"""Synthetic workloads of the shapes BASELINE.json names (no datasets are reachable offline).

(user, item) pairs with Zipf-distributed popularity on both sides, de-duplicated, ids permuted so popularity is not
index-ordered, then re-indexed to the observed uniques like ``rankfm.py:115-116``.
"""
import numpy as np

CONFIGS = {
    # name: U, I, N, factors, loss, max_samples, epochs, P, Q   (BASELINE.json names epochs only for configs[0..1]; the large
    # configurations run 4 epochs per bench step)
    "cfg1": dict(U=10_000, I=5_000, N=100_000, F=16, loss="bpr", max_samples=1, epochs=5, P=0, Q=0,
                 label="synthetic 10k users x 5k items, 100k interactions, factors=16, bpr, 5 epochs"),
    "cfg2": dict(U=6_040, I=3_706, N=1_000_000, F=20, loss="warp", max_samples=20, epochs=20, P=0, Q=0,
                 label="MovieLens-1M shape synthetic (6040x3706, 1M interactions), factors=20, warp, max_samples=20, 20 epochs"),
    "cfg3": dict(U=1_000_000, I=200_000, N=50_000_000, F=64, loss="warp", max_samples=10, epochs=4, P=8, Q=8,
                 label="1M users x 200k items, 50M Zipf interactions, factors=64, warp, 8+8 side features"),
    "cfg3m": dict(U=250_000, I=50_000, N=8_000_000, F=64, loss="warp", max_samples=10, epochs=3, P=8, Q=8,
                  label="1/6 slice of cfg3: 250k users x 50k items, 8M interactions, factors=64, warp max_samples=10, 8+8 dense side features"),
    "cfg3n": dict(U=1_000_000, I=200_000, N=16_000_000, F=64, loss="warp", max_samples=10, epochs=3, P=0, Q=0,
                  label="cfg3 shape without side features: 1M users x 200k items, 16M interactions, factors=64, warp max_samples=10 (tables 308 MB > L2)"),
    "cfg4m": dict(U=1_250_000, I=1_000_000, N=16_000_000, F=128, loss="bpr", max_samples=1, epochs=3, P=0, Q=0,
                  label="DRAM-resident slice of cfg4: 1.25M users x 1M items, 16M interactions, factors=128, bpr (tables 1.2 GB >> L2)"),
    "cfg4s": dict(U=1_250_000, I=1_000_000, N=62_500_000, F=128, loss="bpr", max_samples=1, epochs=4, P=0, Q=0,
                  label="per-GPU shard of 10M users x 1M items, 500M interactions, factors=128, bpr (1/8 of cfg4)"),
}


def zipf_interactions(U, I, N, seed=42, a_u=0.6, a_i=1.0, offset_users=0, reindex=True):
    """int32 [n,2] unique (user,item) pairs in random order, n == N whenever enough unique pairs exist"""
    rng = np.random.default_rng(seed)
    pu = 1.0 / np.arange(1, U + 1) ** a_u
    pi = 1.0 / np.arange(1, I + 1) ** a_i
    cu, ci = np.cumsum(pu / pu.sum()), np.cumsum(pi / pi.sum())
    keys = np.zeros(0, dtype=np.int64)
    draw = int(N * 1.25) + 1024
    for _ in range(8):
        u = np.minimum(np.searchsorted(cu, rng.random(draw, dtype=np.float32).astype(np.float64)), U - 1)
        i = np.minimum(np.searchsorted(ci, rng.random(draw, dtype=np.float32).astype(np.float64)), I - 1)
        k = u.astype(np.int64) * I + i
        k.sort()
        k = k[np.concatenate([[True], k[1:] != k[:-1]])]
        keys = k if len(keys) == 0 else np.union1d(keys, k)
        if len(keys) >= N:
            break
    order = rng.permutation(len(keys))[:N]                 # unbiased trim + final shuffle in one pass
    keys = keys[order]
    u, i = keys // I, keys % I
    u = rng.permutation(U)[u]                               # popularity must not be index-ordered
    i = rng.permutation(I)[i]
    for col, n in ((u, U), (i, I)) if reindex else ():      # re-index to the observed uniques (rankfm.py:115-116)
        present = np.zeros(n, dtype=bool)
        present[col] = True
        col[:] = (np.cumsum(present) - 1)[col]
    return np.ascontiguousarray(np.stack([u + offset_users, i], axis=1).astype(np.int32))


def zipf_interactions_device(U, I, N, seed=42, a_u=0.6, a_i=1.0, offset_users=0, device=0, perm_seed=None, reindex=True):
    """the same workload drawn, de-duplicated, trimmed, permuted and re-indexed on the GPU (``rfm_synth_zipf``): a second or
    two for the 50-62 M-interaction configurations instead of 50-80 s of NumPy.  Same distribution, different random
    stream.  -> (int32 [n,2], observed users, observed items)"""
    import ctypes as C
    from rankfm_b200 import _lib
    out = np.empty((N, 2), dtype=np.int32)
    n, nu, ni = C.c_int64(), C.c_int32(), C.c_int32()
    _lib.check(_lib.lib().rfm_synth_zipf(int(U), int(I), int(N), float(a_u), float(a_i), int(seed), int(seed if perm_seed is None else perm_seed),
                                         int(bool(reindex)), int(offset_users), int(device),
                                         _lib.ptr(out), C.byref(n), C.byref(nu), C.byref(ni)))
    return out[:n.value], nu.value, ni.value


def init_weights(U, I, F, P=0, Q=0, seed=0, sigma=0.1, alpha=0.01, beta=0.1):
    """N(0, sigma) factors like ``rankfm.py:223-244`` (own generator instead of NumPy's global state)"""
    rng = np.random.default_rng(seed)
    scale = (alpha / beta) * sigma
    return dict(w_i=np.zeros(I, np.float32), w_if=np.zeros(max(Q, 1), np.float32),
                v_u=rng.normal(0, sigma, (U, F)).astype(np.float32), v_i=rng.normal(0, sigma, (I, F)).astype(np.float32),
                v_uf=(rng.normal(0, scale, (P, F)) if P else np.zeros((1, F))).astype(np.float32),
                v_if=(rng.normal(0, scale, (Q, F)) if Q else np.zeros((1, F))).astype(np.float32))


def side_features(U, I, P, Q, seed=0):
    rng = np.random.default_rng(seed + 1000)
    x_uf = rng.random((U, P), dtype=np.float32) if P else np.zeros((U, 1), np.float32)
    x_if = rng.random((I, Q), dtype=np.float32) if Q else np.zeros((I, 1), np.float32)
    return x_uf, x_if
