"""Host-prep timing at cfg3's size: id maps + user_items for 50 M interactions, device path vs NumPy path.
usage: python profiles/tools/prep_timing.py [n_interactions]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from rankfm_b200 import _rankfm  # noqa: E402
from rankfm_b200.rankfm import RankFM  # noqa: E402
from rankfm_b200.synthetic import zipf_interactions_device  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
U, I = 1_000_000, 200_000
X, nu, ni = zipf_interactions_device(U, I, N, seed=42)
rng = np.random.default_rng(0)
uid = np.sort(rng.choice(10**10, nu, replace=False))
iid = np.sort(rng.choice(10**9, ni, replace=False))
raw = np.ascontiguousarray(np.stack([uid[X[:, 0]], iid[X[:, 1]]], axis=1))          # raw int64 (user_id, item_id) pairs
print("interactions %d, users %d, items %d" % (len(raw), nu, ni))


def timed(label, fn, repeat=2):
    best = 1e9
    for _ in range(repeat):
        t0 = time.perf_counter(); out = fn(); best = min(best, time.perf_counter() - t0)
    print("%-70s %.3f s" % (label, best))
    return out


timed("device: sorted unique user ids + index of every id (rfm_prep_index_ids)", lambda: _rankfm.prep_index_ids(raw[:, 0]))
timed("device: same for item ids", lambda: _rankfm.prep_index_ids(raw[:, 1]))
timed("device: user_items CSR (rfm_prep_user_items)", lambda: _rankfm.prep_user_items(X, nu, ni))
m = RankFM(factors=64, loss='warp')


def init_all(no_weights):
    mm = RankFM(factors=64, loss='warp')
    if no_weights:
        mm._init_weights = lambda *a, **k: None
    mm._init_all(raw)
    return mm


timed("RankFM._init_all WITHOUT the weight draws (id maps, index pairs, user_items, pandas maps)", lambda: init_all(True))
timed("RankFM._init_all (with np.random.normal weights in the reference's draw order)", lambda: init_all(False), repeat=1)
if N <= 10_000_000:
    timed("host path: UserItems.from_interactions_host", lambda: _rankfm.UserItems.from_interactions_host(X, nu), repeat=1)
