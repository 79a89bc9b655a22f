"""Worker for the multi-process tests (launched once per rank by test_multigpu.py or by torchrun).

mode "gloo-oracle" (CPU): every rank trains ITS user shard with the CPU oracle from identical initial weights, the ranks
  then combine exactly like rfm_session_train does on GPUs -- per epoch a gain-weighted SUM of item-table deltas over the
  process group; user rows are owned by one rank and never exchanged during training -- and the result
  must be identical on every rank and equal to the single-process emulation of the same schedule.
mode "gpu" (GPU): the real thing: one process per GPU, `_fit` with world>1 on the resident training session -- fused
  peer-memory exchange by default, RANKFM_B200_EXCHANGE=nccl for the ncclAllReduce fallback -- then the same invariants,
  agreement with a single-GPU fit within Hogwild tolerances, and a hold-out hit-rate gate (|delta| < 0.02 vs one GPU).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import WEIGHTS, CSRItems, csr_of, features, init_weights, zipf_interactions  # noqa: E402


def problem():
    X = zipf_interactions(600, 300, 20000, seed=3)
    U, I = int(X[:, 0].max()) + 1, int(X[:, 1].max()) + 1
    indptr, indices = csr_of(X, U)
    return X, U, I, CSRItems(indptr, indices), np.ones(len(X), np.float32)


def fold_gain(x, C):
    """gain of each replica's delta when C replicas of a parameter are folded; x = -log(contraction over the epoch).
    Mirrors fold_gain_dev (rankfm_b200/csrc/rfm_comm.cu) / item_delta_kernel (rfm_api.cu)."""
    x = np.asarray(x, np.float64)
    g = np.ones_like(x)
    m = x > 1e-4
    g[m] = (1.0 - np.exp(-x[m] * C)) / (C * (1.0 - np.exp(-x[m])))
    return g.astype(np.float32)


def item_gains(X_shard, I, draws, eta, alpha, C):
    touch = np.bincount(X_shard[:, 1], minlength=I).astype(np.float64) + draws / I
    return fold_gain(eta * (2 * alpha + 0.01) * touch, C), fold_gain(eta * (2 * alpha + 0.15) * touch, C)


def allreduce_sum(a):
    t = torch.from_numpy(np.ascontiguousarray(a))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.numpy()


def gloo_oracle(rank, world):
    from oracle import oracle
    from rankfm_b200 import _rankfm
    X, U, I, ui, sw = problem()
    F, epochs = 8, 3
    x_uf, x_if = features(U, I, 0, 0)
    Xr, swr, (lo, hi) = _rankfm.shard_by_user(X, sw, U, rank, world)
    w = init_weights(U, I, F, seed=1)
    w0_vu = w['v_u'].copy()
    hyper = (0.01, 0.1, 0.1, 'constant', 0.25, 1)
    for e in range(epochs):
        snap = {k: w[k].copy() for k in ('w_i', 'v_i')}
        out = oracle.fit_ex(Xr, swr, ui, x_uf, x_if, *[w[k] for k in WEIGHTS], *hyper, 1, perms=None, sampler="philox", seed=5, epoch_offset=e, max_rejects=64)
        g_f, g_b = item_gains(Xr, I, float(out['draws'][0]), 0.1, 0.01, world)
        # replicated item side: gain-weighted sum of the replicas' deltas (rarely touched rows add up, hot rows average)
        w['v_i'][...] = snap['v_i'] + allreduce_sum(g_f[:, None] * (w['v_i'] - snap['v_i']))
        w['w_i'][...] = snap['w_i'] + allreduce_sum(g_b * (w['w_i'] - snap['w_i']))
    # user rows: owned by exactly one rank, no collective during training; rows outside [lo, hi) are untouched here
    assert np.array_equal(w['v_u'][:lo], w0_vu[:lo]) and np.array_equal(w['v_u'][hi:], w0_vu[hi:])
    assert not np.array_equal(w['v_u'][lo:hi], w0_vu[lo:hi])
    _rankfm.allgather_user_rows(w['v_u'], (lo, hi))              # control plane: assemble the full table on every rank
    # every rank must now hold the same model
    for k in ('w_i', 'v_i', 'v_u'):
        ref = allreduce_sum(w[k].astype(np.float64)) / world
        assert np.allclose(w[k], ref, rtol=0, atol=1e-6), k
    if rank == 0:
        # single-process emulation of the same schedule
        shards = [_rankfm.shard_by_user(X, sw, U, r, world) for r in range(world)]
        ws = init_weights(U, I, F, seed=1)
        for e in range(epochs):
            base = {k: ws[k].copy() for k in WEIGHTS}
            acc = {k: np.zeros_like(ws[k]) for k in ('w_i', 'v_i', 'v_u')}
            for Xs, sws, _ in shards:
                wr = {k: base[k].copy() for k in WEIGHTS}
                out = oracle.fit_ex(Xs, sws, ui, x_uf, x_if, *[wr[k] for k in WEIGHTS], *hyper, 1, perms=None, sampler="philox", seed=5, epoch_offset=e, max_rejects=64)
                g_f, g_b = item_gains(Xs, I, float(out['draws'][0]), 0.1, 0.01, world)
                acc['v_i'] += g_f[:, None] * (wr['v_i'] - base['v_i'])
                acc['w_i'] += g_b * (wr['w_i'] - base['w_i'])
                acc['v_u'] += wr['v_u'] - base['v_u']
            for k in acc:
                ws[k][...] = base[k] + acc[k]
        for k in ('w_i', 'v_i', 'v_u'):
            assert np.allclose(w[k], ws[k], rtol=1e-5, atol=1e-6), k
        print("gloo-oracle ok")


def hit_rate(w, X_train_ui, X_test, U, I, k=10):
    """share of test users with a held-out item among their top-k unseen items (rankfm/evaluation.py:9-33)"""
    from rankfm_b200 import _rankfm
    x_uf, x_if = features(U, I, 0, 0)
    users = np.unique(X_test[:, 0])
    rec = _rankfm._recommend(users.astype(np.float32), X_train_ui, k, True, x_uf, x_if, *[w[n] for n in WEIGHTS])
    held = {}
    for u, i in X_test:
        held.setdefault(int(u), set()).add(int(i))
    return float(np.mean([len(held[int(u)] & set(r[~np.isnan(r)].astype(int).tolist())) > 0 for u, r in zip(users, rec)]))


def gpu(rank, world):
    from rankfm_b200 import _rankfm
    local = int(os.environ.get("LOCAL_RANK", rank))
    _rankfm.set_device(local)
    idt = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        idt = torch.frombuffer(bytearray(_rankfm.nccl_unique_id()), dtype=torch.uint8).clone()
    dist.broadcast(idt, src=0)
    want_path = 2 if os.environ.get("RANKFM_B200_EXCHANGE") == "nccl" else 1

    # ---- invariants on a small problem, BPR ----
    X, U, I, ui, sw = problem()
    F, epochs = 8, 4
    x_uf, x_if = features(U, I, 0, 0)
    hyper = (0.01, 0.1, 0.1, 'constant', 0.25, 1)
    Xr, swr, (lo, hi) = _rankfm.shard_by_user(X, sw, U, rank, world)
    w = init_weights(U, I, F, seed=1)
    w0_vu = w['v_u'].copy()
    _rankfm.set_seed(5)
    _rankfm.set_comm(rank, world, idt.numpy().tobytes(), user_range=(lo, hi))
    _rankfm._fit(Xr, swr, ui, x_uf, x_if, *[w[k] for k in WEIGHTS], *hyper, epochs, False)
    stats = _rankfm.last_stats
    path = _rankfm._training["sess"].exchange_path()
    assert path == want_path, "exchange path %d, wanted %d" % (path, want_path)
    # a second call on the same data reuses the resident session and the communicator (warm start, fit_partial)
    w2 = {k: w[k].copy() for k in WEIGHTS}
    builds = _rankfm._training["builds"]
    _rankfm._fit(Xr, swr, ui, x_uf, x_if, *[w2[k] for k in WEIGHTS], *hyper, 1, False)
    assert _rankfm._training["builds"] == builds and _rankfm._training["hits"] >= 1
    # item side identical on every rank after every call; user rows: only the owned ones moved
    for k in ('w_i', 'v_i'):
        ref = allreduce_sum(w[k].astype(np.float64)) / world
        assert np.allclose(w[k], ref, rtol=0, atol=1e-6), "ranks disagree on " + k
    assert np.array_equal(w['v_u'][:lo], w0_vu[:lo]) and np.array_equal(w['v_u'][hi:], w0_vu[hi:])
    assert not np.array_equal(w['v_u'][lo:hi], w0_vu[lo:hi])
    _rankfm.allgather_user_rows(w['v_u'], (lo, hi))
    ll = np.array([s['log_likelihood'] for s in stats])                 # the library reports whole-job sums
    assert all(s['sync_ms'] > 0 for s in stats)
    _rankfm.set_comm(0, 1, None)
    if rank == 0:
        w1 = init_weights(U, I, F, seed=1)
        s1 = _rankfm.fit_ex(X, sw, ui, x_uf, x_if, *[w1[k] for k in WEIGHTS], *hyper, epochs, mode="production", seed=5)
        ll1 = np.array([s['log_likelihood'] for s in s1])
        assert np.allclose(ll[1:], ll1[1:], rtol=0.05), (ll, ll1)
        for k in ('v_u', 'v_i', 'w_i'):
            assert abs(np.linalg.norm(w[k]) / np.linalg.norm(w1[k]) - 1) < 0.1, k
    dist.barrier()

    # ---- quality gate: hold-out hit rate of the N-GPU model vs the single-GPU model (WARP, every item hot) ----
    Xq = zipf_interactions(3000, 1200, 260000, seed=11)
    Uq, Iq = int(Xq[:, 0].max()) + 1, int(Xq[:, 1].max()) + 1
    rng = np.random.default_rng(0)
    test_mask = rng.random(len(Xq)) < 0.1
    Xtr, Xte = np.ascontiguousarray(Xq[~test_mask]), Xq[test_mask]
    uiq = CSRItems(*csr_of(Xtr, Uq))
    swq = np.ones(len(Xtr), np.float32)
    xq_uf, xq_if = features(Uq, Iq, 0, 0)
    hyper_q = (0.01, 0.1, 0.1, 'invscaling', 0.25, 10)
    Fq, epochs_q = 20, 12
    Xs, sws, (qlo, qhi) = _rankfm.shard_by_user(Xtr, swq, Uq, rank, world)
    wq = init_weights(Uq, Iq, Fq, seed=2)
    _rankfm.set_comm(rank, world, idt.numpy().tobytes(), user_range=(qlo, qhi))
    _rankfm._fit(Xs, sws, uiq, xq_uf, xq_if, *[wq[k] for k in WEIGHTS], *hyper_q, epochs_q, False)
    stats_q = _rankfm.last_stats
    _rankfm.allgather_user_rows(wq['v_u'], (qlo, qhi))
    _rankfm.set_comm(0, 1, None)
    if rank == 0:
        w1 = init_weights(Uq, Iq, Fq, seed=2)
        s1 = _rankfm.fit_ex(Xtr, swq, uiq, xq_uf, xq_if, *[w1[k] for k in WEIGHTS], *hyper_q, epochs_q, mode="production", seed=5)
        hr_n, hr_1 = hit_rate(wq, uiq, Xte, Uq, Iq), hit_rate(w1, uiq, Xte, Uq, Iq)
        ll_n, ll_1 = stats_q[-1]['log_likelihood'], s1[-1]['log_likelihood']
        print("quality: world=%d hit_rate@10 %.4f vs single-GPU %.4f; final log-likelihood %.1f vs %.1f; exchange %.3f ms/epoch (path %d)"
              % (world, hr_n, hr_1, ll_n, ll_1, float(np.mean([s['sync_ms'] for s in stats_q])), path))
        assert abs(hr_n - hr_1) < 0.02, (hr_n, hr_1)
        assert abs(ll_n / ll_1 - 1) < 0.05, (ll_n, ll_1)
        print("gpu ok", ll.tolist())
    dist.barrier()
    _rankfm.release_comms()


if __name__ == "__main__":
    mode = sys.argv[1]
    dist.init_process_group(backend="gloo", init_method="env://")
    try:
        (gloo_oracle if mode == "gloo-oracle" else gpu)(dist.get_rank(), dist.get_world_size())
    finally:
        dist.destroy_process_group()
