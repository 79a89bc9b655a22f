"""Mint golden input/output vectors from the UNMODIFIED reference (compiled into oracle/_ref by oracle/build_ref.py).

Run here (the container that has /root/reference):   python tests/golden/make_golden.py
The .npz files it writes are committed; the GPU box and CI only read them.  The reference's own tests hold no
numeric vectors for this path, so these are the pins for the oracle and the CUDA kernels.

Every case stores: the inputs handed to the reference's `_fit` (`rankfm/_rankfm.pyx:122-142`), the row order of each
epoch (what `np.random.shuffle` produced under the stated seed, `:227`), the trained weights, `_predict` scores
(`:345-390`) and `_recommend` outputs (`:393-460`) of the reference.
"""
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
WEIGHTS = ('w_i', 'w_if', 'v_u', 'v_i', 'v_uf', 'v_if')


def synth(U, I, N, F, P, Q, seed, sw_unit=True):
    rng = np.random.default_rng(seed)
    pu = 1.0 / np.arange(1, U + 1) ** 0.6
    pi = 1.0 / np.arange(1, I + 1) ** 1.0
    X = np.stack([rng.choice(U, N, p=pu / pu.sum()), rng.choice(I, N, p=pi / pi.sum())], 1)
    X = np.concatenate([X, np.stack([np.arange(U), rng.integers(0, I, U)], 1), np.stack([rng.integers(0, U, I), np.arange(I)], 1)])
    X = np.unique(X, axis=0).astype(np.int32)
    rng.shuffle(X)
    N = len(X)
    order = np.lexsort((X[:, 1], X[:, 0]))
    counts = np.bincount(X[:, 0], minlength=U)
    indptr = np.zeros(U + 1, np.int64); np.cumsum(counts, out=indptr[1:])
    indices = X[order, 1].astype(np.int32)
    sw = np.ones(N, np.float32) if sw_unit else rng.uniform(0.5, 1.5, N).astype(np.float32)
    x_uf = rng.uniform(0, 1, (U, P)).astype(np.float32) if P else np.zeros((U, 1), np.float32)
    x_if = rng.uniform(0, 1, (I, Q)).astype(np.float32) if Q else np.zeros((I, 1), np.float32)
    if P: x_uf[rng.uniform(size=x_uf.shape) < 0.3] = 0
    if Q: x_if[rng.uniform(size=x_if.shape) < 0.3] = 0
    w = dict(w_i=np.zeros(I, np.float32), w_if=np.zeros(max(Q, 1), np.float32),
             v_u=rng.normal(0, .1, (U, F)).astype(np.float32), v_i=rng.normal(0, .1, (I, F)).astype(np.float32),
             v_uf=(rng.normal(0, .01, (P, F)) if P else np.zeros((1, F))).astype(np.float32),
             v_if=(rng.normal(0, .01, (Q, F)) if Q else np.zeros((1, F))).astype(np.float32))
    return X, sw, indptr, indices, x_uf, x_if, w


def kernel_case(name, ref, U, I, N, F, P, Q, max_samples, epochs, schedule, seed, sw_unit=True, alpha=0.01, beta=0.1, lr=0.1, expo=0.25):
    X, sw, indptr, indices, x_uf, x_if, w0 = synth(U, I, N, F, P, Q, seed, sw_unit)
    ui = {u: indices[indptr[u]:indptr[u + 1]] for u in range(U)}
    np.random.seed(seed)
    idx = np.arange(len(X), dtype=np.int32)
    perms = np.empty((epochs, len(X)), np.int32)
    for e in range(epochs):
        np.random.shuffle(idx); perms[e] = idx
    w = {k: v.copy() for k, v in w0.items()}
    np.random.seed(seed)
    ref._fit(X, sw, ui, x_uf, x_if, *[w[k] for k in WEIGHTS], alpha, beta, lr, schedule, expo, max_samples, epochs, False)
    rng = np.random.default_rng(seed + 1)
    pairs = np.stack([rng.integers(0, U, 400), rng.integers(0, I, 400)], 1).astype(np.float32)
    pairs[::37, 0] = np.nan; pairs[5::41, 1] = np.nan
    scores = ref._predict(np.ascontiguousarray(pairs), x_uf, x_if, *[w[k] for k in WEIGHTS])
    users = rng.integers(0, U, 24).astype(np.float32); users[3] = np.nan
    rec = ref._recommend(users, ui, 10, False, x_uf, x_if, *[w[k] for k in WEIGHTS])
    rec_f = ref._recommend(users, ui, 10, True, x_uf, x_if, *[w[k] for k in WEIGHTS])
    out = dict(interactions=X, sample_weight=sw, indptr=indptr, indices=indices, x_uf=x_uf, x_if=x_if, perms=perms,
               hyper=np.array([alpha, beta, lr, expo], np.float64), max_samples=max_samples, epochs=epochs, schedule=schedule, seed=seed,
               pairs=pairs, scores=scores, users=users, rec=rec, rec_filtered=rec_f)
    out.update({k + '_init': v for k, v in w0.items()})
    out.update({k + '_ref': v for k, v in w.items()})
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'N=%d' % len(X), {k: float(np.abs(v).max()) for k, v in w.items()})


def api_case(name, RankFM, seed):
    """through the reference's public class: raw ids, feature tables, np.random.seed-driven init"""
    rng = np.random.default_rng(seed)
    U, I, N = 120, 80, 1500
    uid = np.sort(rng.choice(10000, U, replace=False)); iid = np.sort(rng.choice(5000, I, replace=False))
    X = np.unique(np.stack([rng.integers(0, U, N), rng.integers(0, I, N)], 1), axis=0)
    X = np.unique(np.concatenate([X, np.stack([np.arange(U), rng.integers(0, I, U)], 1), np.stack([rng.integers(0, U, I), np.arange(I)], 1)]), axis=0)
    rng.shuffle(X)
    inter = np.stack([uid[X[:, 0]], iid[X[:, 1]]], 1)
    uf = np.concatenate([uid[:, None], rng.integers(0, 2, (U, 2)), rng.uniform(0, 2, (U, 1))], 1)[rng.permutation(U)]
    itf = np.concatenate([iid[:, None], rng.integers(0, 2, (I, 3))], 1)[rng.permutation(I)]
    sw = rng.uniform(0.5, 2.0, len(inter)).astype(np.float32)
    model = RankFM(factors=5, loss='warp', max_samples=6, learning_schedule='invscaling')
    np.random.seed(seed)
    model.fit(inter, user_features=uf, item_features=itf, sample_weight=sw, epochs=3)
    pairs = np.stack([rng.choice(np.append(uid, 99999), 200), rng.choice(np.append(iid, 99999), 200)], 1)
    users = rng.choice(np.append(uid, [77777, 88888]), 20, replace=False)
    out = dict(interactions=inter, user_features=uf, item_features=itf, sample_weight=sw, seed=seed, pairs=pairs,
               scores=model.predict(pairs), users=users,
               rec=model.recommend(users, n_items=7).values.astype(np.float64),
               rec_filtered=model.recommend(users, n_items=7, filter_previous=True).values.astype(np.float64),
               sim_items=model.similar_items(iid[3], 5), sim_users=model.similar_users(uid[7], 5))
    out.update({k + '_ref': getattr(model, k) for k in WEIGHTS})
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'N=%d' % len(inter))


def cfg1_problem(seed=31):
    """BASELINE.json configs[0] at its named size (10k users x 5k items, 100k interactions, factors=16, bpr, 5 epochs),
    regenerated from the seed with NumPy's stream-stable legacy generator so that the golden only has to store OUTPUTS"""
    rs = np.random.RandomState(seed)
    U, I, N, F = 10_000, 5_000, 100_000, 16
    pu = 1.0 / np.arange(1, U + 1) ** 0.6
    pi = 1.0 / np.arange(1, I + 1) ** 1.0
    X = np.stack([rs.choice(U, N, p=pu / pu.sum()), rs.choice(I, N, p=pi / pi.sum())], 1)
    X = np.unique(X, axis=0)
    X[:, 0] = rs.permutation(U)[X[:, 0]]
    X[:, 1] = rs.permutation(I)[X[:, 1]]
    X = X[rs.permutation(len(X))].astype(np.int32)
    order = np.lexsort((X[:, 1], X[:, 0]))
    indptr = np.zeros(U + 1, np.int64); np.cumsum(np.bincount(X[:, 0], minlength=U), out=indptr[1:])
    indices = np.ascontiguousarray(X[order, 1], dtype=np.int32)
    w = dict(w_i=np.zeros(I, np.float32), w_if=np.zeros(1, np.float32), v_u=rs.normal(0, .1, (U, F)).astype(np.float32),
             v_i=rs.normal(0, .1, (I, F)).astype(np.float32), v_uf=np.zeros((1, F), np.float32), v_if=np.zeros((1, F), np.float32))
    epochs = 5
    np.random.seed(seed)
    idx = np.arange(len(X), dtype=np.int32)
    perms = np.empty((epochs, len(X)), np.int32)
    for e in range(epochs):
        np.random.shuffle(idx); perms[e] = idx
    pairs = np.stack([rs.randint(0, U, 2000), rs.randint(0, I, 2000)], 1).astype(np.float32)
    users = rs.randint(0, U, 64).astype(np.float32)
    return dict(X=np.ascontiguousarray(X), sw=np.ones(len(X), np.float32), indptr=indptr, indices=indices, x_uf=np.zeros((U, 1), np.float32),
                x_if=np.zeros((I, 1), np.float32), w=w, perms=perms, epochs=epochs, hyper=(0.01, 0.1, 0.1, 'constant', 0.25), seed=seed,
                pairs=np.ascontiguousarray(pairs), users=users)


def cfg1_case(name, ref):
    p = cfg1_problem()
    ui = {u: p['indices'][p['indptr'][u]:p['indptr'][u + 1]] for u in range(len(p['indptr']) - 1)}
    w = p['w']
    np.random.seed(p['seed'])
    ref._fit(p['X'], p['sw'], ui, p['x_uf'], p['x_if'], *[w[k] for k in WEIGHTS], *p['hyper'], 1, p['epochs'], False)
    scores = ref._predict(p['pairs'], p['x_uf'], p['x_if'], *[w[k] for k in WEIGHTS])
    rec = ref._recommend(p['users'], ui, 10, True, p['x_uf'], p['x_if'], *[w[k] for k in WEIGHTS])
    np.savez_compressed(os.path.join(HERE, name + '.npz'), n_interactions=len(p['X']), checksum=int(p['X'].astype(np.int64).sum()),
                        w_i_ref=w['w_i'], v_u_ref=w['v_u'], v_i_ref=w['v_i'], scores=scores, rec_filtered=rec)
    print(name, 'N=%d' % len(p['X']), {k: float(np.abs(v).max()) for k, v in w.items()})


def eval_case(name, RankFM, evaluation, seed=41):
    """the reference's evaluation.py (`:9-175`) on a model with > 1k users, through the reference's own class"""
    rng = np.random.default_rng(seed)
    U, I, N = 1500, 400, 30000
    uid = np.sort(rng.choice(100000, U, replace=False)); iid = np.sort(rng.choice(50000, I, replace=False))
    pi = 1.0 / np.arange(1, I + 1)
    X = np.unique(np.stack([rng.integers(0, U, N), rng.choice(I, N, p=pi / pi.sum())], 1), axis=0)
    X = np.unique(np.concatenate([X, np.stack([np.arange(U), rng.integers(0, I, U)], 1), np.stack([rng.integers(0, U, I), np.arange(I)], 1)]), axis=0)
    rng.shuffle(X)
    inter = np.stack([uid[X[:, 0]], iid[X[:, 1]]], 1)
    test_mask = rng.random(len(inter)) < 0.15
    train, test = inter[~test_mask], inter[test_mask]
    # make sure every id is in train (the class indexes the ids it has seen), add unseen users / items to the test set
    train = np.concatenate([train, np.stack([uid, iid[rng.integers(0, I, U)]], 1), np.stack([uid[rng.integers(0, U, I)], iid], 1)])
    test = np.concatenate([test, np.array([[777777, iid[0]], [777778, iid[1]], [uid[0], 999999]])])
    model = RankFM(factors=8, loss='warp', max_samples=5, learning_schedule='invscaling')
    np.random.seed(seed)
    model.fit(train, epochs=4)
    out = dict(train=train, test=test, seed=seed)
    out.update({k + '_ref': getattr(model, k) for k in WEIGHTS})
    for k in (5, 10):
        for filt in (False, True):
            tag = "_k%d_%s" % (k, "filt" if filt else "all")
            out["hit_rate" + tag] = evaluation.hit_rate(model, test, k=k, filter_previous=filt)
            out["reciprocal_rank" + tag] = evaluation.reciprocal_rank(model, test, k=k, filter_previous=filt)
            out["dcg" + tag] = evaluation.discounted_cumulative_gain(model, test, k=k, filter_previous=filt)
            out["precision" + tag] = evaluation.precision(model, test, k=k, filter_previous=filt)
            out["recall" + tag] = evaluation.recall(model, test, k=k, filter_previous=filt)
    div = evaluation.diversity(model, test, k=10, filter_previous=True)
    out["diversity_item_id"] = div['item_id'].values
    out["diversity_cnt_users"] = div['cnt_users'].values
    out["diversity_pct_users"] = div['pct_users'].values
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, {k: float(v) for k, v in out.items() if k.startswith(('hit', 'rec', 'dcg', 'prec'))})


if __name__ == '__main__':
    ref = oracle.load_reference()
    assert ref is not None, "oracle/_ref could not be built: is /root/reference present?"
    only = set(sys.argv[1:])                                 # python make_golden.py [case ...]   (default: all)
    want = lambda name: not only or name in only
    if want('bpr_f16'): kernel_case('bpr_f16', ref, U=400, I=250, N=4000, F=16, P=0, Q=0, max_samples=1, epochs=3, schedule='constant', seed=11)
    if want('warp_f20'): kernel_case('warp_f20', ref, U=300, I=200, N=5000, F=20, P=0, Q=0, max_samples=20, epochs=3, schedule='invscaling', seed=12)
    if want('warp_feat'): kernel_case('warp_feat', ref, U=250, I=180, N=3000, F=6, P=3, Q=5, max_samples=5, epochs=2, schedule='invscaling', seed=13, sw_unit=False)
    if want('bpr_uf_only'): kernel_case('bpr_uf_only', ref, U=150, I=120, N=1500, F=10, P=4, Q=0, max_samples=1, epochs=2, schedule='constant', seed=14)
    if want('warp_if_only'): kernel_case('warp_if_only', ref, U=150, I=120, N=1500, F=3, P=0, Q=2, max_samples=4, epochs=2, schedule='constant', seed=15, sw_unit=False)
    if want('api_warp_feat'): api_case('api_warp_feat', oracle.load_reference_class(), seed=21)
    if want('cfg1_bpr'): cfg1_case('cfg1_bpr', ref)
    if want('eval_ref'):
        import rankfm.evaluation as ref_evaluation           # the reference's module, importable next to oracle/_ref's compiled _rankfm
        eval_case('eval_ref', oracle.load_reference_class(), ref_evaluation)
