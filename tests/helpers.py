"""shared test helpers: golden fixtures, synthetic data, metric helpers (test infrastructure, may use oracle/)"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
WEIGHTS = ('w_i', 'w_if', 'v_u', 'v_i', 'v_uf', 'v_if')
KERNEL_CASES = ['bpr_f16', 'warp_f20', 'warp_feat', 'bpr_uf_only', 'warp_if_only']


class CSRItems(dict):
    """minimal dict-with-CSR used as `user_items` in tests"""

    def __init__(self, indptr, indices):
        super().__init__({u: indices[indptr[u]:indptr[u + 1]] for u in range(len(indptr) - 1)})
        self.indptr, self.indices = indptr, indices


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    for k in ('max_samples', 'epochs', 'seed'):
        if k in g:
            g[k] = int(g[k])
    if 'schedule' in g:
        g['schedule'] = str(g['schedule'])
    return g


def golden_fit_args(g, which='init'):
    """positional arguments of `_fit` up to `max_samples` for a golden kernel case; fresh copies of the weights"""
    w = {k: np.ascontiguousarray(g[k + '_' + which]).copy() for k in WEIGHTS}
    ui = CSRItems(g['indptr'], g['indices'])
    alpha, beta, lr, expo = [float(x) for x in g['hyper']]
    args = (g['interactions'], g['sample_weight'], ui, g['x_uf'], g['x_if'], w['w_i'], w['w_if'], w['v_u'], w['v_i'], w['v_uf'], w['v_if'],
            alpha, beta, lr, g['schedule'], expo, g['max_samples'])
    return args, w, ui


def zipf_interactions(U, I, N, seed=42, a_u=0.6, a_i=1.0):
    """synthetic (user,item) pairs per SURVEY.md section 8(d): Zipf users/items, de-duplicated, ids permuted"""
    rng = np.random.default_rng(seed)
    pu = 1.0 / np.arange(1, U + 1) ** a_u
    pi = 1.0 / np.arange(1, I + 1) ** a_i
    X = np.stack([rng.choice(U, N, p=pu / pu.sum()), rng.choice(I, N, p=pi / pi.sum())], 1)
    X = np.unique(X, axis=0)
    X[:, 0] = rng.permutation(U)[X[:, 0]]
    X[:, 1] = rng.permutation(I)[X[:, 1]]
    # re-index to the observed uniques like rankfm.py:115-116
    _, X[:, 0] = np.unique(X[:, 0], return_inverse=True)
    _, X[:, 1] = np.unique(X[:, 1], return_inverse=True)
    rng.shuffle(X)
    return np.ascontiguousarray(X, dtype=np.int32)


def csr_of(X, U):
    order = np.lexsort((X[:, 1], X[:, 0]))
    counts = np.bincount(X[:, 0], minlength=U)
    indptr = np.zeros(U + 1, np.int64)
    np.cumsum(counts, out=indptr[1:])
    return indptr, np.ascontiguousarray(X[order, 1], dtype=np.int32)


def init_weights(U, I, F, P=0, Q=0, seed=0, sigma=0.1):
    rng = np.random.default_rng(seed)
    return dict(w_i=np.zeros(I, np.float32), w_if=np.zeros(max(Q, 1), np.float32),
                v_u=rng.normal(0, sigma, (U, F)).astype(np.float32), v_i=rng.normal(0, sigma, (I, F)).astype(np.float32),
                v_uf=(rng.normal(0, sigma / 10, (P, F)) if P else np.zeros((1, F))).astype(np.float32),
                v_if=(rng.normal(0, sigma / 10, (Q, F)) if Q else np.zeros((1, F))).astype(np.float32))


def features(U, I, P, Q, seed=0, sparsity=0.3):
    rng = np.random.default_rng(seed + 100)
    x_uf = rng.uniform(0, 1, (U, P)).astype(np.float32) if P else np.zeros((U, 1), np.float32)
    x_if = rng.uniform(0, 1, (I, Q)).astype(np.float32) if Q else np.zeros((I, 1), np.float32)
    if P:
        x_uf[rng.uniform(size=x_uf.shape) < sparsity] = 0
    if Q:
        x_if[rng.uniform(size=x_if.shape) < sparsity] = 0
    return x_uf, x_if


def rel_err(a, b):
    """max over entries of |a-b| / max(|b|, 0.1*rms(b)): relative error, where entries smaller than a tenth of the
    array's typical magnitude are judged against that tenth (float32 reassociation noise on a sum of O(rms) terms
    is ~1e-7*rms absolute, so a pure per-entry relative error near zero crossings measures nothing)"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    m = np.isfinite(a) & np.isfinite(b)
    if not m.any():
        return 0.0
    floor = max(0.1 * float(np.sqrt(np.mean(b[m] ** 2))), 1e-12)
    return float(np.max(np.abs(a[m] - b[m]) / np.maximum(np.abs(b[m]), floor)))


def topk_overlap(rec_a, rec_b):
    """mean per-row |set(a) & set(b)| / k over rows where b is not NaN"""
    tot, n = 0.0, 0
    for ra, rb in zip(rec_a, rec_b):
        if np.isnan(rb).all():
            assert np.isnan(ra).all()
            continue
        sb = set(rb[~np.isnan(rb)].tolist())
        sa = set(ra[~np.isnan(ra)].tolist())
        tot += len(sa & sb) / max(len(sb), 1)
        n += 1
    return tot / max(n, 1)


def golden_module():
    """tests/golden/make_golden.py as a module (its problem generators are shared with the tests so that large cases only
    store OUTPUTS: the inputs are regenerated from the seed with NumPy's stream-stable legacy generator)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cfg1_case():
    """BASELINE.json configs[0] at its named size + the reference's outputs for it -> (problem dict, golden dict)"""
    p = golden_module().cfg1_problem()
    z = np.load(os.path.join(GOLDEN, "cfg1_bpr.npz"))
    g = {k: z[k] for k in z.files}
    assert int(g["n_interactions"]) == len(p["X"]) and int(g["checksum"]) == int(p["X"].astype(np.int64).sum()), \
        "the regenerated cfg1 inputs differ from the ones the golden was minted on (NumPy legacy stream changed?)"
    return p, g


def assert_topn_exact_up_to_rounding(got, scores64, n, tol):
    """`got` is THE top-n of `scores64` (float64 ground truth), except that candidates whose scores differ by less than the
    float32 rounding bound `tol` may swap places or swap across the cut: (1) n distinct entries, (2) everything clearly
    above the n-th best score is present, (3) nothing clearly below it is, (4) the order is descending up to tol"""
    got = np.asarray(got)
    assert len(got) == n and len(set(got.tolist())) == n, got
    kth = np.sort(scores64)[::-1][n - 1]
    s = scores64[got]
    must = set(np.flatnonzero(scores64 > kth + tol).tolist())
    assert must <= set(got.tolist()), (sorted(must - set(got.tolist())), kth)
    assert (s >= kth - tol).all(), (s, kth)
    assert (s[:-1] >= s[1:] - tol).all(), s
