"""Debug aid: the three row-threshold modes of the tensor-core recommend path must return identical rows."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
os.environ["RANKFM_B200_TAU_TAIL"] = "4"; os.environ["RANKFM_B200_RECOMMEND"] = "tc"
import test_gpu_parity as t
for ew in ("16", "8"):
  for mode, z, stride in (("estimate", "4.5", "8"), ("head", "4.5", "32"), ("estimate", "0.001", "8")):
    os.environ["RANKFM_B200_GEMM_EW"] = ew
    U = 600
    sess, ui = t._sparse_scoring_session(U, 130000, 24, seed=5)
    users = np.arange(U, dtype=np.float32); users[17] = np.nan
    os.environ["RANKFM_B200_TAU_MODE"] = "safe"; os.environ.pop("RANKFM_B200_TAU_STRIDE", None)
    safe = sess.recommend(users, 20, False)
    r0 = sess.recommend_stats(), sess.recommend_retried()
    os.environ["RANKFM_B200_TAU_MODE"] = mode; os.environ["RANKFM_B200_TAU_Z"] = z; os.environ["RANKFM_B200_TAU_STRIDE"] = stride
    est = sess.recommend(users, 20, False)
    r1 = sess.recommend_stats(), sess.recommend_retried()
    os.environ["RANKFM_B200_RECOMMEND"] = "exact"
    exact = sess.recommend(users, 20, False)
    os.environ["RANKFM_B200_RECOMMEND"] = "tc"
    sess.close()
    bad = [r for r in range(U) if not np.array_equal(est[r], safe[r], equal_nan=True)]
    print("EW", ew, mode, z, stride, "stats safe", r0, "after", r1, "rows differing", len(bad), bad[:10])
    for r in bad[:3]:
        print("  row", r, "est ", est[r].astype(int).tolist()); print("        safe", safe[r].astype(int).tolist()); print("        exct", exact[r].astype(int).tolist())
        print("   set diff est-safe", set(est[r]) - set(safe[r]), "safe-est", set(safe[r]) - set(est[r]))
    print("   safe vs exact rows differing", sum(not np.array_equal(safe[r], exact[r], equal_nan=True) for r in range(U)), " est vs exact", sum(not np.array_equal(est[r], exact[r], equal_nan=True) for r in range(U)))
