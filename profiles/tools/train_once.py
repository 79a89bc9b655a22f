"""Minimal driver for profiler captures: build one bench workload, upload it, train a few epochs on one GPU.
usage: python profiles/tools/train_once.py <workload> [epochs] [sample_n]     (run under ncu; prints per-epoch kernel ms)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rankfm_b200 import _rankfm  # noqa: E402

name = sys.argv[1]
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
c = bench.make_workload(name, device=0)
if len(sys.argv) > 3:
    n = int(sys.argv[3])
    c["X"], c["sw"] = c["X"][:n].copy(), c["sw"][:n].copy()
ui = _rankfm.UserItems.from_interactions(c["X"], c["U_global"], c["I"])
w = bench.alloc_weights(c)
keep = []
prob = _rankfm.fit_problem(c["X"], c["sw"], ui, c["x_uf"], c["x_if"], *[w[k] for k in bench.WEIGHTS], *bench.HYPER_ARGS, c["max_samples"],
                           mode="production", seed=1492, keep=keep)
sess = _rankfm.Session(prob, keep)
for s in sess.train(epochs):
    print("kernel_ms %.3f draws/positive %.2f ll %.1f" % (s["kernel_ms"], s["draws"] / len(c["X"]), s["log_likelihood"]))
sess.close()
