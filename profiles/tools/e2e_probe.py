"""Phase timing of the stateless plug-in call: RANKFM_B200_TIMING=1 makes rfm_fit print create+H2D / train / D2H / destroy.
usage: RANKFM_B200_TIMING=1 python profiles/tools/e2e_probe.py [workload] [calls] [pin]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
import bench  # noqa: E402
from rankfm_b200 import _rankfm  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 20
c = bench.make_workload(name, device=0)
X = c["X"]
ui = _rankfm.UserItems.from_interactions(X, c["U_global"], c["I"])
w = bench.alloc_weights(c)
if len(sys.argv) > 3:
    _rankfm.pin(X, c["sw"], ui.indptr, ui.indices, *[w[k] for k in bench.WEIGHTS])
_rankfm.set_resident_training(os.environ.get("RANKFM_B200_RESIDENT_TRAIN", "0") == "1")


def step():
    bench.reset_weights(c, w)
    t0 = time.perf_counter()
    _rankfm._fit(X, c["sw"], ui, c["x_uf"], c["x_if"], *[w[k] for k in bench.WEIGHTS], *bench.HYPER_ARGS, c["max_samples"], c["epochs"], False)
    return time.perf_counter() - t0


print([round(step() * 1e3, 1) for _ in range(calls)])
