#!/usr/bin/env python
"""One recommend() probe (bench.recommend_probe) under the RANKFM_B200_GEMM_MSUB / RANKFM_B200_TAU_STRIDE variant the
environment selects; prints one JSON line.  `--small` = one batch, one iteration (for ncu captures)."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import bench  # noqa: E402

if __name__ == "__main__":
    small = "--small" in sys.argv
    kw = dict(n_users=18944, iters=1, exact_users=0) if small else dict(exact_users=0)
    for a in sys.argv[1:]:
        if a.startswith("--users="):
            kw["n_users"] = int(a.split("=")[1])
        if a.startswith("--items="):
            kw["n_items_cat"] = int(a.split("=")[1])
        if a.startswith("--factors="):
            kw["factors"] = int(a.split("=")[1])
    print(json.dumps(bench.recommend_probe(**kw)))
