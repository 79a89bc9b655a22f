"""Minimal driver for profiler captures of the tensor-core recommend path: the bench's 65,536 x 262,144 probe, one pass.
usage: python profiles/tools/recommend_once.py [users] [items]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

u = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
i = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
r = bench.recommend_probe(n_users=u, n_items_cat=i, iters=1, exact_users=0)
print({k: r[k] for k in ("ms_total", "ms_gemm_filter", "frac_of_bf16_peak", "users_per_s", "e2e_users_per_s", "e2e_ms", "rows_redone_on_exact_path", "rows_redone_with_provable_threshold", "topk_overlap_vs_exact", "variant")})
