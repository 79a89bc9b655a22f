"""bench.py's own arm, host logic only: the CUDA session is replaced by a stand-in (fixed kernel times, the CPU oracle behind
the plug-in `_fit`), so that the line assembly -- roofline arithmetic, e2e statistic, cpu_baseline leg, the single JSON
line -- is exercised without a GPU.  Numbers mean nothing here; the real run is the driver's."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _FakeSession:
    def __init__(self, problem, keep):
        self.n = problem.n_interactions
        self.launches = 0

    def snapshot(self): pass
    def restore(self): pass
    def flush_l2(self): pass
    def timer_start(self): pass
    def timer_stop(self): return 12.5
    def launch_count(self): return self.launches
    def close(self): pass
    def exchange_path(self): return 0

    def train(self, epochs):
        self.launches += 2 * epochs
        return [dict(log_likelihood=-1.0, penalty=0.1, draws=int(1.5 * self.n), finite=[1] * 6, eta=0.1, kernel_ms=0.5, sync_ms=0.0) for _ in range(epochs)]


def test_run_ours_assembles_the_line(monkeypatch, capsys):
    import bench
    from oracle import oracle
    from rankfm_b200 import _lib, _rankfm

    class _Lib:
        def rfm_device_count(self): return 1
    monkeypatch.setattr(_lib, "lib", lambda: _Lib())
    monkeypatch.setattr(_rankfm, "Session", _FakeSession)
    monkeypatch.setattr(_rankfm, "pin", lambda *a: None)
    monkeypatch.setattr(_rankfm, "unpin", lambda *a: None)
    monkeypatch.setattr(_rankfm, "_fit", oracle._fit)
    monkeypatch.setattr(_rankfm, "set_resident_training", lambda flag: None)
    monkeypatch.setattr(_rankfm, "drop_training", lambda: None)
    monkeypatch.setattr(bench, "ClockSampler", lambda index: type("C", (), {"stop": lambda self: {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 1}})())
    monkeypatch.setenv("BENCH_CPU_BASELINE_S", "2")
    monkeypatch.setitem(bench.CONFIGS, "cfg1", dict(bench.CONFIGS["cfg1"], N=20_000, epochs=2))
    args = argparse.Namespace(gpus=1, steps=3, warmup=3, impl="ours", workload="cfg1", no_cpu_baseline=False, no_recommend=True, sub="all")
    bench.run_ours(args)
    lines = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    N = d["config"]["interactions_per_gpu"]
    assert d["metric"] == "training interactions/sec" and d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3
    assert np.isclose(d["value"], N * 2 * 3 / 12.5e-3) and np.isclose(d["ms_per_step"], 12.5 / 3)
    r = d["roofline"]
    bytes_per_positive = bench.algorithmic_bytes_per_positive(16, 1.5)
    assert r["bound"] == "hbm" and np.isclose(r["achieved"], N * bytes_per_positive / 0.5e-3 / 1e9) and np.isclose(r["frac"], r["achieved"] / r["peak"])
    assert np.isclose(r["mean_draws_per_positive"], 1.5) and d["gpu_launches"] == 12
    e = d["e2e"]
    assert e["unit"] == "interactions/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and len(e["ms_each"]) >= 3
    assert np.isclose(e["value"], N * 2 / (e["ms_per_step"] * 1e-3))
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] == 1 and c["value"] > 0
    assert d["e2e_resident"]["h2d_bytes_per_step"] < e["h2d_bytes_per_step"]
    assert "recommend" not in d and "workloads" not in d and d["clocks"]["reasons"] == []
