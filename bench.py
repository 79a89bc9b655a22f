#!/usr/bin/env python
"""bench.py -- training interactions/sec of the RankFM hot path (`_fit` epoch loop) on N x B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

A "step" is one full pass of the hot path over the workload: `epochs` SGD epochs over all interactions, starting from
the same initial weights every step.  Default workload = BASELINE.json configs[1] (MovieLens-1M shape synthetic, 6040 x
3706, 1M interactions, factors=20, loss='warp', max_samples=20, 20 epochs).

  value   interactions/s (N * epochs * K / device time) with every input already resident in HBM when the timed region
          starts (CUDA events on the library's stream, max over ranks); the region holds, per step: D2D restore of the
          initial weights, an L2 flush (512 MB memset), `epochs` SGD kernel launches + per-epoch weight-stat kernels
  e2e     the same metric through the reference-facing plug-in call `rankfm_b200._rankfm._fit(...)` on HOST buffers:
          H2D of interactions/CSR/weights, all epochs, D2H of the weights, wall clock
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md "Measurement"

Multi-GPU (torchrun, one process per GPU): weak scaling -- every rank owns a cfg2-sized block of users (global U =
6040*N, 1M*N interactions, one shared item catalogue), item-side deltas are summed over NCCL once per epoch.
torch.distributed (gloo) is only the control plane here (NCCL-id broadcast, barrier, max-reduce of the timings).

--impl reference times the UNMODIFIED reference Cython `_fit` (oracle/_ref, built by oracle/build_ref.py) on the
host cores, single-threaded by construction (GIL held, no OpenMP), on a bounded slice of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from rankfm_b200.synthetic import CONFIGS, init_weights, side_features, zipf_interactions  # noqa: E402

WEIGHTS = ('w_i', 'w_if', 'v_u', 'v_i', 'v_uf', 'v_if')
HYPER = dict(alpha=0.01, beta=0.1, learning_rate=0.1, learning_schedule='invscaling', learning_exponent=0.25)
# the reference's only published figure for this path and this configuration (MovieLens-1M,
# factors=20, warp, max_samples=20, invscaling, 20 epochs): 29.7 s wall for 749,724 x 20 interactions on a
# "2.3 GHz i5 MacBook", one thread (README.md:75-77, examples/movielens.ipynb:1073-1080).  Other hardware, real data.
PUBLISHED_CFG2_INTERACTIONS_PER_S = 749_724 * 20 / 29.7


_RESULT_FD = None


def quiet_stdout():
    """the contract is ONE JSON line on stdout: anything a library prints there (NCCL's version banner, for one) is sent
    to stderr instead, and the result line is written to the original stdout at the end"""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_RESULT_FD, data)


def vs_published(value, workload):
    return value / PUBLISHED_CFG2_INTERACTIONS_PER_S if workload == "cfg2" else None



def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes_per_positive(F, S, P=0, Q=0):
    """4F(S+5) + 4S + 28 (+ 4P + 4Q(1+S) with dense side features)"""
    return 4.0 * F * (S + 5.0) + 4.0 * S + 28.0 + (4.0 * P + 4.0 * Q * (1.0 + S) if (P or Q) else 0.0)


def table_note(c):
    mb = 4.0 * (c["U_global"] * (c["F"] + c["P"]) + c["I"] * (c["F"] + 1 + c["Q"])) / 1e6
    if mb < 100:
        return ("this workload's tables (%.1f MB) live in the 126 MB L2: DRAM traffic is only the interaction stream, so the HBM fraction is low by "
                "construction; see roofline_dram_resident for the same kernel on tables that do not fit L2" % mb)
    return "tables %.0f MB (> 126 MB L2): rows come from DRAM except for the Zipf-hot ones" % mb


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs"""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace('.', '').isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 6 for k in range(4) if r[2 + k].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def make_workload(name, rank=0, world=1):
    c = dict(CONFIGS[name])
    X = zipf_interactions(c["U"], c["I"], c["N"], seed=42 + rank, a_u=float(os.environ.get("BENCH_ZIPF_U", 0.6)), a_i=float(os.environ.get("BENCH_ZIPF_I", 1.0)))
    U_alloc = c["U"]                                       # fixed block per rank so every rank agrees on the global U
    X[:, 0] += rank * U_alloc
    c["U_global"] = U_alloc * world
    c["X"] = X
    c["sw"] = np.ones(len(X), np.float32)
    c["x_uf"], c["x_if"] = side_features(c["U_global"], c["I"], c["P"], c["Q"])
    c["w0"] = init_weights(c["U_global"], c["I"], c["F"], c["P"], c["Q"], seed=0)
    return c


def fresh_weights(c):
    return {k: v.copy() for k, v in c["w0"].items()}


# -----------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# -----------------------------------------------------------------------------------------------------------------
def reference_fit_rate(c, epochs, repeats=1):
    """interactions/s of the reference's own Cython `_fit` (single host thread) on this workload; None if unavailable"""
    from oracle import oracle
    ref = oracle.load_reference(build_if_possible=True)
    kind = "reference"
    fit = ref._fit if ref is not None else None
    if fit is None:
        kind, fit = "port", oracle._fit
    X, U = c["X"], c["U_global"]
    order = np.lexsort((X[:, 1], X[:, 0]))
    counts = np.bincount(X[:, 0], minlength=U)
    bounds = np.concatenate([[0], np.cumsum(counts)])
    items = X[order, 1].astype(np.int32)
    ui = {u: items[bounds[u]:bounds[u + 1]] for u in range(U)}
    best = 0.0
    for _ in range(repeats):
        w = fresh_weights(c)
        np.random.seed(0)
        t0 = time.perf_counter()
        fit(X, c["sw"], ui, c["x_uf"], c["x_if"], *[w[k] for k in WEIGHTS], HYPER["alpha"], HYPER["beta"], HYPER["learning_rate"],
            HYPER["learning_schedule"], HYPER["learning_exponent"], c["max_samples"], epochs, False)
        dt = time.perf_counter() - t0
        best = max(best, len(X) * epochs / dt)
    return best, kind


def bounded_reference_sample(c, budget_s):
    """bounded sample of a workload for the single-threaded reference: a quick probe gives this box's rate for
    the workload's row shape, then a run is as many epochs over as many interactions (same U, I, F, features -- the
    per-interaction cost is scale-free apart from cache effects) as fit `budget_s`.  The first epochs are also the
    reference's fastest (fewest WARP draws).  -> (workload dict of the sample, epochs, interactions)"""
    N = len(c["X"])
    probe_n = min(N, 100_000)
    probe = dict(c, X=np.ascontiguousarray(c["X"][:probe_n]), sw=c["sw"][:probe_n])
    t0 = time.perf_counter()
    reference_fit_rate(probe, 1)
    probe_rate = probe_n / (time.perf_counter() - t0)                       # includes the Python set-up over all U users
    sample_epochs = int(min(c["epochs"], max(2, budget_s * probe_rate / N)))
    sample_n = int(min(N, max(probe_n, budget_s * probe_rate / sample_epochs)))
    cs = c if sample_n == N else dict(c, X=np.ascontiguousarray(c["X"][:sample_n]), sw=c["sw"][:sample_n])
    return cs, sample_epochs, sample_n


def run_reference_recommend(args):
    """reference arm of `--workload cfg5`: the reference's own `_recommend` (scalar scoring loop + full argsort per user,
    `_rankfm.pyx:393-460`) on a bounded sample of users of the same synthetic model; users/s scale-free in the user count"""
    from oracle import oracle
    ref = oracle.load_reference(build_if_possible=True)
    kind, rec = ("reference", ref._recommend) if ref is not None else ("port", oracle._recommend)
    U = int(os.environ.get("BENCH_CFG5_USERS", 1_000_000))
    I = int(os.environ.get("BENCH_CFG5_ITEMS", 1_000_000))
    F, topn, sample = 128, 100, int(os.environ.get("BENCH_CFG5_REF_SAMPLE", 24))
    rng = np.random.default_rng(0)                           # same draws as recommend_probe
    w_i = rng.normal(0, 0.3, I).astype(np.float32)
    v_u = rng.standard_normal((U, F), dtype=np.float32) * np.float32(0.1)
    v_i = rng.standard_normal((I, F), dtype=np.float32) * np.float32(0.1)
    zeros = lambda *shape: np.zeros(shape, np.float32)
    users = np.arange(min(sample, U), dtype=np.float32)
    times = []
    for k in range(args.warmup + max(1, args.steps)):
        t0 = time.perf_counter()
        rec(users, {}, topn, False, zeros(U, 1), zeros(I, 1), w_i, zeros(1), v_u, v_i, zeros(1, F), zeros(1, F))
        if k >= args.warmup:
            times.append(time.perf_counter() - t0)
    value = len(users) / float(np.mean(times))
    emit({"impl": "reference", "metric": "recommend users/sec", "value": value, "unit": "users/s", "n_gpus": args.gpus, "steps": max(1, args.steps),
          "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
          "data": "synthetic", "config": {"workload": "recommend top-%d, %d users x %d items, factors=%d" % (topn, U, I, F),
                                          "sample": "%d users per step (the reference's per-user cost does not depend on the number of users)" % len(users)},
          "cpu_baseline": {"value": value, "unit": "users/s", "cores": 1, "kind": kind, "sample": "%d users x %d items per step; single-threaded by construction, 1 thread of %d" % (len(users), I, os.cpu_count())},
          "e2e": {"value": value, "unit": "users/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "cfg5":
        return run_reference_recommend(args)
    c = make_workload(args.workload)
    cs, sample_epochs, sample_n = bounded_reference_sample(c, float(os.environ.get("BENCH_REF_BUDGET_S", 120.0)) / (args.steps + args.warmup))
    N = len(c["X"])
    for _ in range(args.warmup):
        reference_fit_rate(cs, sample_epochs)
    t0 = time.perf_counter()
    rates = [reference_fit_rate(cs, sample_epochs) for _ in range(args.steps)]
    dt = time.perf_counter() - t0
    kind = rates[0][1]
    value = float(np.mean([r[0] for r in rates]))
    what = "all %d interactions" % N if sample_n == N else "the first %d of %d interactions (same users, items, factors, features; extrapolated)" % (sample_n, N)
    line = {
        "impl": "reference", "metric": "training interactions/sec", "value": value, "unit": "interactions/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": vs_published(value, args.workload), "dtype": "f32", "data": "synthetic",
        "config": {"workload": c["label"], "sample": "%d of %d epochs per step, %s" % (sample_epochs, c["epochs"], what)},
        "cpu_baseline": {"value": value, "unit": "interactions/s", "cores": 1, "kind": kind,
                         "sample": "%d epochs x %d interactions per step; the reference holds the GIL and has no OpenMP: 1 thread of %d" % (sample_epochs, sample_n, os.cpu_count())},
        "e2e": {"value": value, "unit": "interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# -----------------------------------------------------------------------------------------------------------------
# our arm
# -----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist                 # control plane only (gloo): id broadcast, barrier, max-reduce
        dist.init_process_group(backend="gloo", init_method="env://")
    from rankfm_b200 import _lib, _rankfm
    assert _lib.lib().rfm_device_count() > 0, "bench.py needs a CUDA device (no CPU fallback)"
    _rankfm.set_device(local_rank)
    if world > 1:
        import torch
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            idt = torch.frombuffer(bytearray(_rankfm.nccl_unique_id()), dtype=torch.uint8).clone()
        dist.broadcast(idt, src=0)
        _rankfm.set_comm(rank, world, idt.numpy().tobytes())

    c = make_workload(args.workload, rank, world)
    X, N, epochs = c["X"], len(c["X"]), c["epochs"]
    ui = _rankfm.UserItems.from_interactions(X, c["U_global"])
    w = fresh_weights(c)
    keep = []
    prob = _rankfm.fit_problem(X, c["sw"], ui, c["x_uf"], c["x_if"], *[w[k] for k in WEIGHTS], HYPER["alpha"], HYPER["beta"],
                               HYPER["learning_rate"], HYPER["learning_schedule"], HYPER["learning_exponent"], c["max_samples"],
                               mode="production", seed=1492, keep=keep)
    sess = _rankfm.Session(prob, keep)
    sess.snapshot()

    def barrier():
        if dist is not None:
            dist.barrier()

    def one_step():
        sess.restore()
        sess.flush_l2()
        return sess.train(epochs)

    for _ in range(args.warmup):
        one_step()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    launches0 = sess.launch_count()
    barrier()
    sess.timer_start()                                   # synchronises the stream, then records the start event
    all_stats = [one_step() for _ in range(args.steps)]
    ms = sess.timer_stop()                               # records + synchronises the stop event
    barrier()
    launches = sess.launch_count() - launches0
    clock_info = clocks.stop() if clocks else None
    if dist is not None:
        import torch
        t = torch.tensor([ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        ln = torch.tensor([launches], dtype=torch.int64)
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)
        launches = int(ln.item())
    total_interactions = N * epochs * args.steps * world
    value = total_interactions / (ms / 1e3)

    # ---- roofline of the dominant kernel (sgd_epoch_kernel), from the per-launch CUDA events of the timed steps ----
    flat = [s for step in all_stats for s in step]
    kern_ms = sum(s["kernel_ms"] for s in flat)
    alg_bytes = sum(N * algorithmic_bytes_per_positive(c["F"], s["draws"] / N, c["P"], c["Q"]) for s in flat)
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (kern_ms / 1e3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_%s.json" % args.workload)
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "sgd_pipe_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes / len(flat),
                "launch_ms": kern_ms / len(flat), "kernel_share_of_step": kern_ms / (ms if world == 1 else max(ms, 1e-9)),
                "mean_draws_per_positive": float(np.mean([s["draws"] / N for s in flat])),
                "note": table_note(c)}

    line = None
    if rank == 0:
        # ---- e2e: the plug-in call on host buffers (H2D + epochs + D2H inside the timed region) ----
        # host buffers of the plug-in call are page-locked in place (the caller's choice; rfm_host_register)
        e2e_w = fresh_weights(c)
        pinned = [X, c["sw"], ui.indptr, ui.indices] + [e2e_w[k] for k in WEIGHTS]

        def e2e_step():
            ww = e2e_w
            for k in WEIGHTS:
                ww[k][...] = c["w0"][k]
            t0 = time.perf_counter()
            _rankfm._fit(X, c["sw"], ui, c["x_uf"], c["x_if"], *[ww[k] for k in WEIGHTS], HYPER["alpha"], HYPER["beta"], HYPER["learning_rate"],
                         HYPER["learning_schedule"], HYPER["learning_exponent"], c["max_samples"], epochs, False)
            return time.perf_counter() - t0
        e2e = None
        if world == 1:
            _rankfm.pin(*pinned)
            for _ in range(5):                            # the block cache / lazy module loading settle over the first calls
                e2e_step()
            dts = [e2e_step() for _ in range(max(15, args.steps))]
            dt_med = float(np.median(dts))
            h2d = X.nbytes + c["sw"].nbytes + ui.indptr.nbytes + ui.indices.nbytes + sum(v.nbytes for v in c["w0"].values()) + \
                (c["x_uf"].nbytes if c["P"] else 0) + (c["x_if"].nbytes if c["Q"] else 0)
            d2h = sum(v.nbytes for v in c["w0"].values())
            _rankfm.unpin(*pinned)
            # wall-clock steps on a shared host see occasional scheduling hiccups (one 470 ms step among 25 ms ones was
            # observed): the value is the MEDIAN step, every sample and the mean are reported next to it
            e2e = {"value": N * epochs / dt_med, "unit": "interactions/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "ms_per_step": 1e3 * dt_med, "statistic": "median of %d steps" % len(dts), "mean_ms_per_step": 1e3 * float(np.mean(dts)),
                   "ms_each": [round(1e3 * d, 2) for d in dts], "call": "rankfm_b200._rankfm._fit(page-locked host ndarray buffers) -> ctypes -> rfm_fit"}
        recommend = None
        if world == 1 and not args.no_recommend:
            try:
                recommend = recommend_probe(device=local_rank)
            except Exception as exc:                    # secondary measurement must never take the headline down
                recommend = {"error": repr(exc)}
        roofline_large = None
        if world == 1 and not args.no_large and args.workload == "cfg2":
            try:
                roofline_large = dram_resident_probe(device=local_rank)
            except Exception as exc:
                roofline_large = {"error": repr(exc)}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cs, sample_epochs, sample_n = bounded_reference_sample(c, float(os.environ.get("BENCH_CPU_BASELINE_S", 20.0)))
            rate, kind = reference_fit_rate(cs, sample_epochs)
            cpu = {"value": rate, "unit": "interactions/s", "cores": 1, "kind": kind,
                   "sample": "%d of %d epochs x %d of %d interactions; single-threaded by construction, %d host cores present" % (sample_epochs, epochs, sample_n, N, os.cpu_count())}
        line = {
            "metric": "training interactions/sec", "value": value, "unit": "interactions/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": vs_published(value, args.workload), "dtype": "f32", "data": "synthetic",
            "config": {"workload": c["label"], "published_baseline": "%.0f interactions/s = reference README/notebook, MovieLens-1M (real data, same shape/hyper-parameters), 2.3 GHz i5 MacBook, 1 thread" % PUBLISHED_CFG2_INTERACTIONS_PER_S,
                       "interactions_per_gpu": N, "epochs_per_step": epochs, "users": c["U_global"], "items": c["I"],
                       "l2": "flushed between steps (512 MB memset inside the timed region)",
                       "parallelism": "user-sharded x%d, per-epoch NCCL sum of item deltas" % world if world > 1 else "single GPU",
                       "schedule": "production: Hogwild lane-group per positive, Philox negatives, on-device Feistel order"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clock_info,
            "recommend": recommend, "roofline_dram_resident": roofline_large,
            "final_log_likelihood": flat[-1]["log_likelihood"],
        }
    sess.close()
    if dist is not None and os.environ.get("BENCH_E2E_MULTI", "1") == "1":
        # ---- e2e at N GPUs: every rank calls the plug-in `_fit` on ITS shard from host buffers; a call builds its own NCCL
        # communicator (fresh unique id per call, broadcast over gloo), trains, exchanges deltas, copies the model back.
        # Guarded by a watchdog: if anything on this secondary path stalls, the line goes out without it.
        def bail():
            if line is not None:
                emit(line)
            os._exit(0)
        dog = threading.Timer(float(os.environ.get("BENCH_E2E_MULTI_TIMEOUT", "90")), bail)
        dog.daemon = True
        dog.start()
        try:
            import torch
            e2e_w = fresh_weights(c)
            dts = []
            for k in range(2 + 5):
                idt = torch.zeros(128, dtype=torch.uint8)
                if rank == 0:
                    idt = torch.frombuffer(bytearray(_rankfm.nccl_unique_id()), dtype=torch.uint8).clone()
                dist.broadcast(idt, src=0)
                _rankfm.set_comm(rank, world, idt.numpy().tobytes())
                for name in WEIGHTS:
                    e2e_w[name][...] = c["w0"][name]
                dist.barrier()
                t0 = time.perf_counter()
                _rankfm._fit(X, c["sw"], ui, c["x_uf"], c["x_if"], *[e2e_w[name] for name in WEIGHTS], HYPER["alpha"], HYPER["beta"], HYPER["learning_rate"],
                             HYPER["learning_schedule"], HYPER["learning_exponent"], c["max_samples"], epochs, False)
                t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                if k >= 2:
                    dts.append(float(t.item()))
            if line is not None:
                dt_med = float(np.median(dts))
                per_rank = X.nbytes + c["sw"].nbytes + ui.indptr.nbytes + ui.indices.nbytes + sum(v.nbytes for v in c["w0"].values())
                line["e2e"] = {"value": N * epochs * world / dt_med, "unit": "interactions/s", "h2d_bytes_per_step": int(per_rank * world),
                               "d2h_bytes_per_step": int(sum(v.nbytes for v in c["w0"].values()) * world), "ms_per_step": 1e3 * dt_med,
                               "statistic": "median of %d steps, max over ranks" % len(dts), "ms_each": [round(1e3 * d, 2) for d in dts],
                               "call": "rankfm_b200._rankfm._fit on every rank (pageable host ndarrays)",
                               "note": "a one-shot call builds and destroys its own NCCL communicator: ~1.3 s of every call at N=2 (profiles/r01_bench_cfg2_2gpu_final.json) "
                                       "against %.0f ms of training -- callers that train repeatedly keep a Session (and its communicator), which is what `value` times" % (ms / args.steps)}
        except Exception as exc:
            if line is not None:
                line["e2e"] = {"error": repr(exc)}
        dog.cancel()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        emit(line)


def dram_resident_probe(device=0, steps=2):
    """the same SGD kernel on a workload whose tables do NOT fit L2 (cfg4m: 1.25M users x 1M items x 16M interactions,
    factors=128, BPR -- a 1/31 slice of BASELINE.json configs[3] with the same row sizes), so that the HBM roofline
    fraction of the kernel is visible next to the L2-resident headline workload"""
    from rankfm_b200 import _rankfm
    c = make_workload("cfg4m")
    X, N, epochs = c["X"], len(c["X"]), c["epochs"]
    ui = _rankfm.UserItems.from_interactions(X, c["U_global"])
    w = fresh_weights(c)
    keep = []
    prob = _rankfm.fit_problem(X, c["sw"], ui, c["x_uf"], c["x_if"], *[w[k] for k in WEIGHTS], HYPER["alpha"], HYPER["beta"], HYPER["learning_rate"],
                               HYPER["learning_schedule"], HYPER["learning_exponent"], c["max_samples"], mode="production", seed=1492, keep=keep)
    prob.device = device
    sess = _rankfm.Session(prob, keep)
    sess.snapshot()
    sess.train(epochs)                                   # warm-up
    flat = []
    for _ in range(steps):
        sess.restore(); sess.flush_l2()
        flat += sess.train(epochs)
    # predict() on the same DRAM-resident tables: 8M random (user, item) pairs, one fat-row gather each side per pair
    rng = np.random.default_rng(3)
    n_pairs = 8_000_000
    pairs = np.stack([rng.integers(0, c["U_global"], n_pairs), rng.integers(0, c["I"], n_pairs)], axis=1).astype(np.float32)
    predict_ms = sess.time_predict(np.ascontiguousarray(pairs), iters=3)
    sess.close()
    kern_ms = sum(s["kernel_ms"] for s in flat)
    alg = sum(N * algorithmic_bytes_per_positive(c["F"], s["draws"] / N) for s in flat)
    peak, src = measured_peaks()
    achieved = alg / (kern_ms / 1e3) / 1e9
    predict_bytes = n_pairs * (8.0 * c["F"] + 4.0 + 8.0 + 4.0)           # v_u row + v_i row + w_i + the pair + the score
    predict = {"workload": "predict(): %d random pairs on the same tables" % n_pairs, "ms": predict_ms, "pairs_per_s": n_pairs / (predict_ms * 1e-3),
               "achieved": predict_bytes / (predict_ms * 1e-3) / 1e9, "unit": "GB/s", "frac": predict_bytes / (predict_ms * 1e-3) / 1e9 / peak, "kernel": "predict_kernel"}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_cfg4m.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    return {"workload": c["label"], "bound": "hbm", "kernel": "sgd_pipe_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_source": src, "algorithmic_bytes_per_launch": alg / len(flat), "launch_ms": kern_ms / len(flat),
            "interactions_per_s": N * len(flat) / (kern_ms / 1e3), "predict": predict}


def recommend_probe(n_users=65536, n_items_cat=262144, factors=128, topn=100, device=0, iters=3, exact_users=4096):
    """secondary measurement (BASELINE.json configs[4] shape; the default is scaled to milliseconds, `--workload cfg5` runs
    the full 1M x 1M): recommend() top-100 through the tcgen05 candidate GEMM.  TFLOP/s = 2*U*I*K / CUDA-event time of the
    GEMM + filter kernels (both passes, thresholds included); inputs resident in HBM."""
    from rankfm_b200 import _rankfm
    rng = np.random.default_rng(0)
    w = dict(w_i=rng.normal(0, 0.3, n_items_cat).astype(np.float32), w_if=np.zeros(1, np.float32),
             v_u=rng.standard_normal((n_users, factors), dtype=np.float32) * np.float32(0.1),
             v_i=rng.standard_normal((n_items_cat, factors), dtype=np.float32) * np.float32(0.1),
             v_uf=np.zeros((1, factors), np.float32), v_if=np.zeros((1, factors), np.float32))
    x_uf, x_if = np.zeros((n_users, 1), np.float32), np.zeros((n_items_cat, 1), np.float32)
    keep = []
    prob = _rankfm._problem(x_uf, x_if, *[w[k] for k in WEIGHTS], keep)
    prob.device = device
    sess = _rankfm.Session(prob, keep)
    users = np.arange(n_users, dtype=np.float32)
    os.environ["RANKFM_B200_RECOMMEND"] = "tc"
    ms, gemm_ms = sess.time_recommend(users, topn, False, iters=iters)
    tc_rows, tc_redone = sess.recommend_stats()
    sample = users[:256]
    fast = sess.recommend(sample, topn, False)
    os.environ["RANKFM_B200_RECOMMEND"] = "exact"
    exact = sess.recommend(sample, topn, False)
    ms_exact = None
    if exact_users:
        ms_exact, _ = sess.time_recommend(users[:exact_users], topn, False, iters=1)
    os.environ.pop("RANKFM_B200_RECOMMEND", None)
    sess.close()
    overlap = float(np.mean([len(set(a.tolist()) & set(b.tolist())) / topn for a, b in zip(fast, exact)]))
    flops = 2.0 * n_users * n_items_cat * factors
    peak = 1693.1
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
    except Exception:
        pass
    return {"workload": "recommend top-%d, %d users x %d items, factors=%d (tcgen05 bf16 candidate GEMM + exact fp32 re-score)" % (topn, n_users, n_items_cat, factors),
            "ms_total": ms, "ms_gemm_filter": gemm_ms, "tflops_gemm_filter": flops / (gemm_ms * 1e-3) / 1e12, "tflops_end_to_end": flops / (ms * 1e-3) / 1e12,
            "bf16_peak_tflops": peak, "frac_of_bf16_peak": flops / (gemm_ms * 1e-3) / 1e12 / peak, "users_per_s": n_users / (ms * 1e-3),
            "exact_fp32_path_users_per_s": (exact_users / (ms_exact * 1e-3)) if ms_exact else None, "topk_overlap_vs_exact": overlap,
            "rows_redone_on_exact_path": "%d of %d" % (tc_redone, tc_rows),
            "variant": {k: os.environ.get(k, "default") for k in ("RANKFM_B200_GEMM_MSUB", "RANKFM_B200_TAU_STRIDE")}}


def run_recommend(args):
    """--workload cfg5: BASELINE.json configs[4] at full size (recommend top-100 for 1M users over 1M items, factors=128);
    one JSON line, metric users/s, roofline = bf16 tensor peak"""
    from rankfm_b200 import _lib, _rankfm
    assert _lib.lib().rfm_device_count() > 0, "bench.py needs a CUDA device (no CPU fallback)"
    _rankfm.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    U = int(os.environ.get("BENCH_CFG5_USERS", 1_000_000))
    I = int(os.environ.get("BENCH_CFG5_ITEMS", 1_000_000))
    clocks = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    r = recommend_probe(n_users=U, n_items_cat=I, factors=128, topn=100, iters=max(1, args.steps), exact_users=1024)
    clock_info = clocks.stop()
    line = {"metric": "recommend users/sec", "value": r["users_per_s"], "unit": "users/s", "n_gpus": 1, "steps": max(1, args.steps), "warmup": 1,
            "ms_per_step": r["ms_total"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 operands, f32 accumulate, f32 exact re-score",
            "data": "synthetic", "config": {"workload": r["workload"]},
            "roofline": {"bound": "tensor", "kernel": "score_filter_kernel (pass 1 + pass 2)", "achieved": r["tflops_gemm_filter"], "peak": r["bf16_peak_tflops"],
                         "unit": "TFLOP/s", "frac": r["frac_of_bf16_peak"], "traffic": None,
                         "note": "useful FLOPs 2*U*I*K over the CUDA-event time of both GEMM passes + threshold kernels"},
            "recommend": r, "clocks": clock_info}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(CONFIGS) + ["cfg5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-recommend", action="store_true", help="skip the secondary recommend() tensor-core measurement")
    ap.add_argument("--no-large", action="store_true", help="skip the DRAM-resident roofline probe (cfg4m)")
    args = ap.parse_args()
    quiet_stdout()
    if args.workload == "cfg5" and args.impl == "ours":
        run_recommend(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
