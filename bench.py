#!/usr/bin/env python
"""bench.py -- training interactions/sec of the RankFM hot path (`_fit` epoch loop) on N x B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2] [--sub all|none|cfg3,cfg5]

A "step" is one full pass of the hot path over the workload: `epochs` SGD epochs over all interactions, starting from
the same initial weights every step.  Headline workload = BASELINE.json configs[1] (MovieLens-1M shape synthetic, 6040 x
3706, 1M interactions, factors=20, loss='warp', max_samples=20, 20 epochs).

  value   interactions/s (N * epochs * K / device time) with every input already resident in HBM when the timed region
          starts (CUDA events on the library's stream, max over ranks); the region holds, per step: D2D restore of the
          initial weights, an L2 flush (512 MB memset), `epochs` SGD kernel launches (+ the multi-GPU exchange kernel) +
          per-epoch weight-stat kernels
  e2e     the same metric through the reference-facing plug-in call `rankfm_b200._rankfm._fit(...)` on HOST buffers:
          H2D of interactions/CSR/weights, all epochs, D2H of the weights, wall clock, every rank on its shard
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md "Measurement"
  workloads   sub-records for the other BASELINE.json configurations, each with its own device-timed value, roofline and
          e2e: "cfg3" (1M x 200k x 50M, F=64, WARP, 8+8 side features: the largest single-GPU training configuration),
          "cfg4s" (one GPU's shard of configs[3]: 1.25M x 1M x 62.5M, F=128, BPR), "cfg5" (recommend top-100, 1M x 1M,
          F=128, tcgen05) at N=1; "cfg4" (configs[3] itself: every rank a 62.5M-interaction shard, 1M-item replicated
          table, one exchange per epoch) at N>1.  Their interactions are drawn on the GPU (rfm_synth_zipf).

Multi-GPU (torchrun, one process per GPU): weak scaling -- every rank owns a block of users (headline: a cfg2-sized block,
global U = 6040*N, 1M*N interactions, one shared item catalogue) and only those users' rows; item-side deltas are folded
once per epoch by the fused peer-memory kernel (NCCL fallback).  torch.distributed (gloo) is only the control plane here
(NCCL-id broadcast, barrier, max-reduce of the timings).

--impl reference times the UNMODIFIED reference Cython `_fit` / `_recommend` (oracle/_ref, built by oracle/build_ref.py)
on the host cores, single-threaded by construction (GIL held, no OpenMP), on bounded samples of the same workloads.
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

from rankfm_b200.synthetic import CONFIGS, zipf_interactions, zipf_interactions_device  # noqa: E402

WEIGHTS = ('w_i', 'w_if', 'v_u', 'v_i', 'v_uf', 'v_if')
HYPER = dict(alpha=0.01, beta=0.1, learning_rate=0.1, learning_schedule='invscaling', learning_exponent=0.25)
HYPER_ARGS = (HYPER["alpha"], HYPER["beta"], HYPER["learning_rate"], HYPER["learning_schedule"], HYPER["learning_exponent"])
# the reference's only published figure for this path and this configuration (MovieLens-1M,
# factors=20, warp, max_samples=20, invscaling, 20 epochs): 29.7 s wall for 749,724 x 20 interactions on a
# "2.3 GHz i5 MacBook", one thread (README.md:75-77, examples/movielens.ipynb:1073-1080).  Other hardware, real data.
PUBLISHED_CFG2_INTERACTIONS_PER_S = 749_724 * 20 / 29.7
DEVICE_GEN_MIN = 4_000_000          # workloads at least this large are drawn on the GPU (the NumPy generator takes minutes there)
ZIPF_U = float(os.environ.get("BENCH_ZIPF_U", 0.6))
ZIPF_I = float(os.environ.get("BENCH_ZIPF_I", 1.0))

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


def measured_bf16_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
    except Exception:
        return 1693.1, "fallback (B200_PROFILING.md)"


def algorithmic_bytes_per_positive(F, S, P=0, Q=0):
    """4F(S+5) + 4S + 28 (+ 4P + 4Q(1+S) with dense side features)"""
    return 4.0 * F * (S + 5.0) + 4.0 * S + 28.0 + (4.0 * P + 4.0 * Q * (1.0 + S) if (P or Q) else 0.0)


def table_note(c):
    mb = 4.0 * (c["U_rows"] * (c["F"] + c["P"]) + c["I"] * (c["F"] + 1 + c["Q"])) / 1e6
    if mb < 100:
        return ("this workload's tables (%.1f MB) live in the 126 MB L2: DRAM traffic is only the interaction stream, so the HBM fraction is low by "
                "construction (the L2, not HBM, is the roof here); workloads.cfg4s / cfg3 are the same kernel on tables that do not fit L2" % mb)
    return "tables %.0f MB per GPU (> 126 MB L2): rows come from DRAM except for the Zipf-hot ones" % mb


def stored_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from a committed `ncu --set full` capture of this round's
    kernel (profiles/traffic_<workload>.json); a STORED figure, not measured by this run (a profiler cannot run inside it)"""
    path = os.path.join(ROOT, "profiles", "traffic_%s.json" % workload)
    try:
        d = json.load(open(path))
        return d.get("dram_bytes_per_launch"), "stored ncu capture: %s" % d.get("source", os.path.basename(path))
    except Exception:
        return None, None


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


# -----------------------------------------------------------------------------------------------------------------
# workloads
# -----------------------------------------------------------------------------------------------------------------
def make_workload(name, rank=0, world=1, device=None, sample_n=None):
    """One rank's block of the named configuration: `U` users per rank (global U = U * world, the rank owns the users
    [rank*U, (rank+1)*U) and holds only their rows), one item catalogue shared by all ranks.

    `device` = CUDA ordinal -> large workloads are drawn on the GPU.  `sample_n` (reference arm) = draw only that many
    interactions over the SAME user / item id spaces (no re-indexing to the observed ids).
    -> dict: X int32 [n,2], sw, x_uf / x_if (global shapes), w0 (initial weights; `v_u` holds the OWNED rows only), ..."""
    c = dict(CONFIGS[name])
    U, I = c["U"], c["I"]
    n_draw = c["N"] if sample_n is None else min(c["N"], int(sample_n))
    lo = rank * U
    if device is not None and n_draw >= DEVICE_GEN_MIN:
        # ranks draw different interactions (seed) over one popularity ranking of the items (perm_seed)
        X, _, _ = zipf_interactions_device(U, I, n_draw, seed=42 + rank, a_u=ZIPF_U, a_i=ZIPF_I, offset_users=lo, device=device, perm_seed=42,
                                           reindex=(world == 1 and sample_n is None))
        c["generator"] = "rfm_synth_zipf (device)"
    else:
        X = zipf_interactions(U, I, n_draw, seed=42 + rank, a_u=ZIPF_U, a_i=ZIPF_I, reindex=sample_n is None)
        X[:, 0] += lo
        c["generator"] = "numpy"
    c["U_rows"] = U                                        # user rows held by this rank
    c["U_global"] = U * world
    c["user_range"] = (lo, lo + U) if world > 1 else None
    c["X"] = X
    c["sw"] = np.ones(len(X), np.float32)
    rng = np.random.default_rng(1000 + rank)
    # feature matrices keep the global [U_global, P] shape of the plug-in API; only the owned rows are ever read
    if c["P"]:
        c["x_uf"] = np.zeros((c["U_global"], c["P"]), np.float32)
        c["x_uf"][lo:lo + U] = rng.random((U, c["P"]), dtype=np.float32)
    else:
        c["x_uf"] = np.zeros((c["U_global"], 1), np.float32)
    c["x_if"] = np.random.default_rng(999).random((I, c["Q"]), dtype=np.float32) if c["Q"] else np.zeros((I, 1), np.float32)
    sigma, scale = np.float32(0.1), np.float32(0.01)       # N(0, sigma) factors, (alpha/beta) sigma for the feature factors (rankfm.py:223-244)
    ri = np.random.default_rng(7)                          # item side: identical on every rank
    F, P, Q = c["F"], c["P"], c["Q"]
    c["w0"] = dict(w_i=np.zeros(I, np.float32), w_if=np.zeros(max(Q, 1), np.float32),
                   v_u=rng.standard_normal((U, F), dtype=np.float32) * sigma,
                   v_i=ri.standard_normal((I, F), dtype=np.float32) * sigma,
                   v_uf=(ri.standard_normal((P, F), dtype=np.float32) * scale) if P else np.zeros((1, F), np.float32),
                   v_if=(ri.standard_normal((Q, F), dtype=np.float32) * scale) if Q else np.zeros((1, F), np.float32))
    return c


def alloc_weights(c):
    """the six weight arrays in the plug-in API's global shapes; v_u [U_global, F] is allocated lazily (np.empty) and only
    the rows this rank owns are ever written or read"""
    w = {k: v.copy() for k, v in c["w0"].items() if k != "v_u"}
    w["v_u"] = np.empty((c["U_global"], c["F"]), np.float32) if c["user_range"] else c["w0"]["v_u"].copy()
    reset_weights(c, w)
    return w


def reset_weights(c, w):
    lo, hi = c["user_range"] or (0, c["U_global"])
    for k in WEIGHTS:
        if k == "v_u":
            w[k][lo:hi] = c["w0"][k]
        else:
            w[k][...] = c["w0"][k]


# -----------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# -----------------------------------------------------------------------------------------------------------------
def reference_fit_rate(c, epochs, repeats=1):
    """interactions/s of the reference's own Cython `_fit` (single host thread) on this workload"""
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
        w = {k: v.copy() for k, v in c["w0"].items()}        # single process: the rank owns every user
        np.random.seed(0)
        t0 = time.perf_counter()
        fit(X, c["sw"], ui, c["x_uf"], c["x_if"], *[w[k] for k in WEIGHTS], *HYPER_ARGS, c["max_samples"], epochs, False)
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


def reference_recommend(steps, warmup):
    """the reference's own `_recommend` (scalar scoring loop + full argsort per user, `_rankfm.pyx:393-460`) on a bounded
    sample of users of the cfg5 model; users/s is scale-free in the user count"""
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
    for k in range(warmup + max(1, steps)):
        t0 = time.perf_counter()
        rec(users, {}, topn, False, zeros(U, 1), zeros(I, 1), w_i, zeros(1), v_u, v_i, zeros(1, F), zeros(1, F))
        if k >= warmup:
            times.append(time.perf_counter() - t0)
    value = len(users) / float(np.mean(times))
    return {"metric": "recommend users/sec", "value": value, "unit": "users/s", "ms_per_step": 1e3 * float(np.mean(times)),
            "config": {"workload": "recommend top-%d, %d users x %d items, factors=%d" % (topn, U, I, F),
                       "sample": "%d users per step (the reference's per-user cost does not depend on the number of users)" % len(users)},
            "cpu_baseline": {"value": value, "unit": "users/s", "cores": 1, "kind": kind,
                             "sample": "%d users x %d items per step; single-threaded by construction, 1 thread of %d" % (len(users), I, os.cpu_count())}}


def reference_training_record(name, budget_s, steps, warmup):
    """reference arm of one training configuration on a bounded sample; large configurations draw only the sample (same id
    spaces, so the tables -- and the cache behaviour of the random row accesses -- have the configuration's size)"""
    big = CONFIGS[name]["N"] >= DEVICE_GEN_MIN
    c = make_workload(name, sample_n=int(os.environ.get("BENCH_REF_SAMPLE_N", 400_000)) if big else None)
    cs, sample_epochs, sample_n = bounded_reference_sample(c, budget_s / max(1, steps + warmup))
    for _ in range(warmup):
        reference_fit_rate(cs, sample_epochs)
    t0 = time.perf_counter()
    rates = [reference_fit_rate(cs, sample_epochs) for _ in range(steps)]
    dt = time.perf_counter() - t0
    kind = rates[0][1]
    value = float(np.mean([r[0] for r in rates]))
    N = CONFIGS[name]["N"]
    note = None
    if big:
        # a sample of a few 100 k interactions over 1 M+ users is dominated by the reference's per-call Python set-up (one
        # ragged-table row per user, _rankfm.pyx:201-212), which a full-size run amortises: report the EPOCH rate instead,
        # (t(k epochs) - t(1 epoch)) / (k - 1) as SURVEY 8(d) prescribes -- the figure most favourable to the reference
        k = max(2, sample_epochs)
        t1 = len(cs["X"]) / reference_fit_rate(cs, 1)[0]
        tk = len(cs["X"]) * k / reference_fit_rate(cs, k)[0]
        if tk > t1:
            note = "epoch rate (t(%d epochs) - t(1 epoch)) / %d = per-call set-up excluded; whole-call rate %.0f interactions/s" % (k, k - 1, value)
            value = len(cs["X"]) * (k - 1) / (tk - t1)
    what = "all %d interactions" % N if sample_n == N else "%d of %d interactions (same users, items, factors, features; extrapolated)" % (sample_n, N)
    return {"metric": "training interactions/sec", "value": value, "unit": "interactions/s", "ms_per_step": 1e3 * dt / max(1, steps),
            "config": {"workload": c["label"], "sample": "%d of %d epochs per step, %s" % (sample_epochs, c["epochs"], what), "note": note},
            "cpu_baseline": {"value": value, "unit": "interactions/s", "cores": 1, "kind": kind,
                             "sample": "%d epochs x %d interactions per step; the reference holds the GIL and has no OpenMP: 1 thread of %d" % (sample_epochs, sample_n, os.cpu_count())}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    common = {"impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
              "dtype": "f32", "data": "synthetic"}
    if args.workload == "cfg5":
        r = reference_recommend(args.steps, args.warmup)
        r.update(common, vs_baseline=None, e2e={"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        return emit(r)
    r = reference_training_record(args.workload, float(os.environ.get("BENCH_REF_BUDGET_S", 120.0)), args.steps, args.warmup)
    r.update(common, vs_baseline=vs_published(r["value"], args.workload),
             e2e={"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    subs = sub_list(args, world=1)
    if subs:
        # the other BASELINE.json configurations our arm reports as `workloads`, on bounded samples (~10 s each)
        r["workloads"] = {}
        for name in subs:
            try:
                if name == "cfg5":
                    r["workloads"][name] = reference_recommend(1, 0)
                else:
                    r["workloads"][name] = reference_training_record(name, float(os.environ.get("BENCH_REF_SUB_BUDGET_S", 10.0)), 1, 0)
            except Exception as exc:
                r["workloads"][name] = {"error": repr(exc)}
    emit(r)


def sub_list(args, world):
    if args.workload != "cfg2" or args.sub == "none":
        return []
    default = ["cfg3", "cfg4s", "cfg5"] if world == 1 else ["cfg4"]
    if args.sub == "all":
        return default
    return [s for s in args.sub.split(",") if s]


# -----------------------------------------------------------------------------------------------------------------
# our arm
# -----------------------------------------------------------------------------------------------------------------
class Job:
    """process-wide context of one bench run: rank / world, the gloo control plane, the library"""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist                 # control plane only (gloo): id broadcast, barrier, max-reduce
            dist.init_process_group(backend="gloo", init_method="env://")
            self.dist = dist
        from rankfm_b200 import _lib, _rankfm
        assert _lib.lib().rfm_device_count() > 0, "bench.py needs a CUDA device (no CPU fallback)"
        self.rfm = _rankfm
        _rankfm.set_device(self.local_rank)
        self.nccl_id = None
        if self.world > 1:
            import torch
            idt = torch.zeros(128, dtype=torch.uint8)
            if self.rank == 0:
                idt = torch.frombuffer(bytearray(_rankfm.nccl_unique_id()), dtype=torch.uint8).clone()
            self.dist.broadcast(idt, src=0)
            self.nccl_id = idt.numpy().tobytes()

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max_over_ranks(self, x):
        if self.dist is None:
            return float(x)
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.int64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return int(t.item())


EXCHANGE_PATHS = {0: "none (single GPU)", 1: "fused peer-memory kernel over NVLink (cudaIpc windows; NCCL only for the bootstrap)", 2: "ncclAllReduce"}


def measure_training(job, name, steps, warmup, e2e_steps, sample_clocks=False):
    """device-timed throughput + roofline + e2e of one training configuration on this job's GPUs"""
    rfm, world = job.rfm, job.world
    c = make_workload(name, job.rank, world, device=job.local_rank)
    X, N, epochs = c["X"], len(c["X"]), c["epochs"]
    if world > 1:
        rfm.set_comm(job.rank, world, job.nccl_id, user_range=c["user_range"])
    t_prep = time.perf_counter()
    ui = rfm.UserItems.from_interactions(X, c["U_global"], c["I"])
    t_prep = time.perf_counter() - t_prep
    w = alloc_weights(c)
    keep = []
    prob = rfm.fit_problem(X, c["sw"], ui, c["x_uf"], c["x_if"], *[w[k] for k in WEIGHTS], *HYPER_ARGS, c["max_samples"],
                           mode="production", seed=1492, keep=keep, user_range=c["user_range"])
    sess = rfm.Session(prob, keep)
    sess.snapshot()
    exchange = EXCHANGE_PATHS[sess.exchange_path()]

    def one_step():
        sess.restore()
        sess.flush_l2()
        return sess.train(epochs)

    for _ in range(warmup):
        one_step()
    clocks = ClockSampler(job.local_rank) if (sample_clocks and job.rank == 0) else None
    launches0 = sess.launch_count()
    job.barrier()
    sess.timer_start()                                   # synchronises the stream, then records the start event
    all_stats = [one_step() for _ in range(steps)]
    ms = sess.timer_stop()                               # records + synchronises the stop event
    job.barrier()
    launches = job.sum_over_ranks(sess.launch_count() - launches0)
    clock_info = clocks.stop() if clocks else None
    ms = job.max_over_ranks(ms)
    value = N * epochs * steps * world / (ms / 1e3)

    # ---- roofline of the dominant kernel (sgd_pipe_kernel), from the per-launch CUDA events of the timed steps ----
    flat = [s for step in all_stats for s in step]
    kern_ms = sum(s["kernel_ms"] for s in flat)
    sync_ms = sum(s["sync_ms"] for s in flat)
    # with world > 1 the library reports whole-job draws; this rank's kernel handled 1/world of them (equal shards)
    alg_bytes = sum(N * algorithmic_bytes_per_positive(c["F"], s["draws"] / (N * world), c["P"], c["Q"]) for s in flat)
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (kern_ms / 1e3) / 1e9
    traffic, traffic_src = stored_traffic(name)
    roofline = {"bound": "hbm", "kernel": "sgd_pipe_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes / len(flat),
                "launch_ms": kern_ms / len(flat), "kernel_share_of_step": kern_ms / max(ms, 1e-9),
                "mean_draws_per_positive": float(np.mean([s["draws"] / (N * world) for s in flat])),
                "note": table_note(c)}
    exchange_rec = None
    if world > 1:
        # one exchange per epoch: rows [I r/C, I (r+1)/C) of every replica are read and the folded rows written to every replica
        nq = (c["F"] + 3) // 4 + 1
        moved = 2.0 * c["I"] * nq * 16.0 * (world - 1) / world           # bytes this rank reads from + writes to its peers per epoch
        exchange_rec = {"path": exchange, "ms_per_epoch": sync_ms / len(flat), "share_of_step": sync_ms / max(ms, 1e-9),
                        "peer_bytes_per_epoch_per_gpu": moved, "peer_gbs": moved / max(sync_ms / len(flat), 1e-9) / 1e6,
                        "table_bytes": c["I"] * nq * 16, "sgd_kernel_ms_per_epoch": kern_ms / len(flat),
                        "limiting_op": "sgd_pipe_kernel (HBM gather/scatter)" if kern_ms > 4 * sync_ms else "exchange_kernel (two in-kernel barriers + peer reads/writes over NVLink)"}
    sess.close()

    # ---- e2e: the plug-in call `_fit` on HOST buffers, every rank on its shard (H2D + epochs + exchange + D2H inside the
    # timed region; wall clock, max over ranks).  `e2e` = stateless calls: every call uploads interactions, CSR, features and
    # weights; `e2e_resident` = the fit_partial pattern: the plug-in keeps the session of the last call, a call moves only
    # the weight arrays.  Host buffers are page-locked in place first (rfm_host_register).
    e2e, e2e_resident = measure_e2e(job, c, X, ui, w, epochs, e2e_steps)
    rec = {"metric": "training interactions/sec", "value": value, "unit": "interactions/s", "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "dtype": "f32", "data": "synthetic",
           "config": {"workload": c["label"], "interactions_per_gpu": N, "epochs_per_step": epochs, "users": c["U_global"], "items": c["I"],
                      "l2": "flushed between steps (512 MB memset inside the timed region)", "generator": c["generator"],
                      "parallelism": "user-sharded x%d (each rank holds only its users' rows), per-epoch fold of item deltas: %s" % (world, exchange) if world > 1 else "single GPU",
                      "schedule": "production: Hogwild lane-group per positive, Philox negatives, on-device Feistel order",
                      "host_prep_user_items_s": round(t_prep, 3)},
           "roofline": roofline, "e2e": e2e, "e2e_resident": e2e_resident, "gpu_launches": int(launches), "clocks": clock_info,
           "exchange": exchange_rec, "final_log_likelihood": flat[-1]["log_likelihood"]}
    return rec, c


def fit_phases():
    """{create+H2D, train, D2H, destroy} milliseconds of this thread's last rfm_fit (None before the first call)"""
    import ctypes as C
    from rankfm_b200 import _lib
    out = (C.c_double * 4)()
    try:
        _lib.lib().rfm_last_fit_phases(out)
    except Exception:
        return None
    return [float(v) for v in out]


def measure_e2e(job, c, X, ui, w, epochs, steps):
    rfm, world = job.rfm, job.world
    N = len(X)
    lo, hi = c["user_range"] or (0, c["U_global"])
    pinned = [X, c["sw"], ui.indptr, ui.indices] + [w[k] if k != "v_u" else w[k][lo:hi] for k in WEIGHTS]
    rfm.pin(*pinned)

    phases = []

    def step():
        reset_weights(c, w)
        job.barrier()
        t0 = time.perf_counter()
        rfm._fit(X, c["sw"], ui, c["x_uf"], c["x_if"], *[w[k] for k in WEIGHTS], *HYPER_ARGS, c["max_samples"], epochs, False)
        dt = time.perf_counter() - t0
        phases.append(fit_phases())
        return job.max_over_ranks(dt)

    w_bytes = sum(v.nbytes for v in c["w0"].values())                    # v_u: the owned rows only
    data_bytes = X.nbytes + c["sw"].nbytes + (hi - lo + 1) * 8 + int(ui.indptr[hi] - ui.indptr[lo]) * 4 + \
        ((hi - lo) * c["P"] * 4 if c["P"] else 0) + (c["x_if"].nbytes if c["Q"] else 0)
    out = []
    for resident, warm, n in ((False, min(10, steps), steps), (True, min(3, steps), max(3, steps // 2))):
        rfm.set_resident_training(resident)
        for _ in range(warm):                             # the block cache / lazy module loading settle over the first calls
            step()
        del phases[:]
        dts = [step() for _ in range(n)]
        dt_med = float(np.median(dts))
        # where a stateless call spends its wall time on THIS rank (library clock): median call and slowest call
        ph = None
        if not resident and phases and phases[0] is not None:
            order = np.argsort(dts)
            names = ("create_h2d_ms", "train_ms", "d2h_ms", "destroy_ms")
            ph = {"median_call": dict(zip(names, [round(v, 2) for v in phases[order[len(order) // 2]]])),
                  "slowest_call": dict(zip(names, [round(v, 2) for v in phases[order[-1]]]))}
        # wall-clock steps on a shared host see occasional scheduling hiccups: the value is the MEDIAN step, every sample
        # and the mean are reported next to it
        out.append({"value": N * epochs * world / dt_med, "unit": "interactions/s",
                    "h2d_bytes_per_step": int((w_bytes + (0 if resident else data_bytes)) * world), "d2h_bytes_per_step": int(w_bytes * world),
                    "ms_per_step": 1e3 * dt_med, "statistic": "median of %d steps%s" % (len(dts), ", max over ranks" if world > 1 else ""),
                    "mean_ms_per_step": 1e3 * float(np.mean(dts)), "ms_each": [round(1e3 * d, 2) for d in dts], "library_phases": ph,
                    "call": "rankfm_b200._rankfm._fit(page-locked host ndarray buffers) -> ctypes -> C ABI, on every rank; " +
                            ("the plug-in keeps the training session of the previous call (fit_partial pattern): only the weight arrays move"
                             if resident else "stateless: every call uploads interactions, user_items CSR, features and weights; the job's communicator is cached by the library")})
    rfm.set_resident_training(True)
    rfm.drop_training()
    rfm.unpin(*pinned)
    return out[0], out[1]


def run_ours(args):
    job = Job()
    rec, c = measure_training(job, args.workload, args.steps, args.warmup, e2e_steps=max(15, args.steps) if args.workload == "cfg2" else 3, sample_clocks=True)
    line = None
    if job.rank == 0:
        cpu = None
        if job.world == 1 and not args.no_cpu_baseline:
            cr = make_workload(args.workload, sample_n=int(os.environ.get("BENCH_REF_SAMPLE_N", 400_000))) if c["N"] >= DEVICE_GEN_MIN else c
            cs, sample_epochs, sample_n = bounded_reference_sample(cr, float(os.environ.get("BENCH_CPU_BASELINE_S", 20.0)))
            rate, kind = reference_fit_rate(cs, sample_epochs)
            cpu = {"value": rate, "unit": "interactions/s", "cores": 1, "kind": kind,
                   "sample": "%d of %d epochs x %d of %d interactions; single-threaded by construction, %d host cores present" % (sample_epochs, c["epochs"], sample_n, c["N"], os.cpu_count())}
        line = dict(rec)
        line["vs_baseline"] = vs_published(rec["value"], args.workload)
        line["config"]["published_baseline"] = "%.0f interactions/s = reference README/notebook, MovieLens-1M (real data, same shape/hyper-parameters), 2.3 GHz i5 MacBook, 1 thread" % PUBLISHED_CFG2_INTERACTIONS_PER_S
        line["cpu_baseline"] = cpu
    # ---- the other BASELINE.json configurations ----
    subs = sub_list(args, job.world)
    workloads = {}
    for name in subs:
        try:
            if name == "cfg5":
                if job.rank == 0:
                    workloads[name] = recommend_record(full=True, device=job.local_rank, iters=2)
            elif name == "cfg4":
                r, _ = measure_training(job, "cfg4s", steps=2, warmup=1, e2e_steps=2)
                r["config"]["workload"] = "BASELINE.json configs[3] on %d GPUs (%d x 62.5M interactions, %d x 1.25M users, 1M items replicated, factors=128, bpr): one shard per GPU, one exchange per epoch" % (job.world, job.world, job.world)
                workloads[name] = r
            else:
                r, _ = measure_training(job, name, steps=2, warmup=1, e2e_steps=2)
                workloads[name] = r
        except Exception as exc:                        # a secondary measurement must never take the headline down
            workloads[name] = {"error": repr(exc)}
            if job.world > 1:
                raise
    if line is not None:
        if subs:
            line["workloads"] = workloads
        if job.world == 1 and not args.no_recommend and args.workload == "cfg2":
            try:
                line["recommend"] = recommend_record(full=False, device=job.local_rank)["recommend"]
            except Exception as exc:
                line["recommend"] = {"error": repr(exc)}
        if job.world == 1 and args.workload == "cfg2":
            try:
                line["e2e_class_fit"] = class_fit_record(c)
            except Exception as exc:
                line["e2e_class_fit"] = {"error": repr(exc)}
    if job.dist is not None:
        job.barrier()
        job.rfm.release_comms()
        job.dist.destroy_process_group()
    if line is not None:
        emit(line)


def class_fit_record(c):
    """the call a user of the reference makes: `RankFM(...).fit(interactions, epochs=...)` on RAW (user_id, item_id) pairs --
    id maps, index pairs, `user_items` (`RankFM._init_all`, rankfm.py:100-137: device radix sorts here, pandas in the
    reference), weight initialisation (NumPy, the reference's draw order), then the plug-in `_fit`; and a following
    `fit_partial` on the same interactions (warm start)"""
    from rankfm_b200 import RankFM
    X, epochs = c["X"], c["epochs"]
    rng = np.random.default_rng(3)
    uid = np.sort(rng.choice(10**9, c["U_global"], replace=False))         # scattered 64-bit ids: the maps do real work
    iid = np.sort(rng.choice(10**8, c["I"], replace=False))
    inter = np.ascontiguousarray(np.stack([uid[X[:, 0]], iid[X[:, 1]]], axis=1))
    loss = "bpr" if c["max_samples"] == 1 else "warp"

    def one():
        model = RankFM(factors=c["F"], loss=loss, max_samples=max(c["max_samples"], 1), alpha=HYPER["alpha"], beta=HYPER["beta"],
                       learning_rate=HYPER["learning_rate"], learning_schedule=HYPER["learning_schedule"], learning_exponent=HYPER["learning_exponent"])
        np.random.seed(0)
        t0 = time.perf_counter()
        model._init_all(inter)
        t_init = time.perf_counter() - t0
        model._reset_state()
        np.random.seed(0)
        t0 = time.perf_counter()
        model.fit(inter, epochs=epochs)
        t_fit = time.perf_counter() - t0
        t0 = time.perf_counter()
        model.fit_partial(inter, epochs=epochs)
        t_partial = time.perf_counter() - t0
        return t_init, t_fit, t_partial
    one()                                                                  # warm-up (lazy loading, block cache)
    runs = [one() for _ in range(3)]
    t_init, t_fit, t_partial = [float(np.median([r[k] for r in runs])) for k in range(3)]
    N = len(X)
    return {"call": "rankfm_b200.RankFM(...).fit(raw int64 id pairs [N,2], epochs=%d), then fit_partial on the same pairs" % epochs,
            "value": N * epochs / t_fit, "unit": "interactions/s", "ms_fit": 1e3 * t_fit, "ms_init_all_alone": 1e3 * t_init, "ms_fit_partial": 1e3 * t_partial,
            "statistic": "median of 3", "h2d_bytes_per_step": int(inter.nbytes + X.nbytes + c["sw"].nbytes), "d2h_bytes_per_step": int(X.nbytes)}


def recommend_probe(n_users=65536, n_items_cat=262144, factors=128, topn=100, device=0, iters=3, exact_users=4096):
    """recommend() top-100 through the tcgen05 candidate GEMM (BASELINE.json configs[4] shape; the default is scaled to
    milliseconds).  TFLOP/s = 2*U*I*K / CUDA-event time of the GEMM + filter kernels (both passes, thresholds included);
    inputs resident in HBM."""
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
    tc_retried = sess.recommend_retried()
    # e2e: the call a user makes -- host user indexes in, host item indexes out (H2D of the user list, D2H of the top-n table)
    t0 = time.perf_counter()
    sess.recommend(users, topn, False)
    e2e_s = time.perf_counter() - t0
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
    peak, peak_src = measured_bf16_peak()
    return {"workload": "recommend top-%d, %d users x %d items, factors=%d (tcgen05 bf16 candidate GEMM + exact fp32 re-score)" % (topn, n_users, n_items_cat, factors),
            "ms_total": ms, "ms_gemm_filter": gemm_ms, "tflops_gemm_filter": flops / (gemm_ms * 1e-3) / 1e12, "tflops_end_to_end": flops / (ms * 1e-3) / 1e12,
            "bf16_peak_tflops": peak, "peak_source": peak_src, "frac_of_bf16_peak": flops / (gemm_ms * 1e-3) / 1e12 / peak, "users_per_s": n_users / (ms * 1e-3),
            "e2e_users_per_s": n_users / e2e_s, "e2e_ms": 1e3 * e2e_s, "h2d_bytes": int(users.nbytes), "d2h_bytes": int(n_users * topn * 4),
            "exact_fp32_path_users_per_s": (exact_users / (ms_exact * 1e-3)) if ms_exact else None, "topk_overlap_vs_exact": overlap,
            "rows_redone_on_exact_path": "%d of %d" % (tc_redone, tc_rows),
            "rows_redone_with_provable_threshold": "%d of %d" % (tc_retried, tc_rows),
            "variant": {k: os.environ.get(k, "default") for k in ("RANKFM_B200_GEMM_MSUB", "RANKFM_B200_GEMM_EW", "RANKFM_B200_TAU_MODE", "RANKFM_B200_TAU_STRIDE",
                                                                  "RANKFM_B200_TAU_MIN", "RANKFM_B200_TAU_Z")}}


def recommend_record(full, device, iters=3):
    """`full`: BASELINE.json configs[4] at full size (recommend top-100 for 1M users over 1M items, factors=128)"""
    if full:
        U = int(os.environ.get("BENCH_CFG5_USERS", 1_000_000))
        I = int(os.environ.get("BENCH_CFG5_ITEMS", 1_000_000))
        r = recommend_probe(n_users=U, n_items_cat=I, factors=128, topn=100, device=device, iters=max(1, iters), exact_users=1024)
    else:
        r = recommend_probe(device=device, iters=iters)
    return {"metric": "recommend users/sec", "value": r["users_per_s"], "unit": "users/s", "n_gpus": 1, "steps": max(1, iters), "warmup": 1,
            "ms_per_step": r["ms_total"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 operands, f32 accumulate, f32 exact re-score",
            "data": "synthetic", "config": {"workload": r["workload"]},
            "roofline": {"bound": "tensor", "kernel": "score_filter_kernel (pass 1 + pass 2)", "achieved": r["tflops_gemm_filter"], "peak": r["bf16_peak_tflops"],
                         "unit": "TFLOP/s", "frac": r["frac_of_bf16_peak"], "traffic": None, "peak_source": r["peak_source"],
                         "note": "useful FLOPs 2*U*I*K over the CUDA-event time of both GEMM passes + threshold kernels; pass 2 is bound by the TMEM read of its fp32 "
                                 "accumulators (tcgen05.ld ~64 B/clk/SM, measured 61): with K=128 that caps it at 0.5 of the tensor pipe's rate (DESIGN.md 3.4)"},
            "e2e": {"value": r["e2e_users_per_s"], "unit": "users/s", "h2d_bytes_per_step": r["h2d_bytes"], "d2h_bytes_per_step": r["d2h_bytes"], "ms_per_step": r["e2e_ms"],
                    "call": "Session.recommend(host float32 user indexes) -> host float32 [users, 100] item indexes"},
            "recommend": r}


def run_recommend(args):
    """--workload cfg5: one JSON line, metric users/s, roofline = bf16 tensor peak"""
    from rankfm_b200 import _lib, _rankfm
    assert _lib.lib().rfm_device_count() > 0, "bench.py needs a CUDA device (no CPU fallback)"
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    _rankfm.set_device(dev)
    clocks = ClockSampler(dev)
    line = recommend_record(full=True, device=dev, iters=max(1, args.steps))
    line["clocks"] = clocks.stop()
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(CONFIGS) + ["cfg5"])
    ap.add_argument("--sub", default="all", help="sub-records of the default run: all | none | comma list of cfg3,cfg4s,cfg5 (N=1) / cfg4 (N>1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-recommend", action="store_true", help="skip the small recommend() tensor-core probe")
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
