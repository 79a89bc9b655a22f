#!/bin/bash
# Full-size single-GPU runs of BASELINE.json configs[2] (cfg3: 1M x 200k, 50M interactions, F=64, WARP, 8+8 features) and
# of one GPU's shard of configs[3] (cfg4s: 1.25M x 1M, 62.5M interactions, F=128, BPR).  Host data generation dominates
# the wall clock (~1 min each); the reference arm is skipped (it runs at ~0.1 M interactions/s on these shapes).
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout 200 -k "resident or degenerate or eighth" > gpurun_out/pytest_new.log 2>&1; echo "pytest(new) rc=$?"; tail -3 gpurun_out/pytest_new.log
E2E_PROBE_STEPS=40 RANKFM_B200_TIMING=1 timeout 200 python profiles/tools/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1; echo "e2e probe rc=$?"; grep -v "^\[rfm_fit\]" gpurun_out/e2e_probe.log | tail -2
for w in cfg3 cfg4s; do
  timeout 500 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-recommend --no-large > gpurun_out/bench_${w}_full.json 2> gpurun_out/bench_${w}_full.err
  echo "$w rc=$?"; tail -c 400 gpurun_out/bench_${w}_full.json; tail -3 gpurun_out/bench_${w}_full.err
done
