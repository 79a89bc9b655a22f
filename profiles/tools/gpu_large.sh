#!/bin/bash
# Full-size single-GPU runs of BASELINE.json configs[2] (cfg3: 1M x 200k, 50M interactions, F=64, WARP, 8+8 features) and
# of one GPU's shard of configs[3] (cfg4s: 1.25M x 1M, 62.5M interactions, F=128, BPR).  Host data generation dominates
# the wall clock (~1 min each); the reference arm is skipped (it runs at ~0.1 M interactions/s on these shapes).
mkdir -p gpurun_out
for w in cfg3 cfg4s; do
  timeout 500 python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-recommend --no-large > gpurun_out/bench_${w}_full.json 2> gpurun_out/bench_${w}_full.err
  echo "$w rc=$?"; tail -c 600 gpurun_out/bench_${w}_full.json; tail -3 gpurun_out/bench_${w}_full.err
done
