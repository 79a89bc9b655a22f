#!/bin/bash
# Verification call: smoke(), full GPU suite, the default bench line + reference arm, cfg5 at full size, ncu launch list of
# the default bench command.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 700 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm rc=$?"
timeout 300 python bench.py --workload cfg5 --steps 2 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "cfg5 rc=$?"
if [ -n "$WITH_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_default.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-large > gpurun_out/launches_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_recommend.csv \
    python profiles/tools/recommend_sweep.py --small > gpurun_out/launches_recommend.log 2>&1; echo "ncu recommend launches rc=$?"
fi
