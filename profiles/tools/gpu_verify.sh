#!/bin/bash
# Verification call: full GPU suite, e2e phase probe, the default bench line, its ncu launch list.
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
E2E_PROBE_STEPS=40 RANKFM_B200_TIMING=1 timeout 200 python profiles/tools/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1; echo "e2e probe rc=$?"; grep -v "^\[rfm_fit\]" gpurun_out/e2e_probe.log | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm rc=$?"
if [ -n "$WITH_NCU" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_default.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-large > gpurun_out/launches_bench.log 2>&1; echo "ncu launches rc=$?"
fi
