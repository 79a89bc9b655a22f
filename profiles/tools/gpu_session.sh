#!/bin/bash
# One gpurun call: GPU parity suite, recommend variant sweep, default bench line, full-size cfg5, ncu captures.
# Everything lands under gpurun_out/ (scratch); summaries are copied into profiles/ by hand afterwards.
mkdir -p gpurun_out
S=gpurun_out/summary.txt
: > $S
timeout 700 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $S
tail -5 gpurun_out/pytest_gpu.log >> $S
: > gpurun_out/recommend_sweep.jsonl
for v in "2 8 head" "1 8 head" "2 8 stride"; do
  set -- $v
  RANKFM_B200_GEMM_MSUB=$1 RANKFM_B200_TAU_STRIDE=$2 RANKFM_B200_TAU_SUBSET=$3 timeout 120 python profiles/tools/recommend_sweep.py >> gpurun_out/recommend_sweep.jsonl 2>> gpurun_out/sweep.err
  echo "sweep msub=$1 stride=$2 subset=$3 rc=$?" >> $S
done
# the exact-path fallback is a performance cliff: only go to full size when the default variant served every row itself
if head -1 gpurun_out/recommend_sweep.jsonl | grep -q '"rows_redone_on_exact_path": "0 of'; then
  if [ -z "$SKIP_BENCH" ]; then
    timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?" >> $S
  fi
  timeout 300 python bench.py --workload cfg5 --steps 2 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "cfg5 rc=$?" >> $S
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_filter -c 4 -f -o gpurun_out/ncu_score_filter \
      python profiles/tools/recommend_sweep.py --small > gpurun_out/ncu_score_filter.log 2>&1; echo "ncu full rc=$?" >> $S
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_recommend.csv \
      python profiles/tools/recommend_sweep.py --small > gpurun_out/launches_recommend.log 2>&1; echo "ncu launches rc=$?" >> $S
else
  echo "default recommend variant fell back to the exact path: skipped bench/cfg5/ncu" >> $S
fi
cat $S
