#!/bin/bash
# r01m: panel-blocked diagonal-block factorisation
OUT=gpurun_out/r01m
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "exit=$?" >> $OUT/pytest_gpu.log
for cfg in "5 1000" "1 1000" "1 100" "5 100"; do
  timeout 120 python tools/run_solve_once.py $cfg 3 >> $OUT/solve_times_inverse.jsonl 2>> $OUT/solve.err
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $OUT/solve_launches.csv \
    python tools/run_solve_once.py 5 1000 1 > $OUT/ncu_solve.log 2>&1
python tools/summarize_launches.py $OUT/solve_launches.csv > $OUT/solve_launches_summary.txt 2>&1
tail -3 $OUT/pytest_gpu.log; cat $OUT/solve_times_inverse.jsonl; head -8 $OUT/solve_launches_summary.txt; tail -3 $OUT/solve.err
