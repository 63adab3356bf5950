#!/bin/bash
# quick check: parity suite + device-resident bench
OUT=gpurun_out/${1:-quick}
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "exit=$?" >> $OUT/pytest_gpu.log
timeout 600 python bench.py --no-cpu --no-solve --no-e2e > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit=$?" >> $OUT/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --captions 984 --no-e2e --no-cpu --no-solve > $OUT/ncu_launch_bench.log 2>&1
python tools/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1
tail -3 $OUT/pytest_gpu.log; cat $OUT/bench.json | cut -c1-200; tail -3 $OUT/bench.err; head -10 $OUT/launches_summary.txt
