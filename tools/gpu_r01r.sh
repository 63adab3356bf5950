#!/bin/bash
# r01r: buffer pool (no cudaFree inside the pass), attention with 3 CTAs per SM
OUT=gpurun_out/r01r
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "exit=$?" >> $OUT/pytest_gpu.log
timeout 600 python bench.py --no-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit=$?" >> $OUT/bench.err
timeout 600 python bench.py --no-cpu --no-solve > $OUT/bench2.json 2> $OUT/bench2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --captions 1024 --no-e2e --no-cpu --no-solve > $OUT/ncu_launch_bench.log 2>&1
python tools/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1
tail -3 $OUT/pytest_gpu.log; cat $OUT/bench.json; tail -3 $OUT/bench.err; cat $OUT/bench2.json; head -12 $OUT/launches_summary.txt
