#!/bin/bash
# One gpurun call: parity tests, bench (both arms), ncu launch list and one full capture of the top kernels.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
python -c 'import __graft_entry__ as g; g.build(); g.smoke()' > $OUT/smoke.log 2>&1; echo "smoke exit=$?" >> $OUT/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "exit=$?" >> $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit=$?" >> $OUT/bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --captions 1024 --no-e2e --no-cpu --no-solve > $OUT/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm3x -s 28 -c 8 -f -o $OUT/prof \
    python bench.py --steps 2 --warmup 1 --captions 1024 --no-e2e --no-cpu --no-solve > $OUT/ncu_full_bench.log 2>&1
fi
tail -3 $OUT/pytest_gpu.log; cat $OUT/smoke.log | tail -2; cat $OUT/bench.json; tail -2 $OUT/bench.err; cat $OUT/bench_reference.json
