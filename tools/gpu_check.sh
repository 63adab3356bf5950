#!/bin/bash
# One gpurun call: smoke, GPU parity tests, bench (both arms) and — unless SKIP_NCU=1 — the ncu launch list and one full
# capture of the statistics step.   usage: gpurun --timeout 2400 -- 'bash tools/gpu_check.sh <tag>'
TAG=${1:-r04}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/smoke.log 2>&1; echo "smoke exit=$?" >> $OUT/smoke.log
timeout ${PYTEST_TIMEOUT:-1500} python -m pytest tests -m gpu -q ${PYTEST_ARGS--x} --durations=15 > $OUT/pytest_gpu.log 2>&1; echo "exit=$?" >> $OUT/pytest_gpu.log
if [ "${SKIP_BENCH:-0}" != "1" ]; then
timeout 900 python bench.py ${BENCH_ARGS:-} > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit=$?" >> $OUT/bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
NCU_CMD="python bench.py --steps 2 --warmup 1 --captions 984 --no-e2e --no-cpu --no-solve --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv \
    $NCU_CMD > $OUT/ncu_launch_bench.log 2>&1
python tools/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1
# full capture: the 5 SYRK launches of one edited layer's block plus the linear layers around them
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm3x|attention" -s 163 -c 14 -f -o $OUT/prof \
    $NCU_CMD > $OUT/ncu_full_bench.log 2>&1
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/prof_raw.csv 2> $OUT/prof_raw.err
python tools/summarize_ncu_raw.py $OUT/prof_raw.csv ncu --set full --clock-control none --import-source on -k "regex:gemm3x|attention" -s 163 -c 14 $NCU_CMD > $OUT/ncu_full_summary.json 2>> $OUT/prof_raw.err
fi
tail -25 $OUT/pytest_gpu.log; cat $OUT/smoke.log | tail -2; [ -f $OUT/bench.json ] && cat $OUT/bench.json; [ -f $OUT/bench.err ] && tail -5 $OUT/bench.err
[ -f $OUT/bench_reference.json ] && cat $OUT/bench_reference.json
if [ -f $OUT/launches_summary.txt ]; then head -14 $OUT/launches_summary.txt; fi
