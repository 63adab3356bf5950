#!/bin/bash
# r01t: bench with the per-kernel-class profile; OpenCLIP bigG shapes (BASELINE configs[3])
OUT=gpurun_out/r01t
mkdir -p $OUT
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit=$?" >> $OUT/bench.err
timeout 900 python bench.py --encoder sdxl-text2 --captions 1024 --block-captions 256 --steps 3 --no-cpu --e2e-captions 20000 > $OUT/bench_bigg.json 2> $OUT/bench_bigg.err; echo "exit=$?" >> $OUT/bench_bigg.err
cat $OUT/bench.json; tail -2 $OUT/bench.err; cat $OUT/bench_bigg.json; tail -3 $OUT/bench_bigg.err
