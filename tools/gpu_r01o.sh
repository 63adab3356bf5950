#!/bin/bash
# r01o: native key extraction + end-to-end edit timing
OUT=gpurun_out/r01o
mkdir -p $OUT
timeout 600 python bench.py --no-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit=$?" >> $OUT/bench.err
EMCID_NATIVE_KEYS=0 timeout 600 python bench.py --no-cpu --no-e2e --steps 2 > $OUT/bench_hf_keys.json 2> $OUT/bench_hf_keys.err
cat $OUT/bench.json; tail -3 $OUT/bench.err; cat $OUT/bench_hf_keys.json; tail -3 $OUT/bench_hf_keys.err
