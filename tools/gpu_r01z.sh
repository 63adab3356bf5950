#!/bin/bash
# r01z: SYRK on MN-major operand tiles (no transposed planes)
OUT=gpurun_out/r01z
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "exit=$?" >> $OUT/pytest_gpu.log
timeout 600 python bench.py --no-cpu --no-solve --no-e2e > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit=$?" >> $OUT/bench.err
EMCID_SYRK_MN=0 timeout 600 python bench.py --no-cpu --no-solve --no-e2e > $OUT/bench_kmajor.json 2> $OUT/bench_kmajor.err
tail -4 $OUT/pytest_gpu.log | cut -c1-300; for f in bench bench_kmajor; do python -c "
import json; d=json.loads(open('$OUT/$f.json').read()); r=d['roofline']; print('$f', d['value'], r['achieved'], r['avg_launch_ms'], {k:round(v['avg_launch_ms'],3) for k,v in r['forward_kernels'].items()})"; done; tail -2 $OUT/bench.err
