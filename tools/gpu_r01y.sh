#!/bin/bash
# r01y: key-extraction continuation
OUT=gpurun_out/r01y
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "exit=$?" >> $OUT/pytest_gpu.log
timeout 600 python bench.py --no-cpu --no-e2e > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit=$?" >> $OUT/bench.err
tail -5 $OUT/pytest_gpu.log; python -c "
import json; d=json.loads(open('$OUT/bench.json').read()); print(d['value'], d['solve']['ms'], d['edit'])"; tail -2 $OUT/bench.err
