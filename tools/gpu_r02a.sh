#!/bin/bash
OUT=gpurun_out/r02a
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "exit=$?" >> $OUT/pytest_gpu.log
timeout 600 python tools/probe_ragged.py 60000 > $OUT/probe_ragged.json 2> $OUT/probe_ragged.err
tail -4 $OUT/pytest_gpu.log | cut -c1-300; cat $OUT/probe_ragged.json; tail -3 $OUT/probe_ragged.err
