#!/bin/bash
# linear-layer TMEM chunk length: accuracy (hidden states, statistics) and speed
OUT=gpurun_out/chunk
mkdir -p $OUT
for c in 2 3 4; do
  EMCID_LINEAR_CHUNK=$c timeout 600 python -m pytest tests -m gpu -q -k "hidden_states or clipl_layer_stats or stats_properties or bigg_width or reference_fixture" > $OUT/pytest_$c.log 2>&1; echo "exit=$?" >> $OUT/pytest_$c.log
  EMCID_LINEAR_CHUNK=$c timeout 600 python bench.py --no-cpu --no-solve --no-e2e > $OUT/bench_$c.json 2> $OUT/bench_$c.err
  EMCID_LINEAR_CHUNK=$c timeout 300 python tools/probe_clip.py > $OUT/probe_clip_$c.log 2>&1
  echo "chunk $c: $(tail -2 $OUT/pytest_$c.log | head -1 | cut -c1-80) | $(python -c "import json; d=json.loads(open('$OUT/bench_$c.json').read()); print(d['value'], {k:round(v['avg_launch_ms'],3) for k,v in d['roofline']['forward_kernels'].items()})")"
  grep "native_vs_fp64\|native_L\|hook_L" $OUT/probe_clip_$c.log | cut -c1-700
done
