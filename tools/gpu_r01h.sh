#!/bin/bash
# r01h: parity suite, bench (default + CTA pairs), e2e stage probe
OUT=gpurun_out/r01h
mkdir -p $OUT
python -c 'import __graft_entry__ as g; g.smoke()' > $OUT/smoke.log 2>&1; echo "smoke exit=$?" >> $OUT/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "exit=$?" >> $OUT/pytest_gpu.log
timeout 300 python tools/probe_e2e.py 16384 512 2 > $OUT/probe_e2e.json 2> $OUT/probe_e2e.err
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit=$?" >> $OUT/bench.err
EMCID_CTA2=1 timeout 600 python bench.py --no-cpu --no-solve > $OUT/bench_cta2.json 2> $OUT/bench_cta2.err
EMCID_CTA2=1 timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_cta2.log 2>&1; echo "exit=$?" >> $OUT/pytest_gpu_cta2.log
tail -3 $OUT/pytest_gpu.log; tail -2 $OUT/smoke.log; cat $OUT/probe_e2e.json; tail -3 $OUT/probe_e2e.err; cat $OUT/bench.json; cat $OUT/bench_cta2.json; tail -3 $OUT/pytest_gpu_cta2.log
