OUT=gpurun_out/r05e
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --durations=5 > $OUT/pytest_gpu.log 2>&1; echo "exit=$?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
bash tools/gpu_ab.sh r05e "pdl|EMCID_PDL=1" "nopdl|EMCID_PDL=0" "pdl2|EMCID_PDL=1" "nopdl2|EMCID_PDL=0"
timeout 600 python bench.py --no-solve --no-cpu > $OUT/bench_e2e.json 2> $OUT/bench_e2e.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r05e/bench_e2e.json"))
print(d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], json.dumps(d["e2e"]["host_timeline_s"]), d["e2e"]["value"], d["e2e"]["seconds"], d["parity"]["max_rel_err"])
PY
