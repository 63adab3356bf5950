SKIP_NCU=1 PYTEST_ARGS="" bash tools/gpu_check.sh r05d > gpurun_out/r05d_stdout.txt 2>&1
OUT=gpurun_out/r05d
python tools/profile_edit.py > $OUT/profile_edit.txt 2>&1
for spec in "1 1000" "1 100" "5 1000"; do echo "$spec: $(timeout 120 python tools/run_solve_once.py $spec 5 2>&1 | tail -1)" >> $OUT/solve_times.txt; done
tail -4 $OUT/pytest_gpu.log; cat $OUT/solve_times.txt
python - <<'PY'
import json
d=json.load(open("gpurun_out/r05d/bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
print(json.dumps(d["solve"]["ms"]), json.dumps({k:v for k,v in d["edit"].items() if k in ("ms","first_call_ms","stages_ms","solve_paths")}))
print(json.dumps(d["edit"]["sequential"]))
PY
