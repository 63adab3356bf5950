SKIP_NCU=1 PYTEST_ARGS="" bash tools/gpu_check.sh r05j > gpurun_out/r05j_stdout.txt 2>&1
OUT=gpurun_out/r05j
tail -4 $OUT/pytest_gpu.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/r05j/bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["seconds"], d["e2e"]["host_timeline_s"], d["clocks"])
print(json.dumps(d["solve"]["ms"]), json.dumps({k:v for k,v in d["edit"].items() if k in ("ms","first_call_ms","stages_ms","solve_paths")}))
print({k:v for k,v in d["edit"]["sequential"].items() if k.endswith("_ms") or k.endswith("per_edit")})
PY
