PYTEST_ARGS="" bash tools/gpu_check.sh r05g > gpurun_out/r05g_stdout.txt 2>&1
OUT=gpurun_out/r05g
python tools/profile_edit.py > $OUT/profile_edit.txt 2>&1
for spec in "1 1000" "1 100" "5 1000"; do echo "$spec: $(timeout 120 python tools/run_solve_once.py $spec 5 2>&1 | tail -1)" >> $OUT/solve_times.txt; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/solve1_launches.csv python tools/run_solve_once.py 1 1000 1 > $OUT/ncu_solve1.log 2>&1
python tools/summarize_launches.py $OUT/solve1_launches.csv > $OUT/solve1_launches_summary.txt 2>&1
tail -4 $OUT/pytest_gpu.log; cat $OUT/solve_times.txt; head -12 $OUT/launches_summary.txt
python - <<'PY'
import json
d=json.load(open("gpurun_out/r05g/bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["host_timeline_s"], d["clocks"])
print(json.dumps(d["solve"]["ms"]), json.dumps({k:v for k,v in d["edit"].items() if k in ("ms","first_call_ms","stages_ms","solve_paths")}))
print({k:v for k,v in d["edit"]["sequential"].items() if k.endswith("_ms") or k.endswith("per_edit")})
PY
