OUT=gpurun_out/r05i
mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --no-cpu > $OUT/bench_n8.json 2> $OUT/bench_n8.err; echo "exit=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r05i/bench_n8.json"))
print(d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["parity"]["max_rel_err"], d["parity"]["ok"])
e=d["e2e"]; print(e["value"], e["seconds"], e["host_timeline_s"]); print(e["strong"])
print(d.get("solve",{}).get("ms"), d.get("solve",{}).get("placed"))
PY
tail -3 $OUT/bench_n8.err
