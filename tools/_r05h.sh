OUT=gpurun_out/r05h
mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_dist.py > $OUT/check_dist_n2.log 2>&1; echo "exit=$?" >> $OUT/check_dist_n2.log
grep -v Warning $OUT/check_dist_n2.log | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cpu > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "exit=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r05h/bench_n2.json"))
print(d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["parity"]["max_rel_err"], d["parity"]["ok"])
e=d["e2e"]; print(e["value"], e["seconds"], e["host_timeline_s"]); print(e["strong"])
PY
