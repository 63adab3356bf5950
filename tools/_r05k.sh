OUT=gpurun_out/r05k
mkdir -p $OUT
timeout 900 python bench.py --encoder sdxl-text2 --steps 4 --warmup 3 > $OUT/bench_sdxl_text2.json 2> $OUT/bench_sdxl_text2.err; echo "exit=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r05k/bench_sdxl_text2.json"))
print(d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["roofline"]["achieved"], d["roofline"]["frac"], (d.get("e2e") or {}).get("value"), (d.get("solve") or {}).get("ms"))
print({k:(round(v["avg_launch_ms"],4), round(v.get("issued_tflops",0))) for k,v in d["roofline"]["forward_kernels"].items()})
PY
tail -3 $OUT/bench_sdxl_text2.err
