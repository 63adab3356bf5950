OUT=gpurun_out/r05l
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -k "bigg" > $OUT/pytest_bigg.log 2>&1; tail -3 $OUT/pytest_bigg.log
timeout 900 python bench.py --encoder sdxl-text2 --steps 4 --warmup 3 --no-solve > $OUT/bench_sdxl_text2.json 2> $OUT/bench_sdxl_text2.err; echo "exit=$?"
EMCID_LINEAR_CHUNK_SHORTK=2 timeout 900 python bench.py --encoder sdxl-text2 --steps 4 --warmup 3 --no-solve --no-e2e > $OUT/bench_sdxl_text2_chunk2.json 2> $OUT/bench_sdxl_text2_chunk2.err; echo "exit=$?"
python - <<'PY'
import json
for f in ("bench_sdxl_text2.json","bench_sdxl_text2_chunk2.json"):
    d=json.load(open("gpurun_out/r05l/"+f))
    print(f, d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["roofline"]["frac"], (d.get("e2e") or {}).get("value"))
    print({k:(round(v["avg_launch_ms"],4), round(v.get("issued_tflops",0))) for k,v in d["roofline"]["forward_kernels"].items()})
PY
