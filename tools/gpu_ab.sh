#!/bin/bash
# A/B of kernel variants on ONE box: short statistics-only bench runs (same clocks, same thermal state).
# usage: gpurun -- 'bash tools/gpu_ab.sh <tag> "<label>|<env assignments>" ...'
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for spec in "$@"; do
  label=${spec%%|*}; envs=${spec#*|}
  env $envs python bench.py --no-e2e --no-cpu --no-solve --steps 8 --warmup 3 > $OUT/ab_$label.json 2> $OUT/ab_$label.err
done
python - "$OUT" <<'PY'
import glob, json, os, sys
for f in sorted(glob.glob(os.path.join(sys.argv[1], "ab_*.json"))):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    fk = d["roofline"]["forward_kernels"]
    print(os.path.basename(f), round(d["value"]), round(d["ms_per_step"], 2), d["clocks"]["sm_mhz"],
          (d.get("parity") or {}).get("max_rel_err"), round(d["roofline"]["avg_launch_ms"], 4),
          {k: round(v["avg_launch_ms"], 4) for k, v in fk.items()})
PY
