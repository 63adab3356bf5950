SKIP_NCU=1 PYTEST_ARGS="" bash tools/gpu_check.sh r05a
OUT=gpurun_out/r05a
python tools/run_solve_once.py 1 1000 5 > $OUT/solve_1x1000.json 2>&1
python tools/run_solve_once.py 1 100 5 > $OUT/solve_1x100.json 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/solve1_launches.csv python tools/run_solve_once.py 1 1000 1 > $OUT/ncu_solve1.log 2>&1
python - <<'PY' > gpurun_out/r05a/solve1_launch_list.txt 2>&1
import csv
rows=[]
with open("gpurun_out/r05a/solve1_launches.csv") as f:
    lines=[l for l in f if not l.startswith("==")]
r=csv.DictReader(lines)
allr=[x for x in r]
# second half = the timed call (warm-up first)
names=[(x["Kernel Name"][:60], float(x["Metric Value"].replace(",",""))) for x in allr if x["Metric Name"]=="gpu__time_duration.sum"]
half=len(names)//2
tot=0
for nme,v in names[half:]:
    print(f"{v/1000:9.1f} us  {nme}")
    tot+=v
print("total us", tot/1000, "launches", len(names)-half)
PY
cat $OUT/solve_1x1000.json $OUT/solve_1x100.json; tail -3 $OUT/solve1_launch_list.txt
