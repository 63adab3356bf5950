OUT=gpurun_out/r05b
mkdir -p $OUT
for cfg in "-1" "0" "32" "16"; do
  for spec in "1 1000" "1 100" "5 1000" "2 1000"; do
    echo "M64_CTAS=$cfg $spec: $(EMCID_SOLVE_M64_CTAS=$cfg timeout 120 python tools/run_solve_once.py $spec 5 2>&1 | tail -1)" >> $OUT/solve_variants.txt
  done
done
cat $OUT/solve_variants.txt
timeout 900 python -m pytest tests -m gpu -q -x -k "solve or factor or edit or execute or sequential or cross_attn or smoke" > $OUT/pytest_solve.log 2>&1; tail -5 $OUT/pytest_solve.log
