set -x
for r in 0 1; do
  PG_K2T_ROLLED=$r timeout 300 python -m pytest tests/test_gpu_mash.py -x -q -k "select or threshold or cfg3 or differential" --timeout 200 --timeout-method thread 2>&1 | tail -1
  PG_K2T_ROLLED=$r timeout 300 python tools/bench_secondary.py --only-k2 2>&1 | cut -c1-140
done
PG_K2T_ROLLED=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4 --csv --log-file gpurun_out/r02_k2_rolled.csv python tools/bench_secondary.py --only-k2 > /dev/null 2>&1
grep -v "^==" gpurun_out/r02_k2_rolled.csv | awk -F'","' '{print substr($5,1,50), $NF}' | head -5
