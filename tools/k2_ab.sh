# A/B of the select-stage knobs (batch width, bucket function): per-kernel times from the ncu launch list
set -x
timeout 600 python -m pytest tests/test_gpu_mash.py -x -q --timeout 200 --timeout-method thread 2>&1 | tail -2
for cfg in "0 1" "8 1" "5 0" "8 0" "4 0" "6 0"; do
  set -- $cfg
  echo "== PG_K2T_SEL_U=$1 PG_K2T_SEL_MUL=$2"
  PG_K2T_SEL_U=$1 PG_K2T_SEL_MUL=$2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 5 --csv --log-file gpurun_out/ab.csv python tools/bench_secondary.py --only-k2 > /dev/null 2>&1
  grep -v "^==" gpurun_out/ab.csv | awk -F'","' '{print substr($5,1,50), $NF}' | sed -n 3,4p
done
PG_K2T_SEL_MUL=0 timeout 300 python -m pytest tests/test_gpu_mash.py -x -q -k "select or threshold" --timeout 200 --timeout-method thread 2>&1 | tail -2
