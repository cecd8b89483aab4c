#!/bin/bash
# A/B the K1 variants on one box: parity (pytest subset) then timing, one process per variant.
set -u
mkdir -p gpurun_out
for cfg in "2 0 32 0" "5 0 32 0" "3 0 32 0"; do
  set -- $cfg
  export PG_K1_VARIANT=$1 PG_K1_FM=$2 PG_K1_R=$3 PG_K1_CTAS=$4
  par=$(timeout 300 python -m pytest tests/test_gpu_mash.py -q -x -k "cfg1 or fast_path or full_size" 2>&1 | tail -1)
  out=$(timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu 2>&1 | tail -1)
  ms=$(echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4f ms  frac %.3f  %s clocks %s' % (d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel'], d['clocks']))" 2>&1)
  echo "variant=$1 fm=$2 R=$3 ctas=$4 :: $ms :: $par" | tee -a gpurun_out/tune_k1.log
done
