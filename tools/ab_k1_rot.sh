# A/B of the K1 variants that move the rotate of the first n body rounds to the FMA pipe (PG_K1_ROTFMA=n).
set -x
for n in 0 1 2 3 5; do
  PG_K1_ROTFMA=$n timeout 200 python -m pytest tests/test_gpu_mash.py -x -q -k "cfg1 or fast_path_all" --timeout 120 --timeout-method thread 2>&1 | tail -1
  PG_K1_ROTFMA=$n timeout 200 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu --no-secondary 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ROTFMA=$n', d['ms_per_step'], d['value'], d['roofline']['frac'], d['parity'])"
done
