# K2 check: the parity tests of the sketch path, timings of cfg3 / genomes / one genome / the large-sketch path, ncu launch list.
set -x
timeout 600 python -m pytest tests/test_gpu_mash.py -x -q --timeout 200 --timeout-method thread 2>&1 | tail -3
timeout 300 python tools/bench_secondary.py --only-k2 2>&1 | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 5 --csv --log-file gpurun_out/r02_k2_launches.csv python tools/bench_secondary.py --only-k2 > /dev/null 2>&1
grep -v "^==" gpurun_out/r02_k2_launches.csv | awk -F'","' '{print substr($5,1,50), $NF}' | head -6
