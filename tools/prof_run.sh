set -x
timeout 200 python -m pytest tests/test_gpu_mash.py -x -q -k "sparse or threshold" --timeout 120 --timeout-method thread 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_k2_launches.csv python tools/bench_secondary.py --only-k2 > gpurun_out/k2_under_ncu.log 2>&1
grep -v "^==" gpurun_out/r02_k2_launches.csv | cut -d, -f5,12- | cut -c1-200 | head -45
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sketch_thresh_walk -c 1 -f -o gpurun_out/r02_k2t_walk python tools/bench_secondary.py --only-k2 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sketch_thresh_select -c 1 -f -o gpurun_out/r02_k2t_select python tools/bench_secondary.py --only-k2 > /dev/null 2>&1
N=50000 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_k3_launches.csv python tools/prof_k3.py > gpurun_out/k3_under_ncu.log 2>&1
grep -v "^==" gpurun_out/r02_k3_launches.csv | cut -d, -f5,12- | cut -c1-200 | tail -16
N=50000 timeout 400 ncu --set full --clock-control none --import-source on -k regex:bucket_join -c 1 -f -o gpurun_out/r02_k3_join python tools/prof_k3.py > /dev/null 2>&1
timeout 200 python tools/prof_k3.py
ls -la gpurun_out
