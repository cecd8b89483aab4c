set -x
timeout 200 python -m pytest tests/test_gpu_mash.py -x -q -k "sparse" --timeout 120 --timeout-method thread 2>&1 | tail -40
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sketch_thresh_walk -c 1 -f -o gpurun_out/r02_k2t_walk_v2 python tools/bench_secondary.py --only-k2 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sketch_thresh_select -c 1 -f -o gpurun_out/r02_k2t_select_v2 python tools/bench_secondary.py --only-k2 > /dev/null 2>&1
timeout 200 python -m pytest tests/test_gpu_mash.py -x -q -k "distance or join" --timeout 120 --timeout-method thread 2>&1 | tail -5
N=50000 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_k3_launches_v3.csv python tools/prof_k3.py > gpurun_out/k3_under_ncu.log 2>&1
grep -v "^==" gpurun_out/r02_k3_launches_v3.csv | awk -F'","' '{print substr($5,1,50), $NF}' | tail -9
timeout 200 python tools/prof_k3.py
