set -x
timeout 300 python -m pytest tests/test_gpu_mash.py -x -q -k "select or threshold or cfg3 or differential" --timeout 200 --timeout-method thread 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 5 --csv --log-file gpurun_out/r02_k2_launches.csv python tools/bench_secondary.py --only-k2 > /dev/null 2>&1
grep -v "^==" gpurun_out/r02_k2_launches.csv | awk -F'","' '{print substr($5,1,50), $NF}' | head -6
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sketch_thresh_select -c 1 -f -o gpurun_out/r02_k2t_select_v4 python tools/bench_secondary.py --only-k2 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sketch_thresh_walk -c 1 -f -o gpurun_out/r02_k2t_walk_v4 python tools/bench_secondary.py --only-k2 > /dev/null 2>&1
