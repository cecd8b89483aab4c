# K2 check after a select/walk change: parity tests, timings, launch list, full captures of both stages,
# and a compute-sanitizer pass (memcheck + racecheck) over the select-regime and join tests.  Every step bounded.
set -x
timeout 600 python -m pytest tests/test_gpu_mash.py -x -q --timeout 200 --timeout-method thread 2>&1 | tail -3
timeout 300 python tools/bench_secondary.py --only-k2 2>&1 | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 5 --csv --log-file gpurun_out/r02_k2_launches.csv python tools/bench_secondary.py --only-k2 > /dev/null 2>&1
grep -v "^==" gpurun_out/r02_k2_launches.csv | awk -F'","' '{print substr($5,1,50), $NF}' | head -6
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sketch_thresh_select -c 1 -f -o gpurun_out/r02_k2t_select_v9 python tools/bench_secondary.py --only-k2 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sketch_thresh_walk -c 1 -f -o gpurun_out/r02_k2t_walk_v9 python tools/bench_secondary.py --only-k2 > /dev/null 2>&1
SEL='select or threshold or sparse or join or all_pairs or distance'
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_mash.py -x -q -k "$SEL" --timeout 500 --timeout-method thread 2>&1 | tail -8 ) > gpurun_out/r02_sanitizer.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer.txt
( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_mash.py -x -q -k "select or threshold" --timeout 800 --timeout-method thread 2>&1 | tail -8 ) >> gpurun_out/r02_sanitizer.txt 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer.txt
cat gpurun_out/r02_sanitizer.txt
