# K2 record after the last select/walk change: full ncu captures of both stages and a compute-sanitizer pass
# (memcheck + racecheck) over the select-regime, large-sketch and join tests.  Every step bounded.
set -x
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sketch_thresh_select -c 1 -f -o gpurun_out/r02_k2t_select_v10 python tools/bench_secondary.py --only-k2 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sketch_thresh_walk -c 1 -f -o gpurun_out/r02_k2t_walk_v10 python tools/bench_secondary.py --only-k2 > /dev/null 2>&1
SEL='select or threshold or sparse or join or all_pairs or distance or beyond'
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_mash.py -x -q -k "$SEL" --timeout 800 --timeout-method thread 2>&1 | tail -8 ) > gpurun_out/r02_sanitizer.txt 2>&1
echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer.txt
( timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_mash.py -x -q -k "select or threshold or beyond" --timeout 1100 --timeout-method thread 2>&1 | tail -8 ) >> gpurun_out/r02_sanitizer.txt 2>&1
echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer.txt
cat gpurun_out/r02_sanitizer.txt
