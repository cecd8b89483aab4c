set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 400 gpurun_out/bench_n1.err
python tools/bench_secondary.py --only-k2 > gpurun_out/k2_thresh.jsonl 2>&1; PG_K2_NO_THRESH=1 python tools/bench_secondary.py --only-k2 > gpurun_out/k2_nothresh.jsonl 2>&1
cat gpurun_out/k2_thresh.jsonl gpurun_out/k2_nothresh.jsonl | cut -c1-300
