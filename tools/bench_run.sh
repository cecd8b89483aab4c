# the driver's two N=1 bench lines, reference arm first (as the driver runs them)
set -x
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_n1.json 2>/dev/null
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err
