# Multi-GPU re-verification with the shipped library (no reference arm): usage  bash tools/multi_run_short.sh N
N=${1:-8}
set -x
timeout 400 python -m pytest tests/test_gpu_multi.py -x -q --timeout 240 --timeout-method thread 2>&1 | tail -4
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    -m pytest tests/test_gpu_dist.py -x -q --timeout 240 --timeout-method thread 2>&1 | grep -E "passed|failed|Error" | tail -10
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 \
    bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 400 gpurun_out/bench_n$N.err; wc -c gpurun_out/bench_n$N.json
