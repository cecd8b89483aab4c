# N = 4 bench lines as the driver launches them (reference arm first), plus the C++ mirror tests
set -x
timeout 300 python -m pytest tests/test_gpu_hostcpp.py -x -q --timeout 240 --timeout-method thread 2>&1 | tail -3
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --impl reference --gpus 4 --steps 5 --warmup 1 > gpurun_out/bench_ref_n4.json 2> gpurun_out/bench_ref_n4.err
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 \
    bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
tail -c 400 gpurun_out/bench_n4.err; wc -c gpurun_out/bench_n4.json gpurun_out/bench_ref_n4.json
