# Multi-GPU pass on one box: usage  bash tools/multi_run.sh N   (N = 2, 4, 8).  Every step is bounded.
N=${1:-2}
set -x
nvidia-smi topo -m 2>/dev/null | head -12
# single process, N devices visible: the pg_*_multi entry points (peer stores across NVLink) + the C++ mirror
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_hostcpp.py -x -q --timeout 240 --timeout-method thread 2>&1 | tail -8
# one process per GPU: NCCL all-gather vs the fused sketch+gather (CUDA IPC), sharded pipeline vs the oracle
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    -m pytest tests/test_gpu_dist.py -x -q --timeout 240 --timeout-method thread 2>&1 | tail -12
# the driver's bench line at N ranks (reference arm first, like the driver)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --impl reference --gpus $N --steps 5 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 \
    bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 600 gpurun_out/bench_n$N.err; wc -c gpurun_out/bench_n$N.json
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"]["value"], d["e2e"].get("host_GBps_aggregate"), d.get("parity"))
    p = d.get("pipeline") or {}
    print(json.dumps({k: (v if not isinstance(v, dict) else {a: b for a, b in v.items() if a != "config"}) for k, v in p.items()})[:3000])
except Exception as e:
    print("no bench line:", e)
PY
