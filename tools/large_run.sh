set -x
timeout 600 python -m pytest tests/test_gpu_mash.py -x -q -k "beyond or threshold" --timeout 400 --timeout-method thread 2>&1 | tail -15
