#!/usr/bin/env python
"""cfg3 all-pairs once (dense, then sparse): the target of `ncu -k regex:bucket_join` captures."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poly_b200 import _lib, synth  # noqa: E402

L = _lib.lib(); _lib.check(L.pg_init(0))
dev = torch.device("cuda", 0); st = torch.cuda.current_stream().cuda_stream
n, RL, k, s = int(os.environ.get("N", 100_000)), 10_000, 31, 2000
reads = torch.empty(n * RL, dtype=torch.uint8, device=dev)
_lib.check(L.pg_synth_reads_dev(reads.data_ptr(), 0, n, RL, synth.SEED_READS, 1, 100, st))
sk = torch.empty((n, s), dtype=torch.int32, device=dev)
_lib.check(L.pg_mash_sketch_uniform_dev(reads.data_ptr(), n, RL, k, s, 0, sk.data_ptr(), s, None, st))
same = torch.empty((n, n), dtype=torch.int32, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(2):
    e0.record(); _lib.check(L.pg_mash_distance_block_dev(sk.data_ptr(), n, s, 0, n, same.data_ptr(), None, st)); e1.record(); torch.cuda.synchronize()
    print("dense all-pairs ms", e0.elapsed_time(e1))
