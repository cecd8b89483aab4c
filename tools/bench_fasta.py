#!/usr/bin/env python
"""FASTA ingest timing (device-resident text, CUDA events): 200k records x 50 lines of 60 residues,
with and without the bufio reader model.  Same leg as in tools/bench_secondary.py, on its own."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poly_b200 import _lib, synth  # noqa: E402

L = _lib.lib()
_lib.check(L.pg_init(0))
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
nfa, nl_per, width = 200_000, 50, 60
body = synth.independent_reads(nfa * nl_per, width).reshape(nfa, nl_per, width)
fa = np.empty((nfa, 12 + nl_per * (width + 1)), dtype=np.uint8)
fa[:, :12] = np.frombuffer(b">seq0000000\n", dtype=np.uint8)
blk = fa[:, 12:].reshape(nfa, nl_per, width + 1)
blk[:, :, :width] = body
blk[:, :, width] = 10
ftext = torch.from_numpy(fa.reshape(-1)).to(dev)
f_bases = torch.empty(nfa * nl_per * width + 64, dtype=torch.uint8, device=dev)
f_names = torch.empty(nfa * 16, dtype=torch.uint8, device=dev)
f_off = torch.empty(nfa + 1, dtype=torch.int64, device=dev)
f_noff = torch.empty(nfa + 1, dtype=torch.int64, device=dev)
nr, tot, ntot, ec, el = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), C.c_int32(0), C.c_uint64(0)
for alias in (0, 1):
    def run():
        _lib.check(L.pg_fasta_ingest_dev(ftext.data_ptr(), ftext.numel(), 65536, alias, f_bases.data_ptr(), f_bases.numel(), f_off.data_ptr(),
                                         f_names.data_ptr(), f_names.numel(), f_noff.data_ptr(), nfa, C.byref(nr), C.byref(tot), C.byref(ntot),
                                         C.byref(ec), C.byref(el), st))
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    same = bool(torch.equal(f_bases[: nfa * nl_per * width].cpu(), torch.from_numpy(body.reshape(-1))))
    print(json.dumps({"kernel": "FASTA ingest (8f.2)", "bufio_alias": alias, "records": nr.value, "text_GB": ftext.numel() / 1e9, "ms": ms,
                      "text_GBps": ftext.numel() / ms / 1e6, "err_code": ec.value, "sequences_equal_generator": same}))
