"""FASTQ ingest on the GPU: text -> (bases, offsets), the layout the sketch entry points take.

Mirrors the record structure and the failure behaviour of io/fastq Parser.ParseNext / ParseN
(/root/reference/io/fastq/fastq.go:88-99,117-214): strict 4-line records, the valid prefix is
returned together with the first error.  `ingest` materialises only the sequences (what the sketch
path needs); `Parse` returns the reference's records (Identifier, Optionals, Sequence, Quality) -- the
sequences from the dense GPU output, the other three cut out of the caller's text with the line spans
the kernel reports.  SURVEY.md 8f.2.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _lib

ERRORS = {0: None, 1: "line {line} failed: unexepcted EOF encountered", 2: "empty fastq sequence, got to line {line}",
          3: "empty quality sequence, got to line {line}", 4: "did not find fastq start '@', got to line {line}",
          5: "reference panics (index out of range) while parsing the identifier at line {line}",
          6: "line {line} too large for buffer, use larger maxLineSize"}


class FastqError(Exception):
    def __init__(self, code: int, line: int):
        super().__init__(ERRORS[code].format(line=line))
        self.code, self.line = code, line


def ingest(text: bytes) -> Tuple[np.ndarray, np.ndarray, Optional[FastqError]]:
    """(bases uint8, offsets uint64[n+1], error or None): the sequences of every record up to the
    first one the reference parser rejects."""
    buf = np.frombuffer(text, dtype=np.uint8)
    cap_rec = text.count(b"\n") // 4 + 1
    bases = np.zeros(max(len(text), 1), dtype=np.uint8)
    offsets = np.zeros(cap_rec + 1, dtype=np.uint64)
    n, tot, ec, el = C.c_uint64(0), C.c_uint64(0), C.c_int32(0), C.c_uint64(0)
    _lib.check(_lib.lib().pg_fastq_ingest(buf.ctypes.data if len(text) else None, len(text), bases.ctypes.data, len(bases), offsets.ctypes.data, cap_rec,
                                          C.byref(n), C.byref(tot), C.byref(ec), C.byref(el)))
    err = FastqError(ec.value, el.value) if ec.value else None
    return bases[: tot.value], offsets[: n.value + 1], err


@dataclass
class Fastq:
    """fastq.Fastq (io/fastq/fastq.go:46-51)."""
    Identifier: str = ""
    Optionals: Dict[str, str] = field(default_factory=dict)
    Sequence: str = ""
    Quality: str = ""


def Parse(text: bytes) -> Tuple[List[Fastq], Optional[FastqError]]:
    """fastq.Parse (fastq.go:54-59): every record up to the first one the reference rejects, and that error."""
    buf = np.frombuffer(text, dtype=np.uint8)
    cap_rec = text.count(b"\n") // 4 + 1
    bases = np.zeros(max(len(text), 1), dtype=np.uint8)
    offsets = np.zeros(cap_rec + 1, dtype=np.uint64)
    spans = np.zeros((cap_rec + 1, 4), dtype=np.uint64)
    n, tot, ec, el = C.c_uint64(0), C.c_uint64(0), C.c_int32(0), C.c_uint64(0)
    _lib.check(_lib.lib().pg_fastq_ingest_records(buf.ctypes.data if len(text) else None, len(text), bases.ctypes.data, len(bases), offsets.ctypes.data,
                                                  spans.ctypes.data, cap_rec, C.byref(n), C.byref(tot), C.byref(ec), C.byref(el)))
    err = FastqError(ec.value, el.value) if ec.value else None
    out: List[Fastq] = []
    seqs = bases.tobytes()
    for i in range(n.value):
        ib, il, qb, ql = (int(x) for x in spans[i])
        tokens = text[ib: ib + il].decode("latin-1").split(" ")            # fastq.go:157
        opts = {}
        for datum in tokens[1:]:                                            # fastq.go:160-165 (every datum holds '=': checked on the GPU)
            kv = datum.split("=")
            opts[kv[0]] = kv[1]
        out.append(Fastq(tokens[0][1:], opts, seqs[int(offsets[i]): int(offsets[i + 1])].decode("latin-1"), text[qb: qb + ql].decode("latin-1")))
    return out, err
