"""FASTA ingest on the GPU: text -> (sequences, offsets, names, name offsets).

Mirrors fasta.Parse / NewParser(r, maxLineSize).ParseAll() of
/root/reference/io/fasta/fasta.go:72-77,96-118,149-243, quirks included (skipped short and ';'
lines, a '>' line right after a name line is sequence text, an unterminated last record is
dropped silently, the valid prefix is returned together with the first error).  SURVEY.md 8f.2.

`bufio_alias=True` (the default of `Parse`, to give what the reference gives) additionally
reproduces the reference's use of the bufio line slice after Peek(1) (fasta.go:192): a line whose
newline is the last byte of a full 64 KiB reader buffer is seen with bytes from one buffer further
on.  `bufio_alias=False` takes every line as written.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from . import _lib

MAX_LINE_SIZE = 2 * 32 * 1024  # fasta.go:74
BUFIO_ALIAS = 1

ERRORS = {0: None, 1: "did not find fasta start '>', got to line {line}", 2: "empty fasta sequence, got to line {line}",
          3: "line {line} too large for buffer, use larger maxLineSize", 4: "bufio: buffer full"}


class FastaError(Exception):
    def __init__(self, code: int, line: int):
        super().__init__(ERRORS[code].format(line=line))
        self.code, self.line = code, line


@dataclass
class Fasta:  # fasta.go:66-69
    Name: str
    Sequence: str


def ingest(text: bytes, max_line_size: int = MAX_LINE_SIZE, bufio_alias: bool = True, names: bool = True):
    """(sequences uint8, offsets uint64[n+1], names uint8, name_offsets uint64[n+1], error or None)
    for every record up to the first error of the reference parser."""
    buf = np.frombuffer(text, dtype=np.uint8)
    cap_rec = text.count(b"\n") + 1
    bases = np.zeros(max(len(text), 1), dtype=np.uint8)
    offsets = np.zeros(cap_rec + 1, dtype=np.uint64)
    nm = np.zeros(max(len(text), 1), dtype=np.uint8) if names else None
    noff = np.zeros(cap_rec + 1, dtype=np.uint64) if names else None
    n, tot, ntot, ec, el = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), C.c_int32(0), C.c_uint64(0)
    _lib.check(_lib.lib().pg_fasta_ingest(buf.ctypes.data if len(text) else None, len(text), max_line_size,
                                          BUFIO_ALIAS if bufio_alias else 0, bases.ctypes.data, len(bases), offsets.ctypes.data,
                                          nm.ctypes.data if names else None, len(nm) if names else 0,
                                          noff.ctypes.data if names else None, cap_rec,
                                          C.byref(n), C.byref(tot), C.byref(ntot), C.byref(ec), C.byref(el)))
    err = FastaError(ec.value, el.value) if ec.value else None
    if not names:
        return bases[: tot.value], offsets[: n.value + 1], None, None, err
    return bases[: tot.value], offsets[: n.value + 1], nm[: ntot.value], noff[: n.value + 1], err


def Parse(text: bytes, max_line_size: int = MAX_LINE_SIZE, bufio_alias: bool = True) -> Tuple[List[Fasta], Optional[FastaError]]:
    """fasta.Parse (fasta.go:72): ([]Fasta, error)."""
    seq, off, nm, noff, err = ingest(text, max_line_size, bufio_alias)
    sb, nb = seq.tobytes(), nm.tobytes()
    recs = [Fasta(nb[int(noff[i]): int(noff[i + 1])].decode("latin-1"), sb[int(off[i]): int(off[i + 1])].decode("latin-1"))
            for i in range(len(off) - 1)]
    return recs, err
