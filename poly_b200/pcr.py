"""Host-side mirror of poly's primers/pcr primer design over libpolyb200.so.

Mirrors /root/reference/primers/pcr/pcr.go:44-66 (`DesignPrimersWithOverhangs`, `DesignPrimers`):
the Tm search (grow the primer one base at a time until MeltingTemp reaches the target) runs on the
GPU for every sequence of a batch and both directions at once (`pg_design_primers_batch`); the
primer strings are assembled here from the returned lengths (pcr.go:55-59).  SURVEY.md 8f.3.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import GoPanic
from .mash import BytesLike, _as_bytes, flatten

_COMP = bytes.maketrans(b"ABCDGHKMNRSTVWYabcdghkmnrstvwy", b"TVGHCDMKNYSABWRtvghcdmknysabwr")


def reverse_complement(seq: bytes) -> bytes:
    """transform.ReverseComplement (transform/transform.go:15-23): unlisted bytes map to 0."""
    table = bytearray(256)
    for a, b in zip(b"ABCDGHKMNRSTVWYabcdghkmnrstvwy", b"TVGHCDMKNYSABWRtvghcdmknysabwr"):
        table[a] = b
    return bytes(seq[::-1].translate(bytes(table)))


def design_primer_lengths(sequences: Sequence[BytesLike], target_tm: float):
    """(fwd_len, rev_len, status) per sequence."""
    bases, offsets = flatten(sequences)
    n = len(sequences)
    fwd, rev, st = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.int32)
    rc = _lib.lib().pg_design_primers_batch(bases.ctypes.data, offsets.ctypes.data, n, float(target_tm), fwd.ctypes.data, rev.ctypes.data,
                                            st.ctypes.data)
    _lib.check(rc, allow=(_lib.PG_ERR_PANIC, _lib.PG_ERR_UNSUPPORTED))
    return fwd, rev, st


def DesignPrimersWithOverhangsBatch(sequences: Sequence[BytesLike], forwardOverhang: str, reverseOverhang: str,
                                    targetTm: float) -> List[Tuple[str, str]]:
    fwd, rev, st = design_primer_lengths(sequences, targetTm)
    out = []
    for i, seq in enumerate(sequences):
        if st[i] == _lib.PG_ITEM_PANIC:
            raise GoPanic(f"slice bounds out of range (sequence {i})")  # pcr.go:48 / :52
        if st[i] == _lib.PG_ITEM_UNSUPPORTED:
            raise ValueError("byte >= 0x80 in sequence: unsupported")
        s = bytes(_as_bytes(seq)).upper()  # pcr.go:45
        f = s[: fwd[i]]
        r = reverse_complement(s[len(s) - int(rev[i]):])
        out.append(((forwardOverhang.encode("latin-1") + f).decode("latin-1"),
                    (reverse_complement(reverseOverhang.encode("latin-1")) + r).decode("latin-1")))  # pcr.go:55-59
    return out


def DesignPrimersWithOverhangs(sequence: BytesLike, forwardOverhang: str, reverseOverhang: str, targetTm: float) -> Tuple[str, str]:
    """pcr.DesignPrimersWithOverhangs (pcr.go:44-60)."""
    return DesignPrimersWithOverhangsBatch([sequence], forwardOverhang, reverseOverhang, targetTm)[0]


def DesignPrimers(sequence: BytesLike, targetTm: float) -> Tuple[str, str]:
    """pcr.DesignPrimers (pcr.go:62-66)."""
    return DesignPrimersWithOverhangs(sequence, "", "", targetTm)
