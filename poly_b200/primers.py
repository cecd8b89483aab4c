"""Host-side mirror of poly's primers.SantaLucia / MeltingTemp over libpolyb200.so.

Mirrors /root/reference/primers/primers.go:70-105 (`SantaLucia`) and :121-128
(`MeltingTemp`); `MeltingTemps` / `santalucia_arrays` are batched additions.  An empty
sequence panics in the reference (primers.go:89) -> `GoPanic`.  Bytes >= 0x80 are
rejected (strings.ToUpper would re-encode them; unsupported domain).
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import GoPanic
from .mash import BytesLike, flatten

DEFAULT_CP, DEFAULT_NA, DEFAULT_MG = 500e-9, 50e-3, 0.0  # primers.go:122-124


def santalucia_arrays(bases: np.ndarray, offsets: np.ndarray, cp: float, na: float, mg: float):
    """(tm, dH, dS, status) per primer."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    tm, dh, ds, st = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n, np.int32)
    rc = _lib.lib().pg_tm_batch(bases.ctypes.data, offsets.ctypes.data, n, cp, na, mg, tm.ctypes.data, dh.ctypes.data,
                                ds.ctypes.data, st.ctypes.data)
    _lib.check(rc, allow=(_lib.PG_ERR_PANIC, _lib.PG_ERR_UNSUPPORTED))
    return tm, dh, ds, st


def _raise_for(st: np.ndarray) -> None:
    if (st == _lib.PG_ITEM_PANIC).any():
        raise GoPanic("index out of range [-1]")  # primers.go:89
    if (st == _lib.PG_ITEM_UNSUPPORTED).any():
        raise ValueError("byte >= 0x80 in primer: unsupported (strings.ToUpper would re-encode it)")


def SantaLucia(sequence: BytesLike, primerConcentration: float, saltConcentration: float,
               magnesiumConcentration: float) -> Tuple[float, float, float]:
    """primers.SantaLucia (primers.go:70-105) -> (meltingTemp, dH, dS)."""
    bases, offsets = flatten([sequence])
    tm, dh, ds, st = santalucia_arrays(bases, offsets, primerConcentration, saltConcentration, magnesiumConcentration)
    _raise_for(st)
    return float(tm[0]), float(dh[0]), float(ds[0])


def MeltingTemp(sequence: BytesLike) -> float:
    """primers.MeltingTemp (primers.go:121-128)."""
    return SantaLucia(sequence, DEFAULT_CP, DEFAULT_NA, DEFAULT_MG)[0]


def MeltingTemps(sequences: Sequence[BytesLike]) -> np.ndarray:
    bases, offsets = flatten(sequences)
    tm, _, _, st = santalucia_arrays(bases, offsets, DEFAULT_CP, DEFAULT_NA, DEFAULT_MG)
    _raise_for(st)
    return tm
