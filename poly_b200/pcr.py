"""Host-side mirror of poly's primers/pcr (primer design and PCR simulation) over libpolyb200.so.

Mirrors /root/reference/primers/pcr/pcr.go:44-66 (`DesignPrimersWithOverhangs`, `DesignPrimers`):
the Tm search (grow the primer one base at a time until MeltingTemp reaches the target) runs on the
GPU for every sequence of a batch and both directions at once (`pg_design_primers_batch`); the
primer strings are assembled here from the returned lengths (pcr.go:55-59).  SURVEY.md 8f.3.

`SimulateSimple` / `Simulate` (pcr.go:73-195): the minimal-primer Tm loops of all primers
(`pg_pcr_minimal_primer_batch`) and the binding-site search the reference does with a suffix array
(`pg_find_sites_batch`: every occurrence of every minimal primer and of its reverse complement in
every upper-cased sequence) run on the GPU; the fragment assembly of pcr.go:117-166,181-195 is
restated here statement by statement (it is bookkeeping over a handful of sites).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import GoPanic
from .mash import BytesLike, _as_bytes, flatten

def _complement_table() -> bytes:
    """complementTable of transform/transform.go:78-109 as a 256-byte translation table: unlisted bytes map to 0."""
    table = bytearray(256)
    for a, b in zip(b"ABCDGHKMNRSTVWYabcdghkmnrstvwy", b"TVGHCDMKNYSABWRtvghcdmknysabwr"):
        table[a] = b
    return bytes(table)


_COMPLEMENT = _complement_table()


def reverse_complement(seq: bytes) -> bytes:
    """transform.ReverseComplement (transform/transform.go:15-23)."""
    return bytes(seq[::-1].translate(_COMPLEMENT))


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


minimalPrimerLength = 7  # pcr.go:35 (designedMinimalPrimerLength = 15, pcr.go:38, is only for DesignPrimers)


def _upper_ascii(b: bytes) -> bytes:
    if any(c >= 0x80 for c in b):
        raise ValueError("byte >= 0x80: strings.ToUpper on non-ASCII input is unsupported")
    return b.upper()


def minimal_primer_lengths(primers: Sequence[BytesLike], target_tm: float):
    """(min_len, status) per primer: the loop of pcr.go:93-100."""
    bases, offsets = flatten(primers)
    n = len(primers)
    ml, st = np.zeros(n, np.uint32), np.zeros(n, np.int32)
    rc = _lib.lib().pg_pcr_minimal_primer_batch(bases.ctypes.data, offsets.ctypes.data, n, float(target_tm), ml.ctypes.data, st.ctypes.data)
    _lib.check(rc, allow=(_lib.PG_ERR_PANIC, _lib.PG_ERR_UNSUPPORTED))
    return ml, st


def find_sites(sequences: Sequence[BytesLike], patterns: Sequence[BytesLike], upper: bool = True):
    """All exact occurrences as rows (sequence index, position, pattern index), sorted."""
    sb, so = flatten(sequences)
    pb, po = flatten(patterns)
    cap = 1024
    while True:
        hs, hp, hq = np.zeros(cap, np.uint32), np.zeros(cap, np.uint64), np.zeros(cap, np.uint32)
        n = C.c_uint64(0)
        rc = _lib.lib().pg_find_sites_batch(sb.ctypes.data, so.ctypes.data, len(sequences), pb.ctypes.data, po.ctypes.data, len(patterns),
                                            1 if upper else 0, hs.ctypes.data, hp.ctypes.data, hq.ctypes.data, cap, C.byref(n))
        if rc == _lib.PG_ERR_ARG and n.value > cap:
            cap = int(n.value)
            continue
        _lib.check(rc)
        k = int(n.value)
        order = np.lexsort((hq[:k], hp[:k], hs[:k]))
        return hs[:k][order], hp[:k][order], hq[:k][order]


def _generate_pcr_fragments(sequence: bytes, fwd_loc: int, rev_loc: int, fwd_idx: List[int], rev_idx: List[int],
                            minimal: List[Optional[bytes]], primers: List[bytes]) -> List[bytes]:
    """generatePcrFragments, pcr.go:181-195."""
    out = []
    for fi in fwd_idx:
        mp = minimal[fi] or b""
        full_fwd = primers[fi]
        for ri in rev_idx:
            out.append(full_fwd[: len(full_fwd) - len(mp)] + sequence[fwd_loc:rev_loc] + reverse_complement(primers[ri]))
    return out


def SimulateSimple(sequences: Sequence[BytesLike], targetTm: float, circular: bool, primerList: List) -> List[str]:
    """pcr.SimulateSimple (pcr.go:73-169).  Like the reference it upper-cases `primerList` in place."""
    as_str = [isinstance(p, str) for p in primerList]
    prim = [_upper_ascii(bytes(_as_bytes(p))) for p in primerList]
    for i, p in enumerate(prim):                                   # pcr.go:76-78
        primerList[i] = p.decode("latin-1") if as_str[i] else p
    seqs = [_upper_ascii(bytes(_as_bytes(s))) for s in sequences]   # pcr.go:82
    fragments: List[bytes] = []
    if not seqs:
        return []
    ml, st = minimal_primer_lengths(prim, targetTm) if prim else (np.zeros(0, np.uint32), np.zeros(0, np.int32))
    if (st == _lib.PG_ITEM_PANIC).any():
        raise GoPanic("slice bounds out of range (primer shorter than 7 nt)")   # primer[len(primer)-index:]
    minimal: List[Optional[bytes]] = [None] * len(prim)              # minimalPrimers, "" in Go where unset
    patterns, owner = [], []
    for i, p in enumerate(prim):
        mp = p[len(p) - int(ml[i]):]                                  # pcr.go:102 (minimalLength 0 -> "")
        if mp != p:                                                   # pcr.go:103
            minimal[i] = mp
            patterns += [mp, reverse_complement(mp)]
            owner += [(i, False), (i, True)]
    hs, hp, hq = find_sites(seqs, patterns, upper=False) if patterns else (np.zeros(0, np.uint32),) * 3
    for si, sequence in enumerate(seqs):
        fwd: Dict[int, List[int]] = {}
        rev: Dict[int, List[int]] = {}
        sel = hs == si
        # primers are visited in list order (pcr.go:92); each appends itself to the sites it binds
        for pos, q in sorted(zip(hp[sel].tolist(), hq[sel].tolist()), key=lambda t: (t[1], t[0])):
            i, is_rev = owner[q]
            (rev if is_rev else fwd).setdefault(pos, []).append(i)
        fwd_locs, rev_locs = sorted(fwd), sorted(rev)                 # pcr.go:127-128
        for index, f in enumerate(fwd_locs):
            if index + 1 != len(fwd_locs):                            # pcr.go:133-143
                for r in rev_locs:
                    if f < r < fwd_locs[index + 1]:
                        fragments += _generate_pcr_fragments(sequence, f, r, fwd[f], rev[r], minimal, prim)
                        break
            else:
                found = False
                for r in rev_locs:                                    # pcr.go:146-151
                    if f < r:
                        fragments += _generate_pcr_fragments(sequence, f, r, fwd[f], rev[r], minimal, prim)
                        found = True
                if circular and not found:                            # pcr.go:153-164
                    for r in rev_locs:
                        if fwd_locs[0] > r:
                            rotated = sequence[f:] + sequence[:f]
                            fragments += _generate_pcr_fragments(rotated, 0, len(sequence[f:]) + r, fwd[f], rev[r], minimal, prim)
    return [f.decode("latin-1") for f in fragments]


class PcrError(Exception):
    pass


def Simulate(sequences: Sequence[BytesLike], targetTm: float, circular: bool, primerList: List) -> Tuple[Optional[List[str]], Optional[PcrError]]:
    """pcr.Simulate (pcr.go:171-186): (fragments, error)."""
    for p in primerList:
        if len(p) < minimalPrimerLength:
            return None, PcrError("Primers are too short.")
    initial = SimulateSimple(sequences, targetTm, circular, primerList)
    subsequent = SimulateSimple(sequences, targetTm, circular, list(primerList) + list(initial))
    if len(initial) != len(subsequent):
        return initial, PcrError("Concatemerization detected in PCR.")
    return initial, None
