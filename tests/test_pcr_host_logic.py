"""CPU tier: the fragment assembly of poly_b200.pcr.SimulateSimple / Simulate (the restatement of
primers/pcr/pcr.go:117-166,181-195) with its two GPU building blocks replaced by CPU stand-ins
(oracle Tm loop, bytes.find): the host logic alone must reproduce the reference's test expectations."""
import numpy as np
import pytest

import pcr_util as P
from poly_b200 import _lib, pcr


@pytest.fixture()
def cpu_blocks(monkeypatch):
    def minimal_primer_lengths(primers, target):
        ml, st = np.zeros(len(primers), np.uint32), np.zeros(len(primers), np.int32)
        for i, p in enumerate(primers):
            p = bytes(p)
            if len(p) < 7:
                st[i] = _lib.PG_ITEM_PANIC
            else:
                ml[i] = P.minimal_length(p.upper(), target)
        return ml, st

    def find_sites(sequences, patterns, upper=True):
        hits = sorted((si, pos, qi) for si, s in enumerate(sequences) for qi, pat in enumerate(patterns)
                      for pos in P.occurrences(bytes(s).upper() if upper else bytes(s), bytes(pat)))
        a = np.array(hits, dtype=np.int64).reshape(-1, 3)
        return a[:, 0].astype(np.uint32), a[:, 1].astype(np.uint64), a[:, 2].astype(np.uint32)

    monkeypatch.setattr(pcr, "minimal_primer_lengths", minimal_primer_lengths)
    monkeypatch.setattr(pcr, "find_sites", find_sites)


def test_reference_expectations_through_the_host_logic(cpu_blocks):
    for seqs, circular, primers, ok in P.reference_cases():
        fr, err = pcr.Simulate(seqs, 55.0, circular, list(primers))
        assert ok(fr, str(err) if err else None), primers
        assert (fr, str(err) if err else None) == P.simulate(seqs, 55.0, circular, primers)


def test_primer_list_is_upper_cased_in_place_and_short_primers_panic(cpu_blocks):
    primers = [P.FWD.decode().lower(), P.REV.decode()]
    assert pcr.SimulateSimple([P.GENE.decode()], 55.0, False, primers) == [P.FULL_AMPLICON]
    assert primers[0] == P.FWD.decode()                       # pcr.go:76-78
    with pytest.raises(_lib.GoPanic):
        pcr.SimulateSimple([P.GENE], 55.0, False, [b"ACGT"])  # primer[len(primer)-7:]
    assert pcr.SimulateSimple([], 55.0, False, [P.FWD]) == []
