"""pcr.SimulateSimple / Simulate through the GPU building blocks (pg_pcr_minimal_primer_batch,
pg_find_sites_batch) vs the reference's test expectations and the CPU restatement."""
import numpy as np
import pytest

import pcr_util as P
from poly_b200 import _lib, pcr, synth
from poly_b200._lib import GoPanic

pytestmark = pytest.mark.gpu


def test_reference_expectations(gpu):
    for seqs, circular, primers, ok in P.reference_cases():
        fr, err = pcr.Simulate(seqs, 55.0, circular, list(primers))
        assert ok(fr, str(err) if err else None), primers
        assert (fr, str(err) if err else None) == P.simulate(seqs, 55.0, circular, primers)
    primers = ["ttatAGGTCTCATACTAATAATTACACCGAGATAACACATCATGG", P.REV.decode()]
    assert pcr.SimulateSimple([P.GENE.decode()], 55.0, False, primers) == [P.FULL_AMPLICON]
    assert primers[0] == P.FWD.decode()                     # upper-cased in place, pcr.go:76-78
    with pytest.raises(GoPanic):
        pcr.SimulateSimple([P.GENE], 55.0, False, [b"ACGT"])


def test_minimal_primer_lengths_vs_oracle(gpu, oracle):
    rng = np.random.default_rng(21)
    primers = [bytes(rng.choice(list(b"ACGTacgtN"), int(rng.integers(7, 60))).tolist()) for _ in range(400)]
    primers += [b"G" * 15 + b"C" * 15, b"A" * 40, b"CTGCAGGTCGACTCTAG", b"ACGT", b"", b"ACGTACGTACGTAC\xc3\xa9A", b"GCGGCCGCGGGCCCGCGGCCGC", b"ACGTACG", b"GCGCGCG", b"ACGTAC"]
    for target in (20.0, 55.0, 72.5):
        ml, st = pcr.minimal_primer_lengths(primers, target)
        for i, p in enumerate(primers):
            if any(c >= 0x80 for c in p):
                assert st[i] == _lib.PG_ITEM_UNSUPPORTED
            elif len(p) < 7:
                assert st[i] == _lib.PG_ITEM_PANIC
            else:
                assert st[i] == 0 and ml[i] == P.minimal_length(p.upper(), target), (p, target)


def test_find_sites_vs_python(gpu):
    rng = np.random.default_rng(22)
    seqs = [bytes(rng.choice(list(b"ACGTacgt"), int(n)).tolist()) for n in (0, 1, 7, 300, 5000, 64, 20000)]
    seqs += [b"A" * 50, b"ACACACACACAC"]
    pats = [b"", b"A", b"AAAA", b"ACAC", b"ACGT", bytes(seqs[3][10:30]).upper(), bytes(seqs[4][-25:]).upper(), b"G" * 70, bytes(seqs[6][:15]).upper()]
    pats.append(bytes(seqs[3][-3:] + seqs[4][:3]).upper())      # straddles a sequence boundary in the flat layout: must not match there
    hs, hp, hq = pcr.find_sites(seqs, pats, upper=True)
    got = sorted(zip(hs.tolist(), hp.tolist(), hq.tolist()))
    want = sorted((si, pos, qi) for si, s in enumerate(seqs) for qi, p in enumerate(pats) for pos in P.occurrences(s.upper(), p))
    assert got == want and len(want) > 1024                      # also exercises the capacity retry
    hs, hp, hq = pcr.find_sites(seqs, pats, upper=False)
    want = sorted((si, pos, qi) for si, s in enumerate(seqs) for qi, p in enumerate(pats) for pos in P.occurrences(s, p))
    assert sorted(zip(hs.tolist(), hp.tolist(), hq.tolist())) == want
    big = [bytes(40000 * b"x") + b"P" * 20 for _ in range(2)]      # patterns beyond the shared-memory staging
    hs, hp, hq = pcr.find_sites([b"zzPPPPPPPPPPPPPPPPPPPPPzz"], big + [b"PPP"], upper=False)
    assert len(hs) == 19 and set(hq.tolist()) == {2}


def test_random_reactions_vs_restatement(gpu):
    rng = np.random.default_rng(23)
    n_frag = 0
    for trial in range(12):
        L = int(rng.integers(600, 3000))
        tpl = bytes(synth.independent_reads(1, L, first_read=100 + trial))
        sites = sorted(int(x) for x in rng.integers(0, L - 40, 4))
        primers = [tpl[sites[0]: sites[0] + 30], P.revcomp(tpl[sites[2]: sites[2] + 32]), b"ACGTTGCAACGTTGCAT" + tpl[sites[1]: sites[1] + 28],
                   P.revcomp(tpl[sites[3]: sites[3] + 35]).lower(), tpl[sites[0]: sites[0] + 30]]
        if trial % 3 == 0:
            primers.append(b"CTGCAGGTCGACTCTAG")
        seqs = [tpl, tpl[L // 3:] + tpl[: L // 3], tpl.lower()[: L // 2]]
        for circular in (False, True):
            got = pcr.SimulateSimple(seqs, 55.0, circular, list(primers))
            assert got == P.simulate_simple(seqs, 55.0, circular, primers)
            n_frag += len(got)
    assert n_frag > 50
