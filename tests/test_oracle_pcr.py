"""Pins the CPU restatement of pcr.Simulate (tests/pcr_util.py, Tm from oracle/) on the expectations of
the reference's own tests: primers/pcr/example_test.go:10-36,58-69 and pcr_test.go:13-101."""
import re

import pytest

import pcr_util as P

REF_EXAMPLE = "/root/reference/primers/pcr/example_test.go"


def test_reference_expectations():
    for seqs, circular, primers, ok in P.reference_cases():
        fr, err = P.simulate(seqs, 55.0, circular, primers)
        assert ok(fr, err), primers


@pytest.mark.skipif(not __import__("os").path.exists(REF_EXAMPLE), reason="reference only exists in the build container")
def test_constants_equal_the_reference_test_files():
    ex = open(REF_EXAMPLE).read()
    want = re.search(r"// Output: \[(\w+)\]", ex).group(1)
    assert P.FULL_AMPLICON == want
    assert P.GENE.decode() == re.search(r'gene := "(\w+)"', ex).group(1)
    assert P.BAD_FRAGMENT.decode() == re.search(r'badFragment := "(\w+)"', ex).group(1)
    t = open("/root/reference/primers/pcr/pcr_test.go").read()
    assert P.CIRCULAR_TARGET == re.search(r'targetFragment := "(\w+)"', t).group(1)
    assert P.FULL_AMPLICON == re.search(r'want := "(\w+)"', t).group(1)


def test_minimal_length_quirks():
    # the loop keeps the last length that FAILED the test: Tm(minimal primer) < target <= Tm(one base more)
    import oracle_ffi as o
    ml = P.minimal_length(P.FWD, 55.0)
    assert 7 <= ml < len(P.FWD)
    assert o.melting_temp(P.FWD[len(P.FWD) - ml:]) < 55.0 <= o.melting_temp(P.FWD[len(P.FWD) - ml - 1:])
    assert P.minimal_length(b"CTGCAGGTCGACTCTAG", 55.0) == 17           # whole primer below target: ignored (pcr.go:103)
    assert P.minimal_length(b"G" * 15 + b"C" * 15, 15.0) == 0            # the 7-mer (Tm 19.6) already reaches the target
    assert P.minimal_length(b"G" * 15 + b"C" * 15, 20.0) == 7            # ... and just misses this one
    # a GC-rich primer whose 15-nt suffix is already above the target still has a minimal part of 7..14 nt
    gc = b"GCGGCCGCGGGCCCGCGGCCGC"
    assert o.melting_temp(gc[-15:]) >= 55.0 and 7 <= P.minimal_length(gc, 55.0) < 15
    assert P.minimal_length(b"ACGTACG", 55.0) == 7                       # 7 nt is legal (pcr.go:35), whole primer below target
    with pytest.raises(IndexError):
        P.minimal_length(b"ACGT", 55.0)
