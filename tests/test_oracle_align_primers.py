"""Pins oracle/align_primers_oracle.c on the reference's own known answers:
search/align/align_test.go:139-292, search/align/example_test.go:49-111,
search/align/matrix/matrix_test.go:11-48, primers/primers_test.go:29-84 and the exact
primer strings of primers/pcr/example_test.go:36,46,54 (Tm threshold crossings)."""
import json
import os

import numpy as np
import pytest

import oracle_ffi as o
from poly_b200 import synth

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "survey_goldens.json")))


def lut(symbols):
    l = np.full(256, -1, np.int16)
    for i, s in enumerate(symbols):
        if len(s) == 1 and ord(s) < 256:
            l[ord(s)] = i
    return l


TEST_MAT = np.array([[0, 0, 0, 0, 0], [0, 3, -3, -3, -3], [0, -3, 3, -3, -3], [0, -3, -3, 3, -3], [0, -3, -3, -3, 3]])
TEST_LUT = lut(["-", "A", "C", "G", "T"])
NUC_4 = np.array([[0, 0, 0, 0, 0], [0, 5, -4, -4, -4], [0, -4, 5, -4, -4], [0, -4, -4, 5, -4], [0, -4, -4, -4, 5]])


def sw(a, b, l=TEST_LUT, m=TEST_MAT, gap=-2):
    return o.sw_score(a, b, l, l, m, gap)


def test_reference_TestSmithWaterman_scores():
    assert sw("TGTTACGG", "GGTTGACTA")[0] == 13     # align_test.go:157-176 (Wikipedia example)
    assert sw("ACACACTA", "AGCACACA")[0] == 17      # :178-196
    assert sw("", "GAT")[0] == 0 and sw("", "")[0] == 0   # :200-236
    assert sw("G", "A")[0] == 0 and sw("G", "G")[0] == 3 and sw("G", "GATTACA")[0] == 3  # :238-291


def test_reference_examples():
    l5 = lut(["A", "C", "G", "T", "U"])
    m5 = 2 * np.eye(5, dtype=np.int64) - 1
    assert o.sw_score("GATTACA", "GCATGCU", l5, l5, m5, -1)[0] == 2      # example_test.go:49-83
    # NUC_4 laid out -,A,C,G,T but addressed through alphabet A,C,G,T,- : literal index mapping -> 15
    ln = lut(["A", "C", "G", "T", "-"])
    assert o.sw_score("GATTACA", "GCATGCT", ln, ln, NUC_4, -1)[0] == 15  # example_test.go:85-111
    # Needleman-Wunsch example (align.go:100-166; example_test.go:14-47): score 0
    assert o.nw_score("GATTACA", "GCATGCU", l5, l5, m5, -1)[0] == 0


def test_substitution_matrix_lookup():
    """search/align/matrix/matrix_test.go:11-48 through the LUT flattening."""
    l = lut(["-", "A", "C", "G", "T"])
    for a, b, want in [("A", "A", 5), ("A", "C", -4), ("C", "T", -4), ("-", "-", 0)]:
        assert NUC_4[l[ord(a)], l[ord(b)]] == want
        assert o.sw_score(a, b, l, l, NUC_4, -100)[0] == max(want, 0)


def test_sw_error_order():
    """align.go:188-191: first failing cell in row-major order; Encode(a) before Encode(b)."""
    assert sw("ANA", "GAT")[3:] == (1, 1) or sw("ANA", "GAT")[3:] == (1, 1)
    assert sw("NAA", "GXT")[3:] == (1, 0)   # a[0] bad -> cell (1,1) fails on a
    assert sw("ANA", "GXT")[3:] == (2, 1)   # row 1 reaches b[1] before row 2 starts
    assert sw("AAN", "GAT")[3:] == (1, 2)
    assert sw("", "X")[3] == 0 and sw("X", "")[3] == 0  # no cell is ever scored
    assert sw("ANA", "GXT")[0] == 0


def test_sw_first_max_position_and_swap_invariance():
    rng = np.random.default_rng(0)
    for _ in range(50):
        a = bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(1, 30))).astype(np.uint8))
        b = bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(1, 60))).astype(np.uint8))
        assert sw(a, b)[0] == sw(b, a)[0]


def test_reference_tm_tests():
    """primers/primers_test.go:29-84 (2 % margins) + SURVEY exact values."""
    rc, tm, dh, ds = o.santalucia("ACGATGGCAGTAGCATGC", 0.1e-6, 350e-3, 0.0)
    assert rc == 0 and abs(62.7 - tm) / 62.7 < 0.02
    assert (tm, dh, ds) == pytest.approx(tuple(GOLD["tm"]["ACGATGGCAGTAGCATGC"]), rel=1e-12)
    pal = "ACGTAGATCTACGT"
    assert o.reverse_complement(pal) == pal.encode()
    rc, tm, dh, ds = o.santalucia(pal, 0.1e-6, 350e-3, 0.0)
    assert abs(47.428514 - tm) / 47.428514 < 0.02
    assert (tm, dh, ds) == pytest.approx(tuple(GOLD["tm"]["ACGTAGATCTACGT"]), rel=1e-12)
    tm = o.melting_temp("GTAAAACGACGGCCAGT")
    assert abs(52.8 - tm) / 52.8 < 0.02 and tm == pytest.approx(GOLD["tm"]["GTAAAACGACGGCCAGT_meltingtemp"], rel=1e-12)
    assert o.melting_temp("gtaaaacgacggccagt") == tm  # strings.ToUpper, primers.go:71
    with pytest.raises(IndexError):
        o.melting_temp("")
    assert o.santalucia(b"AC\xc3\xa9", 1e-6, 1e-2, 0)[0] == o.PO_UNSUPPORTED


GENE = ("aataattacaccgagataacacatcatggataaaccgatactcaaagattctatgaagctatttgaggcacttggtacgatcaagtcgcgctcaatgtttggtggcttcggacttttcgctg"
        "atgaaacgatgtttgcactggttgtgaatgatcaacttcacatacgagcagaccagcaaacttcatctaacttcgagaagcaagggctaaaaccgtacgtttataaaaagcgtggttttcc"
        "agtcgttactaagtactacgcgatttccgacgacttgtgggaatccagtgaacgcttgatagaagtagcgaagaagtcgttagaacaagccaatttggaaaaaaagcaacaggcaagtagta"
        "agcccgacaggttgaaagacctgcctaacttacgactagcgactgaacgaatgcttaagaaagctggtataaaatcagttgaacaacttgaagagaaaggtgcattgaatgcttacaaagcg"
        "atacgtgactctcactccgcaaaagtaagtattgagctactctgggctttagaaggagcgataaacggcacgcactggagcgtcgttcctcaatctcgcagagaagagctggaaaatgcgc"
        "tttcttaa")


def design_primers(melting_temp, sequence, target_tm, fwd_oh="", rev_oh=""):
    """primers/pcr/pcr.go:44-60 restated over a MeltingTemp callable."""
    seq = sequence.upper()
    rc = lambda s: o.reverse_complement(s).decode()
    def sl(lo, hi):  # Go slice expression: out-of-range bounds panic
        if lo < 0 or hi > len(seq) or lo > hi:
            raise IndexError("slice bounds out of range")
        return seq[lo:hi]
    fwd = sl(0, 15)
    add = 0
    while melting_temp(fwd) < target_tm:
        fwd = sl(0, 15 + add)
        add += 1
    rev = rc(sl(len(seq) - 15, len(seq)))
    add = 0
    while melting_temp(rev) < target_tm:
        rev = rc(sl(len(seq) - (15 + add), len(seq)))
        add += 1
    return fwd_oh + fwd, rc(rev_oh) + rev


def test_reference_design_primers_strings():
    """primers/pcr/example_test.go:36,46,54: exact primer strings pin the Tm crossings."""
    assert design_primers(o.melting_temp, GENE, 55.0) == ("AATAATTACACCGAGATAACACATCATGG", "TTAAGAAAGCGCATTTTCCAGC")
    assert design_primers(o.melting_temp, GENE, 55.0, "TTATAGGTCTCATACT", "ATGAAGAGACCATATA") == (
        "TTATAGGTCTCATACTAATAATTACACCGAGATAACACATCATGG", "TATATGGTCTCTTCATTTAAGAAAGCGCATTTTCCAGC")


def test_survey_goldens_cfg5():
    pr = synth.primers(6).reshape(6, 25)
    tpl = bytes(synth.template())
    assert [sw(bytes(p), tpl)[0] for p in pr] == GOLD["cfg5"]["sw"]
    assert [o.melting_temp(bytes(p)) for p in pr] == pytest.approx(GOLD["cfg5"]["tm"], rel=1e-12)
