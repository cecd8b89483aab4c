"""GPU FASTA ingest (pg_fasta_ingest) vs the oracle: records, names, error kind and line, with and
without the bufio aliasing of the reference; then FASTA -> sketch equals sketching the sequences."""
import numpy as np
import pytest

from fasta_util import make_fasta
from poly_b200 import _lib, fasta, mash
from test_oracle_fasta import REFERENCE_CASES, edge_cases, random_fasta

pytestmark = pytest.mark.gpu


def gpu_parse(text, bufsz, alias):
    seq, off, nm, noff, err = fasta.ingest(text, bufsz, alias)
    sb, nb = seq.tobytes(), nm.tobytes()
    recs = [(nb[int(noff[i]): int(noff[i + 1])], sb[int(off[i]): int(off[i + 1])]) for i in range(len(off) - 1)]
    return recs, (err.code if err else 0), (err.line if err else 0)


def test_reference_test_cases(gpu, oracle):
    for text, bufsz, want, code in REFERENCE_CASES:
        for alias in (True, False):
            got = gpu_parse(text, bufsz, alias)
            assert got == oracle.fasta_parse(text, bufsz, alias)
            assert got[1] == code and (want is None or got[0] == want)
    recs, err = fasta.Parse(b">doggy or something\nGATTACA\n\nCATGAT\n>homunculus\nAAAA\n")
    assert err is None and [(r.Name, r.Sequence) for r in recs] == [("doggy or something", "GATTACACATGAT"), ("homunculus", "AAAA")]


def test_edge_cases_and_fuzz_vs_oracle(gpu, oracle):
    for text in edge_cases():
        for bufsz in (16, 65536):
            for alias in (True, False):
                assert gpu_parse(text, bufsz, alias) == oracle.fasta_parse(text, bufsz, alias), (text[:40], bufsz, alias)
    rng = np.random.default_rng(12)
    differ = 0
    for _ in range(400):
        text = random_fasta(rng)
        bufsz = int(rng.choice([2, 16, 17, 24, 40]))
        a, c = gpu_parse(text, bufsz, True), gpu_parse(text, bufsz, False)
        assert a == oracle.fasta_parse(text, bufsz, True) and c == oracle.fasta_parse(text, bufsz, False), (text, bufsz)
        differ += a != c
    assert differ > 10


def test_large_file_both_modes(gpu, oracle):
    rng = np.random.default_rng(13)
    text = make_fasta(60_000, 300, rng=rng)                  # ~18 MB, ~280 reader refills
    assert len(text) > 16 << 20
    a, c = gpu_parse(text, 65536, True), gpu_parse(text, 65536, False)
    assert a == oracle.fasta_parse(text, 65536, True)
    assert c == oracle.fasta_parse(text, 65536, False)
    assert c[1] == 0 and len(c[0]) == 60_000
    small = text[: 3 << 20]                                   # smaller reader: many more refills, and a truncated last record
    for bufsz in (256, 1000):
        assert gpu_parse(small, bufsz, True) == oracle.fasta_parse(small, bufsz, True)
        assert gpu_parse(small, bufsz, False) == oracle.fasta_parse(small, bufsz, False)


def test_names_optional_and_capacity(gpu):
    import ctypes as C
    text = make_fasta(100, 200)
    seq, off, nm, noff, err = fasta.ingest(text, names=False)
    seq2, off2, nm2, noff2, _ = fasta.ingest(text)
    assert nm is None and err is None and np.array_equal(seq, seq2) and np.array_equal(off, off2)
    buf = np.frombuffer(text, dtype=np.uint8)
    bases, offsets = np.zeros(10, np.uint8), np.zeros(101, np.uint64)
    n, tot, ntot, ec, el = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), C.c_int32(0), C.c_uint64(0)
    rc = _lib.lib().pg_fasta_ingest(buf.ctypes.data, len(text), 65536, 0, bases.ctypes.data, 10, offsets.ctypes.data, None, 0, None, 100,
                                    C.byref(n), C.byref(tot), C.byref(ntot), C.byref(ec), C.byref(el))
    assert rc == _lib.PG_ERR_ARG and n.value == 100 and tot.value == len(seq)     # the needs are reported


def test_fasta_then_sketch(gpu):
    text = make_fasta(5_000, 400, width=70)
    seq, off, _, _, err = fasta.ingest(text, names=False)
    assert err is None and len(off) == 5_001
    k, s = 21, 200
    out, count, status = mash.sketch_arrays(seq, off, k, s)
    for i in (0, 1, 17, 4_999):
        m = mash.New(k, s)
        m.Sketch(seq[int(off[i]): int(off[i + 1])].tobytes())
        n = int(off[i + 1] - off[i]) - k
        assert int(count[i]) == max(0, min(n, s))
        assert np.array_equal(out[i, : count[i]], np.asarray(m.Sketches[: count[i]], dtype=np.uint32))
