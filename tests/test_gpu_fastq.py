"""GPU FASTQ ingest (pg_fastq_ingest) vs the oracle: valid prefix, error kind and line; then the
ingest -> sketch pipeline equals sketching the same reads directly."""
import numpy as np
import pytest

from fastq_util import make_fastq, py_parse_records
from poly_b200 import fastq, mash, synth
from test_oracle_fastq import mutations

pytestmark = pytest.mark.gpu


def check(oracle, text):
    bases, offsets, err = fastq.ingest(text)
    seqs, ec, el = oracle.fastq_parse(text)
    assert len(offsets) - 1 == len(seqs)
    assert bytes(bases) == b"".join(seqs)
    assert offsets.tolist() == np.concatenate([[0], np.cumsum([len(s) for s in seqs])]).astype(np.uint64).tolist()
    assert (err.code, err.line) == (ec, el) if ec else err is None


def test_ingest_mutated_vs_oracle(gpu, oracle):
    rng = np.random.default_rng(6)
    for ragged in (False, True):
        base = make_fastq(40, 80, ragged=ragged, rng=rng)
        for text in mutations(base, rng):
            check(oracle, text)


def test_ingest_large_then_sketch(gpu, oracle):
    n, L, k, s = 200_000, 150, 21, 1000
    text = make_fastq(n, L)
    bases, offsets, err = fastq.ingest(text)
    assert err is None and len(offsets) == n + 1 and np.array_equal(bases, synth.independent_reads(n, L))
    out, count, status = mash.sketch_arrays(bases, offsets, k, s)
    direct = mash.sketch_uniform(synth.independent_reads(n, L), n, L, k, s)
    assert np.array_equal(out[:, : L - k], direct)
    # a record broken in the middle: the valid prefix and the error position
    cut = text[: len(text) // 2]
    check(oracle, cut)


def test_parse_records_vs_python_restatement(gpu):
    """fastq.Parse through pg_fastq_ingest_records: Identifier / Optionals / Sequence / Quality of every
    record of the valid prefix, and the error, equal the pure-Python restatement of fastq.go:117-214."""
    rng = np.random.default_rng(8)
    texts = [b"@e3cc70d5-90ef runid=5c88f4 read=13956 ch=53 start_time=2020-11-11T01:49:01Z\nGATGTGCG\n+\n$$&%&%#$\n"
             b"@second\nAC\n+anything\n!!\n@third a=b a=c x=\nT\n+\nI\n", b"", b"@only\nA\n+\n"]
    for ragged in (False, True):
        base = make_fastq(30, 70, ragged=ragged, rng=rng)
        texts += list(mutations(base, rng))
    for text in texts:
        got, err = fastq.Parse(text)
        want, ec, el = py_parse_records(text)
        assert [(r.Identifier, r.Optionals, r.Sequence, r.Quality) for r in got] == want
        assert ((err.code, err.line) if err else (0, 0)) == (ec, el)
    got, err = fastq.Parse(texts[0])
    assert got[0].Identifier == "e3cc70d5-90ef" and got[0].Optionals["ch"] == "53" and got[0].Quality == "$$&%&%#$"
    assert got[2].Optionals == {"a": "c", "x": ""} and err is None
