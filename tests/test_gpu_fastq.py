"""GPU FASTQ ingest (pg_fastq_ingest) vs the oracle: valid prefix, error kind and line; then the
ingest -> sketch pipeline equals sketching the same reads directly."""
import numpy as np
import pytest

from fastq_util import make_fastq
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
