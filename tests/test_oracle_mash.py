"""Pins the CPU oracle (oracle/mash_oracle.c) for search/mash BEFORE it is trusted as the
checker: murmur3 known-answer vectors, a replay of every assertion of the reference's own
tests (search/mash/mash_test.go:9-62, example_test.go:9-22), the SURVEY goldens
(tests/golden/survey_goldens.json), and faithful == closed-form equivalence."""
import json
import os

import numpy as np
import pytest

import oracle_ffi as o
from poly_b200 import synth

A = "ATGCGATCGATCGATCGATCGATCGATCGATCGATCGATCGATCGATCGATCGATCGATCGA"
B = "ATCGATCGATCGATCGATCGATCGATCGATCGATCGAATGCGATCGATCGATCGATCGATCG"
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "survey_goldens.json")))


def test_murmur3_known_answers():
    """github.com/spaolacci/murmur3 v1.1.0 == MurmurHash3_x86_32; upstream/SMHasher vectors."""
    for text, h in GOLD["murmur3_kat"].items():
        assert o.murmur3_32(text) == int(h, 16), text
    # seeds, and every tail length 0..3 against an independent pure-Python restatement
    def py_mm3(data: bytes, seed=0):
        c1, c2, M = 0xCC9E2D51, 0x1B873593, 0xFFFFFFFF
        rot = lambda x, r: ((x << r) | (x >> (32 - r))) & M
        h = seed
        nb = len(data) // 4
        for i in range(nb):
            k = int.from_bytes(data[4 * i:4 * i + 4], "little")
            k = (k * c1) & M; k = rot(k, 15); k = (k * c2) & M
            h ^= k; h = rot(h, 13); h = (h * 5 + 0xE6546B64) & M
        t = data[4 * nb:]
        if t:
            k = int.from_bytes(t, "little")
            k = (k * c1) & M; k = rot(k, 15); k = (k * c2) & M
            h ^= k
        h ^= len(data)
        h ^= h >> 16; h = (h * 0x85EBCA6B) & M; h ^= h >> 13; h = (h * 0xC2B2AE35) & M; h ^= h >> 16
        return h
    rng = np.random.default_rng(0)
    for n in list(range(0, 40)) + [63, 64, 65, 255]:
        d = bytes(rng.integers(0, 256, n, dtype=np.uint8))
        assert o.murmur3_32(d) == py_mm3(d)
        assert o.murmur3_32(d, 0x9747B28C) == py_mm3(d, 0x9747B28C)
    assert o.murmur3_32("", 1) == 0x514E28B7 and o.murmur3_32("", 0xFFFFFFFF) == 0x81F16F39  # SMHasher verification values


def test_reference_TestMash_replay():
    """search/mash/mash_test.go:9-62."""
    for faithful in (True, False):
        f1 = o.OracleMash(17, 10); f1.Sketch(A, faithful)
        f2 = o.OracleMash(17, 9); f2.Sketch(A, faithful)
        assert f1.Distance(f2) == 0 and f2.Distance(f1) == 0           # :16-24
        sp = o.OracleMash(17, 10); sp.Sketches[0] = 0
        assert f1.Distance(sp) == 1                                     # :26-32
        assert f1.Distance(o.OracleMash(17, 9)) == 1                    # :34-39
        f1 = o.OracleMash(17, 10); f1.Sketch(A, faithful)
        f2 = o.OracleMash(17, 5); f2.Sketch(B, faithful)
        d = f1.Distance(f2)
        assert 0.19 < d < 0.21 and d == 0.19999999999999996             # :41-50
        f1 = o.OracleMash(17, 10); f1.Sketch(B, faithful)
        f2 = o.OracleMash(17, 5); f2.Sketch(A, faithful)
        assert f1.Distance(f2) == 0                                     # :52-61


def test_survey_goldens_mash():
    g = GOLD["mash"]
    hashes = [o.murmur3_32(A[i:i + 17]) for i in range(len(A) - 17)]
    assert len(hashes) == 45  # L-k windows, the last k-mer is never hashed (mash.go:73)
    assert [hex(h) for h in hashes[:8]] == [hex(int(x, 16)) for x in g["A_hashes_first8"]]
    for key, (seq, s) in {"A_k17_s10": (A, 10), "B_k17_s5": (B, 5), "B_k17_s10": (B, 10)}.items():
        m = o.OracleMash(17, s); m.Sketch(seq)
        assert [int(x) for x in m.Sketches] == [int(x, 16) for x in g[key]]


def test_survey_goldens_cfg1_and_cfg3():
    c = GOLD["cfg1"]
    reads = synth.independent_reads(1000, 150)
    assert bytes(reads[:50]).decode() == c["read0_prefix"]
    rc, out = o.sketch_batch(reads, synth.uniform_offsets(1000, 150), 21, 1000, variant=0, nthreads=4)
    assert rc == 0
    assert [hex(int(x)) for x in out[0, :4]] == c["first4"] and hex(int(out[0, 128])) == c["w128"]
    assert not out[:, 129:].any()
    assert synth.fnv1a64(out) == int(c["fnv_padded"], 16)
    assert synth.fnv1a64(np.ascontiguousarray(out[:, :129])) == int(c["fnv_compact"], 16)
    assert int(out.astype(np.uint64).sum()) == int(c["sum"], 16)
    c = GOLD["cfg3_sample"]
    reads = synth.family_reads(8, 10000, family=4)
    rc, out = o.sketch_batch(reads, synth.uniform_offsets(8, 10000), 31, 2000, variant=1, nthreads=4)
    assert [hex(int(x)) for x in out[0, :3]] == c["first3"] and hex(int(out[0, 1999])) == c["last"]
    assert len(set(out[0].tolist())) == 2000 and (np.diff(out[0].astype(np.int64)) > 0).all()
    assert synth.fnv1a64(out) == int(c["fnv"], 16)
    row0 = []
    for j in range(8):
        a = o.OracleMash(31, 2000); a.Sketches[:] = out[0]
        b = o.OracleMash(31, 2000); b.Sketches[:] = out[j]
        row0.append(a.SimilarityCount(b)[0])
    assert row0 == c["row0"]


def test_faithful_equals_closed_form_random():
    """Closed form of mash.go:68-104 vs the literal loop, incl. duplicate-heavy
    alphabets, n == s-1 / s / s+1 boundaries and sketching onto a non-fresh Mash."""
    rng = np.random.default_rng(7)
    for trial in range(300):
        k = int(rng.integers(0, 12)); s = int(rng.integers(0, 40)); L = int(rng.integers(0, 120))
        alpha = [b"A", b"AC", b"ACGT", bytes(range(256))][trial % 4]
        seq = bytes(rng.choice(list(alpha), size=L).astype(np.uint8))
        f = o.OracleMash(k, s); c = o.OracleMash(k, s)
        pre = rng.integers(0, 2 ** 32, s, dtype=np.uint64).astype(np.uint32)
        if trial % 3 == 0:
            f.Sketches[:] = pre; c.Sketches[:] = pre
        rf, rcl = f.Sketch(seq, True), c.Sketch(seq, False)
        assert rf == rcl, (k, s, L)
        if rf == 0:
            assert np.array_equal(f.Sketches, c.Sketches), (k, s, L)
    for s in (5, 6, 7):  # L-k = 6 around s
        seq = bytes(rng.choice(list(b"ACGT"), size=10).astype(np.uint8))
        f = o.OracleMash(4, s); c = o.OracleMash(4, s)
        f.Sketch(seq, True); c.Sketch(seq, False)
        assert np.array_equal(f.Sketches, c.Sketches)
        sorted_ = bool((np.diff(f.Sketches.astype(np.int64)) >= 0).all())
        assert sorted_ or s == 7  # sorted iff L-k >= s


def test_panic_paths():
    seq = bytes(synth.independent_reads(1, 60))
    assert o.OracleMash(21, 0).Sketch(seq) == o.PO_PANIC          # mash.go:96 Sketches[-1]
    assert o.OracleMash(21, 0).Sketch(seq[:21]) == 0              # no k-mer, no panic
    hs = [o.murmur3_32(seq[i:i + 21]) for i in range(39)]
    m = o.OracleMash(21, 1)
    assert m.Sketch(seq) == (0 if min(hs) == hs[0] else o.PO_PANIC)
    with pytest.raises(IndexError):
        o.OracleMash(17, 0).Distance(o.OracleMash(17, 3))


def test_similarity_semantics():
    """mash.go:107-135: receiver is 'larger' on ties; early-out; literal walk on unsorted."""
    def py_similarity(a, b):  # independent pure-Python statement of mash.go:107-135
        L, S = (a, b) if len(a) >= len(b) else (b, a)
        if L[-1] < S[0] or S[-1] < L[0]:
            return 0
        same = si = li = 0
        while si < len(S) and li < len(L):
            if S[si] == L[li]:
                same += 1; si += 1; li += 1
            elif S[si] < L[li]:
                si += 1
            else:
                li += 1
        return same
    a = o.OracleMash(1, 4); a.Sketches[:] = [5, 1, 5, 9]
    b = o.OracleMash(1, 4); b.Sketches[:] = [1, 5, 9, 9]
    assert a.SimilarityCount(b) == (2, 0.5)
    rng = np.random.default_rng(11)
    for _ in range(500):
        sa, sb = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        x = o.OracleMash(1, sa); x.Sketches[:] = rng.integers(0, 6, sa)
        y = o.OracleMash(1, sb); y.Sketches[:] = rng.integers(0, 6, sb)
        want = py_similarity(x.Sketches.tolist(), y.Sketches.tolist())
        assert x.SimilarityCount(y) == (want, want / min(sa, sb))
    z = o.OracleMash(21, 1000)
    r = o.OracleMash(21, 1000); r.Sketch(bytes(synth.independent_reads(1, 150)))
    assert r.Distance(r) == 1.0  # identical 150-bp reads: early-out fires (SURVEY "READ THIS FIRST")
    assert z.Distance(z) == 0.0  # all zeros: walk matches everything
