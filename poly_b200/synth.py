"""Counter-based synthetic workloads.

The reference's random.DNASequence (random/random.go:43-63) draws from Go's
math/rand stream, which cannot be reproduced without Go; the benchmark inputs are
therefore defined by splitmix64 over a global base counter so that host (numpy)
and device (pg_synth_* kernels) generate byte-identical data.
"""
from __future__ import annotations

import numpy as np

SEED_READS = 0x706F6C79
SEED_PRIMER = 0x7072696D
SEED_TEMPLATE = 0x74656D70
_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def codes(seed: int, g0: int, count: int) -> np.ndarray:
    """2-bit codes for global base indices [g0, g0+count): one splitmix64 word per 32 bases."""
    if count <= 0:
        return np.zeros(0, dtype=np.uint8)
    w0, w1 = g0 >> 5, (g0 + count - 1) >> 5
    out = np.empty((w1 - w0 + 1) * 32, dtype=np.uint8)
    shifts = (np.arange(32, dtype=np.uint64) * np.uint64(2))[None, :]
    step = 1 << 19  # words per chunk (bounded temporaries)
    for a in range(w0, w1 + 1, step):
        b = min(a + step, w1 + 1)
        with np.errstate(over="ignore"):
            w = splitmix64(np.uint64(seed) + np.arange(a, b, dtype=np.uint64))
        out[(a - w0) * 32:(b - w0) * 32] = ((w[:, None] >> shifts) & np.uint64(3)).astype(np.uint8).reshape(-1)
    lo = g0 & 31
    return out[lo:lo + count]


_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def independent_reads(n_reads: int, length: int, first_read: int = 0, seed: int = SEED_READS) -> np.ndarray:
    """cfg1/cfg2/cfg4 shape: reads back to back, read i position j = code(seed, i*L+j)."""
    return _ACGT[codes(seed, first_read * length, n_reads * length)]


def family_reads(n_reads: int, length: int, family: int = 100, first_read: int = 0, seed: int = SEED_READS) -> np.ndarray:
    """cfg3 shape: R reads per family share a template; ~1/64 positions substituted."""
    out = np.empty(n_reads * length, dtype=np.uint8)
    pos = np.arange(length, dtype=np.uint64)
    for r in range(n_reads):
        i = first_read + r
        t = i // family
        T = codes(seed, t * length, length).astype(np.uint64)
        with np.errstate(over="ignore"):
            m = splitmix64(np.uint64(seed ^ 0xD157) + np.uint64(i * length) + pos)
        sub = (m & np.uint64(63)) == 0
        alt = (T + np.uint64(1) + ((m >> np.uint64(6)) % np.uint64(3))) & np.uint64(3)
        out[r * length:(r + 1) * length] = _ACGT[np.where(sub, alt, T).astype(np.uint8)]
    return out


def primers(n: int, length: int = 25, first: int = 0) -> np.ndarray:
    return _ACGT[codes(SEED_PRIMER, first * length, n * length)]


def template(length: int = 10000) -> np.ndarray:
    return _ACGT[codes(SEED_TEMPLATE, 0, length)]


def uniform_offsets(n: int, length: int) -> np.ndarray:
    return np.arange(n + 1, dtype=np.uint64) * np.uint64(length)


def fnv1a64(data: np.ndarray) -> int:
    """FNV-1a-64 over the little-endian bytes of `data` (used for golden checksums)."""
    h = 0xCBF29CE484222325
    for b in np.ascontiguousarray(data).view(np.uint8).tobytes():
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h
