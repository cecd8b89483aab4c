"""Host-side mirror of poly's search/mash package over libpolyb200.so.

Mirrors /root/reference/search/mash/mash.go: `New`, `Mash{KmerSize, SketchSize,
Sketches}`, `(*Mash).Sketch`, `Similarity`, `Distance` keep their names, argument
meaning and failure behaviour (a Go panic surfaces as `GoPanic`, an IndexError).
Batched entry points (`SketchBatch`, `sketch_uniform`, `DistanceMatrix`, ...) are
additions: a single-call API cannot feed a GPU.

Every result is computed on the GPU through the C ABI -- there is no host compute
path here beyond marshalling.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _lib
from ._lib import GoPanic, PolyError

BytesLike = Union[str, bytes, bytearray, np.ndarray]


def _as_bytes(seq: BytesLike) -> np.ndarray:
    if isinstance(seq, str):
        seq = seq.encode("latin-1")  # Go strings are byte strings
    if isinstance(seq, (bytes, bytearray)):
        return np.frombuffer(bytes(seq), dtype=np.uint8)
    return np.ascontiguousarray(seq, dtype=np.uint8)


def flatten(seqs: Sequence[BytesLike]) -> Tuple[np.ndarray, np.ndarray]:
    """[]string -> (bases, offsets) as the C ABI wants them (cgo would do the same)."""
    arrs = [_as_bytes(s) for s in seqs]
    offsets = np.zeros(len(arrs) + 1, dtype=np.uint64)
    if arrs:
        offsets[1:] = np.cumsum([len(a) for a in arrs], dtype=np.uint64)
    bases = np.concatenate(arrs) if arrs else np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(bases, dtype=np.uint8), offsets


class Mash:
    """mash.Mash (mash.go:52-56): exported, user-mutable fields."""

    __slots__ = ("KmerSize", "SketchSize", "Sketches")

    def __init__(self, kmer_size: int, sketch_size: int):
        if sketch_size < 0:
            raise GoPanic("makeslice: len out of range")  # mash.go:63
        self.KmerSize = int(kmer_size)
        self.SketchSize = int(sketch_size)
        self.Sketches = np.zeros(self.SketchSize, dtype=np.uint32)  # mash.go:59-65

    def Sketch(self, sequence: BytesLike) -> None:
        """(*Mash).Sketch (mash.go:68-104); mutates the receiver in place."""
        seq = _as_bytes(sequence)
        k, s = self.KmerSize, self.SketchSize
        n = len(seq) - k
        if n <= 0:
            return  # loop body never runs (mash.go:73)
        if k < 0:
            raise GoPanic("slice bounds out of range")  # sequence[i:i+k]
        cnt = min(n, s)
        out = np.zeros(max(cnt, 1), dtype=np.uint32)
        status = np.zeros(1, dtype=np.int32)
        rc = _lib.lib().pg_mash_sketch_uniform(seq.ctypes.data, 1, len(seq), k, s, 0, out.ctypes.data, cnt, status.ctypes.data)
        if rc == _lib.PG_ERR_PANIC or status[0] == _lib.PG_ITEM_PANIC:
            raise GoPanic("index out of range [-1]")  # mash.go:96-98 with sketchSize in {0,1}
        _lib.check(rc)
        # n < s: only the first n slots are overwritten (mash.go:81-84); n >= s: all s
        self.Sketches[:cnt] = out[:cnt]

    def _pair(self, other: "Mash") -> Tuple[int, float, float]:
        a = np.ascontiguousarray(self.Sketches[: self.SketchSize], dtype=np.uint32)
        b = np.ascontiguousarray(other.Sketches[: other.SketchSize], dtype=np.uint32)
        if len(a) < self.SketchSize or len(b) < other.SketchSize:
            raise GoPanic("index out of range")
        sk = np.concatenate([a, b])
        off = np.array([0, len(a), len(a) + len(b)], dtype=np.uint64)
        pa, pb = np.array([0], dtype=np.uint32), np.array([1], dtype=np.uint32)
        same, sim, dist = np.zeros(1, np.int64), np.zeros(1, np.float64), np.zeros(1, np.float64)
        st = np.zeros(1, np.int32)
        rc = _lib.lib().pg_mash_similarity_pairs(sk.ctypes.data, off.ctypes.data, 2, pa.ctypes.data, pb.ctypes.data, 1,
                                                 same.ctypes.data, sim.ctypes.data, dist.ctypes.data, st.ctypes.data)
        if rc == _lib.PG_ERR_PANIC:
            raise GoPanic("index out of range [-1]")  # Sketches[SketchSize-1] with size 0
        _lib.check(rc)
        return int(same[0]), float(sim[0]), float(dist[0])

    def Similarity(self, other: "Mash") -> float:
        """(*Mash).Similarity (mash.go:107-135)."""
        return self._pair(other)[1]

    def Distance(self, other: "Mash") -> float:
        """(*Mash).Distance (mash.go:138-140)."""
        return self._pair(other)[2]


def New(kmer_size: int, sketch_size: int) -> Mash:
    """mash.New (mash.go:59-65)."""
    return Mash(kmer_size, sketch_size)


# ---- batched additions ------------------------------------------------------------
def sketch_arrays(bases: np.ndarray, offsets: np.ndarray, k: int, s: int, pad_zero: bool = False):
    """Sketch every read of a flattened batch.  Returns (out[n, stride], count[n], status[n]);
    row i holds count[i] informative words (the zero tail of a fresh Mash is materialised only
    with pad_zero)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    lens = np.diff(offsets.astype(np.int64)) if n else np.zeros(0, np.int64)
    maxn = int(max(int(lens.max()) - k, 0)) if n else 0
    stride = s if pad_zero else min(maxn, s)
    out = np.zeros((n, max(stride, 1)), dtype=np.uint32)
    count = np.zeros(n, dtype=np.uint32)
    status = np.zeros(n, dtype=np.int32)
    rc = _lib.lib().pg_mash_sketch_batch(bases.ctypes.data, offsets.ctypes.data, n, k, s,
                                         _lib.PG_SKETCH_PAD_ZERO if pad_zero else 0, out.ctypes.data, max(stride, 1),
                                         count.ctypes.data, status.ctypes.data)
    _lib.check(rc, allow=(_lib.PG_ERR_PANIC,))
    return out, count, status


def sketch_into(bases: np.ndarray, offsets: np.ndarray, k: int, s: int, sketches: np.ndarray, devices=None):
    """The literal in-place form of (*Mash).Sketch for a batch: `sketches` is the caller's [n, s] uint32 array
    (the Sketches of n Mash values, e.g. one zeroed slab as Go's make gives); row i receives its min(L_i - k, s)
    words and the rest of the row is left as it was (mash.go:73-80 never writes Sketches[n:s]) --
    PG_SKETCH_TAIL_KEEP.  Returns (count, status)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    if sketches.dtype != np.uint32 or not sketches.flags.c_contiguous or sketches.shape != (n, s):
        raise ValueError("sketches must be a C-contiguous uint32 array of shape (n, s)")
    count = np.zeros(n, dtype=np.uint32)
    status = np.zeros(n, dtype=np.int32)
    dev, nd = _device_list(devices) if devices is not None else (None, 0)
    if devices is not None:
        rc = _lib.lib().pg_mash_sketch_batch_multi(bases.ctypes.data, offsets.ctypes.data, n, k, s, _lib.PG_SKETCH_TAIL_KEEP,
                                                   sketches.ctypes.data, max(s, 1), count.ctypes.data, status.ctypes.data, _lib.ptr(dev), nd)
    else:
        rc = _lib.lib().pg_mash_sketch_batch(bases.ctypes.data, offsets.ctypes.data, n, k, s, _lib.PG_SKETCH_TAIL_KEEP,
                                             sketches.ctypes.data, max(s, 1), count.ctypes.data, status.ctypes.data)
    _lib.check(rc, allow=(_lib.PG_ERR_PANIC,))
    return count, status


def sketch_uniform(bases: np.ndarray, n_reads: int, read_len: int, k: int, s: int) -> np.ndarray:
    """Fixed-length reads stored back to back -> compact sketches [n, min(max(L-k,0), s)]."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    cnt = min(max(read_len - k, 0), s)
    out = np.zeros((n_reads, max(cnt, 1)), dtype=np.uint32)
    rc = _lib.lib().pg_mash_sketch_uniform(bases.ctypes.data, n_reads, read_len, k, s, 0, out.ctypes.data, max(cnt, 1), None)
    if rc == _lib.PG_ERR_PANIC:
        raise GoPanic("index out of range [-1]")
    _lib.check(rc)
    return out[:, :cnt]


def SketchBatch(sequences: Sequence[BytesLike], kmer_size: int, sketch_size: int) -> List[Mash]:
    """[]string -> []*Mash, each equal to New(k, s) followed by Sketch(seq)."""
    bases, offsets = flatten(sequences)
    out, count, status = sketch_arrays(bases, offsets, kmer_size, sketch_size)
    res = []
    for i in range(len(sequences)):
        if status[i] == _lib.PG_ITEM_PANIC:
            raise GoPanic(f"index out of range [-1] (sequence {i})")
        m = Mash(kmer_size, sketch_size)
        m.Sketches[: count[i]] = out[i, : count[i]]
        res.append(m)
    return res


def similarity_pairs(mashes: Sequence[Mash], pairs: np.ndarray):
    """Explicit pair list (receiver, argument) over sketches of mixed sizes -> (same, similarity, distance)."""
    arrs = [np.ascontiguousarray(m.Sketches[: m.SketchSize], dtype=np.uint32) for m in mashes]
    off = np.zeros(len(arrs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(a) for a in arrs], dtype=np.uint64)
    sk = np.concatenate(arrs) if arrs else np.zeros(0, np.uint32)
    pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
    pa, pb = np.ascontiguousarray(pairs[:, 0]), np.ascontiguousarray(pairs[:, 1])
    n = len(pa)
    same, sim, dist, st = np.zeros(n, np.int64), np.zeros(n, np.float64), np.zeros(n, np.float64), np.zeros(n, np.int32)
    rc = _lib.lib().pg_mash_similarity_pairs(sk.ctypes.data, off.ctypes.data, len(arrs), pa.ctypes.data, pb.ctypes.data, n,
                                             same.ctypes.data, sim.ctypes.data, dist.ctypes.data, st.ctypes.data)
    if rc == _lib.PG_ERR_PANIC:
        raise GoPanic("index out of range [-1]")
    _lib.check(rc)
    return same, sim, dist


def distance_block(sketches: np.ndarray, row_begin: int = 0, row_end: Optional[int] = None, want_distance: bool = True):
    """Rows [row_begin,row_end) x all columns over equal-size sketches [n, s] (full Go arrays).
    Returns (same uint32 [rows, n], distance float64 [rows, n] or None); receiver = row."""
    sketches = np.ascontiguousarray(sketches, dtype=np.uint32)
    n, s = sketches.shape
    row_end = n if row_end is None else row_end
    rows = row_end - row_begin
    same = np.zeros((rows, n), dtype=np.uint32)
    dist = np.zeros((rows, n), dtype=np.float64) if want_distance else None
    rc = _lib.lib().pg_mash_distance_block(sketches.ctypes.data, n, s, row_begin, row_end, same.ctypes.data,
                                           dist.ctypes.data if dist is not None else None)
    if rc == _lib.PG_ERR_PANIC:
        raise GoPanic("index out of range [-1]")
    _lib.check(rc)
    return same, dist


def distance_sparse(sketches: np.ndarray, row_begin: int = 0, row_end: Optional[int] = None, upper: bool = True):
    """The pairs of a row block that share at least one hash, as (i, j, same) arrays sorted by (i, j);
    every pair that is not listed has same == 0 (Distance 1).  upper: only j > i."""
    sketches = np.ascontiguousarray(sketches, dtype=np.uint32)
    n, s = sketches.shape
    row_end = n if row_end is None else row_end
    cap = max(1024, 4 * (row_end - row_begin))
    while True:
        pi, pj, ps = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
        cnt = C.c_uint64(0)
        rc = _lib.lib().pg_mash_distance_sparse(sketches.ctypes.data, n, s, row_begin, row_end, 1 if upper else 0, pi.ctypes.data, pj.ctypes.data,
                                                ps.ctypes.data, cap, C.byref(cnt))
        if rc == _lib.PG_ERR_ARG and cnt.value > cap:
            cap = int(cnt.value)
            continue
        if rc == _lib.PG_ERR_PANIC:
            raise GoPanic("index out of range [-1]")
        _lib.check(rc)
        k = int(cnt.value)
        order = np.lexsort((pj[:k], pi[:k]))
        return pi[:k][order], pj[:k][order], ps[:k][order]


def DistanceMatrix(mashes: Sequence[Mash]) -> np.ndarray:
    """All-pairs Distance for sketches of one common SketchSize: D[i, j] = mashes[i].Distance(mashes[j])."""
    if not mashes:
        return np.zeros((0, 0))
    s = mashes[0].SketchSize
    if any(m.SketchSize != s for m in mashes):
        raise ValueError("DistanceMatrix needs one common SketchSize; use similarity_pairs for mixed sizes")
    sk = np.stack([np.ascontiguousarray(m.Sketches[:s], dtype=np.uint32) for m in mashes])
    return distance_block(sk)[1]


# ---- single-process multi-GPU additions (pg_*_multi) ----------------------------
def _device_list(devices):
    if devices is None:
        return None, 0
    arr = np.ascontiguousarray(devices, dtype=np.int32)
    return arr, len(arr)


def sketch_uniform_multi(bases: np.ndarray, n_reads: int, read_len: int, k: int, s: int, devices=None, out: Optional[np.ndarray] = None) -> np.ndarray:
    """sketch_uniform with the batch sharded over several GPUs of this process (all visible ones
    by default): one call, no data-path collective."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8) if not hasattr(bases, "data_ptr") else bases
    cnt = min(max(read_len - k, 0), s)
    if out is None:
        out = np.zeros((n_reads, max(cnt, 1)), dtype=np.uint32)
    dev, nd = _device_list(devices)
    rc = _lib.lib().pg_mash_sketch_uniform_multi(_lib.ptr(bases), n_reads, read_len, k, s, 0, _lib.ptr(out), max(cnt, 1), None,
                                                 _lib.ptr(dev), nd)
    if rc == _lib.PG_ERR_PANIC:
        raise GoPanic("index out of range [-1]")
    _lib.check(rc)
    return out[:, :cnt] if isinstance(out, np.ndarray) else out


def sketch_arrays_multi(bases: np.ndarray, offsets: np.ndarray, k: int, s: int, pad_zero: bool = False, devices=None):
    """sketch_arrays (ragged reads) sharded over several GPUs of this process."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    lens = np.diff(offsets.astype(np.int64)) if n else np.zeros(0, np.int64)
    maxn = int(max(int(lens.max()) - k, 0)) if n else 0
    stride = s if pad_zero else min(maxn, s)
    out = np.zeros((n, max(stride, 1)), dtype=np.uint32)
    count = np.zeros(n, dtype=np.uint32)
    status = np.zeros(n, dtype=np.int32)
    dev, nd = _device_list(devices)
    rc = _lib.lib().pg_mash_sketch_batch_multi(bases.ctypes.data, offsets.ctypes.data, n, k, s, _lib.PG_SKETCH_PAD_ZERO if pad_zero else 0,
                                               out.ctypes.data, max(stride, 1), count.ctypes.data, status.ctypes.data, _lib.ptr(dev), nd)
    _lib.check(rc, allow=(_lib.PG_ERR_PANIC,))
    return out, count, status


def sketch_distance_multi(bases: np.ndarray, n_reads: int, read_len: int, k: int, s: int, devices=None, want_sketches: bool = True,
                          want_same: bool = True, want_distance: bool = False):
    """Sketch every read and compute the all-pairs matrix on several GPUs of this process: the sketch
    kernels store their results into every device's gathered buffer (fused all-gather over peer
    memory), device r then computes row block r.  Returns (sketches [n, s], same [n, n], distance [n, n])."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    sk = np.zeros((n_reads, s), dtype=np.uint32) if want_sketches else None
    same = np.zeros((n_reads, n_reads), dtype=np.uint32) if want_same else None
    dist = np.zeros((n_reads, n_reads), dtype=np.float64) if want_distance else None
    dev, nd = _device_list(devices)
    rc = _lib.lib().pg_mash_sketch_distance_multi(bases.ctypes.data, n_reads, read_len, k, s, _lib.ptr(dev), nd, _lib.ptr(sk), _lib.ptr(same),
                                                  _lib.ptr(dist))
    if rc == _lib.PG_ERR_PANIC:
        raise GoPanic("index out of range [-1]")
    _lib.check(rc)
    return sk, same, dist
