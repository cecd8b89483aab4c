"""ctypes binding for oracle/liboracle.so -- the CPU restatement of the reference.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by poly_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = None

PO_OK, PO_PANIC, PO_UNSUPPORTED = 0, -1, -2


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so when missing or stale.  Several ranks of a torchrun job may get here
    at once: the build runs under an exclusive file lock and replaces the library atomically."""
    import fcntl

    so = os.path.join(ORACLE_DIR, "liboracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("mash_oracle.c", "align_primers_oracle.c", "fasta_oracle.c", "batch_drivers.c", "poly_oracle.h")]

    def stale():
        return force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)

    if stale():
        with open(os.path.join(ORACLE_DIR, ".build.lock"), "w") as lk:
            fcntl.flock(lk, fcntl.LOCK_EX)
            if stale():  # nobody built it while we waited
                tmp = f"liboracle.{os.getpid()}.tmp.so"
                subprocess.check_call(["make", "-C", ORACLE_DIR, "-s", "-B", "CC=gcc", f"OUT={tmp}"])
                os.replace(os.path.join(ORACLE_DIR, tmp), so)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u8p, u32p, u64p, i64p = (C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_int64))
        L.po_murmur3_32.restype = C.c_uint32
        L.po_murmur3_32.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32]
        for f in (L.po_mash_sketch_faithful, L.po_mash_sketch_closed):
            f.restype = C.c_int
            f.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]
        L.po_mash_similarity.restype = C.c_int
        L.po_mash_similarity.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, i64p, C.POINTER(C.c_double)]
        L.po_mash_distance.restype = C.c_int
        L.po_mash_distance.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.po_mash_sketch_batch.restype = C.c_int
        L.po_mash_sketch_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, u64p]
        L.po_sw_score.restype = C.c_int
        L.po_sw_score.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_int64, i64p, i64p, i64p, C.POINTER(C.c_int32), i64p]
        L.po_sw_align.restype = C.c_int
        L.po_sw_align.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_int64, i64p, C.c_char_p, C.c_char_p, C.c_int64, i64p, C.POINTER(C.c_int32), i64p]
        L.po_fastq_parse.restype = C.c_int
        L.po_fastq_parse.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, u64p, C.POINTER(C.c_int32), u64p]
        L.po_fasta_parse.restype = C.c_int
        L.po_fasta_parse.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p,
                                     C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, u64p, C.POINTER(C.c_int32), u64p]
        L.po_nw_align.restype = C.c_int
        L.po_nw_align.argtypes = L.po_sw_align.argtypes
        L.po_nw_score.restype = C.c_int
        L.po_nw_score.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                  C.c_int64, i64p, C.POINTER(C.c_int32), i64p]
        L.po_reverse_complement.restype = None
        L.po_reverse_complement.argtypes = [C.c_char_p, C.c_int64, C.c_char_p]
        L.po_santalucia.restype = C.c_int
        L.po_santalucia.argtypes = [C.c_char_p, C.c_int64, C.c_double, C.c_double, C.c_double] + [C.POINTER(C.c_double)] * 3
        L.po_melting_temp.restype = C.c_int
        L.po_melting_temp.argtypes = [C.c_char_p, C.c_int64, C.POINTER(C.c_double)]
        L.po_sw_score_batch.restype = C.c_int
        L.po_sw_score_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                        C.c_int64, C.c_int, C.c_void_p, u64p]
        L.po_melting_temp_batch.restype = C.c_int
        L.po_melting_temp_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, u64p]
        L.po_mash_similarity_block.restype = C.c_int
        L.po_mash_similarity_block.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, u64p]
        _LIB = L
    return _LIB


def _b(s) -> bytes:
    return s.encode("latin-1") if isinstance(s, str) else bytes(s)


def murmur3_32(data, seed: int = 0) -> int:
    d = _b(data)
    return lib().po_murmur3_32(d, len(d), seed)


class OracleMash:
    """Mirror of mash.Mash (search/mash/mash.go:52-56) over the oracle."""

    def __init__(self, kmer_size: int, sketch_size: int):
        self.KmerSize, self.SketchSize = kmer_size, sketch_size
        self.Sketches = np.zeros(sketch_size, dtype=np.uint32)

    def Sketch(self, seq, faithful: bool = True) -> int:
        d = _b(seq)
        f = lib().po_mash_sketch_faithful if faithful else lib().po_mash_sketch_closed
        return f(d, len(d), self.KmerSize, self.SketchSize, self.Sketches.ctypes.data)

    def SimilarityCount(self, other: "OracleMash"):
        same, sim = C.c_int64(0), C.c_double(0)
        rc = lib().po_mash_similarity(self.Sketches.ctypes.data, self.SketchSize, other.Sketches.ctypes.data,
                                      other.SketchSize, C.byref(same), C.byref(sim))
        if rc != PO_OK:
            raise IndexError("reference would panic: index out of range")
        return same.value, sim.value

    def Similarity(self, other) -> float:
        return self.SimilarityCount(other)[1]

    def Distance(self, other) -> float:
        d = C.c_double(0)
        rc = lib().po_mash_distance(self.Sketches.ctypes.data, self.SketchSize, other.Sketches.ctypes.data, other.SketchSize, C.byref(d))
        if rc != PO_OK:
            raise IndexError("reference would panic: index out of range")
        return d.value


def sketch_batch(bases: np.ndarray, offsets: np.ndarray, k: int, s: int, variant: int = 1, nthreads: int = 1,
                 padded: bool = True):
    """Returns (rc, out[n,s] uint32) -- fresh zeroed sketch per read."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    out = np.zeros((n, max(s, 0)), dtype=np.uint32)
    rc = lib().po_mash_sketch_batch(bases.ctypes.data, offsets.ctypes.data, n, k, s, variant, nthreads, out.ctypes.data, None)
    return rc, out


def sketch_batch_timing(bases, offsets, k, s, variant, nthreads):
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    ck = C.c_uint64(0)
    rc = lib().po_mash_sketch_batch(bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1, k, s, variant, nthreads, None, C.byref(ck))
    return rc, ck.value


def sw_score(a, b, lut_a, lut_b, table, gap):
    """Returns (score, max_row, max_col, err_code, err_pos)."""
    a, b = _b(a), _b(b)
    lut_a = np.ascontiguousarray(lut_a, dtype=np.int16)
    lut_b = np.ascontiguousarray(lut_b, dtype=np.int16)
    table = np.ascontiguousarray(table, dtype=np.int64)
    sc, mr, mc, ep = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
    ec = C.c_int32(0)
    rc = lib().po_sw_score(a, len(a), b, len(b), lut_a.ctypes.data, lut_b.ctypes.data, table.ctypes.data, table.shape[1], gap,
                           C.byref(sc), C.byref(mr), C.byref(mc), C.byref(ec), C.byref(ep))
    assert rc == PO_OK
    return sc.value, mr.value, mc.value, ec.value, ep.value


def nw_align(a, b, lut_a, lut_b, table, gap):
    """Returns (score, alignA, alignB, err_code, err_pos) of the full align.NeedlemanWunsch."""
    return sw_align(a, b, lut_a, lut_b, table, gap, _fn="po_nw_align")


def sw_align(a, b, lut_a, lut_b, table, gap, _fn="po_sw_align"):
    """Returns (score, alignA, alignB, err_code, err_pos) of the full align.SmithWaterman."""
    a, b = _b(a), _b(b)
    lut_a = np.ascontiguousarray(lut_a, dtype=np.int16)
    lut_b = np.ascontiguousarray(lut_b, dtype=np.int16)
    table = np.ascontiguousarray(table, dtype=np.int64)
    cap = len(a) + len(b) + 1
    oa, ob = C.create_string_buffer(cap), C.create_string_buffer(cap)
    sc, n, ep, ec = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int32(0)
    rc = getattr(lib(), _fn)(a, len(a), b, len(b), lut_a.ctypes.data, lut_b.ctypes.data, table.ctypes.data, table.shape[1], gap,
                           C.byref(sc), oa, ob, cap, C.byref(n), C.byref(ec), C.byref(ep))
    assert rc == PO_OK, rc
    return sc.value, oa.raw[: n.value], ob.raw[: n.value], ec.value, ep.value


def fastq_parse(text: bytes):
    """Returns (sequences list, err_code, err_line) as fastq.Parse would (valid prefix + error)."""
    cap = text.count(b"\n") // 4 + 2
    st, ln = np.zeros(cap, np.uint64), np.zeros(cap, np.uint64)
    n, ec, el = C.c_uint64(0), C.c_int32(0), C.c_uint64(0)
    rc = lib().po_fastq_parse(text, len(text), st.ctypes.data, ln.ctypes.data, cap, C.byref(n), C.byref(ec), C.byref(el))
    assert rc == PO_OK
    return [text[int(st[i]): int(st[i] + ln[i])] for i in range(n.value)], ec.value, el.value


def fasta_parse(text: bytes, max_line_size: int = 65536, alias: bool = True):
    """Returns ([(name, sequence)], err_code, err_line) as fasta.Parse / NewParser(r, max_line_size).ParseAll would."""
    cap = text.count(b"\n") + 2
    seq, name = np.zeros(len(text) + 1, np.uint8), np.zeros(len(text) + 1, np.uint8)
    so, no = np.zeros(cap + 1, np.uint64), np.zeros(cap + 1, np.uint64)
    n, ec, el = C.c_uint64(0), C.c_int32(0), C.c_uint64(0)
    rc = lib().po_fasta_parse(text, len(text), max_line_size, int(alias), seq.ctypes.data, len(seq), so.ctypes.data,
                              name.ctypes.data, len(name), no.ctypes.data, cap, C.byref(n), C.byref(ec), C.byref(el))
    assert rc == PO_OK
    sb, nb = seq.tobytes(), name.tobytes()
    recs = [(nb[int(no[i]): int(no[i + 1])], sb[int(so[i]): int(so[i + 1])]) for i in range(n.value)]
    return recs, ec.value, el.value


def nw_score(a, b, lut_a, lut_b, table, gap):
    a, b = _b(a), _b(b)
    lut_a = np.ascontiguousarray(lut_a, dtype=np.int16)
    lut_b = np.ascontiguousarray(lut_b, dtype=np.int16)
    table = np.ascontiguousarray(table, dtype=np.int64)
    sc, ep, ec = C.c_int64(0), C.c_int64(0), C.c_int32(0)
    rc = lib().po_nw_score(a, len(a), b, len(b), lut_a.ctypes.data, lut_b.ctypes.data, table.ctypes.data, table.shape[1], gap,
                           C.byref(sc), C.byref(ec), C.byref(ep))
    assert rc == PO_OK
    return sc.value, ec.value, ep.value


def reverse_complement(seq) -> bytes:
    d = _b(seq)
    out = C.create_string_buffer(len(d))
    lib().po_reverse_complement(d, len(d), out)
    return out.raw


def santalucia(seq, cp, na, mg):
    d = _b(seq)
    tm, dh, ds = C.c_double(0), C.c_double(0), C.c_double(0)
    rc = lib().po_santalucia(d, len(d), cp, na, mg, C.byref(tm), C.byref(dh), C.byref(ds))
    return rc, tm.value, dh.value, ds.value


def melting_temp(seq) -> float:
    rc, tm, _, _ = santalucia(seq, 500e-9, 50e-3, 0.0)
    if rc == PO_PANIC:
        raise IndexError("reference would panic: index out of range [-1]")
    if rc != PO_OK:
        raise ValueError("unsupported input")
    return tm


# ---- batch drivers (bench.py's CPU legs) ---------------------------------------------------------
def sw_score_batch(queries: np.ndarray, offsets: np.ndarray, templ: np.ndarray, lut_a, lut_b, table, gap: int, nthreads: int = 1):
    queries = np.ascontiguousarray(queries, dtype=np.uint8); offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    templ = np.ascontiguousarray(templ, dtype=np.uint8)
    lut_a = np.ascontiguousarray(lut_a, dtype=np.int16); lut_b = np.ascontiguousarray(lut_b, dtype=np.int16)
    table = np.ascontiguousarray(table, dtype=np.int64)
    n = len(offsets) - 1
    score = np.zeros(n, dtype=np.int64)
    rc = lib().po_sw_score_batch(queries.ctypes.data, offsets.ctypes.data, n, templ.ctypes.data, len(templ), lut_a.ctypes.data, lut_b.ctypes.data,
                                 table.ctypes.data, table.shape[1], gap, nthreads, score.ctypes.data, None)
    return rc, score


def melting_temp_batch(bases: np.ndarray, offsets: np.ndarray, nthreads: int = 1):
    bases = np.ascontiguousarray(bases, dtype=np.uint8); offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    tm = np.zeros(n, dtype=np.float64)
    rc = lib().po_melting_temp_batch(bases.ctypes.data, offsets.ctypes.data, n, nthreads, tm.ctypes.data, None)
    return rc, tm


def similarity_block(sk: np.ndarray, row_lo: int, row_hi: int, col_lo: int, col_hi: int, nthreads: int = 1):
    sk = np.ascontiguousarray(sk, dtype=np.uint32)
    n, s = sk.shape
    same = np.zeros((row_hi - row_lo, col_hi - col_lo), dtype=np.uint32)
    rc = lib().po_mash_similarity_block(sk.ctypes.data, n, s, row_lo, row_hi, col_lo, col_hi, nthreads, same.ctypes.data, None)
    return rc, same
