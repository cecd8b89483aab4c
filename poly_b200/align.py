"""Host-side mirror of poly's search/align (Smith-Waterman SCORE), alphabet and
search/align/matrix packages over libpolyb200.so.

Mirrors /root/reference/alphabet/alphabet.go:25-61 (`NewAlphabet`, `Encode`, `Error`),
search/align/matrix/matrix.go:13-38 (`NewSubstitutionMatrix`, `Score`, `Default`),
search/align/matrix/matrices.go:33-40 (`NUC_4`) and search/align/align.go:73-95,171-203
(`Scoring`, `NewScoring`, `SmithWaterman`).  `SmithWaterman` returns the score (raising the
reference's `alphabet.Error`); `SmithWatermanAlign` additionally returns the two aligned strings
of the reference's traceback (align.go:205-231; queries of <= 64 symbols); `NeedlemanWunsch`
returns the global score.  All DP cells and the traceback are computed on the GPU.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .mash import BytesLike, _as_bytes, flatten


class AlphabetError(Exception):
    """alphabet.Error (alphabet.go:14-22)."""


class Alphabet:
    """alphabet.Alphabet (alphabet.go:9-12): symbols are strings; a sequence byte b is
    looked up as Go's string(b), i.e. the UTF-8 encoding of code point b (align.go:90)."""

    def __init__(self, symbols: Sequence[str]):
        self.symbols = list(symbols)
        self.encoding = {}
        for i, sym in enumerate(self.symbols):  # alphabet.go:27-30 (later duplicates win)
            self.encoding[sym] = i

    def Encode(self, symbol: str) -> int:
        if symbol not in self.encoding:
            raise AlphabetError(f"Symbol {symbol} not in alphabet")  # alphabet.go:38
        return self.encoding[symbol]

    def Symbols(self) -> List[str]:
        return self.symbols

    def byte_lut(self) -> np.ndarray:
        lut = np.full(256, -1, dtype=np.int16)
        for b in range(256):
            lut[b] = self.encoding.get(chr(b), -1)
        return lut


def NewAlphabet(symbols: Sequence[str]) -> Alphabet:
    return Alphabet(symbols)


class SubstitutionMatrix:
    """matrix.SubstitutionMatrix (matrix.go:13-17)."""

    def __init__(self, first: Alphabet, second: Alphabet, scores):
        scores = np.asarray(scores, dtype=np.int64)
        if scores.ndim != 2 or len(first.Symbols()) != scores.shape[0] or len(second.Symbols()) != scores.shape[1]:
            raise ValueError("invalid dimensions of substitution matrix")  # matrix.go:21-23
        self.FirstAlphabet, self.SecondAlphabet, self.scores = first, second, scores

    def Score(self, a: str, b: str) -> int:  # matrix.go:28-38
        return int(self.scores[self.FirstAlphabet.Encode(a), self.SecondAlphabet.Encode(b)])


def NewSubstitutionMatrix(first: Alphabet, second: Alphabet, scores) -> SubstitutionMatrix:
    return SubstitutionMatrix(first, second, scores)


_LETTERS = [chr(ord("A") + i) for i in range(26)]
Default = SubstitutionMatrix(Alphabet(_LETTERS), Alphabet(_LETTERS), 2 * np.eye(26, dtype=np.int64) - 1)  # matrix.go:41-74
NUC_4 = [[0, 0, 0, 0, 0], [0, 5, -4, -4, -4], [0, -4, 5, -4, -4], [0, -4, -4, 5, -4], [0, -4, -4, -4, 5]]  # matrices.go:33-40


class Scoring:
    """align.Scoring (align.go:73-76)."""

    def __init__(self, substitution_matrix: Optional[SubstitutionMatrix], gap_penalty: int):
        self.SubstitutionMatrix = substitution_matrix if substitution_matrix is not None else Default  # align.go:80-82
        self.GapPenalty = int(gap_penalty)


def NewScoring(substitution_matrix: Optional[SubstitutionMatrix], gap_penalty: int) -> Scoring:
    return Scoring(substitution_matrix, gap_penalty)


def sw_scores_arrays(q_bases: np.ndarray, q_offsets: np.ndarray, template: BytesLike, scoring: Scoring,
                     query_is_a: bool = True, global_alignment: bool = False) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(score, err_code, err_pos) for every query against one template (Smith-Waterman, or the
    Needleman-Wunsch score with global_alignment=True)."""
    m = scoring.SubstitutionMatrix
    lut_a, lut_b = m.FirstAlphabet.byte_lut(), m.SecondAlphabet.byte_lut()
    table = np.ascontiguousarray(m.scores, dtype=np.int64)
    t = _as_bytes(template)
    q_bases = np.ascontiguousarray(q_bases, dtype=np.uint8)
    q_offsets = np.ascontiguousarray(q_offsets, dtype=np.uint64)
    n = len(q_offsets) - 1
    score, ec, ep = np.zeros(n, np.int64), np.zeros(n, np.int32), np.zeros(n, np.int64)
    fn = _lib.lib().pg_nw_score_batch if global_alignment else _lib.lib().pg_sw_score_batch
    rc = fn(q_bases.ctypes.data, q_offsets.ctypes.data, n, t.ctypes.data, len(t), int(query_is_a),
                                      lut_a.ctypes.data, lut_b.ctypes.data, table.ctypes.data, table.shape[0], table.shape[1],
                                      scoring.GapPenalty, score.ctypes.data, ec.ctypes.data, ep.ctypes.data)
    _lib.check(rc)
    return score, ec, ep


def _error_for(ec: int, ep: int, a: np.ndarray, b: np.ndarray) -> AlphabetError:
    byte = int(a[ep]) if ec == 1 else int(b[ep])
    return AlphabetError(f"Symbol {chr(byte)} not in alphabet")  # alphabet.go:38 via align.go:189-191


def SmithWatermanScores(queries: Sequence[BytesLike], template: BytesLike, scoring: Scoring,
                        query_is_a: bool = True, global_alignment: bool = False) -> Tuple[List[int], List[Optional[AlphabetError]]]:
    """Batched addition: score (and error) of SmithWaterman(query, template) per query
    (or SmithWaterman(template, query) with query_is_a=False)."""
    bases, offsets = flatten(queries)
    score, ec, ep = sw_scores_arrays(bases, offsets, template, scoring, query_is_a, global_alignment)
    t = _as_bytes(template)
    errs: List[Optional[AlphabetError]] = []
    for i in range(len(queries)):
        if ec[i]:
            q = bases[int(offsets[i]): int(offsets[i + 1])]
            errs.append(_error_for(int(ec[i]), int(ep[i]), q if query_is_a else t, t if query_is_a else q))
        else:
            errs.append(None)
    return [int(x) for x in score], errs


def SmithWatermanScore(stringA: BytesLike, stringB: BytesLike, scoring: Scoring) -> int:
    """Score of align.SmithWaterman(stringA, stringB, scoring) (align.go:171-203), score kernel only.
    Raises AlphabetError where the reference returns (0, "", "", err)."""
    scores, errs = SmithWatermanScores([stringA], stringB, scoring, query_is_a=True)
    if errs[0] is not None:
        raise errs[0]
    return scores[0]


def NeedlemanWunschScores(queries: Sequence[BytesLike], template: BytesLike, scoring: Scoring, query_is_a: bool = True):
    """Batched score of align.NeedlemanWunsch(query, template) (align.go:100-134,166)."""
    return SmithWatermanScores(queries, template, scoring, query_is_a, global_alignment=True)


def NeedlemanWunschScore(stringA: BytesLike, stringB: BytesLike, scoring: Scoring) -> int:
    """Score of align.NeedlemanWunsch(stringA, stringB, scoring); the aligned strings are not built."""
    scores, errs = NeedlemanWunschScores([stringA], stringB, scoring)
    if errs[0] is not None:
        raise errs[0]
    return scores[0]


def SmithWatermanAligns(queries: Sequence[BytesLike], template: BytesLike, scoring: Scoring, query_is_a: bool = True,
                        global_alignment: bool = False):
    """Batched full align.SmithWaterman (or NeedlemanWunsch with global_alignment=True):
    [(score, alignA, alignB, err)] per query (stringA = query when query_is_a, else the template)."""
    m = scoring.SubstitutionMatrix
    lut_a, lut_b = m.FirstAlphabet.byte_lut(), m.SecondAlphabet.byte_lut()
    table = np.ascontiguousarray(m.scores, dtype=np.int64)
    t = _as_bytes(template)
    bases, offsets = flatten(queries)
    n = len(queries)
    maxq = max((len(_as_bytes(q)) for q in queries), default=0)
    stride = max(2 * maxq + 64, 64) if not global_alignment else maxq + len(t) + 8
    fn = _lib.lib().pg_nw_align_batch if global_alignment else _lib.lib().pg_sw_align_batch
    while True:
        score, ec, ep = np.zeros(n, np.int64), np.zeros(n, np.int32), np.zeros(n, np.int64)
        oa, ob = np.zeros((n, stride), np.uint8), np.zeros((n, stride), np.uint8)
        ln, st = np.zeros(n, np.uint32), np.zeros(n, np.int32)
        rc = fn(bases.ctypes.data, offsets.ctypes.data, n, t.ctypes.data, len(t), int(query_is_a),
                                          lut_a.ctypes.data, lut_b.ctypes.data, table.ctypes.data, table.shape[0], table.shape[1],
                                          scoring.GapPenalty, score.ctypes.data, ec.ctypes.data, ep.ctypes.data, oa.ctypes.data,
                                          ob.ctypes.data, stride, ln.ctypes.data, st.ctypes.data)
        _lib.check(rc)
        if (st == _lib.PG_ITEM_UNSUPPORTED).any():  # an alignment longer than the row: grow and retry
            stride = int(ln.max()) + 8
            continue
        break
    res = []
    for i in range(n):
        if ec[i]:
            q = bases[int(offsets[i]): int(offsets[i + 1])]
            res.append((0, "", "", _error_for(int(ec[i]), int(ep[i]), q if query_is_a else t, t if query_is_a else q)))
        else:
            res.append((int(score[i]), bytes(oa[i, : ln[i]]).decode("latin-1"), bytes(ob[i, : ln[i]]).decode("latin-1"), None))
    return res


def SmithWatermanAlign(stringA: BytesLike, stringB: BytesLike, scoring: Scoring) -> Tuple[int, str, str]:
    """align.SmithWaterman(stringA, stringB, scoring) -> (score, alignA, alignB); raises
    AlphabetError where the reference returns (0, "", "", err).  A string of <= 64 symbols rides in
    registers; when both are longer the whole matrix is kept in HBM, as the reference keeps it on the heap."""
    a, b = _as_bytes(stringA), _as_bytes(stringB)
    if len(a) <= 64:
        score, sa, sb, err = SmithWatermanAligns([a], b, scoring, query_is_a=True)[0]
    else:
        score, sa, sb, err = SmithWatermanAligns([b], a, scoring, query_is_a=False)[0]
    if err is not None:
        raise err
    return score, sa, sb


def NeedlemanWunschAlign(stringA: BytesLike, stringB: BytesLike, scoring: Scoring) -> Tuple[int, str, str]:
    """align.NeedlemanWunsch(stringA, stringB, scoring) -> (score, alignA, alignB) (align.go:100-166),
    including the reference's loop condition (the traceback stops when one string is exhausted)."""
    a, b = _as_bytes(stringA), _as_bytes(stringB)
    if len(a) <= 64:
        score, sa, sb, err = SmithWatermanAligns([a], b, scoring, query_is_a=True, global_alignment=True)[0]
    else:
        score, sa, sb, err = SmithWatermanAligns([b], a, scoring, query_is_a=False, global_alignment=True)[0]
    if err is not None:
        raise err
    return score, sa, sb


def SmithWaterman(stringA: BytesLike, stringB: BytesLike, scoring: Scoring):
    """align.SmithWaterman (align.go:171-232) with the reference's own return shape:
    (score, alignA, alignB, err) -- err is None or the AlphabetError, and (0, "", "", err) on error."""
    try:
        return SmithWatermanAlign(stringA, stringB, scoring) + (None,)
    except AlphabetError as e:
        return 0, "", "", e


def NeedlemanWunsch(stringA: BytesLike, stringB: BytesLike, scoring: Scoring):
    """align.NeedlemanWunsch (align.go:100-166) with the reference's own return shape:
    (score, alignA, alignB, err)."""
    try:
        return NeedlemanWunschAlign(stringA, stringB, scoring) + (None,)
    except AlphabetError as e:
        return 0, "", "", e
