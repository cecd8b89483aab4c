"""Synthetic FASTA text for the ingest tests and an independent line-oriented Python statement of
fasta.Parse (io/fasta/fasta.go:72-77,96-118,149-243) used to cross-check the byte-level oracle."""
import numpy as np

from poly_b200 import synth


def make_fasta(n: int, mean_len=300, width=60, rng=None, crlf=False) -> bytes:
    rng = rng or np.random.default_rng(0)
    lens = rng.integers(1, 2 * mean_len, n)
    codes = synth.independent_reads(1, int(lens.sum()))
    parts, pos = [], 0
    for i in range(n):
        seq = bytes(codes[pos: pos + int(lens[i])])
        pos += int(lens[i])
        parts.append(b">seq%d some description %d\n" % (i, i * 7))
        for j in range(0, len(seq), width):
            parts.append(seq[j: j + width] + b"\n")
    return b"".join(parts)


def py_parse(text: bytes, max_line_size: int = 65536, alias: bool = True):
    """([(name, sequence)], err_code, err_line).  Works on whole lines: the reader's refill points
    are computed up front (they only depend on the newline positions), then one pass of the
    two-state parser.  Error codes as include/poly_b200.h (pg_fasta_ingest)."""
    B = max(16, max_line_size)
    n = len(text)
    nl = [i for i, c in enumerate(text) if c == 10]
    n_lines = len(nl)
    begins = [0] + [p + 1 for p in nl]          # begins[n_lines] = start of the unterminated tail
    frag = text[begins[n_lines]:]
    # refill chain of a reader that fills every Read: lines whose '\n' is the last byte of a full buffer
    corrupt = set()
    if alias:
        start, idx = 0, 0
        while start + B < n:
            end = start + B
            j = idx
            while j < n_lines and nl[j] < end:
                j += 1
            if j == idx:
                break
            if nl[j - 1] == end - 1:
                corrupt.add(j - 1)
            start, idx = nl[j - 1] + 1, j

    def content(i):
        b, e = begins[i], nl[i]
        if i not in corrupt:
            return text[b:e]
        base = e + 1
        got = min(B, n - base)
        p = b - (base - B)
        return bytes(text[base + p + j] if p + j < got else text[b + j] for j in range(e - b))

    recs, seq, name = [], None, None
    looking = True
    line_no = 0
    for i in range(n_lines):
        raw = text[begins[i]: nl[i]]
        line_no = i + 1
        skippable = len(raw) == 0 or raw[:1] == b";"
        if len(raw) >= B:                                     # bufio.ErrBufferFull
            if not skippable:
                return recs, 3, line_no + 1
            if looking:
                return recs, 1, line_no
            return recs, (2 if not seq else 4), line_no
        eff = content(i)
        peek = text[nl[i] + 1: nl[i] + 2]
        if not looking and peek == b">":
            if not skippable:
                seq += eff
            if not seq:
                return recs, 2, line_no
            recs.append((name, seq))
            looking, seq, name = True, None, None
            continue
        if skippable:
            continue
        if looking:
            if eff[:1] == b">":
                name, seq, looking = eff[1:], b"", False
            continue
        seq += eff
    # the end of the text: the unterminated tail is read as one more line
    if n == 0:
        return recs, 0, 0
    line_no = n_lines + 1
    if len(frag) >= B:
        if frag[:1] != b";":
            return recs, 3, line_no + 1
        if looking:
            return recs, 1, line_no
        return recs, (2 if not seq else 4), line_no
    skippable = len(frag) <= 1 or frag[:1] == b";"
    if looking:
        if n_lines == 0 and not frag:
            return recs, 0, 0
        return (recs, 1, line_no) if skippable else (recs, 0, 0)
    if not skippable:
        return recs, 0, 0                                      # returned with io.EOF: dropped by ParseN
    if not seq:
        return recs, 2, line_no
    recs.append((name, seq))
    return recs, 0, 0
