"""Synthetic FASTQ text for the ingest tests (same record shape as the reference's fixtures)."""
import numpy as np

from poly_b200 import synth


def make_fastq(n: int, length=150, seed_read=0, ragged=False, rng=None) -> bytes:
    rng = rng or np.random.default_rng(0)
    reads = synth.independent_reads(n, length, first_read=seed_read)
    parts = []
    for i in range(n):
        L = int(rng.integers(1, length + 1)) if ragged else length
        seq = bytes(reads[i * length: i * length + L])
        qual = bytes(rng.integers(35, 75, L, dtype=np.uint8))
        parts.append(b"@read%d runid=abc%d ch=%d\n" % (i, i, i % 512) + seq + b"\n+\n" + qual + b"\n")
    return b"".join(parts)


def py_parse(text: bytes):
    """Independent pure-Python statement of fastq.ParseAll (io/fastq/fastq.go:88-99,117-214):
    (sequences, err_code, err_line)."""
    pos, line, seqs = 0, 0, []
    n = len(text)
    while pos < n:
        rec, no_at = [], False
        for l in range(4):
            q = text.find(b"\n", pos)
            line += 1
            if q < 0:
                return seqs, (6 if n - pos >= 65536 else 1), line   # bufio fills its 64 KiB buffer before it can see EOF
            if q + 1 - pos > 65536:
                return seqs, 6, line
            ln = text[pos:q]
            pos = q + 1
            if l == 0:
                if not ln:
                    return seqs, 5, line
                no_at = ln[:1] != b"@"
                for datum in ln.split(b" ")[1:]:
                    if b"=" not in datum:
                        return seqs, 5, line
            elif l == 1 and not ln:
                return seqs, 2, line
            elif l == 3 and not ln:
                return seqs, 3, line
            rec.append(ln)
        if no_at:
            return seqs, 4, line
        seqs.append(rec[1])
    return seqs, 0, 0


def py_parse_records(text: bytes):
    """fastq.ParseAll with the record fields (io/fastq/fastq.go:46-51,117-214), pure Python:
    ([(identifier, optionals, sequence, quality)], err_code, err_line)."""
    seqs, ec, el = py_parse(text)
    recs, pos = [], 0
    for _ in range(len(seqs)):
        lines = []
        for _l in range(4):
            q = text.index(b"\n", pos)
            lines.append(text[pos:q].decode("latin-1"))
            pos = q + 1
        splits = lines[0].split(" ")                       # fastq.go:157
        opts = {}
        for datum in splits[1:]:
            kv = datum.split("=")
            opts[kv[0]] = kv[1]                            # fastq.go:161-164 (a later duplicate key wins, as in a Go map)
        recs.append((splits[0][1:], opts, lines[1], lines[3]))
    return recs, ec, el
