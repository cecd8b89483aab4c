"""Sketch persistence.

The reference has no storage format: a `mash.Mash` is three exported fields
(/root/reference/search/mash/mash.go:52-56) and serialises through encoding/json as
`{"KmerSize":k,"SketchSize":s,"Sketches":[...]}`.  Two things live here, both host-side
marshalling only (no compute):

* `mash_to_json` / `mash_from_json`: that JSON shape, byte for byte what Go's json.Marshal
  emits for the struct (field order, no spaces), so single sketches interoperate with the reference;
* a compact binary container for sketch SETS (what `pg_mash_sketch_batch` returns and
  `pg_mash_distance_block` consumes), so gathered sketches can be reused across runs:

      offset  size   field
      0       8      magic  "PGSKETCH"
      8       4      version (1), little endian like everything below
      12      4      KmerSize
      16      4      SketchSize
      20      4      flags   bit 0: rows are dense (every row holds SketchSize words, count omitted)
      24      8      n       number of sketches
      32      8      words   total number of uint32 words stored
      40      4*n    count[i] informative words of sketch i          (absent when dense)
      ..      4*words  the informative words of sketch 0, 1, ... back to back
      end-4   4      CRC-32 (zlib) of everything before it

  Only the informative words are stored: a fill-regime sketch of a 150 bp read (k=21, s=1000)
  costs 129 words instead of 1000; `load` can re-materialise the zero tail of a fresh Mash.
"""
from __future__ import annotations

import json
import struct
import zlib
from typing import Tuple

import numpy as np

from .mash import Mash

MAGIC = b"PGSKETCH"
VERSION = 1
FLAG_DENSE = 1
_HDR = struct.Struct("<8sIIIIQQ")


def mash_to_json(m: Mash) -> str:
    """encoding/json of mash.Mash: exported fields in declaration order, compact separators."""
    return json.dumps({"KmerSize": int(m.KmerSize), "SketchSize": int(m.SketchSize), "Sketches": [int(x) for x in m.Sketches]},
                      separators=(",", ":"))


def mash_from_json(text: str) -> Mash:
    d = json.loads(text)
    m = Mash(d["KmerSize"], d["SketchSize"])
    sk = d.get("Sketches")
    m.Sketches = np.zeros(0, dtype=np.uint32) if sk is None else np.asarray(sk, dtype=np.uint32)   # json null <-> nil slice
    return m


def save(path: str, sketches: np.ndarray, count: np.ndarray, kmer_size: int, sketch_size: int) -> int:
    """Write rows `sketches[i, :count[i]]`.  Returns the number of bytes written."""
    sketches = np.ascontiguousarray(sketches, dtype=np.uint32)
    count = np.ascontiguousarray(count, dtype=np.uint32)
    n = len(count)
    if sketches.ndim != 2 or sketches.shape[0] != n or (n and int(count.max()) > sketches.shape[1]):
        raise ValueError("sketches must be [n, stride] with count[i] <= stride")
    dense = bool(n) and bool((count == sketch_size).all()) and sketches.shape[1] == sketch_size
    if dense:
        words = sketches.reshape(-1)
    else:
        mask = np.arange(sketches.shape[1], dtype=np.uint32)[None, :] < count[:, None]
        words = sketches[mask]
    parts = [_HDR.pack(MAGIC, VERSION, kmer_size, sketch_size, FLAG_DENSE if dense else 0, n, len(words))]
    if not dense:
        parts.append(count.astype("<u4").tobytes())
    parts.append(words.astype("<u4").tobytes())
    crc = 0
    for p in parts:
        crc = zlib.crc32(p, crc)
    parts.append(struct.pack("<I", crc))
    with open(path, "wb") as f:
        for p in parts:
            f.write(p)
    return sum(len(p) for p in parts)


def load(path: str, pad_zero: bool = False) -> Tuple[np.ndarray, np.ndarray, int, int]:
    """(sketches[n, stride], count[n], KmerSize, SketchSize); stride = SketchSize with pad_zero
    (zero tail as in a fresh Mash), else the largest count."""
    blob = open(path, "rb").read()
    if len(blob) < _HDR.size + 4:
        raise ValueError("truncated sketch file")
    magic, version, k, s, flags, n, words = _HDR.unpack_from(blob)
    if magic != MAGIC or version != VERSION:
        raise ValueError("not a PGSKETCH v1 file")
    if zlib.crc32(blob[:-4]) != struct.unpack("<I", blob[-4:])[0]:
        raise ValueError("sketch file checksum mismatch")
    pos = _HDR.size
    if flags & FLAG_DENSE:
        count = np.full(n, s, dtype=np.uint32)
    else:
        count = np.frombuffer(blob, dtype="<u4", count=n, offset=pos).astype(np.uint32)
        pos += 4 * n
    if int(count.sum(dtype=np.uint64)) != words or pos + 4 * words + 4 != len(blob):
        raise ValueError("sketch file is inconsistent")
    flat = np.frombuffer(blob, dtype="<u4", count=words, offset=pos).astype(np.uint32)
    stride = s if pad_zero else (int(count.max()) if n else 0)
    out = np.zeros((n, max(stride, 1)), dtype=np.uint32)
    mask = np.arange(out.shape[1], dtype=np.uint32)[None, :] < count[:, None]
    out[mask] = flat
    return out, count, k, s
