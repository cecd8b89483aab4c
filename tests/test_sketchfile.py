"""Sketch persistence: the JSON shape of mash.Mash and the PGSKETCH container.
Host-side marshalling only: runs on the CPU tier (no library calls)."""
import json

import numpy as np
import pytest

from poly_b200 import sketchfile
from poly_b200.mash import Mash


def test_json_is_what_go_marshals():
    m = Mash(17, 4)
    m.Sketches[:] = [0x096698DE, 1, 0, 0xFFFFFFFF]
    text = sketchfile.mash_to_json(m)
    assert text == '{"KmerSize":17,"SketchSize":4,"Sketches":[157718750,1,0,4294967295]}'   # mash.go:52-56 field order
    back = sketchfile.mash_from_json(text)
    assert (back.KmerSize, back.SketchSize, back.Sketches.tolist()) == (17, 4, m.Sketches.tolist())
    assert sketchfile.mash_from_json('{"KmerSize":3,"SketchSize":0,"Sketches":null}').Sketches.size == 0
    assert json.loads(text)["Sketches"][0] == 0x096698DE


def test_container_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    n, s = 1000, 1000
    count = rng.integers(0, 130, n).astype(np.uint32)
    sk = rng.integers(0, 2**32, (n, 129), dtype=np.uint64).astype(np.uint32)
    sk[np.arange(129)[None, :] >= count[:, None]] = 0
    p = str(tmp_path / "a.pgsk")
    size = sketchfile.save(p, sk, count, 21, s)
    assert size == 40 + 4 * n + 4 * int(count.sum()) + 4
    out, cnt, k, ss = sketchfile.load(p)
    assert (k, ss) == (21, s) and np.array_equal(cnt, count) and np.array_equal(out, sk[:, : out.shape[1]])
    padded, _, _, _ = sketchfile.load(p, pad_zero=True)
    assert padded.shape == (n, s) and np.array_equal(padded[:, :129], sk) and not padded[:, 129:].any()
    dense = np.sort(rng.integers(0, 2**32, (50, 64), dtype=np.uint64).astype(np.uint32), axis=1)
    p2 = str(tmp_path / "b.pgsk")
    assert sketchfile.save(p2, dense, np.full(50, 64, np.uint32), 31, 64) == 40 + 4 * 50 * 64 + 4
    out, cnt, k, ss = sketchfile.load(p2)
    assert np.array_equal(out, dense) and (cnt == 64).all() and (k, ss) == (31, 64)
    p3 = str(tmp_path / "c.pgsk")
    sketchfile.save(p3, np.zeros((0, 1), np.uint32), np.zeros(0, np.uint32), 21, 1000)
    assert sketchfile.load(p3)[0].shape[0] == 0


def test_container_rejects_damage(tmp_path):
    p = str(tmp_path / "a.pgsk")
    sketchfile.save(p, np.arange(12, dtype=np.uint32).reshape(3, 4), np.array([4, 2, 0], np.uint32), 21, 1000)
    blob = bytearray(open(p, "rb").read())
    blob[45] ^= 1
    open(p, "wb").write(bytes(blob))
    with pytest.raises(ValueError, match="checksum"):
        sketchfile.load(p)
    open(p, "wb").write(bytes(blob[:30]))
    with pytest.raises(ValueError):
        sketchfile.load(p)


@pytest.mark.gpu
def test_saved_sketches_feed_the_distance_kernel(gpu, tmp_path):
    from poly_b200 import mash, synth
    n, L, k, s = 64, 600, 21, 200
    reads = synth.family_reads(n, L, family=8)
    offsets = (np.arange(n + 1) * L).astype(np.uint64)
    out, count, status = mash.sketch_arrays(reads, offsets, k, s)
    p = str(tmp_path / "set.pgsk")
    sketchfile.save(p, out, count, k, s)
    back, cnt, kk, ss = sketchfile.load(p, pad_zero=True)
    assert (kk, ss) == (k, s) and np.array_equal(cnt, count) and np.array_equal(back[:, : out.shape[1]], out)
    same0, dist0 = mash.distance_block(np.ascontiguousarray(out[:, :s]), 0, n)
    same1, dist1 = mash.distance_block(back, 0, n)
    assert np.array_equal(same0, same1) and np.array_equal(dist0, dist1)
    m = mash.New(k, s)
    m.Sketch(reads[:L])
    again = sketchfile.mash_from_json(sketchfile.mash_to_json(m))
    assert again.Distance(m) == 0 and np.array_equal(again.Sketches, m.Sketches)
