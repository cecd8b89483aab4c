"""GPU tier: the single-process multi-GPU entry points (pg_*_multi, pg_thread_device) and the
thread-safety of the boundary.  Device lists adapt to what the box shows: with one GPU the shard /
rendezvous / peer-store code still runs (one shard, own buffer); with N GPUs the peer stores cross
NVLink.  Everything is checked bit for bit against the oracle."""
import ctypes as C
import threading

import numpy as np
import pytest

from poly_b200 import _lib, mash, synth

pytestmark = pytest.mark.gpu


def n_devices():
    n = C.c_int(0)
    _lib.lib().pg_device_count(C.byref(n))
    return n.value


def device_sets():
    n = n_devices()
    sets = [None, [0]]
    if n >= 2:
        sets += [[1, 0], list(range(n))]
    return sets


@pytest.mark.parametrize("n,L,k,s", [(1000, 150, 21, 1000), (333, 150, 22, 1000), (37, 3000, 21, 256), (5, 150, 21, 1000)])
def test_sketch_uniform_multi_matches_oracle(gpu, oracle, n, L, k, s):
    reads = synth.family_reads(n, L, family=5)
    rc, want = oracle.sketch_batch(reads, synth.uniform_offsets(n, L), k, s, variant=1)
    assert rc == 0
    cnt = min(L - k, s)
    for devs in device_sets():
        got = mash.sketch_uniform_multi(reads, n, L, k, s, devices=devs)
        assert np.array_equal(got, want[:, :cnt]), devs


def test_sketch_uniform_multi_pinned_buffers(gpu, oracle):
    n, L, k, s = 4096 + 17, 150, 21, 1000
    lib = gpu.lib()
    hin, hout = C.c_void_p(), C.c_void_p()
    node = C.c_int(-7)
    gpu.check(lib.pg_numa_bind_thread(0, C.byref(node)))
    assert node.value >= -1
    gpu.check(lib.pg_host_alloc(C.byref(hin), n * L))
    gpu.check(lib.pg_host_alloc(C.byref(hout), n * (L - k) * 4))
    reads = synth.independent_reads(n, L)
    C.memmove(hin.value, reads.ctypes.data, n * L)
    gpu.check(lib.pg_mash_sketch_uniform_multi(hin.value, n, L, k, s, 0, hout.value, L - k, None, None, 0))
    got = np.ctypeslib.as_array((C.c_uint32 * (n * (L - k))).from_address(hout.value)).reshape(n, L - k).copy()
    rc, want = oracle.sketch_batch(reads, synth.uniform_offsets(n, L), k, s, variant=1)
    assert rc == 0 and np.array_equal(got, want[:, : L - k])
    gpu.check(lib.pg_host_free(hin)); gpu.check(lib.pg_host_free(hout))


def test_sketch_batch_multi_ragged(gpu, oracle):
    rng = np.random.default_rng(5)
    lens = rng.integers(0, 400, 700)
    lens[::50] = 3000                                    # a few select-regime reads among the short ones
    offs = np.zeros(len(lens) + 1, dtype=np.uint64); offs[1:] = np.cumsum(lens)
    bases = synth.independent_reads(int(offs[-1] // 150 + 1), 150)[: int(offs[-1])]
    k, s = 21, 256
    rc, want = oracle.sketch_batch(bases, offs, k, s, variant=1)
    assert rc == 0
    for devs in device_sets():
        out, count, status = mash.sketch_arrays_multi(bases, offs, k, s, pad_zero=True, devices=devs)
        assert (status == 0).all() and np.array_equal(count, np.minimum(np.maximum(lens - k, 0), s))
        assert np.array_equal(out, want), devs


@pytest.mark.parametrize("n,L,k,s,family", [(96, 3000, 21, 256, 8), (70, 150, 21, 1000, 4), (130, 1200, 31, 200, 10)])
def test_sketch_distance_multi_matches_oracle(gpu, oracle, n, L, k, s, family):
    reads = synth.family_reads(n, L, family=family)
    rc, want = oracle.sketch_batch(reads, synth.uniform_offsets(n, L), k, s, variant=1)
    assert rc == 0
    want_same = np.zeros((n, n), dtype=np.uint32)
    want_dist = np.zeros((n, n), dtype=np.float64)
    ms = []
    for i in range(n):
        m = oracle.OracleMash(k, s); m.Sketches[:] = want[i]; ms.append(m)
    for i in range(n):
        for j in range(n):
            want_same[i, j] = ms[i].SimilarityCount(ms[j])[0]
            want_dist[i, j] = ms[i].Distance(ms[j])
    for devs in device_sets():
        sk, same, dist = mash.sketch_distance_multi(reads, n, L, k, s, devices=devs, want_distance=True)
        assert np.array_equal(sk, want), devs
        assert np.array_equal(same, want_same), devs
        assert np.array_equal(dist, want_dist), devs


def test_multi_rejects_bad_device_lists(gpu):
    reads = synth.independent_reads(64, 150)
    with pytest.raises(_lib.PolyError):
        mash.sketch_uniform_multi(reads, 64, 150, 21, 1000, devices=[0, 0])
    with pytest.raises(_lib.PolyError):
        mash.sketch_uniform_multi(reads, 64, 150, 21, 1000, devices=[n_devices()])


def test_thread_device_binding(gpu, oracle):
    """pg_thread_device: a thread bound to device d runs its calls there; results are the same."""
    n, L, k, s = 256, 150, 21, 1000
    reads = synth.independent_reads(n, L)
    rc, want = oracle.sketch_batch(reads, synth.uniform_offsets(n, L), k, s, variant=1)
    results = {}

    def work(d):
        gpu.check(gpu.lib().pg_thread_device(d))
        results[d] = mash.sketch_uniform(reads, n, L, k, s)
        gpu.check(gpu.lib().pg_thread_device(-1))

    ts = [threading.Thread(target=work, args=(d,)) for d in range(n_devices())]
    [t.start() for t in ts]; [t.join() for t in ts]
    for d in range(n_devices()):
        assert np.array_equal(results[d], want[:, : L - k]), d


def test_concurrent_callers(gpu, oracle):
    """8 host threads hammer the host-pointer entry points (fill, select, ragged) while one more
    drives the *_dev entry points on its own stream: the boundary must be callable from many OS
    threads at once (goroutines -> cgo threads), SURVEY.md 8b "Threading"."""
    import torch

    L1, k1, s1 = 150, 21, 1000
    L2, k2, s2 = 3000, 21, 256
    r1 = synth.independent_reads(512, L1)
    r2 = synth.family_reads(24, L2, family=4)
    _, w1 = oracle.sketch_batch(r1, synth.uniform_offsets(512, L1), k1, s1, variant=1)
    _, w2 = oracle.sketch_batch(r2, synth.uniform_offsets(24, L2), k2, s2, variant=1)
    errs = []

    def host_worker(i):
        try:
            for it in range(6):
                if (i + it) % 2 == 0:
                    assert np.array_equal(mash.sketch_uniform(r1, 512, L1, k1, s1), w1[:, : L1 - k1])
                else:
                    out, cnt, st = mash.sketch_arrays(r2, synth.uniform_offsets(24, L2), k2, s2)
                    assert np.array_equal(out, w2)
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    def dev_worker():
        try:
            dev = torch.device("cuda", 0)
            st = torch.cuda.Stream(device=dev)
            d_in = torch.from_numpy(r1).to(dev)
            d_in2 = torch.from_numpy(r2).to(dev)
            for it in range(12):
                with torch.cuda.stream(st):
                    d_out = torch.empty((512, L1 - k1), dtype=torch.int32, device=dev)
                    gpu.check(gpu.lib().pg_mash_sketch_uniform_dev(d_in.data_ptr(), 512, L1, k1, s1, 0, d_out.data_ptr(), L1 - k1, None, st.cuda_stream))
                    d_out2 = torch.empty((24, s2), dtype=torch.int32, device=dev)
                    gpu.check(gpu.lib().pg_mash_sketch_uniform_dev(d_in2.data_ptr(), 24, L2, k2, s2, 0, d_out2.data_ptr(), s2, None, st.cuda_stream))
                st.synchronize()
                assert np.array_equal(d_out.cpu().numpy().view(np.uint32), w1[:, : L1 - k1])
                assert np.array_equal(d_out2.cpu().numpy().view(np.uint32), w2)
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    ts = [threading.Thread(target=host_worker, args=(i,)) for i in range(8)] + [threading.Thread(target=dev_worker)]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert not errs, errs
