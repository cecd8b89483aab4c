"""N>1 path on CPU: world_size-2 gloo run of poly_b200.dist.sharded_sketch_distance with the
oracle injected as the compute (the partition / all-gather / row-block bookkeeping is what is
under test; the CUDA compute is covered by the -m gpu tier)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from poly_b200 import synth
from poly_b200.dist import ShardPlan, all_gather_rows, shard_range, sharded_sketch_distance


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 100, 12345):
        for w in (1, 2, 3, 8):
            edges = [shard_range(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def _oracle_fns():
    import oracle_ffi as o

    def sketch_fn(reads, n, L, k, s):
        rc, out = o.sketch_batch(reads.numpy(), synth.uniform_offsets(n, L), k, s, variant=1)
        assert rc == 0
        cnt = min(max(L - k, 0), s)
        return torch.from_numpy(out[:, :cnt].view(np.int32).copy())

    def distance_fn(gathered, lo, hi):
        g = gathered.numpy().view(np.uint32)
        n, s = g.shape
        same = np.zeros((hi - lo, n), np.int32)
        for i in range(lo, hi):
            a = o.OracleMash(0, s); a.Sketches[:] = g[i]
            for j in range(n):
                b = o.OracleMash(0, s); b.Sketches[:] = g[j]
                same[i - lo, j] = a.SimilarityCount(b)[0]
        return torch.from_numpy(same)

    return sketch_fn, distance_fn


def _worker(rank, world, port, n, L, k, s, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = ShardPlan(n, rank, world)
        reads = torch.from_numpy(synth.family_reads(plan.hi - plan.lo, L, family=3, first_read=plan.lo))
        sketch_fn, distance_fn = _oracle_fns()
        local, gathered, same = sharded_sketch_distance(reads, plan, L, k, s, sketch_fn, distance_fn)
        q.put((rank, plan.lo, plan.hi, local.numpy(), gathered.numpy(), same.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,L,k,s", [(11, 400, 21, 64), (8, 150, 21, 200)])
def test_world2_gloo_pipeline_equals_single_process(n, L, k, s):
    sock = socket.socket(); sock.bind(("127.0.0.1", 0)); port = sock.getsockname()[1]; sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, L, k, s, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference of the same pipeline
    sketch_fn, distance_fn = _oracle_fns()
    reads = torch.from_numpy(synth.family_reads(n, L, family=3))
    _, g1, s1 = sharded_sketch_distance(reads, ShardPlan(n, 0, 1), L, k, s, sketch_fn, distance_fn)
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == n
    for r in res:
        assert np.array_equal(r[4], g1.numpy())                       # every rank holds the full gathered set
        assert np.array_equal(r[5], s1.numpy()[r[1]: r[2]])            # row block == rows of the full matrix
    assert np.array_equal(np.concatenate([res[0][3], res[1][3]]), g1.numpy()[:, : res[0][3].shape[1]])
    if L - k >= s:
        assert s1.numpy().max() == s  # diagonal
