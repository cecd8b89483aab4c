"""GPU tier of the multi-GPU layer: the sharded sketch -> all-gather -> row-block distance
pipeline with the CUDA compute over a NCCL process group (world size = the GPUs this test
process was given: 1 under the driver's `pytest -m gpu`, N under torchrun)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist

from poly_b200 import synth
from poly_b200.dist import ShardPlan, sharded_sketch_distance

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pg(gpu):
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    if not dist.is_initialized():
        if "MASTER_PORT" not in os.environ:
            s = socket.socket(); s.bind(("127.0.0.1", 0)); os.environ["MASTER_PORT"] = str(s.getsockname()[1]); s.close()
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", rank=rank, world_size=world)
    yield rank, world
    dist.destroy_process_group()


@pytest.mark.parametrize("n,L,k,s,family", [(64, 3000, 21, 256, 8), (40, 150, 21, 1000, 4)])
def test_sharded_pipeline_matches_oracle(pg, oracle, n, L, k, s, family):
    rank, world = pg
    plan = ShardPlan(n, rank, world)
    dev = torch.device("cuda", torch.cuda.current_device())
    reads_all = synth.family_reads(n, L, family=family)
    local_reads = torch.from_numpy(reads_all[plan.lo * L: plan.hi * L]).to(dev)
    local, gathered, same = sharded_sketch_distance(local_reads, plan, L, k, s)
    torch.cuda.synchronize()
    rc, want = oracle.sketch_batch(reads_all, synth.uniform_offsets(n, L), k, s, variant=1)
    assert rc == 0
    assert np.array_equal(gathered.cpu().numpy().view(np.uint32), want)            # full Go arrays, zero tail included
    got = same.cpu().numpy()
    for i in range(plan.lo, plan.hi):
        a = oracle.OracleMash(k, s); a.Sketches[:] = want[i]
        for j in range(0, n, 3):
            b = oracle.OracleMash(k, s); b.Sketches[:] = want[j]
            assert got[i - plan.lo, j] == a.SimilarityCount(b)[0]


@pytest.mark.parametrize("n_local,L,k,s", [(64, 150, 21, 1000), (96, 151, 17, 500), (20, 3000, 21, 128), (45, 150, 21, 1000), (33, 150, 22, 64)])
def test_fused_sketch_gather_equals_allgather(pg, oracle, n_local, L, k, s):
    """pg_mash_sketch_uniform_gather_dev: every rank's gathered buffer == all-gather of the
    per-rank sketches == the oracle on the concatenated reads (fill regime via TMA bulk stores
    to peer memory, select regime via per-row peer stores)."""
    from poly_b200.dist import GatheredBuffer, fused_sketch_gather

    rank, world = pg
    n = n_local * world + (world > 1)          # uneven shards when there is more than one rank
    plan = ShardPlan(n, rank, world)
    dev = torch.device("cuda", torch.cuda.current_device())
    reads_all = synth.family_reads(n, L, family=5)
    local_reads = torch.from_numpy(reads_all[plan.lo * L: plan.hi * L]).to(dev)
    cnt = min(max(L - k, 0), s)
    buf = GatheredBuffer(plan, cnt)
    fused_sketch_gather(local_reads, L, k, s, buf)
    got = buf.to_numpy()
    assert np.array_equal(buf.as_tensor().cpu().numpy().view(np.uint32), got)
    buf.close()
    rc, want = oracle.sketch_batch(reads_all, synth.uniform_offsets(n, L), k, s, variant=1)
    assert rc == 0 and np.array_equal(got, want[:, :cnt])
