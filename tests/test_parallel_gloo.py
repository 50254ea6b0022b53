"""world_size = 2 (and 3) gloo runs of the multi-rank host logic on CPU: contiguous packet shards,
rank-order gather reproducing the single-rank row order bit for bit, max-over-ranks timing."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from juliaraytracingsw_b200 import parallel
from oracle import raytrace as oray


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, sqrtN, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N = sqrtN * sqrtN
        lo, hi = parallel.shard_range(N, rank, world)
        xk, sign = parallel.initial_wavepackets_host(2 * np.pi, 5.196152422706632, sqrtN, lo, hi - lo)
        full = parallel.gather_rows(xk, dist)
        full_sign = parallel.gather_rows(sign, dist)
        tmax = parallel.max_over_ranks(10.0 + rank, dist)
        if rank == 0:
            q.put((full, full_sign, tmax))
        else:
            assert full is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,sqrtN", [(2, 12), (3, 7)])
def test_sharded_generation_and_gather_match_single_rank(world, sqrtN):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, sqrtN, q)) for r in range(world)]
    for p in procs:
        p.start()
    full, full_sign, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want, sign = oray.generate_initial_wavepackets(2 * np.pi, 5.196152422706632, sqrtN)
    np.testing.assert_array_equal(full[:, 0:2], want[:, 0:2])
    np.testing.assert_array_equal(full[:, 2:4], want[:, 2:4])
    np.testing.assert_array_equal(full_sign, sign)
    assert tmax == 10.0 + world - 1


def test_shard_ranges_partition_exactly():
    for n in (1, 7, 4096, 16777216):
        for w in (1, 2, 3, 4, 8):
            edges = [parallel.shard_range(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))


def test_pipeline_chunks_are_the_rank_shard_rule():
    """PacketPipeline's row blocks use `shard_range`: contiguous, ordered, exact cover, also when ragged."""
    from juliaraytracingsw_b200.parallel import shard_range
    for n, c in ((2500, 7), (16777216, 8), (5, 8), (1, 1)):
        b = [shard_range(n, i, c) for i in range(c)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(c - 1)) and all(lo <= hi for lo, hi in b)
