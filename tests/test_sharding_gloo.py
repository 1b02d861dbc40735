"""The N>1 host logic on CPU: world_size-2 `gloo` process group, shard ranges from the
C ABI, max-over-ranks timing and digest combination -- what bench.py does under torchrun,
minus the GPU.  No data-path collective exists to test: shards are independent."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n: int, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    import oracle
    from k8s_gpu_hpa_b200 import sharding

    assert sharding.world() == (rank, world, rank)
    assert sharding.init("gloo")
    b, e = sharding.shard_range(n, world, rank)
    # each rank generates ITS shard from the global index and digests its own sum
    a_, b_ = oracle.fill_ctr(e - b, 0x0A, b), oracle.fill_ctr(e - b, 0x0B, b)
    local = oracle.bits_digest(oracle.vadd(a_, b_))
    sharding.barrier()
    glob = sharding.combine_digests(local)
    assert sharding.gather_ints([rank, 7]) == [[r, 7] for r in range(world)]
    slowest = sharding.max_over_ranks(10.0 + rank)
    total = sharding.sum_over_ranks(float(e - b))
    out.put((rank, b, e, glob, slowest, total))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1_000_003, 1 << 16])
def test_two_rank_shards_combine_to_the_unsharded_answer(n):
    import oracle

    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = oracle.bits_digest(oracle.vadd(oracle.fill_ctr(n, 0x0A), oracle.fill_ctr(n, 0x0B)))
    assert got[0][1] == 0 and got[0][2] == got[1][1] and got[1][2] == n      # shards tile [0, n)
    for _, _, _, glob, slowest, total in got:
        assert glob == want                                                   # sharding is invisible in the data
        assert slowest == 11.0                                                # max over ranks
        assert total == float(n)


def test_single_process_fallbacks():
    from k8s_gpu_hpa_b200 import sharding

    assert sharding.init("gloo") is False          # no torchrun environment: stays single-process
    assert sharding.max_over_ranks(3.5) == 3.5
    assert sharding.combine_digests((5, 7)) == (5, 7)
