"""Multi-GPU host logic (replicas only): seed sharding and the max-over-ranks reduction,
exercised with two gloo ranks on the CPU."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vil_fusion_b200 import replicas


def test_seeds_partition_the_job():
    all_seeds = []
    for r in range(4):
        s = replicas.sequence_seeds(r, 4, 3)
        assert len(s) == 3
        all_seeds += s
    assert sorted(all_seeds) == list(range(12))
    assert replicas.job_scans(4, 3, 10) == 120
    with pytest.raises(ValueError):
        replicas.sequence_seeds(4, 4, 1)


def test_max_over_ranks_single_process():
    assert replicas.max_over_ranks([1.5, 2.5]) == [1.5, 2.5]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seeds = replicas.sequence_seeds(rank, world, 2)
    replicas.barrier(sync_cuda=False)
    m = replicas.max_over_ranks([10.0 + rank, 5.0 - rank])
    q.put((rank, seeds, m))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 1] and res[1][1] == [2, 3]
    for _, _, m in res:  # every rank sees the same maxima
        assert m == [11.0, 5.0]
