"""CPU, world_size 2 over gloo: the N>1 host logic (image sharding + the single all_gather of counts)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from superpoint_nerf_pytorch_b200.utils.sharding import gather_export_counts, shard_indices
    mine = list(shard_indices(11, rank, world))
    allc, off = gather_export_counts(len(mine), sum(i * 10 for i in mine))
    q.put((rank, mine, allc.tolist(), off.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_and_count_allgather_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, c0, o0), (r1, m1, c1, o1) = res
    assert sorted(m0 + m1) == list(range(11)) and m0 == list(range(0, 11, 2))
    assert c0 == c1 == [[6, 300], [5, 250]]
    assert o0 == o1 == [0, 300]
