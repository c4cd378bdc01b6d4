"""The N > 1 host path on CPU: world_size-2 gloo processes shard filters f mod G and all_gather the
fixed-size records into global filter order (what bench.py does over NCCL on the GPU box)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from openekfmonoslam_b200.sharding import assign_filters, gather_records

REC = 1504  # sizeof(ekfb_record)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = assign_filters(n_total, world, rank)
    local = torch.zeros(len(mine) * REC, dtype=torch.uint8)
    for i, f in enumerate(mine):  # record content = a function of the global filter index
        local[i * REC:(i + 1) * REC] = torch.from_numpy(((np.arange(REC) * 7 + f * 13) % 251).astype(np.uint8))
    res = gather_records(local, n_total, world, rank, dist, rec_bytes=REC)
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [1, 5, 8])   # 1: a rank that owns no filter at all
def test_two_rank_gather_is_in_global_filter_order(n_total):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.stack([((np.arange(REC) * 7 + f * 13) % 251).astype(np.uint8) for f in range(n_total)])
    for rank, res in out:
        assert res.shape == (n_total, REC) and np.array_equal(res, expect), rank


def test_assignment_is_a_partition():
    for total, world in ((256, 8), (256, 4), (5, 2), (1, 1)):
        seen = sorted(f for r in range(world) for f in assign_filters(total, world, r))
        assert seen == list(range(total))
        sizes = [len(assign_filters(total, world, r)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
