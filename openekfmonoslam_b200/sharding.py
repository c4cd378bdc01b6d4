"""Multi-GPU plumbing for independent filter instances: one process per GPU, filters
partitioned by index, and one all_gather of the fixed-size per-filter result records -- the only collective
of the path.  Works on any torch.distributed backend (NCCL on the GPU box, gloo in the CPU tests)."""
import numpy as np


def assign_filters(n_filters_total, world_size, rank):
    """filter f lives on rank f mod world_size (BASELINE config 4: 256 filters over 1/2/4/8 GPUs)."""
    return [f for f in range(n_filters_total) if f % world_size == rank]


def gather_records(local_records, n_filters_total, world_size, rank, dist=None, rec_bytes=None):
    """local_records: uint8 tensor [n_local * rec_bytes] on this rank's device, in the order of
    assign_filters(...).  Returns a uint8 numpy array [n_filters_total, rec_bytes] in global filter
    order (on every rank).  Ranks may own different numbers of filters -- also none, when there are fewer filters than
    ranks -- so the record size is passed in (capi.RECORD_BYTES), never inferred from a possibly empty local buffer;
    buffers are padded to the largest shard."""
    import torch
    n_local = len(assign_filters(n_filters_total, world_size, rank))
    if rec_bytes is None:
        from .capi import RECORD_BYTES as rec_bytes
    assert local_records.numel() == n_local * rec_bytes, "local record buffer does not match this rank's shard"
    n_max = (n_filters_total + world_size - 1) // world_size
    if world_size == 1:
        return local_records.detach().cpu().numpy().reshape(n_filters_total, rec_bytes)
    padded = torch.zeros(n_max * rec_bytes, dtype=torch.uint8, device=local_records.device)
    padded[: n_local * rec_bytes] = local_records
    out = torch.zeros(world_size * n_max * rec_bytes, dtype=torch.uint8, device=local_records.device)
    dist.all_gather_into_tensor(out, padded)
    host = out.cpu().numpy().reshape(world_size, n_max, rec_bytes)
    res = np.zeros((n_filters_total, rec_bytes), np.uint8)
    for r in range(world_size):
        for i, f in enumerate(assign_filters(n_filters_total, world_size, r)):
            res[f] = host[r, i]
    return res
