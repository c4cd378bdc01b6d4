"""Multi-GPU plumbing for independent filter instances: one process per GPU, filters
partitioned by index, and one all_gather of the fixed-size per-filter result records -- the only collective
of the path.  Works on any torch.distributed backend (NCCL on the GPU box, gloo in the CPU tests)."""
import numpy as np


def assign_filters(n_filters_total, world_size, rank):
    """filter f lives on rank f mod world_size (BASELINE config 4: 256 filters over 1/2/4/8 GPUs)."""
    return [f for f in range(n_filters_total) if f % world_size == rank]


def gather_records(local_records, n_filters_total, world_size, rank, dist=None):
    """local_records: uint8 tensor [n_local * record_bytes] on this rank's device, in the order of
    assign_filters(...).  Returns a uint8 numpy array [n_filters_total, record_bytes] in global filter
    order (on every rank).  Ranks may own different numbers of filters; buffers are padded to the max."""
    import torch
    n_local = len(assign_filters(n_filters_total, world_size, rank))
    rec_bytes = local_records.numel() // max(n_local, 1)
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
