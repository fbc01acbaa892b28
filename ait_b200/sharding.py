"""Multi-GPU: the path shards by (image, query) unit with NO data-path collective.
Units are split into contiguous blocks, one block per rank; the only exchange is the final
host-side gather of [rois, cls_prob, bbox_pred] (12 KB per unit)."""


def shard_units(n_units, rank, world_size):
    """Contiguous block of unit indices owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_units, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def gather_results(local, world_size, group=None):
    """All ranks contribute a picklable per-unit result list; rank 0 receives the concatenation in unit
    order (torch.distributed gather_object on the host: gloo in tests, NCCL-backed group in the bench)."""
    import torch.distributed as dist
    if world_size == 1 or not dist.is_initialized():
        return list(local)
    out = [None] * world_size if dist.get_rank(group) == 0 else None
    dist.gather_object(local, out, dst=0, group=group)
    if out is None:
        return None
    return [x for part in out for x in part]
