"""Packet sharding across GPUs (one process per GPU, ``torch.distributed``).

Packets are independent (reference rk5 is row-wise; ``Input.run`` already loops
chunks serially, ``initial_state/Input.py:243-249``), so a run is sharded by
contiguous GLOBAL packet-id ranges; the Philox counter is the global id, hence
results do not depend on the number of GPUs.  No data-path collective; the only
exchange is one all-reduce (sum) per product: image f64 + counts i64, LOS radiance
f64 + hit counts i64.  NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""


def shard_range(n_total, rank, world):
    """[first_id, first_id + n) of `rank`: sizes differ by at most one."""
    base, rem = divmod(int(n_total), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def allreduce_products(*tensors):
    """In-place SUM all-reduce of result tensors (image, counts, radiance, ...)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for t in tensors:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return tensors
