"""One process per GPU: robots (fans) are sharded over the ranks, the grid is replicated, and the
only exchange step per tick is the moved-block deltas.

Each rank serialises the deltas its own robots produced into a fixed-size slot
(stg_rt_delta_export), the slots are all-gathered (NCCL over NVLink through torch.distributed; gloo
in the CPU tests) and every rank applies all slots, in rank order, straight out of the gathered
buffer (stg_rt_delta_apply_gathered -> K1). Rank order makes every replica - and a single-GPU run
that issues the same maps rank by rank - hold the same cell lists.
"""
import numpy as np


def shard_bounds(n_items, rank, world_size):
    """contiguous shard [lo, hi) of n_items for this rank (sizes differ by at most one)"""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class _CudaArray:
    """zero-copy view of a raw device pointer for torch.as_tensor"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def device_blob_as_tensor(ptr, nbytes):
    import torch
    return torch.as_tensor(_CudaArray(ptr, nbytes), device="cuda")


class DeltaExchange:
    """The per-tick all-gather of moved-block deltas."""

    def __init__(self, ctx, rank, world_size, dist=None, device="cuda", as_tensor=device_blob_as_tensor):
        import torch
        self.ctx, self.rank, self.world_size, self.dist = ctx, rank, world_size, dist
        self.slot = int(ctx.delta_slot_bytes())
        self.as_tensor = as_tensor
        self.gathered = torch.empty(self.slot * world_size, dtype=torch.uint8, device=device)
        self.bytes_exported = 0

    def step(self):
        """export this rank's queued deltas, all-gather, apply every rank's in rank order"""
        ptr, nbytes = self.ctx.delta_export(self.rank)
        self.bytes_exported += nbytes
        mine = self.as_tensor(ptr, self.slot)
        if self.world_size > 1:
            self.dist.all_gather_into_tensor(self.gathered, mine)
        else:
            self.gathered.copy_(mine)
        self.ctx.delta_apply_gathered(self.gathered.data_ptr(), self.world_size, self.slot)
