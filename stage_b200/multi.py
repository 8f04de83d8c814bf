"""One process per GPU: robots (fans) are sharded over the ranks, the grid is replicated, and the
only exchange step per tick is the moved-block deltas.

Each rank serialises the deltas its own robots produced into a fixed-size slot
(stg_rt_delta_export), the slots are all-gathered (NCCL over NVLink through torch.distributed; gloo
in the CPU tests) and every rank applies all slots, in rank order, straight out of the gathered
buffer (stg_rt_delta_apply_gathered -> K1). Rank order makes every replica - and a single-GPU run
that issues the same maps rank by rank - hold the same cell lists.

Fiducial finders (BASELINE config 4) need one more thing on every rank: the poses of ALL models that carry a
fiducial, not only the rank's own (World::models_with_fiducials, world.cc:623-629, is world-wide). They ride in a
second all-gather of fixed-size records (TargetExchange): every rank contributes its own robots' stg_rt_fid_target
records and gets the world's, in rank order - which is model-id order for contiguous shards, the order the
reference visits candidates in (ascending Model*, model_fiducial.cc:245-283).
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

    def _collective_stream(self):
        """the stream torch queues this process's collectives and copies on (the device tests pass a fake ctx / cpu tensors)"""
        if self.gathered.device.type != "cuda":
            return None
        import torch
        return torch.cuda.current_stream(self.gathered.device).cuda_stream

    def step(self):
        """export this rank's queued deltas, all-gather, apply every rank's in rank order"""
        ptr, nbytes = self.ctx.delta_export(self.rank)
        self.bytes_exported += nbytes
        mine = self.as_tensor(ptr, self.slot)
        # The blob is uploaded on the context's stream, the collective runs on torch's current stream and K1 reads
        # the gathered buffer on the context's stream again: order the three explicitly (include/stg_rt.h,
        # stg_rt_delta_export). Both calls are no-ops when the context has been given torch's stream.
        cs = self._collective_stream()
        if cs is not None:
            self.ctx.stream_signal(cs)
        if self.world_size > 1:
            self.dist.all_gather_into_tensor(self.gathered, mine)
        else:
            self.gathered.copy_(mine)
        if cs is not None:
            self.ctx.stream_wait(cs)
        self.ctx.delta_apply_gathered(self.gathered.data_ptr(), self.world_size, self.slot)


class TargetExchange:
    """The per-tick all-gather of fiducial target records (stg_rt_fid_target, 72 bytes each). Shards may differ in
    size: every rank's share is padded to `max_per_rank` records, the true counts travel in a small header gather."""

    def __init__(self, rank, world_size, max_per_rank, record_dtype, dist=None, device="cuda"):
        import torch
        self.rank, self.world_size, self.dist, self.device = rank, world_size, dist, device
        self.dtype = np.dtype(record_dtype)
        self.cap = int(max_per_rank)
        self.mine = torch.zeros(self.cap * self.dtype.itemsize, dtype=torch.uint8, device=device)
        self.all = torch.zeros(world_size * self.cap * self.dtype.itemsize, dtype=torch.uint8, device=device)
        # page-locked staging for this rank's own records; a slot's unused tail is padding records (model id 0xffffffff)
        self.stage = torch.zeros(self.cap * self.dtype.itemsize, dtype=torch.uint8)
        if device != "cpu":
            self.stage = self.stage.pin_memory()
        pad = np.zeros(self.cap, self.dtype)
        pad["model"] = 0xffffffff
        self.pad = pad
        self.count = torch.zeros(1, dtype=torch.int64, device=device)
        self.counts = torch.zeros(world_size, dtype=torch.int64, device=device)
        self.bytes_sent = 0
        self._staged = -1   # how many records the staging buffer's padding was laid out for

    def step_device(self, my_targets, ctx=None):
        """my_targets: this rank's records -> (device pointer, n_records) of the world's, rank after rank, each rank's slot
        padded to `max_per_rank` with records the finders skip: nothing but this rank's own share crosses PCIe, and nothing
        comes back. For stg_rt_fiducials_submit_device_targets; ctx (if given) is made to wait for the collective."""
        import torch
        my_targets = np.ascontiguousarray(my_targets, self.dtype)
        n = len(my_targets)
        if n > self.cap:
            raise ValueError("%d target records exceed the exchange slot of %d" % (n, self.cap))
        # bytes, not records: numpy assigns structured arrays field by field (4 ms for 100,000 records against 0.7 ms)
        host, size = self.stage.numpy(), self.dtype.itemsize
        host[:n * size] = my_targets.view(np.uint8).reshape(-1)
        if n != self._staged:   # the slot's tail beyond this rank's records: padding (written again only when the count changes)
            host[n * size:] = self.pad.view(np.uint8).reshape(-1)[n * size:]
            self._staged = n
        dst = self.mine if self.world_size > 1 else self.all   # alone, the slot is the whole array
        dst.copy_(self.stage, non_blocking=True)
        self.bytes_sent += n * size
        if self.world_size > 1:
            self.dist.all_gather_into_tensor(self.all, self.mine)
        if ctx is not None and self.all.device.type == "cuda":
            ctx.stream_wait(torch.cuda.current_stream(self.all.device).cuda_stream)
        return self.all.data_ptr(), self.world_size * self.cap

    def host_view(self, my_targets_by_rank):
        """the array step_device builds, on the host, from every rank's records (tests / checkers)"""
        parts = []
        for t in my_targets_by_rank:
            slot = self.pad.copy()
            slot[:len(t)] = np.ascontiguousarray(t, self.dtype)
            parts.append(slot)
        return np.concatenate(parts)

    def step(self, my_targets):
        """my_targets: this rank's records (numpy, record_dtype) -> every rank's, concatenated in rank order"""
        import torch
        my_targets = np.ascontiguousarray(my_targets, self.dtype)
        n = len(my_targets)
        if n > self.cap:
            raise ValueError("%d target records exceed the exchange slot of %d" % (n, self.cap))
        raw = torch.from_numpy(my_targets.view(np.uint8).reshape(-1))
        self.mine[:raw.numel()].copy_(raw)
        self.count[0] = n
        self.bytes_sent += raw.numel()
        if self.world_size > 1:
            self.dist.all_gather_into_tensor(self.counts, self.count)
            self.dist.all_gather_into_tensor(self.all, self.mine)
        else:
            self.counts.copy_(self.count)
            self.all.copy_(self.mine)
        counts = self.counts.cpu().numpy()
        flat = self.all.cpu().numpy().reshape(self.world_size, self.cap * self.dtype.itemsize)
        parts = [flat[r, :int(counts[r]) * self.dtype.itemsize].view(self.dtype) for r in range(self.world_size)]
        return np.concatenate(parts) if parts else np.zeros(0, self.dtype)
