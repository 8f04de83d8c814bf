"""What the host takes from N GPUs at once: every rank moves the bytes one bench tick delivers (and 4x that) into
page-locked host memory, all ranks starting together, by kernel stores (the march's way) and by cudaMemcpyAsync.

    python tools/link_ceiling.py                                  # one GPU
    torchrun --nproc-per-node 8 tools/link_ceiling.py             # eight at once

Prints one JSON line (rank 0): GB/s per GPU (min / mean over ranks) and in total, per mode and size. The e2e arm of
bench.py cannot deliver a tick's results faster than bytes_per_tick / that."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stage_b200 import rt  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = rt.Context(50.0, -1, -1, 2, 2, max_models=8, max_blocks=8, max_cell_entries=1 << 10, device=local)
    out = {}
    for label, nbytes in (("tick_9.72MB", 9720000), ("tick_x4_38.9MB", 38880000)):
        for mode, name in ((0, "kernel_stores"), (1, "memcpy_d2h")):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            g = ctx.link_bandwidth(mode, nbytes, reps=40)
            t = torch.tensor([g], dtype=torch.float64, device="cuda")
            if world > 1:
                gathered = [torch.zeros_like(t) for _ in range(world)]
                dist.all_gather(gathered, t)
                vals = [float(x.item()) for x in gathered]
            else:
                vals = [g]
            out["%s/%s" % (name, label)] = {"per_gpu_min": min(vals), "per_gpu_mean": float(np.mean(vals)), "total": float(np.sum(vals))}
    if rank == 0:
        print(json.dumps({"n_gpus": world, "unit": "GB/s", "host_cpus": os.cpu_count(), "results": out}))
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
