"""Host-side logic of the multi-GPU path on CPU: sharding, and the delta exchange run for real over a
2-rank gloo group with a stand-in for the context (the engine needs a GPU, the protocol does not)."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stage_b200 import multi


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 1000, 100001):
        for world in (1, 2, 3, 8):
            spans = [multi.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


class FakeCtx:
    """Mimics the three delta entry points of stage_b200.rt.Context on host memory."""
    SLOT = 4096

    def __init__(self, rank):
        self.rank = rank
        self.buf = np.zeros(self.SLOT, np.uint8)
        self.applied = []
        self.tick = 0

    def delta_slot_bytes(self):
        return self.SLOT

    def delta_export(self, rank):
        assert rank == self.rank
        self.tick += 1
        n = 100 + 10 * rank + self.tick
        self.buf[:] = 0
        self.buf[:n] = (np.arange(n) * (rank + 3) + self.tick) % 251
        return self.buf.ctypes.data, n

    def delta_apply_gathered(self, ptr, n_ranks, slot):
        raw = (ctypes.c_uint8 * (n_ranks * slot)).from_address(ptr)
        self.applied.append(np.frombuffer(raw, np.uint8).copy().reshape(n_ranks, slot))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = FakeCtx(rank)
    ex = multi.DeltaExchange(ctx, rank, world, dist=dist, device="cpu",
                             as_tensor=lambda ptr, n: torch.from_numpy(np.frombuffer((ctypes.c_uint8 * n).from_address(ptr), np.uint8)))
    for _ in range(3):
        ex.step()
    q.put((rank, [a.tolist() for a in ctx.applied], ex.bytes_exported))
    dist.barrier()
    dist.destroy_process_group()


def test_delta_exchange_over_gloo_two_ranks():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        rank, applied, nbytes = q.get(timeout=120)
        got[rank] = (np.array(applied, np.uint8), nbytes)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a0, a1 = got[0][0], got[1][0]
    assert a0.shape == (3, 2, FakeCtx.SLOT)
    assert np.array_equal(a0, a1)  # every rank applies the same bytes ...
    for tick in range(3):          # ... which are the ranks' own blobs, in rank order
        for rank in range(2):
            n = 100 + 10 * rank + (tick + 1)
            want = (np.arange(n) * (rank + 3) + (tick + 1)) % 251
            assert np.array_equal(a0[tick, rank, :n], want.astype(np.uint8))
            assert not a0[tick, rank, n:].any()
    assert got[0][1] == sum(100 + t for t in (1, 2, 3)) and got[1][1] == sum(110 + t for t in (1, 2, 3))


REC = np.dtype([("model", "<u4"), ("ret", "<i4"), ("key", "<i4"), ("pad", "<u4"), ("pose", "<f8", (4,)), ("size", "<f8", (3,))])
assert REC.itemsize == 72  # stg_rt_fid_target


def _target_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ex = multi.TargetExchange(rank, world, max_per_rank=6, record_dtype=REC, dist=dist, device="cpu")
    got = []
    for tick in range(3):
        lo, hi = multi.shard_bounds(11, rank, world)  # 11 robots over 2 ranks: shards of 6 and 5
        mine = np.zeros(hi - lo, REC)
        mine["model"] = 100 + np.arange(lo, hi)
        mine["ret"] = tick + 1
        mine["pose"][:, 0] = np.arange(lo, hi) + 0.25 * tick
        got.append(ex.step(mine))
        # the device-resident form of the same exchange: slots padded with records the finders skip, nothing returned to the host
        ptr, n = ex.step_device(mine)
        raw = np.frombuffer((ctypes.c_uint8 * (n * REC.itemsize)).from_address(ptr), REC).copy()
        shards = []
        for r in range(world):
            a, b = multi.shard_bounds(11, r, world)
            t = np.zeros(b - a, REC)
            t["model"], t["ret"] = 100 + np.arange(a, b), tick + 1
            t["pose"][:, 0] = np.arange(a, b) + 0.25 * tick
            shards.append(t)
        assert n == world * 6 and raw.tobytes() == ex.host_view(shards).tobytes()
        assert np.array_equal(raw["model"][raw["model"] != 0xffffffff], 100 + np.arange(11)) and int((raw["model"] == 0xffffffff).sum()) == 1
    q.put((rank, [g.tobytes() for g in got], ex.bytes_sent))
    dist.barrier()
    dist.destroy_process_group()


def test_target_exchange_over_gloo_two_ranks_uneven_shards():
    """Fiducial target records of every rank reach every rank, in rank (= model id) order, shards of unequal size."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_target_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        rank, blobs, nbytes = q.get(timeout=120)
        got[rank] = ([np.frombuffer(b, REC) for b in blobs], nbytes)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for tick in range(3):
        a, b = got[0][0][tick], got[1][0][tick]
        assert a.tobytes() == b.tobytes() and len(a) == 11
        assert np.array_equal(a["model"], 100 + np.arange(11)) and np.all(a["ret"] == tick + 1)
        assert np.array_equal(a["pose"][:, 0], np.arange(11) + 0.25 * tick)
    assert got[0][1] == 2 * 3 * 6 * 72 and got[1][1] == 2 * 3 * 5 * 72
    with pytest.raises(ValueError):
        multi.TargetExchange(0, 1, 2, REC, device="cpu").step(np.zeros(3, REC))
    one = multi.TargetExchange(0, 1, 4, REC, device="cpu").step(np.zeros(3, REC))  # a single rank gets its own back
    assert len(one) == 3
