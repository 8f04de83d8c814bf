"""Multi-GPU parity check, run under torchrun on N GPUs of one box (not a pytest file):

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py

Robots are sharded over the ranks; every tick each rank re-rasterises its own robots, the deltas are
all-gathered over NCCL and applied on every rank. Checks, against the T1 oracle fed the same maps in
rank order: (1) every rank's device grid equals the oracle's cell for cell (both layers, list order,
region counts) after several ticks; (2) each rank's shard of laser fans matches the oracle bit for
bit (strict trig); (3) every robot is a fiducial and carries a finder: the ranks all-gather the target records
(multi.TargetExchange) and each rank's finders, run against the world's targets on the replicated grid, report the
fiducials the oracle reports (targets, ids, order; ranges to 1e-9). Prints one line per rank and exits non-zero on
any mismatch."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import oracle_lib  # noqa: E402
from stage_b200 import multi, rt, worldgen  # noqa: E402


def main():
    rank, world_size = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_robots, ticks = 64 * world_size, 6
    half = 20.48 if world_size <= 4 else 40.96  # room for every rank's 64 robots
    w = worldgen.make_world(seed=9, n_rects=int(300 * (half / 20.48) ** 2), n_robots=n_robots, half_extent_m=half)
    x0, y0, nx, ny = w.sreg_extent()
    n_bodies = len(w.obstacles) + len(w.robots)
    ctx = rt.Context(w.ppm, x0, y0, nx, ny, max_models=w.n_models + 1, max_blocks=n_bodies, max_cell_entries=1 << 20,
                     device=local, max_delta_bytes=1 << 18)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    worldgen.load_into_context(ctx, w)
    ex = multi.DeltaExchange(ctx, rank, world_size, dist=dist)
    tex = multi.TargetExchange(rank, world_size, max_per_rank=(n_robots + world_size - 1) // world_size,
                               record_dtype=rt.FID_TARGET_DTYPE, dist=dist)
    L = oracle_lib.t1()
    L.t1_fiducial_update.restype = ctypes.c_uint32
    L.t1_fiducial_update.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]
    fid_ok, fid_seen = True, 0

    # the oracle sees the same calls: initial load, then per tick every rank's remaps in rank order
    t1 = oracle_lib.T1World(w.ppm)
    t1.model_upsert(0, 0, 1.0)
    for mid in list(w.obstacle_ids) + list(w.robot_ids):
        t1.model_upsert(int(mid), int(mid), 0.5)
    polys = worldgen.obstacle_polygons(w) + worldgen.robot_polygons(w)
    models = np.concatenate([w.obstacle_ids, w.robot_ids])
    zmax = [w.obstacle_height] * len(w.obstacles) + [w.robot_size[2]] * len(w.robots)
    for layer in (0, 1):
        for b, (p, m, z) in enumerate(zip(polys, models, zmax)):
            t1.block_map(b, int(m), layer, 0.0, z, p)

    lo, hi = multi.shard_bounds(n_robots, rank, world_size)
    poses = w.robots.copy()
    ok = True
    for t in range(1, ticks + 1):
        poses = worldgen.step_robots(w, poses, t, speed_m=0.2)
        layer_w = t % 2
        rp = worldgen.robot_polygons(w, poses)
        # this rank's own robots through the C ABI
        mine = rp[lo:hi]
        ctx.blocks_remap((len(w.obstacles) + np.arange(lo, hi)).astype(np.uint32), w.robot_ids[lo:hi], layer_w,
                         np.stack([np.zeros(hi - lo), np.full(hi - lo, w.robot_size[2])], 1),
                         np.arange(hi - lo + 1, dtype=np.uint32) * 4, np.concatenate(mine, 0))
        ex.step()
        # the oracle: all robots, rank by rank (= robot order, shards are contiguous)
        for j in range(n_robots):
            t1.block_unmap(len(w.obstacles) + j, layer_w)
            t1.block_map(len(w.obstacles) + j, int(w.robot_ids[j]), layer_w, 0.0, w.robot_size[2], rp[j])
        # this rank's fans on the layer that was NOT written this tick
        fans = worldgen.ranger_fans(w, layer=(t + 1) % 2, samples=180, range_max=20.0, robots=poses[lo:hi],
                                    robot_ids=w.robot_ids[lo:hi])
        head = np.concatenate([oracle_lib.ranger_headings(f["oa"], f["fov"], 180) for f in fans])
        g = ctx.trace(fans, want=rt.WANT_RANGES | rt.WANT_HIT_MODEL, trig=oracle_lib.libm_trig(head))
        for i, f in enumerate(fans):
            rg, _, _, hits = t1.ranger_fan(int(f["finder"]), int(f["layer"]), [f["ox"], f["oy"], f["oz"], f["oa"]],
                                           float(f["range"]), float(f["fov"]), 180)
            sl = slice(i * 180, (i + 1) * 180)
            if not (np.array_equal(g.ranges[sl].view(np.uint64), rg.view(np.uint64)) and
                    np.array_equal(g.hit_model[sl], hits["hit_model"])):
                ok = False
        # fiducials: this rank's robots as targets go round, its finders look at everybody's
        mine_t = np.zeros(hi - lo, rt.FID_TARGET_DTYPE)
        mine_t["model"], mine_t["fiducial_return"] = w.robot_ids[lo:hi], 1 + np.arange(lo, hi) % 7
        mine_t["x"], mine_t["y"], mine_t["a"] = poses[lo:hi, 0], poses[lo:hi, 1], poses[lo:hi, 2] * (np.pi / 180)
        mine_t["sx"], mine_t["sy"], mine_t["sz"] = w.robot_size
        targets = tex.step(mine_t)
        assert len(targets) == n_robots and np.array_equal(targets["model"], w.robot_ids)
        finders = np.zeros(hi - lo, rt.FID_FINDER_DTYPE)
        finders["finder"], finders["layer"] = w.robot_ids[lo:hi], (t + 1) % 2
        for k, col in (("x", 0), ("y", 1)):
            finders[k] = finders["o" + k] = poses[lo:hi, col]
        finders["a"] = finders["oa"] = poses[lo:hi, 2] * (np.pi / 180)
        finders["oz"], finders["max_range_anon"], finders["max_range_id"], finders["fov"] = 0.3, 8.0, 5.0, np.pi
        out_f, cnt_f = ctx.fiducials_submit(targets, finders, max_per_finder=32)
        ctx.sync()
        for i in range(hi - lo):
            ref = np.zeros(32, rt.FIDUCIAL_DTYPE)
            f1 = np.ascontiguousarray(finders[i:i + 1])
            n = L.t1_fiducial_update(t1.h, len(targets), targets.ctypes.data, f1.ctypes.data, 32, ref.ctypes.data)
            fid_seen += int(n)
            if not (n == cnt_f[i] and np.array_equal(out_f[i, :n]["target"], ref[:n]["target"]) and
                    np.array_equal(out_f[i, :n]["id"], ref[:n]["id"]) and
                    np.max(np.abs(out_f[i, :n]["range"] - ref[:n]["range"]), initial=0) <= 1e-9):
                fid_ok = False
    ok = ok and fid_ok and fid_seen > 0
    rec_g, cnt_g = ctx.grid_dump()
    rec_o, cnt_o = t1.dump_grid()
    grid_ok = np.array_equal(oracle_lib.sort_grid(rec_g), rec_o) and np.array_equal(oracle_lib.sort_counts(cnt_g), cnt_o)
    flag = torch.tensor([int(ok and grid_ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    print("rank %d/%d: rays + fiducials %s (%d fiducials seen by this rank's finders), grid %s (%d entries), exported %d delta bytes, %d target bytes" % (
        rank, world_size, "ok" if ok else "MISMATCH", fid_seen, "ok" if grid_ok else "MISMATCH", len(rec_g), ex.bytes_exported, tex.bytes_sent), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
