"""The multi-rank delta path on ONE GPU: N contexts stand for N ranks. Each exports the deltas of
its own shard of movers into its slot of a gathered buffer (a device-to-device copy stands in for the
NCCL all-gather), and every context applies all N slots in one K1 pass. After several ticks every
replica's grid must equal the oracle's, which saw the same maps rank by rank - cell for cell, list
order included - and rays must be bit-exact. (tests/multi_gpu_check.py does the same over NCCL.)"""
import numpy as np
import pytest
import torch

import oracle_lib
from stage_b200 import multi, rt, worldgen

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_ranks", [3, 8])
def test_gathered_deltas_from_n_ranks(n_ranks):
    per_rank, ticks = 40, 5
    n_robots = per_rank * n_ranks
    w = worldgen.make_world(seed=21, n_rects=500, n_robots=n_robots, half_extent_m=30.72)
    x0, y0, nx, ny = w.sreg_extent()
    n_bodies = len(w.obstacles) + len(w.robots)
    stream = torch.cuda.Stream()
    ctxs = []
    for r in range(n_ranks):
        ctx = rt.Context(w.ppm, x0, y0, nx, ny, max_models=w.n_models + 1, max_blocks=n_bodies, max_cell_entries=1 << 20,
                         max_delta_bytes=1 << 17)
        ctx.set_stream(stream.cuda_stream)
        worldgen.load_into_context(ctx, w)
        ctxs.append(ctx)
    slot = int(ctxs[0].delta_slot_bytes())
    gathered = torch.empty(slot * n_ranks, dtype=torch.uint8, device="cuda")

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

    poses = w.robots.copy()
    n_obst = len(w.obstacles)
    with torch.cuda.stream(stream):
        for t in range(1, ticks + 1):
            poses = worldgen.step_robots(w, poses, t, speed_m=0.2)
            layer_w = t % 2
            rp = worldgen.robot_polygons(w, poses)
            for r, ctx in enumerate(ctxs):  # every rank re-maps its own robots and exports the deltas
                lo, hi = multi.shard_bounds(n_robots, r, n_ranks)
                for j in range(lo, hi):
                    ctx.block_unmap(n_obst + j, layer_w)
                    ctx.block_map(n_obst + j, int(w.robot_ids[j]), layer_w, 0.0, w.robot_size[2], rp[j])
                    t1.block_unmap(n_obst + j, layer_w)
                    t1.block_map(n_obst + j, int(w.robot_ids[j]), layer_w, 0.0, w.robot_size[2], rp[j])
                ptr, nbytes = ctx.delta_export(r)
                assert 0 < nbytes <= slot
                gathered[r * slot:(r + 1) * slot].copy_(multi.device_blob_as_tensor(ptr, slot))
            for ctx in ctxs:  # ... and applies everybody's
                ctx.delta_apply_gathered(gathered.data_ptr(), n_ranks, slot)
            stream.synchronize()
    rec_o, cnt_o = t1.dump_grid()
    for ctx in ctxs[:: max(1, n_ranks // 3)]:
        rec_g, cnt_g = ctx.grid_dump()
        assert np.array_equal(oracle_lib.sort_grid(rec_g), rec_o)
        assert np.array_equal(oracle_lib.sort_counts(cnt_g), cnt_o)
    digests = [ctx.grid_hash() for ctx in ctxs]
    assert all(d == digests[0] for d in digests) and digests[0] == t1.grid_hash()
    # rays of the last rank's shard, read layer = the one just written
    lo, hi = multi.shard_bounds(n_robots, n_ranks - 1, n_ranks)
    fans = worldgen.ranger_fans(w, layer=ticks % 2, samples=180, range_max=20.0, robots=poses[lo:hi], robot_ids=w.robot_ids[lo:hi])
    head = np.concatenate([oracle_lib.ranger_headings(f["oa"], f["fov"], int(f["n_samples"])) for f in fans])
    g = ctxs[-1].trace(fans, trig=oracle_lib.libm_trig(head))
    want = [t1.ranger_fan(int(f["finder"]), int(f["layer"]), [f["ox"], f["oy"], f["oz"], f["oa"]], float(f["range"]),
                          float(f["fov"]), int(f["n_samples"])) for f in fans]
    assert np.array_equal(g.hit_model, np.concatenate([x[3]["hit_model"] for x in want]))
    assert np.array_equal(g.ranges.view(np.uint64), np.concatenate([x[0] for x in want]).view(np.uint64))
    for ctx in ctxs:
        ctx.close()
