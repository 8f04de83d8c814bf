"""The pose path: stg_rt_block_define + stg_rt_models_pose_update derive Model::LocalToPixels
(model.cc:474-491) on the device. Held to (1) the unmodified reference's own LocalToPixels output for every block of seven
shipped worlds after ticks under their controllers (tests/golden/local_to_pixels.npz, made by
tests/golden/make_l2p_golden.py), bit for bit; (2) the host-rasterised path and the T1 oracle on moving swarms: same
grid cell for cell, same digest, same rays; (3) itself when calls of both kinds are mixed on one block, when a
collision test needs an outline only the device knows, and across N ranks through the delta exchange."""
import math
import os

import numpy as np
import pytest
import torch

import oracle_lib
from stage_b200 import multi, rt, worldgen

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "local_to_pixels.npz")


def test_device_local_to_pixels_equals_the_reference_on_shipped_worlds():
    data = np.load(GOLDEN)
    worlds = sorted({k.split("/")[0] for k in data.files})
    assert len(worlds) >= 6
    total = 0
    for wn in worlds:
        d = {k.split("/")[1]: data[k] for k in data.files if k.startswith(wn + "/")}
        ppm, model, first = float(d["ppm"]), d["model"], d["first"]
        pix = d["pix"]
        lo = (pix.min(0) >> 10) - 1
        hi = (pix.max(0) >> 10) + 1
        ctx = rt.Context(ppm, int(lo[0]), int(lo[1]), int(hi[0] - lo[0] + 1), int(hi[1] - lo[1] + 1), max_models=int(model.max()) + 2,
                         max_blocks=len(model) + 1, max_cell_entries=1 << 22)
        for m in np.unique(model):
            ctx.model_upsert(int(m), int(m), 0.5)
        for b in range(len(model)):
            ctx.block_define(b, int(model[b]), d["local_z"][b, 0], d["local_z"][b, 1], d["pts"][first[b]:first[b + 1]])
        # one pose per model: every block of a model shares GetGlobalPose() + geom.pose
        ids, idx = np.unique(model, return_index=True)
        for layer in (0, 1):
            ctx.models_pose_update(ids, d["gpose"][idx], layer)
        ctx.sync()
        t1 = oracle_lib.T1World(ppm)
        for m in ids:
            t1.model_upsert(int(m), int(m), 0.5)
        for b in range(len(model)):
            want = pix[first[b]:first[b + 1]]
            for layer in (0, 1):
                got = ctx.block_outline(b, layer)
                assert np.array_equal(got, want), (wn, b, layer)
            total += len(want)
        # rendered in model order, each model's blocks in BlockGroup order: the oracle fed the reference's pixels agrees
        for layer in (0, 1):
            for m in ids:
                for b in np.nonzero(model == m)[0]:
                    t1.block_map(int(b), int(m), layer, d["local_z"][b, 0] + d["gpose"][b, 2], d["local_z"][b, 1] + d["gpose"][b, 2], pix[first[b]:first[b + 1]])
        assert ctx.grid_hash() == t1.grid_hash(), wn
        ctx.close()
    assert total > 30000


def swarm(seed=3, n=300):
    w = worldgen.make_world(seed=seed, n_rects=300, n_robots=n, half_extent_m=20.48)
    return w, w.sreg_extent()


def local_box(size):
    sx, sy = size[0], size[1]
    return np.array([[-0.5 * sx, -0.5 * sy], [0.5 * sx, -0.5 * sy], [0.5 * sx, 0.5 * sy], [-0.5 * sx, 0.5 * sy]])


def gposes(robots):
    return np.stack([robots[:, 0], robots[:, 1], np.zeros(len(robots)), robots[:, 2] * (math.pi / 180.0)], 1)


def define_robots(ctx, w):
    n_ob = len(w.obstacles)
    for j, mid in enumerate(w.robot_ids):
        ctx.block_define(n_ob + j, int(mid), 0.0, w.robot_size[2], local_box(w.robot_size))


def test_moving_swarm_pose_path_equals_host_rasterised_path_and_oracle():
    w, ext = swarm()
    n_ob, n = len(w.obstacles), len(w.robots)
    mk = lambda: rt.Context(w.ppm, *ext, max_models=w.n_models + 1, max_blocks=n_ob + n, max_cell_entries=1 << 20)
    a, b = mk(), mk()
    for ctx in (a, b):
        worldgen.load_into_context(ctx, w)
    define_robots(a, w)
    t1 = oracle_lib.T1World(w.ppm)
    t1.model_upsert(0, 0, 1.0)
    for mid in list(w.obstacle_ids) + list(w.robot_ids):
        t1.model_upsert(int(mid), int(mid), 0.5)
    polys = np.concatenate([np.array(worldgen.obstacle_polygons(w)), worldgen.robot_boxes(w)], 0)
    models = np.concatenate([w.obstacle_ids, w.robot_ids])
    z = np.stack([np.zeros(len(polys)), np.concatenate([np.full(n_ob, w.obstacle_height), np.full(n, w.robot_size[2])])], 1)
    first = np.arange(len(polys) + 1, dtype=np.uint32) * 4
    for layer in (0, 1):
        t1.blocks_remap(np.arange(len(polys), dtype=np.uint32), models, layer, z, first, polys.reshape(-1, 2))
    blocks = (n_ob + np.arange(n)).astype(np.uint32)
    poses = w.robots.copy()
    rng = np.random.RandomState(1)
    for t in range(1, 9):
        poses = worldgen.step_robots(w, poses, t, speed_m=0.3)
        poses[:, 2] = np.round(rng.uniform(-180, 180, n) * 1000) / 1000.0  # every angle, not one degree per tick
        if t == 3:
            poses[:4, 2] = [0.0, 90.0, 180.0, -90.0]
        layer = t % 2
        boxes = worldgen.robot_boxes(w, poses)
        a.models_pose_update(w.robot_ids, gposes(poses), layer)
        b.blocks_remap(blocks, w.robot_ids, layer, z[n_ob:], first[:n + 1], boxes.reshape(-1, 2))
        t1.blocks_remap(blocks, w.robot_ids, layer, z[n_ob:], first[:n + 1], boxes.reshape(-1, 2))
        if t in (1, 8):
            for j in range(0, n, 7):
                assert np.array_equal(a.block_outline(n_ob + j, layer), boxes[j])
    for ctx in (a, b):
        ctx.sync()
        assert ctx.grid_hash() == t1.grid_hash()
    rec_a, cnt_a = a.grid_dump()
    rec_o, cnt_o = t1.dump_grid()
    assert np.array_equal(oracle_lib.sort_grid(rec_a), rec_o) and np.array_equal(oracle_lib.sort_counts(cnt_a), cnt_o)
    fans = worldgen.ranger_fans(w, layer=0, samples=90, range_max=12.0, robots=poses[:200], robot_ids=w.robot_ids[:200])
    ga, gb = a.trace(fans), b.trace(fans)
    assert np.array_equal(ga.ranges.view(np.uint64), gb.ranges.view(np.uint64)) and np.array_equal(ga.hit_model, gb.hit_model)
    assert np.array_equal(ga.steps, gb.steps)
    # one more tick, counted: a moved robot costs 48 bytes of deltas instead of 296
    poses = worldgen.step_robots(w, poses, 9, speed_m=0.3)
    st_a, st_b = a.stats(), b.stats()
    a.models_pose_update(w.robot_ids, gposes(poses), 1)
    b.blocks_remap(blocks, w.robot_ids, 1, z[n_ob:], first[:n + 1], worldgen.robot_boxes(w, poses).reshape(-1, 2))
    a.sync()
    b.sync()
    da, db = a.stats()["h2d_bytes"] - st_a["h2d_bytes"], b.stats()["h2d_bytes"] - st_b["h2d_bytes"]
    assert da == 96 + 48 * n and db == 96 + 296 * n
    assert a.grid_hash() == b.grid_hash()
    a.close()
    b.close()


def test_pose_and_outline_calls_mix_on_one_block():
    w, ext = swarm(seed=4, n=60)
    n_ob, n = len(w.obstacles), len(w.robots)
    ctx = rt.Context(w.ppm, *ext, max_models=w.n_models + 1, max_blocks=n_ob + n, max_cell_entries=1 << 20)
    worldgen.load_into_context(ctx, w)
    define_robots(ctx, w)
    t1 = oracle_lib.T1World(w.ppm)
    t1.model_upsert(0, 0, 1.0)
    for mid in list(w.obstacle_ids) + list(w.robot_ids):
        t1.model_upsert(int(mid), int(mid), 0.5)
    polys = worldgen.obstacle_polygons(w) + worldgen.robot_polygons(w)
    models = np.concatenate([w.obstacle_ids, w.robot_ids])
    for layer in (0, 1):
        for bi, (p, m) in enumerate(zip(polys, models)):
            t1.block_map(bi, int(m), layer, 0.0, w.obstacle_height if bi < n_ob else w.robot_size[2], p)
    zr = w.robot_size[2]
    p1 = worldgen.step_robots(w, w.robots, 1, speed_m=0.4)
    b1 = worldgen.robot_boxes(w, p1)
    # 1. host outline -> pose: the old outline leaves through edge records, the new one is derived on the device
    ctx.models_pose_update(w.robot_ids, gposes(p1), 0)
    for j in range(n):
        t1.block_unmap(n_ob + j, 0)
        t1.block_map(n_ob + j, int(w.robot_ids[j]), 0, 0.0, zr, b1[j])
    ctx.sync()
    assert ctx.grid_hash() == t1.grid_hash()
    # 2. pose -> unmap (a removal of an outline only the device knows), for every third robot
    for j in range(0, n, 3):
        ctx.block_unmap(n_ob + j, 0)
        t1.block_unmap(n_ob + j, 0)
    ctx.sync()
    assert ctx.grid_hash() == t1.grid_hash()
    # 3. pose -> host outline again (robots 1, 4, 7 ...), and an outline mapped ON TOP of a pose rendering (robots 2, 5, 8 ...:
    #    Block::Map on a mapped block renders it once more; the library has to bring the device's outline home first)
    p2 = worldgen.step_robots(w, p1, 2, speed_m=0.4)
    b2 = worldgen.robot_boxes(w, p2)
    for j in range(1, n, 3):
        ctx.block_unmap(n_ob + j, 0)
        ctx.block_map(n_ob + j, int(w.robot_ids[j]), 0, 0.0, zr, b2[j])
        t1.block_unmap(n_ob + j, 0)
        t1.block_map(n_ob + j, int(w.robot_ids[j]), 0, 0.0, zr, b2[j])
    for j in range(2, n, 3):
        ctx.block_map(n_ob + j, int(w.robot_ids[j]), 0, 0.0, zr, b2[j])
        t1.block_map(n_ob + j, int(w.robot_ids[j]), 0, 0.0, zr, b2[j])
    ctx.sync()
    rec_g, cnt_g = ctx.grid_dump()
    rec_o, cnt_o = t1.dump_grid()
    assert np.array_equal(oracle_lib.entry_multiset(rec_g), oracle_lib.entry_multiset(rec_o))
    assert np.array_equal(oracle_lib.sort_counts(cnt_g), cnt_o)
    # 4. everything back through poses, then Model::TestCollision for robots the device rendered
    p3 = worldgen.step_robots(w, p2, 3, speed_m=0.4)
    b3 = worldgen.robot_boxes(w, p3)
    ctx.models_pose_update(w.robot_ids, gposes(p3), 0)
    for j in range(n):
        t1.block_unmap(n_ob + j, 0)
        t1.block_map(n_ob + j, int(w.robot_ids[j]), 0, 0.0, zr, b3[j])
    hit = ctx.collisions_test(0, [[n_ob + j] for j in range(n)])
    ctx.sync()
    want = np.array([t1.test_collision(0, [n_ob + j]) for j in range(n)], np.int32)
    assert np.array_equal(hit, want)
    assert ctx.grid_hash() == t1.grid_hash()
    ctx.close()


def test_pose_deltas_cross_the_rank_exchange():
    n_ranks, per = 3, 50
    w, ext = swarm(seed=8, n=n_ranks * per)
    n_ob, n = len(w.obstacles), len(w.robots)
    stream = torch.cuda.Stream()
    ctxs = []
    for r in range(n_ranks):
        ctx = rt.Context(w.ppm, *ext, max_models=w.n_models + 1, max_blocks=n_ob + n, max_cell_entries=1 << 20, max_delta_bytes=1 << 16)
        ctx.set_stream(stream.cuda_stream)
        worldgen.load_into_context(ctx, w)
        define_robots(ctx, w)
        ctxs.append(ctx)
    slot = int(ctxs[0].delta_slot_bytes())
    gathered = torch.zeros(slot * n_ranks, dtype=torch.uint8, device="cuda")
    t1 = oracle_lib.T1World(w.ppm)
    t1.model_upsert(0, 0, 1.0)
    for mid in list(w.obstacle_ids) + list(w.robot_ids):
        t1.model_upsert(int(mid), int(mid), 0.5)
    polys = worldgen.obstacle_polygons(w) + worldgen.robot_polygons(w)
    models = np.concatenate([w.obstacle_ids, w.robot_ids])
    for layer in (0, 1):
        for bi, (p, m) in enumerate(zip(polys, models)):
            t1.block_map(bi, int(m), layer, 0.0, w.obstacle_height if bi < n_ob else w.robot_size[2], p)
    poses = w.robots.copy()
    sent = 0
    with torch.cuda.stream(stream):
        for t in range(1, 6):
            poses = worldgen.step_robots(w, poses, t, speed_m=0.25)
            boxes = worldgen.robot_boxes(w, poses)
            for r, ctx in enumerate(ctxs):
                lo, hi = multi.shard_bounds(n, r, n_ranks)
                ctx.models_pose_update(w.robot_ids[lo:hi], gposes(poses[lo:hi]), t % 2)
                ptr, nbytes = ctx.delta_export(r)
                sent += nbytes
                gathered[r * slot:(r + 1) * slot].copy_(multi.device_blob_as_tensor(ptr, slot))
            for ctx in ctxs:
                ctx.delta_apply_gathered(gathered.data_ptr(), n_ranks, slot)
            for j in range(n):
                t1.block_unmap(n_ob + j, t % 2)
                t1.block_map(n_ob + j, int(w.robot_ids[j]), t % 2, 0.0, w.robot_size[2], boxes[j])
        for ctx in ctxs:
            ctx.sync()
    want = t1.grid_hash()
    for ctx in ctxs:
        assert ctx.grid_hash() == want
    # the first tick of each layer replaces host-rasterised outlines (4 removal edges per robot); after that a robot costs 48 bytes
    assert sent == 2 * n * (4 * 32 + 48) + 3 * n * 48 + 5 * n_ranks * 96
    for ctx in ctxs:
        ctx.close()


def test_bulk_pose_update_on_the_host_threads_equals_the_loop():
    """A swarm's pose update (>= 8192 models of one block each, ascending ids, nothing queued) is checked, queued, sized and
    written by the library's host threads side by side; fewer models per call take the one-by-one loop. The same grid either
    way, tick after tick on alternating layers - the digest (entries, their order in every cell, region counts), outlines,
    bytes uploaded - and the same as a third context that is handed the ids in descending order (the loop again; other
    sequence numbers, so only entries and counts are compared there)."""
    w = worldgen.make_swarm_world(seed=4, n_rects=100, n_robots=9000, half_extent_m=20.48)
    ext = w.sreg_extent()
    n_ob, n = len(w.obstacles), len(w.robots)
    mk = lambda: rt.Context(w.ppm, *ext, max_models=w.n_models + 1, max_blocks=n_ob + n, max_cell_entries=1 << 21)
    a, b, c = mk(), mk(), mk()
    for ctx in (a, b, c):
        worldgen.load_swarm_into_context(ctx, w)
        worldgen.define_boxes(ctx, n_ob, w.robot_ids, w.robot_size)
    poses = w.robots.copy()
    half = n // 2
    back = np.arange(n - 1, -1, -1)
    for t in range(1, 6):
        poses = worldgen.step_robots(w, poses, t, speed_m=0.05)
        layer, gp = t % 2, worldgen.gposes(poses)
        sa, sb = a.stats(), b.stats()
        a.models_pose_update(w.robot_ids, gp, layer)
        b.models_pose_update(w.robot_ids[:half], gp[:half], layer)
        b.models_pose_update(w.robot_ids[half:], gp[half:], layer)
        c.models_pose_update(w.robot_ids[back], np.ascontiguousarray(gp[back]), layer)
        for ctx in (a, b, c):
            ctx.sync()
        if t >= 3:   # (the first Map of a layer also takes the load-time outline out through edge records)
            assert a.stats()["h2d_bytes"] - sa["h2d_bytes"] == b.stats()["h2d_bytes"] - sb["h2d_bytes"] == 96 + 48 * n
        boxes = worldgen.robot_boxes(w, poses)
        for j in (0, 1, half - 1, half, n - 1):
            assert np.array_equal(a.block_outline(n_ob + j, layer), boxes[j]) and np.array_equal(b.block_outline(n_ob + j, layer), boxes[j])
        assert a.grid_hash() == b.grid_hash()
    ha, hc = a.grid_hash(), c.grid_hash()
    assert ha[1] == hc[1] and ha[2] == hc[2]
    rec_a, cnt_a = a.grid_dump()
    rec_c, cnt_c = c.grid_dump()
    assert np.array_equal(oracle_lib.entry_multiset(rec_a), oracle_lib.entry_multiset(rec_c))
    assert np.array_equal(oracle_lib.sort_counts(cnt_a), oracle_lib.sort_counts(cnt_c))
    fans = worldgen.ranger_fans(w, layer=1, samples=64, range_max=6.0, robots=poses[:300], robot_ids=w.robot_ids[:300])
    ga, gb = a.trace(fans), b.trace(fans)
    assert np.array_equal(ga.ranges.view(np.uint64), gb.ranges.view(np.uint64)) and np.array_equal(ga.hit_model, gb.hit_model)
    for ctx in (a, b, c):
        ctx.close()
