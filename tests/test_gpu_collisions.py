"""Model::TestCollision on the device (stg_rt_collisions_test) against the unmodified reference's
recorded answers and against the oracle.

The bumper_150 golden trace (tests/golden/make_golden.py) holds 2100 Model::TestCollision calls of
the reference, 1728 of which found something - walls, a rotated pillar, other robots, a child
model's arm - interleaved with the Map / UnMap calls of ModelPosition::Move, including the undo
re-maps after a hit. Replaying it, every query must return the model the reference returned."""
import lzma
import os
import time

import numpy as np
import pytest

import oracle_lib
import replay
import trace_io
from stage_b200 import rt, worldgen

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name, tmp_path):
    dst = tmp_path / (name + ".trc")
    with lzma.open(os.path.join(GOLDEN, name + ".trc.xz"), "rb") as f:
        dst.write_bytes(f.read())
    return trace_io.load(dst)


def test_collision_replay_equals_reference(tmp_path):
    tr = load("bumper_150", tmp_path)
    want = [tr.collision(ev)[3] for ev in tr.events if ev["type"] == trace_io.EV_COLLISION]
    g = replay.GpuReplay(tr, strict=True)
    assert len(want) == 2100 and sum(w >= 0 for w in want) > 1500
    assert g.collisions == want
    # and the grid those answers were read from equals the oracle's, cell for cell
    t1 = replay.T1Replay(tr)
    assert t1.collisions == want
    rec_g, cnt_g = g.ctx.grid_dump()
    rec_o, cnt_o = t1.world.dump_grid()
    assert np.array_equal(oracle_lib.sort_grid(rec_g), rec_o) and np.array_equal(oracle_lib.sort_counts(cnt_g), cnt_o)
    g.ctx.close()


def test_touching_models_replay_equals_reference(tmp_path):
    """charger_150 (tests/golden/make_golden.py): 600 Model::AppendTouchingModels calls of the reference - a floor
    pad and a strip that robots drive across, a dock built into a wall, a post standing in two other models -
    interleaved with the robots' Map / UnMap and their 2100 collision tests. Same sets, call by call."""
    tr = load("charger_150", tmp_path)
    want = [[int(x) for x in tr.touchers(ev)[3]] for ev in tr.events if ev["type"] == trace_io.EV_TOUCHERS]
    want_col = [tr.collision(ev)[3] for ev in tr.events if ev["type"] == trace_io.EV_COLLISION]
    assert len(want) == 600 and sum(len(w) for w in want) > 700 and max(len(w) for w in want) >= 4
    g = replay.GpuReplay(tr, strict=True)
    assert g.touchers == want
    assert g.collisions == want_col
    t1 = replay.T1Replay(tr)
    assert t1.touchers == want
    g.ctx.close()


def test_collision_replay_no_false_positives_on_shipped_worlds(tmp_path):
    """simple.world: 238 tests, none hits (the wander controller keeps clear) - the device must agree"""
    tr = load("simple_120", tmp_path)
    want = [tr.collision(ev)[3] for ev in tr.events if ev["type"] == trace_io.EV_COLLISION]
    g = replay.GpuReplay(tr, strict=True, want=rt.WANT_RANGES)
    assert len(want) > 200 and g.collisions == want
    g.ctx.close()


def crowded_world(n_robots, seed):
    """config-3 style obstacle field with robots dropped anywhere, bodies of 1..3 blocks: many overlap
    obstacles or each other; some models are no obstacles, some float above the rest, one reaches below the floor"""
    w = worldgen.make_world(seed=seed, n_rects=4096, n_robots=1, half_extent_m=81.92)
    rng = np.random.RandomState(seed + 100)
    xy = rng.uniform(-80, 80, (n_robots, 2))
    heading = rng.uniform(-180, 180, n_robots)
    return w, xy, heading, rng


def test_batched_collisions_equal_oracle_at_scale():
    n = 20000
    w, xy, heading, rng = crowded_world(n, seed=4)
    x0, y0, nx, ny = w.sreg_extent()
    n_obst = len(w.obstacles)
    first_model = w.n_models + 1
    # robot i: parent model first_model + 2i (1 or 2 blocks), child model first_model + 2i + 1 (1 block, every 3rd robot)
    ctx = rt.Context(w.ppm, x0, y0, nx, ny, max_models=first_model + 2 * n + 1, max_blocks=n_obst + 1 + 3 * n, max_cell_entries=1 << 23)
    worldgen.load_into_context(ctx, w)
    t1 = oracle_lib.T1World(w.ppm)
    t1.model_upsert(0, 0, 1.0, flags=rt.VIS_OBSTACLE_RETURN)
    for mid in list(w.obstacle_ids) + list(w.robot_ids):
        t1.model_upsert(int(mid), int(mid), 0.5, flags=rt.VIS_OBSTACLE_RETURN)
    polys = worldgen.obstacle_polygons(w) + worldgen.robot_polygons(w)
    for b, (p, m) in enumerate(zip(polys, np.concatenate([w.obstacle_ids, w.robot_ids]))):
        for layer in (0, 1):
            t1.block_map(b, int(m), layer, 0.0, w.obstacle_height if b < n_obst else w.robot_size[2], p)
    layer = 1
    queries = []
    blk = n_obst + 1
    for i in range(n):
        parent, child = first_model + 2 * i, first_model + 2 * i + 1
        flags = 0 if i % 17 == 0 else rt.VIS_OBSTACLE_RETURN          # a few are no obstacles
        z0 = 2.0 if i % 13 == 0 else (-0.1 if i == 777 else 0.0)       # a few float above everything; one below the floor
        for mid, root, fl in ((parent, parent, flags), (child, parent, rt.VIS_OBSTACLE_RETURN)):
            ctx.model_upsert(mid, root, 0.5, 0, 0, fl)
            t1.model_upsert(mid, root, 0.5, flags=fl)
        a = worldgen.heading_rad(round(heading[i], 3))
        q = []
        bodies = [(parent, 0.0, 0.0, 0.44, 0.38)]
        if i % 2 == 0:
            bodies.append((parent, 0.1, 0.25, 0.2, 0.2))               # second block of the same model
        if i % 3 == 0:
            bodies.append((child, 0.35, 0.0, 0.3, 0.08))               # the child's block: related, never hits its parent
        for mid, ox, oy, sx, sy in bodies:
            cx = round(xy[i, 0] + ox * np.cos(a) - oy * np.sin(a), 3)
            cy = round(xy[i, 1] + ox * np.sin(a) + oy * np.cos(a), 3)
            p = worldgen.pixel_polygon(cx, cy, a, sx, sy, w.ppm)
            ctx.block_map(blk, mid, layer, z0, z0 + 0.3, p)
            t1.block_map(blk, mid, layer, z0, z0 + 0.3, p)
            q.append(blk)
            blk += 1
        queries.append(q)
    ctx.sync()
    first = np.zeros(n + 1, np.uint32)
    first[1:] = np.cumsum([len(q) for q in queries])
    flat = np.concatenate([np.asarray(q, np.uint32) for q in queries])
    warm = ctx.collisions_test(layer, (first, flat))  # sizes the staging buffers
    ctx.sync()
    t0 = time.perf_counter()
    hit = ctx.collisions_test(layer, (first, flat))
    t_call = time.perf_counter() - t0
    ctx.sync()
    dt = time.perf_counter() - t0
    assert np.array_equal(warm, hit)
    t0 = time.perf_counter()
    want = np.array([t1.test_collision(layer, q) for q in queries], np.int32)
    t_cpu = time.perf_counter() - t0
    bad = np.nonzero(hit != want)[0]
    assert len(bad) == 0, "query %d: device %d, oracle %d" % (bad[0], hit[bad[0]], want[bad[0]])
    assert want[777] == 0                                              # the ground model
    idx = np.arange(n)
    assert np.all(want[(idx % 17 == 0) & (idx % 3 != 0)] == -1)       # no obstacle_return, no child: tests nothing
    assert np.any(want[(idx % 17 == 0) & (idx % 3 == 0)] >= 0)        # ... but its child's block still does
    assert (want >= 0).mean() > 0.15 and (want == -1).mean() > 0.15
    floating = want[(idx % 13 == 0) & (idx != 777)]                   # above everything else: can only meet one another
    assert np.all((floating == -1) | (((floating - first_model) // 2) % 13 == 0)) and np.any(floating == -1)
    # Model::AppendTouchingModels for the same crowd: each parent model's own blocks (not the child's), every
    # unrelated model met - obstacle or not, whatever its height
    own = [q[:2] if i % 2 == 0 else q[:1] for i, q in enumerate(queries)]
    ids, counts = ctx.touching_models(layer, own, max_per_query=24)
    ctx.sync()
    n_sets = 0
    for i in range(0, n, 3):
        w_ids = t1.touching_models(layer, own[i])
        assert counts[i] == len(w_ids) and np.array_equal(ids[i, :counts[i]], w_ids), (i, ids[i, :counts[i]], w_ids)
        n_sets += len(w_ids) > 1
    assert n_sets > 100 and counts.max() <= 24
    print("collision tests: %d models, %d hit something; device %.2f ms end to end (the call itself %.2f ms), oracle (1 thread, python loop) %.0f ms"
          % (n, int((want >= 0).sum()), 1e3 * dt, 1e3 * t_call, 1e3 * t_cpu))
    ctx.close()
