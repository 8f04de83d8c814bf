"""ModelPosition::Move for all movers of a tick at once (stg_rt_models_move) against the unmodified
reference's one-by-one loop, replayed from the bumper_150 golden trace: per tick the reference
re-mapped each of the 14 robots (some with a child model), tested it, and put it back if it hit a
wall, the pillar or another robot - 2100 moves, 1728 of them undone. The batched call must stall
exactly the robots the reference stalled, every tick, and leave the grid - cell lists in order -
exactly as the reference's sequence of Map / UnMap calls leaves it."""
import lzma
import os
import time

import numpy as np
import pytest

import oracle_lib
import replay
import trace_io
from stage_b200 import rt

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_batched_move_equals_the_references_loop(tmp_path):
    dst = tmp_path / "bumper.trc"
    with lzma.open(os.path.join(GOLDEN, "bumper_150.trc.xz"), "rb") as f:
        dst.write_bytes(f.read())
    tr = trace_io.load(dst)
    ctx = replay.make_context(tr)
    t1 = replay.make_t1(tr)
    events = list(tr.events)

    def apply(ev, to_ctx=True):
        if ev["type"] == trace_io.EV_MAP:
            args = (int(ev["block"]), int(ev["model"]), int(ev["layer"]), float(ev["zmin"]), float(ev["zmax"]), tr.polygon(ev))
            if to_ctx:
                ctx.block_map(*args)
            t1.block_map(*args)
        elif ev["type"] == trace_io.EV_UNMAP:
            if to_ctx:
                ctx.block_unmap(int(ev["block"]), int(ev["layer"]))
            t1.block_unmap(int(ev["block"]), int(ev["layer"]))

    start = 0
    present = {}  # block -> (zmin, zmax, polygon) of its latest Map on either layer: the outline of the pose it has now
    n_moves = n_stalled = n_ticks = 0
    t_move = 0.0
    i = start
    while i < len(events):
        # one tick: up to the next TICK event
        j = i
        while j < len(events) and events[j]["type"] != trace_io.EV_TICK:
            j += 1
        tick = events[i:j]
        # a Move shows up as [UNMAP x n] [MAP x n] COLLISION and, if it hit, [UNMAP x n] [MAP x n] again (the undo)
        movers, want, claimed = [], [], set()
        layer = None
        for k, ev in enumerate(tick):
            if ev["type"] != trace_io.EV_COLLISION:
                continue
            _, layer, blocks, hit = tr.collision(ev)
            mine = set(int(b) for b in blocks)
            new, m = {}, k - 1
            while m >= 0 and len(new) < len(mine) and tick[m]["type"] == trace_io.EV_MAP and int(tick[m]["block"]) in mine:
                new[int(tick[m]["block"])] = (float(tick[m]["zmin"]), float(tick[m]["zmax"]), np.array(tr.polygon(tick[m])))
                claimed.add(m)
                m -= 1
            assert len(new) == len(mine)
            while m >= 0 and tick[m]["type"] == trace_io.EV_UNMAP and int(tick[m]["block"]) in mine and m not in claimed:
                claimed.add(m)
                m -= 1
            movers.append([(int(b), new[int(b)]) for b in blocks])
            want.append(1 if hit >= 0 else 0)
            if hit >= 0:  # the undo that follows is the expected RESULT, not an input
                m = k + 1
                while m < len(tick) and tick[m]["type"] in (trace_io.EV_UNMAP, trace_io.EV_MAP) and int(tick[m]["block"]) in mine and m - k <= 2 * len(mine):
                    claimed.add(m)
                    m += 1
        for k, ev in enumerate(tick):  # whatever no Move accounts for (the world load, in this trace) goes in as it came
            if k not in claimed and ev["type"] in (trace_io.EV_MAP, trace_io.EV_UNMAP):
                assert not movers or k < min(claimed), "a foreign Map / UnMap in the middle of the movers"
                if ev["type"] == trace_io.EV_MAP:
                    ctx.block_map(int(ev["block"]), int(ev["model"]), int(ev["layer"]), float(ev["zmin"]), float(ev["zmax"]), tr.polygon(ev))
                    present[int(ev["block"])] = (float(ev["zmin"]), float(ev["zmax"]), np.array(tr.polygon(ev)))
                else:
                    ctx.block_unmap(int(ev["block"]), int(ev["layer"]))
        movers = [[(b, new, present[b]) for b, new in m] for m in movers]
        for ev in tick:   # the oracle replays the reference's own sequence, call by call
            apply(ev, to_ctx=False)
            if ev["type"] == trace_io.EV_MAP:
                present[int(ev["block"])] = (float(ev["zmin"]), float(ev["zmax"]), np.array(tr.polygon(ev)))
        if movers:
            t0 = time.perf_counter()
            got = ctx.models_move(layer, movers)
            t_move += time.perf_counter() - t0
            assert list(got) == want, "tick %d: stalled %s, reference %s" % (n_ticks, list(got), want)
            n_moves += len(movers)
            n_stalled += int(sum(want))
        n_ticks += 1
        if n_ticks % 25 == 0 or j >= len(events) - 1:
            rec_g, cnt_g = ctx.grid_dump()
            rec_o, cnt_o = t1.dump_grid()
            assert np.array_equal(oracle_lib.sort_grid(rec_g), rec_o), "grid differs after tick %d" % n_ticks
            assert np.array_equal(oracle_lib.sort_counts(cnt_g), cnt_o)
        i = j + 1
    print("models_move: %d calls, %.2f ms each (14 movers)" % (n_ticks, 1e3 * t_move / max(n_ticks, 1)))
    assert n_moves == 2100 and n_stalled == 1728
    ctx.close()


def test_batched_move_equals_sequential_oracle_at_scale():
    """6000 robots in the config-3 obstacle field, dense enough that many run into obstacles and into each other, some
    in chains (a robot stopped by a robot that was itself stopped). Several ticks on alternating layers; every tick the
    batched call must stall the robots the oracle's one-by-one loop stalls (unmap, map, test, undo) and leave the same grid."""
    from stage_b200 import worldgen
    n, ticks = 6000, 4
    w = worldgen.make_world(seed=2, n_rects=2500, n_robots=1, half_extent_m=61.44)
    rng = np.random.RandomState(8)
    x0, y0, nx, ny = w.sreg_extent()
    n_obst = len(w.obstacles)
    first_model = w.n_models + 1
    ctx = rt.Context(w.ppm, x0, y0, nx, ny, max_models=first_model + n + 1, max_blocks=n_obst + 1 + n, max_cell_entries=1 << 22)
    worldgen.load_into_context(ctx, w)
    t1 = oracle_lib.T1World(w.ppm)
    t1.model_upsert(0, 0, 1.0)
    for mid in list(w.obstacle_ids) + list(w.robot_ids):
        t1.model_upsert(int(mid), int(mid), 0.5)
    polys = worldgen.obstacle_polygons(w) + worldgen.robot_polygons(w)
    for b, (p, m) in enumerate(zip(polys, np.concatenate([w.obstacle_ids, w.robot_ids]))):
        for layer in (0, 1):
            t1.block_map(b, int(m), layer, 0.0, 1.0, p)
    # robots on a jittered lattice 0.8 m apart (bodies 0.44 x 0.38): neighbours are within reach of one step
    side = int(np.ceil(np.sqrt(n)))
    gx, gy = np.meshgrid(np.arange(side), np.arange(side))
    pose = np.zeros((n, 3))
    pose[:, 0] = (gx.ravel()[:n] - side / 2) * 0.8 + rng.uniform(-0.1, 0.1, n)
    pose[:, 1] = (gy.ravel()[:n] - side / 2) * 0.8 + rng.uniform(-0.1, 0.1, n)
    pose[:, 2] = rng.uniform(-180, 180, n)
    pose = np.round(pose, 3)
    blk = n_obst + 1 + np.arange(n)
    mid = first_model + np.arange(n)

    def outline(p):
        return worldgen.pixel_polygon(p[0], p[1], worldgen.heading_rad(p[2]), 0.44, 0.38, w.ppm)
    for i in range(n):
        ctx.model_upsert(int(mid[i]), int(mid[i]), 0.5)
        t1.model_upsert(int(mid[i]), int(mid[i]), 0.5)
        for layer in (0, 1):
            ctx.block_map(int(blk[i]), int(mid[i]), layer, 0.0, 0.3, outline(pose[i]))
            t1.block_map(int(blk[i]), int(mid[i]), layer, 0.0, 0.3, outline(pose[i]))
    ctx.sync()
    total_stalled = chains = 0
    for t in range(1, ticks + 1):
        layer = t % 2
        a = np.deg2rad(pose[:, 2])
        want_pose = pose.copy()
        want_pose[:, 0] = np.round(pose[:, 0] + 0.25 * np.cos(a), 3)
        want_pose[:, 1] = np.round(pose[:, 1] + 0.25 * np.sin(a), 3)
        want_pose[:, 2] = np.round(((pose[:, 2] + 7.0 + 180.0) % 360.0) - 180.0, 3)
        movers = [[(int(blk[i]), (0.0, 0.3, outline(want_pose[i])), (0.0, 0.3, outline(pose[i])))] for i in range(n)]
        t0 = time.perf_counter()
        got = ctx.models_move(layer, movers)
        t_dev = time.perf_counter() - t0
        # the reference's loop, on the oracle
        t0 = time.perf_counter()
        want = np.zeros(n, np.uint8)
        hit_model = np.full(n, -1)
        for i in range(n):
            t1.block_unmap(int(blk[i]), layer)
            t1.block_map(int(blk[i]), int(mid[i]), layer, 0.0, 0.3, movers[i][0][1][2])
            hit_model[i] = t1.test_collision(layer, [int(blk[i])])
            if hit_model[i] >= 0:
                want[i] = 1
                t1.block_unmap(int(blk[i]), layer)
                t1.block_map(int(blk[i]), int(mid[i]), layer, 0.0, 0.3, movers[i][0][2][2])
        t_cpu = time.perf_counter() - t0
        bad = np.nonzero(got != want)[0]
        assert len(bad) == 0, "tick %d: %d movers differ, first %d (device %d, oracle %d, oracle hit model %d)" % (
            t, len(bad), bad[0], got[bad[0]], want[bad[0]], hit_model[bad[0]])
        rec_g, cnt_g = ctx.grid_dump()
        rec_o, cnt_o = t1.dump_grid()
        assert np.array_equal(oracle_lib.sort_grid(rec_g), rec_o) and np.array_equal(oracle_lib.sort_counts(cnt_g), cnt_o)
        by_robot = hit_model >= first_model
        chains += int(np.sum(by_robot & (want[np.clip(hit_model - first_model, 0, n - 1)] == 1)))
        total_stalled += int(want.sum())
        pose = np.where(want[:, None] == 1, pose, want_pose)
        print("tick %d: %d of %d stalled (%d by another robot); device %.1f ms incl. marshalling, oracle loop %.0f ms" % (
            t, int(want.sum()), n, int(by_robot.sum()), 1e3 * t_dev, 1e3 * t_cpu))
    assert total_stalled > n // 4 and chains > 20
    ctx.close()
