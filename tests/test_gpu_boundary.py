"""The C ABI under the conditions SURVEY.md 8(b) names: libstage calls submit / map / unmap from worker threads
concurrently (world.cc:228-262), and the multi-GPU delta slots come from other processes. Concurrent calls must
give the serial result; a slot that does not hold a well-formed blob, or a cell-entry table that runs full on the
device, must be reported by the next stg_rt_sync instead of corrupting the grid quietly."""
import threading

import numpy as np
import pytest
import torch

import oracle_lib
from stage_b200 import multi, rt, worldgen

pytestmark = pytest.mark.gpu


def small_world(seed=5, n_robots=64):
    w = worldgen.make_world(seed=seed, n_rects=300, n_robots=n_robots, half_extent_m=20.48)
    x0, y0, nx, ny = w.sreg_extent()
    return w, (x0, y0, nx, ny)


def test_eight_threads_submitting_and_mapping_equal_the_serial_run():
    w, ext = small_world()
    n_bodies = len(w.obstacles) + len(w.robots) + 64
    n_threads, rounds = 8, 6
    per = len(w.robots) // n_threads

    def fans_of(t, k):
        lo, hi = t * per, (t + 1) * per
        poses = w.robots[lo:hi].copy()
        poses[:, 2] += 7.0 * k
        return worldgen.ranger_fans(w, layer=1, samples=180 + t, range_max=8.0, robots=poses, robot_ids=w.robot_ids[lo:hi])

    # extra blocks live in the far corner, out of reach of every 8 m ray: the rays' answers do not depend on when a
    # thread's maps land, the final grid does not depend on their order
    corner = int((w.half_extent_m - 0.4) * w.ppm)

    def square(i):
        x = corner - 12 * (i % 8)
        y = corner - 12 * (i // 8)
        return np.array([[x, y], [x + 5, y], [x + 5, y + 5], [x, y + 5]], np.int32)

    def run(threaded):
        ctx = rt.Context(w.ppm, *ext, max_models=w.n_models + 1, max_blocks=n_bodies, max_cell_entries=1 << 19)
        worldgen.load_into_context(ctx, w)
        outs = {}
        errors = []

        def work(t):
            try:
                for k in range(rounds):
                    fans = fans_of(t, k)
                    out = rt.Outputs(int(fans["n_samples"].sum()), rt.WANT_RANGES | rt.WANT_HIT_MODEL | rt.WANT_INTENSITIES)
                    ctx.fans_submit(fans, out)
                    outs[(t, k)] = (out, fans)
                    b = len(w.obstacles) + len(w.robots) + t * 8 + k
                    ctx.block_map(b, int(w.obstacle_ids[0]), k % 2, 0.0, 1.0, square(t * 8 + k))
                    if k % 3 == 2:
                        ctx.block_unmap(b - 1, (k - 1) % 2)
            except Exception as e:  # pragma: no cover
                errors.append(e)

        if threaded:
            th = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
            for x in th:
                x.start()
            for x in th:
                x.join()
        else:
            for t in range(n_threads):
                work(t)
        assert not errors, errors
        ctx.sync()
        rec, cnt = ctx.grid_dump()
        res = {k: (v[0].ranges.copy(), v[0].hit_model.copy(), v[0].intensities.copy()) for k, v in outs.items()}
        ctx.close()
        # list order inside a cell follows Map-call order, which threads interleave: compare the entry sets
        return res, oracle_lib.sort_grid(rec)[:, [0, 1, 2, 4]], oracle_lib.sort_counts(cnt)

    serial, grid_s, cnt_s = run(False)
    threaded, grid_t, cnt_t = run(True)
    assert serial.keys() == threaded.keys() and len(serial) == n_threads * rounds
    for k in serial:
        assert np.array_equal(serial[k][0].view(np.uint64), threaded[k][0].view(np.uint64)), k
        assert np.array_equal(serial[k][1], threaded[k][1]) and np.array_equal(serial[k][2], threaded[k][2])
    assert (np.concatenate([v[1] for v in serial.values()]) >= 0).mean() > 0.3
    key = lambda g: g[np.lexsort((g[:, 3], g[:, 2], g[:, 1], g[:, 0]))]
    assert np.array_equal(key(grid_s), key(grid_t)) and np.array_equal(cnt_s, cnt_t)


def _two_rank_setup(entries=1 << 19, entries_b=None):
    w, ext = small_world(seed=9, n_robots=40)
    n_bodies = len(w.obstacles) + len(w.robots)
    stream = torch.cuda.Stream()
    ctxs = []
    for r in range(2):
        ctx = rt.Context(w.ppm, *ext, max_models=w.n_models + 1, max_blocks=n_bodies,
                         max_cell_entries=entries if (r == 0 or entries_b is None) else entries_b, max_delta_bytes=1 << 20)
        ctx.set_stream(stream.cuda_stream)
        ctxs.append(ctx)
    return w, ctxs, stream


def test_a_slot_that_holds_no_blob_is_skipped_and_reported():
    w, ctxs, stream = _two_rank_setup()
    for ctx in ctxs:
        worldgen.load_into_context(ctx, w)
    slot = int(ctxs[0].delta_slot_bytes())
    gathered = torch.zeros(slot * 2, dtype=torch.uint8, device="cuda")
    n_obst = len(w.obstacles)
    poses = worldgen.step_robots(w, w.robots, 1, speed_m=0.3)
    rp = worldgen.robot_polygons(w, poses)
    with torch.cuda.stream(stream):
        before, cnt_before = ctxs[1].grid_dump()
        for r, ctx in enumerate(ctxs):
            lo, hi = multi.shard_bounds(len(w.robots), r, 2)
            for j in range(lo, hi):
                ctx.block_unmap(n_obst + j, 1)
                ctx.block_map(n_obst + j, int(w.robot_ids[j]), 1, 0.0, w.robot_size[2], rp[j])
            ptr, nbytes = ctx.delta_export(r)
            gathered[r * slot:(r + 1) * slot].copy_(multi.device_blob_as_tensor(ptr, slot))
        # rank 0's slot as a gather that ran before the export landed would leave it: counts without their records
        torn = gathered.clone()
        torn[64:slot].zero_()
        torn[0:4] = torch.tensor([0xff, 0xff, 0xff, 0x7f], dtype=torch.uint8, device="cuda")  # n_remove = 2^31 - 1
        ctxs[1].delta_apply_gathered(torn.data_ptr(), 2, slot)
        with pytest.raises(rt.StgRtError) as e:
            ctxs[1].sync()
        assert e.value.code == -1 and "well-formed" in str(e.value)
        # nothing of the torn slot was applied; rank 1's own (sound) slot was
        after, _ = ctxs[1].grid_dump()
        lo, hi = multi.shard_bounds(len(w.robots), 0, 2)
        blocks0 = set(range(n_obst + lo, n_obst + hi))
        sel = lambda g: g[np.isin(g[:, 4], list(blocks0))]
        assert np.array_equal(oracle_lib.sort_grid(sel(before)), oracle_lib.sort_grid(sel(after)))
        assert not np.array_equal(oracle_lib.sort_grid(before), oracle_lib.sort_grid(after))
        # a buffer that never was a blob at all (no magic word)
        junk = torch.randint(0, 255, (slot * 2,), dtype=torch.uint8, device="cuda")
        ctxs[0].delta_apply_gathered(junk.data_ptr(), 2, slot)
        with pytest.raises(rt.StgRtError):
            ctxs[0].sync()
    for ctx in ctxs:
        ctx.close()


def test_a_full_cell_entry_table_on_the_device_is_reported():
    """Rank 1 has a cell-entry table far too small for what rank 0 sends: the host cannot know (the other rank's
    header is booked one exchange late), the device notices, drops the entries without touching counts or masks, and
    the next sync says so."""
    w, ctxs, stream = _two_rank_setup(entries=1 << 19, entries_b=16)
    big, tiny = ctxs
    slot = int(big.delta_slot_bytes())
    gathered = torch.zeros(slot * 2, dtype=torch.uint8, device="cuda")
    with torch.cuda.stream(stream):
        big.model_upsert(0, 0, 1.0)
        tiny.model_upsert(0, 0, 1.0)
        big.model_upsert(1, 1, 0.5)
        tiny.model_upsert(1, 1, 0.5)
        for b in range(8):  # 8 squares of 4 x 60 cell visits: 1920 entries against 64 slots
            x = -400 + 70 * b
            big.block_map(b, 1, 0, 0.0, 1.0, np.array([[x, 0], [x + 60, 0], [x + 60, 60], [x, 60]], np.int32))
        for r, ctx in enumerate(ctxs):
            ptr, nbytes = ctx.delta_export(r)
            gathered[r * slot:(r + 1) * slot].copy_(multi.device_blob_as_tensor(ptr, slot))
        tiny.delta_apply_gathered(gathered.data_ptr(), 2, slot)
        with pytest.raises(rt.StgRtError) as e:
            tiny.sync()
        assert e.value.code == -3 and "cell-entry" in str(e.value)
        rec, cnt = tiny.grid_dump()
        assert len(rec) <= 64 and int(cnt[:, 2].sum()) == len(rec)  # what was counted is what was stored
        big.delta_apply_gathered(gathered.data_ptr(), 2, slot)
        big.sync()
        assert len(big.grid_dump()[0]) == 8 * 240
    for ctx in ctxs:
        ctx.close()


def test_the_local_path_refuses_before_it_touches_anything():
    """STG_RT_ENOMEM from a local flush leaves the shadow and the device as they were: the caller can unmap and go on,
    and the grid is then exactly what the calls that went through describe."""
    ctx = rt.Context(50.0, -1, -1, 2, 2, max_models=8, max_blocks=64, max_cell_entries=256)
    t1 = oracle_lib.T1World(50.0)
    for c in (ctx, t1):
        c.model_upsert(0, 0, 1.0)
        c.model_upsert(1, 1, 0.5)
    sq = lambda x, s: np.array([[x, 0], [x + s, 0], [x + s, s], [x, s]], np.int32)
    ctx.block_map(0, 1, 0, 0.0, 1.0, sq(0, 20))     # 80 entries
    t1.block_map(0, 1, 0, 0.0, 1.0, sq(0, 20))
    ctx.sync()
    for b in range(1, 5):
        ctx.block_map(b, 1, 0, 0.0, 1.0, sq(100 * b - 450, 30))  # 4 x 120 more: over 512 / 2
    with pytest.raises(rt.StgRtError) as e:
        ctx.sync()
    assert e.value.code == -3
    for b in range(2, 5):
        ctx.block_unmap(b, 0)
    t1.block_map(1, 1, 0, 0.0, 1.0, sq(100 - 450, 30))
    ctx.sync()   # 80 + 120 fits
    rec_g, cnt_g = ctx.grid_dump()
    rec_o, cnt_o = t1.dump_grid()
    assert len(rec_g) == 200
    assert np.array_equal(oracle_lib.sort_grid(rec_g), rec_o) and np.array_equal(oracle_lib.sort_counts(cnt_g), cnt_o)
    ctx.close()
