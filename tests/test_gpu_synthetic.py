"""Parity on synthetic worlds of the benchmark's kind (many superregions, long rays, empty-region
jumps, heavily overlapping obstacles) against the T1 oracle, and size-independent properties at the
full BASELINE size."""
import numpy as np
import pytest

import oracle_lib
from stage_b200 import rt, worldgen

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def build_both(w):
    x0, y0, nx, ny = w.sreg_extent()
    ctx = rt.Context(w.ppm, x0, y0, nx, ny, max_models=w.n_models + 1, max_blocks=len(w.obstacles) + len(w.robots),
                     max_cell_entries=1 << 22)
    worldgen.load_into_context(ctx, w)
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
    return ctx, t1


def t1_fans(t1, fans):
    out = []
    for f in fans:
        out.append(t1.ranger_fan(int(f["finder"]), int(f["layer"]), [f["ox"], f["oy"], f["oz"], f["oa"]],
                                 float(f["range"]), float(f["fov"]), int(f["n_samples"])))
    rg = np.concatenate([o[0] for o in out])
    it = np.concatenate([o[1] for o in out])
    hits = np.concatenate([o[3] for o in out])
    return rg, it, hits


@pytest.mark.parametrize("seed,half", [(1, 40.96), (5, 20.48)])
def test_synthetic_world_matches_oracle(seed, half):
    w = worldgen.make_world(seed=seed, n_rects=int(1024 * (half / 40.96) ** 2), n_robots=120, half_extent_m=half)
    ctx, t1 = build_both(w)
    rec_g, cnt_g = ctx.grid_dump()
    rec_o, cnt_o = t1.dump_grid()
    assert np.array_equal(oracle_lib.sort_grid(rec_g), rec_o)
    assert np.array_equal(oracle_lib.sort_counts(cnt_g), cnt_o)
    fans = worldgen.ranger_fans(w, layer=1, samples=360, range_max=30.0)
    rg, it, hits = t1_fans(t1, fans)
    # strict: host trig per beam -> bit-exact
    head = np.concatenate([oracle_lib.ranger_headings(f["oa"], f["fov"], int(f["n_samples"])) for f in fans])
    g = ctx.trace(fans, trig=oracle_lib.libm_trig(head))
    bad = np.nonzero((g.hit_model != hits["hit_model"]) | (bits(g.ranges) != bits(rg)))[0]
    if len(bad):
        i = bad[0]
        print("first mismatch beam", i, "gpu", g.ranges[i], g.hit_model[i], g.hit_cell[i], g.steps[i], "oracle", hits[i])
    assert len(bad) == 0
    assert np.array_equal(g.steps[:, 0], hits["cell_visits"]) and np.array_equal(g.steps[:, 1], hits["region_probes"])
    assert np.array_equal(bits(g.intensities), bits(it))
    # device trig = the host libm's values bit for bit (rt_libm.cuh): identical again
    d = ctx.trace(fans)
    assert np.array_equal(bits(d.ranges), bits(rg))
    assert np.array_equal(d.hit_model, hits["hit_model"])
    h = d.hit_model >= 0
    assert np.array_equal(d.hit_cell[h, 0], hits["hit_x"][h]) and np.array_equal(d.hit_cell[h, 1], hits["hit_y"][h])
    assert np.max(np.abs(d.ranges - rg)) <= 1e-5
    # a resident set gives the same answers as a submit
    s = ctx.set_create(fans, want=rt.WANT_RANGES | rt.WANT_HIT_MODEL)
    s.trace()
    r = s.download()
    assert np.array_equal(bits(r.ranges), bits(d.ranges)) and np.array_equal(r.hit_model, d.hit_model)
    s.destroy()
    ctx.close()


def test_full_baseline_size_matches_oracle():
    """BASELINE config 3 at full size (8192^2 cells, 4096 rectangles, 1000 robots x 1080 beams): every
    beam against the T1 oracle in strict-trig mode, including the step counters."""
    w = worldgen.make_world(seed=1, n_rects=4096, n_robots=1000, half_extent_m=81.92)
    ctx, t1 = build_both(w)
    fans = worldgen.ranger_fans(w, layer=1, samples=1080, range_max=30.0)
    rg, it, hits = t1_fans(t1, fans)
    head = np.concatenate([oracle_lib.ranger_headings(f["oa"], f["fov"], int(f["n_samples"])) for f in fans])
    g = ctx.trace(fans, trig=oracle_lib.libm_trig(head))
    bad = np.nonzero((g.hit_model != hits["hit_model"]) | (bits(g.ranges) != bits(rg)) |
                     (g.steps[:, 0] != hits["cell_visits"]) | (g.steps[:, 1] != hits["region_probes"]))[0]
    for i in bad[:5]:
        f = fans[i // 1080]
        print("mismatch beam", i, "t", i % 1080, "origin", f["ox"], f["oy"], "heading %.17g" % head[i],
              "gpu", g.ranges[i], g.hit_model[i], g.hit_cell[i], g.steps[i], "oracle", hits[i])
    assert len(bad) == 0, "%d of %d beams differ" % (len(bad), len(rg))
    assert np.array_equal(bits(g.intensities), bits(it))
    ctx.close()


def test_resident_set_follows_moved_origins():
    """stg_rt_set_update_origins: a resident fan set re-traced after its robots moved (and after the
    moved bodies were re-rasterised on the other layer) equals a fresh submit of the moved fans."""
    w = worldgen.make_world(seed=4, n_rects=200, n_robots=80, half_extent_m=20.48)
    ctx, t1 = build_both(w)
    fans = worldgen.ranger_fans(w, layer=1, samples=90, range_max=12.0)
    s = ctx.set_create(fans, want=rt.WANT_RANGES | rt.WANT_HIT_MODEL)
    s.trace()
    a = s.download()
    poses = worldgen.step_robots(w, w.robots, 1, speed_m=0.3)
    polys = worldgen.robot_polygons(w, poses)
    n_o = len(w.obstacles)
    ctx.blocks_remap((n_o + np.arange(len(poses))).astype(np.uint32), w.robot_ids, 0,
                     np.tile([0.0, w.robot_size[2]], (len(poses), 1)), np.arange(len(poses) + 1, dtype=np.uint32) * 4,
                     np.concatenate(polys, 0))
    moved = worldgen.ranger_fans(w, layer=0, samples=90, range_max=12.0, robots=poses)
    s.update_origins(np.stack([moved["ox"], moved["oy"], moved["oz"], moved["oa"]], 1), 0)
    s.trace()
    b = s.download()
    c = ctx.trace(moved, want=rt.WANT_RANGES | rt.WANT_HIT_MODEL)
    assert np.array_equal(bits(b.ranges), bits(c.ranges)) and np.array_equal(b.hit_model, c.hit_model)
    assert not np.array_equal(bits(a.ranges), bits(b.ranges))
    s.destroy()
    ctx.close()


def test_cave_tiles_match_oracle():
    """The denser map of SURVEY.md 8(d) in miniature: 2 x 2 tiles of the cave floorplan (concave polygons
    of up to 91 vertices, as the reference traced them from the bitmap), 60 robots, 360-beam scans."""
    w = worldgen.make_cave_world(seed=3, tiles=2, n_robots=60)
    x0, y0, nx, ny = w.sreg_extent()
    ctx = rt.Context(w.ppm, x0, y0, nx, ny, max_models=w.n_models + 1, max_blocks=w.n_static_blocks + len(w.robots),
                     max_cell_entries=1 << 20)
    worldgen.load_into_context(ctx, w)
    t1 = oracle_lib.T1World(w.ppm)
    t1.model_upsert(0, 0, 1.0)
    for mid in list(w.obstacle_ids) + list(w.robot_ids):
        t1.model_upsert(int(mid), int(mid), 0.5)
    st = list(w.static_blocks())
    bodies = [(m, p, w.tile_height) for m, p in st] + [(int(m), p, w.robot_size[2]) for m, p in zip(w.robot_ids, worldgen.robot_polygons(w))]
    for layer in (0, 1):
        for b, (m, p, z) in enumerate(bodies):
            t1.block_map(b, m, layer, 0.0, z, p)
    rec_g, cnt_g = ctx.grid_dump()
    rec_o, cnt_o = t1.dump_grid()
    assert np.array_equal(oracle_lib.sort_grid(rec_g), rec_o) and np.array_equal(oracle_lib.sort_counts(cnt_g), cnt_o)
    fans = worldgen.ranger_fans(w, layer=1, samples=360, range_max=30.0)
    rg, it, hits = t1_fans(t1, fans)
    head = np.concatenate([oracle_lib.ranger_headings(f["oa"], f["fov"], int(f["n_samples"])) for f in fans])
    g = ctx.trace(fans, trig=oracle_lib.libm_trig(head))
    assert np.array_equal(g.hit_model, hits["hit_model"]) and np.array_equal(bits(g.ranges), bits(rg))
    assert np.array_equal(g.steps[:, 0], hits["cell_visits"]) and np.array_equal(g.steps[:, 1], hits["region_probes"])
    assert (g.hit_model >= 0).mean() > 0.9 and 1.0 < g.ranges.mean() < 8.0  # the robots stand on open floor
    g2 = ctx.trace(fans)  # device trig
    assert np.array_equal(g2.hit_model, hits["hit_model"]) and np.array_equal(bits(g2.ranges), bits(rg))
    ctx.close()


def test_random_rays_of_every_kind_match_oracle():
    """The device against the oracle where tests/test_oracle_vs_reference.py pins the oracle to the live reference: a
    generated obstacle field plus models with every attribute the predicates read (invisible to rangers, gripper-visible,
    no obstacle, floating, a related child), robots re-rendered anywhere on alternating layers, and 200,000 random
    rays - three predicate kinds, z test on and off, any finder, origins anywhere, four ranges, both layers. Model,
    range bits, hit cell and both step counters; strict trig and the device's own."""
    w = worldgen.make_world(seed=5, n_rects=120, n_robots=30, half_extent_m=20.48)
    x0, y0, nx, ny = w.sreg_extent()
    n_static = len(w.obstacles) + len(w.robots)
    first_extra = w.n_models + 1
    ctx = rt.Context(w.ppm, x0, y0, nx, ny, max_models=first_extra + 100, max_blocks=n_static + 100, max_cell_entries=1 << 20)
    worldgen.load_into_context(ctx, w)
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
    rng = np.random.RandomState(17)
    finders = [int(i) for i in w.robot_ids] + [int(i) for i in w.obstacle_ids[:20]]
    blk, mid = n_static, first_extra
    for i in range(40):
        rr, flags = [(-1.0, rt.VIS_OBSTACLE_RETURN), (0.5, rt.VIS_OBSTACLE_RETURN | rt.VIS_GRIPPER_RETURN), (0.5, 0),
                     (0.25, rt.VIS_OBSTACLE_RETURN), (1.0, rt.VIS_OBSTACLE_RETURN)][i % 5]
        z0 = 0.0 if i % 3 else 0.7
        x, y, a = rng.uniform(-18, 18), rng.uniform(-18, 18), rng.uniform(-np.pi, np.pi)
        bodies = [(mid, mid, rr, flags, 0.0, rng.uniform(0.2, 1.5), rng.uniform(0.2, 1.5))]
        if i % 4 == 0:  # a child model: same tree, gripper-visible, sticks out
            bodies.append((mid + 1, mid, 1.0, rt.VIS_OBSTACLE_RETURN | rt.VIS_GRIPPER_RETURN, 0.4, 0.5, 0.1))
        for m, root, r_ret, fl, off, sx, sy in bodies:
            for o in (ctx, t1):
                o.model_upsert(m, root, r_ret, 0, 0, fl)
            p = worldgen.pixel_polygon(round(x + off * np.cos(a), 3), round(y + off * np.sin(a), 3), a, sx, sy, w.ppm)
            for layer in (0, 1):
                ctx.block_map(blk, m, layer, z0, z0 + 0.4, p)
                t1.block_map(blk, m, layer, z0, z0 + 0.4, p)
            finders.append(m)
            blk += 1
        mid += len(bodies)
    for t in (1, 2):  # the robots take two more poses, anywhere (across obstacles and each other too)
        layer = t % 2
        for j, rid in enumerate(w.robot_ids):
            p = worldgen.pixel_polygon(round(rng.uniform(-19, 19), 3), round(rng.uniform(-19, 19), 3), rng.uniform(-np.pi, np.pi),
                                       w.robot_size[0], w.robot_size[1], w.ppm)
            b = len(w.obstacles) + j
            ctx.block_unmap(b, layer)
            ctx.block_map(b, int(rid), layer, 0.0, w.robot_size[2], p)
            t1.block_unmap(b, layer)
            t1.block_map(b, int(rid), layer, 0.0, w.robot_size[2], p)
    rec_g, cnt_g = ctx.grid_dump()
    rec_o, cnt_o = t1.dump_grid()
    assert np.array_equal(oracle_lib.sort_grid(rec_g), rec_o) and np.array_equal(oracle_lib.sort_counts(cnt_g), cnt_o)
    n = 100000
    q = np.zeros(n, oracle_lib.T1_RAY)
    q["ox"], q["oy"] = rng.uniform(-20, 20, n), rng.uniform(-20, 20, n)
    q["oz"] = rng.choice([0.1, 0.35, 0.5, 0.9, 1.2], n)
    q["oa"] = rng.uniform(-np.pi, np.pi, n)
    q["oa"][:200] = rng.choice([0.0, np.pi / 2, -np.pi / 2, np.pi, -np.pi], 200)
    q["range"] = rng.choice([0.3, 2.0, 8.0, 30.0], n)
    q["finder"] = rng.choice(finders, n)
    q["kind"] = rng.randint(0, 3, n)
    q["ztest"] = rng.randint(0, 2, n)
    kinds_hit = set()
    for layer in (1, 0):
        q["layer"] = layer
        want = t1.raytrace(q)
        fans = rt.make_fans(n)
        fans["finder"], fans["n_samples"], fans["pred"], fans["ztest"], fans["layer"] = q["finder"], 1, q["kind"], q["ztest"], layer
        fans["angles"], fans["first_heading"] = rt.ANGLES_EXPLICIT, np.arange(n)
        for k in ("ox", "oy", "oz", "oa", "range"):
            fans[k] = q[k]
        head = np.ascontiguousarray(q["oa"])
        for trig in (oracle_lib.libm_trig(head), None):
            g = ctx.trace(fans, headings=head, trig=trig)
            bad = np.nonzero((g.hit_model != want["hit_model"]) | (bits(g.ranges) != bits(want["range"])))[0]
            assert len(bad) == 0, (layer, trig is None, len(bad), q[bad[:3]], g.hit_model[bad[:3]], g.ranges[bad[:3]], want[bad[:3]])
            h = g.hit_model >= 0
            assert np.array_equal(g.hit_cell[h, 0], want["hit_x"][h]) and np.array_equal(g.hit_cell[h, 1], want["hit_y"][h])
            assert np.array_equal(g.steps[:, 0], want["cell_visits"]) and np.array_equal(g.steps[:, 1], want["region_probes"])
        for k in range(3):
            if np.any(want["hit_model"][q["kind"] == k] >= 0):
                kinds_hit.add(k)
        assert 0.2 < np.mean(want["hit_model"] >= 0) < 0.95
    assert kinds_hit == {0, 1, 2}
    ctx.close()


def test_refill_march_kernel_equals_the_plain_one(monkeypatch):
    """k2_ray_march_refill (rt_march.cu: lanes take the next beam of their warp's chunk when their ray has ended; an
    experiment kept off by default because it is slower) must give k2_ray_march's results bit for bit: ranges, intensities,
    hit models, hit cells - with and without sensor noise, on fans that do not fill their chunks."""
    from stage_b200 import worldgen
    w = worldgen.make_world(seed=12, n_rects=500, n_robots=60, half_extent_m=30.72)
    x0, y0, nx, ny = w.sreg_extent()
    ctx = rt.Context(w.ppm, x0, y0, nx, ny, max_models=w.n_models + 1, max_blocks=len(w.obstacles) + len(w.robots), max_cell_entries=1 << 20)
    worldgen.load_into_context(ctx, w)
    want = rt.WANT_RANGES | rt.WANT_INTENSITIES | rt.WANT_HIT_MODEL | rt.WANT_HIT_CELL
    for samples in (1080, 333):
        fans = worldgen.ranger_fans(w, layer=1, samples=samples, range_max=25.0)
        fans["range"][::7] = 0.0   # rays of no length
        monkeypatch.setenv("STG_RT_MARCH", "plain")
        a = ctx.trace(fans, want=want)
        monkeypatch.setenv("STG_RT_MARCH", "refill")
        st0 = ctx.stats()
        b = ctx.trace(fans, want=want)
        assert ctx.stats()["launches_post"] - st0["launches_post"] == 2   # headings + the per-beam trig kernel: the refill path ran
        assert np.array_equal(a.ranges.view(np.uint64), b.ranges.view(np.uint64)) and np.array_equal(a.intensities, b.intensities)
        assert np.array_equal(a.hit_model, b.hit_model) and np.array_equal(a.hit_cell, b.hit_cell)
        assert (a.hit_model >= 0).mean() > 0.5
    ctx.close()


def test_streamed_intensities_equal_the_plain_delivery(monkeypatch):
    """Large batches deliver intensities while the march is still running: the warp that completes a stretch of 8192 beams
    says so through a flag in page-locked memory and host threads expand the palette bytes of finished stretches meanwhile.
    Same doubles as the delivery after the sync, tick after tick (the flags carry a sequence number), into pinned and
    ordinary arrays."""
    from stage_b200 import worldgen
    w = worldgen.make_world(seed=14, n_rects=600, n_robots=220, half_extent_m=40.96)
    x0, y0, nx, ny = w.sreg_extent()
    ctx = rt.Context(w.ppm, x0, y0, nx, ny, max_models=w.n_models + 1, max_blocks=len(w.obstacles) + len(w.robots), max_cell_entries=1 << 20)
    worldgen.load_into_context(ctx, w)
    for k, mid in enumerate(w.robot_ids):   # a few distinct ranger_return values
        ctx.model_upsert(int(mid), int(mid), 0.1 + 0.05 * (k % 9))
    ctx.sync()
    poses = w.robots
    for tick in range(12):
        poses = worldgen.step_robots(w, poses, tick, speed_m=0.3)
        fans = worldgen.ranger_fans(w, layer=1, samples=1080, range_max=25.0, robots=poses)
        res = {}
        for mode in ("0", "1"):
            monkeypatch.setenv("STG_RT_STREAM", mode)
            out = rt.Outputs(len(fans) * 1080, rt.WANT_RANGES | rt.WANT_INTENSITIES | rt.WANT_HIT_MODEL, alloc=ctx.pinned_empty if tick % 2 else None)
            out.intensities[:] = -7.0
            ctx.fans_submit(fans, out)
            ctx.sync()
            res[mode] = out
        a, b = res["0"], res["1"]
        assert np.array_equal(a.intensities, b.intensities) and np.array_equal(a.ranges.view(np.uint64), b.ranges.view(np.uint64))
        assert np.array_equal(a.hit_model, b.hit_model) and (b.intensities >= 0).all() and len(np.unique(b.intensities)) >= 5
    ctx.close()
