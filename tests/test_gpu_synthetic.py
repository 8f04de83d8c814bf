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
