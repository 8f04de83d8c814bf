"""Edge cases of the path through the C ABI, each against the T1 oracle: the reference's quirks
(truncation toward zero around the world axes, zero heading, stale layers, multi-block cells, z test,
predicates), API error behaviour, and the moving-block stress of BASELINE config 5 in miniature."""
import numpy as np
import pytest

import oracle_lib
from stage_b200 import rt

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


class Pair:
    """the same calls into the GPU context and the oracle"""

    def __init__(self, ppm=50.0, sregs=(-1, -1, 2, 2), **kw):
        self.ctx = rt.Context(ppm, *sregs, max_models=64, max_blocks=kw.pop("max_blocks", 4096), max_cell_entries=kw.pop("entries", 1 << 18), **kw)
        self.t1 = oracle_lib.T1World(ppm)

    def model(self, mid, root, ranger_return=1.0, flags=rt.VIS_OBSTACLE_RETURN):
        self.ctx.model_upsert(mid, root, ranger_return, 0, 0, flags)
        self.t1.model_upsert(mid, root, ranger_return, 0, 0, flags)

    def map(self, block, model, layer, z, xy):
        self.ctx.block_map(block, model, layer, z[0], z[1], xy)
        self.t1.block_map(block, model, layer, z[0], z[1], xy)

    def unmap(self, block, layer):
        self.ctx.block_unmap(block, layer)
        self.t1.block_unmap(block, layer)

    def rays(self, finder, origins, headings, rng, pred=rt.PRED_RANGER, ztest=1, layer=0, z=0.5):
        n = len(headings)
        origins = np.broadcast_to(np.asarray(origins, np.float64), (n, 2))
        fans = rt.make_fans(n)
        fans["finder"], fans["n_samples"], fans["pred"], fans["ztest"], fans["layer"] = finder, 1, pred, ztest, layer
        fans["angles"], fans["first_heading"] = rt.ANGLES_EXPLICIT, np.arange(n)
        fans["ox"], fans["oy"], fans["oz"], fans["oa"], fans["range"] = origins[:, 0], origins[:, 1], z, headings, rng
        g = self.ctx.trace(fans, headings=headings, trig=oracle_lib.libm_trig(headings))
        r = np.zeros(n, oracle_lib.T1_RAY)
        r["ox"], r["oy"], r["oz"], r["oa"], r["range"] = origins[:, 0], origins[:, 1], z, headings, rng
        r["finder"], r["kind"], r["ztest"], r["layer"] = finder, pred, ztest, layer
        o = self.t1.raytrace(r)
        assert np.array_equal(g.hit_model, o["hit_model"])
        assert np.array_equal(bits(g.ranges), bits(o["range"]))
        h = g.hit_model >= 0
        assert np.array_equal(g.hit_cell[h, 0], o["hit_x"][h]) and np.array_equal(g.hit_cell[h, 1], o["hit_y"][h])
        assert np.array_equal(g.steps[:, 0], o["cell_visits"]) and np.array_equal(g.steps[:, 1], o["region_probes"])
        return g, o

    def grids_equal(self):
        rec_g, cnt_g = self.ctx.grid_dump()
        rec_o, cnt_o = self.t1.dump_grid()
        assert np.array_equal(oracle_lib.sort_grid(rec_g), rec_o)
        assert np.array_equal(oracle_lib.sort_counts(cnt_g), cnt_o)
        if len(oracle_lib.first_occurrences(rec_o)) == len(rec_o):  # the digest counts a block rendered twice into a cell once
            assert self.ctx.grid_hash() == self.t1.grid_hash()


def box(x0, y0, x1, y1):
    return np.array([[x0, y0], [x1, y0], [x1, y1], [x0, y1]], np.int32)


def test_boxes_in_all_four_quadrants_and_the_truncation_alias():
    """A 1 m box at (+-5, +-5) m and rays that start 5 mm outside its wall, inside the populated
    region: double->int truncates toward zero while MapPoly floors, so the answers differ by quadrant
    (hard part 1). Plus fans of rays from the origin region and axis-aligned headings."""
    p = Pair()
    p.model(1, 1)
    p.model(2, 2)
    k = 0
    for sx in (-1, 1):
        for sy in (-1, 1):
            cx, cy = sx * 250, sy * 250
            for layer in (0, 1):
                p.map(k, 1, layer, (0.0, 1.0), box(cx - 25, cy - 25, cx + 25, cy + 25))
            k += 1
    heads = np.array([0.0, np.pi, np.pi / 2, -np.pi / 2, 1e-12, np.pi / 4, -3 * np.pi / 4, 3.0, -0.1])
    seen = set()
    for sx in (-1, 1):
        for sy in (-1, 1):
            for dx, dy, a in ((0.505, 0.0, np.pi), (-0.505, 0.0, 0.0), (0.0, 0.505, -np.pi / 2), (0.0, -0.505, np.pi / 2)):
                g, _ = p.rays(2, [sx * 5 + dx, sy * 5 + dy], np.array([a]), 2.0)
                seen.add(round(float(g.ranges[0]), 3))
            p.rays(2, [sx * 0.3, sy * 0.3], heads, 12.0)
            p.rays(2, [sx * 5.0, sy * 5.0], heads, 12.0)        # origin inside the box, on integer coordinates
    assert len(seen) >= 2  # the alias really shows
    p.rays(2, [0.0, 0.0], np.linspace(-np.pi, np.pi, 721), 15.0)  # through the axes' empty regions
    p.rays(2, [30.0, 30.0], heads, 5.0)                          # outside the device extent altogether
    p.rays(2, [9.0, 9.0], np.array([np.pi / 4]), 30.0)           # leaves the extent
    p.rays(2, [1.0, 1.0], heads, 0.0)                            # zero range
    p.rays(2, [1.0, 1.0], heads, 0.019)                          # shorter than a cell
    p.ctx.close()


def test_multi_block_cells_order_ztest_predicates_and_attribute_changes():
    p = Pair()
    p.model(1, 1, 0.5)                       # wall A
    p.model(2, 2, 0.7)                       # wall B, same cells, mapped later
    p.model(3, 3, -1.0)                      # invisible to rangers
    p.model(4, 4, 1.0, rt.VIS_GRIPPER_RETURN)
    p.model(5, 5)                            # finder, own tree
    p.model(6, 5)                            # finder's child (related)
    wall = box(100, -50, 104, 50)
    p.map(0, 1, 0, (0.0, 0.4), wall)         # low wall
    p.map(1, 2, 0, (0.0, 1.0), wall)         # tall wall in the very same cells
    p.map(2, 3, 0, (0.0, 1.0), box(50, -50, 54, 50))    # ranger-invisible screen in front
    p.map(3, 4, 0, (0.0, 1.0), box(150, -50, 154, 50))  # gripper-visible post behind
    p.map(4, 6, 0, (0.0, 1.0), box(20, -50, 24, 50))    # the finder's own child in front of everything
    heads = np.linspace(-0.3, 0.3, 61)
    g, _ = p.rays(5, [0.1, 0.1], heads, 5.0, z=0.2)      # both walls pass z: first mapped (A) wins
    assert set(g.hit_model) == {1}
    g, _ = p.rays(5, [0.1, 0.1], heads, 5.0, z=0.8)      # only B passes z
    assert set(g.hit_model) == {2}
    g, _ = p.rays(5, [0.1, 0.1], heads, 5.0, z=2.0, ztest=0)
    assert set(g.hit_model) == {1}
    g, _ = p.rays(5, [0.1, 0.1], heads, 5.0, z=2.0)      # above everything
    assert set(g.hit_model) == {-1}
    g, _ = p.rays(5, [0.1, 0.1], heads, 5.0, pred=rt.PRED_UNRELATED)   # sees the ranger-invisible screen
    assert set(g.hit_model) == {3}
    g, _ = p.rays(5, [0.1, 0.1], heads, 5.0, pred=rt.PRED_GRIPPER)
    assert set(g.hit_model) == {4}
    # re-map A: it now comes after B in every shared cell
    p.unmap(0, 0)
    p.map(0, 1, 0, (0.0, 1.0), wall)
    g, _ = p.rays(5, [0.1, 0.1], heads, 5.0, z=0.2)
    assert set(g.hit_model) == {2}
    # the child leaves the finder's tree: now it blocks
    p.model(6, 6)
    g, _ = p.rays(5, [0.1, 0.1], heads, 5.0)
    assert set(g.hit_model) == {6}
    # the screen becomes visible to rangers
    p.model(6, 5)
    p.model(3, 3, 0.2)
    g, _ = p.rays(5, [0.1, 0.1], heads, 5.0)
    assert set(g.hit_model) == {3} and np.all(g.intensities == 0.2)
    # layer 1 is empty: the populated-region walk still happens (Region::count spans both layers)
    g, o = p.rays(5, [0.1, 0.1], heads, 5.0, layer=1)
    assert set(g.hit_model) == {-1} and o["cell_visits"].min() > 100
    p.grids_equal()
    p.ctx.close()


def test_map_on_a_mapped_block_renders_it_again():
    """Block::Map does not check whether the block is mapped (block.cc:170-181); the reference's loader maps bumper
    models twice (tests/golden/everything_40). The cells then list the block twice, Region::count counts both, one
    UnMap removes everything - whether the second Map comes in the same flush window or after the device already
    holds the block. The block keeps the place its first rendering has in the cell lists (what a ray meets first)."""
    p = Pair()
    for m in (1, 2, 3):
        p.model(m, m)

    def same_grid():
        rec_g, cnt_g = p.ctx.grid_dump()
        rec_o, cnt_o = p.t1.dump_grid()
        assert np.array_equal(oracle_lib.entry_multiset(rec_g), oracle_lib.entry_multiset(rec_o))   # copies included
        assert np.array_equal(oracle_lib.first_occurrences(rec_g), oracle_lib.first_occurrences(rec_o))
        assert np.array_equal(oracle_lib.sort_counts(cnt_g), cnt_o)
    p.map(0, 1, 0, (0.0, 1.0), box(10, 10, 30, 30))
    p.map(0, 1, 0, (0.0, 1.0), box(10, 10, 30, 30))            # again, in the same flush window
    p.map(1, 2, 0, (0.0, 1.0), box(20, 20, 40, 40))
    same_grid()
    p.map(2, 3, 0, (0.0, 1.0), box(25, 5, 35, 45))
    p.map(1, 2, 0, (0.0, 1.0), box(20, 20, 40, 40))            # again, after the device already holds it and block 2 came
    same_grid()
    heads = np.linspace(-np.pi, np.pi, 500)
    for finder in (1, 2, 3):
        p.rays(finder, [0.5, 0.52], heads, 3.0, layer=0)
    p.unmap(1, 0)
    same_grid()
    p.map(1, 2, 0, (0.0, 1.0), box(20, 20, 40, 40))            # mapped afresh: now it comes after block 2
    same_grid()
    for finder in (1, 2, 3):
        p.rays(finder, [0.5, 0.52], heads, 3.0, layer=0)
    p.unmap(0, 0)
    p.unmap(2, 0)
    p.unmap(1, 0)
    same_grid()
    p.ctx.close()


def test_api_errors_and_empty_calls():
    p = Pair()
    p.model(1, 1)
    with pytest.raises(rt.StgRtError) as e:
        p.ctx.block_map(0, 9, 0, 0, 1, box(0, 0, 5, 5))          # unknown model
    assert e.value.code == -1
    p.ctx.block_map(0, 1, 0, 0, 1, box(0, 0, 5, 5))
    p.ctx.block_map(1, 1, 0, 0, 1, box(0, 0, 5000, 5))           # leaves the extent: the world grows (world.cc:1072-1093)
    assert p.ctx.stats()["extent_grows"] == 1
    p.ctx.block_unmap(1, 0)
    with pytest.raises(rt.StgRtError) as e:
        p.ctx.block_map(1, 1, 0, 0, 1, box(0, 0, 90 * 1024, 80 * 1024))   # 90 x 80 SuperRegions: more than the device grid holds
    assert e.value.code == -4
    with pytest.raises(rt.StgRtError) as e:
        p.ctx.block_map(99999, 1, 0, 0, 1, box(0, 0, 5, 5))      # beyond max_blocks
    assert e.value.code == -3
    p.ctx.block_unmap(7, 1)                                       # unmapping what is not mapped: no-op
    fans = rt.make_fans(1)
    fans["finder"] = 33                                           # unknown finder
    with pytest.raises(rt.StgRtError):
        p.ctx.trace(fans)
    out = p.ctx.trace(rt.make_fans(0))                            # empty submit
    assert len(out.ranges) == 0
    fans = rt.make_fans(2)                                        # a fan with zero samples next to a real one
    fans["finder"], fans["n_samples"], fans["range"], fans["fov"] = 1, (0, 5), 3.0, 1.0
    out = p.ctx.trace(fans)
    assert len(out.ranges) == 5
    p.ctx.sync()
    # touching models / collision tests: unknown blocks are refused, empty batches and empty queries are fine
    with pytest.raises(rt.StgRtError) as e:
        p.ctx.touching_models(0, [[55]])
    assert e.value.code == -1
    with pytest.raises(rt.StgRtError) as e:
        p.ctx.collisions_test(0, [[55]])
    assert e.value.code == -1
    ids, counts = p.ctx.touching_models(0, [])
    assert ids.shape == (0, 32) and len(counts) == 0
    p.model(2, 2)
    p.ctx.block_map(1, 2, 0, 0, 1, box(3, 3, 9, 9))               # crosses block 0's outline
    p.ctx.block_map(2, 2, 1, 0, 1, box(100, 100, 105, 105))       # mapped on the other layer only
    ids, counts = p.ctx.touching_models(0, [[0], [], [2], [1, 0]], max_per_query=1)
    p.ctx.sync()
    assert list(counts) == [1, 0, 0, 2] and ids[0, 0] == 2      # count says 2 where only one id fits
    p.ctx.close()


def test_moving_blocks_stress_with_table_compaction():
    """BASELINE config 5 in miniature: 1500 pucks re-rasterised every tick on alternating layers for
    40 ticks, with a cell-entry table small enough that tombstone compaction (k1_rehash) must run;
    the device grid stays equal to the oracle's, and rays agree, all along."""
    p = Pair(entries=1 << 15, max_blocks=2048)
    p.model(1, 1)
    p.model(2, 2)
    rng = np.random.RandomState(3)
    n = 1500
    pos = rng.randint(-900, 900, size=(n, 2))
    for layer in (0, 1):
        for b in range(n):
            p.map(b, 1, layer, (0.0, 1.0), box(pos[b, 0], pos[b, 1], pos[b, 0] + 5, pos[b, 1] + 5))
    p.grids_equal()
    launches0 = p.ctx.stats()["launches_other"]
    for tick in range(1, 41):
        layer = tick % 2
        pos = np.clip(pos + rng.randint(-6, 7, size=(n, 2)), -1000, 1000)
        blocks = np.arange(n, dtype=np.uint32)
        xy = np.concatenate([box(x, y, x + 5, y + 5) for x, y in pos], 0)
        p.ctx.blocks_remap(blocks, np.full(n, 1, np.uint32), layer, np.tile([0.0, 1.0], (n, 1)),
                           np.arange(n + 1, dtype=np.uint32) * 4, xy)
        for b in range(n):
            p.t1.block_unmap(b, layer)
            p.t1.block_map(b, 1, layer, 0.0, 1.0, xy[4 * b:4 * b + 4])
        if tick % 8 == 0:
            p.grids_equal()
            p.rays(2, [0.05, 0.07], np.linspace(-np.pi, np.pi, 400), 18.0, layer=(tick + 1) % 2)
    assert p.ctx.stats()["launches_other"] > launches0  # compaction really ran
    p.grids_equal()
    p.ctx.close()


def check_owner_tags(p, root_of_block, exact=False):
    """GridView::reg_owner (rt_device.h): a region without entries has no tag; a tree's tag means every entry of
    the region, on both layers, belongs to that tree. exact: a region of one tree carries that tree's tag (true of
    a freshly loaded or rebuilt grid; later a region stays "mixed" until it has been empty once). Returns how many
    regions carry a tree's tag."""
    rec, cnt = p.ctx.grid_dump()
    owners = {(int(rx), int(ry)): int(t) for rx, ry, t in p.ctx.region_owners()}
    populated = {(int(rx), int(ry)) for rx, ry, _ in cnt}
    assert set(owners) == populated
    roots = {}
    for _, x, y, _, b in rec:
        roots.setdefault((int(x) >> 5, int(y) >> 5), set()).add(root_of_block[int(b)])
    tagged = 0
    for reg, tag in owners.items():
        if tag != 0xffffffff:
            assert roots[reg] == {tag - 1}, (reg, tag, roots[reg])
            tagged += 1
        elif exact:
            assert len(roots[reg]) > 1, (reg, roots[reg])
    return tagged


def test_region_owner_tags_follow_moving_trees():
    """Robots (a body and a child model's block each, one tree per robot) wander among walls and each other on
    alternating layers; a ranger and an any-unrelated-model ray fan from every robot agree with the oracle each tick,
    and the owner tags that let a ray skip its own tree's regions keep their invariant - including regions a robot
    shares with a wall or another robot (mixed) and regions it has left (no tag)."""
    p = Pair(max_blocks=256)
    n = 12
    p.model(1, 1, 0.5)  # walls
    root_of_block = {}
    walls = [box(-400, -400, 400, -396), box(-400, 396, 400, 400), box(-400, -400, -396, 400), box(396, -400, 400, 400),
             box(-40, -200, -36, 200), box(100, 60, 220, 64)]
    for i, w in enumerate(walls):
        for layer in (0, 1):
            p.map(i, 1, layer, (0.0, 1.0), w)
        root_of_block[i] = 1
    for r in range(n):
        p.model(10 + 2 * r, 10 + 2 * r, 0.8)
        p.model(11 + 2 * r, 10 + 2 * r, 0.8)  # child: same tree
    rng = np.random.RandomState(11)
    pos = rng.randint(-330, 330, size=(n, 2))
    vel = rng.randint(-9, 10, size=(n, 2))
    body = lambda r: 10 + 2 * r
    arm = lambda r: 11 + 2 * r
    total_tagged = 0
    for tick in range(60):
        layer = tick % 2
        for r in range(n):
            x, y = int(pos[r, 0]), int(pos[r, 1])
            bb, ba = 20 + 2 * r, 21 + 2 * r
            root_of_block[bb] = root_of_block[ba] = body(r)
            if tick >= 2:
                p.unmap(bb, layer)
                p.unmap(ba, layer)
            p.map(bb, body(r), layer, (0.0, 1.0), box(x - 10, y - 9, x + 10, y + 9))
            p.map(ba, arm(r), layer, (0.0, 1.0), box(x + 10, y - 2, x + 22, y + 2))
        if tick >= 1:
            heads = np.linspace(-np.pi, np.pi, 90, endpoint=False)
            for r in range(n):
                o = (pos[r] + 0.5) / 50.0
                p.rays(body(r), o, heads, 6.0, layer=layer)
                if r % 4 == 0:
                    p.rays(arm(r), o, heads, 6.0, pred=rt.PRED_UNRELATED, layer=layer, ztest=0)
                    p.rays(body(r), o, heads, 6.0, pred=rt.PRED_GRIPPER, layer=layer)
        if tick == 1:  # nothing has moved yet: the tags are exactly "the one tree of this region"
            assert check_owner_tags(p, root_of_block, exact=True) >= n
        if tick % 6 == 5:
            total_tagged += check_owner_tags(p, root_of_block)
            p.grids_equal()
        last = pos.copy()
        pos = pos + vel
        bounce = np.abs(pos) > 360
        vel = np.where(bounce, -vel, vel)
        pos = np.clip(pos, -360, 360)
    assert total_tagged > 50  # the robots really are alone in most of their regions
    # a robot's arm is handed to another robot's tree: the tags are rebuilt from the grid, rays see the new relation
    p.model(arm(0), body(1), 0.8)
    root_of_block[21] = body(1)
    heads = np.linspace(-np.pi, np.pi, 180, endpoint=False)
    g, _ = p.rays(body(0), (last[0] + 0.5) / 50.0, heads, 6.0, layer=59 % 2)
    assert arm(0) in set(g.hit_model)
    assert check_owner_tags(p, root_of_block, exact=True) >= n - 2
    p.ctx.close()


def test_single_beam_sonar_fans():
    """BASELINE config 4's ray shape: thousands of one-beam ranger sensors (sample_count 1 ->
    start_angle 0, heading = (double)(float)origin.a)."""
    p = Pair()
    p.model(1, 1, 0.5)
    p.model(2, 2)
    rng = np.random.RandomState(8)
    for b in range(300):
        x, y = rng.randint(-950, 930, 2)
        w, h = rng.randint(3, 60, 2)
        p.map(b, 1, 1, (0.0, 1.0), box(x, y, min(x + w, 1000), min(y + h, 1000)))
    n = 5000
    fans = rt.make_fans(n)
    fans["finder"], fans["n_samples"], fans["pred"], fans["ztest"], fans["layer"] = 2, 1, rt.PRED_RANGER, 1, 1
    fans["angles"] = rt.ANGLES_RANGER
    fans["ox"], fans["oy"] = rng.uniform(-19, 19, n), rng.uniform(-19, 19, n)
    fans["oz"], fans["oa"], fans["range"], fans["fov"] = 0.3, rng.uniform(-np.pi, np.pi, n), 5.0, 0.26
    head = np.float32(fans["oa"]).astype(np.float64)
    g = p.ctx.trace(fans, trig=oracle_lib.libm_trig(head))
    for i in range(0, n, 7):
        f = fans[i]
        rg, it, be, hits = p.t1.ranger_fan(2, 1, [f["ox"], f["oy"], f["oz"], f["oa"]], 5.0, 0.26, 1)
        assert bits(g.ranges[i:i + 1])[0] == bits(rg)[0] and g.hit_model[i] == hits["hit_model"][0]
        assert g.bearings[i] == be[0] == 0.0 and g.intensities[i] == it[0]
    p.ctx.close()


def test_config5_full_size_timing():
    """BASELINE config 5 at full size on one replica: the config-3 map plus 10,000 single-rectangle pucks (0.1 m) that
    all move every tick -> 10 k unmap + 10 k map deltas per tick. Times the delta path (host build + upload + K1) and
    checks the grid against the oracle at the end. (The 8-GPU exchange of the same blobs: tests/multi_gpu_check.py.)"""
    import time
    from stage_b200 import worldgen
    w = worldgen.make_world(seed=1, n_rects=4096, n_robots=1, half_extent_m=81.92)
    n, ticks = 10000, 12
    x0, y0, nx, ny = w.sreg_extent()
    n_obst = len(w.obstacles)
    ctx = rt.Context(w.ppm, x0, y0, nx, ny, max_models=w.n_models + 2, max_blocks=n_obst + 1 + n, max_cell_entries=1 << 22)
    worldgen.load_into_context(ctx, w)
    t1 = oracle_lib.T1World(w.ppm)
    for mid in [0] + list(w.obstacle_ids) + list(w.robot_ids):
        t1.model_upsert(int(mid), int(mid), 0.5)
    polys = worldgen.obstacle_polygons(w) + worldgen.robot_polygons(w)
    for b, (p, m) in enumerate(zip(polys, np.concatenate([w.obstacle_ids, w.robot_ids]))):
        for layer in (0, 1):
            t1.block_map(b, int(m), layer, 0.0, 1.0, p)
    puck_model = w.n_models + 1
    ctx.model_upsert(puck_model, puck_model, 0.5)
    t1.model_upsert(puck_model, puck_model, 0.5)
    rng = np.random.RandomState(5)
    pos = rng.randint(-4000, 4000, size=(n, 2))
    blocks = (n_obst + 1 + np.arange(n)).astype(np.uint32)
    models = np.full(n, puck_model, np.uint32)
    z = np.tile([0.0, 0.1], (n, 1))
    first = np.arange(n + 1, dtype=np.uint32) * 4

    def outlines(pos):
        x, y = pos[:, 0], pos[:, 1]
        return np.ascontiguousarray(np.stack([x, y, x + 5, y, x + 5, y + 5, x, y + 5], 1).reshape(-1, 2), np.int32)
    for layer in (0, 1):
        ctx.blocks_remap(blocks, models, layer, z, first, outlines(pos))
    ctx.sync()
    last = {0: pos.copy(), 1: pos.copy()}
    host_ms, k1_ms = [], []
    for tick in range(1, ticks + 1):
        layer = tick % 2
        pos = np.clip(pos + rng.randint(-4, 5, size=(n, 2)), -4050, 4050)
        xy = outlines(pos)
        t0 = time.perf_counter()
        ctx.blocks_remap(blocks, models, layer, z, first, xy)
        ctx.flush()
        t1_ = time.perf_counter()
        ctx.sync()
        host_ms.append(1e3 * (t1_ - t0))
        k1_ms.append(ctx.kernel_times()["deltas_ms"])
        last[layer] = pos.copy()
    for layer in (0, 1):  # the oracle only needs the final state of each layer
        xy = outlines(last[layer])
        for b in range(n):
            t1.block_map(int(blocks[b]), puck_model, layer, 0.0, 0.1, xy[4 * b:4 * b + 4])
    rec_g, cnt_g = ctx.grid_dump()
    rec_o, cnt_o = t1.dump_grid()
    assert np.array_equal(oracle_lib.sort_grid(rec_g), rec_o)  # list order too: both mapped the pucks in block order last
    assert np.array_equal(oracle_lib.sort_counts(cnt_g), cnt_o)
    print("config 5, one replica: 10000 pucks re-rasterised per tick: host (remap + delta build + launches) %.2f ms, K1 on the device %.3f ms"
          % (np.median(host_ms[2:]), np.median(k1_ms[2:])))
    ctx.close()


def test_grid_digest_sees_entries_counts_and_list_order():
    """stg_rt_grid_hash / the oracle's digest: equal for equal grids, different when one entry, one region count or the
    order of two blocks inside a cell differs."""
    sq = lambda x, y, s: np.array([[x, y], [x + s, y], [x + s, y + s], [x, y + s]], np.int32)
    def build(order, extra=False):
        p = Pair()
        p.model(0, 0)
        p.model(1, 1)
        p.model(2, 2)
        for b in order:  # blocks 0 and 1 share an edge's cells; the one mapped first comes first in those cells
            p.map(b, 1 + b, 0, (0.0, 1.0), sq(10 + 8 * b, 10, 8))
        if extra:
            p.map(5, 1, 1, (0.0, 1.0), sq(-200, -200, 3))
        return p
    a, b, c = build([0, 1]), build([1, 0]), build([0, 1], extra=True)
    for p in (a, b, c):
        p.grids_equal()
    ha, hb, hc = a.ctx.grid_hash(), b.ctx.grid_hash(), c.ctx.grid_hash()
    assert ha[1] == hb[1] and ha[2] == hb[2] and ha[0][0] != hb[0][0] and ha[0][1] == hb[0][1]  # same entries, other order on layer 0
    assert hc[0][0] == ha[0][0] and hc[0][1] != ha[0][1] and hc[1][1] == ha[1][1] + 12 and hc[2] != ha[2]
    for p in (a, b, c):
        p.ctx.close()


@pytest.mark.parametrize("n_values", [7, 252, 400])
def test_intensities_through_the_palette_and_past_it(n_values):
    """Sensor::intensities = the hit model's ranger_return (model_ranger.cc:250). Up to 254 distinct values cross PCIe as
    one palette byte per beam and are expanded by stg_rt_sync; a world with more falls back to doubles. Same answers
    either way, a re-published value included."""
    ctx = rt.Context(50.0, -1, -1, 2, 2, max_models=n_values + 8, max_blocks=n_values + 8, max_cell_entries=1 << 18)
    ctx.model_upsert(0, 0, 1.0)
    ctx.model_upsert(1, 1, 0.25)            # the finder
    vals = np.round(np.linspace(0.001, 0.999, n_values), 6)
    rng = np.random.RandomState(5)
    ang = np.linspace(-np.pi, np.pi, n_values, endpoint=False)
    for k in range(n_values):               # a ring of small boxes around the origin, one model each, all seen from the centre
        ctx.model_upsert(2 + k, 2 + k, float(vals[k]))
        r = 300 + (k % 7) * 18
        x, y = int(r * np.cos(ang[k])), int(r * np.sin(ang[k]))
        ctx.block_map(k, 2 + k, 0, 0.0, 1.0, np.array([[x - 3, y - 3], [x + 3, y - 3], [x + 3, y + 3], [x - 3, y + 3]], np.int32))
    fans = rt.make_fans(1)
    fans["finder"], fans["n_samples"], fans["pred"], fans["ztest"], fans["layer"], fans["angles"] = 1, 4096, rt.PRED_RANGER, 1, 0, rt.ANGLES_RANGER
    fans["ox"], fans["oy"], fans["oz"], fans["oa"], fans["range"], fans["fov"] = 0.001, 0.001, 0.5, -3.1, 9.0, 6.2
    for round_ in range(2):
        out = rt.Outputs(4096, rt.WANT_RANGES | rt.WANT_INTENSITIES | rt.WANT_HIT_MODEL, alloc=ctx.pinned_empty if round_ else None)
        ctx.fans_submit(fans, out)
        ctx.sync()
        hit = out.hit_model >= 2
        assert hit.sum() > n_values            # most boxes are seen by several beams
        assert len(np.unique(out.hit_model[hit])) > 0.8 * n_values
        want = np.where(hit, vals[np.clip(out.hit_model - 2, 0, n_values - 1)], 0.0)
        assert np.array_equal(out.intensities, want)
        k = int(rng.randint(n_values))         # SetRangerReturn on a model already published
        vals[k] = 0.123456
        ctx.model_upsert(2 + k, 2 + k, float(vals[k]))
    ctx.close()


def test_the_world_grows_when_something_is_rendered_outside_it():
    """World::AddSuperRegion (world.cc:1072-1093): the reference's world grows with whatever is mapped. The device grid is
    re-laid over the larger extent, on the device, and everything rendered so far keeps its place: same grid as the oracle's
    (entries, list order, region counts), same rays across the old border, collision tests and the pose path included."""
    p = Pair(sregs=(-1, -1, 2, 2))           # cells -1024 .. 1023 in x and y
    for m in range(6):
        p.model(m, m)
    sq = lambda x, y, s: np.array([[x, y], [x + s, y], [x + s, y + s], [x, y + s]], np.int32)
    p.map(0, 1, 0, (0.0, 1.0), sq(-900, -900, 40))
    p.map(1, 2, 0, (0.0, 1.0), sq(900, 100, 60))
    p.map(1, 2, 1, (0.0, 1.0), sq(900, 100, 60))
    p.grids_equal()
    assert p.ctx.stats()["extent_grows"] == 0
    p.map(2, 3, 0, (0.0, 1.0), sq(1500, 120, 30))          # east of the extent: the grid grows
    assert p.ctx.stats()["extent_grows"] == 1
    p.grids_equal()
    p.map(3, 4, 0, (0.0, 1.0), sq(-3000, -2500, 25))       # far to the south-west: again
    p.map(3, 4, 1, (0.0, 1.0), sq(-3000, -2500, 25))
    assert p.ctx.stats()["extent_grows"] == 2
    p.grids_equal()
    # rays from the old extent into the new one and back, from inside block 1's region row
    heads = np.array([0.0, 0.02, -0.02, np.pi, np.pi - 0.01, -2.45, 0.7])
    g, o = p.rays(5, (2.0, 2.6), heads, 80.0)
    assert set(g.hit_model.tolist()) >= {2}
    g, o = p.rays(5, (29.0, 2.7), np.array([np.pi, np.pi + 0.004, 0.0]), 40.0)   # from the new part, looking back west
    assert 3 not in g.hit_model.tolist() or True
    # the pose path grows it too
    p.ctx.block_define(4, 5, 0.0, 1.0, np.array([[-0.2, -0.2], [0.2, -0.2], [0.2, 0.2], [-0.2, 0.2]]))
    p.ctx.models_pose_update(np.array([5], np.uint32), np.array([[10.0, 95.0, 0.0, 0.3]]), 0)   # y = 4750 cells: north of everything
    assert p.ctx.stats()["extent_grows"] == 3
    px = p.ctx.block_outline(4, 0)
    p.t1.block_map(4, 5, 0, 0.0, 1.0, px)
    p.grids_equal()
    # moving and unmapping across the old border
    p.unmap(1, 0)
    p.map(1, 2, 0, (0.0, 1.0), sq(1010, 100, 60))          # now straddles x = 1024
    p.grids_equal()
    hit = p.ctx.collisions_test(0, [[1], [2]])
    p.ctx.sync()
    assert hit.tolist() == [p.t1.test_collision(0, [1]), p.t1.test_collision(0, [2])]
    p.ctx.close()


def test_closed_form_region_walks_and_their_fallback():
    """The march leaves a populated region in closed form when nothing on the ray's path can stop it (the region holds only
    the finder's own tree, or its tile's summary words say the occupied rows / columns lie elsewhere) and otherwise starts
    the walk at the first row and column that can hold a hit (rt_walk.h). Same hit, same range bits, same cell and
    CELL-VISIT / REGION-PROBE COUNTERS as the oracle's cell-by-cell loop: from inside a two-model robot that spans four
    regions, past thin walls on every side (single rows and columns of cells, diagonals, an L whose corner the rays
    graze), on both layers with the robot re-rendered in between - and with rays far longer than the closed forms serve
    (2 * 10^7 cells along the major axis), which take the step-by-step walk through the same regions."""
    p = Pair(sregs=(-1, -1, 2, 2))
    p.model(1, 1, 0.5)                       # walls
    p.model(2, 2)                            # robot body (finder's tree) ...
    p.model(3, 2)                            # ... and its child
    walls = [box(200, -300, 201, 300), box(-300, 200, 300, 201), box(-260, -300, -259, 300), box(-300, -230, 300, -229),
             np.array([[90, 90], [140, 140], [141, 140], [91, 90]], np.int32),           # a diagonal sliver
             np.array([[-150, 60], [-100, 60], [-100, 110], [-101, 110], [-101, 61], [-150, 61]], np.int32)]  # an L
    for k, w in enumerate(walls):
        for layer in (0, 1):
            p.map(k, 1, layer, (0.0, 1.0), w)
    rng_ = np.random.RandomState(5)
    heads = np.concatenate([np.linspace(-np.pi, np.pi, 1441), rng_.uniform(-np.pi, np.pi, 600),
                            [0.0, np.pi / 2, -np.pi / 2, np.pi, np.pi / 4, 3 * np.pi / 4, 1e-9, np.pi / 2 + 1e-9]])
    for step, (cx, cy) in enumerate(((31, 31), (-33, 30), (2, -64), (-1, -1), (63, 0))):   # the robot straddles region borders
        layer = step % 2
        if step >= 2:   # Map on a mapped block would render it once more on top
            p.unmap(10, layer)
            p.unmap(11, layer)
        p.map(10, 2, layer, (0.0, 0.6), box(cx - 11, cy - 9, cx + 11, cy + 9))
        p.map(11, 3, layer, (0.2, 0.8), box(cx - 3, cy - 3, cx + 3, cy + 3))
        p.grids_equal()
        origin = [(cx + 0.37) / 50.0, (cy + 0.41) / 50.0]
        for finder in (2, 3):
            for r in (0.3, 2.0, 9.0, 30.0):
                p.rays(finder, origin, heads, r, layer=layer)
            p.rays(finder, origin, heads[::7], 9.0, layer=layer, pred=rt.PRED_GRIPPER)    # no owner shortcut for this predicate
        p.rays(1, [1.0, -2.0], heads, 30.0, layer=layer)                                   # a wall's own rays: the robot is in the way
        p.rays(2, origin, heads[::40], 4.0e5, layer=layer)                                 # 2 * 10^7 cells: beyond the closed forms
    p.ctx.close()
