"""Synthetic worlds for the raycast path (BASELINE.json configs 3-5).

A world is described in metres exactly as a Stage world file would describe it (every obstacle and
robot is a `model` with a pose and a size, body = the default unit-square block scaled to the size),
so the same description can be (a) turned into pixel-space polygons for the GPU path here and
(b) written out as world-file text. `pixel_polygon` follows Model::LocalToPixels
(model.cc:474-491) after BlockGroup::CalcSize (blockgroup.cc:84-112) operation for operation, with
the host libm's cos/sin, so both routes rasterise identical polygons.

All coordinates live on a 1 mm / 0.001 degree lattice so that their decimal text form parses back
to the same doubles.
"""
import math
import os

import numpy as np

from . import rt


class World:
    def __init__(self, ppm, half_extent_m, obstacles, robots, obstacle_height=1.0, robot_size=(0.44, 0.38, 0.6)):
        self.ppm = ppm
        self.half_extent_m = half_extent_m
        self.obstacles = obstacles      # (n, 4) float64: centre x, centre y, size x, size y  [m]
        self.robots = robots            # (n, 3) float64: x, y [m], heading [degrees]
        self.obstacle_height = obstacle_height
        self.robot_size = robot_size
        # ids as the reference assigns them: 0 is the world's ground model, then file order
        self.obstacle_ids = 1 + np.arange(len(obstacles), dtype=np.uint32)
        self.robot_ids = 1 + len(obstacles) + np.arange(len(robots), dtype=np.uint32)

    @property
    def n_models(self):
        return 1 + len(self.obstacles) + len(self.robots)

    @property
    def n_static_blocks(self):
        """blocks of everything that is not a robot; robot j's body is block n_static_blocks + j"""
        return len(self.obstacles)

    def sreg_extent(self):
        """(x0, y0, nx, ny) in superregions (1024 cells) covering the world"""
        half_cells = int(math.ceil(self.half_extent_m * self.ppm))
        lo = (-half_cells) >> 10
        hi = (half_cells - 1) >> 10
        return lo, lo, hi - lo + 1, hi - lo + 1


def _lattice(v, step=0.001):
    return np.round(np.asarray(v, np.float64) / step) * 1.0 / (1.0 / step)


def make_world(seed=1, n_rects=4096, n_robots=1000, half_extent_m=81.92, ppm=50.0, side_range=(0.5, 4.0)):
    """Seeded random axis-aligned rectangles plus four boundary walls, and robots at poses that are
    outside every rectangle (0.5 m clearance) and at least 0.6 m from each other."""
    rng = np.random.RandomState(seed)
    H = half_extent_m
    margin = 0.5
    sides = rng.uniform(side_range[0], side_range[1], size=(n_rects, 2))
    sides = np.maximum(np.round(sides * 1000) / 1000.0, 0.04)
    cx = rng.uniform(-H + margin, H - margin, size=n_rects)
    cy = rng.uniform(-H + margin, H - margin, size=n_rects)
    # keep rectangles inside the walls
    cx = np.clip(cx, -H + margin + sides[:, 0] / 2, H - margin - sides[:, 0] / 2)
    cy = np.clip(cy, -H + margin + sides[:, 1] / 2, H - margin - sides[:, 1] / 2)
    cx, cy = np.round(cx * 1000) / 1000.0, np.round(cy * 1000) / 1000.0
    wall_t = 0.1
    walls = np.array([[0.0, -H + 0.25, 2 * H - 0.4, wall_t], [0.0, H - 0.25, 2 * H - 0.4, wall_t],
                      [-H + 0.25, 0.0, wall_t, 2 * H - 0.4], [H - 0.25, 0.0, wall_t, 2 * H - 0.4]])
    obstacles = np.concatenate([np.stack([cx, cy, sides[:, 0], sides[:, 1]], 1), walls], 0)

    robots = np.zeros((n_robots, 3))
    placed = 0
    clear = 0.5
    cell = 1.0
    occupied = {}
    ox0, ox1 = obstacles[:, 0] - obstacles[:, 2] / 2 - clear, obstacles[:, 0] + obstacles[:, 2] / 2 + clear
    oy0, oy1 = obstacles[:, 1] - obstacles[:, 3] / 2 - clear, obstacles[:, 1] + obstacles[:, 3] / 2 + clear
    rounds = 0
    while placed < n_robots:
        rounds += 1
        if rounds > 200:
            raise ValueError("cannot place %d robots 0.6 m apart in this world (placed %d)" % (n_robots, placed))
        cand = rng.uniform(-H + 1.5, H - 1.5, size=(4 * (n_robots - placed) + 16, 2))
        cand = np.round(cand * 1000) / 1000.0
        for x, y in cand:
            if np.any((x > ox0) & (x < ox1) & (y > oy0) & (y < oy1)):
                continue
            key = (int(math.floor(x / cell)), int(math.floor(y / cell)))
            near = False
            for dx in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    for (px, py) in occupied.get((key[0] + dx, key[1] + dy), ()):
                        if (px - x) ** 2 + (py - y) ** 2 < 0.6 ** 2:
                            near = True
            if near:
                continue
            occupied.setdefault(key, []).append((x, y))
            robots[placed] = (x, y, np.round(rng.uniform(-180.0, 180.0) * 1000) / 1000.0)
            placed += 1
            if placed == n_robots:
                break
    return World(ppm, half_extent_m, obstacles, robots)


def pixel_polygon(x, y, a_rad, sx, sy, ppm):
    """The 4 pixel-space vertices Model::LocalToPixels yields for a model at global pose (x, y, a)
    whose body is the default unit-square block scaled to (sx, sy).

    model.cc:378-392 AddBlockRect(-0.5,-0.5,1,1): vertices (-.5,-.5) (.5,-.5) (.5,.5) (-.5,.5);
    blockgroup.cc:84-112 CalcSize: (pt - offset) * (modsize / size) with size 1, offset 0;
    model.cc:474-491: gpose = GetGlobalPose() + geom.pose; pt -> gpose + Pose(pt) (stage.hh:297-304:
    x + px*cos(a) - py*sin(a), y + px*sin(a) + py*cos(a)); pixel = floor(coordinate * ppm)."""
    cosa, sina = math.cos(a_rad), math.sin(a_rad)
    # gpose + geom.pose with geom.pose == 0: x + 0*cosa - 0*sina
    gx = x + 0.0 * cosa - 0.0 * sina
    gy = y + 0.0 * sina + 0.0 * cosa
    out = np.zeros((4, 2), np.int32)
    for i, (ux, uy) in enumerate(((-0.5, -0.5), (0.5, -0.5), (0.5, 0.5), (-0.5, 0.5))):
        lx = (ux - 0.0) * (sx / 1.0)
        ly = (uy - 0.0) * (sy / 1.0)
        px = gx + lx * cosa - ly * sina
        py = gy + lx * sina + ly * cosa
        out[i, 0] = int(math.floor(px * ppm))
        out[i, 1] = int(math.floor(py * ppm))
    return out


def heading_rad(deg):
    """world-file degrees -> radians as Worldfile::ReadTuple does (worldfile.cc:63,1616)"""
    return deg * (math.pi / 180.0)


def obstacle_polygons(w):
    return [pixel_polygon(o[0], o[1], 0.0, o[2], o[3], w.ppm) for o in w.obstacles]


def robot_polygons(w, robots=None):
    robots = w.robots if robots is None else robots
    return [pixel_polygon(r[0], r[1], heading_rad(r[2]), w.robot_size[0], w.robot_size[1], w.ppm) for r in robots]


def load_into_context(ctx, w, robot_slice=None):
    """Publish every model and map every body on BOTH layers, as World::LoadWorldPostHook does
    (world.cc:459-465). Block ids: obstacle i -> i, robot j -> n_obstacles + j."""
    if isinstance(w, CaveWorld):
        return w.load_into_context(ctx)
    ctx.model_upsert(0, 0, 1.0)  # the ground model
    for mid in w.obstacle_ids:
        ctx.model_upsert(int(mid), int(mid), 0.5)
    for mid in w.robot_ids:
        ctx.model_upsert(int(mid), int(mid), 0.5)
    polys = obstacle_polygons(w) + robot_polygons(w)
    models = np.concatenate([w.obstacle_ids, w.robot_ids])
    zmax = np.concatenate([np.full(len(w.obstacles), w.obstacle_height), np.full(len(w.robots), w.robot_size[2])])
    first = np.arange(len(polys) + 1, dtype=np.uint32) * 4
    xy = np.concatenate(polys, 0)
    z = np.stack([np.zeros(len(polys)), zmax], 1)
    blocks = np.arange(len(polys), dtype=np.uint32)
    for layer in (0, 1):
        ctx.blocks_remap(blocks, models, layer, z, first, xy)
    ctx.sync()
    return len(polys)


class CaveWorld(World):
    """The denser map of SURVEY.md 8(d): the floorplan of worlds/simple.world (bitmaps/cave.png, 16 m,
    2 cm cells) tiled tiles x tiles, one floorplan model per tile. The outlines are the polygons the
    unmodified reference produced for that bitmap (stage_b200/data/cave_polygons.npz, made by
    tools/extract_cave_polygons.py from a recorded trace), shifted by whole tiles."""
    TILE_CELLS = 800

    def __init__(self, tiles, robots, ppm=50.0):
        data = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "cave_polygons.npz"))
        first, xy = data["first"], data["xy"]
        self.tile_polygons = [xy[first[i]:first[i + 1]] for i in range(len(first) - 1)]
        self.tiles = tiles
        self.tile_height = 0.8
        half_cells = tiles * self.TILE_CELLS // 2
        super().__init__(ppm, max(half_cells / ppm, 81.92), np.zeros((tiles * tiles, 4)), robots)
        # tile (i, j) is centred at cell ((i + 0.5) * 800 - half, (j + 0.5) * 800 - half)
        self.tile_offsets = [(int((i + 0.5) * self.TILE_CELLS) - half_cells, int((j + 0.5) * self.TILE_CELLS) - half_cells)
                             for j in range(tiles) for i in range(tiles)]

    @property
    def n_static_blocks(self):
        return len(self.tile_offsets) * len(self.tile_polygons)

    def static_blocks(self):
        """(model id, polygon) of every floorplan block, in block-id order"""
        for t, (dx, dy) in enumerate(self.tile_offsets):
            for p in self.tile_polygons:
                yield int(self.obstacle_ids[t]), p + np.array([dx, dy], np.int32)

    def load_into_context(self, ctx):
        ctx.model_upsert(0, 0, 1.0)
        for mid in list(self.obstacle_ids) + list(self.robot_ids):
            ctx.model_upsert(int(mid), int(mid), 0.5)
        st = list(self.static_blocks())
        polys = [p for _, p in st] + robot_polygons(self)
        models = np.concatenate([np.array([m for m, _ in st], np.uint32), self.robot_ids])
        zmax = np.concatenate([np.full(len(st), self.tile_height), np.full(len(self.robots), self.robot_size[2])])
        first = np.zeros(len(polys) + 1, np.uint32)
        first[1:] = np.cumsum([len(p) for p in polys])
        xy = np.ascontiguousarray(np.concatenate(polys, 0), np.int32)
        z = np.stack([np.zeros(len(polys)), zmax], 1)
        blocks = np.arange(len(polys), dtype=np.uint32)
        for layer in (0, 1):
            ctx.blocks_remap(blocks, models, layer, z, first, xy)
        ctx.sync()
        return len(polys)


def make_cave_world(seed=1, tiles=10, n_robots=1000, clearance_m=0.35):
    """tiles x tiles cave floorplans and robots dropped on free floor (not within clearance_m of a wall)."""
    from PIL import Image, ImageDraw
    from scipy import ndimage
    rng = np.random.RandomState(seed)
    w = CaveWorld(tiles, np.zeros((n_robots, 3)))
    T = CaveWorld.TILE_CELLS
    img = Image.new("1", (T, T), 0)
    draw = ImageDraw.Draw(img)
    for p in w.tile_polygons[1:]:  # polygon 0 is the bounding square (floorplan `boundary 1`)
        draw.polygon([(int(x) + T // 2, int(y) + T // 2) for x, y in p], fill=1, outline=1)
    wall = np.array(img, bool)
    wall[:2, :] = wall[-2:, :] = wall[:, :2] = wall[:, -2:] = True
    free = ~ndimage.binary_dilation(wall, iterations=int(round(clearance_m * w.ppm)))
    fy, fx = np.nonzero(free)
    assert len(fx) > 1000
    placed, occupied, rounds = 0, {}, 0
    while placed < n_robots:
        rounds += 1
        if rounds > 200:
            raise ValueError("cannot place %d robots in the cave tiles (placed %d)" % (n_robots, placed))
        for _ in range(4 * (n_robots - placed) + 16):
            t = rng.randint(len(w.tile_offsets))
            k = rng.randint(len(fx))
            dx, dy = w.tile_offsets[t]
            x = round((fx[k] - T // 2 + dx + 0.5) / w.ppm, 3)
            y = round((fy[k] - T // 2 + dy + 0.5) / w.ppm, 3)
            key = (int(math.floor(x)), int(math.floor(y)))
            if any((px - x) ** 2 + (py - y) ** 2 < 0.36 for ddx in (-1, 0, 1) for ddy in (-1, 0, 1)
                   for (px, py) in occupied.get((key[0] + ddx, key[1] + ddy), ())):
                continue
            occupied.setdefault(key, []).append((x, y))
            w.robots[placed] = (x, y, round(rng.uniform(-180.0, 180.0), 3))
            placed += 1
            if placed == n_robots:
                break
    return w


def ranger_fans(w, layer, fov_deg=270.0, samples=1080, range_max=30.0, z=0.5, robots=None, robot_ids=None):
    """One laser fan per robot, mounted at the robot centre: the fan record a patched
    ModelRanger::Sensor::Update would submit (origin = pose of the first ray, model_ranger.cc:222-230)."""
    robots = w.robots if robots is None else robots
    robot_ids = w.robot_ids if robot_ids is None else robot_ids
    fov = fov_deg * (math.pi / 180.0)
    fans = rt.make_fans(len(robots))
    fans["finder"] = robot_ids
    fans["n_samples"] = samples
    fans["pred"] = rt.PRED_RANGER
    fans["ztest"] = 1
    fans["layer"] = layer
    fans["angles"] = rt.ANGLES_RANGER
    fans["ox"] = robots[:, 0]
    fans["oy"] = robots[:, 1]
    fans["oz"] = z
    start = -fov / 2.0 if samples > 1 else 0.0
    a = np.array([heading_rad(d) for d in robots[:, 2]]) + start
    # Pose::operator+ normalises the heading into [-pi, pi] (stage.hh:163-170, :303)
    for i in range(len(a)):
        while a[i] < -math.pi:
            a[i] += 2.0 * math.pi
        while a[i] > math.pi:
            a[i] -= 2.0 * math.pi
    fans["oa"] = a
    fans["range"] = range_max
    fans["fov"] = fov
    return fans


def to_worldfile_text(w):
    """The same world as Stage world-file text (resolution = 1/ppm)."""
    lines = ["resolution %.17g" % (1.0 / w.ppm), "interval_sim 100", "threads 1", "quit_time 3600", ""]
    for i, o in enumerate(w.obstacles):
        lines.append('model ( name "o%d" pose [ %.3f %.3f 0 0 ] size [ %.3f %.3f %.3f ] ranger_return 0.5 )'
                     % (i, o[0], o[1], o[2], o[3], w.obstacle_height))
    for i, r in enumerate(w.robots):
        lines.append('model ( name "r%d" pose [ %.3f %.3f 0 %.3f ] size [ %.3f %.3f %.3f ] ranger_return 0.5 )'
                     % (i, r[0], r[1], r[2], w.robot_size[0], w.robot_size[1], w.robot_size[2]))
    return "\n".join(lines) + "\n"


def step_robots(w, robots, tick, speed_m=0.05):
    """A deterministic per-tick motion: every robot creeps forward along its heading by speed_m and
    turns by 1 degree; robots that would leave the free area are turned round instead."""
    out = robots.copy()
    a = out[:, 2] * (math.pi / 180.0)
    nx = np.round((out[:, 0] + speed_m * np.cos(a)) * 1000) / 1000.0
    ny = np.round((out[:, 1] + speed_m * np.sin(a)) * 1000) / 1000.0
    H = w.half_extent_m - 1.5
    ok = (np.abs(nx) < H) & (np.abs(ny) < H)
    out[ok, 0], out[ok, 1] = nx[ok], ny[ok]
    out[:, 2] = np.where(ok, out[:, 2] + 1.0, out[:, 2] + 180.0)
    out[:, 2] = np.round(((out[:, 2] + 180.0) % 360.0 - 180.0) * 1000) / 1000.0
    return out


# ---------------------------------------------------------------------------------------------------
# BASELINE config 4 (SURVEY.md 8d "C4"): a swarm of small robots with sonar rings and fiducial finders;
# config 5 ("C5"): pucks that all move every tick
# ---------------------------------------------------------------------------------------------------

# worlds/pioneer.inc:24-39, `define p2dx_sonar ranger`: pose [x y z heading-in-degrees] of the 16 transducers;
# :9-19 `define p2dxsonar sensor`: size [0.01 0.05 0.01], range [0 5.0], fov 15, samples 1
P2DX_SONAR_POSES = (
    (0.075, 0.130, 90), (0.115, 0.115, 50), (0.150, 0.080, 30), (0.170, 0.025, 10),
    (0.170, -0.025, -10), (0.150, -0.080, -30), (0.115, -0.115, -50), (0.075, -0.130, -90),
    (-0.155, -0.130, -90), (-0.195, -0.115, -130), (-0.230, -0.080, -150), (-0.250, -0.025, -170),
    (-0.250, 0.025, 170), (-0.230, 0.080, 150), (-0.195, 0.115, 130), (-0.155, 0.130, 90))


def p2dx_sonar_sensors():
    """the sensor list of a p2dx_sonar ranger as stg_rt_ranger_sensor records (what Sensor::Load reads, model_ranger.cc:129-156)"""
    s = np.zeros(len(P2DX_SONAR_POSES), rt.RANGER_SENSOR_DTYPE)
    for i, (x, y, deg) in enumerate(P2DX_SONAR_POSES):
        s[i]["px"], s[i]["py"], s[i]["pz"], s[i]["pa"] = x, y, 0.0, heading_rad(deg)
    s["size_z"], s["range_max"], s["fov"], s["n_samples"] = 0.01, 5.0, heading_rad(15.0), 1
    return s


def libm_cos_sin(a):
    """cos / sin of every heading from the C library (math.cos / math.sin), not numpy's own vector routines: the pixel
    outlines must be the ones Model::LocalToPixels computes"""
    a = np.asarray(a, np.float64)
    return np.fromiter(map(math.cos, a), np.float64, len(a)), np.fromiter(map(math.sin, a), np.float64, len(a))


def box_polygons(x, y, a_rad, sx, sy, ppm):
    """pixel_polygon for many boxes at once: (n, 4, 2) int32, operation for operation the same arithmetic"""
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    cosa, sina = libm_cos_sin(a_rad)
    gx = (x + 0.0 * cosa) - 0.0 * sina
    gy = (y + 0.0 * sina) + 0.0 * cosa
    out = np.zeros((len(x), 4, 2), np.int32)
    for i, (ux, uy) in enumerate(((-0.5, -0.5), (0.5, -0.5), (0.5, 0.5), (-0.5, 0.5))):
        lx = (ux - 0.0) * (sx / 1.0)
        ly = (uy - 0.0) * (sy / 1.0)
        px = (gx + lx * cosa) - ly * sina
        py = (gy + lx * sina) + ly * cosa
        out[:, i, 0] = np.floor(px * ppm).astype(np.int32)
        out[:, i, 1] = np.floor(py * ppm).astype(np.int32)
    return out


def robot_boxes(w, robots=None):
    """robot_polygons as one (n, 4, 2) array (swarms: no Python loop per robot)"""
    robots = w.robots if robots is None else robots
    return box_polygons(robots[:, 0], robots[:, 1], robots[:, 2] * (math.pi / 180.0), w.robot_size[0], w.robot_size[1], w.ppm)


def make_swarm_world(seed=1, n_rects=4096, n_robots=100000, half_extent_m=81.92, ppm=50.0, robot_size=(0.1, 0.1, 0.1),
                     pitch_m=0.25, jitter_m=0.03, clearance_m=0.1):
    """The config-3 obstacle field with n_robots small bodies (C4: 0.1 m squares carrying fiducials; C5: pucks) on a
    jittered lattice: never inside or within clearance_m of a rectangle, never closer than pitch - 2 jitter to each other."""
    base = make_world(seed=seed, n_rects=n_rects, n_robots=0, half_extent_m=half_extent_m, ppm=ppm)
    rng = np.random.RandomState(seed + 1000)
    H = half_extent_m - 1.0
    k = int(math.floor(2 * H / pitch_m))
    gx, gy = np.meshgrid(np.arange(k), np.arange(k), indexing="ij")
    cx = -H + (gx.reshape(-1) + 0.5) * pitch_m + rng.uniform(-jitter_m, jitter_m, k * k)
    cy = -H + (gy.reshape(-1) + 0.5) * pitch_m + rng.uniform(-jitter_m, jitter_m, k * k)
    cx, cy = np.round(cx * 1000) / 1000.0, np.round(cy * 1000) / 1000.0
    # a coarse bitmap of where the rectangles (grown by the clearance and half a body diagonal) are
    res = 0.05
    n = int(math.ceil(2 * half_extent_m / res))
    blocked = np.zeros((n, n), bool)
    grow = clearance_m + 0.75 * max(robot_size[0], robot_size[1])
    o = base.obstacles
    x0 = np.clip(np.floor((o[:, 0] - o[:, 2] / 2 - grow + half_extent_m) / res).astype(int), 0, n)
    x1 = np.clip(np.ceil((o[:, 0] + o[:, 2] / 2 + grow + half_extent_m) / res).astype(int), 0, n)
    y0 = np.clip(np.floor((o[:, 1] - o[:, 3] / 2 - grow + half_extent_m) / res).astype(int), 0, n)
    y1 = np.clip(np.ceil((o[:, 1] + o[:, 3] / 2 + grow + half_extent_m) / res).astype(int), 0, n)
    for i in range(len(o)):
        blocked[x0[i]:x1[i], y0[i]:y1[i]] = True
    ix = np.clip(((cx + half_extent_m) / res).astype(int), 0, n - 1)
    iy = np.clip(((cy + half_extent_m) / res).astype(int), 0, n - 1)
    free = np.nonzero(~blocked[ix, iy])[0]
    if len(free) < n_robots:
        raise ValueError("only %d free lattice points for %d robots; lower pitch_m" % (len(free), n_robots))
    pick = np.sort(rng.permutation(free)[:n_robots])
    robots = np.stack([cx[pick], cy[pick], np.round(rng.uniform(-180.0, 180.0, n_robots) * 1000) / 1000.0], 1)
    return World(ppm, half_extent_m, base.obstacles, robots, robot_size=robot_size)


def load_swarm_into_context(ctx, w, fiducial_return=0, chunk=20000):
    """load_into_context for worlds with many robots: same calls, same order, the outlines built with numpy.
    fiducial_return: Model::vis.fiducial_return of every robot (C4: 1)."""
    ctx.model_upsert(0, 0, 1.0)
    for mid in w.obstacle_ids:
        ctx.model_upsert(int(mid), int(mid), 0.5)
    for mid in w.robot_ids:
        ctx.model_upsert(int(mid), int(mid), 0.5, fiducial_return)
    ob = np.array(obstacle_polygons(w), np.int32).reshape(-1, 4, 2)
    polys = np.concatenate([ob, robot_boxes(w)], 0)
    models = np.concatenate([w.obstacle_ids, w.robot_ids])
    zmax = np.concatenate([np.full(len(w.obstacles), w.obstacle_height), np.full(len(w.robots), w.robot_size[2])])
    z = np.stack([np.zeros(len(polys)), zmax], 1)
    blocks = np.arange(len(polys), dtype=np.uint32)
    for layer in (0, 1):
        for lo in range(0, len(polys), chunk):
            hi = min(lo + chunk, len(polys))
            ctx.blocks_remap(blocks[lo:hi], models[lo:hi], layer, z[lo:hi], np.arange(hi - lo + 1, dtype=np.uint32) * 4,
                             np.ascontiguousarray(polys[lo:hi].reshape(-1, 2)))
    ctx.sync()
    return len(polys)


def fiducial_records(w, robots, robot_ids, layer, fiducial_return=1, range_max=8.0, range_max_id=5.0, fov=math.pi):
    """(targets, finders) for robots that each carry a fiducial and a fiducial finder at their centre (model_fiducial.cc:77-78
    defaults: range_max 8, range_max_id 5, fov pi, ignore_zloc 0); the finder looks out at half the body's height"""
    n = len(robots)
    a = robots[:, 2] * (math.pi / 180.0)
    t = np.zeros(n, rt.FID_TARGET_DTYPE)
    t["model"], t["fiducial_return"] = robot_ids, fiducial_return
    t["x"], t["y"], t["a"] = robots[:, 0], robots[:, 1], a
    t["sx"], t["sy"], t["sz"] = w.robot_size
    f = np.zeros(n, rt.FID_FINDER_DTYPE)
    f["finder"], f["layer"] = robot_ids, layer
    for k in ("x", "y", "a"):
        f[k] = f["o" + k] = t[k]
    f["oz"] = w.robot_size[2] / 2.0
    f["max_range_anon"], f["max_range_id"], f["fov"] = range_max, range_max_id, fov
    return t, f


def ranger_poses(robots, robot_ids, layer, z=0.0):
    """stg_rt_ranger_pose records: GetGlobalPose() + geom.pose of a ranger sitting at its robot's origin"""
    p = np.zeros(len(robots), rt.RANGER_POSE_DTYPE)
    p["finder"], p["layer"] = robot_ids, layer
    p["x"], p["y"], p["z"], p["a"] = robots[:, 0], robots[:, 1], z, robots[:, 2] * (math.pi / 180.0)
    return p


def local_box(size):
    """Block::pts of the default unit-square block after BlockGroup::CalcSize scaled it to the model's size (blockgroup.cc:84-112)"""
    sx, sy = size[0], size[1]
    return np.array([[(-0.5 - 0.0) * (sx / 1.0), (-0.5 - 0.0) * (sy / 1.0)], [(0.5 - 0.0) * (sx / 1.0), (-0.5 - 0.0) * (sy / 1.0)],
                     [(0.5 - 0.0) * (sx / 1.0), (0.5 - 0.0) * (sy / 1.0)], [(-0.5 - 0.0) * (sx / 1.0), (0.5 - 0.0) * (sy / 1.0)]])


def define_boxes(ctx, first_block, model_ids, size):
    """stg_rt_block_define for one box-shaped block per model (blocks first_block, first_block + 1, ...): from then on
    stg_rt_models_pose_update renders them from poses alone"""
    pts = local_box(size)
    for j, mid in enumerate(model_ids):
        ctx.block_define(int(first_block) + j, int(mid), 0.0, float(size[2]), pts)


def gposes(robots, z=0.0):
    """GetGlobalPose() + geom.pose of models whose geom.pose is zero: (n, 4) x, y, z, heading in radians"""
    return np.ascontiguousarray(np.stack([robots[:, 0], robots[:, 1], np.full(len(robots), z), robots[:, 2] * (math.pi / 180.0)], 1))
