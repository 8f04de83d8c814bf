"""Regenerate the golden traces from the UNMODIFIED reference (needs oracle/_ref, i.e. a container
that has /root/reference; run oracle/build_ref.sh first). One subprocess per world because a
reference World can never be destroyed.

    python tests/golden/make_golden.py [name ...]

Each trace records every Block::Map / Block::UnMap / World::Raytrace / ModelRanger::Sensor::Update /
ModelFiducial::Update / ModelBlobfinder::Update / Model::TestCollision / Model::AppendTouchingModels the reference performed
(format: oracle/trace_format.h), compressed with xz.

everything_40 is worlds/everything.world (hospital section bitmap, pioneers with 16 sonars, SICK lasers, fiducial
finders, a gripper, bumpers and a blobfinder, all `alwayson`) minus its one `camera(...)` line: the camera model
needs OpenGL and is not part of the headless build (oracle/build_ref.sh).

bumper_150 is not one of the reference's shipped worlds: none of those makes robots collide within
a few hundred ticks, so Model::TestCollision would only ever be seen answering "nothing". It is a
walled arena with a rotated pillar, a bridge the robots fit under (z separation), a model without
obstacle_return, and 14 position models - some with a child model sticking out in front - driven
straight ahead (a few on arcs) by ModelPosition::SetSpeed until they run into things and each other.
"""
import lzma
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
WORLDS = [("simple.world", 120, "simple_120"),     # BASELINE config 1: Pioneer + SICK laser, wander ctrl
          ("fasr.world", 300, "fasr_300"),         # BASELINE config 2: rangers + fiducials, threads 2
          ("lsp_test.world", 20, "lsp_test_20"),   # the Player-plugin test world: fiducial + blobfinder
          ("pioneer_flocking.world", 12, "pioneer_flocking_12"),  # 100 pioneers, 8 sonars + fiducial finder each, threads 4
          ("large.world", 20, "large_20")]  # the widest extent the reference ships: a 1000 m x 700 m floorplan at 2 cm
                                            # (50 x 36 superregions, 1.9 M cell entries per layer), 10 laser robots
# lsp_test.world has no controllers: switch its sensors on the way a Player client would
SUBSCRIBE = {"lsp_test.world": ["r0.fiducial:0", "r1.fiducial:0", "r0.blobfinder:0", "r1.blobfinder:0"]}



def bumper_world_text():
    import numpy as np
    rng = np.random.RandomState(3)
    lines = ["resolution 0.02", "interval_sim 100", "threads 1", "quit_time 3600", ""]

    def box(name, x, y, z, a, sx, sy, sz, extra=""):
        lines.append('model ( name "%s" pose [ %.3f %.3f %.3f %.3f ] size [ %.3f %.3f %.3f ] %s )' % (name, x, y, z, a, sx, sy, sz, extra))
    box("wn", 0, 4.05, 0, 0, 8.2, 0.1, 0.5)
    box("ws", 0, -4.05, 0, 0, 8.2, 0.1, 0.5)
    box("we", 4.05, 0, 0, 0, 0.1, 8.2, 0.5)
    box("ww", -4.05, 0, 0, 0, 0.1, 8.2, 0.5)
    box("pillar", 1.0, 1.0, 0, 30, 0.6, 0.6, 0.5)
    box("bridge", -1.5, 0, 1.0, 0, 0.3, 5.0, 0.2)
    box("ghost", 0, -1.5, 0, 0, 1.0, 1.0, 0.5, "obstacle_return 0")
    for i in range(BUMPERS):
        x, y = rng.uniform(-3.3, 3.3, 2)
        a = rng.uniform(-180, 180)
        child = 'model ( name "arm%d" pose [ 0.25 0 0 0 ] size [ 0.3 0.08 0.1 ] )' % i if i % 3 == 0 else ""
        lines.append('position ( name "r%d" pose [ %.3f %.3f 0 %.3f ] size [ 0.4 0.3 0.3 ] %s )' % (i, x, y, a, child))
    return "\n".join(lines) + "\n"


def charger_world_text():
    """charger_150, not a shipped world either: Model::AppendTouchingModels only runs for models that give charge
    (Model::UpdateCharge), and in fasr.world nothing reaches a charger within a few hundred ticks. Four
    always-on chargers - a floor pad and a thin strip the robots drive across (no obstacle_return), a dock
    built into the east wall with a related child model, a post standing in the rotated pillar and the ghost -
    and the driven robots of the bumper world, some carrying an arm."""
    import numpy as np
    rng = np.random.RandomState(5)
    lines = ["resolution 0.02", "interval_sim 100", "threads 1", "quit_time 3600", ""]

    def box(name, x, y, z, a, sx, sy, sz, extra=""):
        lines.append('model ( name "%s" pose [ %.3f %.3f %.3f %.3f ] size [ %.3f %.3f %.3f ] %s )' % (name, x, y, z, a, sx, sy, sz, extra))
    box("wn", 0, 4.05, 0, 0, 8.2, 0.1, 0.5)
    box("ws", 0, -4.05, 0, 0, 8.2, 0.1, 0.5)
    box("we", 4.05, 0, 0, 0, 0.1, 8.2, 0.5)
    box("ww", -4.05, 0, 0, 0, 0.1, 8.2, 0.5)
    box("pillar", 1.0, 1.0, 0, 30, 0.6, 0.6, 0.5)
    box("ghost", 0.8, 0.8, 0, 0, 1.0, 1.0, 0.5, "obstacle_return 0")
    give = "joules -1 give_watts 1000 alwayson 1"
    box("pad", -1.5, -1.0, 0, 10, 2.4, 2.0, 0.02, "obstacle_return 0 " + give)
    box("strip", 0, 2.4, 0, -4, 7.0, 0.06, 0.02, "obstacle_return 0 " + give)
    box("dock", 3.95, -1.0, 0, 0, 0.2, 0.5, 0.1, give + ' model ( name "dockside" pose [ -0.1 0 0 0 ] size [ 0.1 0.7 0.2 ] )')
    box("post", 1.0, 1.0, 0, 0, 0.9, 0.1, 0.6, "obstacle_return 0 " + give)
    for i in range(BUMPERS):
        x, y = rng.uniform(-3.3, 3.3, 2)
        a = rng.uniform(-180, 180)
        child = 'model ( name "arm%d" pose [ 0.25 0 0 0 ] size [ 0.3 0.08 0.1 ] )' % i if i % 3 == 0 else ""
        pack = "kjoules 10 take_watts 500" if i % 2 == 0 else ""
        lines.append('position ( name "r%d" pose [ %.3f %.3f 0 %.3f ] size [ 0.4 0.3 0.3 ] %s %s )' % (i, x, y, a, pack, child))
    return "\n".join(lines) + "\n"


BUMPERS = 14
BUMPER_CHILD = r"""
import sys, os
sys.path.insert(0, %r)
import oracle_lib
w = oracle_lib.RefWorld(text=open(sys.argv[1]).read())
for i in range(int(sys.argv[4])):
    w.subscribe("r%%d" %% i)
    w.set_speed("r%%d" %% i, 0.5, 0.0, 0.3 if i %% 4 == 0 else 0.0)
w.update(int(sys.argv[2]))
w.trace_end(sys.argv[3])
sys.stdout.flush()
os._exit(0)
"""

CHILD = r"""
import sys, os
sys.path.insert(0, %r)
import oracle_lib
w = oracle_lib.RefWorld(path=os.path.join(oracle_lib.REF, "worlds", sys.argv[1]))
for name in sys.argv[4:]:
    w.subscribe(name)
w.update(int(sys.argv[2]))
w.trace_end(sys.argv[3])
sys.stdout.flush()
os._exit(0)
"""

def everything_world_text():
    sys.path.insert(0, TESTS)
    import oracle_lib
    src = open(os.path.join(oracle_lib.REF, "worlds", "everything.world")).read()
    return "\n".join(line for line in src.split("\n") if not line.strip().startswith("camera(")) + "\n"


TEXT_CHILD = r"""
import sys, os
sys.path.insert(0, %r)
import oracle_lib
w = oracle_lib.RefWorld(text=open(sys.argv[1]).read(), path_hint=os.path.join(oracle_lib.REF, "worlds", "generated.world"))
w.update(int(sys.argv[2]))
w.trace_end(sys.argv[3])
sys.stdout.flush()
os._exit(0)
"""

if __name__ == "__main__":
    for world, ticks, name in WORLDS + [("everything", 40, "everything_40"), ("bumper", 150, "bumper_150"),
                                       ("charger", 150, "charger_150")]:
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        raw = os.path.join(HERE, name + ".trc")
        if world == "everything":
            wf = os.path.join(HERE, "everything.world.tmp")
            with open(wf, "w") as f:
                f.write(everything_world_text())
            subprocess.run([sys.executable, "-c", TEXT_CHILD % TESTS, wf, str(ticks), raw], check=True, stdout=subprocess.DEVNULL)
            os.remove(wf)
        elif world in ("bumper", "charger"):
            wf = os.path.join(HERE, world + ".world.tmp")
            with open(wf, "w") as f:
                f.write(bumper_world_text() if world == "bumper" else charger_world_text())
            subprocess.run([sys.executable, "-c", BUMPER_CHILD % TESTS, wf, str(ticks), raw, str(BUMPERS)], check=True,
                           stdout=subprocess.DEVNULL)
            os.remove(wf)
        else:
            subprocess.run([sys.executable, "-c", CHILD % TESTS, world, str(ticks), raw] + SUBSCRIBE.get(world, []), check=True,
                           stdout=subprocess.DEVNULL)
        with open(raw, "rb") as f, lzma.open(raw + ".xz", "wb", preset=9) as g:
            g.write(f.read())
        os.remove(raw)
        print(name, os.path.getsize(raw + ".xz"), "bytes")
