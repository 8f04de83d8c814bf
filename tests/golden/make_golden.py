"""Regenerate the golden traces from the UNMODIFIED reference (needs oracle/_ref, i.e. a container
that has /root/reference; run oracle/build_ref.sh first). One subprocess per world because a
reference World can never be destroyed.

    python tests/golden/make_golden.py

Each trace records every Block::Map / Block::UnMap / World::Raytrace / ModelRanger::Sensor::Update
the reference performed (format: oracle/trace_format.h), compressed with xz.
"""
import lzma
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
WORLDS = [("simple.world", 120, "simple_120"),     # BASELINE config 1: Pioneer + SICK laser, wander ctrl
          ("fasr.world", 300, "fasr_300"),         # BASELINE config 2: rangers + fiducials, threads 2
          ("lsp_test.world", 20, "lsp_test_20")]   # the Player-plugin test world: fiducial + blobfinder
# lsp_test.world has no controllers: switch its sensors on the way a Player client would
SUBSCRIBE = {"lsp_test.world": ["r0.fiducial:0", "r1.fiducial:0", "r0.blobfinder:0", "r1.blobfinder:0"]}

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

if __name__ == "__main__":
    for world, ticks, name in WORLDS:
        raw = os.path.join(HERE, name + ".trc")
        subprocess.run([sys.executable, "-c", CHILD % TESTS, world, str(ticks), raw] + SUBSCRIBE.get(world, []), check=True,
                       stdout=subprocess.DEVNULL)
        with open(raw, "rb") as f, lzma.open(raw + ".xz", "wb", preset=9) as g:
            g.write(f.read())
        os.remove(raw)
        print(name, os.path.getsize(raw + ".xz"), "bytes")
