"""Extract the pixel-space polygons the UNMODIFIED reference produced when it loaded
worlds/bitmaps/cave.png as the 16 m floorplan of worlds/simple.world (resolution 0.02), from the
golden trace tests/golden/simple_120.trc.xz (the Block::Map calls of model "cave" on layer 0), into
stage_b200/data/cave_polygons.npz: `first` (n_polygons + 1 offsets) and `xy` (int32 vertices, cells,
the tile centred on the origin). stage_b200/worldgen.make_cave_world tiles them.

    python tools/extract_cave_polygons.py
"""
import lzma
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import trace_io  # noqa: E402

if __name__ == "__main__":
    with tempfile.NamedTemporaryFile(suffix=".trc") as f:
        f.write(lzma.open(os.path.join(ROOT, "tests", "golden", "simple_120.trc.xz")).read())
        f.flush()
        tr = trace_io.load(f.name)
    cave = [int(m["id"]) for m in tr.models if m["type"] == b"model" and m["id"] != 0]
    assert len(cave) == 1
    polys = []
    for ev in tr.events:
        if ev["type"] == trace_io.EV_TICK:
            break
        if ev["type"] == trace_io.EV_MAP and ev["model"] == cave[0] and ev["layer"] == 0:
            polys.append(np.array(tr.polygon(ev), np.int32))
    first = np.zeros(len(polys) + 1, np.uint32)
    first[1:] = np.cumsum([len(p) for p in polys])
    out = os.path.join(ROOT, "stage_b200", "data", "cave_polygons.npz")
    np.savez_compressed(out, first=first, xy=np.concatenate(polys))
    print(len(polys), "polygons,", int(first[-1]), "vertices ->", out, os.path.getsize(out), "bytes")
