"""Regenerate tests/golden/local_to_pixels.npz from the UNMODIFIED reference (needs oracle/_ref, i.e. a container that
has /root/reference; run oracle/build_ref.sh first). One subprocess per world (a reference World can never be destroyed).

    python tests/golden/make_l2p_golden.py

For several shipped worlds, after a number of ticks under their real controllers (so that robots stand at arbitrary
angles), every block of every model as the reference holds it: its polygon in model coordinates (Block::pts after
BlockGroup::CalcSize), Block::local_z, the model's frame GetGlobalPose() + geom.pose - the inputs of
stg_rt_block_define / stg_rt_models_pose_update - and the pixels Model::LocalToPixels (model.cc:474-491) gives for
them, the reference's own answer. tests/test_gpu_pose.py holds the device's pose path to these, bit for bit.
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
WORLDS = [("simple.world", 137), ("fasr.world", 211), ("pioneer_flocking.world", 23), ("everything_nocam.world", 17), ("large.world", 9),
          ("SFU.world", 5), ("autolab.world", 5)]


def one(world, ticks, out):
    sys.path.insert(0, TESTS)
    import oracle_lib
    src = os.path.join(oracle_lib.REF, "worlds", world)
    if world == "everything_nocam.world":   # the camera model needs OpenGL: not part of the headless build
        text = "".join(l for l in open(os.path.join(oracle_lib.REF, "worlds", "everything.world")) if "camera" not in l)
        w = oracle_lib.RefWorld(text=text, record=False, path_hint=os.path.join(oracle_lib.REF, "worlds", "everything.world"))
    else:
        w = oracle_lib.RefWorld(path=src, record=False)
    w.update(ticks)
    g = w.block_geometry()
    np.savez(out, ppm=w.ppm, **g)
    os._exit(0)


def main():
    parts = {}
    for world, ticks in WORLDS:
        tmp = "/tmp/l2p_%s.npz" % world
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", world, str(ticks), tmp], capture_output=True, text=True)
        if r.returncode != 0 or not os.path.exists(tmp):
            print("skipped", world, r.stderr[-300:])
            continue
        d = np.load(tmp)
        key = world.split(".")[0]
        for k in d.files:
            parts[key + "/" + k] = d[k]
        print(world, "blocks", len(d["model"]), "points", len(d["pts"]))
    np.savez_compressed(os.path.join(HERE, "local_to_pixels.npz"), **parts)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        one(sys.argv[2], int(sys.argv[3]), sys.argv[4])
    else:
        main()
