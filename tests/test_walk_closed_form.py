"""The closed forms the march uses to leave a region, or to jump ahead inside one, without stepping cell by cell
(stage_b200/csrc/rt_walk.h) against the reference's step-by-step loop (world.cc:823-882 in run form): every slope of a
short ray from a lattice of start cells with every error term the walk can hold, and three million random rays up to
the longest the closed forms serve. Host build of the very functions the kernels inline; no GPU."""
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))


def test_region_exit_and_jumps_equal_the_step_by_step_walk():
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "walk_check")
        subprocess.run(["g++", "-O2", "-Wall", "-Werror", "-o", exe, os.path.join(HERE, "walk_check.cc")], check=True)
        r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("ok "), r.stdout[-2000:] + r.stderr[-2000:]
    assert int(r.stdout.split()[1]) > 100_000_000
