"""The compiled libstage binding (integration/stage_gpu_shim.cc -> integration/_build/libstage_gpu.so) under the
UNMODIFIED reference with its real controllers: `stage_run` (integration/stage_run.cc, a headless front end like
`stage -g`) runs a shipped world twice - plain, and with LD_PRELOAD=libstage_gpu.so STAGE_GPU=1 - and the dumps of
everything a controller can read (every ranger double, every fiducial record, every model pose, tick by tick) must be
byte-identical: the robots' trajectories are driven by the sensor data (wander.cc, fasr.cc) and by the libc random
stream the skipped beam loops would have advanced (stg_rt_ranger_rand_advance), so one wrong double or one missing
rand() draw shows within a few ticks. The CPU side of the comparison is the same process image with the binding
loaded but told to fall back at its first sensor update (see stage_run below for why). STAGE_GPU=2 makes the binding run libstage's own loops too and compare every
answer in place - the check that also works where the reference itself is not reproducible from run to run (fiducial
finders on worker threads racing ModelPosition::Move, pointer-ordered move lists), with the ABI called from four threads."""
import json
import os
import subprocess

import pytest

import stage_dump

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
BUILD = os.path.join(ROOT, "integration", "_build")
RUN, SHIM = os.path.join(BUILD, "stage_run"), os.path.join(BUILD, "libstage_gpu.so")

gpu = pytest.mark.gpu


def have_binding():
    return os.path.exists(RUN) and os.path.exists(SHIM) and os.path.exists(os.path.join(REF, "libstage_ref.so"))


def stage_run(world, ticks, dump=None, gpu=0, fiducials_on_main=False, timeout=600, fallback=False, subscribe_all=False):
    """gpu: 0 = no binding loaded, 1 = through the GPU, 2 = GPU verified in place, 3 = emulated on the CPU.
    fallback: the binding is loaded and behaves like a GPU run until its first sensor update, then leaves everything to
    libstage: the reference itself, with the heap layout of a GPU run. That matters because Stage orders its work by
    POINTER - World::models, the controllers' start-up (world.cc:459-465) and ModelPosition::Move's list are std::sets
    of pointers - so two runs agree on who draws which random() number only if their models sit at the same addresses."""
    env = dict(os.environ)
    env["STAGEPATH"] = "%s/plugins:%s/share/stage" % (REF, REF)
    env.pop("LD_PRELOAD", None)
    if fallback:
        gpu = 1
        env["STAGE_GPU_FORCE_FALLBACK"] = "1"
    if gpu:
        env["LD_PRELOAD"] = SHIM
        env["STAGE_GPU"] = str(gpu)
    if fiducials_on_main:
        env["STAGE_FIDUCIALS_ON_MAIN"] = "1"
    if subscribe_all:   # worlds without controllers: every ranger / fiducial finder / blobfinder updates every tick
        env["STAGE_RUN_SUBSCRIBE_ALL"] = "1"
    cmd = [RUN, os.path.join("worlds", world), str(ticks)] + ([dump] if dump else [])
    r = subprocess.run(cmd, cwd=REF, env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("stage_run: ")][-1]
    return json.loads(line[len("stage_run: "):]), r.stderr


def compare(world, ticks, tmp_path, fiducials_on_main=False, mode=1, subscribe_all=False):
    a, b = str(tmp_path / "cpu.bin"), str(tmp_path / "gpu.bin")
    cpu, _ = stage_run(world, ticks, a, fallback=True, fiducials_on_main=fiducials_on_main, subscribe_all=subscribe_all)
    assert cpu["gpu"]["state"] == "fell back to cpu"
    gpu, err = stage_run(world, ticks, b, gpu=mode, fiducials_on_main=fiducials_on_main, subscribe_all=subscribe_all)
    assert gpu["gpu"]["state"] == ("gpu" if mode == 1 else "emulated on the cpu"), err[-2000:]
    off = stage_dump.first_difference(a, b)
    assert off is None, "dumps differ at byte %d: %r" % (off, stage_dump.describe_first_difference(a, b))
    assert cpu["ticks"] == gpu["ticks"] == ticks
    return cpu, gpu


@gpu
@pytest.mark.skipif(not have_binding(), reason="integration/_build or oracle/_ref not built (integration/build.sh)")
def test_simple_world_1000_ticks_identical_with_the_wander_controller(tmp_path):
    """BASELINE config 1: worlds/simple.world, two Pioneers with 180-beam SICK lasers driven by examples/ctrl/wander.cc"""
    cpu, gpu = compare("simple.world", 1000, tmp_path)
    assert gpu["gpu"]["fans"] == 2 * 1000 and gpu["ranger_samples"] == cpu["ranger_samples"] == 2 * 180 * 1000
    print("C1 simple.world, 1000 ticks: reference %.0f ticks/s, through libstage_gpu.so %.0f ticks/s (host: %d cpus)"
          % (cpu["ticks_per_s"], gpu["ticks_per_s"], os.cpu_count()))


@gpu
@pytest.mark.skipif(not have_binding(), reason="integration/_build or oracle/_ref not built (integration/build.sh)")
def test_fasr_world_300_ticks_identical_with_the_fasr_controllers(tmp_path):
    """BASELINE config 2: worlds/fasr.world, 20 robots with 32-beam lasers + fiducial finders, ctrl fasr / source / sink,
    two worker threads; the finders stay on the main thread's queue so that the reference is reproducible at all."""
    cpu, gpu = compare("fasr.world", 300, tmp_path, fiducials_on_main=True)
    assert gpu["gpu"]["fans"] > 300 and gpu["ranger_samples"] == cpu["ranger_samples"] > 0
    print("C2 fasr.world, 300 ticks: reference %.0f ticks/s, through libstage_gpu.so %.0f ticks/s" % (cpu["ticks_per_s"], gpu["ticks_per_s"]))


@gpu
@pytest.mark.skipif(not have_binding(), reason="integration/_build or oracle/_ref not built (integration/build.sh)")
def test_pioneer_flocking_60_ticks_identical_with_the_flocking_controller(tmp_path):
    """worlds/pioneer_flocking.world: 100 robots steering by what their 8 sonars and their fiducial finder report (800
    one-beam fans and 100 finder updates per tick, ~800 fiducial records per tick): BASELINE config 4's shape as the
    reference itself runs it. Fiducial range and bearing are bit-exact (STG_RT_CFG_HOST_EXACT_FIDUCIALS)."""
    cpu, gpu = compare("pioneer_flocking.world", 60, tmp_path, fiducials_on_main=True)
    assert gpu["gpu"]["fans"] == 800 * 60 and gpu["gpu"]["fiducial_updates"] == 100 * 60
    assert gpu["fiducial_records"] == cpu["fiducial_records"] > 30000
    print("pioneer_flocking.world, 60 ticks: reference %.0f ticks/s, through libstage_gpu.so %.0f ticks/s" % (cpu["ticks_per_s"], gpu["ticks_per_s"]))


@gpu
@pytest.mark.skipif(not have_binding(), reason="integration/_build or oracle/_ref not built (integration/build.sh)")
def test_lsp_test_world_blobfinders_lasers_and_fiducials_identical(tmp_path):
    """worlds/lsp_test.world (the world the reference's own libstageplugin test pins its known answers on: two Pioneers
    with sonar rings, SICK lasers, fiducial finders and six-channel blobfinders in the cave), every sensor subscribed to:
    ModelBlobfinder::Update goes through stg_rt_blobs_submit - colours, column spans, rows and ranges of every blob byte for
    byte as libstage computes them - and, verified in place, every answer of every update agrees."""
    cpu, gpu = compare("lsp_test.world", 30, tmp_path, fiducials_on_main=True, subscribe_all=True)
    assert gpu["blob_records"] == cpu["blob_records"] > 30 and gpu["gpu"]["blob_updates"] == 2 * 30
    assert gpu["gpu"]["fans"] > 0 and gpu["fiducial_records"] == cpu["fiducial_records"] > 0
    res, err = stage_run("lsp_test.world", 30, gpu=2, fiducials_on_main=True, subscribe_all=True)
    g = res["gpu"]
    assert g["state"].startswith("gpu, verified"), err[-2000:]
    assert g["mismatched_blob_updates"] == 0 and g["verified_blobs"] > 30 and g["mismatched_beams"] == 0 and g["mismatched_fiducial_updates"] == 0, (g, err[-1500:])


@gpu
@pytest.mark.skipif(not have_binding(), reason="integration/_build or oracle/_ref not built (integration/build.sh)")
@pytest.mark.parametrize("world,ticks,on_main", [("pioneer_flocking.world", 60, True), ("pioneer_flocking.world", 60, False),
                                                 ("fasr.world", 300, False), ("simple.world", 300, False)])
def test_every_answer_verified_in_place_with_worker_threads(world, ticks, on_main):
    """STAGE_GPU=2: libstage's own loops and the GPU answer every sensor update of the run; the binding compares them in
    place. pioneer_flocking.world as shipped runs its 100 fiducial finders on four worker threads: the ABI is called from
    five threads at once and every ranger beam must still agree; the finders' own answers cannot be checked there,
    because libstage reads the poses its main thread is moving at that moment (world.cc:647 vs model_fiducial.cc:124-127:
    the finder's and the targets' poses differ between its pass and ours) - they are with the finders on the main thread."""
    res, err = stage_run(world, ticks, gpu=2, fiducials_on_main=on_main)
    g = res["gpu"]
    assert g["state"].startswith("gpu, verified"), err[-2000:]
    assert g["mismatched_beams"] == 0, (g, err[-1500:])
    assert g["verified_beams"] > 0 and g["fans"] > 0
    if on_main or world != "pioneer_flocking.world":
        assert g["mismatched_fiducial_updates"] == 0, (g, err[-1500:])
    if world == "pioneer_flocking.world":
        assert g["fiducial_updates"] == 100 * ticks and g["verified_beams"] == 800 * ticks
        if on_main:
            assert g["verified_fiducials"] > 30000
