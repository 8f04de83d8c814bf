"""The libstage binding without a GPU: STAGE_GPU=3 makes integration/stage_gpu_shim.cc take every answer from libstage's
own loops on the side (their libc rand() draws undone), hold it back until the sync point and deliver it there exactly
as the GPU path does - deferred ranges / intensities / fiducials, stg_rt_ranger_rand_advance for the skipped draws. Run
against the same process image falling back to plain libstage, the dumps of everything the controllers read must be
byte-identical: that is the binding's control flow, its deferred delivery and the rand() replay proven on three shipped
worlds with their real controllers; the GPU tier (tests/test_gpu_libstage.py) adds the C ABI underneath."""
import os

import pytest

import stage_dump
from test_gpu_libstage import REF, RUN, SHIM, have_binding, stage_run


@pytest.mark.skipif(not have_binding(), reason="integration/_build or oracle/_ref not built (integration/build.sh)")
@pytest.mark.parametrize("world,ticks", [("simple.world", 1000), ("fasr.world", 300), ("pioneer_flocking.world", 40), ("lsp_test.world", 30)])
def test_emulated_binding_equals_plain_libstage(world, ticks, tmp_path):
    a, b = str(tmp_path / "cpu.bin"), str(tmp_path / "emu.bin")
    sub = world == "lsp_test.world"   # no controllers there: every ranger, fiducial finder and blobfinder is subscribed to instead
    cpu, _ = stage_run(world, ticks, a, fallback=True, fiducials_on_main=True, subscribe_all=sub)
    emu, err = stage_run(world, ticks, b, gpu=3, fiducials_on_main=True, subscribe_all=sub)
    assert cpu["gpu"]["state"] == "fell back to cpu" and emu["gpu"]["state"] == "emulated on the cpu", err[-1000:]
    off = stage_dump.first_difference(a, b)
    assert off is None, "dumps differ at byte %d: %r" % (off, stage_dump.describe_first_difference(a, b))
    assert emu["gpu"]["fans"] > 0 and emu["ranger_samples"] == cpu["ranger_samples"]
    if world == "pioneer_flocking.world":
        assert emu["fiducial_records"] == cpu["fiducial_records"] > 10000
    if world == "lsp_test.world":   # blobfinders: held back and delivered at the sync point like the rest
        assert emu["blob_records"] == cpu["blob_records"] > 30 and emu["gpu"]["blob_updates"] == 2 * ticks


@pytest.mark.skipif(not have_binding(), reason="integration/_build or oracle/_ref not built (integration/build.sh)")
def test_binding_without_a_gpu_falls_back_to_plain_libstage(tmp_path):
    """STAGE_GPU=1 where stg_rt_create fails (no CUDA device here): one warning, then the reference's own behaviour."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    a, b = str(tmp_path / "plain.bin"), str(tmp_path / "nogpu.bin")
    stage_run("simple.world", 200, a, fallback=True)
    res, err = stage_run("simple.world", 200, b, gpu=1)
    assert res["gpu"]["state"] == "fell back to cpu" and "continuing on the CPU path" in err
    assert stage_dump.first_difference(a, b) is None
