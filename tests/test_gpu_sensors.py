"""K3 parity: the fiducial-finder and blobfinder post-processing kernels against what the unmodified
reference produced (golden traces of worlds/fasr.world and worlds/lsp_test.world)."""
import os

import numpy as np
import pytest

import replay
import trace_io
from stage_b200 import rt

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module", params=["fasr_300", "lsp_test_20"])
def run(request):
    tr = trace_io.load(os.path.join(GOLDEN, request.param + ".trc.xz"))
    g = replay.GpuReplay(tr, strict=False, want=rt.WANT_RANGES | rt.WANT_HIT_MODEL)
    yield request.param, tr, g
    g.ctx.close()


def test_fiducials_match_the_reference(run):
    name, tr, g = run
    assert len(g.fiducials) == (40 if name == "lsp_test_20" else len([e for e in tr.events if e["type"] == trace_io.EV_FIDUCIAL]))
    seen = 0
    for ev, d, out, cnt in g.fiducials:
        ref = d["found"]
        assert cnt == len(ref), (ev, cnt, ref)
        for o, r in zip(out, ref):
            assert int(d["targets"][o["target"], 0]) == int(r[0])   # which model (Fiducial::mod), same order
            assert o["id"] == int(r[1])
            assert abs(o["range"] - r[2]) <= 1e-9 and abs(o["bearing"] - r[3]) <= 1e-9
            assert np.allclose(o["geom"], r[4:8], rtol=0, atol=1e-12)
            assert np.array_equal(o["pose"], r[8:12])
            seen += 1
    assert seen > (10 if name == "lsp_test_20" else 100)


def test_blobs_match_the_reference(run):
    name, tr, g = run
    if name != "lsp_test_20":
        pytest.skip("no blobfinder in this world")
    colors = {int(m["id"]): m["color"][:3] for m in tr.models}
    assert len(g.blobs) == 40
    for ev, d, out, cnt in g.blobs:
        ref = d["blobs"]
        assert cnt == len(ref)
        for o, r in zip(out, ref):
            assert np.array_equal(colors[int(o["model"])], r[:3])
            assert (o["left"], o["top"], o["right"], o["bottom"]) == tuple(int(x) for x in r[3:7])
            assert abs(o["range"] - r[7]) <= 1e-5
    assert sum(c for *_, c in g.blobs) >= 60
