"""The per-beam headings kernel against the reference's serial recurrence (model_ranger.cc:237-241,
254: the running angle goes through a `float` every beam). The device cuts the recurrence into 32
chunks per fan and accepts a result only when every chunk's exit state equals the next chunk's
entry state; these fans are chosen to break its guesses: sweeps through 0, through |a| = 1/2, 1, 2
(float binade borders), tiny and huge spacings, sample counts around the chunking's edges."""
import numpy as np
import pytest

import oracle_lib
from stage_b200 import rt

pytestmark = pytest.mark.gpu


def host_headings(oa, fov, n):
    """the recurrence, in numpy scalars (float32 round trip each beam)"""
    incr = np.float64(fov) / np.float64(max(n - 1, 1) if n - 1 > 1 else 1)
    a = np.float64(oa)
    out = np.empty(n)
    for t in range(n):
        fa = np.float32(a)
        out[t] = np.float64(fa)
        a = np.float64(fa) + incr
    return out


def test_ranger_headings_bit_exact_on_adversarial_fans():
    rng = np.random.RandomState(7)
    cases = []
    for n in (1, 2, 3, 31, 32, 33, 63, 64, 65, 180, 361, 1080, 1081, 4096, 20000):
        for oa, fov in ((-2.3561944901923448, 4.71238898038469), (0.0, 3.141592653589793), (-1e-3, 2e-3), (3.1, 6.2),
                        (-3.141592653589793, 6.283185307179586), (0.4999, 1.1), (1.9, 0.3), (-0.26, 0.52)):
            cases.append((oa, fov, n))
    for _ in range(300):
        cases.append((rng.uniform(-np.pi, np.pi), rng.uniform(1e-4, 2 * np.pi), int(rng.choice([8, 16, 180, 270, 683, 1080]))))
    for _ in range(40):  # spacings near or below the float grid: the running angle stalls or crawls
        cases.append((rng.uniform(-3, 3), rng.uniform(1e-9, 1e-5), int(rng.choice([100, 1080]))))
    fans = rt.make_fans(len(cases))
    fans["finder"], fans["pred"], fans["angles"], fans["range"] = 0, rt.PRED_RANGER, rt.ANGLES_RANGER, 1.0
    fans["oa"], fans["fov"], fans["n_samples"] = [c[0] for c in cases], [c[1] for c in cases], [c[2] for c in cases]
    ctx = rt.Context(50.0, -1, -1, 2, 2, max_models=4, max_blocks=4, max_cell_entries=1 << 10)
    ctx.model_upsert(0, 0)
    got = ctx.debug_headings(fans)
    ctx.close()
    pos = 0
    for oa, fov, n in cases:
        want = host_headings(oa, fov, n)
        assert np.array_equal(oracle_lib.ranger_headings(oa, fov, n), want)  # the oracle's own statement of the rule
        g = got[pos:pos + n]
        bad = np.nonzero(g.view(np.uint64) != want.view(np.uint64))[0]
        assert len(bad) == 0, "fan oa=%r fov=%r n=%d: first mismatch at beam %d: %r != %r" % (oa, fov, n, bad[0], g[bad[0]], want[bad[0]])
        pos += n


def test_even_and_explicit_headings():
    """World::Raytrace(fan) spacing (world.cc:731-743) and caller-supplied headings"""
    fans = rt.make_fans(3)
    fans["finder"], fans["range"] = 0, 1.0
    fans["angles"] = [rt.ANGLES_EVEN_FOV, rt.ANGLES_EXPLICIT, rt.ANGLES_EVEN_FOV]
    fans["oa"], fans["fov"], fans["n_samples"] = [0.3, 0.0, -2.0], [1.0471975511965976, 0.0, 3.0], [80, 37, 5]
    fans["first_heading"] = [0, 0, 0]
    explicit = np.linspace(-3, 3, 37)
    ctx = rt.Context(50.0, -1, -1, 2, 2, max_models=4, max_blocks=4, max_cell_entries=1 << 10)
    ctx.model_upsert(0, 0)
    got = ctx.debug_headings(fans, explicit)
    ctx.close()
    def even(oa, fov, n):
        return np.array([(np.float64(s) * fov / np.float64(n - 1)) - (fov / 2.0 - oa) for s in range(n)])
    want = np.concatenate([even(0.3, 1.0471975511965976, 80), explicit, even(-2.0, 3.0, 5)])
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


def test_short_fans_take_the_thread_per_fan_kernel():
    """batches of short fans (sonars: one beam each) are handled one thread per fan; same rule, same bits"""
    rng = np.random.RandomState(11)
    n_fans = 5000
    ns = rng.choice([1, 1, 1, 2, 3, 5, 8], n_fans)
    fans = rt.make_fans(n_fans)
    fans["finder"], fans["pred"], fans["angles"], fans["range"] = 0, rt.PRED_RANGER, rt.ANGLES_RANGER, 1.0
    fans["oa"], fans["fov"], fans["n_samples"] = rng.uniform(-np.pi, np.pi, n_fans), rng.uniform(0.0, 1.0, n_fans), ns
    ctx = rt.Context(50.0, -1, -1, 2, 2, max_models=4, max_blocks=4, max_cell_entries=1 << 10)
    ctx.model_upsert(0, 0)
    got = ctx.debug_headings(fans)
    ctx.close()
    want = np.concatenate([host_headings(f["oa"], f["fov"], int(f["n_samples"])) for f in fans])
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
