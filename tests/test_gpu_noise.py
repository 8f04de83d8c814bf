"""Sensor noise of the ranger (model_ranger.cc:173-200, 232-249). The reference draws from libc rand()
in beam order, which a parallel kernel cannot replay, so parity is statistical: the device noise must
have the reference's distribution - the same discrete uniform U in {-1, -0.998 .. 0.998} for the
proportional and angular terms and N(0, range_noise_const) for the additive term - must leave
beams that reached range.max untouched, and must vanish exactly when the noise parameters are zero."""
import numpy as np
import pytest

from stage_b200 import rt

pytestmark = pytest.mark.gpu


def ring_world():
    """finder at the origin inside a big square room (walls 6..9 m away), part of the fan out of range"""
    ctx = rt.Context(50.0, -1, -1, 2, 2, max_models=8, max_blocks=8, max_cell_entries=1 << 16)
    ctx.model_upsert(1, 1, 0.5)
    ctx.model_upsert(2, 2)
    for layer in (0, 1):
        ctx.block_map(0, 1, layer, 0.0, 1.0, np.array([[-300, -300], [300, -300], [300, 300], [-300, 300]], np.int32))
    n_fans, n = 64, 1000
    fans = rt.make_fans(n_fans)
    fans["finder"], fans["n_samples"], fans["pred"], fans["ztest"], fans["layer"] = 2, n, rt.PRED_RANGER, 1, 0
    fans["angles"] = rt.ANGLES_RANGER
    fans["ox"], fans["oy"] = np.linspace(-1, 1, n_fans), np.linspace(1, -1, n_fans)
    fans["oz"], fans["oa"], fans["range"], fans["fov"] = 0.5, -np.pi, 7.5, 2 * np.pi  # corners (8.5 m) are out of range
    return ctx, fans


def noisy(ctx, fans, c=0.0, p=0.0, a=0.0, seed=0):
    nz = np.zeros(len(fans), rt.FAN_NOISE_DTYPE)
    nz["range_noise_const"], nz["range_noise"], nz["angle_noise"], nz["seed"] = c, p, a, seed
    out = rt.Outputs(int(fans["n_samples"].sum()), rt.WANT_RANGES | rt.WANT_HIT_MODEL)
    ctx.fans_submit_noisy(fans, nz, out)
    ctx.sync()
    return out


def test_zero_noise_is_the_noise_free_result_and_out_of_range_beams_are_untouched():
    ctx, fans = ring_world()
    clean = ctx.trace(fans, want=rt.WANT_RANGES | rt.WANT_HIT_MODEL)
    z = noisy(ctx, fans)
    assert np.array_equal(z.ranges.view(np.uint64), clean.ranges.view(np.uint64)) and np.array_equal(z.hit_model, clean.hit_model)
    far = clean.hit_model < 0
    assert 0.02 < far.mean() < 0.5
    g = noisy(ctx, fans, c=0.01, p=0.05)
    assert np.array_equal(g.ranges[far], clean.ranges[far]) and np.all(clean.ranges[far] == 7.5)
    assert np.all(g.ranges[~far] != clean.ranges[~far]) or np.mean(g.ranges[~far] != clean.ranges[~far]) > 0.99
    ctx.close()


def test_additive_noise_is_normal_with_the_given_variance():
    ctx, fans = ring_world()
    clean = ctx.trace(fans, want=rt.WANT_RANGES | rt.WANT_HIT_MODEL)
    hit = clean.hit_model >= 0
    var = 0.0004  # range_noise_const is a variance (generateGaussianNoise(variance))
    d = (noisy(ctx, fans, c=var).ranges - clean.ranges)[hit]
    n = len(d)
    assert n > 40000
    assert abs(d.mean()) < 5 * np.sqrt(var / n)
    assert abs(d.var() / var - 1) < 0.03
    s = np.sqrt(var)
    assert abs(np.mean(np.abs(d) < s) - 0.6827) < 0.01 and abs(np.mean(np.abs(d) < 2 * s) - 0.9545) < 0.006
    # a second launch draws fresh noise; a fresh context replays the same launches
    d2 = (noisy(ctx, fans, c=var).ranges - clean.ranges)[hit]
    assert np.mean(d2 != d) > 0.99 and abs(np.corrcoef(d, d2)[0, 1]) < 0.02
    ctx.close()
    ctx, fans = ring_world()
    clean2 = ctx.trace(fans, want=rt.WANT_RANGES | rt.WANT_HIT_MODEL)
    d3 = (noisy(ctx, fans, c=var).ranges - clean2.ranges)[hit]
    assert np.array_equal(d3, d)
    ctx.close()


def test_proportional_noise_is_the_references_discrete_uniform():
    ctx, fans = ring_world()
    clean = ctx.trace(fans, want=rt.WANT_RANGES | rt.WANT_HIT_MODEL)
    hit = clean.hit_model >= 0
    p = 0.05
    g = noisy(ctx, fans, p=p, seed=77)
    u = ((g.ranges - clean.ranges) / (clean.ranges * p))[hit]
    k = (u / 2 + 0.5) * 1000  # simpleNoise() = 2 * (k * 0.001 - 0.5), k = rand() % 1000
    assert np.max(np.abs(k - np.round(k))) < 1e-6 and k.min() > -0.5 and k.max() < 999.5
    counts = np.bincount(np.round(k).astype(int), minlength=1000)
    assert counts.min() > 0 and abs(u.mean() + 0.001) < 0.01 and abs(u.var() - 1 / 3) < 0.01
    chi2 = np.sum((counts - len(k) / 1000) ** 2 / (len(k) / 1000))
    assert chi2 < 1150  # 999 degrees of freedom: mean 999, sd 44.7
    ctx.close()


def test_angle_noise_jitters_headings_within_half_a_beam_spacing():
    ctx, fans = ring_world()
    clean = ctx.trace(fans, want=rt.WANT_RANGES | rt.WANT_HIT_MODEL)
    g = noisy(ctx, fans, a=1.0)
    hit = (clean.hit_model >= 0) & (g.hit_model >= 0)
    changed = g.ranges[hit] != clean.ranges[hit]
    assert changed.mean() > 0.9
    # half a beam spacing (0.5 * 2pi/999 rad) at <= 7.5 m moves the hit point by < 2.4 cm along a wall,
    # i.e. the range by about as much (grazing angles here stay above 35 degrees)
    assert np.max(np.abs(g.ranges[hit] - clean.ranges[hit])) < 0.08
    ctx.close()
