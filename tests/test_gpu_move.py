"""ModelPosition::Move for all movers of a tick at once (stg_rt_models_move) against the unmodified
reference's one-by-one loop, replayed from the bumper_150 golden trace: per tick the reference
re-mapped each of the 14 robots (some with a child model), tested it, and put it back if it hit a
wall, the pillar or another robot - 2100 moves, 1728 of them undone. The batched call must stall
exactly the robots the reference stalled, every tick, and leave the grid - cell lists in order -
exactly as the reference's sequence of Map / UnMap calls leaves it."""
import lzma
import os
import time

import numpy as np
import pytest

import oracle_lib
import replay
import trace_io
from stage_b200 import rt

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_batched_move_equals_the_references_loop(tmp_path):
    dst = tmp_path / "bumper.trc"
    with lzma.open(os.path.join(GOLDEN, "bumper_150.trc.xz"), "rb") as f:
        dst.write_bytes(f.read())
    tr = trace_io.load(dst)
    ctx = replay.make_context(tr)
    t1 = replay.make_t1(tr)
    events = list(tr.events)

    def apply(ev, to_ctx=True):
        if ev["type"] == trace_io.EV_MAP:
            args = (int(ev["block"]), int(ev["model"]), int(ev["layer"]), float(ev["zmin"]), float(ev["zmax"]), tr.polygon(ev))
            if to_ctx:
                ctx.block_map(*args)
            t1.block_map(*args)
        elif ev["type"] == trace_io.EV_UNMAP:
            if to_ctx:
                ctx.block_unmap(int(ev["block"]), int(ev["layer"]))
            t1.block_unmap(int(ev["block"]), int(ev["layer"]))

    start = 0
    present = {}  # block -> (zmin, zmax, polygon) of its latest Map on either layer: the outline of the pose it has now
    n_moves = n_stalled = n_ticks = 0
    t_move = 0.0
    i = start
    while i < len(events):
        # one tick: up to the next TICK event
        j = i
        while j < len(events) and events[j]["type"] != trace_io.EV_TICK:
            j += 1
        tick = events[i:j]
        # a Move shows up as [UNMAP x n] [MAP x n] COLLISION and, if it hit, [UNMAP x n] [MAP x n] again (the undo)
        movers, want, claimed = [], [], set()
        layer = None
        for k, ev in enumerate(tick):
            if ev["type"] != trace_io.EV_COLLISION:
                continue
            _, layer, blocks, hit = tr.collision(ev)
            mine = set(int(b) for b in blocks)
            new, m = {}, k - 1
            while m >= 0 and len(new) < len(mine) and tick[m]["type"] == trace_io.EV_MAP and int(tick[m]["block"]) in mine:
                new[int(tick[m]["block"])] = (float(tick[m]["zmin"]), float(tick[m]["zmax"]), np.array(tr.polygon(tick[m])))
                claimed.add(m)
                m -= 1
            assert len(new) == len(mine)
            while m >= 0 and tick[m]["type"] == trace_io.EV_UNMAP and int(tick[m]["block"]) in mine and m not in claimed:
                claimed.add(m)
                m -= 1
            movers.append([(int(b), new[int(b)]) for b in blocks])
            want.append(1 if hit >= 0 else 0)
            if hit >= 0:  # the undo that follows is the expected RESULT, not an input
                m = k + 1
                while m < len(tick) and tick[m]["type"] in (trace_io.EV_UNMAP, trace_io.EV_MAP) and int(tick[m]["block"]) in mine and m - k <= 2 * len(mine):
                    claimed.add(m)
                    m += 1
        for k, ev in enumerate(tick):  # whatever no Move accounts for (the world load, in this trace) goes in as it came
            if k not in claimed and ev["type"] in (trace_io.EV_MAP, trace_io.EV_UNMAP):
                assert not movers or k < min(claimed), "a foreign Map / UnMap in the middle of the movers"
                if ev["type"] == trace_io.EV_MAP:
                    ctx.block_map(int(ev["block"]), int(ev["model"]), int(ev["layer"]), float(ev["zmin"]), float(ev["zmax"]), tr.polygon(ev))
                    present[int(ev["block"])] = (float(ev["zmin"]), float(ev["zmax"]), np.array(tr.polygon(ev)))
                else:
                    ctx.block_unmap(int(ev["block"]), int(ev["layer"]))
        movers = [[(b, new, present[b]) for b, new in m] for m in movers]
        for ev in tick:   # the oracle replays the reference's own sequence, call by call
            apply(ev, to_ctx=False)
            if ev["type"] == trace_io.EV_MAP:
                present[int(ev["block"])] = (float(ev["zmin"]), float(ev["zmax"]), np.array(tr.polygon(ev)))
        if movers:
            t0 = time.perf_counter()
            got = ctx.models_move(layer, movers)
            t_move += time.perf_counter() - t0
            assert list(got) == want, "tick %d: stalled %s, reference %s" % (n_ticks, list(got), want)
            n_moves += len(movers)
            n_stalled += int(sum(want))
        n_ticks += 1
        if n_ticks % 25 == 0 or j >= len(events) - 1:
            rec_g, cnt_g = ctx.grid_dump()
            rec_o, cnt_o = t1.dump_grid()
            assert np.array_equal(oracle_lib.sort_grid(rec_g), rec_o), "grid differs after tick %d" % n_ticks
            assert np.array_equal(oracle_lib.sort_counts(cnt_g), cnt_o)
        i = j + 1
    print("models_move: %d calls, %.2f ms each (14 movers)" % (n_ticks, 1e3 * t_move / max(n_ticks, 1)))
    assert n_moves == 2100 and n_stalled == 1728
    ctx.close()
