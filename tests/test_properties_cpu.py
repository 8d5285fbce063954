"""CPU property tests (hypothesis) of the restated algorithm the CUDA path is checked against -- the cases SURVEY.md
section 4 lists: the random.sample path switch at n = setsize, ring-buffer wrap, obstacle closed-interval edges,
mask-before-horizon ordering, stratified batch layout."""
import random

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import envs as oenvs
from oracle import replay as oreplay

FAST = settings(max_examples=60, deadline=None, derandomize=True, database=None)


def _setsize(k):
    s = 21
    if k > 5:
        p = 1
        while p < 3 * k:
            p *= 4
        s += p
    return s


@FAST
@given(seed=st.integers(0, 2 ** 63), k=st.integers(1, 300), off=st.integers(-3, 3), extra=st.integers(0, 5000))
def test_sample_path_switch_matches_cpython(seed, k, off, extra):
    """Lib/random.py sample(): pool path for n <= setsize(k), set path above; both sides of the switch and far from it"""
    for n in {max(k, _setsize(k) + off), k + extra}:
        random.seed(seed)
        want = random.sample(range(n), k)
        assert oreplay.MT19937(seed).sample_indices(n, k) == want


@FAST
@given(cap=st.integers(1, 40), pushes=st.lists(st.floats(-10, 10, width=32), min_size=0, max_size=120))
def test_ring_buffer_wrap(cap, pushes):
    """replay_memory.py:21-25: append until full, then overwrite at position, position = (position + 1) % capacity"""
    stream = oreplay.SharedStream()
    mem = oreplay.ReplayMemory(cap, 1, stream)
    model, pos = [], 0
    for v in pushes:
        mem.push((v, 0.0), (0.0, 0.0), v, (v, 1.0), 1.0)
        if len(model) < cap:
            model.append(None)
        model[pos] = np.float32(v)
        pos = (pos + 1) % cap
    assert len(mem) == len(model) == min(len(pushes), cap) and mem.position == pos
    assert [mem.buf[i, 0] for i in range(len(model))] == model


@FAST
@given(kind=st.sampled_from([oenvs.NAV1, oenvs.NAV2]), which=st.integers(0, 2), ex=st.integers(0, 1), ey=st.integers(0, 1),
       t=st.floats(0.0, 1.0))
def test_obstacle_edges_are_closed(kind, which, ex, ey, t):
    """obstacle.py:13-15: `lo <= x <= hi` on both axes -- points ON an edge are inside, the next double outside is not
    (unless another rectangle covers it)"""
    rects = oenvs.NAV_RECTS[kind]
    (x0, x1), (y0, y1) = rects[which % len(rects)]
    # a point on the x-edge `ex`, anywhere along y; and one on the y-edge `ey`
    xe, ye = (x0, x1)[ex], (y0, y1)[ey]
    y = y0 + t * (y1 - y0)
    x = x0 + t * (x1 - x0)
    assert oenvs.nav_obstacle(kind, xe, y) and oenvs.nav_obstacle(kind, x, ye)
    out_x = np.nextafter(xe, -np.inf if ex == 0 else np.inf)
    out_y = np.nextafter(ye, -np.inf if ey == 0 else np.inf)
    others = [r for i, r in enumerate(rects) if i != which % len(rects)]

    def covered(px, py):
        return any(a0 <= px <= a1 and b0 <= py <= b1 for (a0, a1), (b0, b1) in others)
    assert bool(oenvs.nav_obstacle(kind, out_x, y)) == covered(out_x, y)
    assert bool(oenvs.nav_obstacle(kind, x, out_y)) == covered(x, out_y)


def test_mask_is_computed_before_the_horizon_truncation():
    """experiment.py:434-435: an episode cut by the horizon is stored with mask = 1 (it still bootstraps), one ended by
    the environment with mask = 0"""
    from oracle.loop import OracleExperiment
    exp = OracleExperiment("navigation1", seed=11, batch_size=16, use_recovery=False, start_steps=0)
    exp.agent.act = lambda *a, **k: [np.zeros((1, 2), np.float32)]             # stay near the start: only the horizon ends it
    infos = [exp.step() for _ in range(100)]
    assert [i["episode_end"] for i in infos] == [False] * 99 + [True]
    assert not infos[-1]["constraint"] and not infos[-1]["success"]
    assert exp.memory.buf[99, 7] == 1.0 and len(exp.memory) == 100
    exp2 = OracleExperiment("navigation1", seed=11, batch_size=16, use_recovery=False, start_steps=0)
    exp2.state = np.array([-75.0, 9.5])          # next to the upper wall: drive into it
    exp2.ep_steps = 0
    exp2.agent.act = lambda *a, **k: [np.array([[0.0, 1.0]], np.float32)]
    info = exp2.step()
    while not info["episode_end"]:
        info = exp2.step()
    assert info["constraint"] and exp2.memory.buf[len(exp2.memory) - 1, 7] == 0.0


@FAST
@given(seed=st.integers(0, 2 ** 32), n=st.integers(40, 400), p=st.floats(0.2, 0.8), B=st.integers(4, 32),
       pf=st.sampled_from([0.25, 0.3, 0.5]))
def test_stratified_batch_layout(seed, n, p, B, pf):
    """replay_memory.py:54-72: int(B * pos_fraction) positives first, then the negatives, all distinct"""
    rs = np.random.RandomState(seed % (2 ** 31))
    flags = (rs.rand(n) < p).astype(np.float64)
    pos, neg = int(B * pf), B - int(B * pf)
    if flags.sum() < pos or (1 - flags).sum() < neg:
        return
    stream = oreplay.SharedStream()
    mem = oreplay.ConstraintReplayMemory(n + 7, seed, stream)
    for i in range(n):
        mem.push((float(i), 0.0), (0.0, 0.0), flags[i], (0.0, 0.0), 1.0)
    idx = mem.sample_slots(B, pf)
    assert len(idx) == B and len(set(idx.tolist())) == B
    assert flags[idx[:pos]].all() and not flags[idx[pos:]].any()
    random.seed(seed)
    pi = random.sample(range(int(flags.sum())), pos)
    ni = random.sample(range(n - int(flags.sum())), neg)
    assert np.array_equal(idx, np.concatenate([np.flatnonzero(flags)[pi], np.flatnonzero(1 - flags)[ni]]))
