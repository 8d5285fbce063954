"""GPU property tests (hypothesis) through the C ABI, mirroring tests/test_properties_cpu.py: the device sampler on
both sides of the random.sample path switch, ring wrap of rrl_replay_push, closed obstacle edges in rrl_env_step."""
import random

import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import envs as oenvs
from test_env_gpu import _step
from test_properties_cpu import _setsize

pytestmark = pytest.mark.gpu
GPU = settings(max_examples=25, deadline=None, derandomize=True, database=None, suppress_health_check=[HealthCheck.function_scoped_fixture])


def _device_sample(native, dev, seed, n, k):
    cap = max(n, k) + 3
    ring = torch.zeros(cap, 8, device=dev)
    ring[:, 0] = torch.arange(cap, device=dev) % 4096
    ring[:, 1] = torch.arange(cap, device=dev) // 4096
    cnt = torch.zeros(native.NUM_COUNTERS, dtype=torch.int64, device=dev)
    cnt[native.C_TASK_LEN] = n
    mt = native.mt19937_seed(seed).to(dev)
    out = [torch.zeros(k, 2, device=dev), torch.zeros(k, 2, device=dev), torch.zeros(k, device=dev),
           torch.zeros(k, 2, device=dev), torch.zeros(k, device=dev)]
    idx = torch.full((k,), -1, dtype=torch.int64, device=dev)
    cfg = native.sample_config(cap, k, False, None, 0, 512, -1.0)
    native.replay_sample(cfg, ring, mt, cnt, native.C_SAC_ROWS, *out, out_idx=idx)
    torch.cuda.synchronize()
    assert int(cnt[native.C_ERROR].item()) == 0 and int(cnt[native.C_SAC_ROWS].item()) == k
    return idx.cpu().numpy().tolist()


@GPU
@given(seed=st.integers(0, 2 ** 63), k=st.integers(1, 256), off=st.integers(-3, 3), extra=st.integers(0, 100000))
def test_device_sampler_path_switch_matches_cpython(native, cuda, seed, k, off, extra):
    for n in {max(k, _setsize(k) + off), k + extra}:
        random.seed(seed)
        assert _device_sample(native, cuda, seed, n, k) == random.sample(range(n), k)


@GPU
@given(cap=st.integers(16, 200), bursts=st.lists(st.integers(1, 16), min_size=1, max_size=40))
def test_device_ring_wrap(native, cuda, cap, bursts):
    ring = torch.zeros(cap, 8, device=cuda)
    cnt = torch.zeros(native.NUM_COUNTERS, dtype=torch.int64, device=cuda)
    model, pos, c = [None] * cap, 0, 0
    length = 0
    for b in bursts:
        rec = np.zeros((b, 8), np.float32)
        rec[:, 0] = np.arange(c, c + b)
        native.replay_push(ring, cap, torch.from_numpy(rec).to(cuda), b, cnt)
        for i in range(b):
            model[pos] = float(c + i)
            pos = (pos + 1) % cap
        c += b
        length = min(length + b, cap)
    torch.cuda.synchronize()
    assert int(cnt[native.C_TASK_POS].item()) == pos and int(cnt[native.C_TASK_LEN].item()) == length
    got = ring[:, 0].cpu().numpy()
    for i, v in enumerate(model):
        if v is not None:
            assert got[i] == v


@GPU
@given(kind=st.sampled_from([oenvs.NAV1, oenvs.NAV2]), which=st.integers(0, 2), ex=st.integers(0, 1), t=st.floats(0.0, 1.0))
def test_device_obstacle_edges_are_closed(native, cuda, kind, which, ex, t):
    """a state exactly ON an obstacle edge is stuck (navigation1.py:99-101) and flagged; the next double outside moves"""
    rects = oenvs.NAV_RECTS[kind]
    (x0, x1), (y0, y1) = rects[which % len(rects)]
    xe = (x0, x1)[ex]
    y = y0 + t * (y1 - y0)
    out_x = np.nextafter(xe, -np.inf if ex == 0 else np.inf)
    state = np.array([[xe, y], [out_x, y]])
    action = np.zeros((2, 2), np.float32)
    noise = np.zeros((2, 2))
    o = _step(native, cuda, kind, state, action, noise)
    ns, r, d, c, su = oenvs.nav_step(kind, state, action, noise)
    assert np.array_equal(o["next_state"], ns) and np.array_equal(o["constraint"], c) and np.array_equal(o["done"], d)
    assert o["constraint"][0]
