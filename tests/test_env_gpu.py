"""GPU parity: rrl_env_step / rrl_env_reset (through the C ABI) against the golden vectors of the
reference's Navigation classes (bit-exact fp64) and against the oracle's restated Maze."""
import os

import numpy as np
import pytest
import torch

from oracle import envs

pytestmark = pytest.mark.gpu


def _soa(x, dev):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, np.float64).T)).to(dev)


def _step(native, dev, kind, state, action, noise=None, ep_steps=None, horizon=100, reset_draws=None, rings=None,
          recovery=None, action_task=None, penalty=0.0):
    n = len(state)
    cfg = native.env_config(kind, n, horizon=horizon, reward_penalty=penalty, seed=5, stream_id=0)
    st = _soa(state, dev)
    a = torch.from_numpy(np.ascontiguousarray(action, np.float32)).to(dev)
    at = a if action_task is None else torch.from_numpy(np.ascontiguousarray(action_task, np.float32)).to(dev)
    es = torch.zeros(n, dtype=torch.int32, device=dev) if ep_steps is None else torch.from_numpy(
        np.asarray(ep_steps, np.int32)).to(dev)
    er = torch.zeros(n, dtype=torch.float64, device=dev)
    cnt = torch.zeros(native.NUM_COUNTERS, dtype=torch.int64, device=dev)
    o_ns = torch.empty(2, n, dtype=torch.float64, device=dev)
    o_r = torch.empty(n, dtype=torch.float64, device=dev)
    o_d = torch.empty(n, dtype=torch.uint8, device=dev)
    o_c = torch.empty(n, dtype=torch.uint8, device=dev)
    o_s = torch.empty(n, dtype=torch.uint8, device=dev)
    kw = {}
    if rings is not None:
        kw = dict(task_ring=rings[0], task_capacity=rings[0].shape[0], cons_ring=rings[1], cons_flags=rings[2],
                  cons_capacity=rings[1].shape[0])
    native.env_step(cfg, at, a, st, es, er, cnt,
                    recovery=None if recovery is None else torch.from_numpy(np.asarray(recovery, np.uint8)).to(dev),
                    noise=None if noise is None else _soa(noise, dev),
                    reset_draws=None if reset_draws is None else _soa(reset_draws, dev),
                    out_next_state=o_ns, out_reward=o_r, out_done=o_d, out_constraint=o_c, out_success=o_s, **kw)
    torch.cuda.synchronize()
    return dict(next_state=o_ns.cpu().numpy().T, reward=o_r.cpu().numpy(), done=o_d.cpu().numpy().astype(bool),
                constraint=o_c.cpu().numpy().astype(bool), success=o_s.cpu().numpy().astype(bool),
                state=st.cpu().numpy().T, ep_steps=es.cpu().numpy(), ep_return=er.cpu().numpy(),
                counters=cnt.cpu().numpy())


@pytest.mark.parametrize("name,kind", [("nav1", 0), ("nav2", 1)])
def test_nav_step_bit_exact_vs_reference(native, cuda, golden_dir, name, kind):
    z = np.load(os.path.join(golden_dir, "nav_step_%s.npz" % name))
    o = _step(native, cuda, kind, z["state"], z["action"], z["noise"])
    assert np.array_equal(o["next_state"], z["next_state"])
    assert np.array_equal(o["reward"], z["reward"])
    assert np.array_equal(o["done"], z["done"].astype(bool))
    assert np.array_equal(o["constraint"], z["constraint"].astype(bool))
    assert np.array_equal(o["success"], z["success"].astype(bool))


def test_nav_horizon_reset_and_counters(native, cuda):
    rs = np.random.RandomState(3)
    n = 5000
    s = np.stack([rs.uniform(-70, 10, n), rs.uniform(-4, 4, n)], 1)
    a = rs.uniform(-1, 1, (n, 2)).astype(np.float32)
    noise = rs.randn(n, 2)
    draws = rs.randn(n, 2)
    steps = rs.randint(90, 100, n)            # some hit steps == horizon
    rec = rs.rand(n) < 0.5
    o = _step(native, cuda, 0, s, a, noise, ep_steps=steps, reset_draws=draws, recovery=rec)
    ns, r, d, c, su = envs.nav_step(envs.NAV1, s, a, noise)
    done_h = d | (steps + 1 == 100)
    assert np.array_equal(o["done"], done_h)
    expect_state = np.where(done_h[:, None], envs.nav_reset(draws), ns)
    assert np.array_equal(o["state"], expect_state)
    assert np.array_equal(o["ep_steps"], np.where(done_h, 0, steps + 1))
    cn = o["counters"]
    assert cn[native.C_EPISODES] == done_h.sum()
    assert cn[native.C_NUM_VIOLS] == (done_h & c).sum()
    assert cn[native.C_NUM_SUCCESSES] == (done_h & su).sum()
    assert cn[native.C_VIOL_RECOVERY] == (done_h & c & rec).sum()
    assert cn[native.C_VIOL_NO_RECOV] == (done_h & c & ~rec).sum()


def test_nav_step_pushes_both_rings(native, cuda):
    rs = np.random.RandomState(4)
    n, cap = 3000, 4096
    s = np.stack([rs.uniform(-70, 10, n), rs.uniform(-8, 8, n)], 1)
    a_real = rs.uniform(-1.3, 1.3, (n, 2)).astype(np.float32)
    a_task = rs.uniform(-1, 1, (n, 2)).astype(np.float32)
    noise = rs.randn(n, 2)
    task = torch.zeros(cap, 8, device=cuda)
    cons = torch.zeros(cap, 8, device=cuda)
    flags = torch.zeros(cap, dtype=torch.uint8, device=cuda)
    o = _step(native, cuda, 0, s, a_real, noise, rings=(task, cons, flags), action_task=a_task, penalty=3.0,
              reset_draws=rs.randn(n, 2))
    ns, r, d, c, su = envs.nav_step(envs.NAV1, s, a_real, noise)
    mask = (~d).astype(np.float32)                              # experiment.py:434 (before horizon)
    r_pen = np.where(c, r - 3.0, r)                             # experiment.py:431-432
    t = task.cpu().numpy()[:n]
    q = cons.cpu().numpy()[:n]
    assert np.array_equal(t[:, 0:2], s.astype(np.float32))
    assert np.array_equal(t[:, 2:4], a_task)                    # relabelled: proposed action (:438-441)
    assert np.array_equal(t[:, 4], r_pen.astype(np.float32))
    assert np.array_equal(t[:, 5:7], ns.astype(np.float32))
    assert np.array_equal(t[:, 7], mask)
    assert np.array_equal(q[:, 2:4], a_real)                    # executed action (:443-445)
    assert np.array_equal(q[:, 4], c.astype(np.float32))
    assert np.array_equal(flags.cpu().numpy()[:n], np.where(c, 1, 2))
    assert (flags.cpu().numpy()[n:] == 0).all()


def test_maze_step_bit_exact_vs_restatement(native, cuda):
    rs = np.random.RandomState(11)
    n = 20000
    s = rs.uniform(-0.28, 0.28, (n, 2))
    s[:4000] = envs.maze_reset_from_uniform(rs.rand(4000, 2))
    # clusters hugging the walls so that contact happens mid-flight
    s[4000:6000, 0] = -0.1 + rs.uniform(-0.06, 0.06, 2000)
    s[6000:8000, 0] = 0.1 + rs.uniform(-0.06, 0.06, 2000)
    s[8000:9000] = (0.25, 0.0) + rs.uniform(-0.04, 0.04, (1000, 2))          # goal region
    a = rs.uniform(-0.13, 0.13, (n, 2)).astype(np.float32)
    steps = rs.randint(0, 100, n)
    o = _step(native, cuda, 2, s, a, ep_steps=steps, reset_draws=rs.rand(n, 2))
    ns, r, d, c, su = envs.maze_step(s, a, steps)
    assert np.array_equal(o["constraint"], c)
    assert np.array_equal(o["next_state"], ns)
    assert np.array_equal(o["reward"], r)
    assert np.array_equal(o["done"], d)
    assert np.array_equal(o["success"], su)
    assert 500 < c.sum() < n - 500


def test_maze_reset_matches_restatement(native, cuda):
    n = 4096
    u = np.random.RandomState(5).rand(n, 2)
    cfg = native.env_config(2, n)
    st = torch.zeros(2, n, dtype=torch.float64, device=cuda)
    native.env_reset(cfg, st, draws=_soa(u, cuda))
    assert np.array_equal(st.cpu().numpy().T, envs.maze_reset_from_uniform(u))


def test_philox_noise_statistics(native, cuda):
    n = 1 << 20
    s = np.zeros((n, 2)); s[:, 0] = -50.0
    o = _step(native, cuda, 0, s, np.zeros((n, 2), np.float32))
    e = (o["next_state"] - s) / 0.05
    assert abs(e.mean()) < 5e-3 and abs(e.std() - 1.0) < 5e-3
    assert abs(np.corrcoef(e[:, 0], e[:, 1])[0, 1]) < 5e-3
    assert abs((e ** 4).mean() - 3.0) < 0.05


def test_maze_contact_edges_and_corners_bit_exact(native, cuda):
    """adversarial placements for the speculative-chunk substep loop: discs within a few ulps .. 1e-7 of touching a
    plane, a wall face or a rounded wall corner, moving towards / along / away from it, and tiny actions."""
    rs = np.random.RandomState(3)
    R = envs.MAZE_R
    pts = []
    for x0, x1, y0, y1 in envs.maze_walls():
        for cx in (x0, x1):
            for cy in (y0, y1):
                if abs(cy) > 0.29:
                    continue
                ang = rs.uniform(0, 2 * np.pi, 300)
                rad = R + rs.choice([0.0, 1e-15, -1e-15, 1e-12, 1e-9, 3e-9, 1e-7, 1e-4, 2e-3], 300)
                pts.append(np.stack([cx + rad * np.cos(ang), cy + rad * np.sin(ang)], 1))      # rounded corners
        for xf, sgn in ((x0, -1.0), (x1, 1.0)):                                                # wall faces
            yy = rs.uniform(max(y0, -0.27), min(y1, 0.27), 300)
            off = R + rs.choice([0.0, 1e-15, 1e-12, 1e-9, 2e-9, 1e-7, 1e-5, 1e-3, 2e-2], 300)
            pts.append(np.stack([xf + sgn * off, yy], 1))
    for sgn in (-1.0, 1.0):                                                                    # outer planes
        off = 0.3 - R - rs.choice([0.0, 1e-16, 1e-13, 1e-9, 1e-6, 1e-3, 2e-2], 300)
        pts.append(np.stack([sgn * off, rs.uniform(-0.27, 0.27, 300)], 1))
        pts.append(np.stack([rs.uniform(-0.27, 0.27, 300), sgn * off], 1))
    s = np.clip(np.concatenate(pts), -0.2999, 0.2999)
    n = len(s)
    a = rs.uniform(-0.12, 0.12, (n, 2)).astype(np.float32)
    a[::7] *= np.float32(1e-3)                  # crawling
    a[::11, 0] = 0.0                            # axis-aligned motion
    a[::13, 1] = 0.0
    steps = rs.randint(0, 100, n)
    o = _step(native, cuda, 2, s, a, ep_steps=steps, reset_draws=rs.rand(n, 2))
    ns, r, d, c, su = envs.maze_step(s, a, steps)
    assert np.array_equal(o["constraint"], c)
    assert np.array_equal(o["next_state"], ns)
    assert np.array_equal(o["reward"], r) and np.array_equal(o["done"], d) and np.array_equal(o["success"], su)
    assert 0.2 * n < c.sum() < 0.9 * n
