"""GPU: the comparison branches of the vectorised acting path (csrc/select.cu, VecEngine) against the oracle:
SQRL's action filter (sac.py:139-161), Q-sampling recovery (qrisk.py:214-225), the second task-buffer push of
--add_both_transitions (experiment.py:446-448) and --policy Deterministic (model.py:447-485)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _smoke():
    spec = importlib.util.spec_from_file_location("rrl_smoke_impl", os.path.join(HERE, "smoke_impl.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _setup(env_name, n, B, seed, tensor_cores, eps_safe=0.3, gamma_safe=0.8, demos_n=400, **engine_kw):
    """engine + oracle agent with identical (pre-trained) weights, env copies reset to the same states"""
    from oracle import envs as oenvs
    from oracle.agent import Agent
    from recovery_rl import native
    from recovery_rl.engine import VecEngine, ACTION_SCALE
    kind = oenvs.KIND_BY_NAME[env_name]
    sc = ACTION_SCALE[env_name]
    torch.manual_seed(seed)
    np.random.seed(seed)
    ora = Agent(action_scale=(np.float32(sc),) * 2, gamma_safe=gamma_safe, eps_safe=eps_safe,
                mf_recovery=engine_kw.get("mf_recovery", True), dgd=engine_kw.get("dgd", False))
    eng = VecEngine(env_name, n, batch_size=B, replay_size=8 * n, safe_replay_size=8 * n, gamma_safe=gamma_safe,
                    eps_safe=eps_safe, seed=seed, host_inputs=True, start_steps=0, use_tensor_cores=tensor_cores, **engine_kw)
    eng.init_agent(ora.nets())
    rs = np.random.RandomState(seed)
    demos = oenvs.maze_offline_data(demos_n, rs) if kind == oenvs.MAZE else oenvs.nav_offline_data(kind, demos_n)
    eng.push_offline(demos)
    eng.pretrain_qrisk(3, n_demos=len(demos))

    def mirror():
        P = {net: eng.agent.params(net) for net in native.NET_NAMES}
        ora.load(lambda net, i: P[net][i])
    mirror()
    draws = rs.rand(n, 2) if kind == oenvs.MAZE else rs.randn(n, 2)
    eng.reset(torch.from_numpy(np.ascontiguousarray(draws.T)).to(eng.device))
    return eng, ora, rs, kind, mirror


def _inputs(eng, rs, kind, n, B):
    from oracle import envs as oenvs
    inp = dict(reset_draws=(rs.rand(n, 2) if kind == oenvs.MAZE else rs.randn(n, 2)).T,
               eps_task=rs.randn(n, 2).astype(np.float32), eps_rec=rs.randn(n, 2).astype(np.float32),
               rand_u=rs.rand(n, 2).astype(np.float32),
               sac_eps_next=rs.randn(B, 2).astype(np.float32), sac_eps_cur=rs.randn(B, 2).astype(np.float32),
               qr_eps_next=rs.randn(B, 2).astype(np.float32), qr_eps_rec=rs.randn(B, 2).astype(np.float32))
    if kind != oenvs.MAZE:
        inp["env_noise"] = rs.randn(n, 2).T
    return inp


def _set_eps_safe(eng, ora, value):
    eng.cfg.eps_safe = float(value)      # the config struct is passed by reference to every call
    ora.eps_safe = float(value)


@pytest.mark.parametrize("tensor_cores", [0, 2])
def test_sqrl_action_filter_matches_oracle(native, cuda, tensor_cores):
    """every env copy's filtered draw == oracle.Agent.select_action_sqrl fed the same candidate noise, with the Categorical
    draw restated as the inverse CDF of the same uniform; both branches (someone passes the filter / nobody does)."""
    n, B, K = 192, 64, 100
    eng, ora, rs, kind, mirror = _setup("navigation1", n, B, 21, tensor_cores, use_recovery=False, dgd=True, update_nu=True,
                                        constraint_sampling=True)
    assert eng._sel_ws is not None and not eng.staged_act
    seen = dict(filtered=0, empty=0, skipped=0)
    for t in range(3):
        inp = _inputs(eng, rs, kind, n, B)
        inp["sqrl_eps"] = rs.randn(n, K, 2).astype(np.float32)
        inp["sqrl_u"] = rs.rand(n).astype(np.float32)
        state = eng.state.cpu().numpy().T.copy()
        if t == 0:
            # put the threshold in the middle of the candidates' Q_risk values so that both branches are exercised
            with torch.no_grad():
                sb = torch.as_tensor(np.repeat(state, K, axis=0), dtype=torch.float32)
                pi, _, _ = ora.policy.sample(sb, torch.as_tensor(inp["sqrl_eps"].reshape(-1, 2)))
                q = torch.max(*ora.qrisk(sb, pi)).numpy().reshape(n, K)
            _set_eps_safe(eng, ora, float(np.median(q.min(1)) + 0.25 * (np.median(q) - np.median(q.min(1)))))
        out = eng.step_host(inp)
        mirror()                                   # the updates precede the action (experiment.py:397-419)
        got = eng.action_task.cpu().numpy()
        assert np.array_equal(out["action"].numpy(), got) and not out["recovery"].numpy().any()
        for i in range(n):
            eps_i = inp["sqrl_eps"][i]
            with torch.no_grad():
                sb = torch.as_tensor(state[i], dtype=torch.float32).unsqueeze(0).repeat(K, 1)
                pi, lp, _ = ora.policy.sample(sb, torch.as_tensor(eps_i))
                q = torch.max(*ora.qrisk(sb, pi)).numpy().ravel().astype(np.float64)
            ok = q <= ora.eps_safe
            if np.abs(q - ora.eps_safe).min() < 2e-5:           # a candidate sits on the threshold: fp32-order dependent
                seen["skipped"] += 1
                continue
            u = float(inp["sqrl_u"][i])

            def categorical(probs, u=u):
                p64 = probs.numpy().astype(np.float64)
                cdf = np.cumsum(p64)
                tgt = u * cdf[-1]
                if np.abs(cdf - tgt).min() < 1e-5 * cdf[-1]:
                    raise FloatingPointError
                return int(np.searchsorted(cdf, tgt, side="right"))
            try:
                want = ora.select_action_sqrl(state[i].astype(np.float32), eps_i, categorical=categorical)
            except FloatingPointError:
                seen["skipped"] += 1
                continue
            if not ok.any():
                srt = np.sort(q)
                if srt[1] - srt[0] < 5e-6:
                    seen["skipped"] += 1
                    continue
            seen["filtered" if ok.any() else "empty"] += 1
            assert np.allclose(got[i], want, rtol=1e-4, atol=1e-5), (t, i, got[i], want, int(ok.sum()))
        assert eng.read_counters()["error"] == 0
    assert seen["filtered"] > 30 and seen["empty"] > 30 and seen["skipped"] < n, seen


@pytest.mark.parametrize("tensor_cores", [0, 2])
def test_q_sampling_recovery_matches_oracle(native, cuda, tensor_cores):
    """env copies whose task action is too risky execute the least risky of 1000 uniform candidate actions (qrisk.py:214-225)."""
    from recovery_rl.engine import ACTION_SCALE
    n, B, K = 96, 64, 1000
    eng, ora, rs, kind, mirror = _setup("maze", n, B, 22, tensor_cores, eps_safe=0.15, gamma_safe=0.5, use_recovery=True,
                                        mf_recovery=False, q_sampling_recovery=True)
    sc = np.float32(ACTION_SCALE["maze"])
    total = 0
    for t in range(3):
        inp = _inputs(eng, rs, kind, n, B)
        inp["qs_u"] = rs.rand(n, K, 2).astype(np.float32)
        state = eng.state.cpu().numpy().T.copy()
        if t == 0:
            with torch.no_grad():
                st = torch.as_tensor(state, dtype=torch.float32)
                a, _, _ = ora.policy.sample(st, torch.as_tensor(inp["eps_task"]))
                q = torch.max(*ora.qrisk(st, a)).numpy().ravel()
            _set_eps_safe(eng, ora, float(np.median(q)))          # about half of the env copies recover
        out = eng.step_host(inp)
        mirror()
        rec = out["recovery"].numpy().astype(bool)
        a_task, a_real = eng.action_task.cpu().numpy(), out["action"].numpy()
        assert np.array_equal(a_real[~rec], a_task[~rec])
        for i in np.flatnonzero(rec):
            cand = ((np.float32(2.0) * inp["qs_u"][i] - np.float32(1.0)) * sc).astype(np.float32)     # Box.sample of the action space
            with torch.no_grad():
                sb = torch.as_tensor(state[i], dtype=torch.float32).unsqueeze(0).repeat(K, 1)
                q = torch.max(*ora.qrisk(sb, torch.as_tensor(cand))).numpy().ravel()
            j = np.flatnonzero(np.abs(cand - a_real[i]).max(1) < 1e-7)
            assert len(j) >= 1, (t, i, a_real[i])                       # the executed action IS one of the candidates ...
            assert q[j[0]] <= q.min() + 5e-6, (t, i, q[j[0]], q.min())   # ... and (up to fp32 ordering of near ties) the least risky
            total += 1
        assert eng.read_counters()["error"] == 0
    assert total > 40


def test_add_both_transitions_appends_the_executed_action_rows(native, cuda):
    """--add_both_transitions: the env copies whose recovery policy acted push (state, real_action, reward, next_state, mask)
    into the task ring as well, in env order after the step's n relabelled rows; positions wrap with the ring."""
    n, B = 256, 64
    eng, ora, rs, kind, mirror = _setup("navigation2", n, B, 23, 0, eps_safe=0.2, gamma_safe=0.65, use_recovery=True,
                                        add_both_transitions=True)
    cap = eng.task_cap
    exp_pos = exp_len = 0
    pushed = 0
    cons0 = int(eng.counters[native.C_CONS_POS])            # the offline demos
    for t in range(14):                      # 8 n slots: the ring wraps
        inp = _inputs(eng, rs, kind, n, B)
        if t == 0:
            state = eng.state.cpu().numpy().T
            with torch.no_grad():
                st = torch.as_tensor(state, dtype=torch.float32)
                a, _, _ = ora.policy.sample(st, torch.as_tensor(inp["eps_task"]))
                q = torch.max(*ora.qrisk(st, a)).numpy().ravel()
            _set_eps_safe(eng, ora, float(np.quantile(q, 0.6)))
        out = eng.step_host(inp)
        rec = out["recovery"].numpy().astype(bool)
        ring = eng.task_ring.cpu().numpy()
        reg = ring[(exp_pos + np.arange(n)) % cap]
        assert np.allclose(reg[:, 2:4], eng.action_task.cpu().numpy())          # the relabelled rows (experiment.py:438-439)
        m = int(rec.sum())
        extra = ring[(exp_pos + n + np.arange(m)) % cap]
        want = reg[rec].copy()
        want[:, 2:4] = out["action"].numpy()[rec]
        assert np.array_equal(extra, want), t
        exp_pos = (exp_pos + n + m) % cap
        exp_len = min(exp_len + n + m, cap)
        pushed += m
        c = eng.counters.cpu().numpy()
        assert c[native.C_TASK_POS] == exp_pos and c[native.C_TASK_LEN] == exp_len, (t, c[native.C_TASK_POS], exp_pos)
        assert c[native.C_CONS_POS] == ((t + 1) * n + cons0) % eng.cons_cap     # the constraint ring gets one row per env copy
        assert c[native.C_ERROR] == 0
    assert pushed > 50 and exp_len == cap


@pytest.mark.parametrize("tensor_cores", [0, 2])
@pytest.mark.parametrize("env_name", ["navigation1", "maze"])
def test_deterministic_policy_vector_step_matches_oracle(native, cuda, env_name, tensor_cores):
    """--policy Deterministic in the vector engine: acting (mean + clamped noise), the SAC update with alpha = 0 and the
    safety-critic update, stage by stage against the oracle."""
    kw = dict(gamma_safe=0.5, eps_safe=0.15, pos_fraction=0.3, demos_n=600) if env_name == "maze" else {}
    assert _smoke().run(env_name=env_name, n=384, B=64, steps=4, seed=6, verbose=False, tensor_cores=tensor_cores,
                        deterministic=True, **kw)


def test_philox_mode_runs_the_comparison_branches_in_a_graph(native, cuda):
    """device RNG + CUDA graph: the SQRL filter, Q-sampling recovery (chunked workspace) and add_both_transitions capture and
    replay; the chunked candidate evaluation draws the same candidates as the unchunked one."""
    from env.maze import get_offline_data
    from recovery_rl.engine import VecEngine

    def make(**kw):
        torch.manual_seed(2)
        e = VecEngine("maze", 1024, batch_size=64, replay_size=16384, safe_replay_size=16384, gamma_safe=0.5, eps_safe=0.15,
                      pos_fraction=0.3, seed=4, start_steps=1024, use_tensor_cores=2, **kw)
        e.init_agent()
        e.push_offline(get_offline_data(2000, rng=np.random.RandomState(4)))
        e.pretrain_qrisk(5)
        e.reset()
        return e

    a = make(use_recovery=False, dgd=True, update_nu=True, constraint_sampling=True)
    b = make(use_recovery=False, dgd=True, update_nu=True, constraint_sampling=True)
    b._sel_ws = torch.empty(native.select_workspace_floats(100, 100), device=b.device)      # 11 chunks of <= 100 env copies
    for e in (a, b):
        e.capture()
        assert e.graph is not None
        for _ in range(5):
            e.replay()
    torch.cuda.synchronize()
    assert torch.equal(a.action_task, b.action_task) and torch.equal(a.state, b.state)
    assert a.read_counters()["error"] == 0 and a.read_counters()["sac_updates"] >= 4
    q = make(use_recovery=True, mf_recovery=False, q_sampling_recovery=True, add_both_transitions=True)
    q.capture()
    assert q.graph is not None
    for _ in range(6):
        q.replay()
    torch.cuda.synchronize()
    c = q.read_counters()
    assert c["error"] == 0 and c["task_len"] >= 6 * 1024 and torch.isfinite(q.action_real).all()
    assert (q.action_real.abs() <= 0.1 + 1e-6).all()


def test_q_sampling_recovery_vs_reference_golden(native, cuda, golden_dir):
    """tests/golden/qsample_nav1.npz holds the candidates the REFERENCE's QRiskWrapper.select_action drew (qrisk.py:214-225, gym
    Box shim of the harness) and the actions it returned: the drop-in QRiskWrapper (N = 1 path) fed the same candidates returns
    the same actions, and so does the vector kernel for 8 env copies at once."""
    import argparse
    from env.spaces import Box
    from oracle.agent import Agent
    from recovery_rl.sac import SAC
    z = np.load(os.path.join(golden_dir, "qsample_nav1.npz"))
    torch.manual_seed(int(z["seed"]))
    np.random.seed(int(z["seed"]))
    ora = Agent(action_scale=(np.float32(1.0),) * 2, gamma_safe=0.8, eps_safe=0.3)       # the golden's xavier init (test_oracle.py)
    args = argparse.Namespace(gamma=0.99, tau=0.005, alpha=0.2, env_name="navigation1", policy="Gaussian", target_update_interval=1,
                              automatic_entropy_tuning=False, gamma_safe=0.8, eps_safe=0.3, nu=0.01, batch_size=64,
                              lr=3e-4, tau_safe=0.0002, MF_recovery=False, Q_sampling_recovery=True, hidden_size=256,
                              DGD_constraints=False, update_nu=False, RCPO=False, use_constraint_sampling=False,
                              lambda_RCPO=0.01, pos_fraction=-1, cnn=False, vismpc_recovery=False)
    one = np.float32(1.0)
    agent = SAC(Box(-np.ones(2) * np.inf, np.ones(2) * np.inf), Box(-np.ones(2) * one, np.ones(2) * one), args, "/tmp/none")
    agent.arena.load_modules(ora.nets())
    qr = agent.safety_critic
    n = len(z["states"])
    sure = z["gap"] > 1e-5                       # runner-up further away than the fp32 ordering noise of two implementations

    class Replayed(object):
        def __init__(self, cands):
            self.c = list(cands)

        def sample(self):
            return self.c.pop(0)
    # (a) the drop-in QRiskWrapper
    for i in range(n):
        qr.ac_space = Replayed(z["candidates"][i])
        a = qr.select_action(z["states"][i])
        if sure[i]:
            assert np.array_equal(a, z["actions"][i]), i
    # (b) the vector kernel: every env copy flagged, candidates handed over as the uniforms they were drawn from
    ar = agent.arena
    dev = ar.arena.device
    state = torch.from_numpy(np.ascontiguousarray(z["states"].T)).to(dev)
    cand_u = torch.from_numpy(((z["candidates"].astype(np.float64) + 1.0) * 0.5).astype(np.float32)).to(dev).contiguous()
    a_real = torch.zeros(n, 2, device=dev)
    flags = torch.ones(n, dtype=torch.uint8, device=dev)
    ws = torch.empty(native.select_workspace_floats(3, 1000), device=dev)      # 3 env copies per chunk: 3 chunks
    native.qsample_recovery_action(ar.cfg, ar.arena, n, 1000, state, ar.counters, ws, a_real, recovery=flags, cand_u=cand_u)
    got = a_real.cpu().numpy()
    assert sure.sum() >= 6
    assert np.allclose(got[sure], z["actions"][sure], rtol=0, atol=2e-7), np.abs(got - z["actions"]).max(1)
    flags[1::2] = 0                              # unflagged env copies keep their action
    a_real.fill_(7.0)
    native.qsample_recovery_action(ar.cfg, ar.arena, n, 1000, state, ar.counters, ws, a_real, recovery=flags, cand_u=cand_u)
    got = a_real.cpu().numpy()
    assert (got[1::2] == 7.0).all() and np.allclose(got[0::2][sure[0::2]], z["actions"][0::2][sure[0::2]], rtol=0, atol=2e-7)
