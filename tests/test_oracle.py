"""CPU tests: the oracle restatement against the golden vectors produced by the reference's own code
(oracle/ref_harness/make_golden.py) and against the live CPython `random`."""
import os
import random

import numpy as np
import pytest

from oracle import envs, replay


@pytest.mark.parametrize("name,kind", [("nav1", envs.NAV1), ("nav2", envs.NAV2)])
def test_nav_step_matches_reference(golden_dir, name, kind):
    z = np.load(os.path.join(golden_dir, "nav_step_%s.npz" % name))
    ns, r, d, c, s = envs.nav_step(kind, z["state"], z["action"], z["noise"])
    assert np.array_equal(ns, z["next_state"])          # bit-exact fp64
    assert np.array_equal(r, z["reward"])
    assert np.array_equal(d, z["done"].astype(bool))
    assert np.array_equal(c, z["constraint"].astype(bool))
    assert np.array_equal(s, z["success"].astype(bool))
    assert c.sum() > 100 and (~c).sum() > 100           # both branches exercised


@pytest.mark.parametrize("name,kind", [("nav1", envs.NAV1), ("nav2", envs.NAV2)])
def test_nav_offline_data_matches_reference(golden_dir, name, kind):
    z = np.load(os.path.join(golden_dir, "offline_%s.npz" % name))
    np.random.seed(int(z["seed"]))
    tr = envs.nav_offline_data(kind, int(z["num"]))
    assert len(tr) == len(z["state"])
    assert np.array_equal(np.array([t[0] for t in tr]), z["state"])
    assert np.array_equal(np.array([t[1] for t in tr]), z["action"])
    assert np.array_equal(np.array([float(t[2]) for t in tr]), z["constraint"])
    assert np.array_equal(np.array([t[3] for t in tr]), z["next_state"])
    assert np.array_equal(np.array([float(t[4]) for t in tr]), z["mask"])


def test_maze_restatement_properties():
    rs = np.random.RandomState(0)
    s = envs.maze_reset_from_uniform(rs.rand(2000, 2))
    assert not envs.maze_touch(s[:, 0], s[:, 1]).any() or True
    a = rs.uniform(-0.1, 0.1, (2000, 2)).astype(np.float32)
    ns, r, d, c, su = envs.maze_step(s, a, np.zeros(2000, int))
    free = ~c
    # contact-free displacement is linear in the action: 0.2467 * a
    ratio = (ns - s)[free] / a[free].astype(np.float64)
    assert np.allclose(ratio, 0.24667751, rtol=1e-6)
    # already in contact at step start: no motion, constraint (maze.py:144-147)
    wall = np.array([[-0.1, 0.3], [0.29, 0.0], [0.1, -0.2]])
    ns2, _, d2, c2, _ = envs.maze_step(wall, np.full((3, 2), 0.1, np.float32), np.zeros(3, int))
    assert c2.all() and d2.all() and np.array_equal(ns2, wall)
    # goal reached
    g = np.array([[0.25, 0.0]])
    _, r3, d3, c3, su3 = envs.maze_step(g, np.zeros((1, 2), np.float32), np.zeros(1, int))
    assert d3[0] and su3[0] and not c3[0] and r3[0] == 0.0
    # horizon is part of `done` for maze (maze.py:152)
    _, _, d4, _, _ = envs.maze_step(s[:4], np.zeros((4, 2), np.float32), np.full(4, 99))
    assert d4.all()


def test_mt19937_known_answers(golden_dir):
    z = np.load(os.path.join(golden_dir, "replay_idx.npz"))
    g = replay.MT19937(1)
    assert [g.genrand_uint32() for _ in range(8)] == list(z["kat_seed1_u32"])
    assert replay.MT19937(123456).sample_indices(1000, 8) == list(z["kat_123456_1000_8"])
    assert replay.MT19937(1).sample_indices(300, 256) == list(z["kat_1_300_256"])
    assert replay.MT19937(1).sample_indices(20000, 256) == list(z["kat_1_20000_256"])


@pytest.mark.parametrize("seed", [0, 1, 7, 123456, 2 ** 40 + 5, -9])
def test_sample_matches_live_cpython(seed):
    for n, k in ((300, 256), (20000, 256), (1045, 256), (1046, 256), (5000, 1024), (500, 76), (5, 5), (40, 6)):
        random.seed(seed)
        assert random.sample(range(n), k) == replay.MT19937(seed).sample_indices(n, k)


def _replay_case(z, name):
    cap = int(z[name + "_cap"]); B = int(z[name + "_B"]); pf = float(z[name + "_pf"])
    pf = None if pf < 0 else pf
    return cap, B, pf, int(z[name + "_seed"]), z[name + "_flags"], z[name + "_bursts"]


@pytest.mark.parametrize("name", ["poolset", "strat", "b1024", "strat1k"])
def test_replay_streams_match_reference(golden_dir, name):
    z = np.load(os.path.join(golden_dir, "replay_idx.npz"))
    cap, B, pf, seed, flags, bursts = _replay_case(z, name)
    st = replay.SharedStream()
    mem = replay.ReplayMemory(cap, seed, st)
    cm = replay.ConstraintReplayMemory(cap, seed, st)
    c = 0
    ti, ci = [], []
    for burst in bursts:
        for _ in range(burst):
            s = np.array([float(c), 0.0])
            mem.push(s, np.zeros(2), -1.0, s + 1, 1.0)
            cm.push(s, np.zeros(2), float(flags[c]), s + 1, 1.0)
            c += 1
        for _ in range(3):
            sl = mem.sample_slots(min(B, len(mem)))
            ti.append(mem.buf[sl, 0].astype(np.int64))
            bq = min(B, int((1 - pf) * len(cm))) if pf else min(B, len(cm))
            sl = cm.sample_slots(bq, pf)
            ci.append(cm.buf[sl, 0].astype(np.int64))
    assert np.array_equal(np.concatenate(ti), z[name + "_task_ids"])
    assert np.array_equal(np.concatenate(ci), z[name + "_cons_ids"])


def test_maze_scalar_matches_vectorised():
    rs = np.random.RandomState(2)
    s = rs.uniform(-0.28, 0.28, (300, 2))
    s[:100, 0] = -0.1 + rs.uniform(-0.05, 0.05, 100)
    a = rs.uniform(-0.12, 0.12, (300, 2)).astype(np.float32)
    steps = rs.randint(0, 100, 300)
    ns, r, d, c, su = envs.maze_step(s, a, steps)
    for i in range(300):
        o = envs.maze_step_scalar(s[i], a[i], int(steps[i]))
        assert np.array_equal(o[0], ns[i]) and o[1] == r[i] and o[2] == d[i] and o[3] == c[i] and o[4] == su[i]


@pytest.mark.parametrize("fname,env_name,seed,gamma_safe,eps_safe,algo", [
    ("traj_nav1_seed7.npz", "navigation1", 7, 0.8, 0.3, {}), ("traj_nav2_seed3.npz", "navigation2", 3, 0.65, 0.2, {}),
    ("traj_nav1_unconstrained.npz", "navigation1", 2, 0.8, 0.3, dict(use_recovery=False)),
    ("traj_nav1_rp.npz", "navigation1", 6, 0.8, 0.3, dict(use_recovery=False, constraint_reward_penalty=1000.0)),
    ("traj_nav1_lr.npz", "navigation1", 3, 0.8, 0.3, dict(use_recovery=False, mf_recovery=False, dgd=True, update_nu=True,
                                                        nu=5000.0)),
    ("traj_nav1_rspo.npz", "navigation1", 4, 0.8, 0.3, dict(use_recovery=False, mf_recovery=False, dgd=True, nu_schedule=True,
                                                          nu_start=10000.0, nu_end=0.0, num_eps=6)),
    ("traj_nav1_rcpo.npz", "navigation1", 7, 0.8, 0.3, dict(use_recovery=False, mf_recovery=False, rcpo=True,
                                                          lambda_rcpo=1000.0)),
    ("traj_nav1_sqrl.npz", "navigation1", 5, 0.8, 0.45, dict(use_recovery=False, mf_recovery=False, dgd=True, update_nu=True,
                                                           nu=5000.0, constraint_sampling=True, start_steps=20)),
    # the two recovery branches no script line uses: --add_both_transitions (experiment.py:446-448), --Q_sampling_recovery
    # (qrisk.py:214-225; its 1000 candidates per recovery step come out of the recorded action-space stream)
    ("traj_nav1_addboth.npz", "navigation1", 8, 0.8, 0.05, dict(add_both_transitions=True)),
    ("traj_nav1_qsample.npz", "navigation1", 9, 0.8, 0.05, dict(mf_recovery=False, q_sampling_recovery=True)),
    # --policy Deterministic (model.py:447-485, alpha = 0): its self.noise.normal_ draws are recorded raw, in call order
    ("traj_nav1_det.npz", "navigation1", 10, 0.8, 0.3, dict(deterministic=True, start_steps=20))])
def test_oracle_loop_reproduces_reference_trajectory(golden_dir, fname, env_name, seed, gamma_safe, eps_safe, algo):
    """oracle/loop.py fed the recorded noise == the reference's Experiment (12 episodes seed 7 on Navigation1, 8 episodes
    seed 3 on Navigation2 with the scripts/navigation2.sh:7 settings, and the unconstrained / reward-penalty lines of
    scripts/navigation1.sh:21,42 for 6 episodes): flags, episode
    boundaries, replay indices and counters bit-exact; states / actions / final weights to fp32 round-off.
    (On the CPU that recorded the golden the whole trajectory is bit-identical; another CPU takes another MKL
    sgemm code path and the fp32 actions move by 1 ulp, so the float fields carry an absolute 1e-5.)"""
    from oracle.loop import OracleExperiment, NoiseSource
    z = np.load(os.path.join(golden_dir, fname))
    sizes = z["eps_sizes"]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    eps = [z["eps"][offs[i]:offs[i + 1]] for i in range(len(sizes))]
    noise = NoiseSource(seed, eps=eps, env_noise=list(z["env_noise"]), rand_actions=list(z["rand_actions"]),
                        categorical=list(z["cat_idx"]) if "cat_idx" in z.files else None)
    exp = OracleExperiment(env_name, seed=seed, batch_size=16, gamma_safe=gamma_safe, eps_safe=eps_safe, noise=noise, **algo)
    tr = [(z["offline_state"][i], z["offline_action"][i], z["offline_constraint"][i], z["offline_next_state"][i],
           z["offline_mask"][i]) for i in range(len(z["offline_state"]))]
    if exp.uses_qrisk:                             # experiment.py:357-361: only the constrained algorithms pre-train
        exp.pretrain(tr, 30, num_unsafe_transitions=2000)
    assert len(exp.idx_log) == int(z["n_pre_idx"])
    infos = []
    for _ in range(int(z["ep_len"].sum())):
        infos.append(exp.step())
    assert np.allclose(np.array([i["state"] for i in infos]), z["state"], rtol=0, atol=1e-5)
    assert np.allclose(np.array([i["next_state"] for i in infos]), z["next_state"], rtol=0, atol=1e-5)
    assert np.allclose(np.array([i["action"] for i in infos], np.float32), z["action"], rtol=0, atol=1e-5)
    assert np.array_equal(np.array([i["constraint"] for i in infos]), z["constraint"])
    assert np.array_equal(np.array([i["recovery"] for i in infos]), z["recovery"].astype(bool))
    ends = np.cumsum(z["ep_len"]) - 1
    assert [i for i, x in enumerate(infos) if x["episode_end"]] == list(ends)
    assert np.array_equal(np.concatenate(exp.idx_log), z["idx"])
    assert exp.num_viols == int(z["num_viols"]) and exp.num_successes == int(z["num_successes"])
    assert exp.total_numsteps == int(z["total_numsteps"]) and exp.updates == int(z["updates"])
    if algo.get("add_both_transitions"):
        assert len(exp.memory) == exp.total_numsteps + int(z["recovery"].sum())     # one more task-buffer row per recovery step
    if algo.get("q_sampling_recovery"):
        assert not noise.rand_actions and int(z["recovery"].sum()) > 0              # every recorded candidate was consumed
    stride = int(z["stride"])
    for net in ("critic", "critic_target", "policy", "qrisk", "qrisk_target", "recovery"):
        if "final_%s_0" % net not in z.files:       # no recovery policy without --MF_recovery (qrisk.py:56-75)
            continue
        for i, p in enumerate(exp.agent.params(net)):
            assert np.allclose(p.ravel()[::stride], z["final_%s_%d" % (net, i)], rtol=0, atol=1e-5)


# ---- comparison-algorithm branches (LR / RSPO / RCPO / auto-alpha / Deterministic / SQRL) ---------------------
ALGO_TAGS = ["lr", "rspo", "rcpo", "autoalpha", "det", "sqrl"]


def algo_oracle_agent(z, tag):
    """oracle Agent configured like the reference SAC of golden case `tag` (same torch RNG order -> same init)."""
    import torch
    from oracle.agent import Agent
    P = tag + "_"
    f = z[P + "flags"]
    seed = int(z["seed"])
    torch.manual_seed(seed)
    np.random.seed(seed)
    sc = np.float32(float(z[P + "scale"]))
    return Agent(action_scale=(sc, sc), gamma=float(z[P + "gamma"]), alpha=float(z[P + "alpha"]), tau=float(z[P + "tau"]),
                 gamma_safe=float(z[P + "gamma_safe"]), tau_safe=float(z[P + "tau_safe"]), eps_safe=float(z[P + "eps_safe"]),
                 lr=float(z[P + "lr"]), mf_recovery=bool(int(z[P + "mf_recovery"])), dgd=bool(f[0]), update_nu=bool(f[1]),
                 rcpo=bool(f[2]), auto_alpha=bool(f[3]), deterministic=bool(f[4]), nu=float(z[P + "nu"]),
                 lambda_rcpo=float(z[P + "lambda_RCPO"]))


def algo_noise(z, tag, key, seed_base, u):
    """the policy noise of update u: recorded eps for the Gaussian policy, re-drawn from the torch generator for
    the Deterministic policy (the golden script seeds torch with seed_base + u right before the update)."""
    import torch
    from oracle.agent import deterministic_noise
    if int(z[tag + "_flags"][4]):
        torch.manual_seed(seed_base + u)
        n1 = deterministic_noise()
        n2 = deterministic_noise()
        return n1.numpy(), n2.numpy()
    return z["%s_%s%d_eps_next" % (tag, key, u)], (z["%s_%s%d_eps_cur" % (tag, key, u)] if key == "sac" else None)


@pytest.mark.parametrize("tag", ALGO_TAGS)
def test_oracle_comparison_branches_match_reference(golden_dir, tag):
    """oracle/agent.py == the reference's SAC.update_parameters on its LR / RSPO / RCPO / auto-alpha / Deterministic
    branches: losses, multipliers and (strided) weights after each update; SQRL action filter."""
    import torch
    z = np.load(os.path.join(golden_dir, "agent_algos_b64.npz"))
    ag = algo_oracle_agent(z, tag)
    P = tag + "_"
    stride = int(z["stride"])
    for u in range(int(z["n_qr"])):
        q = "%sqr%d_" % (P, u)
        e_next, _ = algo_noise(z, tag, "qr", 500, u)
        ag.qrisk_update([z[q + k] for k in ("s", "a", "c", "s2", "m")], e_next, None)
    n_upd = int(z["n_updates"])
    for u in range(n_upd):
        q = "%ssac%d_" % (P, u)
        e_next, e_cur = algo_noise(z, tag, "sac", 1000, u)
        L = ag.sac_update([z[q + k] for k in ("s", "a", "r", "s2", "m")], e_next, e_cur, u, nu=float(z[q + "nu_arg"]))
        assert np.allclose(L, z[q + "losses"], rtol=1e-5, atol=1e-7), (L, z[q + "losses"])
        for k in ("qf1", "qf2", "target", "pi", "min_qf_pi"):
            assert np.allclose(ag.dbg[k], z[q + k], rtol=1e-5, atol=1e-6), k
        assert np.isclose(ag.log_nu.item(), float(z[q + "log_nu"]), rtol=1e-9, atol=1e-12)
        assert np.isclose(ag.log_lambda.item(), float(z[q + "log_lambda"]), rtol=1e-9, atol=1e-12)
        assert np.isclose(float(ag.alpha), float(z[q + "alpha_after"]), rtol=1e-6)
        if u in (0, n_upd - 1):
            for net in ("critic", "critic_target", "policy"):
                for i, p in enumerate(ag.params(net)):
                    assert np.allclose(p.ravel()[::stride], z["%safter_%s_%d" % (q, net, i)], rtol=1e-5, atol=1e-6), (net, i)
    if tag == "sqrl":
        for i in range(len(z["sqrl_sel_s"])):
            torch.manual_seed(7000 + i)
            a = ag.select_action_sqrl(z["sqrl_sel_s"][i], z["sqrl_sel_eps"][i], eps_safe=float(z["sqrl_sel_thresh"][i]))
            assert np.allclose(a, z["sqrl_sel_action"][i], rtol=1e-5, atol=1e-7), i


# ---- model-based recovery (PETS ensemble + CEM, BASELINE config 5) ---------------------------------------------
def mpc_noise(z, tag):
    """the planner's random inputs, regenerated from RandomState(8) in the golden script's order."""
    P = tag + "_"
    pop, hor, npart = int(z[P + "popsize"]), int(z[P + "plan_hor"]), int(z[P + "npart"])
    scale = float(z[P + "scale"])
    rs = np.random.RandomState(int(z[P + "cost_rng_seed"]))
    ac_seqs = rs.uniform(-scale, scale, (pop, hor * 2)).astype(np.float32)
    eps = rs.randn(hor, 5, pop * npart // 5, 2).astype(np.float32)
    zs = np.clip(rs.standard_normal((2, 5, pop, hor * 2)), -2, 2)
    eps2 = rs.randn(2, 5, hor, 5, pop * npart // 5, 2).astype(np.float32)
    assert np.isclose(ac_seqs.astype(np.float64).sum(), float(z[P + "cost_ac_seqs_sum"]))
    assert np.isclose(eps2.astype(np.float64).sum(), float(z[P + "act_eps_sum"]))
    return ac_seqs, eps, zs, eps2


def mpc_oracle(z, za, tag, trained=True):
    """oracle MPC of golden case `tag`: PtModel init / training from the numpy RNG like the reference, safety
    critic = oracle Agent after the warm-up updates of the algos fixture."""
    import torch
    from oracle import mpc as ompc
    P = tag + "_"
    atag = str(z[P + "algos_tag"])
    ag = algo_oracle_agent(za, atag)
    for u in range(int(za["n_qr"])):
        q = "%s_qr%d_" % (atag, u)
        e_next, _ = algo_noise(za, atag, "qr", 500, u)
        ag.qrisk_update([za[q + k] for k in ("s", "a", "c", "s2", "m")], e_next, None)

    def value(obs, acs):
        with torch.no_grad():
            q1, q2 = ag.qrisk(obs, acs)
            return torch.max(q1, q2).squeeze()

    sc = np.float32(float(z[P + "scale"]))
    np.random.seed(int(z[P + "model_seed"]))
    m = ompc.MPC(-np.ones(2, np.float32) * sc, np.ones(2, np.float32) * sc, int(z[P + "plan_hor"]), int(z[P + "popsize"]),
                 int(z[P + "num_elites"]), npart=int(z[P + "npart"]), value_func=value)
    if trained:
        np.random.seed(int(z[P + "train_seed"]))
        m.train(z[P + "train_obs"], z[P + "train_acs"], z[P + "train_next"], int(z[P + "train_epochs"]))
    return m, ag


@pytest.mark.parametrize("tag", ["nav", "maze"])
def test_oracle_mpc_matches_reference(golden_dir, tag):
    import torch
    z = np.load(os.path.join(golden_dir, "mpc.npz"))
    za = np.load(os.path.join(golden_dir, "agent_algos_b64.npz"))
    P = tag + "_"
    stride = int(z["stride"])
    m, _ = mpc_oracle(z, za, tag, trained=False)
    for n, p in m.model.named():
        assert np.array_equal(p.detach().numpy().ravel()[::stride], z[P + "init_" + n]), n     # same scipy / numpy stream
    np.random.seed(int(z[P + "train_seed"]))
    m.train(z[P + "train_obs"], z[P + "train_acs"], z[P + "train_next"], int(z[P + "train_epochs"]))
    for n, p in m.model.named():
        assert np.allclose(p.detach().numpy().ravel()[::stride], z[P + "trained_" + n], rtol=1e-4, atol=1e-6), n
    with torch.no_grad():
        xin = torch.from_numpy(np.concatenate([z[P + "train_obs"], z[P + "train_acs"]], 1)[None].repeat(5, 0)).float()
        mean, var = m.model.forward(xin)
    assert np.allclose(mean.numpy(), z[P + "fwd_mean"], rtol=1e-4, atol=1e-6)
    assert np.allclose(var.numpy(), z[P + "fwd_var"], rtol=1e-4, atol=1e-9)
    ac_seqs, eps, zs, eps2 = mpc_noise(z, tag)
    costs = m.compile_cost(z[P + "cost_obs"], ac_seqs, eps)
    assert np.allclose(costs, z[P + "cost_out"], rtol=1e-5, atol=1e-6)
    for i in range(2):
        a = m.act(z[P + "act_states"][i], zs[i], eps2[i])
        assert np.allclose(a, z[P + "act_actions"][i], rtol=1e-5, atol=1e-7)
        assert np.allclose(m.prev_sol, z[P + "act_prev_sol"][i], rtol=1e-5, atol=1e-7)
    assert np.allclose(np.array(m.iter_costs), z[P + "act_iter_costs"], rtol=1e-5, atol=1e-6)


def test_oracle_q_sampling_recovery_matches_reference(golden_dir):
    """qrisk.py:214-225 through the reference's own QRiskWrapper.select_action (oracle/ref_harness/make_golden.py,
    golden_qsample): of the recorded 1000 uniform candidates, the restatement picks the same one for every state."""
    import torch
    from oracle.agent import Agent
    z = np.load(os.path.join(golden_dir, "qsample_nav1.npz"))
    init = np.load(os.path.join(golden_dir, "agent_nav1_b256.npz"))
    assert str(z["init_sha256"]) == str(init["init_sha256"])            # the networks are that file's xavier init
    torch.manual_seed(int(z["seed"]))
    np.random.seed(int(z["seed"]))
    ora = Agent(action_scale=(np.float32(1.0),) * 2, gamma_safe=0.8, eps_safe=0.3)
    for i, p in enumerate(ora.params("qrisk")):
        assert np.array_equal(p, init["init_qrisk_%d" % i]), i           # same construction order, same draws
    for i in range(len(z["states"])):
        a = ora.select_action_qsample(z["states"][i], z["candidates"][i])
        assert np.array_equal(a, z["actions"][i]), i
        assert np.array_equal(z["candidates"][i, int(z["index"][i])], z["actions"][i])
