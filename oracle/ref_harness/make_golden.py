"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference does not exist on
the GPU box):   python -m oracle.ref_harness.make_golden

Every array written here is an output of the reference's own code (through the shims and
patches P1-P4 of harness.py; SAC update ordering = Variant B).  The oracle restatement in
oracle/*.py and the CUDA path are both tested against these files.
"""
import os
import sys
import random
import hashlib

import numpy as np
import torch

from . import harness

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(ROOT, "tests", "golden")


def _save(name, **arrays):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **arrays)
    print("wrote %-34s %7.1f KB" % (name, os.path.getsize(path) / 1024.0))


# ---------------------------------------------------------------------------------------
# A. Navigation step triples (reference env/navigation1.py:71-110, navigation2.py, obstacle.py)
# ---------------------------------------------------------------------------------------
def golden_nav_steps():
    import env.navigation1 as n1
    import env.navigation2 as n2
    rng = np.random.RandomState(20261017)
    for name, mod, cls in (("nav1", n1, n1.Navigation1), ("nav2", n2, n2.Navigation2)):
        env = cls()
        env.reset()
        n = 6000
        # broad cloud around the corridor / obstacle, near-goal cluster, exact obstacle edges
        s = np.empty((n, 2))
        s[:, 0] = rng.uniform(-110, 30, n)
        s[:, 1] = rng.uniform(-14, 14, n)
        s[:800] = rng.randn(800, 2) * 3.0                       # near goal (cost > -4 region)
        s[800:1200] = (-50, 0) + rng.randn(400, 2)              # reset distribution
        edges_x = [-100.0, -80.0, 150.0, -30.0, -20.0]
        edges_y = [5.0, 10.0, -5.0, -10.0, 7.5, -7.5]
        k = 1200
        for ex in edges_x:
            for ey in edges_y:
                for dx in (0.0, np.nextafter(0.0, 1.0), -1e-12, 1e-12):
                    s[k] = (ex + dx, ey - dx)
                    k += 1
        # states that land exactly on an edge after the update need (s+a)+noise == edge: craft
        a = rng.uniform(-1.6, 1.6, (n, 2)).astype(np.float32)    # beyond +-1 -> exercises the clip
        noise = rng.randn(n, 2)
        s[k:k + 200, 1] = 5.0 - a[k:k + 200, 1].clip(-1, 1).astype(np.float64)
        noise[k:k + 200, 1] = 0.0
        s[k:k + 200, 0] = rng.uniform(-70, 20, 200)
        ns = np.empty((n, 2))
        cost = np.empty(n)
        done = np.empty(n, np.uint8)
        cons = np.empty(n, np.uint8)
        succ = np.empty(n, np.uint8)
        orig_randn = np.random.randn
        devnull = open(os.devnull, "w")
        for i in range(n):
            env.state = s[i].copy()
            np.random.randn = lambda *sh, _i=i: noise[_i].copy()
            so = sys.stdout
            sys.stdout = devnull                                 # "obs ..." prints
            try:
                o, c, d, info = env.step(a[i])
            finally:
                sys.stdout = so
            ns[i] = o
            cost[i] = c
            done[i] = bool(d)
            cons[i] = int(info["constraint"])
            succ[i] = bool(info["success"])
        np.random.randn = orig_randn
        _save("nav_step_%s.npz" % name, state=s, action=a, noise=noise, next_state=ns,
              reward=cost, done=done, constraint=cons, success=succ)


# ---------------------------------------------------------------------------------------
# A'. offline-data generators (reference navigation1.py:133-164, navigation2.py:133-243)
# ---------------------------------------------------------------------------------------
def golden_offline_data():
    import env.navigation1 as n1
    import env.navigation2 as n2
    for name, mod in (("nav1", n1), ("nav2", n2)):
        np.random.seed(5)
        so = sys.stdout
        sys.stdout = open(os.devnull, "w")
        try:
            tr = mod.get_offline_data(2000)
        finally:
            sys.stdout = so
        s = np.array([t[0] for t in tr])
        a = np.array([t[1] for t in tr])
        c = np.array([float(t[2]) for t in tr])
        s2 = np.array([t[3] for t in tr])
        m = np.array([float(t[4]) for t in tr])
        _save("offline_%s.npz" % name, seed=np.int64(5), num=np.int64(2000), state=s, action=a,
              constraint=c, next_state=s2, mask=m)


# ---------------------------------------------------------------------------------------
# B. replay index streams (reference replay_memory.py:11-75 + CPython random)
# ---------------------------------------------------------------------------------------
def golden_replay():
    from recovery_rl.replay_memory import ReplayMemory, ConstraintReplayMemory
    out = {}
    cases = []
    #        name      cap   pushes-before-each-sample-burst              B    pos_fraction  seed
    cases.append(("poolset", 5000, [300, 700, 45, 1, 3000, 2500], 256, None, 7))
    cases.append(("strat",   4000, [600, 800, 2000, 1500],          256, 0.3, 123456))
    cases.append(("b1024",   9000, [1100, 3000, 16, 1, 6000],       1024, None, 2 ** 40 + 5))
    cases.append(("strat1k", 6000, [1500, 2700, 3000],              1024, 0.25, 3))
    for name, cap, bursts, B, pf, seed in cases:
        prng = np.random.RandomState(99)
        mem = ReplayMemory(cap, seed)
        cmem = ConstraintReplayMemory(cap, seed)
        counter = 0
        flags_all = []
        task_idx, cons_idx, lens = [], [], []
        for burst in bursts:
            for _ in range(burst):
                f = float(prng.rand() < 0.25)
                flags_all.append(f)
                st = np.array([float(counter), 0.0])
                mem.push(st, np.zeros(2, np.float32), -1.0, st + 1, 1.0)
                cmem.push(st, np.zeros(2, np.float32), f, st + 1, 1.0)
                counter += 1
            for _rep in range(3):      # SAC sample then Q_risk sample, one shared global stream
                s, *_ = mem.sample(min(B, len(mem)))
                task_idx.append(s[:, 0].astype(np.int64))
                bq = min(B, int((1 - pf) * len(cmem))) if pf else min(B, len(cmem))
                s, *_ = cmem.sample(bq, pos_fraction=pf)
                cons_idx.append(s[:, 0].astype(np.int64))
                lens.append(counter)
        out[name + "_cap"] = np.int64(cap)
        out[name + "_B"] = np.int64(B)
        out[name + "_pf"] = np.float64(-1.0 if pf is None else pf)
        out[name + "_seed"] = np.int64(seed)
        out[name + "_bursts"] = np.array(bursts, np.int64)
        out[name + "_flags"] = np.array(flags_all, np.float32)
        out[name + "_pushed"] = np.array(lens, np.int64)
        # ids are global push counters; slot = id % cap
        out[name + "_task_ids"] = np.concatenate(task_idx)
        out[name + "_task_sizes"] = np.array([len(x) for x in task_idx], np.int64)
        out[name + "_cons_ids"] = np.concatenate(cons_idx)
        out[name + "_cons_sizes"] = np.array([len(x) for x in cons_idx], np.int64)
    out["names"] = np.array([c[0] for c in cases])
    # raw CPython random KATs
    random.seed(1)
    out["kat_seed1_u32"] = np.array([random.getrandbits(32) for _ in range(8)], np.int64)
    random.seed(123456)
    out["kat_123456_1000_8"] = np.array(random.sample(range(1000), 8), np.int64)
    random.seed(1)
    out["kat_1_300_256"] = np.array(random.sample(range(300), 256), np.int64)
    random.seed(1)
    out["kat_1_20000_256"] = np.array(random.sample(range(20000), 256), np.int64)
    _save("replay_idx.npz", **out)


# ---------------------------------------------------------------------------------------
# C/D/E. network init, SAC / Q_risk updates, acting  (sac.py, qrisk.py, model.py)
# ---------------------------------------------------------------------------------------
class _FixedMemory(object):
    def __init__(self):
        self.batch = None

    def sample(self, batch_size, pos_fraction=None):
        s, a, r, s2, m = self.batch
        assert len(s) == batch_size
        return s, a, r, s2, m

    def __len__(self):
        return 10 ** 6


class _DummyEnv(object):
    def reset(self, pos=()):
        return np.zeros(2)


def _params(mod):
    return [p.detach().numpy().copy() for p in mod.parameters()]


STRIDE = 5      # after-update weights / grads are stored as every STRIDE-th element (fixture size)


def _sub(x, stride):
    return x.ravel()[::stride].copy() if stride > 1 else x


def _dump_agent(agent, prefix, out, stride=1, nets=None):
    for nm, mod in (("critic", agent.critic), ("critic_target", agent.critic_target),
                    ("policy", agent.policy), ("qrisk", agent.safety_critic.safety_critic),
                    ("qrisk_target", agent.safety_critic.safety_critic_target),
                    ("recovery", agent.safety_critic.policy)):
        if nets is not None and nm not in nets:
            continue
        for i, p in enumerate(_params(mod)):
            out["%s%s_%d" % (prefix, nm, i)] = _sub(p, stride)


def golden_agent(tag, env_name, scale, argv, B, n_updates=3, seed=11, dump_init=True):
    harness.setup()
    from gym.spaces import Box
    from recovery_rl.sac import SAC
    args = harness.get_args(argv + ["--env-name", env_name, "--seed", str(seed),
                                    "--batch_size", str(B)])
    torch.manual_seed(seed)
    np.random.seed(seed)
    obs_space = Box(-np.ones(2) * float("inf"), np.ones(2) * float("inf"))
    act_space = Box(-np.ones(2) * scale, np.ones(2) * scale)
    agent = SAC(obs_space, act_space, args, "/tmp/none", tmp_env=_DummyEnv())
    qr = agent.safety_critic
    out = {"seed": np.int64(seed), "B": np.int64(B), "scale": np.float64(scale),
           "gamma": np.float64(args.gamma), "gamma_safe": np.float64(args.gamma_safe),
           "alpha": np.float64(args.alpha), "tau": np.float64(args.tau),
           "tau_safe": np.float64(args.tau_safe), "lr": np.float64(args.lr),
           "eps_safe": np.float64(args.eps_safe), "n_updates": np.int64(n_updates)}
    # targets are hard copies of their sources at construction (sac.py:86, qrisk.py:60)
    _dump_agent(agent, "init_", out, nets=("critic", "policy", "qrisk", "recovery") if dump_init else ())
    sha = hashlib.sha256()
    for mod in (agent.critic, agent.policy, qr.safety_critic, qr.policy):
        for p in mod.parameters():
            sha.update(p.detach().numpy().tobytes())
    out["init_sha256"] = np.array(sha.hexdigest())
    out["stride"] = np.int64(STRIDE)

    rng = np.random.RandomState(1234 + seed)
    if env_name == "maze":
        def states(n):
            return rng.uniform(-0.27, 0.27, (n, 2))
        step_scale = 0.02
    else:
        def states(n):
            return np.stack([rng.uniform(-75, 10, n), rng.uniform(-9, 9, n)], 1)
        step_scale = 1.0

    mem = _FixedMemory()
    for u in range(n_updates):
        # ---- SAC update ----
        s = states(B)
        a = rng.uniform(-scale, scale, (B, 2)).astype(np.float32)
        s2 = s + step_scale * a.astype(np.float64) / scale + 0.05 * step_scale * rng.randn(B, 2)
        r = -np.linalg.norm(s, axis=1)
        m = (rng.rand(B) > 0.1).astype(np.float64)
        e_next = rng.randn(B, 2).astype(np.float32)
        e_cur = rng.randn(B, 2).astype(np.float32)
        mem.batch = (s, a, r, s2, m)
        harness.eps_queue.extend([e_next, e_cur])
        losses = agent.update_parameters(mem, B, u, safety_critic=qr)
        assert not harness.eps_queue
        p = "sac%d_" % u
        out.update({p + "s": s, p + "a": a, p + "r": r, p + "s2": s2, p + "m": m,
                    p + "eps_next": e_next, p + "eps_cur": e_cur,
                    p + "losses": np.array(losses, np.float64)})
        for k, v in agent._dbg.items():
            if isinstance(v, list):
                if u == 0:
                    for i, g in enumerate(v):
                        out["%s%s_%d" % (p, k, i)] = _sub(g, STRIDE)
            else:
                out[p + k] = v
        if u in (0, n_updates - 1):
            _dump_agent(agent, "after_sac%d_" % u, out, STRIDE, nets=("critic", "critic_target", "policy"))

        # ---- Q_risk update (+ MF recovery policy) ----
        s = states(B)
        a = rng.uniform(-scale, scale, (B, 2)).astype(np.float32)
        s2 = s + step_scale * a.astype(np.float64) / scale + 0.05 * step_scale * rng.randn(B, 2)
        c = (rng.rand(B) < 0.3).astype(np.float64)
        m = 1.0 - c
        e_next = rng.randn(B, 2).astype(np.float32)
        e_rec = rng.randn(B, 2).astype(np.float32)
        mem.batch = (s, a, c, s2, m)
        # shadow forward on the pre-update weights for intermediate values
        with torch.no_grad():
            harness.eps_queue.append(e_next)
            st = torch.FloatTensor(s)
            at = torch.FloatTensor(a)
            s2t = torch.FloatTensor(s2)
            na, _, _ = agent.policy.sample(s2t)
            q1t, q2t = qr.safety_critic_target(s2t, na)
            tgt = torch.FloatTensor(c).unsqueeze(1) + torch.FloatTensor(m).unsqueeze(1) * \
                qr.gamma_safe * torch.max(q1t, q2t)
            q1, q2 = qr.safety_critic(st, at)
            l1 = torch.nn.functional.mse_loss(q1, tgt).item()
            l2 = torch.nn.functional.mse_loss(q2, tgt).item()
        harness.eps_queue.extend([e_next, e_rec])
        qr.update_parameters(memory=mem, policy=agent.policy, batch_size=B)
        assert not harness.eps_queue
        p = "qr%d_" % u
        out.update({p + "s": s, p + "a": a, p + "c": c, p + "s2": s2, p + "m": m,
                    p + "eps_next": e_next, p + "eps_rec": e_rec,
                    p + "q1": q1.numpy().copy(), p + "q2": q2.numpy().copy(),
                    p + "target": tgt.numpy().copy(), p + "next_action": na.numpy().copy(),
                    p + "losses": np.array([l1, l2], np.float64)})
        # recovery-policy loss on the POST-step critic (qrisk.py:150-155), recomputed here
        if u in (0, n_updates - 1):
            _dump_agent(agent, "after_qr%d_" % u, out, STRIDE, nets=("qrisk", "qrisk_target", "recovery"))

    # ---- acting on the final weights (experiment.py:546-577, sac.py:133-168, qrisk.py:184-213) ----
    N = 777
    s = states(N)
    e_task = rng.randn(N, 2).astype(np.float32)
    e_rec = rng.randn(N, 2).astype(np.float32)
    with torch.no_grad():
        st = torch.FloatTensor(s)
        harness.eps_queue.append(e_task)
        a_task, logp, a_mean = agent.policy.sample(st)
        qv = qr.get_value(st, a_task)
        harness.eps_queue.append(e_rec)
        a_rec, _, a_rec_mean = qr.policy.sample(st)
        # eval-mode threshold uses the mean action (sac.py:166-167)
        qv_mean = qr.get_value(st, a_mean)
    # eps_safe for this fixture = median Q_risk so that both branches of experiment.py:555 occur
    act_thresh = float(np.float32(np.median(qv.numpy())))
    out["act_thresh"] = np.float64(act_thresh)
    rec = (qv.numpy()[:, 0] > act_thresh)
    real = np.where(rec[:, None], a_rec.numpy(), a_task.numpy())
    out.update({"act_s": s, "act_eps_task": e_task, "act_eps_rec": e_rec,
                "act_task": a_task.numpy().copy(), "act_logp": logp.numpy().copy(),
                "act_mean": a_mean.numpy().copy(), "act_qrisk": qv.numpy().copy(),
                "act_qrisk_mean": qv_mean.numpy().copy(),
                "act_rec": a_rec.numpy().copy(), "act_rec_mean": a_rec_mean.numpy().copy(),
                "act_recovery": rec.astype(np.uint8), "act_real": real})
    _save("agent_%s.npz" % tag, **out)


# ---------------------------------------------------------------------------------------
# F. N = 1 whole-trajectory trace through Experiment (experiment.py:356-491)
# ---------------------------------------------------------------------------------------
def golden_trajectory(env_name="navigation1", seed=7, gamma_safe="0.8", eps_safe="0.3", n_eps=12,
                      fname="traj_nav1_seed7.npz", algo=("--use_recovery", "--MF_recovery"), stride=None):
    harness.setup()
    import recovery_rl.replay_memory as rm
    argv = ["--env-name", env_name] + list(algo) + ["--gamma_safe", gamma_safe,
            "--eps_safe", eps_safe, "--num_eps", str(n_eps), "--num_unsafe_transitions", "2000",
            "--critic_safe_pretraining_steps", "30", "--seed", str(seed), "--batch_size", "16",
            "--logdir", "/tmp/rrl_golden_runs"]
    idx_log = []
    orig_sample = random.sample

    def rec_sample(pop, k):
        idx = orig_sample(range(len(pop)), k)
        idx_log.append(np.array(idx, np.int64))
        return [pop[i] for i in idx]

    rm.random.sample = rec_sample
    env_noise = []
    orig_randn = np.random.randn

    def rec_randn(*shape):
        x = orig_randn(*shape)
        env_noise.append(np.array(x, np.float64).ravel().copy())
        return x

    cat_log = []
    orig_cat = torch.distributions.Categorical.sample

    def rec_cat(self, *a, **k):
        j = orig_cat(self, *a, **k)
        cat_log.append(int(j))
        return j

    torch.distributions.Categorical.sample = rec_cat
    so = sys.stdout
    sys.stdout = open(os.devnull, "w")
    try:
        exp = harness.make_experiment(argv)
        offline = exp.constraint_demo_data
        rand_actions = []
        orig_as = exp.env.action_space.sample

        def rec_as():
            x = orig_as()
            rand_actions.append(x.copy())
            return x

        exp.env.action_space.sample = rec_as
        del harness.eps_log[:]
        if "Deterministic" in algo:
            # DeterministicPolicy.sample draws with self.noise.normal_ (model.py:478), not through Normal.rsample: record
            # the raw vector it leaves in self.noise after every call, in call order with the other draws
            pol = exp.agent.policy
            orig_ps = pol.sample

            def rec_ps(state):
                res = orig_ps(state)
                harness.eps_log.append(pol.noise.detach().cpu().numpy().reshape(1, 2).copy())
                return res
            pol.sample = rec_ps
        np.random.randn = rec_randn
        cfg = exp.exp_cfg
        if cfg.use_recovery or cfg.DGD_constraints or cfg.RCPO:     # experiment.py:357-361
            exp.pretrain_critic_recovery()
        n_pre_eps = len(harness.eps_log)
        n_pre_idx = len(idx_log)
        infos = []
        ep_len = []
        for ep in range(1, n_eps + 1):
            info = exp.get_train_rollout(ep)
            infos += info
            ep_len.append(len(info))
    finally:
        sys.stdout = so
        np.random.randn = orig_randn
        rm.random.sample = orig_sample
        torch.distributions.Categorical.sample = orig_cat
    sha = hashlib.sha256()
    for mod in (exp.agent.critic, exp.agent.policy, exp.agent.safety_critic.safety_critic,
                exp.agent.safety_critic.policy):
        for p in mod.parameters():
            sha.update(p.detach().numpy().tobytes())
    out = {"argv": np.array(argv), "ep_len": np.array(ep_len, np.int64),
           "state": np.array([i["state"] for i in infos]),
           "next_state": np.array([i["next_state"] for i in infos]),
           "action": np.array([i["action"] for i in infos], np.float32),
           "reward": np.array([i["reward"] for i in infos]),
           "constraint": np.array([int(i["constraint"]) for i in infos], np.uint8),
           "success": np.array([bool(i["success"]) for i in infos], np.uint8),
           "recovery": np.array([bool(i["recovery"]) for i in infos], np.uint8),
           "env_noise": np.array(env_noise), "rand_actions": np.array(rand_actions, np.float32),
           "n_pre_eps": np.int64(n_pre_eps), "n_pre_idx": np.int64(n_pre_idx),
           "eps_sizes": np.array([e.shape[0] for e in harness.eps_log], np.int64),
           "eps": np.concatenate([e.reshape(-1, 2) for e in harness.eps_log]).astype(np.float32),
           "idx_sizes": np.array([len(x) for x in idx_log], np.int64),
           "idx": np.concatenate(idx_log),
           "offline_state": np.array([t[0] for t in offline]),
           "offline_action": np.array([t[1] for t in offline]),
           "offline_constraint": np.array([float(t[2]) for t in offline]),
           "offline_next_state": np.array([t[3] for t in offline]),
           "offline_mask": np.array([float(t[4]) for t in offline]),
           "num_viols": np.int64(exp.num_viols), "num_successes": np.int64(exp.num_successes),
           "total_numsteps": np.int64(exp.total_numsteps), "updates": np.int64(exp.updates),
           "cat_idx": np.array(cat_log, np.int64),     # SQRL: every Categorical draw of the action filter, in order
           "weights_sha256": np.array(sha.hexdigest())}
    stride = STRIDE if stride is None else stride
    _dump_agent(exp.agent, "final_", out, stride)
    out["stride"] = np.int64(stride)
    _save(fname, **out)


# ---------------------------------------------------------------------------------------
# G. comparison-algorithm branches of SAC (sac.py:139-161, 202-205, 221-228, 241-271; model.py:447-485):
#    LR (--DGD_constraints --update_nu), RSPO (--DGD_constraints, scheduled nu), RCPO, automatic entropy
#    tuning, --policy Deterministic, and the SQRL action filter (--use_constraint_sampling)
# ---------------------------------------------------------------------------------------
ALGO_CASES = [
    # tag        env            scale  extra argv
    ("lr",      "navigation1", 1.0, ["--DGD_constraints", "--nu", "5000", "--update_nu", "--gamma_safe", "0.8", "--eps_safe", "0.3"]),
    ("rspo",    "navigation1", 1.0, ["--DGD_constraints", "--nu_schedule", "--nu_start", "10000", "--gamma_safe", "0.8", "--eps_safe", "0.3"]),
    ("rcpo",    "maze",        0.1, ["--RCPO", "--lambda", "50", "--gamma_safe", "0.5", "--eps_safe", "0.15", "--pos_fraction", "0.3"]),
    ("autoalpha", "maze",      0.1, ["--automatic_entropy_tuning", "1"]),
    ("det",     "navigation1", 1.0, ["--policy", "Deterministic"]),
    ("sqrl",    "maze",        0.1, ["--DGD_constraints", "--use_constraint_sampling", "--nu", "100", "--update_nu", "--gamma_safe", "0.5", "--eps_safe", "0.15", "--pos_fraction", "0.3"]),
]


def golden_algos(B=64, n_qr=12, n_updates=3, seed=23):
    harness.setup()
    from gym.spaces import Box
    from recovery_rl.sac import SAC
    from recovery_rl.utils import linear_schedule
    STR = 13
    out = {"tags": np.array([c[0] for c in ALGO_CASES]), "B": np.int64(B), "n_qr": np.int64(n_qr),
           "n_updates": np.int64(n_updates), "seed": np.int64(seed), "stride": np.int64(STR)}
    for tag, env_name, scale, extra in ALGO_CASES:
        args = harness.get_args(extra + ["--env-name", env_name, "--seed", str(seed), "--batch_size", str(B)])
        torch.manual_seed(seed)
        np.random.seed(seed)
        obs_space = Box(-np.ones(2) * float("inf"), np.ones(2) * float("inf"))
        act_space = Box(-np.ones(2) * scale, np.ones(2) * scale)
        agent = SAC(obs_space, act_space, args, "/tmp/none", tmp_env=_DummyEnv())
        qr = agent.safety_critic
        P = tag + "_"
        out[P + "env"] = np.array(env_name)
        out[P + "scale"] = np.float64(scale)
        for k in ("gamma", "gamma_safe", "alpha", "tau", "tau_safe", "lr", "eps_safe", "nu", "lambda_RCPO"):
            out[P + k] = np.float64(getattr(args, k))
        out[P + "flags"] = np.array([int(args.DGD_constraints), int(args.update_nu), int(args.RCPO),
                                     int(bool(args.automatic_entropy_tuning) and args.policy == "Gaussian"),
                                     int(args.policy != "Gaussian"), int(args.use_constraint_sampling)], np.int64)
        out[P + "mf_recovery"] = np.int64(int(args.MF_recovery))
        if args.nu_schedule:
            sched = linear_schedule(args.nu_start, args.nu_end, args.num_eps)
        else:
            sched = linear_schedule(args.nu, args.nu, 0)
        rng = np.random.RandomState(4321 + seed)
        if env_name == "maze":
            def states(n):
                return rng.uniform(-0.27, 0.27, (n, 2))
            step_scale = 0.02
        else:
            def states(n):
                return np.stack([rng.uniform(-75, 10, n), rng.uniform(-9, 9, n)], 1)
            step_scale = 1.0
        mem = _FixedMemory()
        # a few Q_risk updates first so that the safety critic is not at its (flat) initialisation
        for u in range(n_qr):
            s = states(B)
            a = rng.uniform(-scale, scale, (B, 2)).astype(np.float32)
            s2 = s + step_scale * a.astype(np.float64) / scale + 0.05 * step_scale * rng.randn(B, 2)
            c = (s[:, 1] * (1.0 if env_name != "maze" else 30.0) + rng.randn(B) > 2.0).astype(np.float64)
            m = 1.0 - c
            e_next = rng.randn(B, 2).astype(np.float32)
            mem.batch = (s, a, c, s2, m)
            if args.policy == "Gaussian":
                harness.eps_queue.append(e_next)
            else:
                torch.manual_seed(500 + u)
            if args.MF_recovery:
                harness.eps_queue.append(rng.randn(B, 2).astype(np.float32))
            qr.update_parameters(memory=mem, policy=agent.policy, batch_size=B)
            assert not harness.eps_queue
            q = "%sqr%d_" % (P, u)
            out.update({q + "s": s, q + "a": a, q + "c": c, q + "s2": s2, q + "m": m, q + "eps_next": e_next})
        for u in range(n_updates):
            s = states(B)
            a = rng.uniform(-scale, scale, (B, 2)).astype(np.float32)
            s2 = s + step_scale * a.astype(np.float64) / scale + 0.05 * step_scale * rng.randn(B, 2)
            r = -np.linalg.norm(s, axis=1)
            m = (rng.rand(B) > 0.1).astype(np.float64)
            e_next = rng.randn(B, 2).astype(np.float32)
            e_cur = rng.randn(B, 2).astype(np.float32)
            mem.batch = (s, a, r, s2, m)
            nu_arg = float(sched(1 + 40 * u))           # experiment.py:406 nu=self.nu_schedule(i_episode)
            if args.policy == "Gaussian":
                harness.eps_queue.extend([e_next, e_cur])
            else:
                torch.manual_seed(1000 + u)              # DeterministicPolicy.sample draws self.noise.normal_ (model.py:478)
            losses = agent.update_parameters(mem, B, u, safety_critic=qr, nu=nu_arg)
            assert not harness.eps_queue
            q = "%ssac%d_" % (P, u)
            out.update({q + "s": s, q + "a": a, q + "r": r, q + "s2": s2, q + "m": m, q + "eps_next": e_next,
                        q + "eps_cur": e_cur, q + "nu_arg": np.float64(nu_arg), q + "losses": np.array(losses, np.float64)})
            for k, v in agent._dbg.items():
                if isinstance(v, list):
                    if u == 0:
                        for i, g in enumerate(v):
                            out["%s%s_%d" % (q, k, i)] = _sub(g, STR)
                else:
                    out[q + k] = np.asarray(v)
            out[q + "log_nu"] = np.float64(agent.log_nu.item())
            out[q + "log_lambda"] = np.float64(agent.log_lambda_RCPO.item())
            out[q + "lambda"] = np.float64(float(agent.lambda_RCPO))
            out[q + "alpha_after"] = np.float64(float(agent.alpha))
            if agent.automatic_entropy_tuning:
                out[q + "log_alpha"] = np.float64(agent.log_alpha.item())
            for nm, mod in (("critic", agent.critic), ("critic_target", agent.critic_target), ("policy", agent.policy)):
                for i, pp in enumerate(_params(mod)):
                    if u in (0, n_updates - 1):
                        out["%safter_%s_%d" % (q, nm, i)] = _sub(pp, STR)
        if args.use_constraint_sampling:
            # SQRL action filter (sac.py:139-161): 100 policy samples, keep Q_risk <= eps_safe, Categorical over exp(log_pi)
            n_sel = 24
            st = states(n_sel)
            eps = rng.randn(n_sel, 100, 2).astype(np.float32)
            acts = np.zeros((n_sel, 2), np.float32)
            thr = np.zeros(n_sel)
            with torch.no_grad():
                probe = qr.get_value(torch.FloatTensor(st), torch.zeros(n_sel, 2)).numpy()[:, 0]
            for i in range(n_sel):
                # thresholds around the actual Q_risk range so that all/some/none of the samples pass
                thr[i] = float(probe[i]) + (0.02 * (i % 3 - 1) if i % 4 else -1.0)
                agent.eps_safe = thr[i]
                harness.eps_queue.append(eps[i])
                torch.manual_seed(7000 + i)
                acts[i] = agent.select_action(st[i])
                assert not harness.eps_queue
            agent.eps_safe = args.eps_safe
            out.update({P + "sel_s": st, P + "sel_eps": eps, P + "sel_thresh": thr, P + "sel_action": acts})
    _save("agent_algos_b%d.npz" % B, **out)


# ---------------------------------------------------------------------------------------
# F. Q-sampling recovery (reference recovery_rl/qrisk.py:214-225): 1000 uniform candidate actions from the action
#    space, the one with the smallest max(Q1, Q2)_risk.  The networks are the xavier-initialised ones of
#    agent_nav1_b256.npz (same seed, same construction order: checked through the init sha256), so only the
#    candidates and the chosen actions are stored.
# ---------------------------------------------------------------------------------------
def golden_qsample(seed=11, n_states=8):
    harness.setup()
    from gym.spaces import Box
    from recovery_rl.sac import SAC
    args = harness.get_args(["--use_recovery", "--Q_sampling_recovery", "--gamma_safe", "0.8", "--eps_safe", "0.3",
                             "--env-name", "navigation1", "--seed", str(seed), "--batch_size", "256"])
    torch.manual_seed(seed)
    np.random.seed(seed)
    obs_space = Box(-np.ones(2) * float("inf"), np.ones(2) * float("inf"))
    act_space = Box(-np.ones(2) * 1.0, np.ones(2) * 1.0)
    agent = SAC(obs_space, act_space, args, "/tmp/none", tmp_env=_DummyEnv())
    qr = agent.safety_critic
    sha = hashlib.sha256()
    for mod in (agent.critic, agent.policy, qr.safety_critic, qr.policy):
        for p in mod.parameters():
            sha.update(p.detach().numpy().tobytes())
    ref = np.load(os.path.join(OUT, "agent_nav1_b256.npz"))
    assert sha.hexdigest() == str(ref["init_sha256"]), "construction order changed: the stored init weights are not these"
    rng = np.random.RandomState(4321)
    st = np.stack([rng.uniform(-75, 10, n_states), rng.uniform(-9, 9, n_states)], 1)
    qr.ac_space.seed(2026)
    recorded = []
    inner = qr.ac_space.sample

    def recording_sample():
        a = inner()
        recorded.append(a.copy())
        return a
    qr.ac_space.sample = recording_sample
    acts = np.zeros((n_states, 2), np.float32)
    for i in range(n_states):
        acts[i] = qr.select_action(st[i])
    cands = np.asarray(recorded, np.float32).reshape(n_states, 1000, 2)
    # the chosen action is one of the candidates; store its index and the margin to the runner-up for the tests' tie guard
    with torch.no_grad():
        idx = np.zeros(n_states, np.int64)
        gap = np.zeros(n_states)
        for i in range(n_states):
            sb = torch.FloatTensor(st[i]).unsqueeze(0).repeat(1000, 1)
            q = qr.get_value(sb, torch.FloatTensor(cands[i])).numpy().ravel()
            idx[i] = int(np.argmin(q))
            srt = np.sort(q)
            gap[i] = srt[1] - srt[0]
            assert np.array_equal(cands[i, idx[i]], acts[i])
    _save("qsample_nav1.npz", seed=np.int64(seed), states=st, candidates=cands, actions=acts, index=idx, gap=gap,
          init_sha256=np.array(sha.hexdigest()))


def golden_extra_trajectories():
    # the two recovery branches no script line uses, with eps_safe low enough that the barely trained safety critic
    # triggers recoveries in every episode: --add_both_transitions (experiment.py:446-448) and --Q_sampling_recovery
    # (qrisk.py:214-225; the 1000 candidates per recovery step are part of `rand_actions`, the env's action-space stream)
    golden_trajectory("navigation1", 8, "0.8", "0.05", 6, "traj_nav1_addboth.npz", stride=29,
                      algo=("--use_recovery", "--MF_recovery", "--add_both_transitions"))
    golden_trajectory("navigation1", 9, "0.8", "0.05", 6, "traj_nav1_qsample.npz", stride=29,
                      algo=("--use_recovery", "--Q_sampling_recovery"))
    # --policy Deterministic (model.py:447-485; alpha = 0, sac.py:115-117) under Recovery RL MF, start_steps 20 so that the
    # policy itself acts for most of the run
    golden_trajectory("navigation1", 10, "0.8", "0.3", 6, "traj_nav1_det.npz", stride=29,
                      algo=("--policy", "Deterministic", "--use_recovery", "--MF_recovery", "--start_steps", "20"))


def main():
    os.makedirs(OUT, exist_ok=True)
    harness.setup()
    golden_nav_steps()
    golden_offline_data()
    golden_replay()
    golden_agent("nav1_b256", "navigation1", 1.0,
                 ["--use_recovery", "--MF_recovery", "--gamma_safe", "0.8", "--eps_safe", "0.3"], 256)
    golden_agent("maze_b64", "maze", 0.1,
                 ["--use_recovery", "--MF_recovery", "--gamma_safe", "0.5", "--eps_safe", "0.15",
                  "--pos_fraction", "0.3"], 64, dump_init=False)
    golden_trajectory()
    golden_trajectory("navigation2", 3, "0.65", "0.2", 8, "traj_nav2_seed3.npz")   # scripts/navigation2.sh:7 settings
    golden_trajectory("navigation1", 2, "0.8", "0.3", 6, "traj_nav1_unconstrained.npz", algo=())   # navigation1.sh:21
    golden_trajectory("navigation1", 6, "0.8", "0.3", 6, "traj_nav1_rp.npz", algo=("--constraint_reward_penalty", "1000"))
    # comparison algorithms of scripts/navigation1.sh (LR :28, RSPO :35, RCPO :56), 6 episodes each
    golden_trajectory("navigation1", 3, "0.8", "0.3", 6, "traj_nav1_lr.npz", stride=29,
                      algo=("--DGD_constraints", "--nu", "5000", "--update_nu"))
    golden_trajectory("navigation1", 4, "0.8", "0.3", 6, "traj_nav1_rspo.npz", stride=29,
                      algo=("--DGD_constraints", "--nu_schedule", "--nu_start", "10000"))
    golden_trajectory("navigation1", 7, "0.8", "0.3", 6, "traj_nav1_rcpo.npz", stride=29, algo=("--RCPO", "--lambda", "1000"))
    # SQRL :49, with eps_safe 0.45 / start_steps 20 so that both branches of the action filter occur (6 Categorical draws,
    # 29 argmin fall-backs)
    golden_trajectory("navigation1", 5, "0.8", "0.45", 6, "traj_nav1_sqrl.npz", stride=29,
                      algo=("--DGD_constraints", "--use_constraint_sampling", "--nu", "5000", "--update_nu", "--start_steps", "20"))
    golden_algos()
    golden_qsample()
    golden_extra_trajectories()


if __name__ == "__main__":
    main()
