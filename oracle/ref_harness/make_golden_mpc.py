"""Golden vectors for the model-based recovery path (BASELINE config 5): the reference's own PtModel
(config/maze.py:23-96), MPC (recovery_rl/MPC.py:213-467) and CEMOptimizer (recovery_rl/optimizers.py:73-124),
run on CPU through the harness.   python -m oracle.ref_harness.make_golden_mpc

TEST INFRASTRUCTURE ONLY.  Randomness is pinned the way the reference draws it: PtModel init and MPC.train use the
numpy global RNG (scipy truncnorm.rvs, np.random.randint, shuffle_rows) -> np.random.seed; the particle noise
(torch.randn_like, MPC.py:432) and the CEM candidates (scipy truncnorm.rvs, optimizers.py:100) are injected and
recorded.  The safety critic the planner queries is the Q_risk net of the `agent_algos_b64.npz` cases after
their warm-up updates (reproduced from that fixture, not stored again).
"""
import collections
import os
import sys

import numpy as np
import torch

from . import harness
from .make_golden import _save, _DummyEnv, _FixedMemory, ALGO_CASES

CASES = [
    # tag    algos tag (Q_risk state)  scale  plan_hor popsize elites npart
    ("nav",  "lr",   1.0, 5, 400, 40, 20),          # config/navigation1.py:110,122-125, config/default.py:108-109
    ("maze", "rcpo", 0.1, 15, 50, 5, 20),           # config/maze.py:110 horizon; popsize x npart = 1,000 (BASELINE C5)
]
STR = 29


class _Env(object):
    def __init__(self, scale):
        from gym.spaces import Box
        self.observation_space = Box(-np.ones(2) * float("inf"), np.ones(2) * float("inf"))
        self.action_space = Box(-np.ones(2) * scale, np.ones(2) * scale)


def _qrisk_agent(tag, z):
    """the reference SAC of algos case `tag` after its Q_risk warm-up (same recipe as make_golden.golden_algos)."""
    from gym.spaces import Box
    from recovery_rl.sac import SAC
    case = [c for c in ALGO_CASES if c[0] == tag][0]
    _, env_name, scale, extra = case
    seed, B = int(z["seed"]), int(z["B"])
    args = harness.get_args(extra + ["--env-name", env_name, "--seed", str(seed), "--batch_size", str(B)])
    torch.manual_seed(seed)
    np.random.seed(seed)
    agent = SAC(Box(-np.ones(2) * float("inf"), np.ones(2) * float("inf")), Box(-np.ones(2) * scale, np.ones(2) * scale),
                args, "/tmp/none", tmp_env=_DummyEnv())
    mem = _FixedMemory()
    for u in range(int(z["n_qr"])):
        q = "%s_qr%d_" % (tag, u)
        mem.batch = tuple(z[q + k] for k in ("s", "a", "c", "s2", "m"))
        harness.eps_queue.append(z[q + "eps_next"])
        agent.safety_critic.update_parameters(memory=mem, policy=agent.policy, batch_size=B)
        assert not harness.eps_queue
    return agent


def _make_mpc(scale, plan_hor, popsize, elites, npart):
    from dotmap import DotMap
    import config.maze as cm
    from recovery_rl.MPC import MPC

    def ctor(cfg):
        model = cm.PtModel(cfg.num_nets, 4, 4)
        model.optim = torch.optim.Adam(model.parameters(), lr=0.001)
        return model

    p = DotMap()
    p.env = _Env(scale)
    p.prop_cfg.model_init_cfg.num_nets = 5
    p.prop_cfg.model_init_cfg.model_constructor = ctor
    p.prop_cfg.model_train_cfg = {"epochs": 5}
    p.prop_cfg.mode = "TSinf"
    p.prop_cfg.npart = npart
    p.prop_cfg.obs_postproc = cm.MazeConfigModule.obs_postproc
    p.prop_cfg.targ_proc = cm.MazeConfigModule.targ_proc
    p.opt_cfg.mode = "CEM"
    p.opt_cfg.plan_hor = plan_hor
    p.opt_cfg.obs_cost_fn = lambda obs: None
    p.opt_cfg.ac_cost_fn = cm.MazeConfigModule.ac_cost_fn
    p.opt_cfg.cfg = {"popsize": popsize, "num_elites": elites, "max_iters": 5, "alpha": 0.1}
    so = sys.stdout
    sys.stdout = open(os.devnull, "w")
    try:
        return MPC(p)
    finally:
        sys.stdout = so


def golden_mpc():
    harness.setup()
    import recovery_rl.MPC as mpc_mod
    import recovery_rl.optimizers as opt_mod
    za = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden",
                              "agent_algos_b64.npz"))
    out = {"tags": np.array([c[0] for c in CASES]), "stride": np.int64(STR)}
    mpc_mod.trange = lambda n, **kw: range(n)          # no tqdm bar (and no set_postfix)

    class _Range(object):
        def __init__(self, n):
            self.n = n

        def __iter__(self):
            return iter(range(self.n))

        def set_postfix(self, *a, **k):
            pass

    mpc_mod.trange = lambda n, **kw: _Range(n)
    for tag, atag, scale, hor, pop, elites, npart in CASES:
        P = tag + "_"
        agent = _qrisk_agent(atag, za)
        np.random.seed(100)
        mpc = _make_mpc(scale, hor, pop, elites, npart)           # PtModel init: scipy truncnorm on the numpy global RNG
        mpc.update_value_func(agent.safety_critic)
        out.update({P + "algos_tag": np.array(atag), P + "scale": np.float64(scale), P + "plan_hor": np.int64(hor),
                    P + "popsize": np.int64(pop), P + "num_elites": np.int64(elites), P + "npart": np.int64(npart),
                    P + "model_seed": np.int64(100)})
        names = [n for n, _ in mpc.model.named_parameters()]
        out[P + "param_names"] = np.array(names)
        for n, p_ in mpc.model.named_parameters():
            out[P + "init_" + n] = p_.detach().numpy().ravel()[::STR].copy()
        # ---- MPC.train (MPC.py:213-309) on synthetic transitions of the env's scale ----
        rs = np.random.RandomState(7)
        n_tr = 300
        if tag == "maze":
            obs = rs.uniform(-0.27, 0.27, (n_tr, 2))
            acs = rs.uniform(-0.1, 0.1, (n_tr, 2)).astype(np.float32)
            nxt = obs + 0.2467 * acs + 1e-4 * rs.randn(n_tr, 2)
        else:
            obs = np.stack([rs.uniform(-75, 10, n_tr), rs.uniform(-9, 9, n_tr)], 1)
            acs = rs.uniform(-1, 1, (n_tr, 2)).astype(np.float32)
            nxt = obs + acs + 0.05 * rs.randn(n_tr, 2)
        np.random.seed(200)
        mpc.train(obs, acs, random=True, next_obs=nxt, epochs=3)
        out.update({P + "train_obs": obs, P + "train_acs": acs, P + "train_next": nxt, P + "train_seed": np.int64(200),
                    P + "train_epochs": np.int64(3)})
        for n, p_ in mpc.model.named_parameters():
            out[P + "trained_" + n] = p_.detach().numpy().ravel()[::STR].copy()
        with torch.no_grad():
            xin = torch.from_numpy(np.concatenate([obs, acs], 1)[None].repeat(5, 0)).float()
            mean, var = mpc.model(xin)
        out[P + "fwd_mean"] = mean.numpy().copy()
        out[P + "fwd_var"] = var.numpy().copy()
        # ---- _compile_cost (MPC.py:374-416) with injected particle noise ----
        rs = np.random.RandomState(8)
        cur = obs[3].copy()
        ac_seqs = rs.uniform(-scale, scale, (pop, hor * 2)).astype(np.float32)
        eps = rs.randn(hor, 5, pop * npart // 5, 2).astype(np.float32)
        q = collections.deque(eps)
        real = torch.randn_like
        torch.randn_like = lambda t, **kw: torch.from_numpy(q.popleft())
        try:
            mpc.sy_cur_obs = cur
            costs = mpc._compile_cost(ac_seqs)
        finally:
            torch.randn_like = real
        assert not q
        # ac_seqs / eps are regenerated in the tests from RandomState(8) in this order (not stored: size)
        out.update({P + "cost_obs": cur, P + "cost_rng_seed": np.int64(8), P + "cost_out": costs,
                    P + "cost_ac_seqs_sum": np.float64(ac_seqs.astype(np.float64).sum()),
                    P + "cost_eps_sum": np.float64(eps.astype(np.float64).sum())})
        # ---- CEMOptimizer.obtain_solution through MPC.act (optimizers.py:73-124, MPC.py:322-347) ----
        n_act = 2
        zs = rs.standard_normal((n_act, 5, pop, hor * 2))
        zs = np.clip(zs, -2, 2)                       # any values in [-2, 2] stand in for the truncnorm draws
        eps2 = rs.randn(n_act, 5, hor, 5, pop * npart // 5, 2).astype(np.float32)
        zq = collections.deque(zs.reshape(-1, pop, hor * 2))
        eq = collections.deque(eps2.reshape(-1, 5, pop * npart // 5, 2))

        class _X(object):
            def rvs(self, size):
                z = zq.popleft()
                assert list(z.shape) == list(size)
                return z

        real_tn = opt_mod.stats.truncnorm
        opt_mod.stats.truncnorm = lambda *a, **k: _X()
        torch.randn_like = lambda t, **kw: torch.from_numpy(eq.popleft())
        iter_costs = []
        orig_cost = mpc.optimizer.cost_function

        def rec_cost(samples):
            c = orig_cost(samples)
            iter_costs.append(c.copy())
            return c

        mpc.optimizer.cost_function = rec_cost
        acts, sols = [], []
        try:
            states = [obs[5].copy(), obs[6].copy()]
            for i in range(n_act):
                a = mpc.act(states[i], 0)
                acts.append(np.asarray(a, np.float64).copy())
                sols.append(mpc.prev_sol.copy())
        finally:
            torch.randn_like = real
            opt_mod.stats.truncnorm = real_tn
        out.update({P + "act_states": np.array(states), P + "act_z_sum": np.float64(zs.sum()),
                    P + "act_eps_sum": np.float64(eps2.astype(np.float64).sum()), P + "act_actions": np.array(acts),
                    P + "act_prev_sol": np.array(sols), P + "act_iter_costs": np.array(iter_costs),
                    P + "act_z_used": np.int64(zs.reshape(-1, pop, hor * 2).shape[0] - len(zq))})
    _save("mpc.npz", **out)


if __name__ == "__main__":
    golden_mpc()
