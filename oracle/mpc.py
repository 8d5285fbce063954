"""oracle/mpc.py -- TEST INFRASTRUCTURE: CPU restatement of the model-based recovery policy (BASELINE config 5).

Restates, with the reference's own numeric libraries (torch CPU fp32, numpy, scipy truncnorm):
  config/maze.py:23-96        PtModel: bootstrapped ensemble 4 -> 200 -> 200 -> 200 -> 4, swish, input
                              normalisation, soft-clamped log-variance, weight decays  (same class in
                              config/navigation1.py, navigation2.py)
  config/utils.py:6-27        swish, truncated-normal affine init (scipy truncnorm on the numpy global RNG)
  recovery_rl/MPC.py:213-309  MPC.train (bootstrap indices, batch 32, NLL + decays, Adam lr 1e-3, shuffle_rows)
  recovery_rl/MPC.py:322-347  MPC.act (CEM solution, receding-horizon shift of prev_sol)
  recovery_rl/MPC.py:374-467  _compile_cost / _predict_next_obs / TS-infinity particle bookkeeping
  recovery_rl/optimizers.py:73-124  CEMOptimizer.obtain_solution

PINNED against tests/golden/mpc.npz (outputs of the reference's own classes, oracle/ref_harness/make_golden_mpc.py).
The particle noise (torch.randn_like, MPC.py:432) and the CEM candidates (truncnorm.rvs, optimizers.py:100) are
explicit inputs; PtModel init and MPC.train draw from the numpy global RNG exactly like the reference.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may import this.
"""
import numpy as np
import torch
import torch.nn.functional as F
from scipy.stats import truncnorm

HID = 200


def swish(x):
    return x * torch.sigmoid(x)


def _affine(ens, n_in, n_out):
    w = truncnorm.rvs(-2, 2, size=(ens, n_in, n_out)) * (1.0 / (2.0 * np.sqrt(n_in)))
    return torch.tensor(w, dtype=torch.float32).requires_grad_(), torch.zeros(ens, 1, n_out).requires_grad_()


class PtModel(object):
    NAMES = ["lin0_w", "lin0_b", "lin1_w", "lin1_b", "lin2_w", "lin2_b", "lin3_w", "lin3_b", "inputs_mu", "inputs_sigma",
             "max_logvar", "min_logvar"]

    def __init__(self, ens=5, n_in=4, n_out=4):
        self.num_nets = ens
        self.lin0_w, self.lin0_b = _affine(ens, n_in, HID)
        self.lin1_w, self.lin1_b = _affine(ens, HID, HID)
        self.lin2_w, self.lin2_b = _affine(ens, HID, HID)
        self.lin3_w, self.lin3_b = _affine(ens, HID, n_out)
        self.inputs_mu = torch.zeros(n_in)
        self.inputs_sigma = torch.zeros(n_in)
        self.max_logvar = (torch.ones(1, n_out // 2) / 2.0).requires_grad_()
        self.min_logvar = (-torch.ones(1, n_out // 2) * 10.0).requires_grad_()
        self.n_out = n_out
        self.optim = torch.optim.Adam(self.trainable(), lr=0.001)

    def trainable(self):
        # nn.Module.parameters() order of the reference: registration order; the two input-stat Parameters have
        # requires_grad=False (Adam skips them: grad None)
        return [self.lin0_w, self.lin0_b, self.lin1_w, self.lin1_b, self.lin2_w, self.lin2_b, self.lin3_w, self.lin3_b,
                self.max_logvar, self.min_logvar]

    def named(self):
        return [(n, getattr(self, n)) for n in self.NAMES]

    def compute_decays(self):
        return (0.00025 * (self.lin0_w ** 2).sum() / 2.0 + 0.0005 * (self.lin1_w ** 2).sum() / 2.0 +
                0.0005 * (self.lin2_w ** 2).sum() / 2.0 + 0.00075 * (self.lin3_w ** 2).sum() / 2.0)

    def fit_input_stats(self, data):
        mu = np.mean(data, axis=0, keepdims=True)
        sigma = np.std(data, axis=0, keepdims=True)
        sigma[sigma < 1e-12] = 1.0
        self.inputs_mu = torch.from_numpy(mu).float()
        self.inputs_sigma = torch.from_numpy(sigma).float()

    def forward(self, inputs, ret_logvar=False):
        x = (inputs - self.inputs_mu) / self.inputs_sigma
        x = swish(x.matmul(self.lin0_w) + self.lin0_b)
        x = swish(x.matmul(self.lin1_w) + self.lin1_b)
        x = swish(x.matmul(self.lin2_w) + self.lin2_b)
        x = x.matmul(self.lin3_w) + self.lin3_b
        mean = x[:, :, :self.n_out // 2]
        logvar = x[:, :, self.n_out // 2:]
        logvar = self.max_logvar - F.softplus(self.max_logvar - logvar)
        logvar = self.min_logvar + F.softplus(logvar - self.min_logvar)
        return (mean, logvar) if ret_logvar else (mean, torch.exp(logvar))


def shuffle_rows(arr):
    idxs = np.argsort(np.random.uniform(size=arr.shape), axis=-1)
    return arr[np.arange(arr.shape[0])[:, None], idxs]


class MPC(object):
    """value_func(obs[n,2] fp32 tensor, acs[n,2]) -> max(Q1, Q2)_risk [n] (QRiskWrapper.get_value)."""

    def __init__(self, ac_lb, ac_ub, plan_hor, popsize, num_elites, npart=20, max_iters=5, alpha=0.1, epsilon=0.001,
                 value_func=None):
        self.dO, self.dU = 2, 2
        self.ac_lb, self.ac_ub = np.asarray(ac_lb, np.float32), np.asarray(ac_ub, np.float32)
        self.plan_hor, self.popsize, self.num_elites, self.npart = plan_hor, popsize, num_elites, npart
        self.max_iters, self.alpha, self.epsilon = max_iters, alpha, epsilon
        self.model = PtModel()
        self.value_func = value_func
        self.has_been_trained = False
        self.prev_sol = np.tile((self.ac_lb + self.ac_ub) / 2, [plan_hor])
        self.init_var = np.tile(np.square(self.ac_ub - self.ac_lb) / 16, [plan_hor])
        self.lb = np.tile(self.ac_lb, [plan_hor])
        self.ub = np.tile(self.ac_ub, [plan_hor])
        self.train_in = np.zeros((0, 4))
        self.train_targs = np.zeros((0, 2))
        self.iter_costs = []

    # ---- MPC.py:213-309 ------------------------------------------------------------------------------
    def train(self, obs, acs, next_obs, epochs):
        self.train_in = np.concatenate([self.train_in, np.concatenate([obs, acs], axis=-1)], axis=0)
        self.train_targs = np.concatenate([self.train_targs, next_obs - obs], axis=0)      # targ_proc
        self.has_been_trained = True
        m = self.model
        m.fit_input_stats(self.train_in)
        idxs = np.random.randint(self.train_in.shape[0], size=[m.num_nets, self.train_in.shape[0]])
        batch_size = 32
        num_batch = int(np.ceil(idxs.shape[-1] / batch_size))
        self.losses = []
        for _ in range(epochs):
            for b in range(num_batch):
                bi = idxs[:, b * batch_size:(b + 1) * batch_size]
                loss = 0.01 * (m.max_logvar.sum() - m.min_logvar.sum())
                loss = loss + m.compute_decays()
                tin = torch.from_numpy(self.train_in[bi]).float()
                ttg = torch.from_numpy(self.train_targs[bi]).float()
                mean, logvar = m.forward(tin, ret_logvar=True)
                inv_var = torch.exp(-logvar)
                tl = ((mean - ttg) ** 2) * inv_var + logvar
                loss = loss + tl.mean(-1).mean(-1).sum()
                m.optim.zero_grad()
                loss.backward()
                m.optim.step()
                self.losses.append(loss.item())
            idxs = shuffle_rows(idxs)

    # ---- MPC.py:374-467 ------------------------------------------------------------------------------
    def ts_expand(self, mat):
        d = mat.shape[-1]
        return mat.view(-1, self.model.num_nets, self.npart // self.model.num_nets, d).transpose(0, 1).contiguous() \
            .view(self.model.num_nets, -1, d)

    def ts_flatten(self, arr):
        d = arr.shape[-1]
        return arr.view(self.model.num_nets, -1, self.npart // self.model.num_nets, d).transpose(0, 1).contiguous().view(-1, d)

    @torch.no_grad()
    def compile_cost(self, cur_obs, ac_seqs, eps):
        """ac_seqs fp32 [nopt, hor*dU]; eps [hor][5][nopt*npart/5][2] -> costs fp32 [nopt]."""
        nopt = ac_seqs.shape[0]
        acs = torch.from_numpy(np.asarray(ac_seqs)).float().view(-1, self.plan_hor, self.dU).transpose(0, 1)[:, :, None]
        acs = acs.expand(-1, -1, self.npart, -1).contiguous().view(self.plan_hor, -1, self.dU)
        obs = torch.from_numpy(np.asarray(cur_obs)).float()[None].expand(nopt * self.npart, -1)
        costs = torch.zeros(nopt, self.npart)
        for t in range(self.plan_hor):
            a = acs[t]
            mean, var = self.model.forward(torch.cat((self.ts_expand(obs), self.ts_expand(a)), dim=-1))
            pred = self.ts_flatten(mean + torch.as_tensor(eps[t], dtype=torch.float32) * var.sqrt())
            costs += self.value_func(obs, a).reshape(-1, self.npart)
            obs = obs + pred                                                   # obs_postproc
        costs[costs != costs] = 1e6
        return costs.mean(dim=1).numpy()

    # ---- optimizers.py:73-124 + MPC.py:322-347 ----------------------------------------------------------
    def act(self, obs, zs, eps):
        """zs: iterable of [popsize, sol_dim] truncated-normal draws (one per CEM iteration that runs); eps: iterable
        of particle-noise tensors (one per iteration).  Returns the float64 action, like the reference."""
        mean, var, t = self.prev_sol, self.init_var, 0
        zs, eps = iter(zs), iter(eps)
        while t < self.max_iters and np.max(var) > self.epsilon:
            lb_dist, ub_dist = mean - self.lb, self.ub - mean
            cvar = np.minimum(np.minimum(np.square(lb_dist / 2), np.square(ub_dist / 2)), var)
            samples = (next(zs) * np.sqrt(cvar) + mean).astype(np.float32)
            costs = self.compile_cost(obs, samples, next(eps))
            self.iter_costs.append(costs.copy())
            elites = samples[np.argsort(costs)][:self.num_elites]
            mean = self.alpha * mean + (1 - self.alpha) * np.mean(elites, axis=0)
            var = self.alpha * var + (1 - self.alpha) * np.var(elites, axis=0)
            t += 1
        soln = mean
        self.prev_sol = np.concatenate([np.copy(soln)[self.dU:], np.zeros(self.dU)])
        return soln[:self.dU]
