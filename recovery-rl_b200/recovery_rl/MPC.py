"""Model-based recovery policy: PETS ensemble + CEM planner (reference recovery_rl/MPC.py:55-467,
config/maze.py:23-96) -- BASELINE config 5.

Same class and methods as the reference (`MPC(params)`, `train`, `act`, `update_value_func`, `reset`).  What runs
where:
  * `act` -- the hot part (5 CEM iterations x plan_hor steps x popsize x npart particles through the ensemble and
    the safety critic) -- runs entirely in the CUDA kernels of csrc/mpc.cu, for `n_envs` env copies at once; the
    ensemble is packed once per `train` into a padded k-major image the planner streams.
  * `train` (MPC.py:213-309: bootstrapped NLL fit, batch 32, Adam 1e-3) keeps the reference's data handling and
    numpy RNG calls; every mini-batch is ONE launch of `dyn_train_kernel` (csrc/mpc.cu: forward, Gaussian NLL +
    log-variance bounds + weight decays, backward and torch-Adam for all five nets).  The PtModel parameters are
    views of the flat train arena that kernel updates.  (`train(..., use_torch=True)` runs the same schedule through
    torch autograd on the GPU: kept as a cross-check, not the product path.)
There is no CPU path: construction raises without a CUDA device.
"""
import numpy as np
import torch
from torch import nn as nn
from torch.nn import functional as F

from . import native
from .optimizers import CEMOptimizer
from .utils import get_required_argument

HIDDEN = 200


def swish(x):
    return x * torch.sigmoid(x)


class PtModel(nn.Module):
    """config/maze.py:23-96.  Parameters in the reference's layout and registration order."""

    def __init__(self, ensemble_size, in_features, out_features):
        super().__init__()
        from config.utils import get_affine_params
        native.require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device())
        self.num_nets = ensemble_size
        self.in_features = in_features
        self.out_features = out_features
        # the reference's init draws, in its order (scipy truncnorm on the numpy global RNG) ...
        init = []
        for n_in, n_out in ((in_features, HIDDEN), (HIDDEN, HIDDEN), (HIDDEN, HIDDEN), (HIDDEN, out_features)):
            w, b = get_affine_params(ensemble_size, n_in, n_out)
            init += [w.data, b.data]
        init += [torch.ones(1, out_features // 2, dtype=torch.float32) / 2.0,
                 -torch.ones(1, out_features // 2, dtype=torch.float32) * 10.0]
        # ... stored in ONE flat device arena (the layout dyn_train_kernel updates); the parameters are views of it
        self.train_arena = torch.zeros(native.dyn_train_floats(), device=dev)
        names = ["lin0_w", "lin0_b", "lin1_w", "lin1_b", "lin2_w", "lin2_b", "lin3_w", "lin3_b", "max_logvar", "min_logvar"]
        off = 0
        for name, t in zip(names, init):
            view = self.train_arena[off:off + t.numel()].view(t.shape)
            view.copy_(t)
            self.register_parameter(name, nn.Parameter(view))
            off += t.numel()
        assert off == self.train_arena.numel()
        self.inputs_mu = nn.Parameter(torch.zeros(in_features, device=dev), requires_grad=False)
        self.inputs_sigma = nn.Parameter(torch.zeros(in_features, device=dev), requires_grad=False)
        self.adam_m = torch.zeros_like(self.train_arena)
        self.adam_v = torch.zeros_like(self.train_arena)
        self.wt = torch.zeros(2 * ensemble_size * HIDDEN * HIDDEN, device=dev)
        self.partial = torch.zeros(32, device=dev)
        self.adam_step = torch.zeros(1, dtype=torch.int64, device=dev)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        self.loss_dev = torch.zeros(1, device=dev)

    def compute_decays(self):
        return (0.00025 * (self.lin0_w ** 2).sum() / 2.0 + 0.0005 * (self.lin1_w ** 2).sum() / 2.0 +
                0.0005 * (self.lin2_w ** 2).sum() / 2.0 + 0.00075 * (self.lin3_w ** 2).sum() / 2.0)

    def fit_input_stats(self, data):
        mu = np.mean(data, axis=0, keepdims=True)
        sigma = np.std(data, axis=0, keepdims=True)
        sigma[sigma < 1e-12] = 1.0
        dev = self.lin0_w.device
        self.inputs_mu.data = torch.from_numpy(mu).to(dev).float()
        self.inputs_sigma.data = torch.from_numpy(sigma).to(dev).float()

    def forward(self, inputs, ret_logvar=False):
        inputs = (inputs - self.inputs_mu) / self.inputs_sigma
        inputs = swish(inputs.matmul(self.lin0_w) + self.lin0_b)
        inputs = swish(inputs.matmul(self.lin1_w) + self.lin1_b)
        inputs = swish(inputs.matmul(self.lin2_w) + self.lin2_b)
        inputs = inputs.matmul(self.lin3_w) + self.lin3_b
        mean = inputs[:, :, :self.out_features // 2]
        logvar = inputs[:, :, self.out_features // 2:]
        logvar = self.max_logvar - F.softplus(self.max_logvar - logvar)
        logvar = self.min_logvar + F.softplus(logvar - self.min_logvar)
        if ret_logvar:
            return mean, logvar
        return mean, torch.exp(logvar)


def shuffle_rows(arr):
    idxs = np.argsort(np.random.uniform(size=arr.shape), axis=-1)
    return arr[np.arange(arr.shape[0])[:, None], idxs]


class Controller(object):
    def __init__(self, *args, **kwargs):
        pass


class MPC(Controller):
    optimizers = {"CEM": CEMOptimizer}

    def __init__(self, params, n_envs=1, seed=0, stream_id=0):
        super().__init__(params)
        native.require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.dO, self.dU = params.env.observation_space.shape[0], params.env.action_space.shape[0]
        self.ac_ub, self.ac_lb = params.env.action_space.high, params.env.action_space.low
        self.ac_ub = np.minimum(self.ac_ub, params.get("ac_ub", self.ac_ub))
        self.ac_lb = np.maximum(self.ac_lb, params.get("ac_lb", self.ac_lb))
        self.update_fns = params.get("update_fns", [])
        self.per = params.get("per", 1)
        if self.per != 1 or self.dO != 2 or self.dU != 2:
            raise NotImplementedError("the planner kernels cover the 2-D point envs with per = 1")
        self.model_init_cig = params.prop_cfg.get("model_init_cfg", {})
        self.model_train_cfg = params.prop_cfg.get("model_train_cfg", {})
        self.prop_mode = get_required_argument(params.prop_cfg, "mode", "Must provide propagation method.")
        self.npart = get_required_argument(params.prop_cfg, "npart", "Must provide number of particles.")
        self.opt_mode = get_required_argument(params.opt_cfg, "mode", "Must provide optimization method.")
        self.plan_hor = get_required_argument(params.opt_cfg, "plan_hor", "Must provide planning horizon.")
        self.obs_cost_fn = get_required_argument(params.opt_cfg, "obs_cost_fn", "Must provide cost on observations.")
        self.ac_cost_fn = get_required_argument(params.opt_cfg, "ac_cost_fn", "Must provide cost on actions.")
        assert self.opt_mode == 'CEM'
        assert self.prop_mode == 'TSinf', 'only TSinf propagation mode is supported'
        assert self.npart % self.model_init_cig.num_nets == 0, "Number of particles must be a multiple of the ensemble size."
        opt_cfg = dict(params.opt_cfg.get("cfg", {}))
        self.optimizer = CEMOptimizer(sol_dim=self.plan_hor * self.dU, lower_bound=np.tile(self.ac_lb, [self.plan_hor]),
                                      upper_bound=np.tile(self.ac_ub, [self.plan_hor]), cost_function=self, **opt_cfg)
        self.has_been_trained = params.prop_cfg.get("model_pretrained", False)
        self.ac_buf = np.array([]).reshape(0, self.dU)
        self.train_in = np.array([]).reshape(0, self.dU + self.dO)
        self.train_targs = np.array([]).reshape(0, self.dO)
        print("Created an MPC controller, prop mode %s, %d particles. " % (self.prop_mode, self.npart))
        self.model = get_required_argument(params.prop_cfg.model_init_cfg, "model_constructor",
                                           "Must provide a model constructor.")(params.prop_cfg.model_init_cfg)
        self.value_func = None
        # ---- device-side planner state (csrc/mpc.cu) ----
        o = self.optimizer
        self.n_envs = int(n_envs)
        self.cfg = native.mpc_config(self.plan_hor, o.popsize, o.num_elites, npart=self.npart,
                                     num_nets=self.model_init_cig.num_nets, max_iters=o.max_iters, alpha=o.alpha,
                                     epsilon=o.epsilon, ac_lb=self.ac_lb, ac_ub=self.ac_ub, seed=seed, stream_id=stream_id)
        E, sol, dev = self.n_envs, self.plan_hor * self.dU, self.device
        mid = np.tile((self.ac_lb + self.ac_ub) / 2, [self.plan_hor]).astype(np.float64)
        self.prev_sol_dev = torch.from_numpy(np.tile(mid[None], (E, 1))).to(dev).contiguous()      # MPC.py:179
        self.mean = torch.zeros(E, sol, dtype=torch.float64, device=dev)
        self.var = torch.zeros(E, sol, dtype=torch.float64, device=dev)
        self.active = torch.zeros(E, dtype=torch.int32, device=dev)
        self.samples = torch.zeros(E, o.popsize, sol, device=dev)
        self.row_cost = torch.zeros(E, o.popsize, self.npart, device=dev)
        self.state = torch.zeros(2, E, dtype=torch.float64, device=dev)
        self.action_dev = torch.zeros(E, 2, dtype=torch.float64, device=dev)
        self.dyn_image = torch.zeros(native.dyn_image_floats(), device=dev)
        self.counters = None
        self.agent_cfg = self.agent_arena = None
        self.init_var = np.tile(np.square(self.ac_ub - self.ac_lb) / 16, [self.plan_hor])

    @property
    def prev_sol(self):
        return self.prev_sol_dev[0].cpu().numpy()

    # ---- MPC.py:213-309 -----------------------------------------------------------------------------------
    def train(self, obs_trajs, acs_trajs, random=False, next_obs=False, epochs=None, use_torch=False):
        new_train_in, new_train_targs = [], []
        if random:
            assert next_obs is not None
            new_train_in = [np.concatenate([np.asarray(obs_trajs), np.asarray(acs_trajs)], axis=-1)]
            new_train_targs = [np.asarray(next_obs) - np.asarray(obs_trajs)]                   # targ_proc
        else:
            for obs, acs in zip(obs_trajs, acs_trajs):
                new_train_in.append(np.concatenate([obs[:-1], acs], axis=-1))
                new_train_targs.append(obs[1:] - obs[:-1])
        self.train_in = np.concatenate([self.train_in] + new_train_in, axis=0)
        self.train_targs = np.concatenate([self.train_targs] + new_train_targs, axis=0)
        self.has_been_trained = True
        m = self.model
        m.fit_input_stats(self.train_in)
        idxs = np.random.randint(self.train_in.shape[0], size=[m.num_nets, self.train_in.shape[0]])
        if epochs is None:
            epochs = self.model_train_cfg['epochs']
        batch_size = 32
        num_batch = int(np.ceil(idxs.shape[-1] / batch_size))
        tin_all = torch.from_numpy(self.train_in).to(self.device).float().contiguous()
        ttg_all = torch.from_numpy(self.train_targs).to(self.device).float().contiguous()
        if not use_torch:
            native.dyn_train_sync(m.train_arena, m.wt)
            mu, sigma = m.inputs_mu.data.reshape(-1).contiguous(), m.inputs_sigma.data.reshape(-1).contiguous()
            n_idx = idxs.shape[-1]
            for _ in range(epochs):
                idx_dev = torch.from_numpy(np.ascontiguousarray(idxs)).to(self.device)
                for b in range(num_batch):
                    rows = min(batch_size, n_idx - b * batch_size)
                    native.dyn_train_step(m.train_arena, m.adam_m, m.adam_v, m.wt, m.partial, mu, sigma, tin_all, ttg_all,
                                          idx_dev, b * batch_size, rows, 0.001, m.adam_step, m.ticket, loss_out=m.loss_dev)
                idxs = shuffle_rows(idxs)
            self.last_train_loss = float(m.loss_dev.item()) + float(
                (0.01 * (m.max_logvar.sum() - m.min_logvar.sum()) + m.compute_decays()).item())
            self.pack_model()
            return
        for _ in range(epochs):
            idx_dev = torch.from_numpy(idxs).to(self.device)
            for b in range(num_batch):
                bi = idx_dev[:, b * batch_size:(b + 1) * batch_size]
                loss = 0.01 * (m.max_logvar.sum() - m.min_logvar.sum())
                loss = loss + m.compute_decays()
                mean, logvar = m(tin_all[bi], ret_logvar=True)
                inv_var = torch.exp(-logvar)
                train_losses = ((mean - ttg_all[bi]) ** 2) * inv_var + logvar
                loss = loss + train_losses.mean(-1).mean(-1).sum()
                m.optim.zero_grad()
                loss.backward()
                m.optim.step()
            idxs = shuffle_rows(idxs)
        self.last_train_loss = float(loss.item())
        self.pack_model()

    def pack_model(self):
        """ensemble -> padded k-major image streamed by the planner kernels."""
        m = self.model
        ts = [m.lin0_w, m.lin0_b, m.lin1_w, m.lin1_b, m.lin2_w, m.lin2_b, m.lin3_w, m.lin3_b, m.inputs_mu, m.inputs_sigma,
              m.max_logvar, m.min_logvar]
        native.dyn_pack([t.detach().reshape(-1) for t in ts], HIDDEN, self.dyn_image)

    def reset(self):
        mid = np.tile((self.ac_lb + self.ac_ub) / 2, [self.plan_hor]).astype(np.float64)
        self.prev_sol_dev.copy_(torch.from_numpy(np.tile(mid[None], (self.n_envs, 1))))
        self.optimizer.reset()
        for update_fn in self.update_fns:
            update_fn()

    def update_value_func(self, value_func):
        """value_func: the agent's QRiskWrapper; the planner reads the safety critic straight from the agent arena."""
        self.value_func = value_func
        self.agent_cfg = value_func.arena.cfg
        self.agent_arena = value_func.arena.arena

    # ---- MPC.py:322-347 -----------------------------------------------------------------------------------
    def plan(self, mask=None, z=None, eps=None):
        """one CEM solve for all n_envs env copies from self.state (fp64 [2][n_envs]); returns the device action
        tensor [n_envs][2] fp64.  mask (u8 [n_envs]): only those envs shift their warm start (the reference plans
        only when recovery triggers)."""
        if self.agent_arena is None:
            raise RuntimeError("update_value_func(agent.safety_critic) must be called before act (experiment.py:164-167)")
        E = self.n_envs
        native.mpc_begin(self.cfg, E, self.prev_sol_dev, self.mean, self.var, self.active)
        self.optimizer.obtain_solution(self.mean, self.var, z=z, eps=eps)
        native.mpc_finish(self.cfg, E, self.mean, self.prev_sol_dev, self.action_dev, mask=mask)
        return self.action_dev

    def act(self, obs, t, get_pred_cost=False, z=None, eps=None):
        if not self.has_been_trained:
            return np.random.uniform(self.ac_lb, self.ac_ub, self.ac_lb.shape)
        if self.ac_buf.shape[0] > 0:
            action, self.ac_buf = self.ac_buf[0], self.ac_buf[1:]
            return action
        self.sy_cur_obs = obs
        self.state.copy_(torch.as_tensor(np.asarray(obs, np.float64).reshape(1, 2).T.repeat(self.n_envs, 1)))
        soln = self.plan(z=z, eps=eps)
        self.ac_buf = soln[:1].cpu().numpy().reshape(-1, self.dU)
        return self.act(obs, t)
