"""CEM optimizer of the model-based recovery policy (reference recovery_rl/optimizers.py:28-124).

Same class and `obtain_solution(init_mean, init_var)` contract as the reference, but the loop
(sample -> cost -> elites -> smoothing) runs on the device for a batch of env copies at once: candidates are
drawn, rolled through the ensemble, ranked and reduced by the kernels of csrc/mpc.cu.  `cost_function` is therefore
not a numpy callback but the planner context (recovery_rl.MPC.MPC) that owns the device buffers."""
from . import native


class Optimizer(object):
    def setup(self, cost_function):
        raise NotImplementedError("Must be implemented in subclass.")

    def reset(self):
        raise NotImplementedError("Must be implemented in subclass.")

    def obtain_solution(self, *args, **kwargs):
        raise NotImplementedError("Must be implemented in subclass.")


class CEMOptimizer(Optimizer):
    def __init__(self, sol_dim, max_iters, popsize, num_elites, cost_function, upper_bound=None, lower_bound=None,
                 epsilon=0.001, alpha=0.25):
        self.sol_dim, self.max_iters, self.popsize, self.num_elites = sol_dim, max_iters, popsize, num_elites
        self.ub, self.lb = upper_bound, lower_bound
        self.epsilon, self.alpha = epsilon, alpha
        self.cost_function = cost_function          # the MPC planner context
        if num_elites > popsize:
            raise ValueError("Number of elites must be at most the population size.")

    def reset(self):
        pass

    def obtain_solution(self, init_mean, init_var, z=None, eps=None):
        """init_mean / init_var: fp64 device tensors [n_envs, sol_dim] (updated in place and returned).
        z: optional per-iteration truncated-normal draws [iters][n_envs, popsize, sol_dim] (parity mode);
        eps: optional per-iteration particle noise.  Without them the device Philox streams are used."""
        ctx = self.cost_function
        n = init_mean.shape[0]
        for it in range(self.max_iters):
            zi = None if z is None else z[it]
            ei = None if eps is None else eps[it]
            native.mpc_sample(ctx.cfg, n, it, init_mean, init_var, ctx.samples, ctx.active, z=zi, counters=ctx.counters)
            native.mpc_rollout(ctx.cfg, ctx.agent_cfg, ctx.agent_arena, ctx.dyn_image, n, ctx.state, ctx.samples, ctx.row_cost,
                               active=ctx.active, eps=ei, it=it, counters=ctx.counters)
            native.mpc_update(ctx.cfg, n, it, ctx.samples, ctx.row_cost, ctx.active, init_mean, init_var)
        return init_mean
