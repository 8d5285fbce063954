"""SAC agent (reference recovery_rl/sac.py:25-277) on the device arena.

Same constructor and methods as the reference; the networks, autograd and Adam of the reference are replaced
by the CUDA kernels behind include/rrl.h.  Supported: the default SAC path, the Recovery RL path
(`--use_recovery --MF_recovery | --Q_sampling_recovery`) and the comparison branches of the shipped scripts:
LR / RSPO (`--DGD_constraints [--update_nu | --nu_schedule]`, sac.py:221-228,256-262), SQRL
(`--use_constraint_sampling`, sac.py:139-161), RCPO (`--RCPO`, sac.py:202-205,265-271), automatic entropy
tuning (sac.py:241-253) and `--policy Deterministic` (model.py:447-485).  `--cnn` (image observations) raises.
Update ordering of sac.py:233-239: "Variant B" (all forward expressions, both gradients, then both steps).
"""
import numpy as np
import torch

from . import native
from .arena import AgentArena
from .model import build_reference_modules
from .qrisk import QRiskWrapper


class NetHandle(object):
    """names one network of the arena (stands in for the reference's nn.Module attributes)."""

    def __init__(self, arena, name):
        self.arena, self.name = arena, name

    def parameters(self):
        return [self.arena.tensor(self.name, i) for i in range(self.arena.num_tensors(self.name))]

    def state_dict(self):
        return {"param_%d" % i: p.detach().cpu().clone() for i, p in enumerate(self.parameters())}

    def hard_update_from(self, source):
        self.arena.hard_update(self.name, source.name)

    def soft_update_from(self, source, tau):
        self.arena.soft_update(self.name, source.name, tau)


class SAC(object):
    def __init__(self, observation_space, action_space, args, logdir, im_shape=None, tmp_env=None):
        for flag in ("cnn", "vismpc_recovery"):
            if getattr(args, flag, False):
                raise NotImplementedError("--%s (image observations) is outside this build's hot path (DESIGN.md)" % flag)
        self.gamma, self.tau, self.alpha = args.gamma, args.tau, args.alpha
        self.env_name = args.env_name
        self.logdir = logdir
        self.policy_type = args.policy
        self.target_update_interval = args.target_update_interval
        deterministic = args.policy != "Gaussian"                                   # sac.py:115-117
        self.automatic_entropy_tuning = bool(args.automatic_entropy_tuning) and not deterministic
        if deterministic:
            self.alpha = 0
        self.DGD_constraints, self.update_nu, self.RCPO = args.DGD_constraints, args.update_nu, args.RCPO
        self.use_constraint_sampling = args.use_constraint_sampling
        self.lambda_RCPO = args.lambda_RCPO
        self.deterministic = deterministic
        self.device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        native.require_cuda()
        self.updates = 0
        self.gamma_safe, self.eps_safe = args.gamma_safe, args.eps_safe
        self.nu = args.nu
        scale = ((action_space.high - action_space.low) / 2.).astype(np.float32)
        bias = ((action_space.high + action_space.low) / 2.).astype(np.float32)
        self.arena = AgentArena(self.device, max_batch=args.batch_size, gamma=args.gamma, alpha=args.alpha, tau=args.tau,
                                lr=args.lr, gamma_safe=args.gamma_safe, tau_safe=args.tau_safe, eps_safe=args.eps_safe,
                                target_update_interval=args.target_update_interval, mf_recovery=args.MF_recovery,
                                action_scale=(float(scale[0]), float(scale[1])), action_bias=(float(bias[0]), float(bias[1])),
                                use_tensor_cores=getattr(args, "tensor_cores", 0), dgd=args.DGD_constraints,
                                update_nu=args.update_nu, rcpo=args.RCPO, auto_alpha=self.automatic_entropy_tuning,
                                deterministic=deterministic, nu=args.nu, lambda_rcpo=args.lambda_RCPO,
                                target_entropy=-float(np.prod(action_space.shape)))
        # xavier init in the reference's construction order (sac.py:82-114 then qrisk.py:36-75)
        self.arena.load_modules(build_reference_modules(hidden=args.hidden_size, obs_dim=observation_space.shape[0],
                                                        act_dim=action_space.shape[0], deterministic=deterministic))
        self.critic = NetHandle(self.arena, "critic")
        self.critic_target = NetHandle(self.arena, "critic_target")
        self.policy = NetHandle(self.arena, "policy")
        self.safety_critic = QRiskWrapper(observation_space, action_space, args.hidden_size, logdir, args,
                                          tmp_env=tmp_env, arena=self.arena)
        n = 1
        self._state = torch.zeros(2, n, dtype=torch.float64, device=self.device)
        self._a_task = torch.zeros(n, 2, device=self.device)
        self._a_real = torch.zeros(n, 2, device=self.device)
        self._losses = torch.zeros(16, device=self.device)

    def save(self, path):
        """state_dict-compatible checkpoint of the six networks + Adam state + multipliers (recovery_rl/checkpoint.py)."""
        from . import checkpoint
        return checkpoint.save(path, self.arena)

    def load(self, path):
        from . import checkpoint
        return checkpoint.load(path, self.arena)

    def _policy_noise(self, rows):
        """the draw of policy.sample for `rows` rows, from the torch global (CPU) generator like the reference:
        Normal.rsample -> randn [rows, 2] (model.py:329); DeterministicPolicy: ONE N(0, 0.1) vector clamped to
        +-0.25 and broadcast (model.py:478-480)."""
        if self.deterministic:
            n = torch.Tensor(2).normal_(0., std=0.1).clamp(-0.25, 0.25)
            return n.unsqueeze(0).repeat(rows, 1).contiguous().to(self.device)
        return torch.randn(rows, 2).to(self.device)

    def _sample_policy(self, states, eps):
        """policy.sample on a float32 [n, 2] device batch -> (action, log_pi, mean_action)."""
        n = states.shape[0]
        act = torch.zeros(n, 2, device=self.device)
        lp = torch.zeros(n, device=self.device)
        mean = torch.zeros(n, 2, device=self.device)
        native.policy_sample(self.arena.cfg, self.arena.arena, native.NET_POLICY, n, states.contiguous(), eps, act, lp, mean)
        return act, lp, mean

    def select_action(self, state, eval=False):
        """sac.py:133-168: policy.sample on one state; the eps draw is consumed in eval mode too."""
        if self.use_constraint_sampling:
            # SQRL (sac.py:139-161): 100 policy samples, keep those with Q_risk <= eps_safe, draw one with
            # probability ~ exp(log_pi) (Categorical from the torch generator); none left -> argmin Q_risk
            safe_samples = 100
            sb = torch.as_tensor(np.asarray(state, np.float32).reshape(1, 2)).repeat(safe_samples, 1).to(self.device)
            pi, log_pi, _ = self._sample_policy(sb, self._policy_noise(safe_samples))
            qmax = self.safety_critic.get_value(sb, pi)[:, 0]
            idxs = (qmax <= self.eps_safe).nonzero()[:, 0]
            probs = torch.exp(log_pi[idxs]).flatten().cpu()
            if probs.numel() == 0:
                return pi[torch.argmin(qmax)].cpu().numpy()
            sampled_idx = torch.distributions.Categorical(probs).sample()
            return pi[sampled_idx].cpu().numpy()        # indexes `pi` with the index into the filtered set, as sac.py:157-159
        if self.deterministic:
            sb = torch.as_tensor(np.asarray(state, np.float32).reshape(1, 2)).to(self.device)
            act, _, mean = self._sample_policy(sb, self._policy_noise(1))
            return (mean if eval else act).cpu().numpy()[0]
        eps = torch.randn(1, 2).to(self.device)
        self._state.copy_(torch.as_tensor(np.asarray(state, np.float64).reshape(1, 2).T))
        native.agent_act(self.arena.cfg, self.arena.arena, 1, self._state, None, self._a_task, self._a_real,
                         eps_task=eps, use_recovery=False, eval=bool(eval))
        return self._a_task.cpu().numpy()[0]

    def update_parameters(self, memory, batch_size, updates, nu=None, safety_critic=None):
        """sac.py:170-277 -> (qf1_loss, qf2_loss, policy_loss, alpha_loss, alpha)."""
        ar = self.arena
        if nu is None:
            nu = self.nu
        ar.set_nu_arg(nu)
        memory.sample_into(ar, "sac", batch_size)
        eps_next = self._policy_noise(batch_size)                   # policy.sample(next_state_batch)
        eps_cur = self._policy_noise(batch_size)                    # policy.sample(state_batch)
        ar.counters[native.C_SAC_UPDATES] = int(updates)
        native.sac_backward(ar.cfg, ar.arena, ar.counters, self._losses, eps_next, eps_cur)
        native.sac_apply(ar.cfg, ar.arena, ar.counters)
        l = self._losses[:5].cpu().numpy()
        alpha = float(l[4])
        if self.automatic_entropy_tuning or self.update_nu or self.RCPO:
            f32, f64 = ar.scalars()
            f32, f64 = f32.cpu().numpy(), f64.cpu().numpy()
            if self.automatic_entropy_tuning:
                self.alpha = alpha = float(f32[native.S_ALPHA])     # alpha_tlogs = self.alpha.clone() after the step
            if self.update_nu:
                self.nu = float(f64[native.D_NU_LEARNED])
            if self.RCPO:
                self.lambda_RCPO = float(f64[native.D_LAMBDA])
        return float(l[0]), float(l[1]), float(l[2]), float(l[3]), alpha
