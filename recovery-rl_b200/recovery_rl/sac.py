"""SAC agent (reference recovery_rl/sac.py:25-277) on the device arena.

Same constructor and methods as the reference; the networks, autograd and Adam of the reference are replaced
by the CUDA kernels behind include/rrl.h.  Supported: the default SAC path and the Recovery RL path
(`--use_recovery --MF_recovery | --Q_sampling_recovery`), fixed alpha, Gaussian policy.  The comparison
branches (LR/RSPO `--DGD_constraints`, SQRL `--use_constraint_sampling`, `--RCPO`, automatic entropy tuning,
`--policy Deterministic`, `--cnn`) raise NotImplementedError (DESIGN.md "next").
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


class SAC(object):
    def __init__(self, observation_space, action_space, args, logdir, im_shape=None, tmp_env=None):
        for flag in ("cnn", "DGD_constraints", "use_constraint_sampling", "RCPO", "update_nu", "vismpc_recovery"):
            if getattr(args, flag, False):
                raise NotImplementedError("--%s is outside this build's hot path (DESIGN.md, 'next')" % flag)
        if args.policy != "Gaussian" or args.automatic_entropy_tuning:
            raise NotImplementedError("only --policy Gaussian with fixed alpha is built (DESIGN.md, 'next')")
        self.gamma, self.tau, self.alpha = args.gamma, args.tau, args.alpha
        self.env_name = args.env_name
        self.logdir = logdir
        self.policy_type = args.policy
        self.target_update_interval = args.target_update_interval
        self.automatic_entropy_tuning = False
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
                                use_tensor_cores=getattr(args, "tensor_cores", 0))
        # xavier init in the reference's construction order (sac.py:82-114 then qrisk.py:36-75)
        self.arena.load_modules(build_reference_modules(hidden=args.hidden_size, obs_dim=observation_space.shape[0],
                                                        act_dim=action_space.shape[0]))
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

    def select_action(self, state, eval=False):
        """sac.py:133-168: policy.sample on one state; the eps draw is consumed in eval mode too."""
        eps = torch.randn(1, 2).to(self.device)
        self._state.copy_(torch.as_tensor(np.asarray(state, np.float64).reshape(1, 2).T))
        native.agent_act(self.arena.cfg, self.arena.arena, 1, self._state, None, self._a_task, self._a_real,
                         eps_task=eps, use_recovery=False, eval=bool(eval))
        return self._a_task.cpu().numpy()[0]

    def update_parameters(self, memory, batch_size, updates, nu=None, safety_critic=None):
        """sac.py:170-277 -> (qf1_loss, qf2_loss, policy_loss, alpha_loss, alpha)."""
        ar = self.arena
        memory.sample_into(ar, "sac", batch_size)
        eps_next = torch.randn(batch_size, 2).to(self.device)       # policy.sample(next_state_batch)
        eps_cur = torch.randn(batch_size, 2).to(self.device)        # policy.sample(state_batch)
        ar.counters[native.C_SAC_UPDATES] = int(updates)
        native.sac_backward(ar.cfg, ar.arena, ar.counters, self._losses, eps_next, eps_cur)
        native.sac_apply(ar.cfg, ar.arena, ar.counters)
        l = self._losses[:5].cpu().numpy()
        return float(l[0]), float(l[1]), float(l[2]), float(l[3]), float(l[4])
