"""Safety critic Q_risk + model-free recovery policy (reference recovery_rl/qrisk.py:26-227) on the device
arena.  Same constructor and methods as the reference (`update_parameters`, `get_value`, `select_action`,
`__call__`); `plot` is not built (dead in the reference: always called with plot=0, experiment.py:415)."""
import numpy as np
import torch

from . import native


class QRiskWrapper(object):
    def __init__(self, obs_space, ac_space, hidden_size, logdir, args, tmp_env, arena=None):
        if arena is None:
            raise ValueError("QRiskWrapper shares the agent arena: construct it through SAC")
        self.env_name = args.env_name
        self.logdir = logdir
        self.device = arena.device
        self.ac_space = ac_space
        self.images = False
        self.encoding = False
        self.arena = arena
        self.lr = args.lr
        self.tau = args.tau_safe
        self.gamma_safe = args.gamma_safe
        self.updates = 0
        self.target_update_interval = args.target_update_interval
        self.pos_fraction = args.pos_fraction if args.pos_fraction >= 0 else None
        self.MF_recovery = args.MF_recovery
        self.Q_sampling_recovery = args.Q_sampling_recovery
        self.tmp_env = tmp_env
        self._losses = torch.zeros(16, device=self.device)
        self.torchify = lambda x: torch.FloatTensor(x).to(self.device)

    def update_parameters(self, memory=None, policy=None, batch_size=None, plot=False):
        """qrisk.py:86-182 (policy is the agent's task policy: it lives in the same arena)."""
        if self.pos_fraction:
            batch_size = min(batch_size, int((1 - self.pos_fraction) * len(memory)))
        else:
            batch_size = min(batch_size, len(memory))
        ar = self.arena
        memory.sample_into(ar, "qr", batch_size, pos_fraction=self.pos_fraction)
        # a' ~ policy.sample(next_state_batch): the TASK policy's own noise (Normal.rsample, or the broadcast
        # noise vector of DeterministicPolicy.sample)
        if ar.cfg.algo_flags & native.ALGO_DETERMINISTIC:
            eps_next = torch.Tensor(2).normal_(0., std=0.1).clamp(-0.25, 0.25).unsqueeze(0).repeat(batch_size, 1) \
                .contiguous().to(self.device)
        else:
            eps_next = torch.randn(batch_size, 2).to(self.device)
        ar.counters[native.C_QRISK_UPDATES] = int(self.updates)
        native.qrisk_backward(ar.cfg, ar.arena, ar.counters, self._losses, eps_next)
        native.qrisk_apply(ar.cfg, ar.arena, ar.counters)
        if self.MF_recovery:
            eps_rec = torch.randn(batch_size, 2).to(self.device)
            native.recovery_backward(ar.cfg, ar.arena, ar.counters, self._losses, eps_rec)
        native.recovery_apply(ar.cfg, ar.arena, ar.counters)      # Polyak (qrisk.py:160-162) + updates += 1
        self.updates += 1

    def losses(self):
        return tuple(float(x) for x in self._losses[:3].cpu().numpy())

    def _twin(self, states, actions):
        s = torch.as_tensor(states, dtype=torch.float32).reshape(-1, 2).to(self.device).contiguous()
        a = torch.as_tensor(actions, dtype=torch.float32).reshape(-1, 2).to(self.device).contiguous()
        n = s.shape[0]
        q1 = torch.zeros(n, device=self.device)
        q2 = torch.zeros(n, device=self.device)
        native.twin_q_forward(self.arena.cfg, self.arena.arena, native.NET_QRISK, n, s, a, q1, q2)
        return q1.unsqueeze(1), q2.unsqueeze(1)

    def get_value(self, states, actions, encoded=False):
        q1, q2 = self._twin(states, actions)
        return torch.max(q1, q2)

    def select_action(self, state, eval=False):
        if self.MF_recovery:
            eps = torch.randn(1, 2).to(self.device)
            s = torch.as_tensor(np.asarray(state), dtype=torch.float32).reshape(1, 2).to(self.device)
            act = torch.zeros(1, 2, device=self.device)
            mean = torch.zeros(1, 2, device=self.device)
            native.policy_sample(self.arena.cfg, self.arena.arena, native.NET_RECOVERY, 1, s, eps, act, None, mean)
            return (mean if eval else act).cpu().numpy()[0]
        elif self.Q_sampling_recovery:
            sampled = np.array([self.ac_space.sample() for _ in range(1000)], np.float32)
            states = np.repeat(np.asarray(state, np.float32).reshape(1, 2), 1000, 0)
            q = self.get_value(states, sampled)
            return sampled[int(torch.argmin(q).item())]
        else:
            assert False

    def __call__(self, states, actions):
        return self._twin(states, actions)
