"""oracle/agent.py -- TEST INFRASTRUCTURE: CPU (torch fp32) restatement of the SAC + safety-critic agent.

Restates, with the reference's own numeric library (torch CPU fp32, torch.optim.Adam):
  recovery_rl/model.py:49-76   QNetwork            :172-199 QNetworkConstraint (incl. the dead BatchNorm1d)
  recovery_rl/model.py:295-343 GaussianPolicy      :489-530 StochasticPolicy      :23-26 weights_init_
  recovery_rl/sac.py:133-168   SAC.select_action (incl. the SQRL filter :139-161)
  recovery_rl/sac.py:170-277   SAC.update_parameters: default / --use_recovery, and the comparison branches
                               LR/RSPO (--DGD_constraints, --update_nu :221-228,256-262), RCPO (:202-205,265-271),
                               automatic entropy tuning (:241-253), --policy Deterministic (model.py:447-485)
  recovery_rl/qrisk.py:86-182  QRiskWrapper.update_parameters   :184-213 get_value / select_action
  recovery_rl/utils.py:46-54   soft_update / hard_update
  recovery_rl/experiment.py:546-577 composite action selection

Update ordering of sac.py:233-239: "Variant B" (patch P4) -- the reference steps the critic
optimizer before policy_loss.backward(), which is only legal on its pinned torch 1.4; here all forward
expressions are evaluated as written, then critic grads, policy grads (w.r.t. policy parameters only),
critic step, policy step.  The golden vectors were generated from the reference under the same patch.

PINNED against tests/golden/agent_nav1_b256.npz, agent_maze_b64.npz, agent_algos_b64.npz and traj_nav1_seed7.npz (outputs of
the reference's own classes, oracle/ref_harness/make_golden.py).  Noise (eps) is always an explicit input.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may import this.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.optim import Adam

LOG_SIG_MAX = 2
LOG_SIG_MIN = -20
EPSILON = 1e-6
NET_NAMES = ["critic", "critic_target", "policy", "qrisk", "qrisk_target", "recovery"]


def weights_init_(m):
    if isinstance(m, nn.Linear):
        torch.nn.init.xavier_uniform_(m.weight, gain=1)
        torch.nn.init.constant_(m.bias, 0)


class QNetwork(nn.Module):
    def __init__(self, num_inputs, num_actions, hidden_dim, constraint=False):
        super().__init__()
        if constraint:
            self.bn1 = nn.BatchNorm1d(num_inputs + num_actions)   # never used in forward (model.py:175)
        self.linear1 = nn.Linear(num_inputs + num_actions, hidden_dim)
        self.linear2 = nn.Linear(hidden_dim, hidden_dim)
        self.linear3 = nn.Linear(hidden_dim, 1)
        self.linear4 = nn.Linear(num_inputs + num_actions, hidden_dim)
        self.linear5 = nn.Linear(hidden_dim, hidden_dim)
        self.linear6 = nn.Linear(hidden_dim, 1)
        self.constraint = constraint
        self.apply(weights_init_)

    def forward(self, state, action):
        xu = torch.cat([state, action], 1)
        x1 = self.linear3(F.relu(self.linear2(F.relu(self.linear1(xu)))))
        x2 = self.linear6(F.relu(self.linear5(F.relu(self.linear4(xu)))))
        if self.constraint:
            return torch.sigmoid(x1), torch.sigmoid(x2)
        return x1, x2


def _normal_log_prob(value, loc, scale):
    var = scale ** 2
    return -((value - loc) ** 2) / (2 * var) - scale.log() - math.log(math.sqrt(2 * math.pi))


class GaussianPolicy(nn.Module):
    def __init__(self, num_inputs, num_actions, hidden_dim, action_scale, action_bias):
        super().__init__()
        self.linear1 = nn.Linear(num_inputs, hidden_dim)
        self.linear2 = nn.Linear(hidden_dim, hidden_dim)
        self.mean_linear = nn.Linear(hidden_dim, num_actions)
        self.log_std_linear = nn.Linear(hidden_dim, num_actions)
        self.apply(weights_init_)
        self.action_scale = torch.as_tensor(action_scale, dtype=torch.float32)
        self.action_bias = torch.as_tensor(action_bias, dtype=torch.float32)

    def forward(self, state):
        x = F.relu(self.linear2(F.relu(self.linear1(state))))
        return self.mean_linear(x), torch.clamp(self.log_std_linear(x), min=LOG_SIG_MIN, max=LOG_SIG_MAX)

    def sample(self, state, eps):
        mean, log_std = self.forward(state)
        std = log_std.exp()
        x_t = mean + eps * std                                   # Normal.rsample
        y_t = torch.tanh(x_t)
        action = y_t * self.action_scale + self.action_bias
        log_prob = _normal_log_prob(x_t, mean, std)
        log_prob = log_prob - torch.log(self.action_scale * (1 - y_t.pow(2)) + EPSILON)
        log_prob = log_prob.sum(1, keepdim=True)
        mean = torch.tanh(mean) * self.action_scale + self.action_bias
        return action, log_prob, mean


class StochasticPolicy(nn.Module):
    def __init__(self, num_inputs, num_actions, hidden_dim, action_scale, action_bias):
        super().__init__()
        self.linear1 = nn.Linear(num_inputs, hidden_dim)
        self.linear2 = nn.Linear(hidden_dim, hidden_dim)
        self.mean = nn.Linear(hidden_dim, num_actions)
        self.log_std = nn.Parameter(torch.as_tensor([np.log(0.1)] * num_actions, dtype=torch.float32))  # P2
        self.min_log_std = np.log(1e-6)
        self.apply(weights_init_)
        self.action_scale = torch.as_tensor(action_scale, dtype=torch.float32)
        self.action_bias = torch.as_tensor(action_bias, dtype=torch.float32)

    def sample(self, state, eps):
        x = F.relu(self.linear2(F.relu(self.linear1(state))))
        mean = torch.tanh(self.mean(x)) * self.action_scale + self.action_bias
        log_std = torch.clamp(self.log_std, min=self.min_log_std).unsqueeze(0).repeat([len(mean), 1])
        std = torch.exp(log_std)
        action = mean + eps * std
        return action, _normal_log_prob(action, mean, std).sum(-1), mean


class DeterministicPolicy(nn.Module):
    """model.py:447-485.  `noise` is ONE [num_actions] vector per sample() call, N(0, 0.1) clamped to +-0.25 and
    broadcast over the batch; it is an explicit input here (already scaled and clamped)."""

    def __init__(self, num_inputs, num_actions, hidden_dim, action_scale, action_bias):
        super().__init__()
        self.linear1 = nn.Linear(num_inputs, hidden_dim)
        self.linear2 = nn.Linear(hidden_dim, hidden_dim)
        self.mean = nn.Linear(hidden_dim, num_actions)
        self.apply(weights_init_)
        self.action_scale = torch.as_tensor(action_scale, dtype=torch.float32)
        self.action_bias = torch.as_tensor(action_bias, dtype=torch.float32)

    def sample(self, state, noise):
        x = F.relu(self.linear2(F.relu(self.linear1(state))))
        mean = torch.tanh(self.mean(x)) * self.action_scale + self.action_bias
        return mean + torch.as_tensor(noise, dtype=torch.float32), torch.tensor(0.), mean


def deterministic_noise():
    """the draw of DeterministicPolicy.sample (model.py:478-479) from the torch global generator."""
    return torch.Tensor(2).normal_(0., std=0.1).clamp(-0.25, 0.25)


def soft_update(target, source, tau):
    for tp, p in zip(target.parameters(), source.parameters()):
        tp.data.copy_(tp.data * (1.0 - tau) + p.data * tau)


def hard_update(target, source):
    for tp, p in zip(target.parameters(), source.parameters()):
        tp.data.copy_(p.data)


class Agent(object):
    """SAC (sac.py:25-131) + QRiskWrapper (qrisk.py:27-84) with the reference's construction order
    (which fixes the xavier draws from the torch global RNG)."""

    def __init__(self, action_scale=(1.0, 1.0), action_bias=(0.0, 0.0), hidden=256, gamma=0.99, alpha=0.2, tau=0.005,
                 gamma_safe=0.5, tau_safe=0.0002, eps_safe=0.1, lr=3e-4, target_update_interval=1, mf_recovery=True,
                 dgd=False, update_nu=False, rcpo=False, auto_alpha=False, deterministic=False, nu=0.01,
                 lambda_rcpo=0.01):
        self.gamma, self.alpha, self.tau = gamma, alpha, tau
        self.dgd, self.update_nu, self.rcpo = bool(dgd), bool(update_nu), bool(rcpo)
        self.deterministic = bool(deterministic)
        self.auto_alpha = bool(auto_alpha) and not self.deterministic
        # sac.py:56-72: float64 scalars (np.log of a python float), Adam lr 0.1*lr
        self.nu, self.lambda_rcpo = nu, lambda_rcpo
        self.log_nu = torch.tensor(np.log(nu), requires_grad=True)
        self.nu_optim = Adam([self.log_nu], lr=0.1 * lr)
        self.log_lambda = torch.tensor(np.log(lambda_rcpo), requires_grad=True)
        self.lambda_optim = Adam([self.log_lambda], lr=0.1 * lr)
        if self.auto_alpha:                                   # sac.py:95-101
            self.target_entropy = -2.0
            self.log_alpha = torch.zeros(1, requires_grad=True)
            self.alpha_optim = Adam([self.log_alpha], lr=lr)
        if self.deterministic:                                # sac.py:116-117
            self.alpha = 0
        self.gamma_safe, self.tau_safe, self.eps_safe = gamma_safe, tau_safe, eps_safe
        self.target_update_interval = target_update_interval
        self.mf_recovery = mf_recovery
        self.critic = QNetwork(2, 2, hidden)
        self.critic_target = QNetwork(2, 2, hidden)
        self.critic_optim = Adam(self.critic.parameters(), lr=lr)
        hard_update(self.critic_target, self.critic)
        pol_cls = DeterministicPolicy if self.deterministic else GaussianPolicy
        self.policy = pol_cls(2, 2, hidden, action_scale, action_bias)
        self.policy_optim = Adam(self.policy.parameters(), lr=lr)
        self.qrisk = QNetwork(2, 2, hidden, constraint=True)
        self.qrisk_target = QNetwork(2, 2, hidden, constraint=True)
        self.qrisk_optim = Adam(self.qrisk.parameters(), lr=lr)
        hard_update(self.qrisk_target, self.qrisk)
        self.recovery = StochasticPolicy(2, 2, hidden, action_scale, action_bias)
        self.recovery_optim = Adam(self.recovery.parameters(), lr=lr)
        self.qrisk_updates = 0
        self.dbg = {}

    def nets(self):
        return dict(critic=self.critic, critic_target=self.critic_target, policy=self.policy, qrisk=self.qrisk,
                    qrisk_target=self.qrisk_target, recovery=self.recovery)

    def load(self, getter):
        """getter(net_name, i) -> np array for parameter i (None: keep)."""
        with torch.no_grad():
            for name, mod in self.nets().items():
                for i, p in enumerate(mod.parameters()):
                    v = getter(name, i)
                    if v is not None:
                        p.copy_(torch.as_tensor(np.asarray(v), dtype=torch.float32).reshape(p.shape))

    def params(self, name):
        return [p.detach().numpy().copy() for p in self.nets()[name].parameters()]

    # ---- sac.py:170-277 (Variant B) -------------------------------------------------------------
    def sac_update(self, batch, eps_next, eps_cur, updates, nu=None):
        """eps_next / eps_cur: [B, 2] standard normals (Gaussian policy) or the two [2] noise vectors of
        DeterministicPolicy.sample.  nu: the argument experiment.py:406 passes (nu_schedule(i_episode))."""
        if nu is None:
            nu = self.nu
        s, a, r, s2, m = [torch.as_tensor(np.asarray(x), dtype=torch.float32) for x in batch]
        r, m = r.reshape(-1, 1), m.reshape(-1, 1)
        eps_next = torch.as_tensor(eps_next, dtype=torch.float32)
        eps_cur = torch.as_tensor(eps_cur, dtype=torch.float32)
        with torch.no_grad():
            na, nlp, _ = self.policy.sample(s2, eps_next)
            q1n, q2n = self.critic_target(s2, na)
            y = r + m * self.gamma * (torch.min(q1n, q2n) - self.alpha * nlp)
            if self.rcpo:                                      # sac.py:202-205
                qsafe = torch.max(*self.qrisk(s, a))
                y -= self.lambda_rcpo * qsafe
        qf1, qf2 = self.critic(s, a)
        qf1_loss, qf2_loss = F.mse_loss(qf1, y), F.mse_loss(qf2, y)
        pi, log_pi, _ = self.policy.sample(s, eps_cur)
        qf1_pi, qf2_pi = self.critic(s, pi)
        min_qf_pi = torch.min(qf1_pi, qf2_pi)
        sqf1_pi, sqf2_pi = self.qrisk(s, pi)                   # sac.py:221-222
        max_sqf_pi = torch.max(sqf1_pi, sqf2_pi)
        if self.dgd:                                           # sac.py:224-228
            policy_loss = ((self.alpha * log_pi) + nu * (max_sqf_pi - self.eps_safe) - 1. * min_qf_pi).mean()
        else:
            policy_loss = ((self.alpha * log_pi) - min_qf_pi).mean()
        self.critic_optim.zero_grad()
        (qf1_loss + qf2_loss).backward(retain_graph=True)
        pol_params = list(self.policy.parameters())
        pol_grads = torch.autograd.grad(policy_loss, pol_params)
        self.critic_optim.step()
        self.policy_optim.zero_grad()
        for p, g in zip(pol_params, pol_grads):
            p.grad = g
        self.policy_optim.step()
        alpha_loss, alpha_log = 0.0, float(self.alpha)
        if self.auto_alpha:                                    # sac.py:241-250
            a_loss = -(self.log_alpha * (log_pi + self.target_entropy).detach()).mean()
            self.alpha_optim.zero_grad()
            a_loss.backward()
            self.alpha_optim.step()
            self.alpha = self.log_alpha.exp().detach()
            alpha_loss, alpha_log = a_loss.item(), self.alpha.item()
        if self.update_nu:                                     # sac.py:256-262
            nu_loss = (self.log_nu * (self.eps_safe - max_sqf_pi).detach()).mean()
            self.nu_optim.zero_grad()
            nu_loss.backward()
            self.nu_optim.step()
            self.nu = self.log_nu.exp().detach()
        if self.rcpo:                                          # sac.py:265-271
            l_loss = (self.log_lambda * (self.eps_safe - qsafe).detach()).mean()
            self.lambda_optim.zero_grad()
            l_loss.backward()
            self.lambda_optim.step()
            self.lambda_rcpo = self.log_lambda.exp().detach()
        if updates % self.target_update_interval == 0:
            soft_update(self.critic_target, self.critic, self.tau)
        self.dbg = dict(qf1=qf1.detach().numpy(), qf2=qf2.detach().numpy(), target=y.numpy(), pi=pi.detach().numpy(),
                        log_pi=log_pi.detach().numpy(), min_qf_pi=min_qf_pi.detach().numpy(), next_action=na.numpy(),
                        next_log_pi=nlp.numpy(), max_sqf_pi=max_sqf_pi.detach().numpy(),
                        critic_grads=[p.grad.detach().numpy().copy() for p in self.critic.parameters()],
                        policy_grads=[g.detach().numpy().copy() for g in pol_grads])
        return qf1_loss.item(), qf2_loss.item(), policy_loss.item(), alpha_loss, alpha_log

    # ---- sac.py:139-161: SQRL action filter ------------------------------------------------------
    def select_action_sqrl(self, state, eps, eps_safe=None, safe_samples=100, categorical=None):
        """eps: [safe_samples, 2].  The Categorical draw comes from the torch global generator, as in the reference,
        unless `categorical(probs) -> index` is given (replay of a recorded run)."""
        eps_safe = self.eps_safe if eps_safe is None else eps_safe
        with torch.no_grad():
            sb = torch.as_tensor(np.asarray(state), dtype=torch.float32).unsqueeze(0).repeat(safe_samples, 1)
            pi, log_pi, _ = self.policy.sample(sb, torch.as_tensor(eps, dtype=torch.float32))
            qmax = torch.max(*self.qrisk(sb, pi))
            idxs = (qmax <= eps_safe).nonzero()[:, 0]
            probs = torch.exp(log_pi[idxs]).flatten()
            if probs.numel() == 0:
                return pi[torch.argmin(qmax)].numpy()
            # NB sac.py:157-159 indexes `pi` with the index INTO THE FILTERED SET (not thresh_idxs[sampled_idx]):
            # reproduced as written
            j = torch.distributions.Categorical(probs).sample() if categorical is None else int(categorical(probs))
            return pi[j].numpy()

    # ---- qrisk.py:214-225: Q-sampling recovery -----------------------------------------------------
    def select_action_qsample(self, state, candidates):
        """candidates: [samples, 2] float32 draws of `ac_space.sample()` (1000 in the reference, gym Box: uniform in
        [low, high]); returns the one with the smallest max(Q1, Q2)_risk(state, .)  (qrisk.py:196, 222-224).
        PINNED against tests/golden/qsample_nav1.npz (the reference's own QRiskWrapper.select_action)."""
        cand = torch.as_tensor(np.asarray(candidates), dtype=torch.float32)
        with torch.no_grad():
            sb = torch.as_tensor(np.asarray(state), dtype=torch.float32).unsqueeze(0).repeat(len(cand), 1)
            qmax = torch.max(*self.qrisk(sb, cand))
            return cand[torch.argmin(qmax)].numpy()

    # ---- qrisk.py:86-182 ---------------------------------------------------------------------------
    def qrisk_update(self, batch, eps_next, eps_rec):
        s, a, c, s2, m = [torch.as_tensor(np.asarray(x), dtype=torch.float32) for x in batch]
        c, m = c.reshape(-1, 1), m.reshape(-1, 1)
        eps_next = torch.as_tensor(eps_next, dtype=torch.float32)
        with torch.no_grad():
            na, _, _ = self.policy.sample(s2, eps_next)          # TASK policy (experiment.py:413)
            q1n, q2n = self.qrisk_target(s2, na)
            y = c + m * self.gamma_safe * torch.max(q1n, q2n)
        q1, q2 = self.qrisk(s, a)
        l1, l2 = F.mse_loss(q1, y), F.mse_loss(q2, y)
        self.qrisk_optim.zero_grad()
        (l1 + l2).backward()
        self.qrisk_optim.step()
        rec_loss = 0.0
        if self.mf_recovery:
            pi, _, _ = self.recovery.sample(s, torch.as_tensor(eps_rec, dtype=torch.float32))
            qp1, qp2 = self.qrisk(s, pi)                         # POST-step critic (qrisk.py:150-152)
            policy_loss = torch.max(qp1, qp2).mean()
            self.recovery_optim.zero_grad()
            policy_loss.backward()
            self.recovery_optim.step()
            rec_loss = policy_loss.item()
        if self.qrisk_updates % self.target_update_interval == 0:
            soft_update(self.qrisk_target, self.qrisk, self.tau_safe)
        self.qrisk_updates += 1
        self.dbg = dict(q1=q1.detach().numpy(), q2=q2.detach().numpy(), target=y.numpy(), next_action=na.numpy())
        return l1.item(), l2.item(), rec_loss

    # ---- experiment.py:546-577 (vectorised over rows) ---------------------------------------------
    def act(self, state, eps_task, eps_rec, use_recovery=True, eval=False, eps_safe=None):
        eps_safe = self.eps_safe if eps_safe is None else eps_safe
        with torch.no_grad():
            st = torch.as_tensor(np.asarray(state), dtype=torch.float32)      # torch.FloatTensor(state)
            a_s, logp, a_mean = self.policy.sample(st, torch.as_tensor(eps_task, dtype=torch.float32))
            a_task = a_mean if eval else a_s
            if not use_recovery:
                return a_task.numpy(), a_task.numpy().copy(), np.zeros(len(st), bool), np.zeros(len(st), np.float32)
            q1, q2 = self.qrisk(st, a_task)
            qv = torch.max(q1, q2)[:, 0]
            rec = qv > eps_safe
            a_rec, _, _ = self.recovery.sample(st, torch.as_tensor(eps_rec, dtype=torch.float32))
            real = torch.where(rec[:, None], a_rec, a_task)
        return a_task.numpy(), real.numpy(), rec.numpy(), qv.numpy()
