"""Parameter containers of the four MLP families (host side).

The arithmetic of these networks runs in librrl.so (csrc/agent.cu); the torch modules below exist only to
(a) draw the initial weights from the torch global RNG in exactly the reference's order -- nn.Linear's own
init draws followed by xavier_uniform_ / zero bias (reference recovery_rl/model.py:23-26), module by module
in the construction order of sac.py:82-131 and qrisk.py:36-75 -- and (b) give checkpoints the reference's
state_dict names.  Parameter order (== arena tensor order): model.py:49-76, 172-199, 295-343, 489-530.
"""
import numpy as np
import torch
import torch.nn as nn


def weights_init_(m):
    if isinstance(m, nn.Linear):
        torch.nn.init.xavier_uniform_(m.weight, gain=1)
        torch.nn.init.constant_(m.bias, 0)


class QNetwork(nn.Module):
    """twin Q: cat[s, a](4) -> 256 -> 256 -> 1, two heads (linear1-3, linear4-6)."""

    def __init__(self, num_inputs, num_actions, hidden_dim):
        super().__init__()
        for i in (0, 3):
            setattr(self, "linear%d" % (i + 1), nn.Linear(num_inputs + num_actions, hidden_dim))
            setattr(self, "linear%d" % (i + 2), nn.Linear(hidden_dim, hidden_dim))
            setattr(self, "linear%d" % (i + 3), nn.Linear(hidden_dim, 1))
        self.apply(weights_init_)


class QNetworkConstraint(nn.Module):
    """Q_risk: same twin MLP with sigmoid outputs; carries the reference's unused BatchNorm1d(4), whose
    parameters still sit first in parameters() and take part in the Polyak average (qrisk.py:160-162)."""

    def __init__(self, num_inputs, num_actions, hidden_dim):
        super().__init__()
        self.bn1 = nn.BatchNorm1d(num_inputs + num_actions)
        for i in (0, 3):
            setattr(self, "linear%d" % (i + 1), nn.Linear(num_inputs + num_actions, hidden_dim))
            setattr(self, "linear%d" % (i + 2), nn.Linear(hidden_dim, hidden_dim))
            setattr(self, "linear%d" % (i + 3), nn.Linear(hidden_dim, 1))
        self.apply(weights_init_)


class GaussianPolicy(nn.Module):
    def __init__(self, num_inputs, num_actions, hidden_dim, action_space=None):
        super().__init__()
        self.linear1 = nn.Linear(num_inputs, hidden_dim)
        self.linear2 = nn.Linear(hidden_dim, hidden_dim)
        self.mean_linear = nn.Linear(hidden_dim, num_actions)
        self.log_std_linear = nn.Linear(hidden_dim, num_actions)
        self.apply(weights_init_)


class DeterministicPolicy(nn.Module):
    """model.py:447-485 (`--policy Deterministic`): three Linear layers; the arena keeps the Gaussian layout with
    the log_std head unused, so the parameter list is padded with two zero tensors for the loader."""

    def __init__(self, num_inputs, num_actions, hidden_dim, action_space=None):
        super().__init__()
        self.linear1 = nn.Linear(num_inputs, hidden_dim)
        self.linear2 = nn.Linear(hidden_dim, hidden_dim)
        self.mean = nn.Linear(hidden_dim, num_actions)
        self.apply(weights_init_)
        self._pad = [torch.zeros(num_actions, hidden_dim), torch.zeros(num_actions)]

    def parameters(self, recurse=True):
        return list(super().parameters(recurse)) + self._pad


class StochasticPolicy(nn.Module):
    def __init__(self, num_inputs, num_actions, hidden_dim, action_space=None):
        super().__init__()
        self.linear1 = nn.Linear(num_inputs, hidden_dim)
        self.linear2 = nn.Linear(hidden_dim, hidden_dim)
        self.mean = nn.Linear(hidden_dim, num_actions)
        # float32 as on the reference's pinned torch 1.4
        self.log_std = nn.Parameter(torch.as_tensor([np.log(0.1)] * num_actions, dtype=torch.float32))
        self.apply(weights_init_)


def build_reference_modules(hidden=256, action_scale=(1.0, 1.0), obs_dim=2, act_dim=2, deterministic=False):
    """{net name: module}, constructed in the reference's order: critic, critic_target, policy (sac.py:82-114),
    then safety_critic, safety_critic_target, recovery policy (qrisk.py:36-75).  Targets are hard copies."""
    critic = QNetwork(obs_dim, act_dim, hidden)
    critic_target = QNetwork(obs_dim, act_dim, hidden)
    critic_target.load_state_dict(critic.state_dict())
    policy = (DeterministicPolicy if deterministic else GaussianPolicy)(obs_dim, act_dim, hidden)
    qrisk = QNetworkConstraint(obs_dim, act_dim, hidden)
    qrisk_target = QNetworkConstraint(obs_dim, act_dim, hidden)
    qrisk_target.load_state_dict(qrisk.state_dict())
    recovery = StochasticPolicy(obs_dim, act_dim, hidden)
    return dict(critic=critic, critic_target=critic_target, policy=policy, qrisk=qrisk, qrisk_target=qrisk_target,
                recovery=recovery)
