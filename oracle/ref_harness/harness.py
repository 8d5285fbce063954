"""Out-of-tree driver that makes the UNMODIFIED reference (/root/reference) run on CPU.

TEST INFRASTRUCTURE ONLY (never imported by the product path).  It exists so that
golden vectors can be generated from the reference's own code in the build
container (the reference cannot travel to the GPU box), see `make_golden.py`.

What it does:
  * puts inert/functional shim packages (gym, dotmap, matplotlib, moviepy, plotly,
    mujoco_py) and /root/reference on sys.path;
  * P1  env/navigation1.py:63      np.float removed in numpy>=1.24  -> np.float = float
  * P2  recovery_rl/model.py:497   StochasticPolicy.log_std becomes float64 on modern torch
                                   -> cast to float32 after construction (torch 1.4 behaviour)
  * P3  recovery_rl/experiment.py:26  torchify hard-codes .to('cuda') -> CPU tensor
  * P4  recovery_rl/sac.py:233-239 critic_optim.step() before policy_loss.backward() is an
        in-place-modification error on torch>=1.5.  "Variant B": every forward expression is
        evaluated exactly as written (:192-231), then critic grads, policy grads (w.r.t. the
        policy parameters only), critic step, policy step.
  * noise capture/injection: every agent-side Gaussian draw goes through
    torch.distributions.normal._standard_normal; env noise through np.random.randn.
"""
import os
import sys
import collections

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("RRL_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIMS = os.path.join(_HERE, "shims")

_state = {"ready": False}

# ---------------------------------------------------------------------------------------
# noise plumbing
# ---------------------------------------------------------------------------------------
eps_log = []            # every _standard_normal draw, in order (np arrays)
eps_queue = collections.deque()   # if non-empty, draws are popped from here instead


def _install_noise_hooks():
    import torch.distributions.normal as tdn
    orig = tdn._standard_normal

    def hooked(shape, dtype, device):
        if eps_queue:
            e = eps_queue.popleft()
            e = torch.as_tensor(e, dtype=dtype, device=device)
            assert tuple(e.shape) == tuple(shape), (tuple(e.shape), tuple(shape))
        else:
            e = orig(shape, dtype=dtype, device=device)
        eps_log.append(e.detach().cpu().numpy().copy())
        return e

    tdn._standard_normal = hooked


# ---------------------------------------------------------------------------------------
# P4: Variant-B restatement of SAC.update_parameters (reference sac.py:170-277)
# ---------------------------------------------------------------------------------------
def _sac_update_variant_b(self, memory, batch_size, updates, nu=None, safety_critic=None):
    import torch.nn.functional as F
    from recovery_rl.utils import soft_update
    if nu is None:
        nu = self.nu
    state_batch, action_batch, reward_batch, next_state_batch, mask_batch = memory.sample(
        batch_size=batch_size)
    state_batch = torch.FloatTensor(state_batch).to(self.device)
    next_state_batch = torch.FloatTensor(next_state_batch).to(self.device)
    action_batch = torch.FloatTensor(action_batch).to(self.device)
    reward_batch = torch.FloatTensor(reward_batch).to(self.device).unsqueeze(1)
    mask_batch = torch.FloatTensor(mask_batch).to(self.device).unsqueeze(1)

    with torch.no_grad():
        next_state_action, next_state_log_pi, _ = self.policy.sample(next_state_batch)
        qf1_next_target, qf2_next_target = self.critic_target(next_state_batch, next_state_action)
        min_qf_next_target = torch.min(qf1_next_target, qf2_next_target) - self.alpha * next_state_log_pi
        next_q_value = reward_batch + mask_batch * self.gamma * (min_qf_next_target)
        if self.RCPO:
            qsafe_batch = torch.max(*safety_critic(state_batch, action_batch))
            next_q_value -= self.lambda_RCPO * qsafe_batch
    qf1, qf2 = self.critic(state_batch, action_batch)
    qf1_loss = F.mse_loss(qf1, next_q_value)
    qf2_loss = F.mse_loss(qf2, next_q_value)

    pi, log_pi, _ = self.policy.sample(state_batch)
    qf1_pi, qf2_pi = self.critic(state_batch, pi)
    min_qf_pi = torch.min(qf1_pi, qf2_pi)
    sqf1_pi, sqf2_pi = self.safety_critic(state_batch, pi)
    max_sqf_pi = torch.max(sqf1_pi, sqf2_pi)
    if self.DGD_constraints:
        policy_loss = ((self.alpha * log_pi) + nu * (max_sqf_pi - self.eps_safe) - 1. * min_qf_pi).mean()
    else:
        policy_loss = ((self.alpha * log_pi) - min_qf_pi).mean()

    # --- Variant B ordering -----------------------------------------------------------
    self.critic_optim.zero_grad()
    (qf1_loss + qf2_loss).backward(retain_graph=True)
    pol_params = list(self.policy.parameters())
    pol_grads = torch.autograd.grad(policy_loss, pol_params)
    self.critic_optim.step()
    self.policy_optim.zero_grad()
    for p, g in zip(pol_params, pol_grads):
        p.grad = g
    self.policy_optim.step()
    # ------------------------------------------------------------------------------------

    if self.automatic_entropy_tuning:
        alpha_loss = -(self.log_alpha * (log_pi + self.target_entropy).detach()).mean()
        self.alpha_optim.zero_grad()
        alpha_loss.backward()
        self.alpha_optim.step()
        self.alpha = self.log_alpha.exp()
        alpha_tlogs = self.alpha.clone()
    else:
        alpha_loss = torch.tensor(0.).to(self.device)
        alpha_tlogs = torch.tensor(self.alpha)
    if self.update_nu:
        nu_loss = (self.log_nu * (self.eps_safe - max_sqf_pi).detach()).mean()
        self.nu_optim.zero_grad()
        nu_loss.backward()
        self.nu_optim.step()
        self.nu = self.log_nu.exp()
    if self.RCPO:
        lambda_RCPO_loss = (self.log_lambda_RCPO * (self.eps_safe - qsafe_batch).detach()).mean()
        self.lambda_RCPO_optim.zero_grad()
        lambda_RCPO_loss.backward()
        self.lambda_RCPO_optim.step()
        self.lambda_RCPO = self.log_lambda_RCPO.exp()
    if updates % self.target_update_interval == 0:
        soft_update(self.critic_target, self.critic, self.tau)

    # extra outputs for the golden dump (not part of the reference's return value)
    self._dbg = dict(qf1=qf1.detach().numpy().copy(), qf2=qf2.detach().numpy().copy(),
                     target=next_q_value.detach().numpy().copy(),
                     pi=pi.detach().numpy().copy(), log_pi=log_pi.detach().numpy().copy(),
                     min_qf_pi=min_qf_pi.detach().numpy().copy(),
                     next_action=next_state_action.numpy().copy(),
                     next_log_pi=next_state_log_pi.numpy().copy(),
                     critic_grads=[p.grad.detach().numpy().copy() for p in self.critic.parameters()],
                     policy_grads=[g.detach().numpy().copy() for g in pol_grads])
    return qf1_loss.item(), qf2_loss.item(), policy_loss.item(), alpha_loss.item(), alpha_tlogs.item()


def setup():
    """Idempotent: shims + patches + import of the reference modules."""
    if _state["ready"]:
        return
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError("reference tree not found at %s (golden vectors can only be "
                           "regenerated in the build container)" % REFERENCE_ROOT)
    np.float = float                                   # P1
    sys.path[:0] = [_SHIMS, REFERENCE_ROOT]
    import warnings
    warnings.filterwarnings("ignore")
    import recovery_rl.model as model
    import recovery_rl.experiment as experiment
    import recovery_rl.sac as sac

    orig_init = model.StochasticPolicy.__init__

    def patched_init(self, *a, **k):                   # P2
        orig_init(self, *a, **k)
        self.log_std.data = self.log_std.data.float()

    model.StochasticPolicy.__init__ = patched_init
    experiment.torchify = lambda x: torch.FloatTensor(x)   # P3
    sac.SAC.update_parameters = _sac_update_variant_b      # P4
    _install_noise_hooks()
    _state["ready"] = True


def get_args(argv):
    setup()
    import arg_utils
    old = sys.argv
    sys.argv = ["rrl_main"] + list(argv)
    try:
        return arg_utils.get_args()
    finally:
        sys.argv = old


def make_experiment(argv):
    setup()
    from recovery_rl.experiment import Experiment
    return Experiment(get_args(argv))
