"""oracle/loop.py -- TEST INFRASTRUCTURE: CPU restatement of the reference's training loop at N = 1.

Restates recovery_rl/experiment.py:261-296 (Q_risk pre-training on the constraint demos), :379-461
(get_train_rollout: update gates, composite action, env step, reward penalty, mask-before-horizon,
relabelled pushes incl. the second push of --add_both_transitions, episode statistics) and :546-577 (get_action) on top of oracle/{agent,replay,envs}.py.
It is the `cpu_baseline` / `--impl reference` arm of bench.py ("port": the reference itself is Python and
cannot travel to the GPU box) and the checker for the whole-trajectory parity test.

PINNED against tests/golden/traj_nav1_seed7.npz (a 12-episode run of the reference's own Experiment):
same states, actions, flags, replay indices and final weights when fed the recorded noise.  Maze uses the
restated physics of oracle/envs.py (parity unpinned, see there).
"""
import time

import numpy as np
import torch

from . import envs
from .agent import Agent
from .replay import SharedStream, ReplayMemory, ConstraintReplayMemory

ACTION_SCALE = {"navigation1": 1.0, "navigation2": 1.0, "maze": np.float32(0.1)}


class NoiseSource(object):
    """Gaussian / uniform draws for the loop: either recorded arrays (parity) or live numpy/torch RNGs in the
    reference's call order (torch for agent eps, numpy for env noise and resets, a Box RandomState for
    random start actions)."""

    def __init__(self, seed, eps=None, env_noise=None, rand_actions=None, categorical=None):
        self.cat = None if categorical is None else list(categorical)
        self.eps = None if eps is None else list(eps)
        self.env_noise = None if env_noise is None else list(env_noise)
        self.rand_actions = None if rand_actions is None else list(rand_actions)
        self.box_rng = np.random.RandomState(seed)

    def categorical(self, probs):
        """the SQRL action filter's Categorical draw (sac.py:156-158): recorded index, else the live torch generator"""
        if self.cat is not None:
            return int(self.cat.pop(0))
        return int(torch.distributions.Categorical(probs).sample())

    def agent_eps(self, rows):
        if self.eps is not None:
            e = self.eps.pop(0)
            assert e.shape[0] == rows, (e.shape, rows)
            return e
        return torch.randn(rows, 2).numpy()

    def det_noise(self):
        """DeterministicPolicy.sample's draw (model.py:478-479): ONE N(0, 0.1) vector per call, clamped to +-0.25.  Recorded
        draws are the raw vectors (the reference clamps out of place), live ones come from the torch global generator."""
        if self.eps is not None:
            e = np.asarray(self.eps.pop(0), np.float32).reshape(-1)
            assert e.shape == (2,), e.shape
            return np.clip(e, np.float32(-0.25), np.float32(0.25))
        from .agent import deterministic_noise
        return deterministic_noise().numpy()

    def randn2(self):
        if self.env_noise is not None:
            return self.env_noise.pop(0)
        return np.random.randn(2)

    def uniform2(self):
        return np.random.rand(2)

    def random_action(self, scale):
        if self.rand_actions is not None:
            return self.rand_actions.pop(0)
        return self.box_rng.uniform(-scale, scale, 2).astype(np.float32)


class OracleExperiment(object):
    def __init__(self, env_name, seed=0, batch_size=256, replay_size=1000000, gamma=0.99, alpha=0.2, tau=0.005, lr=3e-4,
                 gamma_safe=0.5, tau_safe=0.0002, eps_safe=0.1, use_recovery=True, mf_recovery=True, pos_fraction=-1.0,
                 constraint_reward_penalty=0.0, start_steps=100, noise=None, dgd=False, update_nu=False, rcpo=False,
                 nu=0.01, nu_schedule=False, nu_start=1e3, nu_end=0.0, num_eps=1000000, lambda_rcpo=0.01,
                 constraint_sampling=False, add_both_transitions=False, q_sampling_recovery=False, q_samples=1000,
                 deterministic=False):
        self.env_name = env_name
        self.kind = envs.KIND_BY_NAME[env_name]
        self.B = batch_size
        self.use_recovery = use_recovery
        self.gate_pos_fraction = pos_fraction
        self.pos_fraction = pos_fraction if pos_fraction >= 0 else None
        self.penalty = constraint_reward_penalty
        self.start_steps = start_steps
        self.eps_safe = eps_safe
        torch.manual_seed(seed)
        np.random.seed(seed)
        sc = ACTION_SCALE[env_name]
        self.scale = sc
        self.agent = Agent(action_scale=(sc, sc), gamma=gamma, alpha=alpha, tau=tau, gamma_safe=gamma_safe,
                           tau_safe=tau_safe, eps_safe=eps_safe, lr=lr, mf_recovery=mf_recovery, dgd=dgd,
                           update_nu=update_nu, rcpo=rcpo, nu=nu, lambda_rcpo=lambda_rcpo, deterministic=deterministic)
        # the safety critic is trained for Recovery RL and for the LR / RSPO / SQRL / RCPO comparisons (experiment.py:357-361,443)
        self.uses_qrisk = bool(use_recovery or dgd or rcpo)
        # experiment.py:80-87 + utils.py:62-64: the multiplier handed to every SAC update
        if nu_schedule:
            self.nu_fn = lambda t: nu_start + t / num_eps * (nu_end - nu_start) if t < num_eps else nu_end
        else:
            self.nu_fn = lambda t: nu
        self.i_episode = 1
        self.constraint_sampling = bool(constraint_sampling)      # SQRL action filter (sac.py:139-161)
        self.add_both = bool(add_both_transitions)                # experiment.py:446-448
        self.q_sampling = bool(q_sampling_recovery) and not mf_recovery      # qrisk.py:207-225: MF_recovery is tested first
        self.q_samples = int(q_samples)                           # qrisk.py:216, 220 hard-code 1000
        stream = SharedStream()
        self.memory = ReplayMemory(replay_size, seed, stream)
        self.recovery_memory = ConstraintReplayMemory(replay_size, seed, stream)
        self.noise = noise or NoiseSource(seed)
        self.total_numsteps = 0
        self.updates = 0
        self.num_constraint_violations = 0
        self.num_viols = self.num_successes = self.viol_and_recovery = self.viol_and_no_recovery = 0
        self.idx_log = []
        self.state = None
        self.ep_steps = 0
        self.last_losses = None

    # experiment.py:277-296
    def pretrain(self, transitions, steps, num_unsafe_transitions=None):
        n = 0
        for t in transitions:
            self.recovery_memory.push(*t)
            self.num_constraint_violations += int(t[2])
            n += 1
            if num_unsafe_transitions is not None and n == num_unsafe_transitions:
                break
        for _ in range(steps):
            self._qrisk_update(min(self.B, len(transitions)))

    def _qrisk_update(self, batch_size):
        mem = self.recovery_memory
        if self.pos_fraction:
            batch_size = min(batch_size, int((1 - self.pos_fraction) * len(mem)))
        else:
            batch_size = min(batch_size, len(mem))
        idx = mem.sample_slots(batch_size, self.pos_fraction)
        self.idx_log.append(idx)
        batch = mem.gather(idx)
        e_next = self.noise.det_noise() if self.agent.deterministic else self.noise.agent_eps(batch_size)
        e_rec = self.noise.agent_eps(batch_size) if self.agent.mf_recovery else None
        return self.agent.qrisk_update(batch, e_next, e_rec)

    def _reset(self):
        if self.kind == envs.MAZE:
            while True:
                s = envs.maze_reset_from_uniform(self.noise.uniform2())
                if not envs.maze_touch(s[0], s[1]):       # maze.py:208-212: resample while in contact
                    break
        else:
            s = envs.nav_reset(self.noise.randn2())
        self.state = s
        self.ep_steps = 0

    def _env_step(self, action):
        if self.kind == envs.MAZE:
            ns, r, d, c, su = envs.maze_step_scalar(self.state, action, self.ep_steps)
            return ns, r, d, c, su
        stuck = bool(envs.nav_obstacle(self.kind, self.state[0], self.state[1]))
        noise = np.zeros(2) if stuck else self.noise.randn2()
        ns, r, d, c, su = envs.nav_step(self.kind, self.state[None], np.asarray(action)[None], noise[None])
        return ns[0], r[0], bool(d[0]), bool(c[0]), bool(su[0])

    # experiment.py:546-577
    def _get_action(self):
        if self.start_steps > self.total_numsteps:
            action = self.noise.random_action(self.scale)
        elif self.constraint_sampling:
            e = self.noise.agent_eps(100)
            action = self.agent.select_action_sqrl(np.asarray(self.state, np.float32), e, categorical=self.noise.categorical)
        else:
            e = self.noise.det_noise()[None] if self.agent.deterministic else self.noise.agent_eps(1)
            action = self.agent.act(self.state[None], e, np.zeros((1, 2), np.float32), use_recovery=False)[0][0]
        if not self.use_recovery:
            return action, np.copy(action), False
        with torch.no_grad():
            st = torch.as_tensor(self.state[None], dtype=torch.float32)
            at = torch.as_tensor(np.asarray(action)[None], dtype=torch.float32)
            q1, q2 = self.agent.qrisk(st, at)
            risky = bool(torch.max(q1, q2) > self.eps_safe)
            if risky and self.q_sampling:
                # qrisk.py:214-225: 1000 x ac_space.sample() (the env's own action space: the stream of the random start actions)
                cands = np.array([self.noise.random_action(self.scale) for _ in range(self.q_samples)], np.float32)
                return action, self.agent.select_action_qsample(np.asarray(self.state, np.float32), cands), True
            if risky:
                e = self.noise.agent_eps(1)
                real, _, _ = self.agent.recovery.sample(st, torch.as_tensor(e, dtype=torch.float32))
                return action, real.numpy()[0], True
        return action, np.copy(action), False

    # experiment.py:396-452: ONE env step (with the updates that precede it)
    def step(self):
        if self.state is None:
            self._reset()
        if len(self.memory) > self.B:
            idx = self.memory.sample_slots(min(self.B, len(self.memory)))
            self.idx_log.append(idx)
            batch = self.memory.gather(idx)
            if self.agent.deterministic:                     # --policy Deterministic: one noise vector per sample() call
                e_next, e_cur = self.noise.det_noise(), self.noise.det_noise()
            else:
                e_next = self.noise.agent_eps(len(idx))
                e_cur = self.noise.agent_eps(len(idx))
            self.last_losses = self.agent.sac_update(batch, e_next, e_cur, self.updates, nu=self.nu_fn(self.i_episode))
            if len(self.recovery_memory) > self.B and \
                    (self.num_viols + self.num_constraint_violations) / self.B > self.gate_pos_fraction:
                self._qrisk_update(self.B)
            self.updates += 1
        action, real_action, recovery_used = self._get_action()
        state = self.state
        next_state, reward, done, constraint, success = self._env_step(real_action)
        lim = 0.1 if self.kind == envs.MAZE else 1
        info = dict(state=state, next_state=next_state, action=np.clip(np.asarray(real_action), -lim, lim), reward=reward,
                    constraint=int(constraint), success=bool(success), recovery=bool(recovery_used))
        self.ep_steps += 1
        self.total_numsteps += 1
        if constraint:
            reward = reward - self.penalty
        mask = float(not done)
        horizon = envs.MAZE_HORIZON if self.kind == envs.MAZE else envs.NAV_HORIZON
        done = done or self.ep_steps == horizon
        self.memory.push(state, action, reward, next_state, mask)
        if self.uses_qrisk:
            self.recovery_memory.push(state, real_action, float(constraint), next_state, mask)
            if recovery_used and self.add_both:                   # experiment.py:446-448
                self.memory.push(state, real_action, reward, next_state, mask)
        self.state = next_state
        if done:
            if constraint:
                self.num_viols += 1
                if recovery_used:
                    self.viol_and_recovery += 1
                else:
                    self.viol_and_no_recovery += 1
            self.num_successes += int(success)
            self.state = None
            self.i_episode += 1
        info["episode_end"] = bool(done)
        return info

    def run_steps(self, n):
        t0 = time.perf_counter()
        for _ in range(n):
            self.step()
        return time.perf_counter() - t0
