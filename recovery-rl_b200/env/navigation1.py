"""Navigation1 (reference env/navigation1.py): 2-D point robot, s' = s + clip(a) + 0.05 N(0, I) in fp64,
three rectangular obstacles, reward -||s||.  Same class / attribute / function names as the reference;
the arithmetic of step() runs in the CUDA kernel (csrc/env.cu) through PointEnvBatch, the random draws
come from the numpy global RandomState in the reference's order (so seeded runs and the offline-data
generator reproduce the reference's streams bit for bit).
"""
import numpy as np

from env.obstacle import ComplexObstacle
from env.spaces import Box
from env.vec_env import PointEnvBatch

START_POS = [-50, 0]
END_POS = [0, 0]
GOAL_THRESH = 1.
START_STATE = START_POS
GOAL_STATE = END_POS
MAX_FORCE = 1
HORIZON = 100
NOISE_SCALE = 0.05
HARD_MODE = False
OBSTACLE = ComplexObstacle([[[-100, 150], [5, 10]], [[-100, -80], [-10, 10]], [[-100, 150], [-10, -5]]])
CAUTION_ZONE = ComplexObstacle([[[-100, 150], [4, 5]], [[-100, 150], [-5, -4]]])
ENV_NAME = "navigation1"


def process_action(a):
    return np.clip(a, -MAX_FORCE, MAX_FORCE)


class _NavigationBase(object):
    env_name = None
    obstacle = None
    caution_zone = None

    def __init__(self):
        self.hist = self.cost = self.done = self.time = self.state = None
        self.A = np.eye(2)
        self.B = np.eye(2)
        self.horizon = HORIZON
        self.action_space = Box(-np.ones(2) * MAX_FORCE, np.ones(2) * MAX_FORCE)
        self.observation_space = Box(-np.ones(2) * float('inf'), np.ones(2) * float('inf'))
        self._max_episode_steps = HORIZON
        self.goal = GOAL_STATE
        self._dev = PointEnvBatch(self.env_name, n=1, horizon=1 << 30)   # horizon is the caller's (experiment.py:435)

    def seed(self, seed=None):
        return [seed]

    def step(self, a):
        a = process_action(a)
        old_state = self.state.copy()
        self._dev.set_state(self.state)
        noise = np.zeros(2) if self._stuck(self.state) else np.random.randn(len(self.state))
        ns, cost, done, constraint, success = self._dev.step(a, noise)
        next_state = ns[0]
        cur_cost = cost[0]
        self.cost.append(cur_cost)
        self.state = next_state
        self.time += 1
        self.hist.append(self.state)
        self.done = bool(done[0])
        return self.state, cur_cost, self.done, {
            "constraint": int(constraint[0]),
            "reward": cur_cost,
            "state": old_state,
            "next_state": next_state,
            "action": a,
            "success": bool(success[0])
        }

    def _stuck(self, s):
        # the reference draws NO noise when the state is already inside an obstacle (navigation1.py:99-101)
        return bool(self.obstacle(s))

    def reset(self):
        self.state = self._dev.reset_from_draws(np.random.randn(2))[0]
        self.time = 0
        self.cost = []
        self.done = False
        self.hist = [self.state]
        return self.state

    def _next_state(self, s, a, override=False):
        """(A s + B a) + 0.05 n on the device; no draw and no motion when s is inside an obstacle."""
        if self._stuck(s):
            return s
        self._dev.set_state(s)
        ns, _, _, _, _ = self._dev.step(a, np.random.randn(len(s)))
        return ns[0]

    def step_cost(self, s, a):
        self._dev.set_state(s)
        _, cost, _, _, _ = self._dev.step(np.zeros(2, np.float32), np.zeros(2))
        return cost[0]

    def sample(self):
        return np.random.random(2) * 2 * MAX_FORCE - MAX_FORCE


class Navigation1(_NavigationBase):
    env_name = ENV_NAME
    obstacle = OBSTACLE
    caution_zone = CAUTION_ZONE

    def __init__(self):
        _NavigationBase.__init__(self)
        self.transition_function = get_offline_data


def _rollout(env, state, action_fn, transitions, rollouts):
    for _ in range(10):
        action = action_fn()
        next_state = env._next_state(state, action, override=True)
        constraint = env.obstacle(next_state)
        transitions.append((state, action, constraint, next_state, not constraint))
        rollouts[-1].append((state, action, constraint, next_state, not constraint))
        state = next_state
        if constraint:
            break


def get_offline_data(num_transitions, task_demos=False, save_rollouts=False):
    """Constraint demos (reference navigation1.py:133-164): num//10 rollouts of <= 10 steps from the
    corridor edges with clipped Gaussian actions, stopping at the first violation."""
    env = Navigation1()
    transitions, rollouts = [], []
    for _ in range(num_transitions // 10):
        rollouts.append([])
        if np.random.uniform(0, 1) < 0.5:
            state = np.array([np.random.uniform(-80, 50), np.random.uniform(-5, -2)])
        else:
            state = np.array([np.random.uniform(-80, 50), np.random.uniform(2, 5)])
        _rollout(env, state, lambda: np.clip(np.random.randn(2), -1, 1), transitions, rollouts)
    return rollouts if save_rollouts else transitions
