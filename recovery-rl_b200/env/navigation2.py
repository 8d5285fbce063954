"""Navigation2 (reference env/navigation2.py): same dynamics as Navigation1 with one central obstacle
[[-30,-20],[-7.5,7.5]]; offline data from five start regions with biased actions (:133-243)."""
import numpy as np

from env.obstacle import ComplexObstacle
from env.navigation1 import _NavigationBase, _rollout, process_action, START_STATE, GOAL_STATE, MAX_FORCE, HORIZON  # noqa: F401

OBSTACLE = ComplexObstacle([[[-30, -20], [-7.5, 7.5]]])
CAUTION_ZONE = ComplexObstacle([[[-32, -18], [-12, 12]]])
ENV_NAME = "navigation2"


class Navigation2(_NavigationBase):
    env_name = ENV_NAME
    obstacle = OBSTACLE
    caution_zone = CAUTION_ZONE

    def __init__(self):
        _NavigationBase.__init__(self)
        self.transition_function = get_offline_data


def get_offline_data(num_transitions, task_demos=False, save_rollouts=False):
    env = Navigation2()
    transitions, rollouts = [], []
    U, R = np.random.uniform, np.random.randn
    for _ in range(num_transitions // 10 // 3):
        rollouts.append([])
        state = np.array([U(-40, 10), U(-25, 25)])
        while env.obstacle(state):
            state = np.array([U(-40, 10), U(-25, 25)])
        _rollout(env, state, lambda: np.clip(R(2), -1, 1), transitions, rollouts)
    regions = [
        ((-35, -30), (-12, 12), lambda: np.clip(np.array([U(0.5, 1, 1), R(1)]), -1, 1).ravel()),
        ((-20, -15), (-12, 12), lambda: np.clip(np.array([U(-1, -0.5, 1), R(1)]), -1, 1).ravel()),
        ((-30, -20), (10, 15), lambda: np.clip(np.array([R(1), U(-1, -0.5, 1)]), -1, 1).ravel()),
        ((-30, -20), (-15, -10), lambda: np.clip(np.array([R(1), U(0.5, 1, 1)]), -1, 1).ravel()),
    ]
    for xr, yr, act in regions:
        for _ in range(num_transitions // 10 * 1 // 4):
            rollouts.append([])
            state = np.array([U(*xr), U(*yr)])
            _rollout(env, state, act, transitions, rollouts)
    return rollouts if save_rollouts else transitions
