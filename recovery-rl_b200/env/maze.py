"""MazeNavigation (reference env/maze.py): planar disc pushed by two motors through a two-wall maze.

The reference runs the physics in MuJoCo 1.50 (closed binary, absent here).  The step arithmetic is the
restatement documented in DESIGN.md / oracle/envs.py, executed by the CUDA kernel (csrc/env.cu): 500
substeps of semi-implicit Euler with implicit joint damping per env step, contact = touching any wall.
Names, constants and the info dict follow env/maze.py:14-26,110-232.
"""
import numpy as np

from env.spaces import Box
from env.vec_env import PointEnvBatch

HORIZON = 100
MAX_FORCE = 0.1
FAILURE_COST = 0
GOAL_THRESH = 3e-2
GT_STATE = True
DENSE_REWARD = True
_RESET_RANGES = {'e': (0.14, 0.22), 'm': (-0.04, 0.04), 'h': (-0.22, -0.13), None: (-0.27, 0.27)}


def process_action(a):
    return np.clip(a, -MAX_FORCE, MAX_FORCE)


class MazeNavigation(object):
    def __init__(self, n=1):
        self.horizon = HORIZON
        self._max_episode_steps = self.horizon
        self.transition_function = get_offline_data
        self.steps = 0
        self.images = not GT_STATE
        self.action_space = Box(-MAX_FORCE * np.ones(2), MAX_FORCE * np.ones(2))
        self.observation_space = Box(-0.3, 0.3, shape=(2,))
        self.dense_reward = DENSE_REWARD
        self.gain = 1.05
        self.goal = np.array([0.25, 0.0])
        self._dev = PointEnvBatch("maze", n=1, horizon=HORIZON)
        self.qpos = np.zeros(2)
        self._in_contact = False

    def seed(self, seed=None):
        return [seed]

    def _get_obs(self, images=False):
        return self.qpos.copy()

    def step(self, action):
        action = process_action(action)
        cur_obs = self._get_obs()
        self._dev.set_state(self.qpos, ep_steps=[self.steps])
        ns, reward, done, constraint, success = self._dev.step(action)
        self.qpos = ns[0]
        self.steps += 1
        self.done = bool(done[0])
        obs = self._get_obs()
        info = {
            "constraint": int(constraint[0]),
            "reward": reward[0],
            "state": cur_obs,
            "next_state": obs,
            "action": action,
            "success": bool(success[0])
        }
        return obs, reward[0], self.done, info

    def _touching(self):
        """ncon > 3 at the current position: a zero-force step reports the contact flag without moving."""
        self._dev.set_state(self.qpos, ep_steps=[0])
        _, _, _, constraint, _ = self._dev.step(np.zeros(2, np.float32))
        return bool(constraint[0])

    def reset(self, difficulty='h', check_constraint=True, pos=()):
        if len(pos):
            self.qpos = np.array([pos[0], pos[1]], np.float64)
        else:
            lo, hi = _RESET_RANGES[difficulty]
            x = np.random.uniform(lo, hi)
            y = np.random.uniform(-0.22, 0.22)
            self.qpos = np.array([x, y])
        self.steps = 0
        if check_constraint and not len(pos) and self._touching():
            self.reset(difficulty)
        return self._get_obs()

    def get_distance_score(self):
        return np.sqrt(np.mean((self.goal - self.qpos) ** 2))

    def expert_action(self):
        st = self.qpos
        if st[0] <= -0.151:
            delt = (np.array([-0.15, -0.125]) - st)
        elif st[0] <= 0.149:
            delt = (np.array([0.15, 0.125]) - st)
        else:
            delt = (np.array([self.goal[0], self.goal[1]]) - st)
        return self.gain * delt


def _expert_actions(states, goal=(0.25, 0.0), gain=1.05):
    st = np.asarray(states, np.float64)
    tgt = np.where((st[:, 0] <= -0.151)[:, None], np.array([-0.15, -0.125]),
                   np.where((st[:, 0] <= 0.149)[:, None], np.array([0.15, 0.125]), np.array(goal)))
    return gain * (tgt - st)


def get_offline_data(num_transitions, images=False, save_rollouts=False, rng=None):
    """Constraint demos (reference maze.py:34-107): num//2 transitions with uniformly random actions and
    num//2 with the waypoint expert, in 20-step segments that start from a fresh reset of random difficulty
    (30% 'e', 30% 'm', 40% 'h', no contact check) and keep stepping after a contact.  The segments are
    independent, so they are stepped as one batch on the device (one env copy per segment).  The reference
    ships these demos as a pickle produced with MuJoCo; here they are regenerated with the restated physics."""
    rng = rng if rng is not None else np.random
    half = num_transitions // 2
    out = []
    for phase in (0, 1):
        n_seg = (half + 19) // 20
        if n_seg == 0:
            continue
        u = rng.uniform(0, 1, n_seg)
        lo = np.where(u < 0.3, 0.14, np.where(u < 0.6, -0.04, -0.22))
        hi = np.where(u < 0.3, 0.22, np.where(u < 0.6, 0.04, -0.13))
        state = np.stack([rng.uniform(lo, hi), rng.uniform(-0.22, 0.22, n_seg)], 1)
        env = PointEnvBatch("maze", n=n_seg, horizon=HORIZON)
        env.set_state(state)
        seg = [[] for _ in range(n_seg)]
        for t in range(20):
            if phase == 0:
                action = rng.uniform(-MAX_FORCE, MAX_FORCE, (n_seg, 2)).astype(np.float32)
            else:
                action = _expert_actions(state)
            ns, _, done, cons, _ = env.step(action)
            for i in range(n_seg):
                seg[i].append((state[i], action[i], int(cons[i]), ns[i], not bool(done[i])))
            state = ns
        flat = [tr for s in seg for tr in s][:half]
        out.append((seg, flat))
    if save_rollouts:
        return [s for seg, _ in out for s in seg]
    return [tr for _, flat in out for tr in flat]
