"""oracle/envs.py -- TEST INFRASTRUCTURE: CPU (numpy fp64) restatement of the point environments.

Navigation1/2: restates env/navigation1.py:50-51,71-110, env/navigation2.py:49-50,70-110 and
env/obstacle.py:13-15,44-45 of the reference.  PINNED against tests/golden/nav_step_nav{1,2}.npz
and offline_nav{1,2}.npz, which were produced by the reference's own classes
(oracle/ref_harness/make_golden.py).

Maze: PARITY UNPINNED.  The reference delegates the physics to MuJoCo 1.50 through mujoco_py
1.50.1.68 (env/maze.py:10,117,141-151; install.sh:13), a closed binary that is not vendored and
not installable here.  `maze_*` below restates env/maze.py:25-26,139-168,184-220 plus the model in
env/assets/simple_maze.xml:6-40 under the rules of SURVEY.md §8c (semi-implicit Euler with implicit
joint damping, collision phase before integration, contact = touching any wall, inelastic freeze).
The CUDA kernel is bit-exact against THIS restatement only; no MuJoCo parity is claimed.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may import this.
"""
import numpy as np

NAV1, NAV2, MAZE = 0, 1, 2
KIND_BY_NAME = {"navigation1": NAV1, "navigation2": NAV2, "maze": MAZE}

# env/navigation1.py:41-42, env/navigation2.py:41  ([[x0,x1],[y0,y1]] closed rectangles)
NAV_RECTS = {
    NAV1: [((-100.0, 150.0), (5.0, 10.0)), ((-100.0, -80.0), (-10.0, 10.0)), ((-100.0, 150.0), (-10.0, -5.0))],
    NAV2: [((-30.0, -20.0), (-7.5, 7.5))],
}
NAV_HORIZON = 100          # navigation1.py:34
NAV_NOISE_SCALE = 0.05     # navigation1.py:36
NAV_START = (-50.0, 0.0)   # navigation1.py:27


def nav_obstacle(kind, x, y):
    """obstacle.py:13-15 (closed intervals), :44-45 (max over rectangles).  Vectorised."""
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    hit = np.zeros(np.broadcast(x, y).shape, bool)
    for (x0, x1), (y0, y1) in NAV_RECTS[kind]:
        hit |= (x0 <= x) & (x <= x1) & (y0 <= y) & (y <= y1)
    return hit


def nav_step(kind, state, action, noise):
    """Navigation*.step (navigation1.py:71-89) for a batch of independent envs.
    state [n,2] f64, action [n,2] f32 (any float), noise [n,2] f64 standard normal.
    Returns next_state f64, reward f64, done, constraint, success (bool arrays); `done` excludes the
    horizon truncation, which the caller applies (experiment.py:435)."""
    s = np.asarray(state, np.float64)
    a = np.clip(np.asarray(action), -1, 1)                       # :50-51 (dtype preserved)
    stuck = nav_obstacle(kind, s[:, 0], s[:, 1])                  # :99-101
    moved = (s + a.astype(np.float64)) + NAV_NOISE_SCALE * np.asarray(noise, np.float64)  # :102-104
    ns = np.where(stuck[:, None], s, moved)
    cost = np.array([-np.linalg.norm(np.subtract((0, 0), s[i])) for i in range(len(s))], np.float64)
    constraint = nav_obstacle(kind, ns[:, 0], ns[:, 1])
    success = cost > -4
    done = success | constraint
    return ns, cost, done, constraint, success


def nav_reset(draws):
    """navigation1.py:91-92: START_STATE + np.random.randn(2)."""
    return np.asarray(NAV_START, np.float64) + np.asarray(draws, np.float64)


class _NavOfflineEnv(object):
    def __init__(self, kind):
        self.kind = kind

    def obstacle(self, s):
        return int(nav_obstacle(self.kind, s[0], s[1]))

    def next_state(self, s, a):
        if self.obstacle(s):
            return s
        return np.eye(2).dot(s) + np.eye(2).dot(a) + NAV_NOISE_SCALE * np.random.randn(len(s))


def _nav_rollout(env, state, action_fn, transitions):
    for _ in range(10):
        action = action_fn()
        next_state = env.next_state(state, action)
        constraint = env.obstacle(next_state)
        transitions.append((state, action, constraint, next_state, not constraint))
        state = next_state
        if constraint:
            break


def nav_offline_data(kind, num_transitions):
    """get_offline_data of navigation1.py:133-164 / navigation2.py:133-243, drawing from the numpy
    GLOBAL RandomState in the same order as the reference."""
    env = _NavOfflineEnv(kind)
    tr = []
    U, R = np.random.uniform, np.random.randn
    if kind == NAV1:
        for _ in range(num_transitions // 10):
            if U(0, 1) < 0.5:
                state = np.array([U(-80, 50), U(-5, -2)])
            else:
                state = np.array([U(-80, 50), U(2, 5)])
            _nav_rollout(env, state, lambda: np.clip(R(2), -1, 1), tr)
        return tr
    for _ in range(num_transitions // 10 // 3):
        state = np.array([U(-40, 10), U(-25, 25)])
        while env.obstacle(state):
            state = np.array([U(-40, 10), U(-25, 25)])
        _nav_rollout(env, state, lambda: np.clip(R(2), -1, 1), tr)
    quarter = num_transitions // 10 * 1 // 4
    regions = [
        ((-35, -30), (-12, 12), lambda: np.clip(np.array([U(0.5, 1, 1), R(1)]), -1, 1).ravel()),
        ((-20, -15), (-12, 12), lambda: np.clip(np.array([U(-1, -0.5, 1), R(1)]), -1, 1).ravel()),
        ((-30, -20), (10, 15), lambda: np.clip(np.array([R(1), U(-1, -0.5, 1)]), -1, 1).ravel()),
        ((-30, -20), (-15, -10), lambda: np.clip(np.array([R(1), U(0.5, 1, 1)]), -1, 1).ravel()),
    ]
    for xr, yr, act in regions:
        for _ in range(quarter):
            state = np.array([U(*xr), U(*yr)])
            _nav_rollout(env, state, act, tr)
    return tr


# ------------------------------------------------------------------------------------------------
# Maze (restated physics; see module docstring)
# ------------------------------------------------------------------------------------------------
MAZE_HORIZON = 100                 # maze.py:16
MAZE_MAX_FORCE = 0.1               # maze.py:17 (np.clip keeps the policy's float32: bound = float32(0.1))
MAZE_GOAL_THRESH = 3e-2            # maze.py:19
MAZE_R = 0.025                     # simple_maze.xml:28 cylinder radius
MAZE_SUBSTEPS = 500                # maze.py:147
MAZE_H = 0.002                     # MuJoCo default timestep
MAZE_GEAR = 0.05                   # simple_maze.xml:8
MAZE_DAMPING = 0.01                # simple_maze.xml:29-30
MAZE_MASS = 1000.0 * 3.141592653589793 * 0.025 * 0.025 * 0.05   # default density * cylinder volume
MAZE_CA = MAZE_MASS / (MAZE_MASS + MAZE_H * MAZE_DAMPING)
MAZE_CB = MAZE_H / (MAZE_MASS + MAZE_H * MAZE_DAMPING)
MAZE_GOAL = (0.25, 0.0)            # maze.py:135-137


def maze_walls():
    """simple_maze.xml:22-25 (boxes, half-sizes (.02,.2,.005), thin axis along world x) with the y
    centres written by MazeNavigation.reset (maze.py:199-206).  Order: 1A, 1B, 2A, 2B."""
    w1, w2 = -0.08, 0.08
    cy = [0.5 + w1, -0.25 + w1, 0.4 + w2, -0.25 + w2]
    cx = [-0.1, -0.1, 0.1, 0.1]
    return [(cx[i] - 0.005, cx[i] + 0.005, cy[i] - 0.2, cy[i] + 0.2) for i in range(4)]


def maze_touch(x, y):
    """ncon > 3 (maze.py:144,150): disc of radius MAZE_R touching an outer plane (closed) or
    overlapping one of the four wall rectangles (strict).  Vectorised."""
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    hit = (x - MAZE_R <= -0.3) | (x + MAZE_R >= 0.3) | (y - MAZE_R <= -0.3) | (y + MAZE_R >= 0.3)
    for x0, x1, y0, y1 in maze_walls():
        dx = np.maximum(np.maximum(x0 - x, 0.0), x - x1)
        dy = np.maximum(np.maximum(y0 - y, 0.0), y - y1)
        hit = hit | ((dx * dx + dy * dy) < (MAZE_R * MAZE_R))
    return hit


def maze_distance(x, y):
    """get_distance_score maze.py:215-220: sqrt(mean((goal - qpos)**2)), unfused."""
    d0 = MAZE_GOAL[0] - np.asarray(x, np.float64)
    d1 = MAZE_GOAL[1] - np.asarray(y, np.float64)
    return np.sqrt((d0 * d0 + d1 * d1) * 0.5)


def maze_step(state, action, ep_steps, substeps=MAZE_SUBSTEPS):
    """MazeNavigation.step (maze.py:139-168) for a batch.  ep_steps = steps taken BEFORE this one.
    Returns next_state, reward, done (includes steps >= horizon, maze.py:152), constraint, success."""
    s = np.asarray(state, np.float64)
    a = np.asarray(action)
    a = np.clip(a, -0.1, 0.1)                # maze.py:25-26; dtype preserved (fp32 from the policy)
    fbx = MAZE_CB * (MAZE_GEAR * a[:, 0].astype(np.float64))
    fby = MAZE_CB * (MAZE_GEAR * a[:, 1].astype(np.float64))
    x = s[:, 0].copy()
    y = s[:, 1].copy()
    vx = np.zeros_like(x)
    vy = np.zeros_like(y)
    contact = np.zeros(x.shape, bool)
    for _ in range(substeps):
        contact |= maze_touch(x, y)          # collision phase precedes integration; contact freezes
        live = ~contact
        vx_n = MAZE_CA * vx + fbx
        vy_n = MAZE_CA * vy + fby
        x_n = x + MAZE_H * vx_n
        y_n = y + MAZE_H * vy_n
        vx = np.where(live, vx_n, vx)
        vy = np.where(live, vy_n, vy)
        x = np.where(live, x_n, x)
        y = np.where(live, y_n, y)
    dist = maze_distance(x, y)
    reward = -dist
    steps = np.asarray(ep_steps) + 1
    done = (steps >= MAZE_HORIZON) | contact | (dist < MAZE_GOAL_THRESH)
    success = reward > -0.03
    return np.stack([x, y], 1), reward, done, contact, success


def maze_reset_from_uniform(u, difficulty="h"):
    """maze.py:189-197: np.random.uniform(lo, hi) = lo + (hi - lo) * u, u ~ U[0,1)."""
    u = np.asarray(u, np.float64)
    lo, hi = {"e": (0.14, 0.22), "m": (-0.04, 0.04), "h": (-0.22, -0.13), None: (-0.27, 0.27)}[difficulty]
    x = lo + (hi - lo) * u[..., 0]
    y = -0.22 + (0.22 - (-0.22)) * u[..., 1]
    return np.stack([x, y], -1)


def maze_expert_action(st):
    """maze.py:222-232 waypoint P-controller."""
    st = np.asarray(st, np.float64)
    if st[0] <= -0.151:
        delt = np.array([-0.15, -0.125]) - st
    elif st[0] <= 0.149:
        delt = np.array([0.15, 0.125]) - st
    else:
        delt = np.array(MAZE_GOAL) - st
    return 1.05 * delt


def maze_offline_data(num_transitions, action_rng=None):
    """get_offline_data maze.py:34-107 on the restated env: half random actions, half expert,
    20-step segments with a fresh reset (random difficulty, check_constraint=False), no break on
    contact.  Random actions come from `action_rng` (the reference uses gym's Box.sample, whose
    stream is unpinned -- SURVEY.md §8c); resets come from the numpy global RandomState."""
    if action_rng is None:
        action_rng = np.random.RandomState(0)
    tr = []
    for phase in range(2):
        state = None
        steps = 0
        for i in range(num_transitions // 2):
            if i % 20 == 0:
                sample = np.random.uniform(0, 1, 1)[0]
                mode = "e" if sample < 0.3 else ("m" if sample < 0.6 else "h")
                lo, hi = {"e": (0.14, 0.22), "m": (-0.04, 0.04), "h": (-0.22, -0.13)}[mode]
                x = np.random.uniform(lo, hi)
                y = np.random.uniform(-0.22, 0.22)
                state = np.array([x, y])
                steps = 0
            if phase == 0:
                action = action_rng.uniform(-0.1, 0.1, 2).astype(np.float32)
            else:
                action = maze_expert_action(state)
            ns, _, done, cons, _ = maze_step(state[None], np.asarray(action)[None], np.array([steps]))
            steps += 1
            tr.append((state, action, int(cons[0]), ns[0], not bool(done[0])))
            state = ns[0]
    return tr


_WALLS = None


def maze_step_scalar(state, action, ep_steps, substeps=MAZE_SUBSTEPS):
    """maze_step for ONE env in plain Python floats (same IEEE-754 double operations, same order) -- the
    N = 1 CPU baseline loop uses it; tests check it against the vectorised maze_step."""
    global _WALLS
    if _WALLS is None:
        _WALLS = maze_walls()
    a = np.clip(np.asarray(action), -0.1, 0.1)
    fbx = MAZE_CB * (MAZE_GEAR * float(a[0]))
    fby = MAZE_CB * (MAZE_GEAR * float(a[1]))
    x, y = float(state[0]), float(state[1])
    vx = vy = 0.0
    contact = False
    r2 = MAZE_R * MAZE_R
    for _ in range(substeps):
        if (x - MAZE_R <= -0.3) or (x + MAZE_R >= 0.3) or (y - MAZE_R <= -0.3) or (y + MAZE_R >= 0.3):
            contact = True
            break
        hit = False
        for x0, x1, y0, y1 in _WALLS:
            dx = max(max(x0 - x, 0.0), x - x1)
            dy = max(max(y0 - y, 0.0), y - y1)
            if dx * dx + dy * dy < r2:
                hit = True
                break
        if hit:
            contact = True
            break
        vx = MAZE_CA * vx + fbx
        vy = MAZE_CA * vy + fby
        x = x + MAZE_H * vx
        y = y + MAZE_H * vy
    d0 = MAZE_GOAL[0] - x
    d1 = MAZE_GOAL[1] - y
    dist = float(np.sqrt((d0 * d0 + d1 * d1) * 0.5))
    reward = -dist
    done = (ep_steps + 1 >= MAZE_HORIZON) or contact or (dist < MAZE_GOAL_THRESH)
    return np.array([x, y]), reward, bool(done), bool(contact), bool(reward > -0.03)
