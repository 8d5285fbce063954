"""Environment factory of the drop-in surface: `register_env(name)` / `make_env(name)` as called by
experiment.py:157-158 (reference env/make_utils.py:1-31 does this through the gym registry; here a plain table maps
the three point envs of the hot path to their host-side classes, which step through librrl.so)."""
import importlib

# env-name flag -> (gym id the reference registers, module under env/, class)
_TABLE = {
    'navigation1': ('Navigation-v0', 'env.navigation1', 'Navigation1'),
    'navigation2': ('Navigation-v1', 'env.navigation2', 'Navigation2'),
    'maze': ('Maze-v0', 'env.maze', 'MazeNavigation'),
}
ENV_ID = {name: row[0] for name, row in _TABLE.items()}
ENV_CLASS = {name: row[2] for name, row in _TABLE.items()}
_constructors = {}


def register_env(env_name):
    if env_name not in _TABLE:
        raise AssertionError("unknown environment %r (have: %s)" % (env_name, ", ".join(sorted(_TABLE))))
    _, module, cls = _TABLE[env_name]
    _constructors[env_name] = getattr(importlib.import_module(module), cls)
    return _constructors[env_name]


def make_env(env_name):
    ctor = _constructors.get(env_name) or register_env(env_name)
    return ctor()
