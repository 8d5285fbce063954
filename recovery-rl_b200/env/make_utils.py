"""env id <-> class maps (reference env/make_utils.py:1-31), without the gym registry."""
ENV_ID = {
    'navigation1': 'Navigation-v0',
    'navigation2': 'Navigation-v1',
    'maze': 'Maze-v0',
}

ENV_CLASS = {
    'navigation1': 'Navigation1',
    'navigation2': 'Navigation2',
    'maze': 'MazeNavigation',
}

_REGISTRY = {}


def register_env(env_name):
    assert env_name in ENV_ID, "unknown environment"
    import importlib
    module = importlib.import_module("env." + env_name)
    _REGISTRY[ENV_ID[env_name]] = getattr(module, ENV_CLASS[env_name])


def make_env(env_name):
    if ENV_ID[env_name] not in _REGISTRY:
        register_env(env_name)
    return _REGISTRY[ENV_ID[env_name]]()
