"""Small utilities with the reference's names (recovery_rl/utils.py:46-88)."""


def soft_update(target, source, tau):
    """utils.py:46-49 on device nets: `target` / `source` are (arena, net_name) handles (see sac.py)."""
    target.soft_update_from(source, tau)


def hard_update(target, source):
    target.hard_update_from(source)


def linear_schedule(startval, endval, endtime):
    return lambda t: startval + t / endtime * (endval - startval) if t < endtime else endval


def get_required_argument(dotmap, key, message, default=None):
    val = dotmap.get(key, default)
    if val is default:
        raise ValueError(message)
    return val


def recovery_config_setup(exp_cfg, logdir):
    """utils.py:84-88: controller configuration of the model-based recovery policy."""
    from config import create_config
    from recovery_rl.dotmap_lite import DotMap
    ctrl_args = DotMap(**{key: val for (key, val) in exp_cfg.ctrl_arg})
    cfg = create_config(exp_cfg.env_name, "MPC", ctrl_args, exp_cfg.override, logdir)
    return cfg
