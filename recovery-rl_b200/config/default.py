"""config/default.py of the reference (:15-119): builds the controller configuration tree for an env.
As in the reference, `ctrl_args` / `overrides` (-ca / -o) are accepted and never applied (default.py:15-51 ignores
them), the model is a 5-net probabilistic ensemble ('PE'), propagation is TS-infinity with 20 particles, and the
optimizer is CEM with the env module's OPT_CFG."""
import importlib

from recovery_rl.dotmap_lite import DotMap


def create_config(env_name, ctrl_type, ctrl_args, overrides, logdir):
    cfg = DotMap()
    cfg_module = importlib.import_module("config." + env_name).CONFIG_MODULE()
    _create_exp_config(cfg.exp_cfg, cfg_module, logdir)
    _create_ctrl_config(cfg.ctrl_cfg, cfg_module, ctrl_type, ctrl_args)
    return cfg


def _create_exp_config(exp_cfg, cfg_module, logdir):
    exp_cfg.sim_cfg.env = cfg_module.ENV
    exp_cfg.sim_cfg.task_hor = cfg_module.TASK_HORIZON
    exp_cfg.exp_cfg.ntrain_iters = cfg_module.NTRAIN_ITERS
    exp_cfg.exp_cfg.nrollouts_per_iter = cfg_module.NROLLOUTS_PER_ITER
    exp_cfg.log_cfg.logdir = logdir


def _create_ctrl_config(ctrl_cfg, cfg_module, ctrl_type, ctrl_args):
    assert ctrl_type == 'MPC'
    ctrl_cfg.env = cfg_module.ENV
    if hasattr(cfg_module, "UPDATE_FNS"):
        ctrl_cfg.update_fns = cfg_module.UPDATE_FNS
    if hasattr(cfg_module, "obs_preproc"):
        ctrl_cfg.prop_cfg.obs_preproc = cfg_module.obs_preproc
    if hasattr(cfg_module, "obs_postproc"):
        ctrl_cfg.prop_cfg.obs_postproc = cfg_module.obs_postproc
    if hasattr(cfg_module, "targ_proc"):
        ctrl_cfg.prop_cfg.targ_proc = cfg_module.targ_proc
    ctrl_cfg.opt_cfg.plan_hor = cfg_module.PLAN_HOR
    ctrl_cfg.opt_cfg.obs_cost_fn = cfg_module.obs_cost_fn
    ctrl_cfg.opt_cfg.ac_cost_fn = cfg_module.ac_cost_fn
    model_init_cfg = ctrl_cfg.prop_cfg.model_init_cfg
    ctrl_args["model-type"] = 'PE'
    model_init_cfg.num_nets = 5
    ctrl_cfg.prop_cfg.model_train_cfg = cfg_module.NN_TRAIN_CFG
    model_init_cfg.model_constructor = cfg_module.nn_constructor
    ctrl_cfg.prop_cfg.mode = "TSinf"
    ctrl_cfg.prop_cfg.npart = 20
    ctrl_cfg.opt_cfg.mode = "CEM"
    ctrl_cfg.opt_cfg.cfg = cfg_module.OPT_CFG[ctrl_cfg.opt_cfg.mode]


def make_bool(arg):
    return not (arg == "False" or arg == "false" or not bool(arg))
