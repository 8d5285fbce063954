"""`config/` of the reference holds the DotMap trees of the PETS/CEM model-based recovery policy
(config/default.py, config/<env>.py).  That controller is a 'next' row of the scope table;
the package stays importable and names the per-env constants the reference defines there."""
PLAN_HOR = {"maze": 15, "navigation1": 5, "navigation2": 5}       # config/maze.py:110, config/navigation1.py
CEM = dict(popsize=400, num_elites=40, max_iters=5, alpha=0.1)    # config/maze.py:122-127
ENSEMBLE = dict(num_nets=5, npart=20, prop_mode="TSinf")           # config/default.py:91,108-109


def create_config(env_name, ctrl_type, ctrl_args, overrides, logdir):
    raise NotImplementedError("PETS/CEM recovery (config tree of the reference) is not part of this build; "
                              "use --MF_recovery")
