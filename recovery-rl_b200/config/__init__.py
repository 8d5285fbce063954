"""`config/` of the reference: the DotMap trees of the PETS / CEM model-based recovery policy
(config/default.py:15-119, config/utils.py, config/<env>.py).  Same module names, `create_config` entry point and
per-env constants; the ensemble itself (PtModel) lives in recovery_rl/MPC.py and runs through csrc/mpc.cu."""
from .default import create_config  # noqa: F401

PLAN_HOR = {"maze": 15, "navigation1": 5, "navigation2": 5}       # config/maze.py:110, config/navigation1.py:110
CEM = dict(popsize=400, num_elites=40, max_iters=5, alpha=0.1)    # config/maze.py:122-127
ENSEMBLE = dict(num_nets=5, npart=20, prop_mode="TSinf")           # config/default.py:91,108-109
