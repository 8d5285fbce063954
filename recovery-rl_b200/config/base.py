"""Shared body of the per-env config modules of the model-based recovery policy (the reference repeats it in
config/maze.py:104-165, config/navigation1.py:104-160, config/navigation2.py:104-160): CEM constants, pre/post-processing
of the ensemble's inputs and targets, and the model constructor.  Each env module only sets ENV_NAME / TASK_HORIZON /
PLAN_HOR."""
import torch

from env.make_utils import make_env
from recovery_rl.utils import get_required_argument


class PointEnvConfigModule:
    ENV_NAME = None
    TASK_HORIZON = 100
    NTRAIN_ITERS = 100
    NROLLOUTS_PER_ITER = 1
    PLAN_HOR = 5
    MODEL_IN, MODEL_OUT = 4, 2

    def __init__(self):
        self.ENV = make_env(self.ENV_NAME)
        self.ENV.reset()
        self.NN_TRAIN_CFG = {"epochs": 5}
        self.OPT_CFG = {"Random": {"popsize": 2000},
                        "CEM": {"popsize": 400, "num_elites": 40, "max_iters": 5, "alpha": 0.1}}
        self.UPDATE_FNS = []

    @staticmethod
    def obs_postproc(obs, pred):
        return obs + pred

    @staticmethod
    def targ_proc(obs, next_obs):
        return next_obs - obs

    def obs_cost_fn(self, obs):      # stored, never called (MPC.py:406-412 uses the safety critic only)
        pass

    @staticmethod
    def ac_cost_fn(acs):
        return 0.01 * (acs ** 2).sum(dim=1)

    def nn_constructor(self, model_init_cfg):
        from recovery_rl.MPC import PtModel
        ensemble_size = get_required_argument(model_init_cfg, "num_nets", "Must provide ensemble size")
        assert model_init_cfg.get("load_model", False) is False, 'Has yet to support loading model'
        model = PtModel(ensemble_size, self.MODEL_IN, self.MODEL_OUT * 2).to(torch.device("cuda"))
        model.optim = torch.optim.Adam(model.parameters(), lr=0.001)
        return model
