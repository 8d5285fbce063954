"""Entry point, same contract as the reference's rrl_main.py:5-9:  python -m rrl_main --env-name ... """
from arg_utils import get_args
from recovery_rl.experiment import Experiment

if __name__ == '__main__':
    exp_cfg = get_args()
    experiment = Experiment(exp_cfg)
    experiment.run()
