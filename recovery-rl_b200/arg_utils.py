"""Command-line surface of Recovery RL (reference arg_utils.py:8-257): every flag name, type and default is
kept, including the quirks the shipped scripts rely on (argparse prefix abbreviation, `type=bool` flags,
the dash in --env-name), so scripts/navigation1.sh, navigation2.sh and maze.sh run unchanged.

Added knobs (defaults keep the reference behaviour; each can also come from the environment so that the
unmodified scripts can be vectorised):
  --num_envs N      env copies stepped per vector step on each GPU     (RRL_NUM_ENVS, default 1)
  --tensor_cores b  run the acting 256x256 contractions on tcgen05     (RRL_TENSOR_CORES, default 1)
  --log_envs k      env copies whose per-step info is logged at N > 1  (RRL_LOG_ENVS, default 1)
  --checkpoint_every k / --resume path   agent (+ engine) checkpoints in the logdir (recovery_rl/checkpoint.py)
"""
import argparse
import os


def get_args(argv=None):
    parser = argparse.ArgumentParser(description='Recovery RL Arguments')
    parser.add_argument('--env-name', default='maze', help='Gym environment (default: maze)')
    parser.add_argument('--logdir', default='runs', help='exterior log directory')
    parser.add_argument('--logdir_suffix', default='', help='log directory suffix')
    parser.add_argument('--cuda', action='store_true', help='run on CUDA (always true here: there is no CPU path)')
    parser.add_argument('--cnn', action='store_true', help='visual observations (not supported by this build)')

    # SAC
    parser.add_argument('--lr', type=float, default=0.0003)
    parser.add_argument('--updates_per_step', type=int, default=1)
    parser.add_argument('--start_steps', type=int, default=100)
    parser.add_argument('--target_update_interval', type=int, default=1)
    parser.add_argument('--policy', default='Gaussian')
    parser.add_argument('--eval', type=bool, default=True)
    parser.add_argument('--gamma', type=float, default=0.99)
    parser.add_argument('--tau', type=float, default=0.005)
    parser.add_argument('--alpha', type=float, default=0.2)
    parser.add_argument('--automatic_entropy_tuning', type=bool, default=False)
    parser.add_argument('--seed', type=int, default=123456)
    parser.add_argument('--batch_size', type=int, default=256)
    parser.add_argument('--num_steps', type=int, default=1000000)
    parser.add_argument('--num_eps', type=int, default=1000000)
    parser.add_argument('--hidden_size', type=int, default=256)
    parser.add_argument('--replay_size', type=int, default=1000000)
    parser.add_argument('--task_demos', action='store_true')
    parser.add_argument('--num_task_transitions', type=int, default=10000000)
    parser.add_argument('--critic_pretraining_steps', type=int, default=3000)

    # Q_risk / recovery
    parser.add_argument('--pos_fraction', type=float, default=-1)
    parser.add_argument('--gamma_safe', type=float, default=0.5)
    parser.add_argument('--eps_safe', type=float, default=0.1)
    parser.add_argument('--tau_safe', type=float, default=0.0002)
    parser.add_argument('--safe_replay_size', type=int, default=1000000)
    parser.add_argument('--num_unsafe_transitions', type=int, default=10000)
    parser.add_argument('--critic_safe_pretraining_steps', type=int, default=10000)
    parser.add_argument('--use_recovery', action='store_true')
    parser.add_argument('--MF_recovery', action='store_true')
    parser.add_argument('--Q_sampling_recovery', action='store_true')
    parser.add_argument('-ca', '--ctrl_arg', action='append', nargs=2, default=[])
    parser.add_argument('-o', '--override', action='append', nargs=2, default=[])
    parser.add_argument('--recovery_policy_update_freq', type=int, default=1)
    parser.add_argument('--vismpc_recovery', action='store_true')
    parser.add_argument('--load_vismpc', action='store_true')
    parser.add_argument('--model_fname', default='image_maze_dynamics')
    parser.add_argument('--beta', type=float, default=10)

    # ablations / comparisons
    parser.add_argument('--disable_offline_updates', action='store_true')
    parser.add_argument('--disable_online_updates', action='store_true')
    parser.add_argument('--disable_action_relabeling', action='store_true')
    parser.add_argument('--add_both_transitions', action='store_true')
    parser.add_argument('--constraint_reward_penalty', type=float, default=0)
    parser.add_argument('--DGD_constraints', action='store_true')
    parser.add_argument('--use_constraint_sampling', action='store_true')
    parser.add_argument('--nu', type=float, default=0.01)
    parser.add_argument('--update_nu', action='store_true')
    parser.add_argument('--nu_schedule', action='store_true')
    parser.add_argument('--nu_start', type=float, default=1e3)
    parser.add_argument('--nu_end', type=float, default=0)
    parser.add_argument('--RCPO', action='store_true')
    parser.add_argument('--lambda_RCPO', type=float, default=0.01)

    # B200 build
    parser.add_argument('--num_envs', type=int, default=int(os.environ.get('RRL_NUM_ENVS', '1')))
    parser.add_argument('--tensor_cores', type=int, default=int(os.environ.get('RRL_TENSOR_CORES', '1')))
    parser.add_argument('--log_envs', type=int, default=int(os.environ.get('RRL_LOG_ENVS', '1')))
    parser.add_argument('--mpc_popsize', type=int, default=None,
                        help='CEM candidates per env copy for the model-based recovery policy at --num_envs > 1 '
                             '(default: the reference\'s 400; BASELINE config 5 uses 50 x 20 particles)')
    parser.add_argument('--checkpoint_every', type=int, default=0,
                        help='write <logdir>/checkpoint.pt every k episodes (k reports at --num_envs > 1); 0 = off')
    parser.add_argument('--resume', default='', help='checkpoint to load before training')
    return parser.parse_args(argv)
