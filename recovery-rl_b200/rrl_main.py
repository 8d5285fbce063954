"""Command-line entry point of the drop-in surface.

Keeps the calling convention of the reference's rrl_main.py:5-9 (`python rrl_main.py --env-name maze --use_recovery ...`):
the flags are parsed by arg_utils.get_args and handed to recovery_rl.experiment.Experiment, whose run() drives the CUDA
path (librrl.so).  The process fails at start-up if the extension or a GPU is missing; there is no CPU fallback.
"""
import sys


def main(argv=None):
    from recovery_rl import native
    native.require_cuda()
    import arg_utils
    from recovery_rl import experiment as rrl_experiment
    if argv is not None:
        sys.argv = [sys.argv[0]] + list(argv)
    cfg = arg_utils.get_args()
    rrl_experiment.Experiment(cfg).run()
    return 0


if __name__ == '__main__':
    sys.exit(main())
