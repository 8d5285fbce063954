"""Run a few vector steps of the bench workload for ncu (profiling range = cudaProfilerStart/Stop).

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python profiles/profile_step.py --steps 2
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:act -c 2 \
      -o gpurun_out/prof_act python profiles/profile_step.py --steps 1
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "recovery-rl_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--tc", type=int, default=0)
    args = ap.parse_args()
    from recovery_rl.engine import VecEngine
    from env.maze import get_offline_data
    torch.manual_seed(1)
    eng = VecEngine("maze", args.envs, batch_size=256, gamma_safe=0.5, eps_safe=0.15, pos_fraction=0.3, seed=1,
                    use_tensor_cores=args.tc)
    eng.init_agent()
    eng.push_offline(get_offline_data(10000, rng=np.random.RandomState(1)))
    eng.pretrain_qrisk(20)
    eng.reset()
    for _ in range(4):
        eng.step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for _ in range(args.steps):
        eng.step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print(eng.read_counters())


if __name__ == "__main__":
    main()
