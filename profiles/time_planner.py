"""Time one planning call of the model-based recovery policy at BASELINE config 5 scale:
Maze, 2,048 env copies x (50 candidates x 20 particles = 1,000 CEM particles), horizon 15, 5 CEM iterations.
  python profiles/time_planner.py [--envs 2048] [--popsize 50]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "recovery-rl_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=2048)
    ap.add_argument("--popsize", type=int, default=50)
    ap.add_argument("--tc", type=int, default=1, help="rollout contractions on tcgen05 (1) or fp32 SIMT (0)")
    args = ap.parse_args()
    from recovery_rl.engine import VecEngine
    from env.maze import get_offline_data
    torch.manual_seed(1)
    np.random.seed(1)
    eng = VecEngine("maze", args.envs, batch_size=256, gamma_safe=0.5, eps_safe=0.15, pos_fraction=0.3, seed=1,
                    mf_recovery=False, mb_recovery=True, mpc_popsize=args.popsize, replay_size=200000, safe_replay_size=200000,
                    use_tensor_cores=args.tc)
    eng.init_agent()
    demos = get_offline_data(4000, rng=np.random.RandomState(1))
    eng.push_offline(demos)
    eng.pretrain_qrisk(50, n_demos=len(demos))
    t0 = time.time()
    eng.train_mb(demos, epochs=5)
    torch.cuda.synchronize()
    print("ensemble training (5 epochs, %d transitions): %.2f s" % (len(demos), time.time() - t0))
    eng.reset()
    mpc = eng.mpc
    for _ in range(2):
        mpc.plan()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record(); mpc.plan(); ev[1].record(); mpc.plan(); ev[2].record()
    torch.cuda.synchronize()
    ms = [ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])]
    rows = args.envs * args.popsize * 20
    flop = rows * 15 * 5 * 2.0 * (4 * 200 + 200 * 200 * 2 + 200 * 4 + 2 * (4 * 256 + 256 * 256 + 256))
    print("tensor cores: %d" % args.tc)
    print("MPC.plan: %d envs x %d particles, %.1f ms per call (%.1f / %.1f), %.1f TFLOP/s algorithmic, %.0f planned env-steps/s"
          % (args.envs, args.popsize * 20, np.mean(ms), ms[0], ms[1], flop / (np.mean(ms) * 1e-3) / 1e12, args.envs / (np.mean(ms) * 1e-3)))
    for _ in range(3):
        eng.step()
    eng.capture()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(3):
        eng.replay()
    ev[1].record()
    torch.cuda.synchronize()
    print("vector step with model-based recovery (graph replay): %.1f ms" % (ev[0].elapsed_time(ev[1]) / 3))
    print(eng.read_counters())


if __name__ == "__main__":
    main()
