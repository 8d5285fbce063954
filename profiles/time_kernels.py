"""Time the two big kernels of the bench workload alone (CUDA events, L2 flushed): quick iteration helper.
  python profiles/time_kernels.py [--envs 65536] [--tc 1]"""
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
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--tc", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    args = ap.parse_args()
    from recovery_rl import native
    from recovery_rl.engine import VecEngine
    from env.maze import get_offline_data
    torch.manual_seed(1)
    eng = VecEngine("maze", args.envs, batch_size=256, gamma_safe=0.5, eps_safe=0.15, pos_fraction=0.3, seed=1,
                    use_tensor_cores=args.tc)
    eng.init_agent()
    eng.push_offline(get_offline_data(10000, rng=np.random.RandomState(1)))
    eng.pretrain_qrisk(200)
    eng.reset()
    for _ in range(args.steps):
        eng.step()
    torch.cuda.synchronize()
    flush = torch.zeros(64 * 1024 * 1024, device="cuda")

    def timeit(fn, K=20):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for i in range(K):
            flush.add_(1.0)
            ev[i][0].record(); fn(); ev[i][1].record()
        torch.cuda.synchronize()
        t = [a.elapsed_time(b) for a, b in ev]
        return float(np.mean(t)), float(np.min(t))

    snap = eng.snapshot()
    env = timeit(lambda: native.env_step(eng.env_cfg, eng.action_task, eng.action_real, eng.state, eng.ep_steps, eng.ep_return,
                                         eng.counters, recovery=eng.recovery, task_ring=eng.task_ring,
                                         task_capacity=eng.task_cap, cons_ring=eng.cons_ring, cons_flags=eng.cons_flags,
                                         cons_capacity=eng.cons_cap))
    eng.restore(snap)
    act = timeit(lambda: native.agent_act(eng.cfg, eng.arena, eng.n, eng.state, eng.counters, eng.action_task, eng.action_real,
                                          eng.recovery, eng.qrisk, use_recovery=True, start_steps=eng.start_steps, seed=eng.seed))
    step = timeit(lambda: eng.step(), K=10)
    print("env_step_kernel  mean %.1f us  min %.1f us" % (env[0] * 1e3, env[1] * 1e3))
    print("act kernel       mean %.1f us  min %.1f us" % (act[0] * 1e3, act[1] * 1e3))
    print("eager step       mean %.1f us  min %.1f us" % (step[0] * 1e3, step[1] * 1e3))
    print(eng.read_counters())


if __name__ == "__main__":
    main()
