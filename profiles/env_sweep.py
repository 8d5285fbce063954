"""N-sweep of the env-step + replay-push kernel (the HBM fraction is structurally small at the BASELINE
N because one launch moves only 120 B x N; the sweep shows the asymptote).  CUDA events, L2 flushed between launches.
  python profiles/env_sweep.py [--env-name navigation1] [--max-log2 24] > gpurun_out/env_sweep.txt
Algorithmic bytes per env-step: 120 (production RNG)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "recovery-rl_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

BYTES_PER_ENV_STEP = 120.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--env-name", default="navigation1")
    ap.add_argument("--min-log2", type=int, default=16)
    ap.add_argument("--max-log2", type=int, default=24)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    from recovery_rl import native
    from recovery_rl.engine import VecEngine
    peak = 6555.5
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f).get("hbm_gbs", peak))
    except Exception:
        pass
    flush = torch.zeros(64 * 1024 * 1024, device="cuda")
    print("env %s, %d B/env-step algorithmic, HBM peak %.1f GB/s" % (args.env_name, BYTES_PER_ENV_STEP, peak))
    print("%10s %12s %12s %10s %14s" % ("N", "us/launch", "GB/s", "frac", "env-steps/s"))
    for lg in range(args.min_log2, args.max_log2 + 1):
        n = 1 << lg
        torch.manual_seed(1)
        eng = VecEngine(args.env_name, n, batch_size=256, replay_size=2 * n, safe_replay_size=2 * n, seed=1, start_steps=0,
                        gamma_safe=0.8, eps_safe=0.3)
        eng.init_agent()
        eng.reset()
        eng.action_task.uniform_(-0.05, 0.05)
        eng.action_real.copy_(eng.action_task)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.reps)]
        for i in range(args.reps + 3):
            flush.add_(1.0)
            k = max(0, i - 3)
            if i >= 3:
                ev[k][0].record()
            native.env_step(eng.env_cfg, eng.action_task, eng.action_real, eng.state, eng.ep_steps, eng.ep_return,
                            eng.counters, recovery=eng.recovery, task_ring=eng.task_ring, task_capacity=eng.task_cap,
                            cons_ring=eng.cons_ring, cons_flags=eng.cons_flags, cons_capacity=eng.cons_cap)
            if i >= 3:
                ev[k][1].record()
        torch.cuda.synchronize()
        t = float(np.median([a.elapsed_time(b) for a, b in ev])) * 1e-3
        gbs = BYTES_PER_ENV_STEP * n / t * 1e-9
        print("%10d %12.1f %12.1f %10.4f %14.3e" % (n, t * 1e6, gbs, gbs / peak, n / t))
        del eng
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
