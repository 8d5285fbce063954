"""Whole N = 1 runs of the reference's Experiment for the comparison algorithms of scripts/navigation1.sh
(unconstrained, LR, RSPO, SQRL, RP, RCPO), shortened (6 episodes, 2,000 offline transitions, 30 pre-training updates,
batch 16) -> tests/golden/runs_nav1.npz.  The runs use the LIVE generators (numpy, torch, Box, CPython random), all
seeded by --seed as experiment.py:88-93 does, so a drop-in that consumes the same streams reproduces them.
Build container only (imports /root/reference through the harness).  TEST INFRASTRUCTURE."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.ref_harness import harness  # noqa: E402

COMMON = ["--env-name", "navigation1", "--num_eps", "6", "--num_unsafe_transitions", "2000",
          "--critic_safe_pretraining_steps", "30", "--batch_size", "16", "--logdir", "/tmp/rrl_golden_runs"]
RUNS = {   # tag -> (seed, the algorithm flags of the script line)
    "unconstrained": (2, []),
    "lr": (3, ["--gamma_safe", "0.8", "--eps_safe", "0.3", "--DGD_constraints", "--nu", "5000", "--update_nu"]),
    "rspo": (4, ["--gamma_safe", "0.8", "--eps_safe", "0.3", "--DGD_constraints", "--nu_schedule", "--nu_start", "10000"]),
    "sqrl": (5, ["--gamma_safe", "0.8", "--eps_safe", "0.3", "--DGD_constraints", "--use_constraint_sampling", "--nu", "5000",
                 "--update_nu"]),
    "rp": (6, ["--constraint_reward_penalty", "1000"]),
    "rcpo": (7, ["--gamma_safe", "0.8", "--eps_safe", "0.3", "--RCPO", "--lambda", "1000"]),
}
# recovery branches that no script line uses (experiment.py:446-448, qrisk.py:214-225) -> tests/golden/runs_nav1_extra.npz;
# eps_safe low enough that the barely trained safety critic triggers recoveries within 6 episodes
EXTRA_RUNS = {
    "addboth": (8, ["--use_recovery", "--MF_recovery", "--gamma_safe", "0.8", "--eps_safe", "0.05", "--add_both_transitions"]),
    "qsample": (9, ["--use_recovery", "--Q_sampling_recovery", "--gamma_safe", "0.8", "--eps_safe", "0.05"]),
    "det": (10, ["--policy", "Deterministic", "--use_recovery", "--MF_recovery", "--gamma_safe", "0.8", "--eps_safe", "0.3",
                 "--start_steps", "20"]),
}
STRIDE = 37


def one_run(tag, seed, flags):
    argv = COMMON + flags + ["--seed", str(seed), "--logdir_suffix", tag]
    so = sys.stdout
    sys.stdout = open(os.devnull, "w")
    try:
        exp = harness.make_experiment(argv)
        cfg = exp.exp_cfg
        if not cfg.disable_offline_updates and (cfg.use_recovery or cfg.DGD_constraints or cfg.RCPO):
            exp.pretrain_critic_recovery()          # experiment.py:357-361
        infos, ep_len = [], []
        for ep in range(1, 7):
            info = exp.get_train_rollout(ep)
            infos += info
            ep_len.append(len(info))
    finally:
        sys.stdout = so
    P = tag + "_"
    out = {P + "argv": np.array(argv), P + "ep_len": np.array(ep_len, np.int64),
           P + "state": np.array([i["state"] for i in infos]),
           P + "action": np.array([i["action"] for i in infos], np.float32),
           P + "reward": np.array([i["reward"] for i in infos]),
           P + "constraint": np.array([int(i["constraint"]) for i in infos], np.uint8),
           P + "recovery": np.array([bool(i.get("recovery", False)) for i in infos], np.uint8),
           P + "num_viols": np.int64(exp.num_viols), P + "num_successes": np.int64(exp.num_successes),
           P + "total_numsteps": np.int64(exp.total_numsteps), P + "updates": np.int64(exp.updates),
           P + "memory_len": np.int64(len(exp.memory)), P + "recovery_memory_len": np.int64(len(exp.recovery_memory))}
    nets = {"critic": exp.agent.critic, "policy": exp.agent.policy, "qrisk": exp.agent.safety_critic.safety_critic}
    for name, mod in nets.items():
        for k, p in enumerate(mod.parameters()):
            out["%sfinal_%s_%d" % (P, name, k)] = p.detach().numpy().ravel()[::STRIDE].copy()
    for name in ("nu", "lambda_RCPO", "alpha"):
        v = getattr(exp.agent, name, None)
        if v is not None:
            out[P + name] = np.float64(float(v))
    print("%-14s seed %d: %d steps, ep_len %s, viols %d, successes %d, updates %d" % (
        tag, seed, len(infos), ep_len, exp.num_viols, exp.num_successes, exp.updates))
    return out


def main_extra():
    out = {"stride": np.int64(STRIDE), "tags": np.array(sorted(EXTRA_RUNS))}
    for tag in sorted(EXTRA_RUNS):
        seed, flags = EXTRA_RUNS[tag]
        out.update(one_run(tag, seed, flags))
    path = os.path.join(ROOT, "tests", "golden", "runs_nav1_extra.npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024.0))


def main():
    out = {"stride": np.int64(STRIDE), "tags": np.array(sorted(RUNS))}
    for tag in sorted(RUNS):
        seed, flags = RUNS[tag]
        out.update(one_run(tag, seed, flags))
    path = os.path.join(ROOT, "tests", "golden", "runs_nav1.npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024.0))


if __name__ == "__main__":
    if "--extra" in sys.argv:
        main_extra()           # leaves runs_nav1.npz untouched
    else:
        main()
        main_extra()
