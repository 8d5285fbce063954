"""GPU: the drop-in surface (rrl_main / arg_utils / Experiment / SAC / QRiskWrapper / ReplayMemory / env classes)
against the reference: offline-data generators bit-exact, replay API, and the reference's own 12-episode
Navigation1 run (seed 7) reproduced through `Experiment` with live RNGs."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["nav1", "nav2"])
def test_offline_data_generators_bit_exact(native, cuda, golden_dir, name):
    import importlib
    mod = importlib.import_module("env.navigation%s" % name[-1])
    z = np.load(os.path.join(golden_dir, "offline_%s.npz" % name))
    np.random.seed(int(z["seed"]))
    tr = mod.get_offline_data(int(z["num"]))
    assert len(tr) == len(z["state"])
    assert np.array_equal(np.array([t[0] for t in tr]), z["state"])
    assert np.array_equal(np.array([t[1] for t in tr]), z["action"])
    assert np.array_equal(np.array([float(t[2]) for t in tr]), z["constraint"])
    assert np.array_equal(np.array([t[3] for t in tr]), z["next_state"])


def test_env_classes_follow_gym_contract(native, cuda):
    from env.make_utils import make_env
    from oracle import envs as oenvs
    for name in ("navigation1", "navigation2", "maze"):
        env = make_env(name)
        np.random.seed(3)
        s = env.reset()
        assert s.shape == (2,) and env._max_episode_steps == 100 and env.action_space.shape == (2,)
        env.action_space.seed(3)
        for _ in range(5):
            a = env.action_space.sample()
            st = np.random.get_state()
            s2, r, done, info = env.step(a)
            assert set(info) >= {"constraint", "reward", "state", "next_state", "action", "success"}
            if name == "maze":
                ns, rr, d, c, su = oenvs.maze_step_scalar(s, a, env.steps - 1)
            else:
                np.random.set_state(st)
                ns, rr, d, c, su = oenvs.nav_step(oenvs.KIND_BY_NAME[name], s[None], a[None], np.random.randn(1, 2))
                ns, rr, d, c, su = ns[0], rr[0], d[0], c[0], su[0]
            assert np.array_equal(s2, ns) and r == rr and bool(done) == bool(d) and info["constraint"] == int(c)
            s = s2


def test_replay_memory_api_matches_cpython(native, cuda):
    from recovery_rl.replay_memory import ReplayMemory, ConstraintReplayMemory
    mem = ReplayMemory(1000, 5)
    cmem = ConstraintReplayMemory(1000, 5)
    random.seed(5)
    rs = np.random.RandomState(0)
    for i in range(300):
        s = np.array([float(i), 0.5])
        mem.push(s, np.zeros(2, np.float32), -1.0, s + 1, 1.0)
        cmem.push(s, np.zeros(2, np.float32), float(rs.rand() < 0.3), s + 1, 0.0)
    assert len(mem) == 300 and len(cmem) == 300
    st, a, r, s2, m = mem.sample(64)
    assert np.array_equal(st[:, 0].astype(int), random.sample(range(300), 64))
    assert st.shape == (64, 2) and r.shape == (64,) and np.all(s2 == st + 1)
    st, a, c, s2, m = cmem.sample(64, pos_fraction=0.25)
    pos = np.flatnonzero(cmem.pos_idx)
    neg = np.flatnonzero(1 - cmem.pos_idx[:300])
    ref = [pos[j] for j in random.sample(range(len(pos)), 16)] + [neg[j] for j in random.sample(range(len(neg)), 48)]
    assert np.array_equal(st[:, 0].astype(int), ref)
    assert c[:16].all() and not c[16:].any()
    with pytest.raises(ValueError):
        mem.sample(301)


# free-running runs are compared on the prefix up to the first borderline decision (see the tests); the prefix must at least
# cover the first 100 steps (the random-action phase, start_steps = 100, during which updates already run) of a run
MIN_FREE_RUNNING_PREFIX = 100


@pytest.mark.parametrize("fname,n_eps", [("traj_nav1_seed7.npz", 12), ("traj_nav2_seed3.npz", 8)])
def test_experiment_reproduces_reference_run(native, cuda, golden_dir, tmp_path, fname, n_eps):
    """scripts/navigation1.sh-style command through the drop-in Experiment with LIVE RNGs (numpy, torch, Box,
    CPython-compatible sampler) == the reference's own run at seed 7: same episode lengths, constraint and
    recovery flags; states to fp32 round-off of the recovery actions."""
    import arg_utils
    from recovery_rl.experiment import Experiment
    z = np.load(os.path.join(golden_dir, fname))
    argv = [str(x) for x in z["argv"]]
    argv[argv.index("--logdir") + 1] = str(tmp_path)
    args = arg_utils.get_args(argv + ["--tensor_cores", "0"])
    assert args.num_envs == 1
    exp = Experiment(args)
    off = exp.constraint_demo_data
    assert np.array_equal(np.array([t[0] for t in off]), z["offline_state"])
    assert np.array_equal(np.array([t[3] for t in off]), z["offline_next_state"])
    exp.pretrain_critic_recovery()
    infos, ep_len = [], []
    for ep in range(1, n_eps + 1):
        info = exp.get_train_rollout(ep)
        infos += info
        ep_len.append(len(info))
    rec = np.array([bool(i["recovery"]) for i in infos])
    con = np.array([i["constraint"] for i in infos])
    ref_rec, ref_con = z["recovery"].astype(bool), z["constraint"]
    n = min(len(rec), len(ref_rec))
    st_all = np.array([i["state"] for i in infos]); ac_all = np.array([i["action"] for i in infos])
    drift = (np.abs(st_all[:n] - z["state"][:n]).max(1) > 1e-4) | (np.abs(ac_all[:n] - z["action"][:n]).max(1) > 1e-4)
    bad = np.flatnonzero((rec[:n] != ref_rec[:n]) | (con[:n] != ref_con[:n]) | drift)
    first = int(bad[0]) if len(bad) else n
    if fname == "traj_nav1_seed7.npz":
        # the whole 12-episode run is reproduced decision for decision
        assert first == len(ref_rec) == len(rec), (first, len(rec), len(ref_rec))
    else:
        # A FREE-RUNNING run follows the reference only until the first decision whose margin is below the fp32
        # difference between two implementations (here: `Q_risk > eps_safe`, experiment.py:555, after the 10,000
        # pre-training updates of scripts/navigation2.sh the safety critic sits within ~1e-6 of eps_safe on some states);
        # from there both runs are valid but different trajectories, and WHERE that happens moves with every change of a
        # summation order: measured on this run at step 790, 377 and 21 of 800 with three reduction orders of the same
        # kernels.  So this run is a smoke test of the Navigation2 script line: offline data bit-exact (above), identical
        # decisions and states / actions within 1e-4 on whatever prefix precedes the first borderline decision (or the
        # first step at which the accumulated fp32 differences of the updates move a state by more than 1e-4).  Update arithmetic is held to
        # 1e-4 per update by the teacher-forced tests in test_agent_gpu.py / test_algos_gpu.py; whole-run identity is held
        # by the Navigation1 seed-7 run (12 episodes) above and the four comparison runs that match in full below.
        print("traj_nav2_seed3: identical decisions, states within 1e-4 for %d of %d steps" % (first, len(ref_rec)))
        assert first >= 16, (first, len(ref_rec))
    assert np.allclose(np.array([i["state"] for i in infos[:first]]), z["state"][:first], rtol=0, atol=1e-4)
    assert np.allclose(np.array([i["action"] for i in infos[:first]]), z["action"][:first], rtol=0, atol=1e-4)
    if first == len(ref_rec):
        assert ep_len == list(z["ep_len"])
        assert exp.num_viols == int(z["num_viols"]) and exp.total_numsteps == int(z["total_numsteps"])
        assert exp.updates == int(z["updates"])
        stride = int(z["stride"])
        for net in ("critic", "policy", "qrisk", "recovery"):
            for i, p in enumerate(exp.agent.arena.params(net)):
                ref = z["final_%s_%d" % (net, i)]
                assert np.allclose(p.ravel()[::stride], ref, rtol=0, atol=2e-3), (net, i, np.abs(p.ravel()[::stride] - ref).max())


def test_vectorised_experiment_runs(native, cuda, tmp_path):
    import arg_utils
    from recovery_rl.experiment import Experiment
    args = arg_utils.get_args(["--env-name", "maze", "--use_recovery", "--MF_recovery", "--gamma_safe", "0.5",
                               "--eps_safe", "0.15", "--pos_fraction", "0.3", "--num_unsafe_transitions", "2000",
                               "--critic_safe_pretraining_steps", "20", "--batch_size", "64", "--num_envs", "1024",
                               "--num_steps", "60000", "--seed", "3", "--logdir", str(tmp_path), "--replay_size", "100000",
                               "--safe_replay_size", "100000"])
    exp = Experiment(args)
    stats = exp.run()
    assert stats[-1]["total_numsteps"] > 60000 and stats[-1]["error"] == 0
    assert stats[-1]["sac_updates"] > 40 and stats[-1]["qrisk_updates"] > 40
    assert os.path.exists(os.path.join(exp.logdir, "run_stats.pkl")) and os.path.exists(os.path.join(exp.logdir, "args.pkl"))


# the algorithm lines of the reference's scripts/navigation1.sh / maze.sh (flags verbatim, incl. the `--lambda`
# prefix abbreviation of --lambda_RCPO)
SCRIPT_LINES = {
    "RRL_MF": ["--use_recovery", "--MF_recovery"],
    "RRL_MB": ["--use_recovery", "--recovery_policy_update_freq", "2"],
    "unconstrained": [],
    "LR": ["--DGD_constraints", "--nu", "5000", "--update_nu"],
    "RSPO": ["--DGD_constraints", "--nu_schedule", "--nu_start", "10000"],
    "SQRL": ["--DGD_constraints", "--use_constraint_sampling", "--nu", "5000", "--update_nu"],
    "RP": ["--constraint_reward_penalty", "1000"],
    "RCPO": ["--RCPO", "--lambda", "1000"],
}


@pytest.mark.parametrize("algo", sorted(SCRIPT_LINES))
@pytest.mark.parametrize("env_name", ["navigation1", "maze"])
def test_script_algorithm_lines_run(native, cuda, tmp_path, algo, env_name):
    """every algorithm of the shipped scripts runs through rrl_main's surface on the CUDA path (N = 1 loop)."""
    import arg_utils
    from recovery_rl.experiment import Experiment
    extra = ["--gamma_safe", "0.8", "--eps_safe", "0.3"] if env_name == "navigation1" else \
        ["--gamma_safe", "0.5", "--eps_safe", "0.15", "--pos_fraction=0.3"]
    argv = ["--cuda", "--env-name", env_name] + SCRIPT_LINES[algo] + (extra if algo not in ("unconstrained", "RP") else []) + \
        ["--logdir", str(tmp_path), "--logdir_suffix", algo, "--num_eps", "3", "--num_unsafe_transitions", "600",
         "--critic_safe_pretraining_steps", "10", "--batch_size", "16", "--seed", "2"]
    exp = Experiment(arg_utils.get_args(argv))
    exp.run()
    assert exp.total_numsteps > 3 and exp.updates > 0
    l = exp.agent._losses[:3].cpu().numpy()
    assert np.isfinite(l).all()
    if algo in ("LR", "SQRL"):
        assert exp.agent.arena.counters[native.C_ADAM_T_NU].item() == exp.updates
    if algo == "RCPO":
        assert exp.agent.arena.counters[native.C_ADAM_T_LAMBDA].item() == exp.updates and exp.agent.lambda_RCPO != 1000
    if algo == "RRL_MB":        # PETS ensemble trained on the demos + online episodes, CEM planner used for recovery
        rp = exp.recovery_policy
        assert rp.has_been_trained and len(rp.train_in) > 400 and np.isfinite(rp.last_train_loss)
        assert torch.isfinite(rp.dyn_image).all()


# flags no script line uses, through the vector engine as well (device RNG + CUDA graph)
EXTRA_LINES = {
    "DET": ["--policy", "Deterministic", "--use_recovery", "--MF_recovery"],
    "QSAMPLE": ["--use_recovery", "--Q_sampling_recovery"],
    "ADDBOTH": ["--use_recovery", "--MF_recovery", "--add_both_transitions"],
}


@pytest.mark.parametrize("algo", ["LR", "RSPO", "SQRL", "RCPO", "RP", "unconstrained", "DET", "QSAMPLE", "ADDBOTH"])
def test_vectorised_comparison_algorithms_run(native, cuda, tmp_path, algo):
    import arg_utils
    from recovery_rl.experiment import Experiment
    argv = ["--env-name", "navigation1", "--gamma_safe", "0.8", "--eps_safe", "0.3"] + dict(SCRIPT_LINES, **EXTRA_LINES)[algo] + \
        ["--num_unsafe_transitions", "2000", "--critic_safe_pretraining_steps", "20", "--batch_size", "64",
         "--num_envs", "512", "--num_steps", "30000", "--seed", "3", "--logdir", str(tmp_path), "--replay_size", "60000",
         "--safe_replay_size", "60000"]
    exp = Experiment(arg_utils.get_args(argv))
    stats = exp.run()
    assert stats[-1]["total_numsteps"] > 30000 and stats[-1]["error"] == 0 and stats[-1]["sac_updates"] > 20
    uses_qrisk = algo in ("LR", "RSPO", "SQRL", "RCPO", "DET", "QSAMPLE", "ADDBOTH")
    assert (stats[-1]["qrisk_updates"] > 20) == uses_qrisk
    assert torch.isfinite(exp.engine.arena[:exp.engine.agent.grad_off]).all()


@pytest.mark.parametrize("tag", ["unconstrained", "lr", "rspo", "sqrl", "rp", "rcpo", "addboth", "qsample", "det"])
def test_experiment_reproduces_reference_comparison_runs(native, cuda, golden_dir, tmp_path, tag):
    """the comparison-algorithm lines of scripts/navigation1.sh through the drop-in Experiment with LIVE RNGs == the
    reference's own (shortened) runs recorded by oracle/ref_harness/make_golden_runs.py; `addboth` / `qsample` / `det`: the
    branches no script line uses (--add_both_transitions, experiment.py:446-448; --Q_sampling_recovery, qrisk.py:214-225;
    --policy Deterministic, model.py:447-485), recorded the same way into runs_nav1_extra.npz."""
    import arg_utils
    from recovery_rl.experiment import Experiment
    z = np.load(os.path.join(golden_dir, "runs_nav1_extra.npz" if tag in ("addboth", "qsample", "det") else "runs_nav1.npz"))
    P = tag + "_"
    argv = [str(x) for x in z[P + "argv"]]
    argv[argv.index("--logdir") + 1] = str(tmp_path)
    args = arg_utils.get_args(argv + ["--tensor_cores", "0"])
    exp = Experiment(args)
    if not args.disable_offline_updates and (args.use_recovery or args.DGD_constraints or args.RCPO):
        exp.pretrain_critic_recovery()
    infos, ep_len = [], []
    for ep in range(1, len(z[P + "ep_len"]) + 1):
        info = exp.get_train_rollout(ep)
        infos += info
        ep_len.append(len(info))
    # free-running runs (see test_experiment_reproduces_reference_run): compared on the prefix up to the first step at
    # which the states drift past 1e-4 or a flag differs; that prefix must cover the random-action phase and the first
    # 100 updates, and a run that never drifts must also end with the reference's counters
    con = np.array([int(i["constraint"]) for i in infos])
    st = np.array([i["state"] for i in infos])
    ac = np.array([i["action"] for i in infos])
    n = min(len(con), len(z[P + "constraint"]))
    bad = (con[:n] != z[P + "constraint"][:n]) | (np.abs(st[:n] - z[P + "state"][:n]).max(1) > 1e-4) | \
        (np.abs(ac[:n] - z[P + "action"][:n]).max(1) > 1e-4)
    first = int(np.flatnonzero(bad)[0]) if bad.any() else n
    print("%s: within 1e-4 of the reference run for %d of %d steps" % (tag, first, len(z[P + "constraint"])))
    # Q-sampling picks the arg-min of 1,000 candidates every step: two candidates closer than the fp32 ordering noise of the two
    # implementations (the golden of tests/test_select_gpu.py has gaps down to 3e-6) may swap and end the common prefix early
    need = 5 if tag == "qsample" else min(MIN_FREE_RUNNING_PREFIX, len(z[P + "constraint"]))
    assert first >= need, (tag, first)
    if first == len(z[P + "constraint"]) == len(con):
        assert ep_len == list(z[P + "ep_len"])
        assert exp.num_viols == int(z[P + "num_viols"]) and exp.total_numsteps == int(z[P + "total_numsteps"])
        assert exp.updates == int(z[P + "updates"])
        if P + "memory_len" in z.files:       # add_both_transitions: two task-buffer pushes per recovery step
            assert len(exp.memory) == int(z[P + "memory_len"]) and len(exp.recovery_memory) == int(z[P + "recovery_memory_len"])
        if tag in ("addboth", "qsample"):
            rec = np.array([bool(i.get("recovery", False)) for i in infos])
            assert np.array_equal(rec, z[P + "recovery"].astype(bool)) and rec.any()


def test_utils_soft_and_hard_update_delegates(native, cuda):
    """recovery_rl.utils.soft_update / hard_update with the reference's signature (utils.py:46-54) on the device nets."""
    from recovery_rl import utils
    from recovery_rl.arena import AgentArena
    from recovery_rl.model import build_reference_modules
    from recovery_rl.sac import NetHandle
    torch.manual_seed(0)
    ar = AgentArena(cuda, max_batch=64)
    ar.load_modules(build_reference_modules())
    src, tgt = NetHandle(ar, "critic"), NetHandle(ar, "critic_target")
    utils.hard_update(tgt, src)
    for p, q in zip(src.parameters(), tgt.parameters()):
        assert torch.equal(p, q)
    for p in src.parameters():
        p.add_(0.25)
    before = [q.clone() for q in tgt.parameters()]
    utils.soft_update(tgt, src, 0.1)
    torch.cuda.synchronize()
    for p, q0, q in zip(src.parameters(), before, tgt.parameters()):
        assert torch.allclose(q, q0 * (1.0 - 0.1) + p * 0.1, rtol=0, atol=1e-7)
