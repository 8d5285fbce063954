"""GPU: the whole vector step (VecEngine) against the oracle, eager and as a replayed CUDA graph."""
import importlib.util
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _smoke():
    spec = importlib.util.spec_from_file_location("rrl_smoke_impl", os.path.join(HERE, "smoke_impl.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("env_name", ["navigation1", "navigation2", "maze"])
def test_vector_step_matches_oracle(native, cuda, env_name):
    assert _smoke().run(env_name=env_name, n=384, B=64, steps=4, seed=5, verbose=False)


# BASELINE.json configs 2-4 at their own sizes: every stage of the vector step against the oracle
BASELINE_CASES = {
    "C2_nav1_4096_b256": dict(env_name="navigation1", n=4096, B=256, gamma_safe=0.8, eps_safe=0.3, demos_n=2000),
    "C3_nav2_8192_b1024": dict(env_name="navigation2", n=8192, B=1024, gamma_safe=0.65, eps_safe=0.2, demos_n=4000),
    "C4_maze_shard_8192_b256": dict(env_name="maze", n=8192, B=256, gamma_safe=0.5, eps_safe=0.15, pos_fraction=0.3,
                                    demos_n=2000),
    "C4_maze_65536_b256": dict(env_name="maze", n=65536, B=256, gamma_safe=0.5, eps_safe=0.15, pos_fraction=0.3,
                               demos_n=2000, steps=2),
}


@pytest.mark.parametrize("tensor_cores", [0, 1, 2])
@pytest.mark.parametrize("case", sorted(BASELINE_CASES))
def test_baseline_configs_match_oracle(native, cuda, case, tensor_cores):
    kw = dict(steps=3, seed=11, verbose=False, tensor_cores=tensor_cores)
    kw.update(BASELINE_CASES[case])
    assert _smoke().run(**kw)


def test_graph_replay_equals_eager(native, cuda):
    """the captured CUDA graph replays to exactly the state an eager run reaches (Philox mode, same seed)."""
    from recovery_rl.engine import VecEngine

    def make():
        torch.manual_seed(1)
        e = VecEngine("maze", 2048, batch_size=64, replay_size=16384, safe_replay_size=16384, gamma_safe=0.5,
                      eps_safe=0.15, pos_fraction=0.3, seed=9, start_steps=100)
        e.init_agent()
        from env.maze import get_offline_data
        e.push_offline(get_offline_data(2000, rng=np.random.RandomState(4)))
        e.pretrain_qrisk(5)
        e.reset()
        return e

    a, b = make(), make()
    for _ in range(6):
        a.step()
    b.capture()
    for _ in range(6):
        b.replay()
    torch.cuda.synchronize()
    ca, cb = a.counters.clone(), b.counters.clone()
    ca[native.C_RETURN_SUM_BITS] = cb[native.C_RETURN_SUM_BITS] = 0      # fp64 atomicAdd: order-dependent bits
    assert torch.equal(ca, cb), (ca.tolist(), cb.tolist())
    assert abs(a.read_counters()["return_sum"] - b.read_counters()["return_sum"]) < 1e-6 * (1 + abs(a.read_counters()["return_sum"]))
    assert torch.equal(a.state, b.state)
    assert torch.equal(a.arena[:a.agent.grad_off], b.arena[:b.agent.grad_off])       # all six networks, bit-exact
    assert torch.equal(a.mt_state, b.mt_state)
    c = a.read_counters()
    assert c["sac_updates"] == 5 and c["qrisk_updates"] == 5 + 5 and c["error"] == 0
    assert c["total_numsteps"] == 6 * 2048


@pytest.mark.gpu
def test_peer_grads_match_nccl():
    """world 2: the optimizer-step kernel that sums the ranks' gradient blocks over NVLink peer memory leaves exactly the
    parameters the NCCL all-reduce path leaves (tests/p2p_check.py).  Needs two GPUs."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(here, "p2p_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "P2P_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.gpu
def test_pipelined_host_steps_match_serial():
    """submit/collect (copies on their own streams, two staging slots) returns exactly what step_host returns for the
    same inputs, step by step."""
    import numpy as np
    import torch
    from recovery_rl.engine import VecEngine
    n, B = 1024, 64
    engines = []
    for _ in range(2):
        torch.manual_seed(9)
        e = VecEngine("navigation1", n, batch_size=B, replay_size=8192, safe_replay_size=8192, gamma_safe=0.8, eps_safe=0.3,
                      seed=9, host_inputs=True, start_steps=0)
        e.init_agent()
        e.reset(torch.zeros(2, n, dtype=torch.float64, device=e.device))
        engines.append(e)
    rs = np.random.RandomState(4)

    def draw():
        return dict(reset_draws=rs.randn(2, n), env_noise=rs.randn(2, n), eps_task=rs.randn(n, 2).astype(np.float32),
                    eps_rec=rs.randn(n, 2).astype(np.float32), rand_u=rs.rand(n, 2).astype(np.float32),
                    sac_eps_next=rs.randn(B, 2).astype(np.float32), sac_eps_cur=rs.randn(B, 2).astype(np.float32),
                    qr_eps_next=rs.randn(B, 2).astype(np.float32), qr_eps_rec=rs.randn(B, 2).astype(np.float32))
    inputs = [draw() for _ in range(7)]
    a, b = engines
    for e in engines:
        e.step_host(inputs[0])
        e.capture()
    want = [{k: v.clone() for k, v in a.step_host(x).items()} for x in inputs[1:]]
    b.enable_pipeline()
    got = []
    for i, x in enumerate(inputs[1:]):
        t = b.submit(x)
        if i >= 1:
            got.append({k: v.clone() for k, v in b.collect(t - 1).items()})
    got.append({k: v.clone() for k, v in b.collect(len(inputs) - 2).items()})
    assert len(got) == len(want)
    from recovery_rl import native
    for i, (g, w) in enumerate(zip(got, want)):
        for k in w:
            a, b = g[k].clone(), w[k].clone()
            if k == "counters":          # the return sum is an fp64 atomicAdd over the env copies: order-dependent last bits
                ra, rb = a[native.C_RETURN_SUM_BITS:native.C_RETURN_SUM_BITS + 1].view(torch.float64), \
                    b[native.C_RETURN_SUM_BITS:native.C_RETURN_SUM_BITS + 1].view(torch.float64)
                assert abs(float(ra) - float(rb)) <= 1e-9 * (1 + abs(float(rb))), (i, float(ra), float(rb))
                a[native.C_RETURN_SUM_BITS] = b[native.C_RETURN_SUM_BITS] = 0
            assert torch.equal(a, b), (i, k)


@pytest.mark.gpu
def test_eval_rollout_is_side_effect_free_and_uses_the_mean_action():
    """VecEngine.eval_rollout (experiment.py:493-538): eval-mode task action = the policy's mean action, same Q_risk
    threshold; episodes chain and end at done / the horizon; the training state is untouched."""
    import numpy as np
    import torch
    from oracle import envs as oenvs
    from oracle.agent import Agent
    from recovery_rl import native
    from recovery_rl.engine import VecEngine
    torch.manual_seed(3)
    np.random.seed(3)
    ora = Agent(action_scale=(np.float32(1.0),) * 2, gamma_safe=0.8, eps_safe=0.3)
    eng = VecEngine("navigation1", 256, batch_size=64, replay_size=8192, safe_replay_size=8192, gamma_safe=0.8, eps_safe=0.3,
                    seed=3, start_steps=0, use_tensor_cores=2)
    eng.init_agent(ora.nets())
    eng.push_offline(oenvs.nav_offline_data(oenvs.KIND_BY_NAME["navigation1"], 400))
    eng.pretrain_qrisk(3, n_demos=400)
    eng.reset()
    for _ in range(3):
        eng.step()
    torch.cuda.synchronize()
    P = {net: eng.agent.params(net) for net in native.NET_NAMES}
    ora.load(lambda net, i: P[net][i])
    before = dict(arena=eng.arena.clone(), counters=eng.counters.clone(), state=eng.state.clone(), mt=eng.mt_state.clone(),
                  ep_steps=eng.ep_steps.clone())
    eps = eng.eval_rollout(16)
    torch.cuda.synchronize()
    # bitwise (the arena also holds fp16 operand images, whose bit patterns read as float32 may be NaN)
    assert torch.equal(eng.arena.view(torch.int32), before["arena"].view(torch.int32)) and torch.equal(eng.counters, before["counters"])
    assert torch.equal(eng.state, before["state"]) and torch.equal(eng.mt_state, before["mt"])
    assert torch.equal(eng.ep_steps, before["ep_steps"])
    assert len(eps) == 16
    s0 = np.array([e[0]["state"] for e in eps])
    a_task, a_real, rec, qv = ora.act(s0, np.zeros((16, 2), np.float32), np.zeros((16, 2), np.float32), eps_safe=0.3)
    got_rec = np.array([e[0]["recovery"] for e in eps])
    sure = np.abs(qv - 0.3) > 1e-4
    assert np.array_equal(got_rec[sure], rec[sure])
    plain = sure & ~rec                                  # no recovery: executed action == the policy's mean action
    got_a = np.array([e[0]["action"] for e in eps])
    assert np.allclose(got_a[plain], a_task[plain], rtol=1e-4, atol=1e-5)
    for e in eps:
        assert 1 <= len(e) <= 100
        for a, b in zip(e, e[1:]):
            assert np.array_equal(a["next_state"], b["state"])          # the episode chains (no reset inside it)
        assert all(k in e[0] for k in ("constraint", "reward", "state", "next_state", "action", "success", "recovery"))
    eps2 = eng.eval_rollout(16)
    assert len(eps2) == 16 and not np.array_equal(np.array([e[0]["state"] for e in eps2]), s0)    # fresh reset draws


@pytest.mark.gpu
def test_staged_acting_equals_fused_launch():
    """the three acting stages enqueued next to the updates (policy after the SAC step, Q_risk after the safety-critic step,
    recovery + select after the recovery step) leave exactly the state the single fused acting launch leaves."""
    import numpy as np
    import torch
    from recovery_rl import native
    from recovery_rl.engine import VecEngine
    from env.maze import get_offline_data

    def make(staging):
        torch.manual_seed(1)
        e = VecEngine("maze", 4096, batch_size=64, replay_size=32768, safe_replay_size=32768, gamma_safe=0.5, eps_safe=0.15,
                      pos_fraction=0.3, seed=9, start_steps=4096 * 2, use_tensor_cores=2)      # two steps of the random phase
        e.act_staging = staging
        e.init_agent()
        e.push_offline(get_offline_data(2000, rng=np.random.RandomState(4)))
        e.pretrain_qrisk(5)
        e.reset()
        return e

    a, b = make(False), make(True)
    assert b.staged_act and not a.staged_act
    for _ in range(3):
        a.step(); b.step()
    b.capture()
    a.capture()
    for _ in range(4):
        a.replay(); b.replay()
    torch.cuda.synchronize()
    for name in ("action_task", "action_real", "recovery", "qrisk", "state", "ep_steps", "mt_state"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    ca, cb = a.counters.clone(), b.counters.clone()
    ca[native.C_RETURN_SUM_BITS] = cb[native.C_RETURN_SUM_BITS] = 0      # fp64 atomicAdd: order-dependent bits
    assert torch.equal(ca, cb)
    g = a.agent.grad_off
    assert torch.equal(a.arena[:g].view(torch.int32), b.arena[:g].view(torch.int32))     # all six networks, bit-exact
    c = b.read_counters()
    assert c["error"] == 0 and c["sac_updates"] >= 6 and c["qrisk_updates"] >= 6 + 5
    assert int(b.recovery.sum()) > 0                                      # the recovery branch was exercised
