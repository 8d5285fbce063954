"""GPU parity of the model-based recovery policy (BASELINE config 5: PETS ensemble + CEM planner, csrc/mpc.cu and
recovery_rl/MPC.py) against golden vectors recorded from the reference's own PtModel / MPC / CEMOptimizer
(tests/golden/mpc.npz).  Planner costs and actions within 1e-4 (fp32 path); ensemble init bit-exact."""
import os

import numpy as np
import pytest
import torch

from test_oracle import mpc_noise, mpc_oracle

pytestmark = pytest.mark.gpu


def _product_mpc(z, tag, n_envs=1):
    from config import create_config
    from recovery_rl.dotmap_lite import DotMap
    from recovery_rl.MPC import MPC
    P = tag + "_"
    env_name = "maze" if tag == "maze" else "navigation1"
    cfg = create_config(env_name, "MPC", DotMap(), [], "/tmp/none")
    cc = cfg.ctrl_cfg
    assert cc.opt_cfg.plan_hor == int(z[P + "plan_hor"]) and cc.prop_cfg.npart == 20 and cc.opt_cfg.cfg["popsize"] == 400
    cc.opt_cfg.cfg = {"popsize": int(z[P + "popsize"]), "num_elites": int(z[P + "num_elites"]), "max_iters": 5, "alpha": 0.1}
    np.random.seed(int(z[P + "model_seed"]))            # PtModel init: scipy truncnorm on the numpy global RNG
    return MPC(cc, n_envs=n_envs)


def _agent_arena(native, cuda, ag, z, tag, tensor_cores=0):
    from recovery_rl.arena import AgentArena
    sc = float(np.float32(float(z[tag + "_scale"])))
    ar = AgentArena(cuda, max_batch=64, action_scale=(sc, sc), use_tensor_cores=tensor_cores)
    ar.load_modules(ag.nets())

    class VF(object):
        arena = ar
    return VF()


def _load_ensemble(mpc, ora, cuda):
    """copy the oracle's (== the reference's) ensemble into the product model, by parameter name."""
    src = dict(ora.model.named())
    with torch.no_grad():
        for n, p in mpc.model.named_parameters():
            p.copy_(src[n].detach().reshape(p.shape).to(cuda))
    mpc.pack_model()


@pytest.mark.parametrize("tensor_cores", [0, 1])
@pytest.mark.parametrize("tag", ["nav", "maze"])
def test_planner_matches_reference(native, cuda, golden_dir, tag, tensor_cores):
    """tensor_cores = 1: the rollout's 256x256 contractions run on tcgen05 (mpc_tc.cu) and meet the same bar."""
    z = np.load(os.path.join(golden_dir, "mpc.npz"))
    za = np.load(os.path.join(golden_dir, "agent_algos_b64.npz"))
    P = tag + "_"
    stride = int(z["stride"])
    ora, ag = mpc_oracle(z, za, tag, trained=True)
    mpc = _product_mpc(z, tag)
    for n, p in mpc.model.named_parameters():                       # same scipy / numpy stream as the reference
        assert np.array_equal(p.detach().cpu().numpy().ravel()[::stride], z[P + "init_" + n]), n
    # MPC.train through the drop-in class: same bootstrap indices, same schedule; every mini-batch is one launch of
    # dyn_train_kernel (forward + NLL + backward + Adam for the five nets)
    np.random.seed(int(z[P + "train_seed"]))
    mpc.train(z[P + "train_obs"], z[P + "train_acs"], random=True, next_obs=z[P + "train_next"],
              epochs=int(z[P + "train_epochs"]))
    n_steps = int(z[P + "train_epochs"]) * int(np.ceil(len(z[P + "train_obs"]) / 32))
    assert int(mpc.model.adam_step.item()) == n_steps and int(mpc.model.ticket.item()) == 0
    for n, p in mpc.model.named_parameters():
        got = p.detach().cpu().numpy().ravel()[::stride]
        assert np.allclose(got, z[P + "trained_" + n], rtol=2e-3, atol=2e-5), (n, np.abs(got - z[P + "trained_" + n]).max())
    assert abs(mpc.last_train_loss - ora.losses[-1]) < 1e-3 * (1 + abs(ora.losses[-1])), (mpc.last_train_loss, ora.losses[-1])
    # cross-check: the same schedule through torch autograd on the GPU lands on the same parameters
    mpc_t = _product_mpc(z, tag)
    np.random.seed(int(z[P + "train_seed"]))
    mpc_t.train(z[P + "train_obs"], z[P + "train_acs"], random=True, next_obs=z[P + "train_next"],
                epochs=int(z[P + "train_epochs"]), use_torch=True)
    # (Adam turns a near-zero gradient into a step of up to lr regardless of its size, so isolated entries may differ
    # by a few lr = 1e-3 steps between two fp32 evaluation orders: bound those, require the bulk to agree tightly)
    for (n, p), (_, pt) in zip(mpc.model.named_parameters(), mpc_t.model.named_parameters()):
        diff = (p - pt).abs()
        ok = diff <= 2e-5 + 2e-3 * pt.abs()
        assert ok.float().mean().item() > 0.999 and diff.max().item() < 5e-3, (n, ok.float().mean().item(), diff.max().item())
    # the planner kernels on the reference's exact ensemble (teacher forcing) and safety critic
    _load_ensemble(mpc, ora, cuda)
    mpc.update_value_func(_agent_arena(native, cuda, ag, z, tag, tensor_cores))
    ac_seqs, eps, zs, eps2 = mpc_noise(z, tag)
    pop, npart = int(z[P + "popsize"]), int(z[P + "npart"])
    # _compile_cost (MPC.py:374-416)
    mpc.state.copy_(torch.from_numpy(z[P + "cost_obs"].reshape(2, 1)))
    mpc.samples.copy_(torch.from_numpy(ac_seqs)[None])
    native.mpc_rollout(mpc.cfg, mpc.agent_cfg, mpc.agent_arena, mpc.dyn_image, 1, mpc.state, mpc.samples, mpc.row_cost,
                       eps=torch.from_numpy(eps).to(cuda).contiguous())
    torch.cuda.synchronize()
    rc = mpc.row_cost[0].cpu().numpy()
    assert np.isfinite(rc).all()
    assert np.allclose(rc.mean(1), z[P + "cost_out"], rtol=1e-4, atol=1e-5), np.abs(rc.mean(1) - z[P + "cost_out"]).max()
    # MPC.act x2 (CEM loop, warm start shift): injected candidates and particle noise
    scale = float(z[P + "scale"])
    for i in range(2):
        a = mpc.act(z[P + "act_states"][i], 0, z=[torch.from_numpy(zs[i][k][None]).to(cuda).contiguous() for k in range(5)],
                    eps=[torch.from_numpy(eps2[i][k]).to(cuda).contiguous() for k in range(5)])
        assert np.allclose(a, z[P + "act_actions"][i], rtol=1e-3, atol=2e-4 * scale), (a, z[P + "act_actions"][i])
        assert np.allclose(mpc.prev_sol, z[P + "act_prev_sol"][i], rtol=1e-3, atol=2e-4 * scale)
    # production mode: device Philox for candidates and particle noise
    a = mpc.act(z[P + "act_states"][0], 0)
    assert np.isfinite(a).all() and (np.abs(a) <= scale + 1e-6).all()


@pytest.mark.parametrize("tensor_cores", [0, 1])
def test_planner_batched_envs_equal_single(native, cuda, golden_dir, tensor_cores):
    """n_envs copies planned in one launch == each env planned alone (rows of different envs never mix)."""
    z = np.load(os.path.join(golden_dir, "mpc.npz"))
    za = np.load(os.path.join(golden_dir, "agent_algos_b64.npz"))
    tag, P = "maze", "maze_"
    ora, ag = mpc_oracle(z, za, tag, trained=True)
    vf = _agent_arena(native, cuda, ag, z, tag, tensor_cores)
    E = 3
    ms = [_product_mpc(z, tag, n_envs=1) for _ in range(E)] + [_product_mpc(z, tag, n_envs=E)]
    for m in ms:
        _load_ensemble(m, ora, cuda)
        m.update_value_func(vf)
        m.has_been_trained = True
    rs = np.random.RandomState(0)
    pop, hor, npart = int(z[P + "popsize"]), int(z[P + "plan_hor"]), 20
    states = rs.uniform(-0.2, 0.2, (E, 2))
    zs = np.clip(rs.standard_normal((5, E, pop, hor * 2)), -2, 2)
    eps = rs.randn(5, E, hor, 5, pop * npart // 5, 2).astype(np.float32)
    big = ms[-1]
    big.state.copy_(torch.from_numpy(states.T.copy()))
    act = big.plan(z=[torch.from_numpy(zs[k]).to(cuda).contiguous() for k in range(5)],
                   eps=[torch.from_numpy(eps[k]).to(cuda).contiguous() for k in range(5)]).cpu().numpy()
    for e in range(E):
        a = ms[e].act(states[e], 0, z=[torch.from_numpy(zs[k][e][None]).to(cuda).contiguous() for k in range(5)],
                      eps=[torch.from_numpy(eps[k][e]).to(cuda).contiguous() for k in range(5)])
        assert np.array_equal(a, act[e]), (e, a, act[e])


def test_vectorised_model_based_recovery_runs(native, cuda, tmp_path):
    """BASELINE config 5 in small: Maze, model-based recovery, N env copies x (popsize x 20) CEM particles through
    the vector engine (planner inside the captured CUDA graph), ensemble pre-trained on the demos."""
    import arg_utils
    from recovery_rl.experiment import Experiment
    args = arg_utils.get_args(["--env-name", "maze", "--use_recovery", "--gamma_safe", "0.5", "--eps_safe", "0.15",
                               "--pos_fraction", "0.3", "--num_unsafe_transitions", "1500", "--critic_safe_pretraining_steps", "20",
                               "--batch_size", "64", "--num_envs", "128", "--mpc_popsize", "20", "--num_steps", "7000",
                               "--seed", "3", "--logdir", str(tmp_path), "--replay_size", "50000", "--safe_replay_size", "50000",
                               "--recovery_policy_update_freq", "1"])
    exp = Experiment(args)
    stats = exp.run()
    eng = exp.engine
    assert eng.mpc is not None and eng.mpc.has_been_trained and eng.graph is not None
    assert stats[-1]["total_numsteps"] > 7000 and stats[-1]["error"] == 0
    assert stats[-1]["viol_and_recovery"] + stats[-1]["viol_and_no_recovery"] == stats[-1]["num_viols"]
    assert torch.isfinite(eng.arena[:eng.agent.grad_off]).all() and torch.isfinite(eng.mpc.dyn_image).all()
    assert (eng.action_real.abs() <= 0.1 + 1e-6).all()
