"""GPU parity of the comparison-algorithm branches of SAC.update_parameters / select_action (LR, RSPO, RCPO,
automatic entropy tuning, --policy Deterministic, SQRL filter) through the C ABI, against golden vectors recorded
from the reference's own SAC class (tests/golden/agent_algos_b64.npz) with the oracle as teacher (every update
starts from the reference's exact weights, Adam moments and multipliers).  fp32 losses / Q-values within 1e-4."""
import os

import numpy as np
import pytest
import torch

from test_agent_gpu import close, logp_cond, _dev
from test_oracle import ALGO_TAGS, algo_oracle_agent, algo_noise

pytestmark = pytest.mark.gpu


def _arena(native, dev, z, tag, B, tc=1):
    from recovery_rl.arena import AgentArena
    P = tag + "_"
    f = z[P + "flags"]
    sc = float(np.float32(float(z[P + "scale"])))
    return AgentArena(dev, max_batch=B, gamma=float(z[P + "gamma"]), alpha=float(z[P + "alpha"]), tau=float(z[P + "tau"]),
                      gamma_safe=float(z[P + "gamma_safe"]), tau_safe=float(z[P + "tau_safe"]),
                      eps_safe=float(z[P + "eps_safe"]), lr=float(z[P + "lr"]), action_scale=(sc, sc),
                      mf_recovery=bool(int(z[P + "mf_recovery"])), dgd=bool(f[0]), update_nu=bool(f[1]), rcpo=bool(f[2]),
                      auto_alpha=bool(f[3]), deterministic=bool(f[4]), nu=float(z[P + "nu"]),
                      lambda_rcpo=float(z[P + "lambda_RCPO"]), use_tensor_cores=tc)   # update GEMMs on tcgen05 (2: + fused stages)


def _pad_det(mods):
    """the arena keeps the Gaussian layout: a DeterministicPolicy's 6 tensors + a zero log_std head."""
    class Padded(object):
        def __init__(self, m):
            self.m = m

        def parameters(self):
            return list(self.m.parameters()) + [torch.zeros(2, 256), torch.zeros(2)]
    out = dict(mods)
    out["policy"] = Padded(mods["policy"])
    return out


def _opt_state(opt):
    p = opt.param_groups[0]["params"][0]
    st = opt.state.get(p, {})
    if "exp_avg" not in st:
        return 0.0, 0.0, 0
    return float(st["exp_avg"]), float(st["exp_avg_sq"]), int(st["step"])


def _sync(native, ar, ora):
    nets = ora.nets()
    ar.load_modules(_pad_det(nets) if ora.deterministic else nets)
    T0 = native.C_ADAM_T0
    ar.load_optimizer("critic", ora.critic_optim, T0 + 0)
    if not ora.deterministic:
        ar.load_optimizer("policy", ora.policy_optim, T0 + 1)
    else:       # 6 tensors of torch state, the padded head stays zero
        step = 0
        for i, p in enumerate(ora.policy_optim.param_groups[0]["params"]):
            st = ora.policy_optim.state.get(p, {})
            m, v = ar.adam_state("policy", i)
            if "exp_avg" in st:
                m.copy_(st["exp_avg"].reshape(-1).to(ar.device)); v.copy_(st["exp_avg_sq"].reshape(-1).to(ar.device))
                step = int(st["step"])
        ar.counters[T0 + 1] = step
    ar.load_optimizer("qrisk", ora.qrisk_optim, T0 + 2)
    ar.load_optimizer("recovery", ora.recovery_optim, T0 + 3)
    f32, f64 = ar.scalars()
    f32[native.S_ALPHA] = float(ora.alpha)
    f64[native.D_LAMBDA] = float(ora.lambda_rcpo)
    f64[native.D_LOG_NU] = float(ora.log_nu.item())
    f64[native.D_LOG_LAMBDA] = float(ora.log_lambda.item())
    for opt, (mi, vi, ti) in ((ora.nu_optim, (native.D_M_NU, native.D_V_NU, native.C_ADAM_T_NU)),
                              (ora.lambda_optim, (native.D_M_LAMBDA, native.D_V_LAMBDA, native.C_ADAM_T_LAMBDA))):
        m, v, t = _opt_state(opt)
        f64[mi] = m; f64[vi] = v
        ar.counters[ti] = t
    if ora.auto_alpha:
        f32[native.S_LOG_ALPHA] = float(ora.log_alpha.item())
        m, v, t = _opt_state(ora.alpha_optim)
        f32[native.S_M_ALPHA] = m; f32[native.S_V_ALPHA] = v
        ar.counters[native.C_ADAM_T_ALPHA] = t


def _noise_dev(e, B, dev):
    e = np.asarray(e, np.float32)
    if e.ndim == 1:                              # Deterministic policy: one noise vector for the whole batch
        e = np.repeat(e[None], B, 0)
    return _dev(e, dev)


@pytest.mark.parametrize("tc", [1, 2])
@pytest.mark.parametrize("tag", ALGO_TAGS)
def test_comparison_branches_vs_reference(native, cuda, golden_dir, tag, tc):
    z = np.load(os.path.join(golden_dir, "agent_algos_b64.npz"))
    B = int(z["B"])
    P = tag + "_"
    stride = int(z["stride"])
    ora = algo_oracle_agent(z, tag)
    ar = _arena(native, cuda, z, tag, B, tc)
    losses = torch.zeros(16, device=cuda)
    det = ora.deterministic
    sc = float(z[P + "scale"])
    for u in range(int(z["n_qr"])):              # Q_risk warm-up: GPU path checked against the oracle teacher
        q = "%sqr%d_" % (P, u)
        e_next, _ = algo_noise(z, tag, "qr", 500, u)
        batch = [z[q + k] for k in ("s", "a", "c", "s2", "m")]
        _sync(native, ar, ora)
        Lq = ora.qrisk_update(batch, e_next, None)
        ar.set_batch("qr", *batch)
        native.qrisk_backward(ar.cfg, ar.arena, ar.counters, losses, _noise_dev(e_next, B, cuda))
        native.qrisk_apply(ar.cfg, ar.arena, ar.counters)
        native.recovery_apply(ar.cfg, ar.arena, ar.counters)
        torch.cuda.synchronize()
        close(losses[:2].cpu(), Lq[:2])
        if u == int(z["n_qr"]) - 1:
            for w, o in zip(ar.params("qrisk"), ora.params("qrisk")):
                close(w, o, rtol=2e-4, atol=2e-6)
    n_upd = int(z["n_updates"])
    for u in range(n_upd):
        q = "%ssac%d_" % (P, u)
        e_next, e_cur = algo_noise(z, tag, "sac", 1000, u)
        batch = [z[q + k] for k in ("s", "a", "r", "s2", "m")]
        nu_arg = float(z[q + "nu_arg"])
        _sync(native, ar, ora)
        alpha_used = float(ora.alpha)
        ora.sac_update(batch, e_next, e_cur, u, nu=nu_arg)
        ar.set_batch("sac", *batch)
        ar.set_nu_arg(nu_arg)
        native.sac_backward(ar.cfg, ar.arena, ar.counters, losses, _noise_dev(e_next, B, cuda), _noise_dev(e_cur, B, cuda))
        torch.cuda.synchronize()
        c_cur = np.zeros(B) if det else logp_cond(z[q + "pi"], sc)
        c_next = np.zeros(B) if det else logp_cond(z[q + "next_action"], sc)
        close(ar.scratch("qf1")[:B].cpu(), z[q + "qf1"])
        close(ar.scratch("qf2")[:B].cpu(), z[q + "qf2"])
        close(ar.scratch("pi", 2)[:B].cpu(), z[q + "pi"])
        close(ar.scratch("target")[:B].cpu(), z[q + "target"], atol=1e-6 + alpha_used * c_next)
        close(ar.scratch("minq")[:B].cpu(), z[q + "min_qf_pi"])
        dy = alpha_used * c_next
        tol_q = [1e-6 + np.mean(2 * np.abs(z[q + k][:, 0] - z[q + "target"][:, 0]) * dy + dy * dy) for k in ("qf1", "qf2")]
        tol_pi = 1e-6 + alpha_used * c_cur.mean()
        gl = z[q + "losses"]
        close(losses[:3].cpu(), gl[:3], atol=np.array(tol_q + [tol_pi]))
        if int(z[P + "flags"][3]):               # alpha_loss = -(log_alpha * (log_pi + target_entropy)).mean()
            close(losses[3:4].cpu(), gl[3:4], atol=1e-6 + abs(float(ora.log_alpha.item())) * c_cur.mean() + 1e-5)
        if u == 0:
            for net, cnt in (("critic", 12), ("policy", 6 if det else 8)):
                for i in range(cnt):
                    ref = z["%s%s_grads_%d" % (q, net, i)]
                    g = ar.grad(net, i).cpu().numpy().ravel()[::stride]
                    close(g, ref, rtol=2e-4, atol=1e-6 * (1 + np.abs(ref).max()))
        native.sac_apply(ar.cfg, ar.arena, ar.counters)
        torch.cuda.synchronize()
        f32, f64 = ar.scalars()
        f32, f64 = f32.cpu().numpy(), f64.cpu().numpy()
        assert np.isclose(f64[native.D_LOG_NU], float(z[q + "log_nu"]), rtol=1e-7, atol=1e-9)
        assert np.isclose(f64[native.D_LOG_LAMBDA], float(z[q + "log_lambda"]), rtol=1e-7, atol=1e-9)
        assert np.isclose(f64[native.D_LAMBDA], float(z[q + "lambda"]), rtol=1e-7)
        assert np.isclose(f32[native.S_ALPHA], float(z[q + "alpha_after"]), rtol=1e-5, atol=1e-7)
        if u in (0, n_upd - 1):
            for net in ("critic", "critic_target", "policy"):
                for i, w in enumerate(ar.params(net)[:6 if (det and net == "policy") else None]):
                    close(w.ravel()[::stride], z["%safter_%s_%d" % (q, net, i)], rtol=2e-4, atol=2e-6)
    c = ar.counters.cpu().numpy()
    f = z[P + "flags"]
    assert c[native.C_ADAM_T_NU] == (n_upd if f[1] else 0) and c[native.C_ADAM_T_LAMBDA] == (n_upd if f[2] else 0)
    assert c[native.C_ADAM_T_ALPHA] == (n_upd if f[3] else 0)


def test_sqrl_action_filter_drop_in(native, cuda, golden_dir):
    """SAC.select_action with --use_constraint_sampling (sac.py:139-161) through the drop-in class: same 100 policy
    samples, same Categorical draw from the torch generator -> the reference's chosen action."""
    import argparse
    from recovery_rl.sac import SAC
    from env.spaces import Box
    z = np.load(os.path.join(golden_dir, "agent_algos_b64.npz"))
    tag, P = "sqrl", "sqrl_"
    ora = algo_oracle_agent(z, tag)
    for u in range(int(z["n_qr"])):
        q = "%sqr%d_" % (P, u)
        ora.qrisk_update([z[q + k] for k in ("s", "a", "c", "s2", "m")], z[q + "eps_next"], None)
    for u in range(int(z["n_updates"])):
        q = "%ssac%d_" % (P, u)
        ora.sac_update([z[q + k] for k in ("s", "a", "r", "s2", "m")], z[q + "eps_next"], z[q + "eps_cur"], u,
                       nu=float(z[q + "nu_arg"]))
    args = argparse.Namespace(gamma=0.99, tau=0.005, alpha=0.2, env_name="maze", policy="Gaussian", target_update_interval=1,
                              automatic_entropy_tuning=False, gamma_safe=0.5, eps_safe=0.15, nu=100.0, batch_size=64,
                              lr=3e-4, tau_safe=0.0002, MF_recovery=False, Q_sampling_recovery=False, hidden_size=256,
                              DGD_constraints=True, update_nu=True, RCPO=False, use_constraint_sampling=True,
                              lambda_RCPO=0.01, pos_fraction=0.3, cnn=False, vismpc_recovery=False)
    sc = np.float32(0.1)
    agent = SAC(Box(-np.ones(2) * np.inf, np.ones(2) * np.inf), Box(-np.ones(2) * sc, np.ones(2) * sc), args, "/tmp/none")
    agent.arena.load_modules(ora.nets())
    real_randn = torch.randn
    n_match = 0
    n = len(z["sqrl_sel_s"])
    try:
        for i in range(n):
            eps = torch.from_numpy(z["sqrl_sel_eps"][i])
            torch.randn = lambda *shape, _e=eps: _e.clone()
            agent.eps_safe = float(z["sqrl_sel_thresh"][i])
            torch.manual_seed(7000 + i)
            a = agent.select_action(z["sqrl_sel_s"][i])
            n_match += int(np.allclose(a, z["sqrl_sel_action"][i], rtol=1e-4, atol=2e-6))
    finally:
        torch.randn = real_randn
    # a Q_risk value within fp32 round-off of the threshold can move one sample in or out of the filtered set and
    # with it the Categorical draw; everything else must be the reference's action
    assert n_match >= n - 2, (n_match, n)
