"""GPU parity of the SAC / Q_risk / recovery-policy kernels (through the C ABI) against golden vectors
recorded from the reference's own torch modules, and against the oracle restatement.
Tolerance (BASELINE.json north_star): fp32 losses and Q-values within 1e-4 rtol."""
import os

import numpy as np
import pytest
import torch

from oracle.agent import Agent

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def close(a, b, rtol=RTOL, atol=1e-6):
    """|a - b| <= atol + rtol * |b| elementwise; atol may be a per-element array."""
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    assert a.shape == b.shape, (a.shape, b.shape)
    # absolute floor: 1e-5 of the tensor's largest magnitude -- a 256-term fp32 dot product whose terms cancel
    # (Q-values near 0 from O(1) summands) cannot agree to 1e-4 of its own value across summation orders
    floor = 0.1 * rtol * (np.abs(b).max() if b.size else 0.0)
    err = np.abs(a - b) - (np.asarray(atol, np.float64).ravel() + floor + rtol * np.abs(b))
    assert (err <= 0).all(), "max rel err %.3e (abs %.3e) at %d" % (
        np.max(np.abs(a - b) / (np.abs(b) + 1e-12)), np.max(np.abs(a - b)), int(np.argmax(err)))


def _oracle_agent(z):
    seed = int(z["seed"])
    torch.manual_seed(seed)
    np.random.seed(seed)
    sc = float(z["scale"])
    return Agent(action_scale=(np.float32(sc),) * 2, gamma=float(z["gamma"]), alpha=float(z["alpha"]), tau=float(z["tau"]),
                 gamma_safe=float(z["gamma_safe"]), tau_safe=float(z["tau_safe"]), eps_safe=float(z["eps_safe"]),
                 lr=float(z["lr"]))


def _arena(native, dev, z, B, **kw):
    from recovery_rl.arena import AgentArena
    sc = float(np.float32(float(z["scale"])))
    return AgentArena(dev, max_batch=B, gamma=float(z["gamma"]), alpha=float(z["alpha"]), tau=float(z["tau"]),
                      gamma_safe=float(z["gamma_safe"]), tau_safe=float(z["tau_safe"]), eps_safe=float(z["eps_safe"]),
                      lr=float(z["lr"]), action_scale=(sc, sc), **kw)


def _sync_from_oracle(native, ar, ora):
    """teacher forcing: every update starts from the reference's exact weights and Adam state, so the
    comparison is 'one update from identical weights / batch / eps' (SURVEY.md 8c)."""
    ar.load_modules(ora.nets())
    T0 = native.C_ADAM_T0
    ar.load_optimizer("critic", ora.critic_optim, T0 + 0)
    ar.load_optimizer("policy", ora.policy_optim, T0 + 1)
    ar.load_optimizer("qrisk", ora.qrisk_optim, T0 + 2)
    ar.load_optimizer("recovery", ora.recovery_optim, T0 + 3)


def logp_cond(action, scale):
    """fp32 conditioning of log pi (model.py:333-335): log(scale*(1 - y^2) + 1e-6) with y = tanh(x) within
    2 ulp of +-1 has absolute error ~ 2|y| dy / (1 - y^2 + 1e-6); ANY fp32 evaluation order (torch CPU vs
    torch CUDA included) differs by this much when the action saturates.  Per-row bound, summed over dims."""
    y = np.clip(np.asarray(action, np.float64) / scale, -1, 1)
    return 1e-5 + (6e-7 / (1.0 - y * y + 1e-6)).sum(1)


def _dev(x, dev):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32).to(dev)


@pytest.mark.parametrize("tc_updates", [0, 1, 2])
@pytest.mark.parametrize("tag", ["nav1_b256", "maze_b64"])
def test_updates_and_acting_vs_reference(native, cuda, golden_dir, tag, tc_updates):
    """tc_updates = 1: the forward passes of the updates run their 256x256 contractions on tcgen05 (fp16 hi/lo
    split, fwd_tc_kernel) and must meet the same 1e-4 bar as the fp32 SIMT tiles; tc_updates = 2: additionally the
    loss / sample-backward stages run as last-CTA tails of the producing kernels (the path bench.py times)."""
    z = np.load(os.path.join(golden_dir, "agent_%s.npz" % tag))
    B = int(z["B"])
    stride = int(z["stride"])
    n_upd = int(z["n_updates"])
    ora = _oracle_agent(z)                       # bit-identical init to the reference (same torch RNG order)
    ar = _arena(native, cuda, z, B, use_tensor_cores=tc_updates)
    ar.load_modules(ora.nets())
    if tag == "nav1_b256":                       # the dumped reference init must equal what we loaded
        for net in ("critic", "policy", "qrisk", "recovery"):
            for i, p in enumerate(ar.params(net)):
                assert np.array_equal(p, z["init_%s_%d" % (net, i)].reshape(p.shape))
    losses = torch.zeros(16, device=cuda)
    for u in range(n_upd):
        p = "sac%d_" % u
        _sync_from_oracle(native, ar, ora)
        ora.sac_update([z[p + k] for k in ("s", "a", "r", "s2", "m")], z[p + "eps_next"], z[p + "eps_cur"], u)
        ar.set_batch("sac", z[p + "s"], z[p + "a"], z[p + "r"], z[p + "s2"], z[p + "m"])
        native.sac_backward(ar.cfg, ar.arena, ar.counters, losses, _dev(z[p + "eps_next"], cuda), _dev(z[p + "eps_cur"], cuda))
        torch.cuda.synchronize()
        close(ar.scratch("next_a", 2)[:B].cpu(), z[p + "next_action"])
        sc = float(z["scale"])
        c_next = logp_cond(z[p + "next_action"], sc)
        c_cur = logp_cond(z[p + "pi"], sc)
        close(ar.scratch("next_logp")[:B].cpu(), z[p + "next_log_pi"], atol=c_next)
        close(ar.scratch("target")[:B].cpu(), z[p + "target"], atol=1e-6 + float(z["alpha"]) * c_next)
        close(ar.scratch("qf1")[:B].cpu(), z[p + "qf1"])
        close(ar.scratch("qf2")[:B].cpu(), z[p + "qf2"])
        close(ar.scratch("pi", 2)[:B].cpu(), z[p + "pi"])
        close(ar.scratch("logp")[:B].cpu(), z[p + "log_pi"], atol=c_cur)
        close(ar.scratch("minq")[:B].cpu(), z[p + "min_qf_pi"])
        al = float(z["alpha"])
        dy = al * c_next                                        # bound on the TD-target deviation per row
        tol_q1 = 1e-6 + np.mean(2 * np.abs(z[p + "qf1"][:, 0] - z[p + "target"][:, 0]) * dy + dy * dy)
        tol_q2 = 1e-6 + np.mean(2 * np.abs(z[p + "qf2"][:, 0] - z[p + "target"][:, 0]) * dy + dy * dy)
        close(losses[:3].cpu(), z[p + "losses"][:3], atol=np.array([tol_q1, tol_q2, 1e-6 + al * c_cur.mean()]))
        if u == 0:                               # gradients of the first update (strided in the fixture)
            for net, cnt in (("critic", 12), ("policy", 8)):
                for i in range(cnt):
                    ref = z["%s%s_grads_%d" % (p, net, i)]
                    g = ar.grad(net, i).cpu().numpy().ravel()[::stride]
                    close(g, ref, rtol=2e-4, atol=1e-6 * (1 + np.abs(ref).max()))
        native.sac_apply(ar.cfg, ar.arena, ar.counters)
        if u in (0, n_upd - 1):
            for net in ("critic", "critic_target", "policy"):
                for i, w in enumerate(ar.params(net)):
                    close(w.ravel()[::stride], z["after_sac%d_%s_%d" % (u, net, i)], rtol=2e-4, atol=2e-6)
        p = "qr%d_" % u
        _sync_from_oracle(native, ar, ora)
        Lq = ora.qrisk_update([z[p + k] for k in ("s", "a", "c", "s2", "m")], z[p + "eps_next"], z[p + "eps_rec"])
        ar.set_batch("qr", z[p + "s"], z[p + "a"], z[p + "c"], z[p + "s2"], z[p + "m"])
        native.qrisk_backward(ar.cfg, ar.arena, ar.counters, losses, _dev(z[p + "eps_next"], cuda))
        torch.cuda.synchronize()
        close(ar.scratch("qr_next_a", 2)[:B].cpu(), z[p + "next_action"])
        close(ar.scratch("qr_q1")[:B].cpu(), z[p + "q1"])
        close(ar.scratch("qr_q2")[:B].cpu(), z[p + "q2"])
        close(ar.scratch("qr_target")[:B].cpu(), z[p + "target"])
        close(losses[:2].cpu(), z[p + "losses"])
        native.qrisk_apply(ar.cfg, ar.arena, ar.counters)
        native.recovery_backward(ar.cfg, ar.arena, ar.counters, losses, _dev(z[p + "eps_rec"], cuda))
        native.recovery_apply(ar.cfg, ar.arena, ar.counters)
        torch.cuda.synchronize()
        close(losses[2:3].cpu(), Lq[2:3])                      # recovery-policy loss on the post-step critic
        if u in (0, n_upd - 1):
            for net in ("qrisk", "qrisk_target", "recovery"):
                for i, w in enumerate(ar.params(net)):
                    close(w.ravel()[::stride], z["after_qr%d_%s_%d" % (u, net, i)], rtol=2e-4, atol=2e-6)
    c = ar.counters.cpu().numpy()
    assert c[native.C_SAC_UPDATES] == n_upd and c[native.C_QRISK_UPDATES] == n_upd
    assert list(c[native.C_ADAM_T0:native.C_ADAM_T0 + 4]) == [n_upd] * 4
    _sync_from_oracle(native, ar, ora)           # final weights == the reference's (bit-exact oracle)

    for tc in (0, 1):
        _check_acting(native, cuda, ar, z, tc)
    ar.cfg.use_tensor_cores = tc_updates


def _check_acting(native, cuda, ar, z, tc):
    """composite action selection (experiment.py:546-577) on the reference's final weights; tc = 1 runs the
    256x256 contractions on the tensor cores (tcgen05, fp16 hi/lo split) and must meet the same bar."""
    ar.cfg.use_tensor_cores = tc
    N = len(z["act_s"])
    ar.cfg.eps_safe = float(z["act_thresh"])
    st = torch.from_numpy(np.ascontiguousarray(z["act_s"].T)).to(cuda)
    a_task = torch.zeros(N, 2, device=cuda)
    a_real = torch.zeros(N, 2, device=cuda)
    rec = torch.zeros(N, dtype=torch.uint8, device=cuda)
    qv = torch.zeros(N, device=cuda)
    e_task, e_rec = _dev(z["act_eps_task"], cuda), _dev(z["act_eps_rec"], cuda)
    native.agent_act(ar.cfg, ar.arena, N, st, None, a_task, a_real, rec, qv, e_task, e_rec)
    torch.cuda.synchronize()
    close(a_task.cpu(), z["act_task"], atol=2e-6)
    close(qv.cpu(), z["act_qrisk"], atol=2e-6)
    g_rec = z["act_recovery"].astype(bool)
    margin = np.abs(z["act_qrisk"][:, 0] - float(z["act_thresh"])) > 1e-5     # flags away from the threshold
    got_rec = rec.cpu().numpy().astype(bool)
    assert np.array_equal(got_rec[margin], g_rec[margin])
    same = got_rec == g_rec
    close(a_real.cpu().numpy()[same], z["act_real"][same], atol=2e-6)
    assert g_rec.sum() > 50 and (~g_rec).sum() > 50
    # eval mode: mean action (sac.py:166-167)
    native.agent_act(ar.cfg, ar.arena, N, st, None, a_task, a_real, rec, qv, e_task, e_rec, eval=True)
    torch.cuda.synchronize()
    close(a_task.cpu(), z["act_mean"], atol=2e-6)
    close(qv.cpu(), z["act_qrisk_mean"], atol=2e-6)
    # stand-alone entry points
    q1 = torch.zeros(N, device=cuda)
    q2 = torch.zeros(N, device=cuda)
    s32 = _dev(z["act_s"], cuda)
    native.twin_q_forward(ar.cfg, ar.arena, native.NET_QRISK, N, s32, _dev(z["act_task"], cuda), q1, q2)
    close(torch.maximum(q1, q2).cpu(), z["act_qrisk"], atol=2e-6)
    act = torch.zeros(N, 2, device=cuda)
    lp = torch.zeros(N, device=cuda)
    mean = torch.zeros(N, 2, device=cuda)
    native.policy_sample(ar.cfg, ar.arena, native.NET_POLICY, N, s32, e_task, act, lp, mean)
    close(act.cpu(), z["act_task"], atol=2e-6)
    close(lp.cpu(), z["act_logp"], atol=logp_cond(z["act_task"], float(z["scale"])))
    native.policy_sample(ar.cfg, ar.arena, native.NET_RECOVERY, N, s32, e_rec, act, lp, mean)
    close(act.cpu(), z["act_rec"], atol=2e-6)
    close(mean.cpu(), z["act_rec_mean"], atol=2e-6)


def test_partial_batch_and_closed_gate(native, cuda, golden_dir):
    """rows < max_batch (qrisk.py:100-104, pre-training with few demos) and rows == 0 (gate closed)."""
    z = np.load(os.path.join(golden_dir, "agent_nav1_b256.npz"))
    ora = _oracle_agent(z)
    ar = _arena(native, cuda, z, 256)
    ar.load_modules(ora.nets())
    rows = 77
    losses = torch.zeros(16, device=cuda)
    p = "sac0_"
    batch = [z[p + k][:rows] for k in ("s", "a", "r", "s2", "m")]
    e_next, e_cur = _dev(z[p + "eps_next"][:rows], cuda), _dev(z[p + "eps_cur"][:rows], cuda)
    before = ar.arena[:ar.grad_off].clone()
    ar.set_batch("sac", *batch)
    ar.counters[native.C_SAC_ROWS] = 0
    native.sac_backward(ar.cfg, ar.arena, ar.counters, losses, e_next, e_cur)
    native.sac_apply(ar.cfg, ar.arena, ar.counters)
    torch.cuda.synchronize()
    assert torch.equal(before, ar.arena[:ar.grad_off])          # nothing moved, no Adam step counted
    assert int(ar.counters[native.C_SAC_UPDATES]) == 0
    ar.counters[native.C_SAC_ROWS] = rows
    native.sac_backward(ar.cfg, ar.arena, ar.counters, losses, e_next, e_cur)
    native.sac_apply(ar.cfg, ar.arena, ar.counters)
    L = ora.sac_update(batch, z[p + "eps_next"][:rows], z[p + "eps_cur"][:rows], 0)
    torch.cuda.synchronize()
    close(losses[:3].cpu(), L[:3])
    for net in ("critic", "critic_target", "policy"):
        for w, o in zip(ar.params(net), ora.params(net)):
            close(w, o, rtol=2e-4, atol=2e-6)
    p = "qr0_"
    batch = [z[p + k][:rows] for k in ("s", "a", "c", "s2", "m")]
    ar.set_batch("qr", *batch)
    native.qrisk_backward(ar.cfg, ar.arena, ar.counters, losses, _dev(z[p + "eps_next"][:rows], cuda))
    native.qrisk_apply(ar.cfg, ar.arena, ar.counters)
    native.recovery_backward(ar.cfg, ar.arena, ar.counters, losses, _dev(z[p + "eps_rec"][:rows], cuda))
    native.recovery_apply(ar.cfg, ar.arena, ar.counters)
    Lq = ora.qrisk_update(batch, z[p + "eps_next"][:rows], z[p + "eps_rec"][:rows])
    torch.cuda.synchronize()
    close(losses[:3].cpu(), Lq)
    for net in ("qrisk", "qrisk_target", "recovery"):
        for w, o in zip(ar.params(net), ora.params(net)):
            close(w, o, rtol=2e-4, atol=2e-6)


def test_act_tensor_cores_match_simt_many_tiles(native, cuda, golden_dir):
    """tcgen05 acting kernel vs the fp32 SIMT kernel on 40,001 rows (several tiles per CTA, ragged last tile):
    actions / Q_risk within 1e-4, recovery flags equal away from the threshold."""
    z = np.load(os.path.join(golden_dir, "agent_maze_b64.npz"))
    ora = _oracle_agent(z)
    ar = _arena(native, cuda, z, 64)
    ar.load_modules(ora.nets())
    rs = np.random.RandomState(0)
    N = 40001
    st = torch.from_numpy(rs.uniform(-0.28, 0.28, (2, N))).to(cuda)
    e_task, e_rec = _dev(rs.randn(N, 2), cuda), _dev(rs.randn(N, 2), cuda)
    outs = []
    for tc in (0, 1):
        ar.cfg.use_tensor_cores = tc
        q = torch.zeros(N, device=cuda)
        native.agent_act(ar.cfg, ar.arena, N, st, None, torch.zeros(N, 2, device=cuda), torch.zeros(N, 2, device=cuda),
                         torch.zeros(N, dtype=torch.uint8, device=cuda), q, e_task, e_rec)
        med = float(q.median())
        ar.cfg.eps_safe = med                      # both branches of experiment.py:555 occur
        a_task = torch.zeros(N, 2, device=cuda); a_real = torch.zeros(N, 2, device=cuda)
        rec = torch.zeros(N, dtype=torch.uint8, device=cuda)
        native.agent_act(ar.cfg, ar.arena, N, st, None, a_task, a_real, rec, q, e_task, e_rec)
        torch.cuda.synchronize()
        outs.append((a_task.cpu().numpy(), a_real.cpu().numpy(), rec.cpu().numpy().astype(bool), q.cpu().numpy(), med))
    (t0, r0, f0, q0, m0), (t1, r1, f1, q1, m1) = outs
    close(t1, t0, atol=2e-6)
    close(q1, q0, atol=2e-6)
    sure = np.abs(q0 - m0) > 1e-5
    assert abs(m0 - m1) < 1e-6
    assert np.array_equal(f0[sure], f1[sure]) and 1000 < f0.sum() < N - 1000
    same = f0 == f1
    close(r1[same], r0[same], atol=2e-6)
