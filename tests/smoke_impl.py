"""smoke: a few vector steps of the whole hot path on cuda:0 (act -> env step -> replay push -> sample ->
SAC + Q_risk + recovery update), every stage checked against the oracle on the same inputs.
Called by __graft_entry__.smoke() and by tests/test_engine_gpu.py."""
import numpy as np
import torch


def _close(a, b, what, rtol=1e-4, atol=1e-5):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    floor = 0.1 * rtol * (np.abs(b).max() if b.size else 0)
    bad = np.abs(a - b) > atol + floor + rtol * np.abs(b)
    if bad.any():
        raise AssertionError("%s: %d/%d mismatches, max abs err %.3e" % (what, bad.sum(), bad.size, np.abs(a - b).max()))


def run(env_name="navigation1", n=512, B=64, steps=4, seed=3, verbose=True, gamma_safe=0.8, eps_safe=0.3, pos_fraction=-1.0,
        replay=8192, demos_n=400, tensor_cores=0, deterministic=False):
    from oracle import envs as oenvs
    from oracle.agent import Agent
    from recovery_rl import native
    from recovery_rl.engine import VecEngine, ACTION_SCALE

    native.require_cuda()
    kind = oenvs.KIND_BY_NAME[env_name]
    sc = ACTION_SCALE[env_name]
    torch.manual_seed(seed)
    np.random.seed(seed)
    ora = Agent(action_scale=(np.float32(sc),) * 2, gamma_safe=gamma_safe, eps_safe=eps_safe, deterministic=deterministic)
    eng = VecEngine(env_name, n, batch_size=B, replay_size=max(replay, 2 * n), safe_replay_size=max(replay, 2 * n),
                    gamma_safe=gamma_safe, eps_safe=eps_safe, pos_fraction=pos_fraction, seed=seed, host_inputs=True,
                    start_steps=0, use_tensor_cores=tensor_cores, deterministic=deterministic)
    eng.init_agent(ora.nets())
    rs = np.random.RandomState(seed)
    if kind == oenvs.MAZE:
        demos = oenvs.maze_offline_data(demos_n, rs)
    else:
        demos = oenvs.nav_offline_data(kind, demos_n)
    eng.push_offline(demos)
    eng.pretrain_qrisk(3, n_demos=len(demos))
    # mirror the pre-trained weights into the oracle so that both sides start the rollout identically
    def mirror():
        P = {net: eng.agent.params(net) for net in native.NET_NAMES}
        ora.load(lambda net, i: P[net][i])
    mirror()
    draws = rs.rand(n, 2) if kind == oenvs.MAZE else rs.randn(n, 2)
    eng.reset(torch.from_numpy(np.ascontiguousarray(draws.T)).to(eng.device))
    state = oenvs.maze_reset_from_uniform(draws) if kind == oenvs.MAZE else oenvs.nav_reset(draws)
    ep_steps = np.zeros(n, int)
    assert np.array_equal(eng.state.cpu().numpy().T, state), "reset mismatch"
    for t in range(steps):
        inp = dict(reset_draws=(rs.rand(n, 2) if kind == oenvs.MAZE else rs.randn(n, 2)).T,
                   eps_task=rs.randn(n, 2).astype(np.float32), eps_rec=rs.randn(n, 2).astype(np.float32),
                   rand_u=rs.rand(n, 2).astype(np.float32),
                   sac_eps_next=rs.randn(B, 2).astype(np.float32), sac_eps_cur=rs.randn(B, 2).astype(np.float32),
                   qr_eps_next=rs.randn(B, 2).astype(np.float32), qr_eps_rec=rs.randn(B, 2).astype(np.float32))
        if kind != oenvs.MAZE:
            inp["env_noise"] = rs.randn(n, 2).T
        if deterministic:
            # --policy Deterministic (model.py:475-481): the policy's draw is N(0, 0.1) clamped to +-0.25 -- one vector per
            # sample() call, i.e. per env copy when acting and ONE for the whole batch inside an update
            clamp = lambda z: np.clip(np.float32(0.1) * z, -0.25, 0.25).astype(np.float32)
            inp["eps_task"] = clamp(inp["eps_task"])
            for k in ("sac_eps_next", "sac_eps_cur", "qr_eps_next"):
                inp[k] = np.repeat(clamp(inp[k][:1]), B, axis=0)
        before = eng.read_counters()
        out = eng.step_host(inp)
        cn = eng.read_counters()
        # ---- updates that preceded the action (experiment.py:397-415), checked on the sampled batches ----
        if before["task_len"] > B:
            a = eng.agent
            batch = [a.scratch(k, w)[:B].cpu().numpy() for k, w in (("sac_s", 2), ("sac_a", 2), ("sac_r", None),
                                                                     ("sac_s2", 2), ("sac_m", None))]
            L = ora.sac_update(batch, inp["sac_eps_next"], inp["sac_eps_cur"], before["sac_updates"])
            _close(out["losses"][:3].numpy(), L[:3], "SAC losses", atol=1e-4)
            assert cn["sac_updates"] == before["sac_updates"] + 1
            if cn["qrisk_updates"] > before["qrisk_updates"]:     # the online gate (experiment.py:407-410) opened
                batch = [a.scratch(k, w)[:B].cpu().numpy() for k, w in (("qr_s", 2), ("qr_a", 2), ("qr_c", None),
                                                                         ("qr_s2", 2), ("qr_m", None))]
                Lq = ora.qrisk_update(batch, inp["qr_eps_next"], inp["qr_eps_rec"])
                _close(out["losses"][8:10].numpy(), Lq[:2], "Q_risk losses", atol=1e-5)
            # keep the oracle's weights identical to the device's for the acting check below
            mirror()
        else:
            assert cn["sac_updates"] == before["sac_updates"]
        # ---- composite action (experiment.py:546-577) ----
        a_task, a_real, rec, qv = ora.act(state, inp["eps_task"], inp["eps_rec"], eps_safe=eps_safe)
        _close(eng.action_task.cpu().numpy(), a_task, "task action")
        _close(eng.qrisk.cpu().numpy(), qv, "Q_risk value")
        g_rec = out["recovery"].numpy().astype(bool)
        sure = np.abs(qv - eps_safe) > 1e-4
        assert np.array_equal(g_rec[sure], rec[sure]), "recovery flags"
        same = g_rec == rec
        _close(out["action"].numpy()[same], a_real[same], "executed action")
        # ---- env step on the DEVICE's executed action: bit-exact (constraint flags, next state, reward) ----
        act = out["action"].numpy()
        if kind == oenvs.MAZE:
            ns, r, d, c, su = oenvs.maze_step(state, act, ep_steps)
            done_h = d
        else:
            ns, r, d, c, su = oenvs.nav_step(kind, state, act, inp["env_noise"].T)
            done_h = d | (ep_steps + 1 == 100)
        assert np.array_equal(out["next_state"].numpy().T, ns), "next_state"
        assert np.array_equal(out["reward"].numpy(), r), "reward"
        assert np.array_equal(out["constraint"].numpy().astype(bool), c), "constraint flags"
        assert np.array_equal(out["done"].numpy().astype(bool), done_h), "done flags"
        assert np.array_equal(out["success"].numpy().astype(bool), su), "success flags"
        fresh = oenvs.maze_reset_from_uniform(inp["reset_draws"].T) if kind == oenvs.MAZE else oenvs.nav_reset(inp["reset_draws"].T)
        state = np.where(done_h[:, None], fresh, ns)
        ep_steps = np.where(done_h, 0, ep_steps + 1)
        assert np.array_equal(eng.state.cpu().numpy().T, state), "state after auto-reset"
        assert cn["total_numsteps"] == (t + 1) * n and cn["task_len"] == min((t + 1) * n, eng.task_cap)
        assert cn["error"] == 0
        if verbose:
            print("smoke step %d ok: viols=%d recoveries=%d sac_updates=%d qrisk_updates=%d" % (
                t, int(c.sum()), int(g_rec.sum()), cn["sac_updates"], cn["qrisk_updates"]))
    return True
