"""Agent / engine checkpoints.  The reference has none for the agent (only args.pkl, run_stats.pkl
and VisualMPC's torch.save); this is the `state_dict`-compatible save / resume named in SURVEY.md §8f.2.

A checkpoint is one torch.save'd dict:
  nets      {net: OrderedDict(reference state_dict name -> tensor)}   -- loadable into the reference's own
            QNetwork / QNetworkConstraint / GaussianPolicy / StochasticPolicy modules (model.py:49-76,172-199,
            295-343,489-530; the dead BatchNorm1d buffers of QNetworkConstraint are emitted at their init values)
  optim     {net: {"exp_avg": [...], "exp_avg_sq": [...], "step": int}}   torch.optim.Adam state per parameter
  scalars   the 32-float multiplier block (alpha / nu / lambda, their logs and Adam moments)
  counters  the int64 device counter block (step / episode / violation counts, ring positions, Adam steps)
  engine    (VecEngine only) env state, both replay rings (valid prefix) + flags, the CPython-compatible sampler
            state -- Philox draws are keyed by the vector-step counter, so a resumed run continues bit-identically.
            Sharded runs (world > 1) write ONE FILE PER RANK (`shard_path`): env copies, replay shards and the sampler
            stream (seeded seed + rank) are per-rank state; loading a shard written by another rank / world raises.
"""
import collections

import torch

from . import native
from . import model as model_mod

TRAINABLE = ("critic", "policy", "qrisk", "recovery")
_T_COUNTER = {"critic": native.C_ADAM_T0 + 0, "policy": native.C_ADAM_T0 + 1, "qrisk": native.C_ADAM_T0 + 2,
              "recovery": native.C_ADAM_T0 + 3}


def _param_names(net, deterministic=False):
    """reference state_dict parameter names in parameters() order (== arena tensor order)."""
    cls = {"critic": model_mod.QNetwork, "critic_target": model_mod.QNetwork, "qrisk": model_mod.QNetworkConstraint,
           "qrisk_target": model_mod.QNetworkConstraint, "recovery": model_mod.StochasticPolicy,
           "policy": model_mod.DeterministicPolicy if deterministic else model_mod.GaussianPolicy}[net]
    rng = torch.get_rng_state()                      # module construction draws init weights: keep the stream intact
    try:
        mod = cls(2, 2, 4) if net in ("policy", "recovery") else cls(2, 2, 4)
    finally:
        torch.set_rng_state(rng)
    return [n for n, _ in mod.named_parameters()]


def agent_state(arena):
    """arena: recovery_rl.arena.AgentArena -> checkpoint dict (host tensors)."""
    det = bool(arena.cfg.algo_flags & native.ALGO_DETERMINISTIC)
    nets, optim = collections.OrderedDict(), collections.OrderedDict()
    for net in native.NET_NAMES:
        names = _param_names(net, det)
        sd = collections.OrderedDict()
        for i, name in enumerate(names):
            sd[name] = arena.tensor(net, i).detach().cpu().clone()
        if net in ("qrisk", "qrisk_target"):       # buffers of the unused bn1 (never updated: forward never calls it)
            sd["bn1.running_mean"] = torch.zeros(4)
            sd["bn1.running_var"] = torch.ones(4)
            sd["bn1.num_batches_tracked"] = torch.tensor(0)
        nets[net] = sd
    c = arena.counters.cpu()
    for net in TRAINABLE:
        ms, vs = [], []
        for i in range(len(_param_names(net, det))):
            m, v = arena.adam_state(net, i)
            ms.append(m.detach().cpu().clone()); vs.append(v.detach().cpu().clone())
        optim[net] = {"exp_avg": ms, "exp_avg_sq": vs, "step": int(c[_T_COUNTER[net]])}
    return {"format": "rrl-b200-1", "nets": nets, "optim": optim, "scalars": arena.scratch("scalars").cpu().clone(),
            "counters": c.clone(), "algo_flags": int(arena.cfg.algo_flags)}


def load_agent_state(arena, state, load_counters=True):
    det = bool(arena.cfg.algo_flags & native.ALGO_DETERMINISTIC)
    if int(state.get("algo_flags", 0)) != int(arena.cfg.algo_flags):
        raise ValueError("checkpoint was written with algo_flags %d, this agent has %d"
                         % (int(state.get("algo_flags", 0)), int(arena.cfg.algo_flags)))
    for net in native.NET_NAMES:
        sd = state["nets"][net]
        for i, name in enumerate(_param_names(net, det)):
            t = arena.tensor(net, i)
            t.copy_(sd[name].to(arena.device).reshape(t.shape))
    for net in TRAINABLE:
        st = state["optim"][net]
        for i, (m_src, v_src) in enumerate(zip(st["exp_avg"], st["exp_avg_sq"])):
            m, v = arena.adam_state(net, i)
            m.copy_(m_src.reshape(-1).to(arena.device)); v.copy_(v_src.reshape(-1).to(arena.device))
    arena.scratch("scalars").copy_(state["scalars"].to(arena.device))
    if load_counters:
        arena.counters.copy_(state["counters"].to(arena.device))
    else:
        for net in TRAINABLE:
            arena.counters[_T_COUNTER[net]] = int(state["optim"][net]["step"])
    arena.refresh()                                   # k-major / fp16 operand images of the loaded weights


def shard_path(path, rank, world):
    """file of rank `rank` of a `world`-rank run: the path itself for one GPU, else <path>.rank<r>of<w>."""
    return path if int(world) <= 1 else "%s.rank%dof%d" % (path, int(rank), int(world))


def check_shard(e, rank, world):
    """the engine dict `e` of a checkpoint must have been written by this rank of an equally sized run."""
    r, w = int(e.get("rank", 0)), int(e.get("world", 1))
    if (r, w) != (int(rank), int(world)):
        raise ValueError("checkpoint shard was written by rank %d of %d, this process is rank %d of %d "
                         "(env copies, replay shards and the sampler stream are per-rank state)" % (r, w, rank, world))


def engine_state(eng):
    """recovery_rl.engine.VecEngine -> checkpoint dict."""
    st = agent_state(eng.agent)
    c = eng.counters.cpu()
    tl, cl = int(c[native.C_TASK_LEN]), int(c[native.C_CONS_LEN])
    st["engine"] = {
        "env_name": eng.env_name, "num_envs": eng.n, "task_cap": eng.task_cap, "cons_cap": eng.cons_cap,
        "rank": int(eng.rank), "world": int(eng.world),
        "state": eng.state.cpu().clone(), "ep_steps": eng.ep_steps.cpu().clone(), "ep_return": eng.ep_return.cpu().clone(),
        "task_ring": eng.task_ring[:tl].cpu().clone(), "cons_ring": eng.cons_ring[:cl].cpu().clone(),
        "cons_flags": eng.cons_flags[:cl].cpu().clone(), "mt_state": eng.mt_state.cpu().clone(),
        "action_task": eng.action_task.cpu().clone(), "action_real": eng.action_real.cpu().clone(),
        "recovery": eng.recovery.cpu().clone(),
    }
    return st


def load_engine_state(eng, state):
    e = state["engine"]
    if e["env_name"] != eng.env_name or int(e["num_envs"]) != eng.n or int(e["task_cap"]) != eng.task_cap \
            or int(e["cons_cap"]) != eng.cons_cap:
        raise ValueError("checkpoint is for %s x %d envs (rings %d / %d)" % (e["env_name"], e["num_envs"], e["task_cap"],
                                                                           e["cons_cap"]))
    check_shard(e, eng.rank, eng.world)
    load_agent_state(eng.agent, state, load_counters=True)
    dev = eng.device
    eng.state.copy_(e["state"].to(dev)); eng.ep_steps.copy_(e["ep_steps"].to(dev)); eng.ep_return.copy_(e["ep_return"].to(dev))
    eng.task_ring.zero_(); eng.cons_ring.zero_(); eng.cons_flags.zero_()
    eng.task_ring[:len(e["task_ring"])].copy_(e["task_ring"].to(dev))
    eng.cons_ring[:len(e["cons_ring"])].copy_(e["cons_ring"].to(dev))
    eng.cons_flags[:len(e["cons_flags"])].copy_(e["cons_flags"].to(dev))
    eng.mt_state.copy_(e["mt_state"].to(dev))
    eng.action_task.copy_(e["action_task"].to(dev)); eng.action_real.copy_(e["action_real"].to(dev))
    eng.recovery.copy_(e["recovery"].to(dev))


def save(path, obj):
    """obj: VecEngine or AgentArena (or anything with `.arena` being an AgentArena, e.g. the drop-in SAC)."""
    from .arena import AgentArena
    if hasattr(obj, "task_ring"):
        st = engine_state(obj)
        path = shard_path(path, obj.rank, obj.world)
    else:
        st = agent_state(obj if isinstance(obj, AgentArena) else obj.arena)
    torch.save(st, path)
    return path


def load(path, obj):
    from .arena import AgentArena
    if hasattr(obj, "task_ring"):
        path = shard_path(path, obj.rank, obj.world)
    st = torch.load(path, map_location="cpu", weights_only=False)
    if hasattr(obj, "task_ring"):
        load_engine_state(obj, st)
    else:
        load_agent_state(obj if isinstance(obj, AgentArena) else obj.arena, st)
    return st
